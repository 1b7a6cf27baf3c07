#!/bin/bash
# 2-GPU diagnostic: is the slab step slower with UBGL_PDL=0? (an 8-GPU run measured 16.8 vs 9.4 ms once)
cd $GRAFT_REPO_ROOT
O=gpurun_out
run() { n=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > $O/pdl2_$n.json 2> $O/pdl2_$n.err; tail -c 150 $O/pdl2_$n.err; }
PORT=29701 run off1 UBGL_PDL=0
PORT=29702 run on1 UBGL_PDL=1
PORT=29703 run off2 UBGL_PDL=0
PORT=29704 run on2 UBGL_PDL=1
python - <<PY
import json
for f in ("off1","on1","off2","on2"):
    try:
        d=json.loads([l for l in open("$O/pdl2_"+f+".json").read().strip().splitlines() if l.startswith("{")][-1])
        print(f, round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],1), "equiv", d["equiv"]["bitwise_ok"])
    except Exception as e: print(f,"ERR",e)
PY
