#!/bin/bash
# 8-GPU call: bench line at the new defaults (distributed depth 256 rows, main-thread clock samples), and with weighted cuts
cd $GRAFT_REPO_ROOT
O=gpurun_out
run() { n=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > $O/w2_$n.json 2> $O/w2_$n.err; tail -c 200 $O/w2_$n.err; }
PORT=29671 run default UBGL_TIMELINE=$O/w2_tl
PORT=29672 run balanced UBGL_SLAB_BALANCE=1
python - <<PY
import json
for f in ("w2_default","w2_balanced"):
    try:
        d=json.loads([l for l in open("$O/"+f+".json").read().strip().splitlines() if l.startswith("{")][-1])
        print(f, round(d["ms_per_step"],3), round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"],1), "equiv", d["equiv"]["bitwise_ok"], d["run_info"]["rows_per_rank"], "ex", d["run_info"]["exchanges_per_step"], d["clocks"])
        print("   halo", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["halo"].items() if k!="note"})
    except Exception as e: print(f,"ERR",e)
PY
