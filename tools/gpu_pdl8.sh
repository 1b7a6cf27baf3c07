#!/bin/bash
# 8-GPU call: slab tests on 8 GPUs + bench line with and without programmatic dependent launch
cd $GRAFT_REPO_ROOT
O=gpurun_out
(time timeout 600 python -m pytest tests/test_slab.py -m gpu -x -q) > $O/pdl8_pytest_slab.log 2>&1; tail -3 $O/pdl8_pytest_slab.log
run() { n=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > $O/pdl8_$n.json 2> $O/pdl8_$n.err; tail -c 200 $O/pdl8_$n.err; }
PORT=29691 run on UBGL_PDL=1
PORT=29692 run off UBGL_PDL=0
python - <<PY
import json
for f in ("pdl8_on","pdl8_off"):
    try:
        d=json.loads([l for l in open("$O/"+f+".json").read().strip().splitlines() if l.startswith("{")][-1])
        print(f, round(d["ms_per_step"],3), round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"],1), "equiv", d["equiv"]["bitwise_ok"], "ex", d["run_info"]["exchanges_per_step"])
    except Exception as e: print(f,"ERR",e)
PY
