#!/bin/bash
# 8-GPU call: distributed-depth sweep (UBGL_SLAB_MIN_ROWS) + the bench line at the default
cd $GRAFT_REPO_ROOT
O=gpurun_out
UBGL_SLAB_SWEEP_MIN_ROWS=64,128,256,512,1024 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > $O/t2_bench8.json 2> $O/t2_bench8.err; grep "\[sweep\]" $O/t2_bench8.err; tail -c 300 $O/t2_bench8.err
python - <<PY
import json
for f in ("t2_bench8",):
    try:
        d=json.loads([l for l in open("$O/"+f+".json").read().strip().splitlines() if l.startswith("{")][-1])
        print(f, round(d["ms_per_step"],3), round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"],1), "equiv", d["equiv"]["bitwise_ok"], d["run_info"]["rows_per_rank"], "ex", d["run_info"]["exchanges_per_step"])
    except Exception as e: print(f,"ERR",e)
PY
