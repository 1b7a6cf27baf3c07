#!/bin/bash
# 4-GPU call: N=4 bench line for the scaling table
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29681 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline > $O/z2_bench4.json 2> $O/z2_bench4.err; tail -c 200 $O/z2_bench4.err
python - <<PY
import json
d=json.loads([l for l in open("$O/z2_bench4.json").read().strip().splitlines() if l.startswith("{")][-1])
print("N=4", round(d["ms_per_step"],3), round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"],1), "equiv", d["equiv"]["bitwise_ok"], d["run_info"]["rows_per_rank"], "ex", d["run_info"]["exchanges_per_step"], d["run_info"]["decomposition"])
PY
