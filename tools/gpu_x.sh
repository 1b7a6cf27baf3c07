#!/bin/bash
# 2-GPU call: slab tests incl. N-GPU == 1-GPU, N=2 bench line (process-group shutdown at exit)
cd $GRAFT_REPO_ROOT
O=gpurun_out
(time timeout 900 python -m pytest tests/test_slab.py -m gpu -x -q) > $O/x2_pytest_slab.log 2>&1; tail -3 $O/x2_pytest_slab.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 > $O/x2_bench2.json 2> $O/x2_bench2.err; echo "rc=$?"; tail -c 300 $O/x2_bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/x2_ref2.json 2> $O/x2_ref2.err; echo "rc=$?"; tail -c 200 $O/x2_ref2.json
python - <<PY
import json
d=json.loads([l for l in open("$O/x2_bench2.json").read().strip().splitlines() if l.startswith("{")][-1])
print("N=2", round(d["ms_per_step"],3), round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"],1), "equiv", d["equiv"]["bitwise_ok"], "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
PY
