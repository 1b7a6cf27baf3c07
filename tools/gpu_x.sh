#!/bin/bash
# 2-GPU call (round-2 final state): slab tests incl. N-GPU == 1-GPU, memcheck of the rewritten halo kernel, N=2 bench line
cd $GRAFT_REPO_ROOT
O=gpurun_out
(time timeout 900 python -m pytest tests/test_slab.py -m gpu -x -q) > $O/x2_pytest_slab.log 2>&1; tail -3 $O/x2_pytest_slab.log
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 --no-python $CS --tool memcheck --print-limit 20 python tests/mgpu_equiv.py 600 512 1 0.002 > $O/x2_san_memcheck_slab2.log 2>&1; grep -E "MGPU_EQUIV|ERROR SUMMARY" $O/x2_san_memcheck_slab2.log | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $O/x2_bench2.json 2> $O/x2_bench2.err; tail -c 200 $O/x2_bench2.err
python - <<PY
import json
d=json.loads([l for l in open("$O/x2_bench2.json").read().strip().splitlines() if l.startswith("{")][-1])
print("N=2", round(d["ms_per_step"],3), round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"],1), "equiv", d["equiv"]["bitwise_ok"], d["run_info"]["rows_per_rank"], "ex", d["run_info"]["exchanges_per_step"], d["run_info"]["decomposition"])
PY
