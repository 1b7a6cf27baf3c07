#!/bin/bash
# 2-GPU call: slab equivalence tests, 2-rank bench line, compute-sanitizer summaries
cd $GRAFT_REPO_ROOT
O=gpurun_out
nvidia-smi -L
(time timeout 900 python -m pytest tests/test_slab.py -m gpu -x -q) > $O/g2_pytest_slab.log 2>&1; tail -5 $O/g2_pytest_slab.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > $O/g2_bench2.json 2> $O/g2_bench2.err; tail -c 400 $O/g2_bench2.err
python - <<PY
import json
try:
    d=json.loads(open("$O/g2_bench2.json").read().strip().splitlines()[-1])
    print("N=2", d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], "equiv", d["equiv"]["bitwise_ok"], d["run_info"]["rows_per_rank"], d["halo"])
except Exception as e: print("ERR", e)
PY
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 600 $CS --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $O/g2_san_memcheck_smoke.log 2>&1; tail -4 $O/g2_san_memcheck_smoke.log
timeout 900 $CS --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $O/g2_san_racecheck_smoke.log 2>&1; tail -4 $O/g2_san_racecheck_smoke.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 --no-python $CS --tool memcheck --print-limit 20 python tests/mgpu_equiv.py 600 512 1 0.002 > $O/g2_san_memcheck_slab2.log 2>&1; grep -E "MGPU_EQUIV|ERROR SUMMARY" $O/g2_san_memcheck_slab2.log | head
