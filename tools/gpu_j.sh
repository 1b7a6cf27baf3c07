#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
(time timeout 1200 python -m pytest tests/test_gpu_sim.py tests/test_gpu_fused.py -x -q) > $O/j2_pytest.log 2>&1; tail -4 $O/j2_pytest.log
timeout 400 python bench.py --steps 20 --warmup 3 --no-strong-base > $O/j2_bench.json 2> $O/j2_bench.err; tail -c 300 $O/j2_bench.err
timeout 300 python bench.py --workload game --steps 200 --warmup 20 --no-cpu-baseline > $O/j2_game.json 2> $O/j2_game.err
python - <<PY
import json
for f in ("j2_bench","j2_game"):
    try:
        d=json.loads(open("$O/"+f+".json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],4), "e2e", d["e2e"]["ms_per_step"], "pipe", d.get("e2e_pipelined"), [(k["kernel"],k["level"],k["ms"]) for k in d["kernels_ms_per_step"][:6]])
    except Exception as e: print(f,"ERR",e)
PY
