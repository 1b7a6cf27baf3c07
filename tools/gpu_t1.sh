#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
(time timeout 1500 python -m pytest tests/test_gpu_fused.py tests/test_gpu_mg.py tests/test_gpu_sim.py tests/test_gpu_fullsize.py tests/test_host_mirror.py -m gpu -x -q) > $O/t1_pytest.log 2>&1; tail -4 $O/t1_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong-base > $O/t1_bench.json 2> $O/t1_bench.err || tail -5 $O/t1_bench.err
timeout 300 python bench.py --workload game --steps 200 --warmup 20 --no-cpu-baseline > $O/t1_game.json 2> $O/t1_game.err || tail -5 $O/t1_game.err
(ubootgl_b200/host/_build/mgtest 1025 --resident; ubootgl_b200/host/_build/mgtest 1024 --resident) 2>&1 | tail -4
python - <<PY
import json
for n in ("bench","game"):
    d=json.loads(open("$O/t1_%s.json"%n).read().strip().splitlines()[-1])
    print(n, round(d["ms_per_step"],4), "vcycle", d["vcycle"]["ms"], [(k["kernel"],k["level"],k["ms"]) for k in d["kernels_ms_per_step"] if "coarse" in k["kernel"]])
PY
