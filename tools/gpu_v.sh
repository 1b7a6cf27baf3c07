#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_sim.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong-base > $O/v2_bench.json 2> $O/v2_bench.err || tail -5 $O/v2_bench.err
python - <<PY
import json
for n in ("bench",):
    d=json.loads(open("$O/v2_%s.json"%n).read().strip().splitlines()[-1])
    print(n, round(d["ms_per_step"],4), [(k["kernel"],k["level"],k["ms"]) for k in d["kernels_ms_per_step"][:6]], d["e2e"]["ms_per_step"] if d.get("e2e") else None)
PY
