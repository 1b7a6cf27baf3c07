#!/bin/bash
# disc-restricted flag-pyramid update after craters: tests + explosion frame A/B
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_next.py tests/test_gpu_vs_reference.py tests/test_dropin.py -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do
UBGL_DISC_UPDATE=$v timeout 300 python bench.py --workload explosion4096 --steps 10 --warmup 3 --no-cpu-baseline > $O/t1_expl$v.json 2> $O/t1_expl$v.err || tail -5 $O/t1_expl$v.err
done
python - <<PY
import json
for n in ("expl0","expl1"):
    d=json.loads(open("$O/t1_%s.json"%n).read().strip().splitlines()[-1])
    print(n, round(d["ms_per_step"],4), d["particles"]["terrain_ms"], d["e2e"]["ms_per_step"])
PY
