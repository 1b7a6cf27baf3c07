#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_next.py tests/test_dropin.py tests/test_gpu_vs_reference.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
timeout 300 python bench.py --workload explosion4096 --steps 10 --warmup 3 --no-cpu-baseline > $O/t1_expl$i.json 2> $O/t1_expl$i.err || tail -5 $O/t1_expl$i.err
done
python - <<PY
import json
for n in ("expl1","expl2"):
    d=json.loads(open("$O/t1_%s.json"%n).read().strip().splitlines()[-1])
    print(n, round(d["ms_per_step"],4), d["particles"]["items_ms"], d["e2e"]["ms_per_step"])
PY
