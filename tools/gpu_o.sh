#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
(time UBGL_MG_ASYNC=2 timeout 900 python -m pytest tests/test_gpu_fused.py -x -q) > $O/o2_pytest_async2.log 2>&1; tail -3 $O/o2_pytest_async2.log
run() { n=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong-base > $O/o2_$n.json 2> $O/o2_$n.err || tail -5 $O/o2_$n.err; }
run async1 X=1
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/o2_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], round(d["ms_per_step"],4), "vc", round(d["vcycle"]["ms"],4), [(k["kernel"],k["level"],k["ms"]) for k in d["kernels_ms_per_step"][:8]])
PY
