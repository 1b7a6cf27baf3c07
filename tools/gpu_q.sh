#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
(time UBGL_MG_ROWS=2 timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_mg.py tests/test_gpu_sim.py tests/test_gpu_next.py tests/test_dropin.py -x -q) > $O/q2_pytest_rows2.log 2>&1; tail -5 $O/q2_pytest_rows2.log
run() { n=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong-base > $O/q2_$n.json 2> $O/q2_$n.err || tail -5 $O/q2_$n.err; }
run rows4 UBGL_MG_ROWS=4
run rows2 UBGL_MG_ROWS=2
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/q2_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], round(d["ms_per_step"],4), "vc", round(d["vcycle"]["ms"],4), [(k["kernel"],k["level"],k["ms"]) for k in d["kernels_ms_per_step"][:10]])
PY
grep "rigid bodies: acc" $O/parity_errors.jsonl | tail -3
