#!/bin/bash
# 1-GPU call: divergence in the advect epilogue -- full GPU suite, A/B bench
cd $GRAFT_REPO_ROOT
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $O/y2_pytest.log 2>&1; tail -4 $O/y2_pytest.log
for v in 1 0; do
UBGL_ADVECT_DIV=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong-base > $O/y2_div$v.json 2> $O/y2_div$v.err || tail -5 $O/y2_div$v.err
done
UBGL_ADVECT_DIV=1 timeout 600 python bench.py --workload game --steps 200 --warmup 20 --no-cpu-baseline > $O/y2_game1.json 2> $O/y2_game1.err || tail -5 $O/y2_game1.err
UBGL_ADVECT_DIV=0 timeout 600 python bench.py --workload game --steps 200 --warmup 20 --no-cpu-baseline > $O/y2_game0.json 2> $O/y2_game0.err || tail -5 $O/y2_game0.err
python - <<PY
import json
for n in ("div1","div0","game1","game0"):
    try:
        d=json.loads(open("$O/y2_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],4), [(k["kernel"],k["level"],k["launches"],k["ms"]) for k in d["kernels_ms_per_step"] if k["kernel"] in ("advect","divergence","prestep_fused","finish_fused")])
    except Exception as e: print(n,"ERR",e)
PY
