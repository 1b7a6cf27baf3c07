#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
UBGL_MG_PDL=1 timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_mg.py tests/test_gpu_sim.py -m gpu -x -q 2>&1 | tail -2
for v in 0 1; do
UBGL_MG_PDL=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong-base > $O/pdl_bench$v.json 2> $O/pdl_bench$v.err || tail -5 $O/pdl_bench$v.err
UBGL_MG_PDL=$v timeout 300 python bench.py --workload game --steps 200 --warmup 20 --no-cpu-baseline > $O/pdl_game$v.json 2> $O/pdl_game$v.err || tail -5 $O/pdl_game$v.err
done
python - <<PY
import json
for n in ("bench0","bench1","game0","game1"):
    try:
        d=json.loads(open("$O/pdl_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],4), "vcycle", d["vcycle"]["ms"] if d.get("vcycle") else None)
    except Exception as e: print(n,"ERR",e)
PY
