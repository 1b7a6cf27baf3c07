#!/bin/bash
# programmatic dependent launch on every kernel of the step chain: tests + A/B
cd $GRAFT_REPO_ROOT
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $O/pdl_pytest.log 2>&1; tail -3 $O/pdl_pytest.log
for v in 0 1; do
UBGL_PDL=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong-base > $O/pdl_bench$v.json 2> $O/pdl_bench$v.err || tail -5 $O/pdl_bench$v.err
UBGL_PDL=$v timeout 300 python bench.py --workload game --steps 200 --warmup 20 --no-cpu-baseline > $O/pdl_game$v.json 2> $O/pdl_game$v.err || tail -5 $O/pdl_game$v.err
UBGL_PDL=$v timeout 300 python bench.py --workload explosion4096 --steps 10 --warmup 3 --no-cpu-baseline > $O/pdl_expl$v.json 2> $O/pdl_expl$v.err || tail -5 $O/pdl_expl$v.err
done
python - <<PY
import json
for n in ("bench0","bench1","game0","game1","expl0","expl1"):
    try:
        d=json.loads(open("$O/pdl_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],4), "vcycle", d["vcycle"]["ms"] if d.get("vcycle") else None, "e2e", d["e2e"]["ms_per_step"], d.get("roofline_items",{}).get("frac"))
    except Exception as e: print(n,"ERR",e)
PY
