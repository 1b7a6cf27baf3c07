#!/bin/bash
# round-2 final bench lines at HEAD (programmatic dependent launch on, fixed per-launch byte accounting)
cd $GRAFT_REPO_ROOT
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/g2f_pytest.log 2>&1; tail -3 $O/g2f_pytest.log
(time timeout 300 python -c "import __graft_entry__ as g; g.smoke()") > $O/g2f_smoke.log 2>&1; tail -2 $O/g2f_smoke.log
timeout 900 python bench.py > $O/g2f_bench.json 2> $O/g2f_bench.err || tail -5 $O/g2f_bench.err
timeout 600 python bench.py --workload game --steps 200 --warmup 20 > $O/g2f_game.json 2> $O/g2f_game.err || tail -5 $O/g2f_game.err
timeout 600 python bench.py --workload explosion4096 --steps 10 --warmup 3 > $O/g2f_expl.json 2> $O/g2f_expl.err || tail -5 $O/g2f_expl.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/g2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong-base > $O/g2f_ncu_bench.log 2>&1
python - <<PY
import json
for n in ("bench","game","expl"):
    try:
        d=json.loads(open("$O/g2f_%s.json"%n).read().strip().splitlines()[-1])
        print(n, d.get("ms_per_step"), d.get("value"), (d.get("e2e") or {}).get("ms_per_step"), (d.get("strong_scaling_base") or {}).get("ms_per_step"), (d.get("roofline") or {}).get("frac"), (d.get("roofline_items") or {}).get("frac"))
    except Exception as e: print(n,"ERR",e)
PY
