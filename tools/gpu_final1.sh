#!/bin/bash
# round-2 final evidence, one B200: tests + smoke, default bench line, reference arm, the other configs, launch list
cd $GRAFT_REPO_ROOT
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/f2_pytest.log 2>&1; tail -4 $O/f2_pytest.log
(time timeout 300 python -c "import __graft_entry__ as g; g.smoke()") > $O/f2_smoke.log 2>&1; tail -3 $O/f2_smoke.log
timeout 900 python bench.py > $O/f2_bench.json 2> $O/f2_bench.err || tail -5 $O/f2_bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > $O/f2_reference.json 2> $O/f2_reference.err || tail -5 $O/f2_reference.err
timeout 600 python bench.py --workload game --steps 200 --warmup 20 > $O/f2_game.json 2> $O/f2_game.err || tail -5 $O/f2_game.err
timeout 600 python bench.py --workload explosion4096 --steps 10 --warmup 3 > $O/f2_expl.json 2> $O/f2_expl.err || tail -5 $O/f2_expl.err
(ubootgl_b200/host/_build/mgtest 1025; ubootgl_b200/host/_build/mgtest 1024; python tools/mgtest_cpu.py 1025) > $O/f2_mgtest.txt 2>&1; tail -4 $O/f2_mgtest.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/f2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong-base > $O/f2_ncu_bench.log 2>&1
python - <<PY
import json
for n in ("bench","reference","game","expl"):
    try:
        d=json.loads(open("$O/f2_%s.json"%n).read().strip().splitlines()[-1])
        print(n, d.get("ms_per_step"), d.get("value"), (d.get("e2e") or {}).get("ms_per_step"), (d.get("strong_scaling_base") or {}).get("ms_per_step"), (d.get("roofline") or {}).get("frac"))
    except Exception as e: print(n,"ERR",e)
PY
