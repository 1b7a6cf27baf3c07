#!/bin/bash
# final check at HEAD: whole GPU suite + smoke + the explosion frame line (disc-restricted pyramid update)
cd $GRAFT_REPO_ROOT
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/g3f_pytest.log 2>&1; tail -3 $O/g3f_pytest.log
(time timeout 300 python -c "import __graft_entry__ as g; g.smoke()") > $O/g3f_smoke.log 2>&1; tail -2 $O/g3f_smoke.log
for i in 1 2; do
timeout 600 python bench.py --workload explosion4096 --steps 10 --warmup 3 > $O/g3f_expl$i.json 2> $O/g3f_expl$i.err || tail -5 $O/g3f_expl$i.err
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-strong-base > $O/g3f_bench.json 2> $O/g3f_bench.err || tail -5 $O/g3f_bench.err
python - <<PY
import json
for n in ("expl1","expl2","bench"):
    try:
        d=json.loads(open("$O/g3f_%s.json"%n).read().strip().splitlines()[-1])
        print(n, d.get("ms_per_step"), (d.get("e2e") or {}).get("ms_per_step"), (d.get("particles") or {}).get("terrain_ms"))
    except Exception as e: print(n,"ERR",e)
PY
