#!/bin/bash
# 1-GPU call: full GPU test suite + bench line (prestep loads one iteration ahead, lazy *_current, display export)
cd $GRAFT_REPO_ROOT
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $O/u2_pytest.log 2>&1; tail -4 $O/u2_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong-base > $O/u2_bench.json 2> $O/u2_bench.err || tail -5 $O/u2_bench.err
timeout 600 python bench.py --workload game --steps 200 --warmup 20 --no-cpu-baseline > $O/u2_game.json 2> $O/u2_game.err || tail -5 $O/u2_game.err
python - <<PY
import json
for n in ("bench","game"):
    try:
        d=json.loads(open("$O/u2_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],4), [(k["kernel"],k["level"],k["ms"]) for k in d["kernels_ms_per_step"][:6]], d["e2e"]["ms_per_step"] if d.get("e2e") else None)
    except Exception as e: print(n,"ERR",e)
PY
