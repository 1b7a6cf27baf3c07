#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) > $O/h2_pytest.log 2>&1; tail -6 $O/h2_pytest.log
timeout 300 python bench.py --workload game --steps 200 --warmup 20 --no-cpu-baseline > $O/h2_game.json 2> $O/h2_game.err
timeout 400 python bench.py --workload explosion4096 --steps 10 --warmup 3 --no-cpu-baseline > $O/h2_expl.json 2> $O/h2_expl.err
UBGL_ITEMS_VARIANT=1 timeout 400 python bench.py --workload explosion4096 --steps 10 --warmup 3 --no-cpu-baseline > $O/h2_expl_v1.json 2>> $O/h2_expl.err
python - <<PY
import json
for f in ("h2_game","h2_expl","h2_expl_v1"):
    try:
        d=json.loads(open("$O/"+f+".json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],4), d.get("particles"), [(k["kernel"],k["level"],k["ms"]) for k in d["kernels_ms_per_step"][:8]])
    except Exception as e: print(f,"ERR",e)
PY
ls /usr/lib/x86_64-linux-gnu 2>/dev/null | grep -i -E "egl|osmesa|libGL\.|glvnd|gbm" | head -20
python - <<PY
import ctypes
for n in ("libEGL.so.1","libEGL.so","libOSMesa.so.8","libOSMesa.so","libGL.so.1","libGLESv2.so.2"):
    try:
        ctypes.CDLL(n); print("loadable:", n)
    except OSError as e: print("absent:", n)
PY
which eglinfo glxinfo 2>&1 | head -3
