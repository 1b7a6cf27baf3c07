#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
run() { n=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong-base > $O/e2_$n.json 2> $O/e2_$n.err || tail -5 $O/e2_$n.err; }
run dbg0 X=1
run dbg1 UBGL_MG_DBG=1
run dbg2 UBGL_MG_DBG=2
run dbg3 UBGL_MG_DBG=3
run pf0 UBGL_MG_PREFETCH=0
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/e2_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], round(d["ms_per_step"],4), "vc", round(d["vcycle"]["ms"],4), [(k["kernel"],k["level"],k["ms"]) for k in d["kernels_ms_per_step"] if k["kernel"].startswith("mg_")][:6])
    except Exception as e: print(f,"ERR",e)
PY
