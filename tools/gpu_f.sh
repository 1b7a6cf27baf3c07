#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
run() { n=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong-base > $O/f2_$n.json 2> $O/f2_$n.err || tail -5 $O/f2_$n.err; }
run st0 X=1
run st2000 UBGL_MG_STAGGER_NS=2000
run st4000 UBGL_MG_STAGGER_NS=4000
run st6000 UBGL_MG_STAGGER_NS=6000
run st9000 UBGL_MG_STAGGER_NS=9000
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/f2_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], round(d["ms_per_step"],4), "vc", round(d["vcycle"]["ms"],4), [(k["kernel"],k["level"],k["ms"]) for k in d["kernels_ms_per_step"] if k["kernel"].startswith("mg_")][:6])
    except Exception as e: print(f,"ERR",e)
PY
