#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
run() { n=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong-base > $O/p2_$n.json 2> $O/p2_$n.err || tail -5 $O/p2_$n.err; }
run persist_noland_small UBGL_MG_DBG=12
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/p2_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], round(d["ms_per_step"],4), "vc", round(d["vcycle"]["ms"],4), [(k["kernel"],k["level"],k["ms"]) for k in d["kernels_ms_per_step"][:8]])
PY
