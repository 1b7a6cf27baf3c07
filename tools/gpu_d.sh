#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_mg.py tests/test_gpu_fullsize.py -x -q) > $O/d2_pytest.log 2>&1; tail -6 $O/d2_pytest.log
run() { n=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong-base > $O/d2_$n.json 2> $O/d2_$n.err || tail -5 $O/d2_$n.err; }
run pair1 X=1
run pair0 UBGL_MG_PAIR_SYNC=0
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/d2_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        ks={}
        for k in d["kernels_ms_per_step"]: ks[k["kernel"]]=ks.get(k["kernel"],0)+k["ms"]
        print(f.split("/")[-1], round(d["ms_per_step"],4), "vc", round(d["vcycle"]["ms"],4), {k:round(v,3) for k,v in ks.items()})
        print("   ", [(k["kernel"],k["level"],k["ms"]) for k in d["kernels_ms_per_step"][:10]])
    except Exception as e: print(f,"ERR",e)
PY
