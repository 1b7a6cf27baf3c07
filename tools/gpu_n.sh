#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
(time UBGL_MG_ASYNC=2 timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_mg.py tests/test_gpu_sim.py tests/test_slab.py -x -q) > $O/n2_pytest_async2.log 2>&1; tail -5 $O/n2_pytest_async2.log
(time timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_next.py -x -q) > $O/n2_pytest.log 2>&1; tail -5 $O/n2_pytest.log
run() { n=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong-base > $O/n2_$n.json 2> $O/n2_$n.err || tail -5 $O/n2_$n.err; }
run async1 X=1
run async0 UBGL_MG_ASYNC=0
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/n2_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        ks={}
        for k in d["kernels_ms_per_step"]: ks[k["kernel"]]=ks.get(k["kernel"],0)+k["ms"]
        print(f.split("/")[-1], round(d["ms_per_step"],4), "vc", round(d["vcycle"]["ms"],4), {k:round(v,3) for k,v in ks.items()})
        print("   ", [(k["kernel"],k["level"],k["ms"]) for k in d["kernels_ms_per_step"][:10]])
    except Exception as e: print(f,"ERR",e)
PY
grep "rigid bodies" $O/parity_errors.jsonl | tail -9
