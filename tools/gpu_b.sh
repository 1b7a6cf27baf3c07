#!/bin/bash
# round-2 GPU call B: tests, advect / prestep A/B runs, ncu captures
cd $GRAFT_REPO_ROOT
O=gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) > $O/b2_pytest.log 2>&1; tail -8 $O/b2_pytest.log
run() { # name, env...
  n=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong-base > $O/b2_$n.json 2> $O/b2_$n.err || tail -5 $O/b2_$n.err
}
run default X=1
run occ3 UBGL_ADVECT_OCC=3
run occ5 UBGL_ADVECT_OCC=5
run occ6 UBGL_ADVECT_OCC=6
run pre0 UBGL_PRESTEP_VARIANT=0
run tail16k UBGL_MG_TAIL_CELLS=16384
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/b2_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        ks={}
        for k in d["kernels_ms_per_step"]: ks[k["kernel"]]=ks.get(k["kernel"],0)+k["ms"]
        print(f.split("/")[-1], round(d["ms_per_step"],4), "vc", round(d["vcycle"]["ms"],4), {k:round(v,3) for k,v in ks.items()})
    except Exception as e: print(f,"ERR",e)
PY
# ncu: launch list of 2 steps after warm-up, and full captures of the new kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/b2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong-base > $O/b2_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_advect_xy|k_prestep_run' -s 4 -c 3 -o $O/b2_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong-base > $O/b2_ncu_full.log 2>&1
ls -la $O/b2_prof.ncu-rep
