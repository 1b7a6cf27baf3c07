#!/bin/bash
# round-2 GPU call R: ncu captures of the r02 kernels (launch list, full set of one step as raw csv, source pages of the top kernels)
cd $GRAFT_REPO_ROOT
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $O/r2_pytest.log 2>&1; tail -4 $O/r2_pytest.log
UBGL_LAZY_CURRENT=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong-base > $O/r2_nolazy.json 2> $O/r2_nolazy.err || tail -5 $O/r2_nolazy.err
for v in 0 1; do
UBGL_ADVECT_FORCE_SLAB=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong-base > $O/r2_slabadv$v.json 2> $O/r2_slabadv$v.err || tail -5 $O/r2_slabadv$v.err
done
timeout 600 python bench.py --steps 20 --warmup 3 > $O/r2_bench.json 2> $O/r2_bench.err || tail -5 $O/r2_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong-base > $O/r2_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/r2_step python tools/prof_step.py --steps 1 > $O/r2_ncu_step.log 2>&1
ncu -i /tmp/r2_step.ncu-rep --page raw --csv > $O/r2_step_raw.csv 2> $O/r2_step_raw.err
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_advect_xy|k_mg_run|k_prestep_run' -c 4 -o $O/r2_src python tools/prof_step.py --steps 1 > $O/r2_ncu_src.log 2>&1
ls -la $O/r2_* /tmp/r2_step.ncu-rep
du -sh $O
python - <<PY
import json
for n in ("nolazy","slabadv0","slabadv1","bench"):
    try:
        d=json.loads(open("$O/r2_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],4), [(k["kernel"],k["level"],k["ms"]) for k in d["kernels_ms_per_step"][:3]], d.get("strong_scaling_base",{}).get("ms_per_step"))
    except Exception as e: print(n,"ERR",e)
PY
