#!/bin/bash
# 8-GPU call: N=8 bench lines (weighted and equal cuts)
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 --steps 10 --warmup 3 > $O/m2_bench8.json 2> $O/m2_bench8.err; tail -c 300 $O/m2_bench8.err
UBGL_SLAB_BALANCE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > $O/m2_bench8_nobal.json 2> $O/m2_bench8_nobal.err
python - <<PY
import json
for f in ("m2_bench8","m2_bench8_nobal"):
    try:
        d=json.loads(open("$O/"+f+".json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"],1), "equiv", d["equiv"]["bitwise_ok"], d["run_info"]["rows_per_rank"], "ex", d["run_info"]["exchanges_per_step"])
        print("   halo", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["halo"].items() if k!="note"})
        print("   ", [(k["kernel"],k["level"],k["ms"]) for k in d["kernels_ms_per_step_rank0"][:8]])
    except Exception as e: print(f,"ERR",e)
PY
