#!/bin/bash
# 8-GPU diagnostic call: box topology, per-rank launch timeline of the slab step, e2e host-thread sweep
cd $GRAFT_REPO_ROOT
O=gpurun_out
(nvidia-smi topo -m; echo; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)|thread"; echo; nproc; taskset -p $$; for n in /sys/devices/system/node/node*; do echo $n $(grep MemTotal $n/meminfo); done; cat /proc/meminfo | head -3) > $O/s2_topo.txt 2>&1
UBGL_TIMELINE=$O/s2_tl UBGL_SLAB_E2E_SWEEP=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > $O/s2_bench8.json 2> $O/s2_bench8.err; tail -c 300 $O/s2_bench8.err
grep "e2e sweep" $O/s2_bench8.json $O/s2_bench8.err
python - <<PY
import json
for f in ("s2_bench8",):
    try:
        d=json.loads(open("$O/"+f+".json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"],1), "equiv", d["equiv"]["bitwise_ok"], d["run_info"]["rows_per_rank"], "ex", d["run_info"]["exchanges_per_step"])
    except Exception as e: print(f,"ERR",e)
PY
ls $O/s2_tl* | head
