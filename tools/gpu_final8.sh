#!/bin/bash
# 8-GPU final line at HEAD (as the driver launches it)
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 20 --warmup 3 > $O/f8_bench8.json 2> $O/f8_bench8.err; echo "rc=$?"; tail -c 200 $O/f8_bench8.err
python - <<PY
import json
d=json.loads([l for l in open("$O/f8_bench8.json").read().strip().splitlines() if l.startswith("{")][-1])
print("N=8", round(d["ms_per_step"],3), round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"],1), "equiv", d["equiv"]["bitwise_ok"], "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None, d["clocks"])
PY
