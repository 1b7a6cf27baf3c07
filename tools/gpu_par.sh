#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_sim.py tests/test_gpu_advect_variants.py -m gpu -q 2>&1 | tail -3
