#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_sim.py -m gpu -x -q 2>&1 | tail -3
