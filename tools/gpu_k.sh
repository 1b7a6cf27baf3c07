#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 300 python tools/pcie_probe.py > $O/k2_pcie.txt 2>&1; cat $O/k2_pcie.txt
UBGL_PIPE_DEBUG=1 timeout 300 python tools/pipe_probe.py > $O/k2_pipe.txt 2>&1; grep -v "^\[pipe\]" $O/k2_pipe.txt; grep "^\[pipe\]" $O/k2_pipe.txt | sed -n '4,8p;12,14p;20,22p;28,30p'
