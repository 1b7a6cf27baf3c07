"""Times the UNMODIFIED reference's MG::solve (oracle/_ref) on the mgtest problem
(mgtest.cpp:10-61) on this box's host cores, for the line next to host/mgtest's GPU numbers.
    python tools/mgtest_cpu.py [N=1025]
"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bind as ob

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1025
chk = ob.Ref() if ob.have_ref() else ob.Port()
kind = "reference (oracle/_ref)" if ob.have_ref() else "C restatement"
nproc = chk.num_procs()
chk.set_threads(nproc)
h = 1.0 / (N - 1)
u = np.zeros((N, N), np.float32)
x = np.arange(1, N - 1)
u[N - 1, 1:N - 1] = np.sinh(np.pi) * np.sin(x / (N - 1.0) * np.pi)
rhs = np.zeros((N, N), np.float32)
flag = np.ones((N, N), np.float32)
m = chk.MG(N, N)
m.update_fields(flag)
m.set(u, rhs, flag)
hist = []
for i in range(5):
    m.solve(h, False)
    hist.append(m.residual(h))
t0 = time.perf_counter()
for i in range(10):
    m.solve(h, False)
t = (time.perf_counter() - t0) / 10
path = "pipelined path 3" if N // nproc >= 100 else "canonical red-black"
print(f"{kind}, N={N}, {nproc} OMP threads ({path}): residual after cycles 1-5 {['%.6g' % r for r in hist]}, "
      f"{t * 1e3:.2f} ms per MG::solve")
