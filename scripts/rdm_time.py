"""Times tcc_make_rdm12 at (n e, n o) on a random CI vector.  usage: rdm_time.py [n]"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tencirchem_b200.plan import Plan
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
p = Plan(2 * n, n, (), ())
c = torch.randn(p.size, dtype=torch.float64, device="cuda")
c /= c.norm()
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d1, d2 = p.make_rdm12_dev(c)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("rdm12 (%de,%do) N=%d: %.3f s, trace rdm1 %.12f, %.2f TFLOP/s on 2 n^4 N" % (n, n, p.size, dt, np.trace(d1), 2 * n**4 * p.size / dt / 1e12))
