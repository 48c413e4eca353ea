"""Times sigma = H.c at (n e, n o) through tcc_apply_h (device-resident).  usage: sigma_time.py [n]"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tencirchem_b200.hamiltonian import random_integral
from tencirchem_b200.plan import Plan
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
i1, i2 = random_integral(n)
p = Plan(2 * n, n, (), ())
p.set_hamiltonian(i1, i2)
c = torch.randn(p.size, dtype=torch.float64, device="cuda")
s = torch.empty_like(c)
p.profile_enable(True)
for _ in range(2):
    p.apply_h_dev(c, s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    p.apply_h_dev(c, s)
e1.record(); torch.cuda.synchronize()
print("sigma ms per call %.1f" % (e0.elapsed_time(e1) / 3), "checksum %.10e" % float(s.double().sum()))
