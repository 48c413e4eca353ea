"""Two UCCSD energy_and_grad evaluations at (n e, n o) for profiler captures.  usage: prof_eval.py [n_orbitals]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tencirchem_b200 import UCCSD
from tencirchem_b200.hamiltonian import random_integral
from tencirchem_b200.plan import Plan

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
i1, i2 = random_integral(n)
u = UCCSD.from_integral(i1, i2, n)
p = Plan(2 * n, n, u.ex_ops, u.param_ids)
p.set_hamiltonian(i1, i2)
np.random.seed(1)
params = np.random.rand(u.n_params) - 0.5
for _ in range(2):
    e, g = p.energy_and_grad_dev(params)
print("E =", e, "|g| =", float(np.linalg.norm(g)))
