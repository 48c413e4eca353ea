"""Debug helper: compare the CUDA path with the oracle for one small case under the current TCC_B200_* env.
usage: dbg_parity.py [kind no nv]"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import civector_numpy as O
from tencirchem_b200 import UCCSD, KUPCCGSD, PUCCD
from tencirchem_b200.hamiltonian import random_integral
from tencirchem_b200.plan import Plan
kind = sys.argv[1] if len(sys.argv) > 1 else "uccsd"
no = int(sys.argv[2]) if len(sys.argv) > 2 else 4
nv = int(sys.argv[3]) if len(sys.argv) > 3 else 4
n = no + nv
i1, i2 = random_integral(n)
u = {"uccsd": UCCSD, "kupccgsd": KUPCCGSD, "puccd": PUCCD}[kind].from_integral(i1, i2, 2 * no)
np.random.seed(2077); params = np.random.rand(u.n_params) - 0.5
c_ref = O.get_civector(params, u.n_qubits, u.n_elec, u.ex_ops, u.param_ids, u.hcb)
h = O.get_h_from_integral(i1, i2, u.n_elec, u.hcb)
e_ref, g_ref = O.get_energy_and_grad(params, h, u.n_qubits, u.n_elec, u.ex_ops, u.param_ids, u.hcb)
p = Plan(u.n_qubits, u.n_elec, u.ex_ops, u.param_ids, hcb=u.hcb)
p.set_hamiltonian(i1, i2)
c = p.civector(params)
x = np.random.default_rng(0).standard_normal(c.size)
ds = np.abs(p.apply_h(x) - h(x)).max()
e, g = p.energy_and_grad(params)
e2 = p.energy(params)
print(kind, no, nv, "TILE", os.environ.get("TCC_B200_TILE_ELEMS"), "THR", os.environ.get("TCC_B200_BATCH_THREADS"), "batches", p.n_batches,
      "dc %.2e dsigma %.2e de %.2e de2 %.2e dg %.2e" % (np.abs(c - c_ref).max(), ds, abs(e - e_ref), abs(e2 - e_ref), np.abs(g - g_ref).max()))
