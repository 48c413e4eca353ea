"""UCCSD.kernel() (L-BFGS-B over energy_and_grad) on literal H_n chains, STO-3G, 0.8 A -- the reference's quickstart
table (docs/source/quickstart.rst:75-116; screened ansatz there, unscreened here).  usage: hchain_kernel.py n [n ...]"""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tencirchem_b200 import UCCSD
from tencirchem_b200.hchain import h_chain_integrals
for n in [int(a) for a in sys.argv[1:]] or [4, 6, 8]:
    t0 = time.time(); int1e, int2e, e_nuc, e_hf = h_chain_integrals(n); t_int = time.time() - t0
    ucc = UCCSD.from_integral(int1e, int2e, n, e_nuc)
    ucc.energy_and_grad(np.zeros(ucc.n_params))  # staging (plan + Hamiltonian upload)
    t0 = time.time(); e = ucc.kernel(); t_opt = time.time() - t0
    print(f"H{n}: dets {ucc.civector_size} ops {len(ucc.ex_ops)} params {ucc.n_params} e_hf {e_hf:.10f} e_uccsd {e:.10f} "
          f"nfev {ucc.opt_res['nfev']} kernel() {t_opt:.3f} s (integrals+RHF {t_int:.2f} s)")
