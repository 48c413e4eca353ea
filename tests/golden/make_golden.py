"""Regenerates tests/golden/oracle_vectors.npz from the CPU oracle.

The Python reference cannot be imported in the build container (pyscf / tensorcircuit / openfermion
are absent), so these vectors come from ``oracle/civector_numpy.py``, whose
fidelity is pinned by tests/test_oracle_kat.py against the reference's published values
(tests/golden/published_kat.json).  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import civector_numpy as O  # noqa: E402
from oracle import pools  # noqa: E402

out = {}
cases = [("uccsd", 2, 2), ("uccsd", 3, 3), ("uccsd", 2, 3), ("kupccgsd", 2, 2), ("kupccgsd", 3, 2), ("puccd", 2, 3), ("puccd", 3, 3)]
for kind, no, nv in cases:
    n = no + nv
    int1e, int2e = O.random_integral(n)
    hcb = kind == "puccd"
    if kind == "uccsd":
        ops, pids = pools.uccsd_ex_ops(no, nv)
    elif kind == "kupccgsd":
        ops, pids = pools.kupccgsd_ex_ops(no, nv, 3)
    else:
        ops, pids = pools.puccd_ex_ops(no, nv)
    nq = n if hcb else 2 * n
    np.random.seed(2077)
    params = np.random.rand(max(pids) + 1) - 0.5
    h = O.get_h_from_integral(int1e, int2e, 2 * no, hcb)
    e, g = O.get_energy_and_grad(params, h, nq, 2 * no, ops, pids, hcb)
    c = O.get_civector(params, nq, 2 * no, ops, pids, hcb)
    tag = f"{kind}_{no}_{nv}"
    out[tag + "_params"] = params
    out[tag + "_civector"] = c
    out[tag + "_energy"] = np.array(e)
    out[tag + "_grad"] = g
    out[tag + "_sigma"] = h(c)
    O.clear_cache()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors.npz"), **out)
print("wrote", len(out), "arrays")
