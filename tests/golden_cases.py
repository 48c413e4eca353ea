"""Loader of the sparse-oracle known answers (tests/golden/sparse_parity_*.json, made by make_sparse_golden.py).
Shared by the GPU parity tests and by bench.py's `--parity-golden` check; reads only committed fixtures, never the
oracle or /root/reference."""
from __future__ import annotations

import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def spin_ref_address(strings: np.ndarray, n_orb: int) -> np.ndarray:
    """Reference address (rank among the ascending bit strings with the same popcount, ci_utils.py:22-28) of each
    per-spin bit string, by the combinatorial number system."""
    from math import comb

    out = np.zeros(len(strings), dtype=np.int64)
    for i, s in enumerate(strings):
        s = int(s)
        r, k, o = 0, 1, 0
        while s:
            if s & 1:
                r += comb(o, k)
                k += 1
            s >>= 1
            o += 1
        out[i] = r
    return out


class SparseCase:
    def __init__(self, n_orb: int):
        path = os.path.join(HERE, "golden", f"sparse_parity_{n_orb}o.json")
        d = json.load(open(path))
        self.d = d
        self.n_orb = n_orb
        self.n_qubits, self.n_elec, self.n_params = d["n_qubits"], d["n_elec"], d["n_params"]
        self.params = np.zeros(self.n_params)
        self.params[d["param_ids_nonzero"]] = d["param_values"]
        self.energy = d["energy"]
        self.grad_ids = np.asarray(d["grad_ids"])
        self.grad_values = np.asarray(d["grad_values"])
        self.norm2 = d["norm2"]
        mask = (1 << n_orb) - 1
        s = np.asarray(d["init_strings"], dtype=np.uint64)
        self.init_alpha = spin_ref_address((s >> np.uint64(n_orb)).astype(np.int64), n_orb)
        self.init_beta = spin_ref_address((s & np.uint64(mask)).astype(np.int64), n_orb)
        self.init_amps = np.asarray(d["init_amps"])
        self.nstr = None

    def init_sparse(self):
        return self.init_alpha, self.init_beta, self.init_amps

    def flat_address(self, strings, nstr):
        s = np.asarray(strings, dtype=np.uint64)
        mask = (1 << self.n_orb) - 1
        a = spin_ref_address((s >> np.uint64(self.n_orb)).astype(np.int64), self.n_orb)
        b = spin_ref_address((s & np.uint64(mask)).astype(np.int64), self.n_orb)
        return a * nstr + b

    def check(self, energy, grad, e_tol=1e-10, g_tol=1e-8):
        """-> dict of absolute errors; raises AssertionError beyond the north_star tolerances"""
        e_err = abs(energy - self.energy)
        g_err = float(np.abs(np.asarray(grad)[self.grad_ids] - self.grad_values).max())
        res = {"energy_abs_err": e_err, "grad_max_abs_err": g_err, "n_grad_checked": int(len(self.grad_ids)),
               "e_tol": e_tol, "g_tol": g_tol, "ok": bool(e_err < e_tol and g_err < g_tol)}
        return res
