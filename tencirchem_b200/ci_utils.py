"""Host-side CI-string helpers with the reference's names and results
(tencirchem/static/ci_utils.py:16-87), bit-exact.  PySCF's ``cistring.make_strings`` is replaced by
the plan's own enumeration (ascending integers of fixed popcount, include/tcc_b200.h
``tcc_ci_strings_spin``); the ``[N]``-long arrays are only materialised when a caller asks for them.
"""
from __future__ import annotations

from math import comb

import numpy as np

from .plan import Plan

UINT = np.uint64


def _spin_strings(n_qubits, n_elec, hcb):
    return Plan(n_qubits, n_elec, (), (), hcb=hcb, device=None).ci_strings_spin()


def get_ci_strings(n_qubits, n_elec, hcb, strs2addr=False):
    """ci_utils.py:16-37."""
    if 2**n_qubits > np.iinfo(UINT).max:
        raise ValueError(f"Too many qubits: {n_qubits}, try using complex128 datatype")
    spin = _spin_strings(n_qubits, n_elec, hcb)
    if not hcb:
        half = n_qubits // 2
        ci_strings = ((spin << UINT(half)).reshape(-1, 1) + spin.reshape(1, -1)).ravel()
        table_bits = half
    else:
        ci_strings = spin
        table_bits = n_qubits
    if strs2addr:
        if table_bits > 32:
            raise ValueError(f"Too many qubits: {n_qubits} for a dense string lookup table")
        table = np.zeros(2**table_bits, dtype=UINT)
        table[spin] = np.arange(len(spin), dtype=UINT)
        return ci_strings, table
    return ci_strings


def get_addr(excitation, n_qubits, n_elec, strs2addr, hcb, num_strings=None):
    """ci_utils.py:40-49."""
    excitation = np.asarray(excitation, dtype=UINT)
    if hcb:
        return strs2addr[excitation]
    half = n_qubits // 2
    alpha = excitation >> UINT(half)
    beta = excitation & UINT(2**half - 1)
    if num_strings is None:
        num_strings = comb(half, n_elec // 2)
    return strs2addr[alpha] * UINT(num_strings) + strs2addr[beta]


def get_ex_bitstring(n_qubits, n_elec, ex_op, hcb):
    """ci_utils.py:52-71."""
    if not hcb:
        bits = (["0"] * (n_qubits // 2 - n_elec // 2) + ["1"] * (n_elec // 2)) * 2
    else:
        bits = ["0"] * (n_qubits - n_elec // 2) + ["1"] * (n_elec // 2)
    bits = bits[::-1]
    assert len(ex_op) in (2, 4)
    half = len(ex_op) // 2
    for i in ex_op[half:]:
        bits[i] = "0"
    for i in ex_op[:half]:
        bits[i] = "1"
    return "".join(reversed(bits))


def civector_to_statevector(civector, n_qubits, ci_strings):
    """ci_utils.py:74-76.  Raises instead of allocating 2**n_qubits doubles beyond 32 qubits."""
    if n_qubits > 32:
        raise ValueError(f"statevector of {n_qubits} qubits is too large to materialise; use the CI vector")
    sv = np.zeros(2**n_qubits, dtype=np.float64)
    sv[np.asarray(ci_strings)] = np.asarray(civector, dtype=np.float64)
    return sv


def statevector_to_civector(statevector, ci_strings):
    """ci_utils.py:79-80."""
    return np.asarray(statevector)[np.asarray(ci_strings)]


def get_init_civector(len_ci):
    """ci_utils.py:83-87 (host copy; the engine creates it on the device)."""
    c = np.zeros(len_ci, dtype=np.float64)
    c[0] = 1.0
    return c
