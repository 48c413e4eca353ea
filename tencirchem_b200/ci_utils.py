"""Host-side CI-string helpers: same names, arguments and results as tencirchem/static/ci_utils.py:16-87
(bit-exact), built on the engine's own string enumeration instead of PySCF's ``cistring`` module.

The ``[N]``-long arrays are only materialised when a caller asks for them: the engine itself never needs them
(everything factorises per spin string).
"""
from __future__ import annotations

from math import comb

import numpy as np

from .plan import Plan

UINT = np.uint64
_MAX_TABLE_BITS = 32


def _reference_order_strings(n_qubits: int, n_elec: int, hcb: bool) -> np.ndarray:
    """Occupation strings of one spin (all orbitals for hcb) in the reference's address order, from
    ``tcc_ci_strings_spin`` (ascending integers with n_elec/2 bits set)."""
    host_plan = Plan(n_qubits, n_elec, (), (), hcb=hcb, device=None)
    return host_plan.ci_strings_spin()


def get_ci_strings(n_qubits, n_elec, hcb, strs2addr=False):
    """ci_utils.py:16-37: determinant bit strings (alpha string in the high half, alpha-major) and, on request,
    the string -> address lookup table (zero for strings outside the space)."""
    if (1 << n_qubits) > np.iinfo(UINT).max:
        raise ValueError(f"Too many qubits: {n_qubits}, try using complex128 datatype")
    one_spin = _reference_order_strings(n_qubits, n_elec, hcb)
    if hcb:
        bits_in_table, determinants = n_qubits, one_spin
    else:
        bits_in_table = n_qubits // 2
        high = one_spin << UINT(bits_in_table)
        determinants = np.add.outer(high, one_spin).reshape(-1)
    if not strs2addr:
        return determinants
    if bits_in_table > _MAX_TABLE_BITS:
        raise ValueError(f"Too many qubits: {n_qubits} for a dense string lookup table")
    lookup = np.zeros(1 << bits_in_table, dtype=UINT)
    lookup[one_spin] = np.arange(one_spin.size, dtype=UINT)
    return determinants, lookup


def get_addr(excitation, n_qubits, n_elec, strs2addr, hcb, num_strings=None):
    """ci_utils.py:40-49: address of each bit string; strings outside the space map to 0, like upstream."""
    strings = np.asarray(excitation, dtype=UINT)
    if hcb:
        return strs2addr[strings]
    half = n_qubits // 2
    n_per_spin = comb(half, n_elec // 2) if num_strings is None else num_strings
    row = strs2addr[strings >> UINT(half)]
    col = strs2addr[strings & UINT((1 << half) - 1)]
    return row * UINT(n_per_spin) + col


def get_ex_bitstring(n_qubits, n_elec, ex_op, hcb):
    """ci_utils.py:52-71: the HF determinant with the excitation applied, as a bit string (qubit 0 rightmost)."""
    if len(ex_op) not in (2, 4):
        raise AssertionError("excitation tuples have 2 or 4 indices")
    n_occ = n_elec // 2
    hf = (1 << n_occ) - 1
    det = hf if hcb else hf | (hf << (n_qubits // 2))
    n_create = len(ex_op) // 2
    for q in ex_op[n_create:]:
        det &= ~(1 << q)
    for q in ex_op[:n_create]:
        det |= 1 << q
    return format(det, f"0{n_qubits}b")


def civector_to_statevector(civector, n_qubits, ci_strings):
    """ci_utils.py:74-76.  Raises instead of allocating 2**n_qubits doubles beyond 32 qubits."""
    if n_qubits > _MAX_TABLE_BITS:
        raise ValueError(f"statevector of {n_qubits} qubits is too large to materialise; use the CI vector")
    full = np.zeros(1 << n_qubits, dtype=np.float64)
    full[np.asarray(ci_strings)] = np.asarray(civector, dtype=np.float64)
    return full


def statevector_to_civector(statevector, ci_strings):
    """ci_utils.py:79-80."""
    return np.asarray(statevector)[np.asarray(ci_strings)]


def get_init_civector(len_ci):
    """ci_utils.py:83-87 (host copy; the engine creates the HF vector on the device)."""
    hf = np.zeros(len_ci, dtype=np.float64)
    hf[0] = 1.0
    return hf
