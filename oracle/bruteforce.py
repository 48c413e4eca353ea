"""Independent second oracle: dense operators by brute-force second quantisation.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Shares no code with
``civector_numpy.py``: operators are applied ladder by ladder on occupation
integers with the Jordan-Wigner sign (-1)^(occupied orbitals below j), which is
the definition behind both ``ex_op_to_fop`` (utils/misc.py:121-129) and the
fermionic Hamiltonian of ``get_hop_from_integral`` (static/hamiltonian.py:51-96).
Usable up to about (6e,6o) (N = 400).
"""
from __future__ import annotations

from itertools import combinations

import numpy as np
from scipy.linalg import expm


def basis(n_qubits: int, n_elec: int, hcb: bool = False):
    """Determinants in the reference's CI order (alpha = high half, alpha-major)."""
    if hcb:
        strs = sorted(sum(1 << o for o in occ) for occ in combinations(range(n_qubits), n_elec // 2))
        return strs
    half = n_qubits // 2
    spin = sorted(sum(1 << o for o in occ) for occ in combinations(range(half), n_elec // 2))
    return [(a << half) | b for a in spin for b in spin]


def _apply_ladder(det: int, idx: int, dag: bool):
    occ = (det >> idx) & 1
    if dag == bool(occ):
        return 0, 0
    sign = -1 if bin(det & ((1 << idx) - 1)).count("1") & 1 else 1
    return sign, det ^ (1 << idx)


def apply_word(det: int, word):
    """Apply ``a^(+)_{w0} a^(+)_{w1} ...`` (rightmost operator first)."""
    sign = 1
    for idx, dag in reversed(word):
        s, det = _apply_ladder(det, idx, dag)
        if s == 0:
            return 0, 0
        sign *= s
    return sign, det


def dense_from_words(dets, words_coeffs):
    index = {d: i for i, d in enumerate(dets)}
    mat = np.zeros((len(dets), len(dets)))
    for word, coeff in words_coeffs:
        if coeff == 0.0:
            continue
        for j, d in enumerate(dets):
            s, new = apply_word(d, word)
            if s and new in index:
                mat[index[new], j] += coeff * s
    return mat


def dense_excitation_generator(ex_op, n_qubits, n_elec):
    """G = T - T^dagger for an excitation tuple (fermionic, non-hcb)."""
    half = len(ex_op) // 2
    word = [(i, True) for i in ex_op[:half]] + [(i, False) for i in ex_op[half:]]
    t = dense_from_words(basis(n_qubits, n_elec), [(word, 1.0)])
    return t - t.T


def dense_excitation_generator_hcb(ex_op, n_qubits, n_elec):
    """hcb: pair creation/annihilation without fermionic sign, b+_p b_q - h.c."""
    dets = basis(n_qubits, n_elec, hcb=True)
    index = {d: i for i, d in enumerate(dets)}
    half = len(ex_op) // 2
    t = np.zeros((len(dets), len(dets)))
    for j, d in enumerate(dets):
        if all((d >> a) & 1 for a in ex_op[half:]) and not any((d >> c) & 1 for c in ex_op[:half]):
            new = d
            for b in ex_op:
                new ^= 1 << b
            t[index[new], j] = 1.0
    return t - t.T


def dense_hamiltonian(int1e, int2e, n_elec):
    """H = sum h_pq a+_ps a_qs + 1/2 sum (pq|rs) a+_ps a+_rt a_st a_qs on the
    (n_elec/2, n_elec/2) sector (static/hamiltonian.py:51-96 semantics)."""
    n = len(int1e)
    dets = basis(2 * n, n_elec)
    words = []
    # spin-orbital index: beta p -> p, alpha p -> n + p
    for off in (0, n):
        for p in range(n):
            for q in range(n):
                words.append(([(off + p, True), (off + q, False)], float(int1e[p, q])))
    for off1 in (0, n):
        for off2 in (0, n):
            for p in range(n):
                for q in range(n):
                    for r in range(n):
                        for s in range(n):
                            v = 0.5 * float(int2e[p, q, r, s])
                            words.append(
                                ([(off1 + p, True), (off2 + r, True), (off2 + s, False), (off1 + q, False)], v)
                            )
    return dense_from_words(dets, words)


def dense_hamiltonian_hcb(int1e, int2e, n_elec):
    """Hard-core-boson Hamiltonian of static/hamiltonian.py:111-125 on pair strings."""
    n = len(int1e)
    dets = basis(n, n_elec, hcb=True)
    index = {d: i for i, d in enumerate(dets)}
    h = np.zeros((len(dets), len(dets)))
    for j, d in enumerate(dets):
        for p in range(n):
            if (d >> p) & 1:
                h[j, j] += 2 * int1e[p, p] + int2e[p, p, p, p]
            for q in range(p):
                if (d >> p) & 1 and (d >> q) & 1:
                    h[j, j] += 4 * int2e[p, p, q, q] - 2 * int2e[p, q, p, q]
                if ((d >> p) & 1) != ((d >> q) & 1):
                    h[index[d ^ (1 << p) ^ (1 << q)], j] += int2e[p, q, p, q]
    return h


def ucc_state(params, ex_ops, param_ids, n_qubits, n_elec, hcb=False, init_state=None):
    dets = basis(n_qubits, n_elec, hcb)
    c = np.zeros(len(dets))
    c[0] = 1.0
    if init_state is not None:
        c = np.array(init_state, dtype=float)
    gen = dense_excitation_generator_hcb if hcb else dense_excitation_generator
    for op, pid in zip(ex_ops, param_ids):
        c = expm(params[pid] * gen(op, n_qubits, n_elec)) @ c
    return c
