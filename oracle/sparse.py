"""Sparse (determinant -> amplitude) restatement of the ``civector`` engine, for parity checks at sizes where the
dense oracle cannot hold one vector: (12e,12o) ... (20e,20o).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Parity: pinned against ``oracle/civector_numpy.py`` (itself
pinned by the reference's known-answer values) on every excitation kind and on sigma = H c up to (8e,8o)
(tests/test_oracle_kat.py::test_sparse_oracle_*).

A state is a pair ``(strings, amps)``: ``strings`` = sorted unique uint64 determinant bit strings in the reference's
format (alpha string << n_orb | beta string, ci_utils.py:22-25), ``amps`` = float64 amplitudes.  All arithmetic is
the reference's, restricted to the non-zeros:

* one UCC factor (evolve_civector.py:165-177):  c += (1-cos t) f2 (.) c + sin t (c[perm] (.) phase) with
  perm(I) = I xor mask (evolve_civector.py:24-29), the support classes of :32-45 and the fermionic sign of :51-88
  (``civector_numpy.get_fket_phase / get_fermion_phase`` are evaluated on the stored bit strings themselves);
* sigma = sum_{pq,rs} h2e[pq,rs] E_pq E_rs c with the absorbed integrals of hamiltonian.py:146-155 (Knowles-Handy
  three-step form, enumerated from the non-zero determinants only, collected on a given target set);
* the gradient of evolve_civector.py:221-243 in its equivalent forward form
  grad_j = <bra_j| G_j |ket_j> = <H psi| U_M ... U_j G_j |psi_{j-1}>  (bra_j = U_j^+ ... U_M^+ H psi): the sparse vector
  G_j psi_{j-1} is pushed through the remaining factors and contracted with sigma, so nothing dense is ever formed.
"""
from __future__ import annotations

import numpy as np

from . import civector_numpy as O

UINT = np.uint64


def _merge(strings, amps):
    """sorted unique strings, amplitudes of duplicates summed"""
    if len(strings) == 0:
        return strings.astype(UINT), amps.astype(np.float64)
    uniq, inv = np.unique(strings, return_inverse=True)
    out = np.zeros(len(uniq))
    np.add.at(out, inv, amps)
    return uniq, out


def _lookup(strings, amps, keys):
    """amplitude of every key (0 where the determinant is not stored)"""
    if len(strings) == 0:
        return np.zeros(len(keys))
    pos = np.searchsorted(strings, keys)
    pos[pos >= len(strings)] = len(strings) - 1
    hit = strings[pos] == keys
    return np.where(hit, amps[pos], 0.0)


def from_dense(civector, n_qubits, n_elec, hcb=False):
    ci = O.get_ci_strings(n_qubits, n_elec, hcb)
    c = np.asarray(civector, dtype=np.float64)
    nz = np.nonzero(c)[0]
    return _merge(ci[nz], c[nz])


def to_dense(state, n_qubits, n_elec, hcb=False):
    ci = O.get_ci_strings(n_qubits, n_elec, hcb)
    return _lookup(state[0], state[1], ci)


def hf_state(n_qubits, n_elec, hcb=False):
    """ci_utils.py:83-87: unit vector on address 0 = the lowest string of both spins"""
    k = n_elec // 2
    low = (1 << k) - 1
    s = low if hcb else (low << (n_qubits // 2)) | low
    return np.array([s], dtype=UINT), np.array([1.0])


def _mask(f_idx):
    m = 0
    for i in f_idx:
        m |= 1 << i
    return UINT(m)


def _phases(f_idx, n_qubits, strings, hcb):
    """(fket_phase, f2ket_phase) of evolve_civector.py:91-106 on the given bit strings"""
    if len(set(f_idx)) != len(f_idx):
        raise ValueError(f"Excitation {f_idx} not supported")
    positive, negative = O.get_fket_phase(tuple(f_idx), strings)
    phase = np.zeros(len(strings))
    phase -= positive
    phase += negative
    if not hcb:
        phase *= O.get_fermion_phase(tuple(f_idx), n_qubits, strings.copy())
    f2 = np.zeros(len(strings))
    f2 -= positive
    f2 -= negative
    return phase, f2


def apply_excitation(state, n_qubits, f_idx, hcb=False):
    """G|c> (evolve_civector.py:311-313): out_I = c[perm(I)] * phase_I"""
    strings, amps = state
    phase, _ = _phases(f_idx, n_qubits, strings, hcb)
    sup = phase != 0
    # the partner J = I xor mask of a support determinant I receives phase_J * c_I
    partners = strings[sup] ^ _mask(f_idx)
    ph_partner, _ = _phases(f_idx, n_qubits, partners, hcb)
    return _merge(partners, ph_partner * amps[sup])


def evolve_excitation(state, n_qubits, f_idx, theta, hcb=False):
    """one factor exp(theta G) on the non-zeros (evolve_civector.py:169-175)"""
    strings, amps = state
    if theta == 0.0:
        return strings, amps
    phase, _ = _phases(f_idx, n_qubits, strings, hcb)
    partners = strings[phase != 0] ^ _mask(f_idx)
    allstr = np.union1d(strings, partners)
    c = _lookup(strings, amps, allstr)
    phase, f2 = _phases(f_idx, n_qubits, allstr, hcb)
    fket = _lookup(strings, amps, allstr ^ _mask(f_idx)) * phase
    out = c + (1.0 - np.cos(theta)) * (f2 * c) + np.sin(theta) * fket
    return allstr, out


def get_civector(params, n_qubits, n_elec, ex_ops, param_ids, hcb=False, init_state=None):
    """evolve_civector.py:180-198 on a sparse state"""
    state = hf_state(n_qubits, n_elec, hcb) if init_state is None else (np.asarray(init_state[0], dtype=UINT), np.asarray(init_state[1], dtype=np.float64))
    for f_idx, pid in zip(ex_ops, param_ids):
        state = evolve_excitation(state, n_qubits, tuple(f_idx), float(params[pid]), hcb)
    return state


# ---- sigma = H c on a target set -------------------------------------------------------------------------------
def _single_replacements(strings, n_orb, shift):
    """all E_pq (acting on the spin whose bits start at `shift`) of every string:
    yields (pq, index of the source string, new string, sign) arrays per (p, q)"""
    sp = (strings >> UINT(shift)) & UINT((1 << n_orb) - 1)
    for p in range(n_orb):
        for q in range(n_orb):
            occ_q = (sp >> UINT(q)) & UINT(1) == 1
            if p == q:
                src = np.nonzero(occ_q)[0]
                yield p * n_orb + q, src, strings[src], np.ones(len(src))
                continue
            valid = occ_q & ((sp >> UINT(p)) & UINT(1) == 0)
            src = np.nonzero(valid)[0]
            if len(src) == 0:
                continue
            new = strings[src] ^ UINT(((1 << p) | (1 << q)) << shift)
            lo, hi = min(p, q), max(p, q)
            between = UINT(((1 << hi) - 1) & ~((1 << (lo + 1)) - 1))
            par = sp[src] & between
            par ^= par >> UINT(16)
            par ^= par >> UINT(8)
            par ^= par >> UINT(4)
            par ^= par >> UINT(2)
            par ^= par >> UINT(1)
            yield p * n_orb + q, src, new, 1.0 - 2.0 * (par & UINT(1)).astype(np.float64)


def apply_h(state, int1e, int2e, n_elec, targets, chunk=1 << 17):
    """sigma_I for every I in `targets` (sorted unique uint64) of sigma = sum h2e[pq,rs] E_pq E_rs c
    (hamiltonian.py:146-155; pyscf direct_nosym.absorb_h1e / contract_2e restated in civector_numpy).  E_pq is the
    spin-summed single replacement: the alpha part acts on the high half of the bit string, the beta part on the low
    half, with no cross-spin sign (civector_numpy.contract_2e)."""
    strings, amps = state
    n_orb = len(int1e)
    h2e = O.absorb_h1e(int1e, int2e, n_orb, n_elec, 0.5).reshape(n_orb * n_orb, n_orb * n_orb)
    targets = np.asarray(targets, dtype=UINT)
    sigma = np.zeros(len(targets))
    # step 1: D[rs, K] = sum_J <K|E_rs|J> c_J as (K, rs, value) triples
    ks, rss, vals = [], [], []
    for shift in (n_orb, 0):
        for rs, src, new, sgn in _single_replacements(strings, n_orb, shift):
            ks.append(new)
            rss.append(np.full(len(new), rs, dtype=np.int64))
            vals.append(sgn * amps[src])
    if not ks:
        return sigma
    ks, rss, vals = np.concatenate(ks), np.concatenate(rss), np.concatenate(vals)
    uniq_k, inv = np.unique(ks, return_inverse=True)
    order = np.argsort(inv, kind="stable")
    inv, rss, vals = inv[order], rss[order], vals[order]
    bounds = np.searchsorted(inv, np.arange(0, len(uniq_k) + chunk, chunk))
    for ci in range(len(bounds) - 1):
        k_lo = ci * chunk
        k_str = uniq_k[k_lo:k_lo + chunk]
        if len(k_str) == 0:
            break
        sl = slice(bounds[ci], bounds[ci + 1])
        d = np.zeros((n_orb * n_orb, len(k_str)))
        np.add.at(d, (rss[sl], inv[sl] - k_lo), vals[sl])
        g = h2e @ d  # step 2: G[pq, K]
        # step 3: sigma_I += <I|E_pq|K> G[pq, K], kept where I is a target
        for shift in (n_orb, 0):
            for pq, src, new, sgn in _single_replacements(k_str, n_orb, shift):
                pos = np.searchsorted(targets, new)
                pos[pos >= len(targets)] = len(targets) - 1
                hit = targets[pos] == new
                if hit.any():
                    np.add.at(sigma, pos[hit], sgn[hit] * g[pq, src[hit]])
    return sigma


def dot(a, b):
    """<a|b> of two sparse states"""
    return float(np.dot(a[1], _lookup(b[0], b[1], a[0])))


def get_energy_and_grad(params, int1e, int2e, n_qubits, n_elec, ex_ops, param_ids, init_state=None, grad_params=None):
    """evolve_civector.py:201-218 on sparse states.  Returns (energy, {param id: gradient}) for the parameter ids in
    `grad_params` (default: all).  Factors with theta == 0 are the identity and are skipped in every product; their
    gradients are not zero and are computed like the others."""
    params = np.asarray(params, dtype=np.float64)
    n_ops = len(ex_ops)
    if grad_params is None:
        grad_params = sorted(set(param_ids))
    grad_params = sorted(set(int(p) for p in grad_params))
    want = set(grad_params)
    state = hf_state(n_qubits, n_elec) if init_state is None else (np.asarray(init_state[0], dtype=UINT), np.asarray(init_state[1], dtype=np.float64))
    # forward sweep; remember G_j psi_{j-1} of every wanted factor (G_j commutes with its own rotation)
    pending = []  # (j, sparse vector G_j psi_{j-1})
    for j in range(n_ops):
        f_idx = tuple(ex_ops[j])
        if param_ids[j] in want:
            pending.append((j, apply_excitation(state, n_qubits, f_idx)))
        state = evolve_excitation(state, n_qubits, f_idx, float(params[param_ids[j]]))
    psi = state
    # push every pending vector through the factors j .. M-1 (only the non-identity ones do anything)
    nz_ops = [j for j in range(n_ops) if params[param_ids[j]] != 0.0]
    pushed = []
    for j, v in pending:
        for i in nz_ops:
            if i >= j:
                v = evolve_excitation(v, n_qubits, tuple(ex_ops[i]), float(params[param_ids[i]]))
        pushed.append((j, v))
    targets = psi[0]
    for _, v in pushed:
        targets = np.union1d(targets, v[0])
    sigma = (targets, apply_h(psi, int1e, int2e, n_elec, targets))
    energy = dot(psi, sigma)
    grads = {p: 0.0 for p in grad_params}
    for j, v in pushed:
        grads[param_ids[j]] += 2.0 * dot(v, sigma)
    return energy, grads, psi, sigma
