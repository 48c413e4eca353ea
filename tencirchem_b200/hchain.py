"""PySCF-free inputs for the hydrogen-chain configurations of BASELINE.json (H4 / H8 / H12, STO-3G).

The reference builds its molecules, RHF orbitals and MO integrals with PySCF
(tencirchem/molecule.py:72-99 ``h_chain``; tencirchem/static/hamiltonian.py:30-48 ``get_integral_from_hf``;
utils/misc.py:103-108 ``canonical_mo_coeff``).  PySCF is not available here, and for s-type Gaussians every
integral has a closed form (Boys function F0), so this module restates that input path for hydrogen chains only:
STO-3G contracted s functions, overlap / kinetic / nuclear-attraction / electron-repulsion integrals, restricted
Hartree-Fock, transformation to the MO basis.  It is pinned by the MO integrals the reference publishes for H2
(docs/source/tutorial_jupyter/ucc_functions.ipynb, tests/golden/published_kat.json).

This is host-side setup (the reference runs it once in the constructor, ucc.py:241-300), not part of the hot path.
"""
from __future__ import annotations

from math import erf, pi, sqrt

import numpy as np

BOHR = 0.52917721092  # Angstrom; the value PySCF uses (reproduces the published e_nuc of H2 to the last digit)
STO3G_H_EXP = (3.42525091, 0.62391373, 0.16885540)
STO3G_H_COEF = (0.15432897, 0.53532814, 0.44463454)


def _boys0(t: float) -> float:
    if t < 1e-12:
        return 1.0 - t / 3.0
    return 0.5 * sqrt(pi / t) * erf(sqrt(t))


def _primitives(centers):
    """Flattened primitive list: (center index, exponent, contraction coefficient x primitive norm)."""
    prims = []
    for ic in range(len(centers)):
        for a, d in zip(STO3G_H_EXP, STO3G_H_COEF):
            prims.append((ic, a, d * (2.0 * a / pi) ** 0.75))
    return prims


def ao_integrals(centers_bohr: np.ndarray):
    """Overlap, core Hamiltonian and (mu nu|lambda sigma) for one contracted STO-3G s function per center."""
    centers = np.asarray(centers_bohr, dtype=np.float64)
    n = len(centers)
    prims = _primitives(centers)
    npr = len(prims)
    s = np.zeros((n, n))
    t = np.zeros((n, n))
    v = np.zeros((n, n))
    # pair quantities of primitive products
    pair_p = np.zeros((npr, npr))
    pair_k = np.zeros((npr, npr))
    pair_c = np.zeros((npr, npr, 3))
    for i, (ci, a, da) in enumerate(prims):
        for j, (cj, b, db) in enumerate(prims):
            p = a + b
            r2 = float(np.sum((centers[ci] - centers[cj]) ** 2))
            k = np.exp(-a * b / p * r2)
            pc = (a * centers[ci] + b * centers[cj]) / p
            pair_p[i, j] = p
            pair_k[i, j] = da * db * k
            pair_c[i, j] = pc
            sij = (pi / p) ** 1.5 * k
            s[ci, cj] += da * db * sij
            t[ci, cj] += da * db * (a * b / p) * (3.0 - 2.0 * a * b / p * r2) * sij
            for cz in centers:  # unit nuclear charges
                v[ci, cj] += -da * db * 2.0 * pi / p * k * _boys0(p * float(np.sum((pc - cz) ** 2)))
    # electron repulsion integrals over primitives, contracted on the fly
    idx = np.array([pr[0] for pr in prims])
    p_flat = pair_p.reshape(-1)
    k_flat = pair_k.reshape(-1)
    c_flat = pair_c.reshape(-1, 3)
    boys = np.vectorize(_boys0)
    eri = np.zeros((n, n, n, n))
    pair_mu = np.repeat(idx, npr)
    pair_nu = np.tile(idx, npr)
    for ab in range(npr * npr):
        p = p_flat[ab]
        q = p_flat
        rpq2 = np.sum((c_flat[ab] - c_flat) ** 2, axis=1)
        val = 2.0 * pi**2.5 / (p * q * np.sqrt(p + q)) * k_flat[ab] * k_flat * boys(p * q / (p + q) * rpq2)
        np.add.at(eri[pair_mu[ab], pair_nu[ab]], (pair_mu, pair_nu), val)
    return s, t + v, eri


def rhf(s, hcore, eri, n_elec, tol=1e-12, max_iter=200):
    """Restricted Hartree-Fock with DIIS; returns (electronic energy, MO coefficients, orbital energies)."""
    n_occ = n_elec // 2
    w, u = np.linalg.eigh(s)
    x = u @ np.diag(w**-0.5) @ u.T
    def diag(f):
        e, c = np.linalg.eigh(x.T @ f @ x)
        return e, x @ c
    e_orb, c = diag(hcore)
    d = 2.0 * c[:, :n_occ] @ c[:, :n_occ].T
    focks, errs = [], []
    e_old = 0.0
    for _ in range(max_iter):
        j = np.einsum("pqrs,rs->pq", eri, d)
        k = np.einsum("prqs,rs->pq", eri, d)
        f = hcore + j - 0.5 * k
        e_el = 0.5 * float(np.sum(d * (hcore + f)))
        err = x.T @ (f @ d @ s - s @ d @ f) @ x
        focks.append(f)
        errs.append(err)
        focks, errs = focks[-8:], errs[-8:]
        if len(focks) > 1:
            m = len(focks)
            b = -np.ones((m + 1, m + 1))
            b[m, m] = 0.0
            for a_ in range(m):
                for b_ in range(m):
                    b[a_, b_] = float(np.sum(errs[a_] * errs[b_]))
            rhs = np.zeros(m + 1)
            rhs[m] = -1.0
            try:
                coef = np.linalg.solve(b, rhs)[:m]
                f = sum(cf * fk for cf, fk in zip(coef, focks))
            except np.linalg.LinAlgError:
                pass
        e_orb, c = diag(f)
        d = 2.0 * c[:, :n_occ] @ c[:, :n_occ].T
        if abs(e_el - e_old) < tol and float(np.abs(err).max()) < 1e-9:
            break
        e_old = e_el
    j = np.einsum("pqrs,rs->pq", eri, d)
    k = np.einsum("prqs,rs->pq", eri, d)
    f = hcore + j - 0.5 * k
    e_orb, c = diag(f)
    d = 2.0 * c[:, :n_occ] @ c[:, :n_occ].T
    j = np.einsum("pqrs,rs->pq", eri, d)
    k = np.einsum("prqs,rs->pq", eri, d)
    e_el = 0.5 * float(np.sum(d * (2.0 * hcore + j - 0.5 * k)))
    return e_el, c, e_orb


def canonical_mo_coeff(mo_coeff):
    """utils/misc.py:103-108: make the first element larger than 1e-5 of every orbital positive."""
    first = np.argmax(np.abs(mo_coeff) > 1e-5, axis=0)
    sign = np.sign(mo_coeff[first, np.arange(mo_coeff.shape[1])])
    return mo_coeff * sign.reshape(1, -1)


def h_chain_integrals(n_h: int = 4, bond_distance: float = 0.8, with_mo_energy: bool = False):
    """MO integrals of a linear hydrogen chain in STO-3G (molecule.py:72-73; hamiltonian.py:30-48 with the full
    space as active space).  Returns ``int1e [n,n], int2e [n,n,n,n] (chemist notation), e_nuc, e_hf`` and, with
    ``with_mo_energy``, the RHF orbital energies (what MP2 needs, see :func:`mp2_t2`)."""
    if n_h % 2:
        raise ValueError("UCC class does not support open shell system")  # ucc.py:207-208
    z = np.arange(n_h) * bond_distance / BOHR
    centers = np.stack([np.zeros(n_h), np.zeros(n_h), z], axis=1)
    s, hcore, eri = ao_integrals(centers)
    e_nuc = sum(1.0 / abs(z[i] - z[j]) for i in range(n_h) for j in range(i))
    e_el, c, e_orb = rhf(s, hcore, eri, n_h)
    c = canonical_mo_coeff(c)
    int1e = c.T @ hcore @ c
    int2e = np.einsum("up,vq,uvls,lr,st->pqrt", c, c, eri, c, c, optimize=True)
    if with_mo_energy:
        return int1e, int2e, float(e_nuc), float(e_el + e_nuc), np.asarray(e_orb)
    return int1e, int2e, float(e_nuc), float(e_el + e_nuc)


def mp2_t2(int2e, mo_energy, n_occ: int):
    """Closed-shell MP2 doubles amplitudes in PySCF's layout and convention (what ``UCC.__init__`` takes from
    ``mp2.kernel()``, ucc.py:259-265): ``t2[i, j, a, b] = (ia|jb) / (e_i + e_j - e_a - e_b)`` for canonical RHF
    orbitals, and the correlation energy ``sum t2[i,j,a,b] * (2 (ia|jb) - (ib|ja))``.  Returns ``(e_corr, t2)``."""
    int2e = np.asarray(int2e, dtype=np.float64)
    e = np.asarray(mo_energy, dtype=np.float64)
    o, v = slice(0, n_occ), slice(n_occ, len(e))
    ovov = int2e[o, v, o, v]  # (ia|jb) as [i, a, j, b]
    d = e[o][:, None, None, None] - e[v][None, :, None, None] + e[o][None, None, :, None] - e[v][None, None, None, :]
    t_iajb = ovov / d
    e_corr = float(np.einsum("iajb,iajb->", t_iajb, 2.0 * ovov - ovov.transpose(0, 3, 2, 1)))
    return e_corr, np.ascontiguousarray(t_iajb.transpose(0, 2, 1, 3))


def active_space_integrals(int1e, int2e, e_nuc, n_elec_total, active_space):
    """CASCI effective integrals, the quantities ``get_integral_from_hf`` (hamiltonian.py:30-48) takes from PySCF's
    ``CASCI.get_h1eff / get_h2eff``: the doubly occupied inactive orbitals are folded into an effective one-body
    term and ``e_core`` (which also carries the nuclear repulsion, as in the reference).
    ``active_space = (n_active_electrons, n_active_orbitals)`` (ucc.py:218-228)."""
    n_act_elec, n_act = active_space
    n_core = n_elec_total // 2 - n_act_elec // 2
    core = list(range(n_core))
    act = list(range(n_core, n_core + n_act))
    e_core = float(e_nuc)
    for i in core:
        e_core += 2.0 * int1e[i, i]
        for j in core:
            e_core += 2.0 * int2e[i, i, j, j] - int2e[i, j, j, i]
    h_eff = int1e[np.ix_(act, act)].copy()
    for i in core:
        h_eff += 2.0 * int2e[np.ix_(act, act, [i], [i])][:, :, 0, 0] - int2e[np.ix_(act, [i], [i], act)][:, 0, 0, :]
    return h_eff, int2e[np.ix_(act, act, act, act)].copy(), e_core
