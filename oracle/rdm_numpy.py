"""TEST INFRASTRUCTURE ONLY (never imported by the product path): CPU restatement of the reduced density matrices
the reference takes from PySCF.

Reference call sites: UCC.make_rdm1 / make_rdm2, tencirchem/static/ucc.py:755-852, which call
``pyscf.fci.direct_spin1.make_rdm1(civector, norb, nelec)`` and ``make_rdm12(...)[1]`` (pyscf==2.7.0, requirements.txt:6;
the source is not under /root/reference).  Published semantics restated here: spin-traced
``dm1[p,q] = <q^+ p>`` and ``dm2[p,q,r,s] = <p^+ r^+ s q>`` so that
``E = einsum('pq,pq', h1, dm1) + 0.5 * einsum('pqrs,pqrs', eri, dm2)`` — the identity the reference's own docstring
example checks (ucc.py:838-846).  Pinned in tests/ by that identity, by traces, and against brute-force ladder
operators (oracle/bruteforce.py).
"""
import numpy as np

from .civector_numpy import make_strings, single_replacements


def single_replacement_intermediate(civector, norb, nelec):
    """D[p, q, a, b] = <(a,b)| E_pq |c>, E_pq the spin-summed single replacement (rows alpha, columns beta)."""
    k = nelec // 2
    strings = make_strings(norb, k)
    ns = len(strings)
    c = np.asarray(civector, dtype=np.float64).reshape(ns, ns)
    table = single_replacements(strings, norb)
    d = np.zeros((norb, norb, ns, ns))
    for (p, q), (src, dst, sign) in table.items():
        # E_pq |src> = sign |dst>  =>  <dst| E_pq |c> picks c[src]
        np.add.at(d[p, q], (dst, slice(None)), sign[:, None] * c[src, :])        # alpha part: rows
        np.add.at(d[p, q].T, (dst, slice(None)), sign[:, None] * c[:, src].T)    # beta part: columns
    return d


def make_rdm12(civector, norb, nelec):
    """(dm1, dm2) with PySCF's direct_spin1 conventions (see the module docstring)."""
    d = single_replacement_intermediate(civector, norb, nelec)
    c = np.asarray(civector, dtype=np.float64).reshape(d.shape[2:])
    e1 = np.einsum("ab,pqab->pq", c, d)                 # <E_pq>
    d2 = d.reshape(norb * norb, -1)
    gram = (d2 @ d2.T).reshape(norb, norb, norb, norb)  # gram[q,p,r,s] = <E_pq E_rs>
    dm2 = gram.transpose(1, 0, 2, 3).copy()             # [p,q,r,s] = <E_pq E_rs>
    for q in range(norb):
        dm2[:, q, q, :] -= e1
    return e1.T.copy(), dm2


def make_rdm1(civector, norb, nelec):
    return make_rdm12(civector, norb, nelec)[0]


def make_rdm12_hcb(civector, norb, nelec):
    """pUCCD RDMs on the hard-core-boson CI vector, restating tencirchem/static/puccd.py:148-203 (MO basis, before
    embed_rdm_cas / rdm_mo2ao): the mask loops over the pair-occupation strings, the partner address through the
    string -> address map (cistring.strs2addr upstream)."""
    strings = np.asarray(make_strings(norb, nelec // 2), dtype=np.int64)
    addr = {int(s): i for i, s in enumerate(strings)}
    c = np.asarray(civector, dtype=np.float64)
    rdm1 = np.zeros((norb, norb))
    for i in range(norb):
        m = 1 << i
        rdm1[i, i] = 2 * float(c @ (((strings & m) == m) * c))
    rdm2 = np.zeros((norb,) * 4)
    for p in range(norb):
        for q in range(p + 1):
            mq, mp = 1 << q, 1 << p
            mask = (strings & mq) == mq
            if p == q:
                value = float(c @ (mask * c))
            else:
                mask = mask & (((~strings) & mp) == mp)
                partner = np.array([addr.get(int(s), 0) for s in strings ^ (mp | mq)])
                value = float(c @ (mask * c[partner]))
            rdm2[p, q, p, q] = rdm2[q, p, q, p] = value
            if p == q:
                continue
            both = (strings & (mp | mq)) == (mp | mq)
            value = float(c @ (both * c))
            rdm2[p, p, q, q] = rdm2[q, q, p, p] = 2 * value
            rdm2[p, q, q, p] = rdm2[q, p, p, q] = -value
    return rdm1, 2 * rdm2
