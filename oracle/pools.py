"""Excitation-operator pools of the reference ansatz classes, restated (CPU oracle).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  These generators are the
INPUTS of the hot path (ordered ``ex_ops`` tuples + ``param_ids``).  Spin-orbital
numbering (docs/source/faq.rst:16-24): beta = 0..n-1, alpha = n..2n-1, occupied
first inside each spin block.  A tuple lists creation indices then annihilation
indices, T = a+_p a+_q a_r a_s (utils/misc.py:121-129).

``uccsd_ex_ops`` returns the unscreened lists in generator order (``pick_ex2=False,
sort_ex2=False``).  ``uccsd_ex_ops_screened`` restates the amplitude-based screening and
sorting (uccsd.py:107-178) from spatial MP2 amplitudes; the amplitudes themselves come from
PySCF in the reference (ucc.py:259-265) and from ``mp2_t2`` below here -- the textbook
closed-shell MP2 formula in PySCF's ``t2[i, j, a, b]`` layout, pinned by the MP2 energy of H4
and the initial guess of H2 that the reference's notebooks print
(adapt_vqe.ipynb:411, ucc_functions.ipynb:115).
"""
from __future__ import annotations


def ex1_ops(no: int, nv: int):
    """UCC.get_ex1_ops, tencirchem/static/ucc.py:889-936."""
    ops, pids = [], []
    pid = 0
    for i in range(no):
        for a in range(nv):
            ops.append((2 * no + nv + a, no + nv + i))  # alpha -> alpha
            ops.append((no + a, i))  # beta -> beta
            pids += [pid, pid]
            pid += 1
    return ops, pids


def ex2_ops(no: int, nv: int):
    """UCC.get_ex2_ops, tencirchem/static/ucc.py:938-1031."""
    a_o = lambda i: no + nv + i
    a_v = lambda i: 2 * no + nv + i
    b_o = lambda i: i
    b_v = lambda i: no + i
    ops, pids = [], []
    pid = 0
    for i in range(no):  # same-spin doubles
        for j in range(i):
            for a in range(nv):
                for b in range(a):
                    ops.append((a_v(b), a_v(a), a_o(i), a_o(j)))
                    ops.append((b_v(b), b_v(a), b_o(i), b_o(j)))
                    pids += [pid, pid]
                    pid += 1
    for i in range(no):  # opposite-spin doubles
        for j in range(i + 1):
            for a in range(nv):
                for b in range(a + 1):
                    if i == j and a == b:
                        ops.append((b_v(a), a_v(a), a_o(i), b_o(i)))
                        pids.append(pid)
                        pid += 1
                        continue
                    ops.append((b_v(b), a_v(a), a_o(i), b_o(j)))
                    ops.append((a_v(b), b_v(a), b_o(i), a_o(j)))
                    pids += [pid, pid]
                    pid += 1
                    if i != j and a != b:
                        ops.append((b_v(a), a_v(b), a_o(i), b_o(j)))
                        ops.append((a_v(a), b_v(b), b_o(i), a_o(j)))
                        pids += [pid, pid]
                        pid += 1
    return ops, pids


def uccsd_ex_ops(no: int, nv: int):
    """UCCSD.get_ex_ops without screening, tencirchem/static/uccsd.py:107-157."""
    o1, p1 = ex1_ops(no, nv)
    o2, p2 = ex2_ops(no, nv)
    off = max(p1) + 1 if p1 else 0
    return o1 + o2, p1 + [p + off for p in p2]


def spatial2spin_t2(t2):
    """PySCF ``cc.addons.spatial2spin`` for a closed-shell t2 (the call at ucc.py:967-968), restated from its
    published algorithm: spin orbitals interleaved alpha (even), beta (odd);
    t2aa = t2 - t2.transpose(0, 1, 3, 2), t2ab = t2, and every spin block that one alpha-beta swap reaches."""
    import numpy as np

    t2 = np.asarray(t2, dtype=np.float64)
    no, _, nv, _ = t2.shape
    t2aa = t2 - t2.transpose(0, 1, 3, 2)
    out = np.zeros((2 * no, 2 * no, 2 * nv, 2 * nv))
    out[0::2, 0::2, 0::2, 0::2] = t2aa
    out[1::2, 1::2, 1::2, 1::2] = t2aa
    out[0::2, 1::2, 0::2, 1::2] = t2
    out[1::2, 0::2, 1::2, 0::2] = t2
    out[0::2, 1::2, 1::2, 0::2] = -t2.transpose(0, 1, 3, 2)
    out[1::2, 0::2, 0::2, 1::2] = -t2.transpose(0, 1, 3, 2)
    return out


def ex2_init_guess(t2):
    """The amplitude UCC.get_ex2_ops reads for each two-body parameter, ucc.py:1001, :1015, :1022, :1029."""
    t2s = spatial2spin_t2(t2)
    no, nv = t2.shape[0], t2.shape[2]
    guess = []
    for i in range(no):
        for j in range(i):
            for a in range(nv):
                for b in range(a):
                    guess.append(t2s[2 * i, 2 * j, 2 * a, 2 * b])
    for i in range(no):
        for j in range(i + 1):
            for a in range(nv):
                for b in range(a + 1):
                    if i == j and a == b:
                        guess.append(t2s[2 * i, 2 * i + 1, 2 * a, 2 * a + 1])
                        continue
                    guess.append(t2s[2 * i, 2 * j + 1, 2 * a, 2 * b + 1])
                    if i != j and a != b:
                        guess.append(t2s[2 * i, 2 * j + 1, 2 * b, 2 * a + 1])
    return [float(x) for x in guess]


def pick_and_sort(ex_ops, param_ids, init_guess, do_pick=True, do_sort=True, eps=1e-12):
    """UCCSD.pick_and_sort, tencirchem/static/uccsd.py:159-178."""
    import numpy as np

    if do_sort:
        sorted_ex_ops = sorted(zip(ex_ops, param_ids), key=lambda x: -np.abs(init_guess[x[1]]))
    else:
        sorted_ex_ops = list(zip(ex_ops, param_ids))
    ret_ex_ops, ret_param_ids = [], []
    for ex_op, param_id in sorted_ex_ops:
        if do_pick and np.abs(init_guess[param_id]) < eps:
            continue
        ret_ex_ops.append(ex_op)
        ret_param_ids.append(param_id)
    unique_ids = np.unique(ret_param_ids)
    ret_init_guess = np.array(init_guess)[unique_ids]
    id_mapping = {old: new for new, old in enumerate(unique_ids)}
    return ret_ex_ops, [id_mapping[i] for i in ret_param_ids], list(ret_init_guess)


def uccsd_ex_ops_screened(no: int, nv: int, t2, do_pick=True, do_sort=True, eps=1e-12):
    """UCCSD.get_ex_ops with MP2 amplitudes (t1 = 0), tencirchem/static/uccsd.py:107-157."""
    o1, p1 = ex1_ops(no, nv)
    o2, p2 = ex2_ops(no, nv)
    o2, p2, g2 = pick_and_sort(o2, p2, ex2_init_guess(t2), do_pick, do_sort, eps)
    return o1 + o2, p1 + [i + max(p1) + 1 for i in p2], [0.0] * (max(p1) + 1) + g2


def mp2_t2(int2e, mo_energy, n_occ):
    """Closed-shell MP2 amplitudes in PySCF's layout: t2[i,j,a,b] = (ia|jb) / (e_i + e_j - e_a - e_b); plain loops."""
    import numpy as np

    n = len(mo_energy)
    nv = n - n_occ
    t2 = np.zeros((n_occ, n_occ, nv, nv))
    e_corr = 0.0
    for i in range(n_occ):
        for j in range(n_occ):
            for a in range(nv):
                for b in range(nv):
                    g = int2e[i, n_occ + a, j, n_occ + b]
                    t2[i, j, a, b] = g / (mo_energy[i] + mo_energy[j] - mo_energy[n_occ + a] - mo_energy[n_occ + b])
                    e_corr += t2[i, j, a, b] * (2.0 * g - int2e[i, n_occ + b, j, n_occ + a])
    return e_corr, t2


def kupccgsd_ex_ops(no: int, nv: int, k: int = 3):
    """KUPCCGSD.get_ex_ops / get_ex1_ops / get_ex2_ops, tencirchem/static/kupccgsd.py:127-260."""
    n = no + nv
    o1, p1, o2, p2 = [], [], [], []
    pid = 0
    for a in range(n):
        for i in range(a):
            o1 += [(n + a, n + i), (a, i)]
            p1 += [pid, pid]
            o2.append((a, n + a, n + i, i))
            p2.append(pid)
            pid += 1
    ops, pids = [], [-1]
    for _ in range(k):
        ops += o2 + o1
        pids += [i + pids[-1] + 1 for i in p2]
        pids += [i + pids[-1] + 1 for i in p1]
    return ops, pids[1:]


def puccd_ex_ops(no: int, nv: int):
    """PUCCD.get_ex_ops (hard-core-boson pair excitations), tencirchem/static/puccd.py:101-146."""
    ops = []
    for i in range(no):
        for a in range(nv - 1, -1, -1):
            ops.append((no + a, i))
    return ops, list(range(len(ops)))
