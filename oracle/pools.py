"""Excitation-operator pools of the reference ansatz classes, restated (CPU oracle).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  These generators are the
INPUTS of the hot path (ordered ``ex_ops`` tuples + ``param_ids``).  Spin-orbital
numbering (docs/source/faq.rst:16-24): beta = 0..n-1, alpha = n..2n-1, occupied
first inside each spin block.  A tuple lists creation indices then annihilation
indices, T = a+_p a+_q a_r a_s (utils/misc.py:121-129).

Amplitude-based screening/sorting (uccsd.py:159-178) needs MP2 amplitudes from
PySCF and is not reproduced: the generators return the unscreened lists in
generator order (``pick_ex2=False, sort_ex2=False``).
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
