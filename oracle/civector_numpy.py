"""numpy restatement of TenCirChem's ``civector`` UCC engine (CPU oracle).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Parity: pinned
(tests/test_oracle_kat.py checks every PySCF-free known-answer value the
reference ships for this path).

Every function cites the reference lines it follows (paths relative to
``/root/reference``).  Three third-party calls on the path have no source in
the reference tree and are restated from their published algorithms:

* ``pyscf.fci.cistring.make_strings / strs2addr / num_strings`` (pyscf==2.7.0)
  -> :func:`make_strings`, :func:`strs2addr_table`: occupation bit strings of
  k electrons in n orbitals in ascending integer order.
* ``openfermion.jordan_wigner`` (openfermion==1.6.1) -> :func:`jordan_wigner_product`:
  a_j^dagger = (X_j - iY_j)/2 * Z_0..Z_{j-1},  a_j = (X_j + iY_j)/2 * Z_0..Z_{j-1}.
* ``pyscf.fci.direct_nosym.absorb_h1e / contract_2e`` (pyscf==2.7.0)
  -> :func:`absorb_h1e`, :func:`contract_2e`: Knowles-Handy direct CI,
  sigma = sum_{pq,rs} h2e[p,q,r,s] E_pq E_rs c  with spin-summed E_pq.

The tables are kept in the reference's own (wasteful) form on purpose --
full-length uint64 permutation + int8 phase vectors per excitation -- because
this module is the statement of WHAT the engine computes, not how the GPU
path computes it.
"""
from __future__ import annotations

import itertools
from math import comb

import numpy as np

UINT = np.uint64


# --------------------------------------------------------------------------
# CI strings and addresses                      tencirchem/static/ci_utils.py
# --------------------------------------------------------------------------
def make_strings(n_orb: int, n_occ: int) -> np.ndarray:
    """pyscf ``cistring.make_strings(range(n_orb), n_occ)``: all bit strings with
    ``n_occ`` bits set among the lowest ``n_orb`` bits, ascending as integers
    (call sites ci_utils.py:22, :31; hamiltonian.py:161)."""
    if n_occ == 0:
        return np.zeros(1, dtype=np.int64)
    out = np.empty(comb(n_orb, n_occ), dtype=np.int64)
    for i, occ in enumerate(itertools.combinations(range(n_orb), n_occ)):
        v = 0
        for o in occ:
            v |= 1 << o
        out[i] = v
    out.sort()
    return out


def get_ci_strings(n_qubits: int, n_elec: int, hcb: bool, strs2addr: bool = False):
    """ci_utils.py:16-37.  Non-hcb: alpha string in the high half, beta in the
    low half, alpha-major; hcb: plain strings over ``n_qubits`` orbitals."""
    if 2**n_qubits > np.iinfo(UINT).max:  # ci_utils.py:19-20
        raise ValueError(f"Too many qubits: {n_qubits}, try using complex128 datatype")
    if not hcb:
        half = n_qubits // 2
        beta = make_strings(half, n_elec // 2).astype(UINT)
        alpha = beta << UINT(half)
        ci_strings = (alpha.reshape(-1, 1) + beta.reshape(1, -1)).ravel()
        if strs2addr:
            table = np.zeros(2**half, dtype=UINT)
            table[beta] = np.arange(len(beta), dtype=UINT)
            return ci_strings, table
    else:
        ci_strings = make_strings(n_qubits, n_elec // 2).astype(UINT)
        if strs2addr:
            table = np.zeros(2**n_qubits, dtype=UINT)
            table[ci_strings] = np.arange(len(ci_strings), dtype=UINT)
            return ci_strings, table
    return ci_strings


def get_addr(excitation, n_qubits, n_elec, strs2addr, hcb, num_strings=None):
    """ci_utils.py:40-49.  Strings that are not valid map to address 0 because
    the lookup table is zero-initialised (the reference relies on their phase
    being 0)."""
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
        base = (["0"] * (n_qubits // 2 - n_elec // 2) + ["1"] * (n_elec // 2)) * 2
    else:
        base = ["0"] * (n_qubits - n_elec // 2) + ["1"] * (n_elec // 2)
    bits = base[::-1]
    half = len(ex_op) // 2
    for idx in ex_op[half:]:
        bits[idx] = "0"
    for idx in ex_op[:half]:
        bits[idx] = "1"
    return "".join(reversed(bits))


def get_init_civector(len_ci):
    """ci_utils.py:83-87: unit vector on address 0 (the HF determinant)."""
    c = np.zeros(len_ci)
    c[0] = 1.0
    return c


def civector_to_statevector(civector, n_qubits, ci_strings):
    """ci_utils.py:74-76."""
    sv = np.zeros(2**n_qubits)
    sv[ci_strings] = civector
    return sv


def statevector_to_civector(statevector, ci_strings):
    """ci_utils.py:79-80."""
    return statevector[ci_strings]


# --------------------------------------------------------------------------
# Jordan-Wigner of a product of ladder operators (openfermion restated)
# --------------------------------------------------------------------------
_PAULI_MUL = {  # (left, right) -> (phase, result)
    ("X", "X"): (1, "I"), ("Y", "Y"): (1, "I"), ("Z", "Z"): (1, "I"),
    ("X", "Y"): (1j, "Z"), ("Y", "X"): (-1j, "Z"),
    ("Y", "Z"): (1j, "X"), ("Z", "Y"): (-1j, "X"),
    ("Z", "X"): (1j, "Y"), ("X", "Z"): (-1j, "Y"),
}


def _pauli_string_mul(left: dict, right: dict):
    phase = 1
    out = dict(left)
    for q, p in right.items():
        if q not in out:
            out[q] = p
            continue
        ph, res = _PAULI_MUL[(out[q], p)]
        phase *= ph
        if res == "I":
            del out[q]
        else:
            out[q] = res
    return phase, out


def jordan_wigner_product(ladder_ops):
    """JW transform of ``prod_k a^(dagger)_{j_k}``; ``ladder_ops`` is a sequence of
    ``(index, is_creation)``.  Returns ``{((idx, 'X'|'Y'|'Z'), ...): coeff}`` with
    keys sorted by qubit index, the key format of openfermion's QubitOperator."""
    terms = {(): 1.0 + 0j}
    for idx, dag in ladder_ops:
        zs = {q: "Z" for q in range(idx)}
        comp = [({**zs, idx: "X"}, 0.5), ({**zs, idx: "Y"}, -0.5j if dag else 0.5j)]
        new_terms = {}
        for key, coeff in terms.items():
            left = dict(key)
            for pstr, c in comp:
                ph, res = _pauli_string_mul(left, pstr)
                k = tuple(sorted(res.items()))
                new_terms[k] = new_terms.get(k, 0) + coeff * c * ph
        terms = {k: v for k, v in new_terms.items() if abs(v) > 1e-14}
    return terms


def ex_op_ladder(f_idx):
    """utils/misc.py:121-129: T = a+_{f0} a_{f1}  or  a+_{f0} a+_{f1} a_{f2} a_{f3}."""
    half = len(f_idx) // 2
    return [(i, True) for i in f_idx[:half]] + [(i, False) for i in f_idx[half:]]


def fermion_phase_mask_and_sign(f_idx):
    """evolve_civector.py:55-71: Z positions of the JW string of T and the
    global sign (-1 when the first term in sorted order has positive real
    coefficient)."""
    qop = jordan_wigner_product(ex_op_ladder(f_idx))
    first_key = next(iter(qop.keys()))
    mask = 0
    for idx, pauli in first_key:
        if pauli != "Z":
            assert idx in f_idx
            continue
        mask |= 1 << idx
    sign = -1 if sorted(qop.items())[0][1].real > 0 else 1
    return mask, sign


def fermion_phase_closed_form(f_idx):
    """SURVEY.md finding F2: the same (mask, sign) without the Pauli algebra."""
    srt = sorted(f_idx)
    mask = 0
    for lo, hi in zip(srt[0::2], srt[1::2]):
        for b in range(lo + 1, hi):
            mask |= 1 << b
    if len(f_idx) == 2:
        s_t = 1
    else:
        p, q, r, s = f_idx
        s_t = (-1) ** (int(r < s) + int(q < p))
    return mask, -s_t


# --------------------------------------------------------------------------
# Excitation tables                     tencirchem/static/evolve_civector.py
# --------------------------------------------------------------------------
def get_fket_permutation(f_idx, n_qubits, n_elec, ci_strings, strs2addr, hcb):
    """evolve_civector.py:24-29: flip every bit named by the excitation."""
    mask = 0
    for i in f_idx:
        mask += 1 << i
    return get_addr(ci_strings ^ UINT(mask), n_qubits, n_elec, strs2addr, hcb)


def get_fket_phase(f_idx, ci_strings):
    """evolve_civector.py:32-45: support classes.  positive <=> creation bits
    empty and annihilation bits occupied; negative <=> the reverse."""
    half = len(f_idx) // 2
    assert len(f_idx) in (2, 4)
    mask1 = sum(1 << i for i in f_idx[:half])
    mask2 = sum(1 << i for i in f_idx[half:])
    mask = mask1 | mask2
    masked = (ci_strings ^ UINT(mask1)) & UINT(mask)
    return masked == UINT(mask), masked == UINT(0)


_PHASE_CACHE: dict = {}


def get_fermion_phase(f_idx, n_qubits, ci_strings):
    """evolve_civector.py:51-88: ``sign * (+1 if popcount(ci & zmask) odd else -1)``."""
    if f_idx not in _PHASE_CACHE:
        _PHASE_CACHE[f_idx] = fermion_phase_mask_and_sign(f_idx)
    mask, sign = _PHASE_CACHE[f_idx]
    parity = ci_strings & UINT(mask)
    # xor-fold + multiply trick, evolve_civector.py:75-86
    parity ^= parity >> UINT(1)
    parity ^= parity >> UINT(2)
    parity = (parity & UINT(0x1111111111111111)) * UINT(0x1111111111111111)
    parity = (parity >> UINT(60)) & UINT(1)
    return sign * np.sign(parity.astype(np.float64) - 0.5).astype(np.int8)


def get_operators(n_qubits, n_elec, strs2addr, f_idx, ci_strings, hcb):
    """evolve_civector.py:91-106: (partner index, G matrix element, diag of G^2)."""
    if len(set(f_idx)) != len(f_idx):
        raise ValueError(f"Excitation {f_idx} not supported")
    perm = get_fket_permutation(f_idx, n_qubits, n_elec, ci_strings, strs2addr, hcb)
    positive, negative = get_fket_phase(f_idx, ci_strings)
    fket_phase = np.zeros(len(ci_strings))
    fket_phase -= positive
    fket_phase += negative
    if not hcb:
        fket_phase *= get_fermion_phase(f_idx, n_qubits, ci_strings)
    f2ket_phase = np.zeros(len(ci_strings))
    f2ket_phase -= positive
    f2ket_phase -= negative
    return perm, fket_phase, f2ket_phase


_TENSOR_CACHE: dict = {}


def clear_cache():
    _TENSOR_CACHE.clear()
    _PHASE_CACHE.clear()


def get_operator_tensors(n_qubits, n_elec, ex_ops, hcb=False):
    """evolve_civector.py:113-150: stacked ``[M, N]`` tables, cached per ansatz."""
    key = (n_qubits, n_elec, tuple(ex_ops), hcb)
    if key in _TENSOR_CACHE:
        return _TENSOR_CACHE[key]
    ci_strings, strs2addr = get_ci_strings(n_qubits, n_elec, hcb, strs2addr=True)
    m, n = len(ex_ops), len(ci_strings)
    perm_t = np.zeros((m, n), dtype=UINT)
    phase_t = np.zeros((m, n), dtype=np.int8)
    f2_t = np.zeros((m, n), dtype=np.int8)
    for i, f_idx in enumerate(ex_ops):
        perm_t[i], phase_t[i], f2_t[i] = get_operators(n_qubits, n_elec, strs2addr, tuple(f_idx), ci_strings, hcb)
    ret = ci_strings, perm_t, phase_t, f2_t
    _TENSOR_CACHE[key] = ret
    return ret


def get_theta_tensors(params, param_ids):
    """evolve_civector.py:153-162."""
    theta = np.asarray([params[i] for i in param_ids], dtype=np.float64)
    return np.sin(theta), 1 - np.cos(theta)


def evolve_excitation(civector, perm, phase, f2, theta_1mcos, theta_sin):
    """One UCC factor, evolve_civector.py:169-175 / :246-250:
    c += (1-cos t) G^2 c + sin t G c."""
    fket = civector[perm] * phase
    f2ket = civector * f2
    civector += theta_1mcos * f2ket + theta_sin * fket
    return civector


def get_civector(params, n_qubits, n_elec, ex_ops, param_ids, hcb=False, init_state=None):
    """evolve_civector.py:180-198 (cached-table engine ``"civector"``)."""
    ci_strings, perm_t, phase_t, f2_t = get_operator_tensors(n_qubits, n_elec, tuple(ex_ops), hcb)
    theta_sin, theta_1mcos = get_theta_tensors(params, param_ids)
    if init_state is None:
        civector = get_init_civector(len(ci_strings))
    else:
        civector = np.array(init_state, dtype=np.float64)  # copy: do not alias the caller's array (F5 quirk i)
    for j in range(len(perm_t)):
        civector = evolve_excitation(civector, perm_t[j], phase_t[j], f2_t[j], theta_1mcos[j], theta_sin[j])
    return civector.reshape(-1)


def get_civector_nocache(params, n_qubits, n_elec, ex_ops, param_ids, hcb=False, init_state=None):
    """evolve_civector.py:253-271 (engine ``"civector-large"``): tables rebuilt per excitation."""
    theta_sin, theta_1mcos = get_theta_tensors(params, param_ids)
    ci_strings, strs2addr = get_ci_strings(n_qubits, n_elec, hcb, strs2addr=True)
    if init_state is None:
        civector = get_init_civector(len(ci_strings))
    else:
        civector = np.array(init_state, dtype=np.float64)
    for ts, tc_, f_idx in zip(theta_sin, theta_1mcos, ex_ops):
        perm, phase, f2 = get_operators(n_qubits, n_elec, strs2addr, tuple(f_idx), ci_strings, hcb)
        civector = evolve_excitation(civector, perm, phase, f2, tc_, ts)
    return civector.reshape(-1)


def apply_op(hamiltonian, ket):
    """hamiltonian.py:240-252 restricted to the ``fcifunc`` branch."""
    return hamiltonian(ket)


def get_gradients(bra, ket, perm_t, phase_t, f2_t, theta_sin, theta_1mcos):
    """evolve_civector.py:221-243: reverse sweep.  For j = M-1..0 undo factor j on
    both vectors, then grad_j = <bra| G_j |ket>."""
    m = len(perm_t)
    grads = np.zeros(m)
    bra = bra.copy()
    ket = ket.copy()
    for j in range(m - 1, -1, -1):
        ket = evolve_excitation(ket, perm_t[j], phase_t[j], f2_t[j], theta_1mcos[j], -theta_sin[j])
        bra = evolve_excitation(bra, perm_t[j], phase_t[j], f2_t[j], theta_1mcos[j], -theta_sin[j])
        fket = ket[perm_t[j]] * phase_t[j]
        grads[j] = bra @ fket
    return grads


def get_energy_and_grad(params, hamiltonian, n_qubits, n_elec, ex_ops, param_ids, hcb=False, init_state=None):
    """evolve_civector.py:201-218.  Energy excludes ``e_core``; the gradient is
    accumulated over shared parameters and doubled."""
    params = np.asarray(params, dtype=np.float64)
    ket = get_civector(params, n_qubits, n_elec, ex_ops, param_ids, hcb, init_state)
    bra = apply_op(hamiltonian, ket)
    energy = bra @ ket
    _, perm_t, phase_t, f2_t = get_operator_tensors(n_qubits, n_elec, tuple(ex_ops), hcb)
    theta_sin, theta_1mcos = get_theta_tensors(params, param_ids)
    before_sum = get_gradients(bra, ket, perm_t, phase_t, f2_t, theta_sin, theta_1mcos)
    gradients = np.zeros(params.shape)
    for g, pid in zip(before_sum, param_ids):
        gradients[pid] += g
    return energy, 2 * gradients


def get_energy_and_grad_nocache(params, hamiltonian, n_qubits, n_elec, ex_ops, param_ids, hcb=False, init_state=None):
    """evolve_civector.py:274-308."""
    params = np.asarray(params, dtype=np.float64)
    ket = get_civector_nocache(params, n_qubits, n_elec, ex_ops, param_ids, hcb, init_state)
    bra = apply_op(hamiltonian, ket)
    energy = bra @ ket
    ci_strings, strs2addr = get_ci_strings(n_qubits, n_elec, hcb, True)
    theta_sin, theta_1mcos = get_theta_tensors(params, param_ids)
    grads = np.zeros(len(ex_ops))
    for j in range(len(ex_ops) - 1, -1, -1):
        perm, phase, f2 = get_operators(n_qubits, n_elec, strs2addr, tuple(ex_ops[j]), ci_strings, hcb)
        bra = evolve_excitation(bra, perm, phase, f2, theta_1mcos[j], -theta_sin[j])
        ket = evolve_excitation(ket, perm, phase, f2, theta_1mcos[j], -theta_sin[j])
        grads[j] = bra @ (ket[perm] * phase)
    gradients = np.zeros(params.shape)
    for g, pid in zip(grads, param_ids):
        gradients[pid] += g
    return energy, 2 * gradients


def get_energy(params, hamiltonian, n_qubits, n_elec, ex_ops, param_ids, hcb=False, init_state=None):
    """engine_ucc.py:79-89."""
    ket = get_civector(np.asarray(params, dtype=np.float64), n_qubits, n_elec, ex_ops, param_ids, hcb, init_state)
    return ket @ apply_op(hamiltonian, ket)


def apply_excitation(civector, n_qubits, n_elec, f_idx, hcb):
    """evolve_civector.py:311-313: G|c> for one excitation."""
    _, perm_t, phase_t, _ = get_operator_tensors(n_qubits, n_elec, (tuple(f_idx),), hcb)
    return np.asarray(civector, dtype=np.float64)[perm_t[0]] * phase_t[0]


# --------------------------------------------------------------------------
# Hamiltonian action                    tencirchem/static/hamiltonian.py
# --------------------------------------------------------------------------
def absorb_h1e(h1e, eri, norb, nelec, fac=1.0):
    """pyscf ``direct_nosym.absorb_h1e`` (call site hamiltonian.py:148): fold the
    one-body term into the two-body array so that H = sum h2e[pq,rs] E_pq E_rs."""
    h2e = np.array(eri, dtype=np.float64).reshape(norb, norb, norb, norb).copy()
    f1e = np.asarray(h1e, dtype=np.float64) - np.einsum("jiik->jk", h2e) * 0.5
    f1e = f1e * (1.0 / (nelec + 1e-100))
    for k in range(norb):
        h2e[k, k, :, :] += f1e
        h2e[:, :, k, k] += f1e
    return h2e * fac


def single_replacements(strings: np.ndarray, n_orb: int):
    """Per (p, q): arrays (src_addr, dst_addr, sign) with E_pq|src> = sign|dst>
    (pyscf ``cistring.gen_linkstr_index`` content, grouped by orbital pair)."""
    strings = np.asarray(strings, dtype=np.int64)
    lookup = {int(s): i for i, s in enumerate(strings)}
    table = {}
    for p in range(n_orb):
        for q in range(n_orb):
            occ_q = (strings >> q) & 1 == 1
            if p == q:
                src = np.nonzero(occ_q)[0]
                table[(p, q)] = (src, src.copy(), np.ones(len(src)))
                continue
            valid = occ_q & ((strings >> p) & 1 == 0)
            src = np.nonzero(valid)[0]
            new = strings[src] ^ ((1 << p) | (1 << q))
            dst = np.fromiter((lookup[int(s)] for s in new), dtype=np.int64, count=len(new))
            lo, hi = min(p, q), max(p, q)
            between = ((1 << hi) - 1) & ~((1 << (lo + 1)) - 1)
            par = np.array([bin(int(s) & between).count("1") & 1 for s in strings[src]], dtype=np.int64)
            table[(p, q)] = (src, dst, 1.0 - 2.0 * par)
    return table


def contract_2e(h2e, civector, norb, nelec, links=None):
    """pyscf ``direct_nosym.contract_2e`` (call site hamiltonian.py:152), three-step
    Knowles-Handy form: D[rs,K] = sum_J <K|E_rs|J> c_J ; G = h2e . D ;
    sigma_I = sum_{pq,K} <I|E_pq|K> G[pq,K].  Rows are alpha strings, columns beta."""
    k = nelec // 2
    strings = make_strings(norb, k)
    ns = len(strings)
    if links is None:
        links = single_replacements(strings, norb)
    c = np.asarray(civector, dtype=np.float64).reshape(ns, ns)
    d = np.zeros((norb, norb, ns, ns))
    for (r, s), (src, dst, sgn) in links.items():
        d[r, s][dst, :] += sgn[:, None] * c[src, :]  # alpha string = row
        d[r, s][:, dst] += sgn[None, :] * c[:, src]  # beta string = column
    g = (np.asarray(h2e).reshape(norb * norb, norb * norb) @ d.reshape(norb * norb, ns * ns)).reshape(norb, norb, ns, ns)
    sigma = np.zeros((ns, ns))
    for (p, q), (src, dst, sgn) in links.items():
        sigma[dst, :] += sgn[:, None] * g[p, q][src, :]
        sigma[:, dst] += sgn[None, :] * g[p, q][:, src]
    return sigma.reshape(-1)


def get_h_fcifunc_from_integral(int1e, int2e, n_elec):
    """hamiltonian.py:146-155."""
    n_orb = len(int1e)
    h2e = absorb_h1e(int1e, int2e, n_orb, n_elec, 0.5)
    links = single_replacements(make_strings(n_orb, n_elec // 2), n_orb)

    def fci_func(civector):
        return contract_2e(h2e, civector, n_orb, n_elec, links)

    return fci_func


def get_h_fcifunc_hcb_from_integral(int1e, int2e, n_elec):
    """hamiltonian.py:158-183 (hard-core-boson / pUCCD Hamiltonian action)."""
    n_orb = len(int1e)
    ci_strings = make_strings(n_orb, n_elec // 2)
    lookup = np.zeros(2**n_orb, dtype=np.int64)
    lookup[ci_strings] = np.arange(len(ci_strings))

    def fci_func(civector):
        civector = np.asarray(civector, dtype=np.float64)
        res = np.zeros(len(civector))
        for p in range(n_orb):
            for q in range(p + 1):
                if p == q:
                    occ = (ci_strings & (1 << p)) == (1 << p)
                    res += (civector * occ) * (2 * int1e[p, p] + int2e[p, p, p, p])
                else:
                    bitmask = (1 << p) + (1 << q)
                    addr = lookup[ci_strings ^ bitmask]
                    masked_flip = (ci_strings ^ (1 << p)) & bitmask
                    hop = (masked_flip == bitmask) | (masked_flip == 0)
                    res += civector[addr] * hop * int2e[p, q, p, q]
                    both = (ci_strings & bitmask) == bitmask
                    res += (civector * both) * (4 * int2e[p, p, q, q] - 2 * int2e[p, q, p, q])
        return res

    return fci_func


def get_h_from_integral(int1e, int2e, n_elec, hcb):
    """hamiltonian.py:186-196, ``htype="fcifunc"`` only."""
    if hcb:
        return get_h_fcifunc_hcb_from_integral(int1e, int2e, n_elec)
    return get_h_fcifunc_from_integral(int1e, int2e, n_elec)


def symmetrize_int2e(int2e):
    """hamiltonian.py:264-269."""
    int2e = 0.25 * (int2e + int2e.transpose((0, 1, 3, 2)) + int2e.transpose((1, 0, 2, 3)) + int2e.transpose((2, 3, 0, 1)))
    return 0.5 * (int2e + int2e.transpose(3, 2, 1, 0))


def random_integral(nao: int, seed: int = 2077):
    """hamiltonian.py:255-261 (legacy global numpy RNG stream, as in the reference)."""
    np.random.seed(seed)
    int1e = np.random.uniform(-1, 1, size=(nao, nao))
    int2e = np.random.uniform(-1, 1, size=(nao, nao, nao, nao))
    int1e = 0.5 * (int1e + int1e.T)
    return int1e, symmetrize_int2e(int2e)
