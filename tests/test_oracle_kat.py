"""CPU: pins the oracle against every PySCF-free known-answer value the reference publishes for this
path (tests/golden/published_kat.json) and against the independent brute-force
second oracle."""
import json
import os

import numpy as np
import pytest

from oracle import bruteforce as bf
from oracle import civector_numpy as O
from oracle import pools

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "published_kat.json")))


def h2_integrals():
    k = KAT["h2_sto3g"]
    int1e = np.diag([k["h00"], k["h11"]])
    int2e = np.zeros((2, 2, 2, 2))
    int2e[0, 0, 0, 0] = k["J00"]
    int2e[1, 1, 1, 1] = k["J11"]
    int2e[0, 0, 1, 1] = int2e[1, 1, 0, 0] = k["J01"]
    for idx in [(0, 1, 0, 1), (0, 1, 1, 0), (1, 0, 0, 1), (1, 0, 1, 0)]:
        int2e[idx] = k["K01"]
    return int1e, int2e


def test_ci_string_doctests():
    k = KAT["ci_strings"]
    cs, s2a = O.get_ci_strings(4, 2, False, strs2addr=True)
    assert cs.tolist() == k["h2_ci_strings"]
    assert s2a.tolist() == k["h2_strs2addr"]
    for bits, addr in k["h2_addr"].items():
        assert int(O.get_addr(np.array([int(bits, 2)], dtype=np.uint64), 4, 2, s2a, False)[0]) == addr
    cs_h, s2a_h = O.get_ci_strings(2, 2, True, strs2addr=True)
    for bits, addr in k["h2_hcb_addr"].items():
        assert int(O.get_addr(np.array([int(bits, 2)], dtype=np.uint64), 2, 2, s2a_h, True)[0]) == addr
    # faq.rst:38-43: HF determinant and one excited determinant of H4
    assert format(int(O.get_ci_strings(8, 4, False)[0]), "08b") == k["h4_hf_bitstring"]
    assert O.get_ex_bitstring(8, 4, (6, 2, 0, 4), False) == k["h4_ex_6204_bitstring"]


def test_too_many_qubits():
    with pytest.raises(ValueError, match="Too many qubits"):
        O.get_ci_strings(64, 2, False)


def test_apply_excitation_doctest():
    k = KAT["apply_excitation"]
    res = O.apply_excitation(np.array(k["state"], dtype=float), 4, 2, tuple(k["ex_op"]), False)
    assert res.tolist() == k["result"]


def test_pool_doctests():
    p = KAT["pools"]
    ops, ids = pools.uccsd_ex_ops(1, 1)
    assert [list(o) for o in ops] == p["uccsd_h2"]["ex_ops"] and ids == p["uccsd_h2"]["param_ids"]
    ops, ids = pools.kupccgsd_ex_ops(1, 1, 3)
    assert [list(o) for o in ops] == p["kupccgsd_h2"]["ex_ops"] and ids == p["kupccgsd_h2"]["param_ids"]
    ops, ids = pools.puccd_ex_ops(1, 1)
    assert [list(o) for o in ops] == p["puccd_h2"]["ex_ops"] and ids == p["puccd_h2"]["param_ids"]
    ops, ids = pools.ex1_ops(2, 2)
    assert [list(o) for o in ops] == p["h4_ex1"]["ex_ops"] and ids == p["h4_ex1"]["param_ids"]
    ops, ids = pools.ex2_ops(2, 2)
    assert [list(o) for o in ops] == p["h4_ex2"]["ex_ops"] and ids == p["h4_ex2"]["param_ids"]


def test_h2_published_state_and_energy():
    k = KAT["h2_sto3g"]
    int1e, int2e = h2_integrals()
    h = O.get_h_fcifunc_from_integral(int1e, int2e, 2)
    ops = [tuple(o) for o in k["ex_ops"]]
    params = np.array(k["params"])
    c = O.get_civector(params, 4, 2, ops, k["param_ids"])
    np.testing.assert_allclose(c, k["civector"], atol=k["civector_atol"])
    e, g = O.get_energy_and_grad(params, h, 4, 2, ops, k["param_ids"])
    assert abs(e + k["e_nuc"] - k["e_ucc"]) < 1e-9  # published params are rounded to 9 digits
    assert np.abs(g).max() < 1e-7  # the published point is the optimum
    e_hf = O.get_energy(np.zeros(2), h, 4, 2, ops, k["param_ids"])
    assert abs(e_hf + k["e_nuc"] - k["e_hf"]) < 1e-13
    # the nocache ("civector-large") engine gives the same numbers
    e2, g2 = O.get_energy_and_grad_nocache(params, h, 4, 2, ops, k["param_ids"])
    assert abs(e - e2) < 1e-14 and np.abs(g - g2).max() < 1e-14


def _all_test_ops(no, nv):
    ops, _ = pools.uccsd_ex_ops(no, nv)
    gops, _ = pools.kupccgsd_ex_ops(no, nv, 1)
    extra = [(4, 0, 3, 7), (7, 3, 0, 4), (6, 2, 0, 4), (2, 7, 5, 0)] if no + nv == 4 else []
    return list(dict.fromkeys(ops + gops + extra))


@pytest.mark.parametrize("no,nv", [(1, 1), (2, 2), (2, 3), (3, 3)])
def test_sign_rule_closed_form_matches_jordan_wigner(no, nv):
    for op in _all_test_ops(no, nv):
        assert O.fermion_phase_mask_and_sign(op) == O.fermion_phase_closed_form(op), op


@pytest.mark.parametrize("no,nv", [(2, 2), (3, 3), (2, 3)])
def test_tables_are_matrix_elements_of_G(no, nv):
    """fket_phase[I] == <I| T - T^+ |perm(I)> and f2 == diag(G^2), against brute force."""
    n = no + nv
    cs, s2a = O.get_ci_strings(2 * n, 2 * no, False, True)
    for op in _all_test_ops(no, nv):
        perm, phase, f2 = O.get_operators(2 * n, 2 * no, s2a, op, cs, False)
        g = bf.dense_excitation_generator(op, 2 * n, 2 * no)
        dense = np.zeros_like(g)
        dense[np.arange(len(cs)), perm.astype(np.int64)] = phase
        # rows outside the support have phase 0 whatever their (meaningless) partner index
        np.testing.assert_array_equal(dense, g, err_msg=str(op))
        np.testing.assert_array_equal(np.diag(g @ g), f2, err_msg=str(op))


def test_hcb_tables_against_brute_force():
    cs, s2a = O.get_ci_strings(5, 4, True, True)
    ops, _ = pools.puccd_ex_ops(2, 3)
    for op in ops + [(1, 3), (4, 0)]:
        perm, phase, f2 = O.get_operators(5, 4, s2a, op, cs, True)
        g = bf.dense_excitation_generator_hcb(op, 5, 4)
        dense = np.zeros_like(g)
        dense[np.arange(len(cs)), perm.astype(np.int64)] = phase
        np.testing.assert_array_equal(dense, g)


@pytest.mark.parametrize("no,nv", [(2, 2), (2, 3)])
def test_state_matches_expm_product(no, nv):
    n = no + nv
    ops, pids = pools.uccsd_ex_ops(no, nv)
    np.random.seed(7)
    params = np.random.rand(max(pids) + 1) - 0.5
    c = O.get_civector(params, 2 * n, 2 * no, ops, pids)
    c_bf = bf.ucc_state(params, ops, pids, 2 * n, 2 * no)
    assert np.abs(c - c_bf).max() < 1e-13
    c2 = O.get_civector_nocache(params, 2 * n, 2 * no, ops, pids)
    assert np.abs(c - c2).max() < 1e-15


@pytest.mark.parametrize("n,nelec", [(2, 2), (4, 4), (4, 2), (5, 4)])
def test_sigma_equals_dense_hamiltonian(n, nelec):
    """test_hamiltonian_fcifunc (tests/static/test_hamiltonian.py:44-64) pins the lowest eigenvalue; here the
    whole matrix is checked against the second-quantised definition of hamiltonian.py:51-96."""
    int1e, int2e = O.random_integral(n)
    h = O.get_h_fcifunc_from_integral(int1e, int2e, nelec)
    dense = bf.dense_hamiltonian(int1e, int2e, nelec)
    dim = dense.shape[0]
    applied = np.stack([h(np.eye(dim)[i]) for i in range(dim)], axis=1)
    assert np.abs(applied - dense).max() < 1e-12
    if (n, nelec) == (4, 4):
        # session-derived check value recorded in SURVEY.md appendix
        assert abs(np.linalg.eigvalsh(dense)[0] - (-5.118408240910737)) < 1e-12


def test_sigma_hcb_equals_dense():
    int1e, int2e = O.random_integral(5)
    h = O.get_h_fcifunc_hcb_from_integral(int1e, int2e, 4)
    dense = bf.dense_hamiltonian_hcb(int1e, int2e, 4)
    dim = dense.shape[0]
    applied = np.stack([h(np.eye(dim)[i]) for i in range(dim)], axis=1)
    assert np.abs(applied - dense).max() < 1e-12


def test_random_44_check_values_and_gradient():
    int1e, int2e = O.random_integral(4)
    ops, pids = pools.uccsd_ex_ops(2, 2)
    np.random.seed(2077)
    params = np.random.rand(15) - 0.5
    h = O.get_h_fcifunc_from_integral(int1e, int2e, 4)
    e, g = O.get_energy_and_grad(params, h, 8, 4, ops, pids)
    assert abs(e - 0.06402312039336584) < 1e-12
    eps = 1e-6
    fd = np.zeros_like(g)
    for i in range(len(params)):
        d = np.zeros_like(params)
        d[i] = eps
        fd[i] = (O.get_energy(params + d, h, 8, 4, ops, pids) - O.get_energy(params - d, h, 8, 4, ops, pids)) / (2 * eps)
    assert np.abs(fd - g).max() < 1e-8


def test_golden_file_is_current():
    """The committed fixtures are what the oracle produces today."""
    gold = np.load(os.path.join(HERE, "golden", "oracle_vectors.npz"))
    int1e, int2e = O.random_integral(5)
    ops, pids = pools.uccsd_ex_ops(2, 3)
    h = O.get_h_fcifunc_from_integral(int1e, int2e, 4)
    e, g = O.get_energy_and_grad(gold["uccsd_2_3_params"], h, 10, 4, ops, pids)
    assert abs(e - float(gold["uccsd_2_3_energy"])) < 1e-13
    np.testing.assert_allclose(g, gold["uccsd_2_3_grad"], atol=1e-13)


# ---- literal hydrogen chains (BASELINE.json configs 0-2): PySCF-free STO-3G integrals + RHF ----------------------
def test_hchain_integrals_reproduce_published_h2():
    """tencirchem_b200/hchain.py (s-Gaussian integrals, RHF, MO transform) against the MO integrals and energies the
    reference publishes for H2 (ucc_functions.ipynb)."""
    from tencirchem_b200.hchain import h_chain_integrals

    k = KAT["h2_sto3g"]
    int1e, int2e, e_nuc, e_hf = h_chain_integrals(2, 0.741)
    assert abs(e_nuc - k["e_nuc"]) < 1e-14 and abs(e_hf - k["e_hf"]) < 1e-12
    assert abs(int1e[0, 0] - k["h00"]) < 1e-12 and abs(int1e[1, 1] - k["h11"]) < 1e-12 and abs(int1e[0, 1]) < 1e-12
    for idx, name in [((0, 0, 0, 0), "J00"), ((1, 1, 1, 1), "J11"), ((0, 0, 1, 1), "J01"), ((0, 1, 0, 1), "K01"), ((0, 1, 1, 0), "K01")]:
        assert abs(int2e[idx] - k[name]) < 1e-12, name


def test_h8_active_space_published_energies():
    from scipy.optimize import minimize

    from tencirchem_b200.hchain import active_space_integrals, h_chain_integrals

    k = KAT["h_chain_sto3g"]
    int1e, int2e, e_nuc, e_hf = h_chain_integrals(8)
    assert abs(e_hf - k["h8_e_hf"]) < 1e-11
    h1, h2, e_core = active_space_integrals(int1e, int2e, e_nuc, 8, (2, 2))
    ops, pids = pools.uccsd_ex_ops(1, 1)
    h = O.get_h_fcifunc_from_integral(h1, h2, 2)
    e0, g0 = O.get_energy_and_grad(np.zeros(2), h, 4, 2, ops, pids)
    assert abs(e0 + e_core - k["h8_e_hf"]) < 1e-11  # ucc.energy(zeros) == e_hf, cell 37
    np.testing.assert_allclose(g0, k["h8_cas22_grad_at_zero"], atol=3e-7)  # cell 38
    res = minimize(lambda x: tuple(np.asarray(v) for v in O.get_energy_and_grad(x, h, 4, 2, ops, pids)), np.zeros(2), jac=True, method="L-BFGS-B")
    assert abs(res.fun + e_core - k["h8_cas22_uccsd"]) < 1e-7  # cells 31/34


def test_h4_adapt_vqe_pool_gradients_published():
    """Iteration 0 of the reference's ADAPT-VQE tutorial (adapt_vqe.ipynb cell 9): 2 <HF| H G_k |HF> for every pool
    operator of the literal H4 chain -- pins the sign convention of every excitation kind on a real molecule."""
    from tencirchem_b200.hchain import h_chain_integrals

    k = KAT["h_chain_sto3g"]
    int1e, int2e, e_nuc, e_hf = h_chain_integrals(4)
    assert abs(e_hf - k["h4_e_hf"]) < 1e-6
    assert abs(np.linalg.eigvalsh(bf.dense_hamiltonian(int1e, int2e, 4))[0] + e_nuc - k["h4_e_fci"]) < 1e-6
    h = O.get_h_fcifunc_from_integral(int1e, int2e, 4)
    psi = O.get_init_civector(36)
    bra = h(psi)
    grads = {}
    for ops_, g_pub in k["h4_adapt_iter0"]:
        g = 2 * sum(bra @ O.apply_excitation(psi, 8, 4, tuple(op), False) for op in ops_)
        assert abs(g - g_pub) < 2e-7, (ops_, g, g_pub)
        grads[str(ops_)] = g
    singles = [2 * sum(bra @ O.apply_excitation(psi, 8, 4, op, False) for op in grp) for grp in ([(6, 4), (2, 0)], [(7, 5), (3, 1)], [(7, 4), (3, 0)], [(6, 5), (2, 1)])]
    assert abs(np.linalg.norm(list(grads.values()) + singles) - k["h4_adapt_iter0_norm"]) < 2e-7


def test_rdm_oracle_against_bruteforce_and_energy_identity():
    """oracle/rdm_numpy.py (PySCF direct_spin1.make_rdm12 conventions used by ucc.py:755-852): element-wise against
    brute-force ladder operators, plus the identity of the reference's own docstring example (ucc.py:838-846)."""
    from oracle import bruteforce as bf
    from oracle import rdm_numpy as R

    n, ne = 3, 4
    int1e, int2e = O.random_integral(n)
    ns = len(O.make_strings(n, ne // 2))
    rng = np.random.default_rng(11)
    c = rng.standard_normal(ns * ns)
    c /= np.linalg.norm(c)
    dm1, dm2 = R.make_rdm12(c, n, ne)
    dets = bf.basis(2 * n, ne)
    assert len(dets) == ns * ns
    for p in range(n):
        for q in range(n):
            words = [([(off + q, True), (off + p, False)], 1.0) for off in (0, n)]
            assert abs(c @ bf.dense_from_words(dets, words) @ c - dm1[p, q]) < 1e-12
            for r in range(n):
                for s in range(n):
                    words = [([(o1 + p, True), (o2 + r, True), (o2 + s, False), (o1 + q, False)], 1.0)
                             for o1 in (0, n) for o2 in (0, n)]
                    assert abs(c @ bf.dense_from_words(dets, words) @ c - dm2[p, q, r, s]) < 1e-12
    h = O.get_h_fcifunc_from_integral(int1e, int2e, ne)
    e = np.einsum("pq,pq", int1e, dm1) + 0.5 * np.einsum("pqrs,pqrs", int2e, dm2)
    assert abs(e - c @ h(c)) < 1e-12
    assert abs(np.trace(dm1) - ne) < 1e-12


@pytest.mark.parametrize("n,ne", [(4, 4), (5, 4), (5, 6)])
def test_hcb_rdm_oracle_against_the_fermionic_rdm_oracle(n, ne):
    """oracle make_rdm12_hcb (puccd.py:148-203) against the fermionic RDM oracle on the same state embedded in the
    full CI space (alpha string == beta string; the pair-creation reordering sign is common to all strings) -- the
    comparison the reference's own test makes through UCC with paired excitations (tests/static/test_puccd.py:48-57) --
    plus the energy identity with the hcb Hamiltonian (hamiltonian.py:158-183)."""
    from oracle import rdm_numpy as R

    int1e, int2e = O.random_integral(n)
    ns = len(O.make_strings(n, ne // 2))
    rng = np.random.default_rng(5)
    c = rng.standard_normal(ns)
    c /= np.linalg.norm(c)
    full = np.zeros((ns, ns))
    full[np.arange(ns), np.arange(ns)] = c
    f1, f2 = R.make_rdm12(full.ravel(), n, ne)
    h1, h2 = R.make_rdm12_hcb(c, n, ne)
    assert np.abs(h1 - f1).max() < 1e-12
    assert np.abs(h2 - f2).max() < 1e-12
    h = O.get_h_from_integral(int1e, int2e, ne, True)
    e = np.einsum("pq,pq", int1e, h1) + 0.5 * np.einsum("pqrs,pqrs", int2e, h2)
    assert abs(e - c @ h(c)) < 1e-12


def test_rdm_oracle_reproduces_the_reference_docstring_example_h2():
    """ucc.py:838-846 (make_rdm2 doctest): for the H2 HF state [1, 0, 0, 0],
    int1e.rdm1 + 1/2 int2e.rdm2 + e_nuc equals the published HF energy."""
    from oracle import rdm_numpy as R

    k = KAT["h2_sto3g"]
    int1e, int2e = h2_integrals()
    rdm1, rdm2 = R.make_rdm12(np.array([1.0, 0.0, 0.0, 0.0]), 2, 2)
    e_hf = int1e.ravel() @ rdm1.ravel() + 0.5 * int2e.ravel() @ rdm2.ravel()
    assert abs(e_hf + k["e_nuc"] - k["e_hf"]) < 1e-10  # the doctest's own tolerance
    # and at the published UCC optimum
    c = np.array(k["civector"])
    rdm1, rdm2 = R.make_rdm12(c, 2, 2)
    e = int1e.ravel() @ rdm1.ravel() + 0.5 * int2e.ravel() @ rdm2.ravel()
    assert abs(e + k["e_nuc"] - k["e_ucc"]) < 1e-8  # the published amplitudes carry 9 digits


# ---- sparse oracle (oracle/sparse.py): pinned against the dense restatement ---------------------------------------
@pytest.mark.parametrize("n", [2, 3, 4, 6])
def test_sparse_oracle_matches_dense_oracle(n):
    """Energy, all gradients, state, sigma on the full space and G|c> of the sparse (dict-of-determinants) oracle equal
    the dense oracle's for full UCCSD with dense random parameters."""
    from oracle import sparse as S

    no = n // 2
    ex_ops, pids = pools.uccsd_ex_ops(no, n - no)
    int1e, int2e = O.random_integral(n)
    np.random.seed(2077)
    params = np.random.rand(max(pids) + 1) - 0.5
    h = O.get_h_fcifunc_from_integral(int1e, int2e, n)
    e_ref, g_ref = O.get_energy_and_grad(params, h, 2 * n, n, ex_ops, pids)
    c_ref = O.get_civector(params, 2 * n, n, ex_ops, pids)
    e, g, psi, _ = S.get_energy_and_grad(params, int1e, int2e, 2 * n, n, ex_ops, pids)
    assert abs(e - e_ref) < 1e-13
    assert np.abs(np.array([g[p] for p in range(len(params))]) - g_ref).max() < 1e-13
    assert np.abs(S.to_dense(psi, 2 * n, n) - c_ref).max() < 1e-14
    ci = O.get_ci_strings(2 * n, n, False)
    assert np.abs(S.apply_h(psi, int1e, int2e, n, ci) - h(c_ref)).max() < 1e-13
    for f in list(ex_ops[:4]) + list(ex_ops[-4:]):
        a = S.to_dense(S.apply_excitation(psi, 2 * n, f), 2 * n, n)
        assert np.abs(a - O.apply_excitation(c_ref, 2 * n, n, f, False)).max() < 1e-14
    O.clear_cache()


def test_sparse_golden_8o_matches_dense_oracle():
    """The committed sparse known answers are what the dense oracle gives: the (8e,8o) case of make_sparse_golden.py
    (sparse random initial state, a few non-zero parameters of every kind, full UCCSD op list) re-evaluated densely."""
    import json
    import os

    from oracle import sparse as S

    case = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sparse_parity_8o.json")))
    n = case["n_orb"]
    ex_ops, pids = pools.uccsd_ex_ops(n // 2, n - n // 2)
    int1e, int2e = O.random_integral(n)
    params = np.zeros(case["n_params"])
    params[case["param_ids_nonzero"]] = case["param_values"]
    ci = O.get_ci_strings(2 * n, n, False)
    init = S._lookup(np.array(case["init_strings"], dtype=np.uint64), np.array(case["init_amps"]), ci)
    h = O.get_h_fcifunc_from_integral(int1e, int2e, n)
    e, g = O.get_energy_and_grad(params, h, 2 * n, n, ex_ops, pids, init_state=init)
    assert abs(e - case["energy"]) < 1e-13
    assert np.abs(g[case["grad_ids"]] - case["grad_values"]).max() < 1e-13
    c = O.get_civector(params, 2 * n, n, ex_ops, pids, init_state=init)
    pos = np.searchsorted(ci, np.array(case["psi_sample_strings"], dtype=np.uint64))
    assert np.abs(c[pos] - case["psi_sample_amps"]).max() < 1e-14
    pos = np.searchsorted(ci, np.array(case["sigma_sample_strings"], dtype=np.uint64))
    assert np.abs(h(c)[pos] - case["sigma_sample_values"]).max() < 1e-13
    O.clear_cache()
