"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the same
seeded inputs, against the committed golden fixtures and the reference's published values, plus
size-independent properties at larger sizes.

Tolerances (BASELINE.json north_star): strings / index maps / signs bit-exact (tests/test_host_cpu.py);
energies within 1e-10 Ha; gradients within 1e-8 Ha/rad; CI amplitudes within 1e-12.
"""
import json
import os

import numpy as np
import pytest

from oracle import bruteforce as bf
from oracle import civector_numpy as O
from oracle import pools
from tencirchem_b200 import KUPCCGSD, PUCCD, UCC, UCCSD, engine_ucc
from tencirchem_b200.evolve_civector import clear_cache
from tencirchem_b200.hamiltonian import get_h_from_integral, random_integral
from tencirchem_b200.plan import Plan

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "published_kat.json")))
GOLD = np.load(os.path.join(HERE, "golden", "oracle_vectors.npz"))

E_TOL = 1e-10
G_TOL = 1e-8
C_TOL = 1e-12


def _params(n, seed=2077):
    np.random.seed(seed)
    return np.random.rand(n) - 0.5


def test_native_library_is_what_runs():
    import ctypes

    from tencirchem_b200 import _lib

    lib = _lib.load()
    assert lib.tcc_device_count() >= 1
    assert isinstance(lib, ctypes.CDLL)
    maps = open(f"/proc/{os.getpid()}/maps").read()
    assert "libtcc_b200.so" in maps


def test_h2_published_values():
    k = KAT["h2_sto3g"]
    int1e = np.diag([k["h00"], k["h11"]])
    int2e = np.zeros((2, 2, 2, 2))
    int2e[0, 0, 0, 0] = k["J00"]
    int2e[1, 1, 1, 1] = k["J11"]
    int2e[0, 0, 1, 1] = int2e[1, 1, 0, 0] = k["J01"]
    for idx in [(0, 1, 0, 1), (0, 1, 1, 0), (1, 0, 0, 1), (1, 0, 1, 0)]:
        int2e[idx] = k["K01"]
    ucc = UCCSD.from_integral(int1e, int2e, 2, k["e_nuc"])
    assert [list(o) for o in ucc.ex_ops] == k["ex_ops"] and ucc.param_ids == k["param_ids"]
    np.testing.assert_array_equal(ucc.civector([0, 0]), [1.0, 0.0, 0.0, 0.0])  # ucc.py:466-467 doctest
    assert round(ucc.energy([0, 0]), 8) == -1.11670614  # ucc.py:633-637 doctest
    c = ucc.civector(k["params"])
    np.testing.assert_allclose(c, k["civector"], atol=k["civector_atol"])
    e, g = ucc.energy_and_grad(k["params"])
    assert abs(e - k["e_ucc"]) < 1e-9
    assert np.abs(g).max() < 1e-7
    assert abs(ucc.energy([0, 0]) - k["e_hf"]) < 1e-13
    e_opt = ucc.kernel()
    assert abs(e_opt - k["e_ucc"]) < 1e-8


def test_apply_excitation_doctest_and_statevector_form():
    k = KAT["apply_excitation"]
    res = engine_ucc.apply_excitation(k["state"], 4, 2, tuple(k["ex_op"]), hcb=False)
    assert res.tolist() == k["result"]
    # statevector in, statevector out (engine_ucc.py:145-159)
    ci = O.get_ci_strings(4, 2, False)
    sv = np.zeros(16)
    sv[ci] = k["state"]
    out = engine_ucc.apply_excitation(sv, 4, 2, tuple(k["ex_op"]), hcb=False)
    assert out.shape == (16,) and out[ci].tolist() == k["result"]


CASES = [("uccsd", 1, 1), ("uccsd", 2, 2), ("uccsd", 3, 3), ("uccsd", 2, 3), ("uccsd", 3, 1), ("uccsd", 4, 4),
         ("kupccgsd", 2, 2), ("kupccgsd", 3, 2), ("puccd", 2, 3), ("puccd", 3, 3), ("puccd", 4, 4)]


def _build(kind, no, nv):
    n = no + nv
    int1e, int2e = random_integral(n)
    cls = {"uccsd": UCCSD, "kupccgsd": KUPCCGSD, "puccd": PUCCD}[kind]
    ucc = cls.from_integral(int1e, int2e, 2 * no)
    return ucc, int1e, int2e


@pytest.mark.parametrize("kind,no,nv", CASES)
def test_energy_grad_civector_against_oracle(kind, no, nv):
    ucc, int1e, int2e = _build(kind, no, nv)
    params = _params(ucc.n_params)
    h = O.get_h_from_integral(int1e, int2e, ucc.n_elec, ucc.hcb)
    e_ref, g_ref = O.get_energy_and_grad(params, h, ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, ucc.hcb)
    c_ref = O.get_civector(params, ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, ucc.hcb)
    e, g = ucc.energy_and_grad(params)
    c = ucc.civector(params)
    assert np.abs(c - c_ref).max() < C_TOL
    assert abs(e - e_ref) < E_TOL
    assert np.abs(g - g_ref).max() < G_TOL
    assert abs(ucc.energy(params) - e_ref) < E_TOL
    O.clear_cache()


@pytest.mark.parametrize("tag", ["uccsd_2_2", "uccsd_3_3", "uccsd_2_3", "kupccgsd_2_2", "kupccgsd_3_2", "puccd_2_3", "puccd_3_3"])
def test_against_committed_golden_vectors(tag):
    kind, no, nv = tag.split("_")
    ucc, _, _ = _build(kind, int(no), int(nv))
    params = GOLD[tag + "_params"]
    e, g = ucc.energy_and_grad(params)
    c = ucc.civector(params)
    assert np.abs(c - GOLD[tag + "_civector"]).max() < C_TOL
    assert abs(e - float(GOLD[tag + "_energy"])) < E_TOL
    assert np.abs(g - GOLD[tag + "_grad"]).max() < G_TOL
    sigma = ucc.hamiltonian(c)
    assert np.abs(sigma - GOLD[tag + "_sigma"]).max() < 1e-11


@pytest.mark.parametrize("n,nelec,hcb", [(2, 2, False), (4, 4, False), (4, 2, False), (5, 4, False), (5, 6, False), (6, 6, False), (5, 4, True), (6, 6, True)])
def test_sigma_against_dense_hamiltonian(n, nelec, hcb):
    """The GPU H.c against the second-quantised dense matrix (what test_hamiltonian_fcifunc pins upstream)."""
    int1e, int2e = random_integral(n)
    h = get_h_from_integral(int1e, int2e, nelec, hcb)
    dense = bf.dense_hamiltonian_hcb(int1e, int2e, nelec) if hcb else bf.dense_hamiltonian(int1e, int2e, nelec)
    rng = np.random.default_rng(1)
    for _ in range(3):
        x = rng.standard_normal(dense.shape[0])
        assert np.abs(h(x) - dense @ x).max() < 1e-11
    if not hcb and n <= 4:
        w = np.linalg.eigvalsh(dense)[0]
        applied = np.stack([h(np.eye(dense.shape[0])[i]) for i in range(dense.shape[0])], axis=1)
        assert abs(np.linalg.eigvalsh(0.5 * (applied + applied.T))[0] - w) < 1e-7 * max(1, abs(w))  # rtol of the upstream test


def test_sigma_nonsymmetric_integrals():
    """direct_nosym semantics: no permutational symmetry of int2e is assumed (ucc.py:70-71)."""
    rng = np.random.default_rng(3)
    n, nelec = 4, 4
    int1e = rng.standard_normal((n, n))
    int2e = rng.standard_normal((n, n, n, n))
    h_ref = O.get_h_fcifunc_from_integral(int1e, int2e, nelec)
    h = get_h_from_integral(int1e, int2e, nelec, False)
    x = rng.standard_normal(36)
    assert np.abs(h(x) - h_ref(x)).max() < 1e-11


def test_every_excitation_kind_single_rotation():
    """One rotation of each kind (alpha single, beta single, alpha-alpha, beta-beta, alpha-beta, mixed order)
    on a random vector against the oracle's table form."""
    n, nelec = 5, 4
    ops = [(7, 5), (2, 0), (8, 7, 6, 5), (3, 2, 1, 0), (2, 7, 5, 0), (7, 2, 0, 5), (5, 0, 3, 7), (9, 3, 0, 6), (6, 9, 8, 5)]
    rng = np.random.default_rng(5)
    x = rng.standard_normal(100)
    for op in ops:
        plan = Plan(2 * n, nelec, [op], [0])
        for theta in (0.3, -1.2):
            c = plan.civector([theta], init_state=x)
            c_ref = O.get_civector(np.array([theta]), 2 * n, nelec, [op], [0], init_state=x)
            assert np.abs(c - c_ref).max() < 1e-14, op
        g = engine_ucc.apply_excitation(x, 2 * n, nelec, op, hcb=False)
        assert np.abs(g - O.apply_excitation(x, 2 * n, nelec, op, False)).max() == 0.0, op
    O.clear_cache()


def test_reference_test_excitation_case():
    """tests/static/test_engine.py:44-71: civector(params) then apply_excitation (4,0,3,7) on an H4-sized space."""
    ucc, int1e, int2e = _build("uccsd", 2, 2)
    params = _params(ucc.n_params)
    state = ucc.apply_excitation(ucc.civector(params), (4, 0, 3, 7))
    ref = O.apply_excitation(O.get_civector(params, 8, 4, ucc.ex_ops, ucc.param_ids), 8, 4, (4, 0, 3, 7), False)
    assert np.abs(state - ref).max() < 1e-13
    dense = bf.dense_excitation_generator((4, 0, 3, 7), 8, 4) @ bf.ucc_state(params, ucc.ex_ops, ucc.param_ids, 8, 4)
    assert np.abs(state - dense).max() < 1e-12


def test_init_state_and_input_not_modified():
    ucc, int1e, int2e = _build("uccsd", 2, 2)
    params = _params(ucc.n_params)
    rng = np.random.default_rng(0)
    init = rng.standard_normal(36)
    init /= np.linalg.norm(init)
    keep = init.copy()
    ucc.init_state = init
    e, g = ucc.energy_and_grad(params)
    h = O.get_h_fcifunc_from_integral(int1e, int2e, 4)
    e_ref, g_ref = O.get_energy_and_grad(params, h, 8, 4, ucc.ex_ops, ucc.param_ids, init_state=init)
    assert abs(e - e_ref) < E_TOL and np.abs(g - g_ref).max() < G_TOL
    np.testing.assert_array_equal(init, keep)  # the reference aliases and mutates; we must not
    # statevector-form init state (engine_ucc.py:163-174)
    sv = np.zeros(256)
    sv[O.get_ci_strings(8, 4, False)] = init
    ucc.init_state = sv
    e2, _ = ucc.energy_and_grad(params)
    assert abs(e2 - e_ref) < E_TOL


def test_error_behaviour():
    ucc, _, _ = _build("uccsd", 2, 2)
    with pytest.raises(ValueError, match="Incompatible parameter shape"):
        ucc.energy(np.zeros(3))
    with pytest.raises(ValueError, match="not supported"):
        ucc.energy(np.zeros(ucc.n_params), engine="tensornetwork")
    with pytest.raises(ValueError, match="not supported"):
        engine_ucc.apply_excitation(np.zeros(36), 8, 4, (5, 0), hcb=False)
    with pytest.raises(ValueError, match="not supported"):
        engine_ucc.get_energy_and_grad(np.zeros(1), ucc.hamiltonian, 8, 4, [(1, 1)], [0], False, None)
    empty = UCC.from_integral(*random_integral(3), 2, ex_ops=[], param_ids=[])
    assert np.array_equal(empty.civector(np.zeros(0)), np.eye(1, 9)[0])  # no excitations: HF determinant
    e, g = empty.energy_and_grad(np.zeros(0))
    assert g.shape == (0,)


def test_clear_cache_does_not_change_energy():
    """tests/test_misc.py:8-13."""
    ucc, _, _ = _build("uccsd", 2, 2)
    params = _params(ucc.n_params)
    e1 = ucc.energy(params)
    clear_cache()
    e2 = ucc.energy(params)
    assert abs(e1 - e2) < 1e-12  # sigma accumulates with fp64 atomics: last-bit differences run to run


def test_kernel_reaches_fci():
    """test_gradient_opt / test_ucc upstream: UCCSD reaches FCI for a 4-orbital 4-electron system (atol 1e-4 .. 3e-3)."""
    int1e, int2e = random_integral(4)
    # weakly correlated: scale the two-body part so that index 0 is the dominant determinant
    int1e = int1e + np.diag([-4.0, -3.0, 3.0, 4.0])
    ucc = UCCSD.from_integral(int1e, int2e, 4)
    e = ucc.kernel()
    e_fci = np.linalg.eigvalsh(bf.dense_hamiltonian(int1e, int2e, 4))[0]
    assert e >= e_fci - 1e-9
    assert abs(e - e_fci) < 3e-3
    assert abs(ucc.energy() - e) < 1e-12


def test_puccd_matches_explicit_pair_ucc():
    """tests/static/test_puccd.py:19-58: pUCCD (hcb) == fermionic UCC with explicit paired excitations."""
    no, nv = 2, 2
    int1e, int2e = random_integral(4)
    pu = PUCCD.from_integral(int1e, int2e, 4)
    params = _params(pu.n_params)
    paired = [(no + a, 4 + no + a, 4 + i, i) for i in range(no) for a in range(nv - 1, -1, -1)]
    assert paired[0] == (3, 7, 4, 0)
    ferm = UCC.from_integral(int1e, int2e, 4, ex_ops=paired, param_ids=list(range(len(paired))))
    assert abs(pu.energy(params) - ferm.energy(params)) < 1e-10


def test_properties_medium_size():
    """(10e,10o)-size properties the oracle is too slow to check directly in full: norm conservation,
    <x|H y> symmetry, energy invariance under the reverse sweep bookkeeping (grad vs finite differences)."""
    n = 8
    int1e, int2e = random_integral(n)
    ucc = UCCSD.from_integral(int1e, int2e, n)
    params = 0.1 * _params(ucc.n_params)
    c = ucc.civector(params)
    assert abs(np.linalg.norm(c) - 1.0) < 1e-12
    rng = np.random.default_rng(2)
    x, y = rng.standard_normal(c.size), rng.standard_normal(c.size)
    hx, hy = ucc.hamiltonian(x), ucc.hamiltonian(y)
    assert abs(y @ hx - x @ hy) < 1e-8 * abs(y @ hx)
    e, g = ucc.energy_and_grad(params)
    assert abs(e - c @ ucc.hamiltonian(c)) < 1e-10
    eps = 1e-5
    for i in (0, 17, ucc.n_params - 1):
        d = np.zeros_like(params)
        d[i] = eps
        fd = (ucc.energy(params + d) - ucc.energy(params - d)) / (2 * eps)
        assert abs(fd - g[i]) < 1e-7
    # oracle on the same inputs (4,900 determinants: seconds on the CPU)
    h = O.get_h_fcifunc_from_integral(int1e, int2e, n)
    e_ref, g_ref = O.get_energy_and_grad(params, h, 2 * n, n, ucc.ex_ops, ucc.param_ids)
    assert abs(e - e_ref) < E_TOL and np.abs(g - g_ref).max() < G_TOL
    O.clear_cache()


@pytest.mark.parametrize("tile_elems,threads", [(64, 32), (400, 64), (1500, 128), (1500, 256)])
def test_small_tiles_many_batches(tile_elems, threads, monkeypatch):
    """Force the batch planner into many batches / many tiles per batch (what large active spaces look like)
    on a size the oracle can check: state, energy AND gradient (tile-partial reduction, spectator signs)."""
    monkeypatch.setenv("TCC_B200_TILE_ELEMS", str(tile_elems))
    monkeypatch.setenv("TCC_B200_BATCH_THREADS", str(threads))
    clear_cache()
    for kind, no, nv in [("uccsd", 4, 4), ("uccsd", 3, 4), ("kupccgsd", 3, 3), ("puccd", 4, 4)]:
        ucc, int1e, int2e = _build(kind, no, nv)
        params = _params(ucc.n_params)
        plan = Plan(ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, hcb=ucc.hcb)
        plan.set_hamiltonian(int1e, int2e)
        if kind == "uccsd" and no == 4 and tile_elems <= 400:
            assert plan.n_batches > 2
        h = O.get_h_from_integral(int1e, int2e, ucc.n_elec, ucc.hcb)
        e_ref, g_ref = O.get_energy_and_grad(params, h, ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, ucc.hcb)
        c_ref = O.get_civector(params, ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, ucc.hcb)
        e, g = plan.energy_and_grad(params)
        assert np.abs(plan.civector(params) - c_ref).max() < C_TOL
        assert abs(e - e_ref) < E_TOL
        assert np.abs(g - g_ref).max() < G_TOL
        O.clear_cache()
    clear_cache()


@pytest.mark.parametrize("persist,persist_fwd,sched", [("2", "1", "2"), ("0", "0", "0"), ("2", "0", "1"), ("1", "0", "0")])
@pytest.mark.parametrize("tile_elems", [64, 400, 6144])
def test_persistent_kernels_and_schedulers_against_oracle(tile_elems, persist, persist_fwd, sched, monkeypatch):
    """The persistent batch kernel (lean_persist_kernel: ticketed tiles, next tile resolved during the gather) in the
    reverse sweep (default) AND in the forward sweep (TCC_B200_PERSIST_FWD=1), the classic one-tile-per-CTA kernels
    (TCC_B200_PERSIST=0; =2 forces the persistent kernel also for launches with few tiles, =1 is the default: from
    four tiles per resident CTA on), and the three batch schedulers (TCC_B200_SCHED), from a handful to thousands of tiles per
    launch (more tiles than resident CTAs, so every CTA takes several tickets): state, energy, gradient, with an
    init_state, twice in a row (the ticket counters reset themselves between launches)."""
    monkeypatch.setenv("TCC_B200_TILE_ELEMS", str(tile_elems))
    monkeypatch.setenv("TCC_B200_PERSIST", persist)
    monkeypatch.setenv("TCC_B200_PERSIST_FWD", persist_fwd)
    monkeypatch.setenv("TCC_B200_SCHED", sched)
    clear_cache()
    for kind, no, nv in [("uccsd", 4, 4), ("uccsd", 2, 4), ("kupccgsd", 3, 3)]:
        ucc, int1e, int2e = _build(kind, no, nv)
        params = _params(ucc.n_params)
        rng = np.random.default_rng(11)
        init = rng.standard_normal(ucc.civector_size)
        init /= np.linalg.norm(init)
        plan = Plan(ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids)
        plan.set_hamiltonian(int1e, int2e)
        h = O.get_h_from_integral(int1e, int2e, ucc.n_elec, False)
        e_ref, g_ref = O.get_energy_and_grad(params, h, ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, False)
        ei_ref, gi_ref = O.get_energy_and_grad(params, h, ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, False, init_state=init)
        c_ref = O.get_civector(params, ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, False)
        for _ in range(2):
            e, g = plan.energy_and_grad(params)
            assert abs(e - e_ref) < E_TOL and np.abs(g - g_ref).max() < G_TOL
            ei, gi = plan.energy_and_grad(params, init)
            assert abs(ei - ei_ref) < E_TOL and np.abs(gi - gi_ref).max() < G_TOL
            assert np.abs(plan.civector(params) - c_ref).max() < C_TOL
        O.clear_cache()
    clear_cache()


@pytest.mark.parametrize("tile_elems", [64, 400, 1500])
def test_dynamic_column_layout_against_static_and_oracle(tile_elems, monkeypatch):
    """Dynamic column layout (every batch gathers its tiles from contiguous columns and scatters in the order of the
    next batch, out of place; plan.hpp) against the in-place static order and the oracle: state, energy, gradient, with
    and without an init_state, odd and even numbers of batches."""
    monkeypatch.setenv("TCC_B200_TILE_ELEMS", str(tile_elems))
    seen_batches = set()
    for kind, no, nv in [("uccsd", 4, 4), ("uccsd", 3, 4), ("uccsd", 2, 4), ("kupccgsd", 3, 3)]:
        ucc, int1e, int2e = _build(kind, no, nv)
        params = _params(ucc.n_params)
        rng = np.random.default_rng(7)
        init = rng.standard_normal(ucc.civector_size)
        init /= np.linalg.norm(init)
        res = {}
        for dyn in ("1", "0"):
            monkeypatch.setenv("TCC_B200_DYN_LAYOUT", dyn)
            clear_cache()
            plan = Plan(ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids)
            plan.set_hamiltonian(int1e, int2e)
            seen_batches.add(plan.n_batches)
            e, g = plan.energy_and_grad(params)
            c = plan.civector(params)
            ci = plan.civector(params, init)
            ei, gi = plan.energy_and_grad(params, init)
            res[dyn] = (e, g, c, ci, ei, gi)
        for x, y in zip(res["1"], res["0"]):
            assert np.abs(np.asarray(x) - np.asarray(y)).max() < 1e-13
        h = O.get_h_from_integral(int1e, int2e, ucc.n_elec, False)
        e_ref, g_ref = O.get_energy_and_grad(params, h, ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, False)
        c_ref = O.get_civector(params, ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, False)
        ci_ref = O.get_civector(params, ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, False, init_state=init)
        e, g, c, ci, ei, gi = res["1"]
        assert abs(e - e_ref) < E_TOL and np.abs(g - g_ref).max() < G_TOL
        assert np.abs(c - c_ref).max() < C_TOL and np.abs(ci - ci_ref).max() < C_TOL
        O.clear_cache()
    assert any(b > 1 for b in seen_batches)
    clear_cache()


def test_batch_schedule_is_a_valid_reordering():
    """Every excitation appears exactly once, and two excitations whose relative order changed act on disjoint
    spin-orbitals (so their generators commute and the product of factors is unchanged)."""
    ucc, _, _ = _build("uccsd", 4, 4)
    for tile in (64, 400, 6144):
        os.environ["TCC_B200_TILE_ELEMS"] = str(tile)
        try:
            plan = Plan(ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, device=None)
        finally:
            del os.environ["TCC_B200_TILE_ELEMS"]
        order = [j for b in range(plan.n_batches) for j in plan.batch_info(b)["op_indices"]]
        assert sorted(order) == list(range(len(ucc.ex_ops)))
        pos = {j: i for i, j in enumerate(order)}
        sets = [set(op) for op in ucc.ex_ops]
        for a in range(len(order)):
            for b in range(a + 1, len(order)):
                if pos[a] > pos[b]:
                    assert not (sets[a] & sets[b]), (ucc.ex_ops[a], ucc.ex_ops[b])


@pytest.mark.parametrize("kind,no,nv", [("uccsd", 2, 2), ("uccsd", 3, 3), ("kupccgsd", 3, 2), ("puccd", 3, 3), ("uccsd", 4, 4)])
def test_batched_parameter_sets(kind, no, nv):
    """energy_and_grad_batch (many theta in the same launches) == one call per theta == oracle."""
    ucc, int1e, int2e = _build(kind, no, nv)
    plan = Plan(ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, hcb=ucc.hcb)
    plan.set_hamiltonian(int1e, int2e)
    rng = np.random.default_rng(11)
    thetas = rng.random((7, ucc.n_params)) - 0.5
    e_b, g_b = plan.energy_and_grad_batch(thetas)
    e_only, none = plan.energy_and_grad_batch(thetas, with_grad=False)
    assert none is None
    h = O.get_h_from_integral(int1e, int2e, ucc.n_elec, ucc.hcb)
    for i, th in enumerate(thetas):
        e1, g1 = plan.energy_and_grad(th)
        assert abs(e_b[i] - e1) < 1e-12 and np.abs(g_b[i] - g1).max() < 1e-12
        assert abs(e_only[i] - e1) < 1e-12
        if i < 2:
            e_ref, g_ref = O.get_energy_and_grad(th, h, ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, ucc.hcb)
            assert abs(e_b[i] - e_ref) < E_TOL and np.abs(g_b[i] - g_ref).max() < G_TOL
    O.clear_cache()
    with pytest.raises(ValueError, match="Incompatible parameter shape"):
        plan.energy_and_grad_batch(np.zeros((3, max(ucc.n_params - 1, 0))))


def test_batched_parameter_sets_through_the_ansatz_class_and_rank_sharding():
    from tencirchem_b200.parallel import energy_and_grad_batch, shard_bounds

    ucc, int1e, int2e = _build("kupccgsd", 2, 2)
    rng = np.random.default_rng(3)
    thetas = rng.random((9, ucc.n_params)) - 0.5
    e, g = ucc.energy_and_grad_batch(thetas)
    for i in (0, 8):
        e1, g1 = ucc.energy_and_grad(thetas[i])
        assert abs(e[i] - e1) < 1e-12 and np.abs(g[i] - g1).max() < 1e-12
    with pytest.raises(ValueError, match="Incompatible parameter shape"):
        ucc.energy_and_grad_batch(thetas[:, :-1])
    # the per-rank slice of a 2-rank job evaluates exactly its rows (world == 1 path of each rank)
    plan = Plan(ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids)
    plan.set_hamiltonian(int1e, int2e)
    lo, hi = shard_bounds(len(thetas), 1, 2)
    e_r, g_r = energy_and_grad_batch(plan, thetas[lo:hi])
    np.testing.assert_allclose(e_r, e[lo:hi], atol=1e-12)
    np.testing.assert_allclose(g_r, g[lo:hi], atol=1e-12)


def test_batched_parameter_sets_chunking(monkeypatch):
    """More sets than fit the memory budget at once are processed in chunks."""
    monkeypatch.setenv("TCC_B200_BATCH_MEM_MB", "1")
    ucc, int1e, int2e = _build("uccsd", 3, 3)
    plan = Plan(ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids)
    plan.set_hamiltonian(int1e, int2e)
    rng = np.random.default_rng(12)
    thetas = rng.random((50, ucc.n_params)) - 0.5
    e_b, g_b = plan.energy_and_grad_batch(thetas)
    for i in (0, 17, 49):
        e1, g1 = plan.energy_and_grad(thetas[i])
        assert abs(e_b[i] - e1) < 1e-12 and np.abs(g_b[i] - g1).max() < 1e-12


@pytest.mark.parametrize("kind,no,nv", [("uccsd", 1, 4), ("uccsd", 4, 1), ("uccsd", 1, 2), ("kupccgsd", 1, 3), ("puccd", 1, 4), ("puccd", 4, 1)])
def test_ragged_active_spaces(kind, no, nv):
    """Very unequal occupied / virtual counts (single electron pair, single virtual orbital)."""
    ucc, int1e, int2e = _build(kind, no, nv)
    params = _params(ucc.n_params)
    h = O.get_h_from_integral(int1e, int2e, ucc.n_elec, ucc.hcb)
    e_ref, g_ref = O.get_energy_and_grad(params, h, ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, ucc.hcb)
    e, g = ucc.energy_and_grad(params)
    assert abs(e - e_ref) < E_TOL and np.abs(g - g_ref).max() < G_TOL
    c_ref = O.get_civector(params, ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, ucc.hcb)
    assert np.abs(ucc.civector(params) - c_ref).max() < C_TOL
    O.clear_cache()


def test_trivial_spaces():
    """No electrons / no virtual orbitals: a one-determinant CI space and an empty excitation list."""
    int1e, int2e = random_integral(3)
    for n_elec in (0, 6):
        ucc = UCC.from_integral(int1e, int2e, n_elec, ex_ops=[], param_ids=[])
        assert ucc.civector_size == 1
        np.testing.assert_array_equal(ucc.civector(np.zeros(0)), [1.0])
        h = O.get_h_fcifunc_from_integral(int1e, int2e, n_elec)
        assert abs(ucc.energy(np.zeros(0)) - h(np.ones(1))[0]) < 1e-12


def test_properties_full_size_16o(monkeypatch):
    """BASELINE.json's (16e,16o) UCCSD at full size (1.66e8 determinants, 5792 excitations): the oracle would need
    a day, so parity is checked through size-independent properties -- norm conservation, symmetry of H,
    E == <c|H|c>, gradient vs central finite differences, and agreement of the two independent GPU code paths
    (batched tiles vs one launch per excitation)."""
    import torch

    n = 16
    int1e, int2e = random_integral(n)
    ucc = UCCSD.from_integral(int1e, int2e, n)
    params = 0.05 * _params(ucc.n_params)
    clear_cache()
    plan = Plan(ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids)
    plan.set_hamiltonian(int1e, int2e)
    assert plan.size == 165636900 and plan.n_ops == 5792 and plan.n_batches > 100
    c = torch.empty(plan.size, dtype=torch.float64, device="cuda")
    plan.civector_dev(params, c)
    assert abs(float(torch.linalg.vector_norm(c)) - 1.0) < 1e-11
    hc = torch.empty_like(c)
    plan.apply_h_dev(c, hc)
    e_direct = float(torch.dot(c, hc))
    e, g = plan.energy_and_grad_dev(params)
    assert abs(e - e_direct) < 1e-9
    # H symmetric: <x|H c> == <c|H x> for a random x (reuse hc's storage for H x afterwards)
    x = torch.randn(plan.size, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    x_hc = float(torch.dot(x, hc))
    plan.apply_h_dev(x, hc)
    c_hx = float(torch.dot(c, hc))
    assert abs(x_hc - c_hx) < 1e-8 * max(1.0, abs(x_hc))
    del x, hc, c
    torch.cuda.empty_cache()
    # gradient of two parameters (one single, one opposite-spin double) against finite differences of the energy:
    # central differences at two step sizes, Richardson-extrapolated (truncation error O(eps^4); what is left is the
    # rounding of the energies, ~1e-13 / eps).  The north_star's 1e-8 on the gradient itself is checked against the
    # sparse oracle at this size in tests/test_gpu_large_parity.py; this is the oracle-free cross-check.
    eps = 2e-3
    for i in (3, ucc.n_params - 5):
        d = np.zeros_like(params)
        d[i] = eps
        fd1 = (plan.energy(params + d) - plan.energy(params - d)) / (2 * eps)
        fd2 = (plan.energy(params + 0.5 * d) - plan.energy(params - 0.5 * d)) / eps
        fd = (4.0 * fd2 - fd1) / 3.0
        assert abs(fd - g[i]) < 2e-8, (i, fd1, fd2, fd, g[i])
    del plan
    # the per-excitation kernels are an independent implementation of the same sweeps
    monkeypatch.setenv("TCC_B200_PER_OP", "1")
    plan2 = Plan(ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids)
    plan2.set_hamiltonian(int1e, int2e)
    e2, g2 = plan2.energy_and_grad_dev(params)
    assert abs(e - e2) < E_TOL
    assert np.abs(g - g2).max() < G_TOL
    del plan2
    torch.cuda.empty_cache()


def test_h_chain_published_energies_on_the_engine():
    """Literal BASELINE molecules (PySCF-free STO-3G integrals, tencirchem_b200/hchain.py) through the engine against
    the numbers the reference publishes: H8 (2e,2o) active space (ucc_functions.ipynb cells 31-42)."""
    from tencirchem_b200.hchain import active_space_integrals, h_chain_integrals

    k = KAT["h_chain_sto3g"]
    int1e, int2e, e_nuc, e_hf = h_chain_integrals(8)
    h1, h2, e_core = active_space_integrals(int1e, int2e, e_nuc, 8, (2, 2))
    ucc = UCCSD.from_integral(h1, h2, 2, e_core)
    assert abs(ucc.energy(np.zeros(2)) - k["h8_e_hf"]) < 1e-11
    e, g = ucc.energy_and_grad(np.zeros(2))
    np.testing.assert_allclose(g, k["h8_cas22_grad_at_zero"], atol=3e-7)
    assert abs(ucc.kernel() - k["h8_cas22_uccsd"]) < 1e-7


def test_screened_uccsd_reproduces_the_published_hydrogen_chain_table():
    """The reference's headline UCC table (docs/source/quickstart.rst:75-104): MP2-screened, amplitude-sorted UCCSD on
    H4 / H6 / H10 chains (STO-3G, 0.8 A) optimised with kernel() on the engine; the published columns "Parameters" and
    "Error (mH)" = E_UCCSD - E_FCI, with E_FCI from the engine's own H.c (Lanczos, tencirchem_b200/fci.py), itself
    checked against the dense oracle Hamiltonian."""
    from tencirchem_b200.hchain import h_chain_integrals, mp2_t2

    published = {4: (11, 0.01), 6: (39, 0.27), 10: (246, 1.37)}  # parameters, error in mH
    for n, (n_params, err_mh) in published.items():
        int1e, int2e, e_nuc, e_hf, e_orb = h_chain_integrals(n, 0.8, True)
        _, t2 = mp2_t2(int2e, e_orb, n // 2)
        ucc = UCCSD.from_integral(int1e, int2e, n, e_nuc, t2=t2)
        assert ucc.n_params == n_params
        assert abs(ucc.energy(np.zeros(ucc.n_params)) - e_hf) < 1e-10
        e = ucc.kernel()
        if n <= 6:
            assert abs(ucc.e_fci - (np.linalg.eigvalsh(bf.dense_hamiltonian(int1e, int2e, n))[0] + e_nuc)) < 1e-9
        assert round(1e3 * (e - ucc.e_fci), 2) == err_mh
    clear_cache()


def test_adapt_vqe_tutorial_h4_on_the_engine():
    """The reference's ADAPT-VQE tutorial (adapt_vqe.ipynb cells 3-11) replayed on the B200 engine for the literal H4
    chain: pool gradients through civector / hamiltonian / apply_excitation, operator picks and the optimised
    energy of every iteration against the published output."""
    from collections import defaultdict

    from tencirchem_b200.hchain import h_chain_integrals

    k = KAT["h_chain_sto3g"]
    int1e, int2e, e_nuc, _ = h_chain_integrals(4)
    probe = UCCSD.from_integral(int1e, int2e, 4, e_nuc)
    ex1_ops, ex1_ids, _ = probe.get_ex1_ops()
    ex2_ops, ex2_ids, _ = probe.get_ex2_ops()
    pool = defaultdict(list)
    for op, pid in zip(ex1_ops, ex1_ids):
        pool[(1, pid)].append(op)
    for op, pid in zip(ex2_ops, ex2_ids):
        pool[(2, pid)].append(op)
    pool = list(pool.values())
    ucc = UCC.from_integral(int1e, int2e, 4, e_nuc, ex_ops=[], param_ids=[])
    ucc.params = []
    for it, (pick_pub, e_pub) in enumerate(zip(k["h4_adapt_picks"], k["h4_adapt_energies"])):
        psi = ucc.civector()
        bra = ucc.hamiltonian(psi)
        grads = [2 * sum(bra @ ucc.apply_excitation(psi, op) for op in ops_) for ops_ in pool]
        # the whole pool in one batched launch set (tcc_pool_gradients), psi and H psi kept on the GPU
        np.testing.assert_allclose(ucc.pool_gradients(pool), grads, atol=1e-12)
        if it == 0:
            assert abs(np.linalg.norm(grads) - k["h4_adapt_iter0_norm"]) < 2e-7
            pub = {str([list(o) for o in ops_]): g for ops_, g in k["h4_adapt_iter0"]}
            for ops_, g in zip(pool, grads):
                key = str([list(o) for o in ops_])
                assert abs(g - pub.get(key, 0.0)) < 3e-7, (ops_, g)
        chosen = pool[int(np.argmax(np.abs(grads)))]
        assert [list(o) for o in chosen] == pick_pub, (it, chosen, pick_pub)
        ucc.ex_ops.extend(chosen)
        ucc.params = list(ucc.params) + [0]
        ucc.param_ids = list(ucc.param_ids) + [len(ucc.params) - 1] * len(chosen)
        ucc.init_guess = ucc.params
        ucc.kernel()
        assert abs(ucc.e_ucc - e_pub) < 2e-6, (it, ucc.e_ucc, e_pub)


@pytest.mark.parametrize("kind,no,nv", [("uccsd", 3, 3), ("uccsd", 2, 4), ("puccd", 3, 3)])
def test_pool_gradients_against_oracle(kind, no, nv):
    """tcc_pool_gradients: 2 <H psi| G_k |psi> for every operator of a pool (all UCCSD excitations, or pUCCD pair
    excitations on the hard-core-boson space) at a non-trivial state, against the oracle's apply_excitation
    (evolve_civector.py:311-313) and H psi."""
    ucc, int1e, int2e = _build(kind, no, nv)
    params = _params(ucc.n_params)
    c = O.get_civector(params, ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, ucc.hcb)
    hc = O.get_h_from_integral(int1e, int2e, ucc.n_elec, ucc.hcb)(c)
    pool = [[op] for op in ucc.ex_ops]
    ref = np.array([2 * hc @ O.apply_excitation(c, ucc.n_qubits, ucc.n_elec, op, ucc.hcb) for (op,) in pool])
    got = ucc.pool_gradients(pool, params)
    assert np.abs(got - ref).max() < G_TOL
    # grouped entries (operators sharing a parameter) add up
    pairs = [list(ucc.ex_ops[i:i + 2]) for i in range(0, len(ucc.ex_ops) - 1, 2)]
    got2 = ucc.pool_gradients(pairs, params)
    assert np.abs(got2 - np.array([ref[i] + ref[i + 1] for i in range(0, len(ucc.ex_ops) - 1, 2)])).max() < G_TOL
    O.clear_cache()


@pytest.mark.parametrize("no,nv", [(2, 2), (3, 3), (2, 5)])
def test_puccd_hcb_rdms_on_gpu(no, nv):
    """tcc_make_rdm12_hcb + PUCCD.make_rdm1 / make_rdm2 against the restated puccd.py:148-203, on an optimiser-free
    random-theta pUCCD state and on a random hcb vector; and, as in the reference's own test
    (tests/static/test_puccd.py:48-57), against the fermionic UCC RDMs of the same state built from paired excitations."""
    from oracle import rdm_numpy as R

    ucc, int1e, int2e = _build("puccd", no, nv)
    n, ne = no + nv, 2 * no
    params = _params(ucc.n_params)
    c = ucc.civector(params)
    r1, r2 = R.make_rdm12_hcb(c, n, ne)
    assert np.abs(ucc.make_rdm1(c) - r1).max() < 1e-13
    assert np.abs(ucc.make_rdm2(c) - r2).max() < 1e-13
    e = np.einsum("pq,pq", int1e, r1) + 0.5 * np.einsum("pqrs,pqrs", int2e, ucc.make_rdm2(c))
    assert abs(e - ucc.energy(params)) < E_TOL
    ucc.params = params
    assert np.abs(ucc.make_rdm2() - r2).max() < 1e-13  # default: the state at self.params
    v = np.random.default_rng(3).standard_normal(len(c))
    a1, a2 = R.make_rdm12_hcb(v, n, ne)
    assert np.abs(ucc.make_rdm1(v) - a1).max() < 1e-13 * np.abs(a1).max()
    assert np.abs(ucc.make_rdm2(v) - a2).max() < 1e-13 * np.abs(a2).max()
    # the same state as a fermionic UCC with paired excitations (b+_i = a+_(n+i) a+_i)
    from tencirchem_b200 import UCC

    fer = UCC.from_integral(int1e, int2e, ne)
    fer.ex_ops = tuple((n + a, a, i, n + i) for (a, i) in ucc.ex_ops)
    fer.param_ids = list(range(len(fer.ex_ops)))
    fer.params = params
    assert abs(fer.energy(params) - ucc.energy(params)) < E_TOL
    assert np.abs(fer.make_rdm1() - ucc.make_rdm1()).max() < 1e-12
    assert np.abs(fer.make_rdm2() - ucc.make_rdm2()).max() < 1e-12
    with pytest.raises(ValueError):
        ucc.make_rdm1(c, basis="AO")
    O.clear_cache()


def test_device_resident_api():
    import torch

    ucc, int1e, int2e = _build("uccsd", 3, 3)
    params = _params(ucc.n_params)
    plan = Plan(ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids)
    plan.set_hamiltonian(int1e, int2e)
    out = torch.empty(plan.size, dtype=torch.float64, device="cuda")
    plan.civector_dev(params, out)
    sigma = torch.empty_like(out)
    plan.apply_h_dev(out, sigma)
    e_dev = float(out @ sigma)
    e, g = plan.energy_and_grad_dev(params)
    assert abs(e - e_dev) < 1e-11
    c_ref = O.get_civector(params, ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids)
    assert np.abs(out.cpu().numpy() - c_ref).max() < C_TOL
    assert plan.launch_count > 0
    O.clear_cache()


# ---- alpha-row sharded engine, all ranks as virtual ranks on one GPU ----------------------
@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("tile_elems", [400, 6144])
def test_sharded_virtual_ranks_match_oracle(world, tile_elems, monkeypatch):
    """The sharded driver with every rank in this process (LocalComm): each rank's plan holds only its alpha rows,
    rows move between layouts, sigma is summed over row blocks, gradients over ranks.  State, energy and gradient
    against the oracle."""
    from tencirchem_b200.sharded import LocalComm, ShardedUCC

    monkeypatch.setenv("TCC_B200_TILE_ELEMS", str(tile_elems))
    for kind, no, nv in [("uccsd", 3, 3), ("uccsd", 4, 4), ("kupccgsd", 3, 3)]:
        ucc, int1e, int2e = _build(kind, no, nv)
        params = _params(ucc.n_params)
        h = O.get_h_from_integral(int1e, int2e, ucc.n_elec, False)
        e_ref, g_ref = O.get_energy_and_grad(params, h, ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, False)
        c_ref = O.get_civector(params, ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, False)
        # H c either with the vector kept sharded (three alpha-beta passes + column transpose) or all-reduced
        for sigma_mode in (("distributed", "replicated") if world > 1 else ("replicated",)):
            # a small staging budget forces the row moves and the column transposes to run in several rounds
            eng = ShardedUCC(ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, LocalComm(world), sigma_mode=sigma_mode,
                             chunk_bytes=(4 << 30) if tile_elems == 6144 else 1024,
                             fuse_reverse_moves=(sigma_mode == "replicated"))
            eng.set_hamiltonian(int1e, int2e)
            assert np.abs(eng.civector(params) - c_ref).max() < C_TOL
            e, g = eng.energy_and_grad(params)
            assert abs(e - e_ref) < E_TOL, (sigma_mode, e, e_ref)
            assert np.abs(g - g_ref).max() < G_TOL, sigma_mode
            if world > 1 and kind == "uccsd":
                assert eng.reshards > 0  # the layouts really changed
        O.clear_cache()


def test_sigma_by_row_blocks_sums_to_h_c():
    """tcc_apply_h_rows over a partition of the rows == tcc_apply_h (both through the C ABI, storage order)."""
    import torch

    n = 5
    int1e, int2e = random_integral(n)
    plan = Plan(2 * n, n + 1 if (n + 1) % 2 == 0 else n - 1, (), ())
    plan.set_hamiltonian(int1e, int2e)
    na = plan.n_strings
    rng = np.random.default_rng(5)
    c_ref = torch.as_tensor(rng.standard_normal(plan.size), device="cuda")
    c_sto = torch.empty_like(c_ref)
    plan.to_storage_order_dev(c_ref, c_sto)
    sig_ref = torch.empty_like(c_ref)
    plan.apply_h_dev(c_ref, sig_ref)
    sig_sto = torch.zeros_like(c_ref)
    for lo, hi in [(0, 3), (3, na // 2), (na // 2, na)]:
        plan.apply_h_rows(c_sto, sig_sto, lo, hi - lo)
    out = torch.empty_like(c_ref)
    plan.from_storage_order_dev(sig_sto, out)
    torch.cuda.synchronize()
    assert (out - sig_ref).abs().max().item() < 1e-12


def test_ucc_class_on_sharded_engine_kernel():
    """UCCSD.shard(comm): the reference-shaped class API (energy, energy_and_grad, kernel) on the sharded engine gives
    the same optimum as the single-GPU engine."""
    from tencirchem_b200.sharded import LocalComm

    ucc1, int1e, int2e = _build("uccsd", 3, 3)
    ucc4 = UCCSD.from_integral(int1e, int2e, 6).shard(LocalComm(4))
    params = _params(ucc1.n_params)
    ea, ga = ucc1.energy_and_grad(params)
    eb, gb = ucc4.energy_and_grad(params)
    assert abs(ea - eb) < E_TOL and np.abs(ga - gb).max() < G_TOL, (ea, eb)
    assert abs(ucc1.energy(params) - ucc4.energy(params)) < E_TOL
    assert np.abs(ucc1.civector(params) - ucc4.civector(params)).max() < C_TOL
    # L-BFGS-B driven by the sharded engine: it converges, and the single-GPU engine agrees at the optimum (two
    # optimiser runs are not compared with each other: rounding-level differences may end in different local minima)
    e_start = ucc4.energy(np.zeros(ucc4.n_params))
    e4 = ucc4.kernel()
    assert e4 < e_start
    e_opt, g_opt = ucc1.energy_and_grad(ucc4.params)
    assert abs(e_opt - e4) < E_TOL
    assert np.abs(g_opt).max() < 1e-3
    with pytest.raises(ValueError):
        PUCCD.from_integral(int1e, int2e, 6).shard(LocalComm(2))


@pytest.mark.parametrize("no,nv", [(1, 1), (2, 2), (2, 3), (3, 3), (4, 5)])
def test_rdm12_on_gpu_against_oracle(no, nv):
    """tcc_make_rdm12 (Gram matrix of the single-replacement intermediate on the FP64 tensor path) against the CPU
    restatement of PySCF's conventions, plus the reference's own identity E = h.rdm1 + 1/2 eri.rdm2 (ucc.py:838-846)
    through the class API.  (4e... ,9 orbitals) has n^2 = 81 > 64: two row blocks of the Gram matrix."""
    from oracle import rdm_numpy as R

    ucc, int1e, int2e = _build("uccsd", no, nv)
    params = _params(ucc.n_params)
    c = ucc.civector(params)
    dm1_ref, dm2_ref = R.make_rdm12(c, no + nv, 2 * no)
    dm1 = ucc.make_rdm1(c)
    dm2 = ucc.make_rdm2(c)
    assert np.abs(dm1 - dm1_ref).max() < 1e-12
    assert np.abs(dm2 - dm2_ref).max() < 1e-12
    e = np.einsum("pq,pq", int1e, dm1) + 0.5 * np.einsum("pqrs,pqrs", int2e, dm2)
    assert abs(e - ucc.energy(params)) < E_TOL
    # a non-normalised, non-symmetric random vector too
    rng = np.random.default_rng(3)
    v = rng.standard_normal(ucc.civector_size)
    a1, a2 = R.make_rdm12(v, no + nv, 2 * no)
    assert np.abs(ucc.make_rdm1(v) - a1).max() < 1e-13 * np.abs(a1).max()  # |v|^2 ~ N: relative tolerance
    assert np.abs(ucc.make_rdm2(v) - a2).max() < 1e-13 * np.abs(a2).max()
    with pytest.raises(ValueError):
        ucc.make_rdm1(c, basis="AO")


def test_sharded_full_size_16o_matches_single_gpu():
    """BASELINE.json's (16e,16o) UCCSD with the CI vector sharded over 2 virtual ranks (both H c modes) against the
    single-GPU engine on the same parameters: energy and all 2928 gradients."""
    import torch

    from tencirchem_b200.sharded import LocalComm, ShardedUCC

    n = 16
    int1e, int2e = random_integral(n)
    ucc = UCCSD.from_integral(int1e, int2e, n)
    params = 0.05 * _params(ucc.n_params)
    clear_cache()
    plan = Plan(ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids)
    plan.set_hamiltonian(int1e, int2e)
    e1, g1 = plan.energy_and_grad_dev(params)
    del plan
    torch.cuda.empty_cache()
    for mode in ("replicated", "distributed"):
        # the reverse sweep moves ket and bra separately in one run and as one interleaved block in the other
        eng = ShardedUCC(ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, LocalComm(2), sigma_mode=mode,
                         chunk_bytes=256 << 20, fuse_reverse_moves=(mode == "distributed"))
        eng.set_hamiltonian(int1e, int2e)
        e2, g2 = eng.energy_and_grad(params)
        assert abs(e1 - e2) < E_TOL, (mode, e1, e2)
        assert np.abs(g1 - g2).max() < G_TOL, mode
        assert eng.reshards > 0
        del eng
        torch.cuda.empty_cache()


def test_sharded_distributed_sigma_peak_memory():
    """Every shard-sized block of the sharded driver lives in a fixed pool of FOUR slots per rank (ket + sigma + one moved
    copy + its partial); a fifth live shard is an out-of-memory error at (20e,20o) on 8 GPUs (33.5 GiB shards, 180 GB
    HBM).  The distributed H c must run inside the pool (exhaustion raises) and allocate next to nothing beyond it."""
    import torch

    from tencirchem_b200.sharded import LocalComm, ShardedUCC

    n, world = 12, 2
    int1e, int2e = random_integral(n)
    ucc = UCCSD.from_integral(int1e, int2e, n)
    ex_ops, param_ids = ucc.ex_ops[:40], list(range(40))
    eng = ShardedUCC(2 * n, n, ex_ops, param_ids, LocalComm(world), sigma_mode="distributed", chunk_bytes=256 << 10,
                     fuse_reverse_moves=False)
    assert eng.pool.n_slots == 4
    eng.set_hamiltonian(int1e, int2e)
    params = 0.1 * _params(40)
    kets, layout = eng._forward(params)
    shard = eng.pool.slot * 8
    torch.cuda.synchronize()
    base = torch.cuda.memory_allocated()
    torch.cuda.reset_peak_memory_stats()
    bras = eng._sigma_distributed(kets, layout)
    torch.cuda.synchronize()
    extra = torch.cuda.max_memory_allocated() - base  # outside the pool: the staging of the column transposes only
    assert extra < world * 0.5 * shard, (extra / shard, "shards allocated outside the pool")
    assert sum(eng.pool._used) == 2  # ket and sigma; every temporary slot was given back
    # and the result is H c: compare with the replicated mode
    eng2 = ShardedUCC(2 * n, n, ex_ops, param_ids, LocalComm(world), sigma_mode="replicated")
    eng2.set_hamiltonian(int1e, int2e)
    e1 = sum(float(torch.dot(b.reshape(-1), k.reshape(-1))) for b, k in zip(bras, kets))
    assert abs(e1 - eng2.energy(params)) < E_TOL
    # the whole evaluation, gradient included, runs in the four slots, and equally through the pack / exchange / unpack
    # transport of a system without peer mapping
    e_p2p, g_p2p = eng.energy_and_grad(params)
    eng3 = ShardedUCC(2 * n, n, ex_ops, param_ids, LocalComm(world, peer_memory=False), sigma_mode="distributed",
                      chunk_bytes=256 << 10, fuse_reverse_moves=True)
    assert eng3.pool.n_slots == 6 and not eng3.pool.p2p and eng.pool.p2p
    eng3.set_hamiltonian(int1e, int2e)
    e_st, g_st = eng3.energy_and_grad(params)
    assert abs(e_p2p - e_st) < E_TOL and np.abs(g_p2p - g_st).max() < G_TOL
