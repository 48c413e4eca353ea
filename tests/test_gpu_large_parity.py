"""GPU parity at the BASELINE.json sizes no dense oracle vector reaches: the full unscreened UCCSD plan of (12e,12o) and
(16e,16o) [(18e,18o) with TCC_B200_BIG=1] against the SPARSE oracle's committed known answers
(tests/golden/sparse_parity_*.json <- oracle/sparse.py, pinned against the dense oracle in tests/test_oracle_kat.py),
and the full (12e,12o) UCCSD with dense random parameters against the DENSE oracle (tests/golden/dense_12o.npz).

The sparse cases run the REAL plan of the benchmark (all 1818 / 5792 / 9315 excitations: every batch, tile table, hop
list and row layout is exercised) on a sparse random initial state spread over the whole CI space, with a few non-zero
parameters of every excitation kind; zero-angle factors are exact identities but their gradients are checked too.
Tolerances are the north_star's: energy 1e-10 Ha, gradient 1e-8 Ha/rad, amplitudes 1e-12.
"""
import os

import numpy as np
import pytest

from golden_cases import SparseCase
from tencirchem_b200 import UCCSD
from tencirchem_b200.evolve_civector import clear_cache
from tencirchem_b200.hamiltonian import random_integral
from tencirchem_b200.plan import Plan

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
BIG = os.environ.get("TCC_B200_BIG", "0") == "1"


def _ucc(n):
    int1e, int2e = random_integral(n)
    return UCCSD.from_integral(int1e, int2e, n), int1e, int2e


def test_full_uccsd_12o_against_dense_oracle():
    """BASELINE.json configs[2]: (12e,12o), 853,776 determinants, 1818 excitations, 927 dense random parameters."""
    gold = np.load(os.path.join(HERE, "golden", "dense_12o.npz"))
    ucc, _, _ = _ucc(12)
    params = gold["params"]
    e, g = ucc.energy_and_grad(params)
    assert abs(e - float(gold["energy"])) < 1e-10
    assert np.abs(g - gold["grad"]).max() < 1e-8
    c = ucc.civector(params)
    assert np.abs(c[gold["c_index"]] - gold["c_values"]).max() < 1e-12
    assert abs(float(c @ c) - float(gold["c_norm2"])) < 1e-12
    assert abs(float(c.sum()) - float(gold["c_sum"])) < 1e-10


@pytest.mark.parametrize("n", [12, 16] + ([18] if BIG else []))
def test_full_plan_against_sparse_oracle_single_gpu(n):
    import torch

    case = SparseCase(n)
    ucc, int1e, int2e = _ucc(n)
    assert len(ucc.ex_ops) == case.d["n_ops"] and ucc.n_params == case.n_params
    clear_cache()
    plan = Plan(ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids)
    plan.set_hamiltonian(int1e, int2e)
    nstr = plan.n_strings
    init = torch.zeros(plan.size, dtype=torch.float64, device="cuda")
    ia, ib, iv = case.init_sparse()
    init[torch.as_tensor(ia * nstr + ib, device="cuda")] = torch.as_tensor(iv, device="cuda")
    e, g = plan.energy_and_grad_dev(case.params, init=init)
    res = case.check(e, g)
    assert res["ok"], res
    # amplitudes of psi and of H psi at the sampled determinants
    c = torch.empty_like(init)
    plan.civector_dev(case.params, c, init=init)
    pos = torch.as_tensor(case.flat_address(case.d["psi_sample_strings"], nstr), device="cuda")
    assert np.abs(c[pos].cpu().numpy() - np.asarray(case.d["psi_sample_amps"])).max() < 1e-12
    assert abs(float(torch.dot(c, c)) - case.norm2) < 1e-11
    sig = torch.empty_like(init)
    plan.apply_h_dev(c, sig)
    pos = torch.as_tensor(case.flat_address(case.d["sigma_sample_strings"], nstr), device="cuda")
    assert np.abs(sig[pos].cpu().numpy() - np.asarray(case.d["sigma_sample_values"])).max() < 1e-11
    del plan, init, c, sig
    torch.cuda.empty_cache()


@pytest.mark.parametrize("n,world,mode", [(12, 4, "distributed"), (12, 8, "replicated"), (16, 2, "distributed")]
                         + ([(16, 8, "distributed")] if BIG else []))
def test_full_plan_against_sparse_oracle_sharded(n, world, mode):
    """The alpha-row sharded engine (virtual ranks on one GPU: every layout change, ownership map and partial sum of
    the N-GPU run) against the sparse oracle, with the sparse initial state scattered over the ranks."""
    import torch

    from tencirchem_b200.sharded import LocalComm, ShardedUCC

    case = SparseCase(n)
    ucc, int1e, int2e = _ucc(n)
    clear_cache()
    eng = ShardedUCC(ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, LocalComm(world), sigma_mode=mode,
                     chunk_bytes=256 << 20)
    eng.set_hamiltonian(int1e, int2e)
    e, g = eng.energy_and_grad(case.params, init_sparse=case.init_sparse())
    res = case.check(e, g)
    assert res["ok"], res
    assert eng.reshards > 0
    del eng
    torch.cuda.empty_cache()
