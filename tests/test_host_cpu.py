"""CPU: host logic of the product package and the C-ABI surface (no compute calls without a GPU)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from oracle import civector_numpy as O
from oracle import pools
from tencirchem_b200 import KUPCCGSD, PUCCD, UCCSD, _lib, ci_utils
from tencirchem_b200.plan import Plan

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
KAT = json.load(open(os.path.join(HERE, "golden", "published_kat.json")))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "tcc_b200.h")).read()
    declared = set(re.findall(r"\b(tcc_[a-z0-9_]+)\s*\(", header))
    declared.discard("tcc_plan")
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in _lib.load().tcc_version()


def test_ci_strings_bit_exact():
    k = KAT["ci_strings"]
    cs, s2a = ci_utils.get_ci_strings(4, 2, False, strs2addr=True)
    assert cs.tolist() == k["h2_ci_strings"] and s2a.tolist() == k["h2_strs2addr"]
    assert cs.dtype == np.uint64
    for nq, ne, hcb in [(8, 4, False), (12, 6, False), (10, 4, False), (16, 8, False), (6, 4, True), (12, 6, True), (4, 0, False)]:
        a, ta = ci_utils.get_ci_strings(nq, ne, hcb, strs2addr=True)
        b, tb = O.get_ci_strings(nq, ne, hcb, strs2addr=True)
        np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(ta, tb)
    assert ci_utils.get_ex_bitstring(8, 4, (6, 2, 0, 4), False) == k["h4_ex_6204_bitstring"]
    for bits, addr in k["h2_addr"].items():
        assert int(ci_utils.get_addr(np.array([int(bits, 2)]), 4, 2, s2a, False)[0]) == addr
    with pytest.raises(ValueError, match="Too many qubits"):
        ci_utils.get_ci_strings(64, 2, False)
    with pytest.raises(ValueError, match="Too many qubits"):
        Plan(64, 2, (), (), device=None)


@pytest.mark.parametrize("no,nv", [(1, 1), (2, 2), (3, 3), (2, 4), (3, 2)])
def test_excitation_tables_bit_exact(no, nv):
    """tcc_excitation_tables == get_operators (evolve_civector.py:91-106) for every entry, including the
    reference's partner index of determinants outside the support."""
    n = no + nv
    ops, _ = pools.uccsd_ex_ops(no, nv)
    gops, _ = pools.kupccgsd_ex_ops(no, nv, 1)
    extra = [(4, 0, 3, 7), (7, 3, 0, 4), (6, 2, 0, 4), (2, 7, 5, 0)] if n == 4 else []
    all_ops = list(dict.fromkeys(ops + gops + extra))
    plan = Plan(2 * n, 2 * no, all_ops, device=None)
    cs, s2a = O.get_ci_strings(2 * n, 2 * no, False, True)
    for j, op in enumerate(all_ops):
        perm, phase, f2 = O.get_operators(2 * n, 2 * no, s2a, op, cs, False)
        perm2, phase2, f22 = plan.excitation_tables(j)
        np.testing.assert_array_equal(perm, perm2, err_msg=str(op))
        np.testing.assert_array_equal(phase.astype(np.int8), phase2, err_msg=str(op))
        np.testing.assert_array_equal(f2.astype(np.int8), f22, err_msg=str(op))


@pytest.mark.parametrize("no,nv", [(1, 1), (2, 2), (3, 4)])
def test_excitation_tables_hcb_bit_exact(no, nv):
    n = no + nv
    ops, pids = pools.puccd_ex_ops(no, nv)
    plan = Plan(n, 2 * no, ops, pids, hcb=True, device=None)
    cs, s2a = O.get_ci_strings(n, 2 * no, True, True)
    for j, op in enumerate(ops):
        perm, phase, f2 = O.get_operators(n, 2 * no, s2a, op, cs, True)
        perm2, phase2, f22 = plan.excitation_tables(j)
        np.testing.assert_array_equal(perm, perm2)
        np.testing.assert_array_equal(phase.astype(np.int8), phase2)
        np.testing.assert_array_equal(f2.astype(np.int8), f22)


def test_bad_excitations_raise_like_the_reference():
    for bad in [(1, 1), (0, 1, 1, 2)]:  # evolve_civector.py:92-93
        with pytest.raises(ValueError, match="not supported"):
            Plan(8, 4, [bad], device=None)
    for bad in [(5, 0), (4, 5, 0, 1), (0, 1, 2)]:  # silently wrong upstream / assert at :37
        with pytest.raises(ValueError, match="not supported"):
            Plan(8, 4, [bad], device=None)
    with pytest.raises(ValueError):
        Plan(8, 4, [(4, 0)], [0, 1], device=None)


def test_op_pools_match_published_lists():
    p = KAT["pools"]
    int1e, int2e = O.random_integral(2)
    u = UCCSD(int1e, int2e, 2)
    assert [list(o) for o in u.ex_ops] == p["uccsd_h2"]["ex_ops"] and u.param_ids == p["uccsd_h2"]["param_ids"]
    k = KUPCCGSD(int1e, int2e, 2)
    assert [list(o) for o in k.ex_ops] == p["kupccgsd_h2"]["ex_ops"] and k.param_ids == p["kupccgsd_h2"]["param_ids"]
    pu = PUCCD(int1e, int2e, 2)
    assert [list(o) for o in pu.ex_ops] == p["puccd_h2"]["ex_ops"] and pu.param_ids == p["puccd_h2"]["param_ids"]
    int1e, int2e = O.random_integral(4)
    u = UCCSD(int1e, int2e, 4)
    o1, i1, _ = u.get_ex1_ops()
    o2, i2, _ = u.get_ex2_ops()
    assert [list(o) for o in o1] == p["h4_ex1"]["ex_ops"] and i1 == p["h4_ex1"]["param_ids"]
    assert [list(o) for o in o2] == p["h4_ex2"]["ex_ops"] and i2 == p["h4_ex2"]["param_ids"]
    for no, nv in [(2, 2), (3, 5), (4, 4)]:
        i1, i2 = np.zeros((no + nv,) * 2), np.zeros((no + nv,) * 4)
        u = UCCSD(i1, i2, 2 * no)
        assert (u.ex_ops, u.param_ids) == tuple(pools.uccsd_ex_ops(no, nv))
        k = KUPCCGSD(i1, i2, 2 * no)
        assert (k.ex_ops, k.param_ids) == tuple(pools.kupccgsd_ex_ops(no, nv, 3))
        pu = PUCCD(i1, i2, 2 * no)
        assert (pu.ex_ops, pu.param_ids) == tuple(pools.puccd_ex_ops(no, nv))


def test_sizes_at_baseline_configs():
    """SURVEY.md section 8 size table: (n, M, P) of unscreened UCCSD."""
    for n, m, p, size in [(4, 26, 15, 36), (8, 360, 188, 4900), (12, 1818, 927, 853776), (16, 5792, 2928, 165636900)]:
        ops, pids = pools.uccsd_ex_ops(n // 2, n // 2)
        assert len(ops) == m and max(pids) + 1 == p
        u = UCCSD(np.zeros((n, n)), np.zeros((n,) * 4), n)
        assert u.civector_size == size and len(u.ex_ops) == m and u.n_params == p


def test_no_cpu_fallback():
    """Compute entry points must fail loudly without a GPU instead of silently computing on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    plan = Plan(4, 2, [(3, 2), (1, 0), (1, 3, 2, 0)], [0, 0, 1], device=None)
    with pytest.raises(_lib.EngineUnavailable):
        plan.civector(np.zeros(2))
    with pytest.raises(_lib.EngineUnavailable):
        Plan(4, 2, [(3, 2)], [0], device=0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tencirchem_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".hpp", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` runs on the host cores (no GPU) and prints ONE JSON line with the contract's keys;
    under torchrun only rank 0 prints."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "4o", "--steps", "1", "--warmup", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "uccsd_energy_and_grad_evals_per_s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # a non-zero rank of a multi-rank launch exits 0 without output
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run(cmd + ["--gpus", "2"], capture_output=True, text=True, timeout=300, cwd=root, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_mp2_screened_uccsd_matches_oracle_and_published_counts():
    """MP2-screened, amplitude-sorted UCCSD (uccsd.py:107-178, the reference's default ansatz): the product's ex_ops /
    param_ids / init_guess against the oracle restatement (which goes through the full spatial2spin tensor), the MP2
    energy of H4 and the H2 initial guess the reference prints, and the parameter counts of its quickstart table
    (docs/source/quickstart.rst:75-104: 11 / 39 / 108 / 246 / 495)."""
    from oracle import pools as OP
    from tencirchem_b200.ansatz import UCCSD
    from tencirchem_b200.hchain import h_chain_integrals, mp2_t2

    int1e, int2e, e_nuc, e_hf, e_orb = h_chain_integrals(2, 0.741, True)
    e_corr, t2 = mp2_t2(int2e, e_orb, 1)
    u = UCCSD(int1e, int2e, 2, e_nuc, t2=t2)
    assert u.ex_ops == [(3, 2), (1, 0), (1, 3, 2, 0)] and u.param_ids == [0, 0, 1]  # uccsd.py:139-144
    assert abs(u.init_guess[1] - (-0.07260814651571333)) < 1e-12  # ucc_functions.ipynb:115
    published = {4: 11, 6: 39, 8: 108, 10: 246, 12: 495}
    for n, n_params in published.items():
        int1e, int2e, e_nuc, e_hf, e_orb = h_chain_integrals(n, 0.8, True)
        e_corr, t2 = mp2_t2(int2e, e_orb, n // 2)
        if n == 4:
            assert abs(e_hf + e_corr - (-2.151794)) < 5e-7  # adapt_vqe.ipynb:411
        u = UCCSD(int1e, int2e, n, e_nuc, t2=t2)
        assert u.n_params == n_params
        if n <= 8:
            e_o, t2_o = OP.mp2_t2(int2e, e_orb, n // 2)
            assert abs(e_o - e_corr) < 1e-12 and np.abs(t2_o - t2).max() < 1e-13
            ops, pids, guess = OP.uccsd_ex_ops_screened(n // 2, n // 2, t2)
            assert u.ex_ops == ops and list(u.param_ids) == pids
            assert np.abs(np.asarray(u.init_guess) - np.asarray(guess)).max() < 1e-15
            for pick, sort in [(False, True), (True, False), (False, False)]:
                v = UCCSD(int1e, int2e, n, e_nuc, t2=t2, pick_ex2=pick, sort_ex2=sort)
                ops, pids, guess = OP.uccsd_ex_ops_screened(n // 2, n // 2, t2, pick, sort)
                assert v.ex_ops == ops and list(v.param_ids) == pids
                assert np.abs(np.asarray(v.init_guess) - np.asarray(guess)).max() < 1e-15


def _canon_pairs(entries, valid_bit):
    """multiset of rotations in orientation-free form: (min index, max index, sign as seen from the smaller index)"""
    out = []
    for e in entries:
        e = int(e)
        if valid_bit and not (e >> 31) & 1:
            assert e == 0  # idle slot
            continue
        i1, i2, sg = e & 0x7FFF, (e >> 15) & 0x7FFF, (e >> 30) & 1
        assert i1 != i2
        out.append((i1, i2, sg) if i1 < i2 else (i2, i1, sg ^ 1))  # (i1, i2, s) and (i2, i1, -s) are the same rotation
    return sorted(out)


@pytest.mark.parametrize("n,tile_elems", [(6, 400), (8, 6144)])
def test_pair_lists_and_streams_hold_the_same_rotations(n, tile_elems, monkeypatch):
    """The host orders the pairs of every excitation for conflict-free shared-memory access (greedy bank order, or
    bipartite matching with pair orientation and idle slots: HostPlan::build_batch / bank_matched).  Whatever the order:
    the plain list, the 8-byte stream and the 16-byte stream of an excitation must hold exactly the same rotations as
    the unordered list (each pair once, (i1, i2, s) == (i2, i1, -s)), idle slots only in the LAST round of a stream
    (the lean kernels rely on a lane with an entry in round r having entries in all rounds before it), and the
    matching order must not have more bank clashes per quarter-warp group than the unordered list."""
    from tencirchem_b200 import UCCSD
    from tencirchem_b200.hamiltonian import random_integral
    from tencirchem_b200.plan import Plan

    int1e, int2e = random_integral(n)
    ucc = UCCSD.from_integral(int1e, int2e, n)
    monkeypatch.setenv("TCC_B200_TILE_ELEMS", str(tile_elems))

    def clashes(stream, group):
        total = 0
        for g0 in range(0, len(stream), group):
            grp = [int(e) for e in stream[g0:g0 + group] if int(e) >> 31 & 1]
            for shift in (0, 15):
                banks = [(e >> shift) & 0x7FFF & (group - 1) for e in grp]
                total += max([banks.count(b) for b in set(banks)] or [1]) - 1
        return total

    plans = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("TCC_B200_BANK_ORDER", mode)
        plans[mode] = Plan(2 * n, n, ucc.ex_ops, ucc.param_ids, device=None)
    ordered, plain = plans["1"], plans["0"]
    assert ordered.n_batches == plain.n_batches
    checked = padded = 0
    clash_ordered = clash_plain = 0
    for b in range(ordered.n_batches):
        info = ordered.batch_info(b)
        for pos in range(info["n_ops"]):
            for ea in range(info["active_alpha"] + 1):
                for eb in range(info["active_beta"] + 1):
                    ref = _canon_pairs(plain.batch_pairs(b, pos, ea, eb, 0), False)
                    assert len(set((i, j) for i, j, _ in ref)) == len(ref)  # every pair once
                    assert _canon_pairs(ordered.batch_pairs(b, pos, ea, eb, 0), False) == ref
                    for which in (1, 2):
                        s_ord, tps = ordered.batch_pairs(b, pos, ea, eb, which, with_round_len=True)
                        s_pl = plain.batch_pairs(b, pos, ea, eb, which)
                        if len(s_ord) == 0:
                            continue
                        assert _canon_pairs(s_ord, True) == ref and _canon_pairs(s_pl, True) == ref
                        # per lane, the rounds that hold an entry are a prefix: idle slots only where no later round has one
                        assert tps > 0 and len(s_ord) % tps == 0
                        valid = ((s_ord >> 31) & 1).reshape(-1, tps)
                        assert (np.diff(valid.astype(np.int8), axis=0) <= 0).all()
                        padded += int((valid[: max(1, -(-len(ref) // tps)) - 1] == 0).sum())
                        if which == 2 and len(ref) >= 64:
                            clash_ordered += clashes(s_ord, 8)
                            clash_plain += clashes(s_pl, 8)
                        checked += 1
    assert checked > 0 and padded == 0  # no idle slot before the last round
    assert clash_ordered <= clash_plain
    if n == 8:
        assert clash_ordered < 0.5 * clash_plain  # the matching removes most clashes of the unordered streams


def test_bench_arms_share_the_config_keys(capsys):
    """Both bench arms print `config` with the same workload keys (the driver compares them); what only describes how
    this engine ran the workload (sharding, batches, staging) moves to `engine_config`."""
    import importlib.util
    import json

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    line = {"metric": bench.METRIC, "config": {"workload": "w", "n_determinants": 36, "n_excitations": 26, "n_params": 15,
                                               "l2": bench.l2_note(36), "multi_gpu": "single GPU", "n_batches": 3, "staging_s": 0.1}}
    bench.emit(line)
    d = json.loads(capsys.readouterr().out)
    assert tuple(d["config"]) == bench.CONFIG_KEYS
    assert d["engine_config"] == {"multi_gpu": "single GPU", "n_batches": 3, "staging_s": 0.1}
    assert bench.l2_note(165636900) == bench.l2_note(165636900, 1, True) != bench.l2_note(165636900, 8, True)
