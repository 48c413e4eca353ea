"""Generates tests/golden/sparse_parity_*.json: known answers of the SPARSE oracle (oracle/sparse.py) for the full
unscreened UCCSD op list at (12e,12o) ... (20e,20o), sizes where no dense oracle vector fits.

    python tests/golden/make_sparse_golden.py [n_orb ...]        (default 12 16 18 20; minutes of CPU per case)

A case = the FULL UCCSD excitation list of the active space (so the GPU engine builds the same plan -- batches, tiles,
hop tables, row layouts -- as for the benchmark), a sparse random initial state spread over the whole CI space, and a
parameter vector that is zero except for a few parameters covering every excitation kind (alpha / beta single,
alpha-alpha / beta-beta double, paired and mixed opposite-spin doubles).  Zero-angle factors are exact identities in
both implementations, but their gradients are not zero: the expected gradient is stored for the non-zero parameters
plus a random sample of the others.  Stored: energy <psi|H|psi>, the sampled gradient entries, |psi|^2, and samples of
psi and of sigma = H psi (largest entries + random entries).

Also writes tests/golden/dense_12o.npz with the DENSE oracle (oracle/civector_numpy.py, nocache engine) on the full
(12e,12o) UCCSD with dense random parameters: BASELINE.json configs[2].
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import civector_numpy as O  # noqa: E402
from oracle import pools  # noqa: E402
from oracle import sparse as S  # noqa: E402


def random_strings(rng, n_orb, k, count):
    out = np.zeros(count, dtype=np.uint64)
    for i in range(count):
        occ = rng.choice(n_orb, size=k, replace=False)
        out[i] = sum(1 << int(o) for o in occ)
    return out


def pick_params(rng, ex_ops, param_ids, n_orb):
    """a few parameter ids of every excitation kind"""
    kinds = {}
    for op, pid in zip(ex_ops, param_ids):
        spins = tuple(int(i >= n_orb) for i in op)
        if len(op) == 2:
            kind = "single"
        elif len(set(spins)) == 1:
            kind = "same"
        else:
            kind = "mixed"
        kinds.setdefault((kind, pid), []).append(op)
    by_kind = {"single": [], "same": [], "paired": [], "mixed": []}
    for (kind, pid), ops in kinds.items():
        if kind == "mixed" and len(ops) == 1:
            kind = "paired"
        by_kind[kind].append(pid)
    chosen = []
    for kind, cnt in (("single", 2), ("same", 2), ("paired", 2), ("mixed", 4)):
        ids = sorted(set(by_kind[kind]))
        chosen += [int(x) for x in rng.choice(ids, size=min(cnt, len(ids)), replace=False)]
    return sorted(set(chosen))


def make_case(n_orb, nnz0, n_extra_grad, seed=2077):
    rng = np.random.RandomState(seed + n_orb)
    no = n_orb // 2
    nv = n_orb - no
    n_qubits, n_elec = 2 * n_orb, n_orb
    ex_ops, param_ids = pools.uccsd_ex_ops(no, nv)
    n_params = max(param_ids) + 1
    int1e, int2e = O.random_integral(n_orb)  # bench.py's integrals (seed 2077)
    a = random_strings(rng, n_orb, n_elec // 2, nnz0)
    b = random_strings(rng, n_orb, n_elec // 2, nnz0)
    strings, amps = S._merge((a << np.uint64(n_orb)) | b, rng.randn(nnz0))
    amps /= np.linalg.norm(amps)
    nz_ids = pick_params(rng, ex_ops, param_ids, n_orb)
    params = np.zeros(n_params)
    params[nz_ids] = rng.rand(len(nz_ids)) - 0.5
    others = [p for p in range(n_params) if p not in nz_ids]
    grad_ids = sorted(set(nz_ids) | set(int(x) for x in rng.choice(others, size=min(n_extra_grad, len(others)), replace=False)))
    t0 = time.time()
    e, grads, psi, sigma = S.get_energy_and_grad(params, int1e, int2e, n_qubits, n_elec, ex_ops, param_ids,
                                                 init_state=(strings, amps), grad_params=grad_ids)
    dt = time.time() - t0

    def sample(state, count):
        s, v = state
        top = np.argsort(-np.abs(v))[:count]
        rnd = rng.choice(len(s), size=min(count, len(s)), replace=False)
        idx = np.unique(np.concatenate([top, rnd]))
        return [int(x) for x in s[idx]], [float(x) for x in v[idx]]

    psi_s, psi_v = sample(psi, 48)
    # sigma is only known on the oracle's target set; sample there
    sig_s, sig_v = sample(sigma, 48)
    case = dict(
        n_orb=n_orb, n_qubits=n_qubits, n_elec=n_elec, n_ops=len(ex_ops), n_params=n_params, integral_seed=2077,
        init_strings=[int(x) for x in strings], init_amps=[float(x) for x in amps],
        param_ids_nonzero=nz_ids, param_values=[float(params[i]) for i in nz_ids],
        energy=float(e), grad_ids=grad_ids, grad_values=[float(grads[i]) for i in grad_ids],
        norm2=float(np.dot(psi[1], psi[1])), psi_nnz=int(len(psi[0])), sigma_targets=int(len(sigma[0])),
        psi_sample_strings=psi_s, psi_sample_amps=psi_v, sigma_sample_strings=sig_s, sigma_sample_values=sig_v,
        oracle_seconds=round(dt, 1),
    )
    return case


def make_dense_12o():
    n = 12
    ex_ops, param_ids = pools.uccsd_ex_ops(n // 2, n // 2)
    n_params = max(param_ids) + 1
    int1e, int2e = O.random_integral(n)
    np.random.seed(2077)
    params = np.random.rand(n_params) - 0.5
    h = O.get_h_fcifunc_from_integral(int1e, int2e, n)
    t0 = time.time()
    e, g = O.get_energy_and_grad_nocache(params, h, 2 * n, n, ex_ops, param_ids)
    c = O.get_civector_nocache(params, 2 * n, n, ex_ops, param_ids)
    idx = np.unique(np.concatenate([np.argsort(-np.abs(c))[:64], np.random.RandomState(1).choice(len(c), 192, replace=False)]))
    np.savez(os.path.join(HERE, "dense_12o.npz"), params=params, energy=e, grad=g, c_index=idx, c_values=c[idx],
             c_norm2=float(c @ c), c_sum=float(c.sum()), seconds=time.time() - t0)
    print("dense 12o:", e, "in", round(time.time() - t0, 1), "s")


if __name__ == "__main__":
    sizes = [int(x) for x in sys.argv[1:] if x != "dense12"] or [12, 16, 18, 20]
    if "dense12" in sys.argv[1:] or len(sys.argv) == 1:
        make_dense_12o()
    for n_orb in sizes if any(x != "dense12" for x in sys.argv[1:]) or len(sys.argv) == 1 else []:
        nnz0 = {8: 40, 12: 400, 16: 400, 18: 300, 20: 300}.get(n_orb, 200)
        case = make_case(n_orb, nnz0, n_extra_grad=24)
        path = os.path.join(HERE, f"sparse_parity_{n_orb}o.json")
        with open(path, "w") as f:
            json.dump(case, f)
        print(n_orb, "E =", case["energy"], "psi nnz", case["psi_nnz"], "targets", case["sigma_targets"], case["oracle_seconds"], "s")
