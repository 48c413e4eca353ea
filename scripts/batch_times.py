"""Per-batch device times of one UCCSD energy_and_grad (diagnostic).  usage: batch_times.py [n_orbitals]"""
import os, sys, json, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tencirchem_b200 import UCCSD
from tencirchem_b200.hamiltonian import random_integral
from tencirchem_b200.plan import Plan
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
i1, i2 = random_integral(n)
u = UCCSD.from_integral(i1, i2, n)
p = Plan(2 * n, n, u.ex_ops, u.param_ids)
p.set_hamiltonian(i1, i2)
np.random.seed(1); params = np.random.rand(u.n_params) - 0.5
p.energy_and_grad_dev(params)
p.profile_batches(True)
p.energy_and_grad_dev(params)
fwd, rev = p.profile_batches(False)
infos = [p.batch_info(b) for b in range(p.n_batches)]
nops = np.array([i["n_ops"] for i in infos], dtype=float)
A = np.vstack([nops, np.ones_like(nops)]).T
print("batches", len(infos), "fwd total ms %.1f rev total ms %.1f" % (fwd.sum(), rev.sum()))
print("fwd ms = %.4f * nops + %.3f" % tuple(np.linalg.lstsq(A, fwd, rcond=None)[0]))
print("rev ms = %.4f * nops + %.3f" % tuple(np.linalg.lstsq(A, rev, rcond=None)[0]))
kinds = {}
for i, f, r in zip(infos, fwd, rev):
    k = (i["active_alpha"], i["active_beta"], i["n_tiles"], i["max_tile_rows"], i["max_tile_cols"])
    kinds.setdefault(k, []).append((i["n_ops"], f, r))
for k, v in kinds.items():
    v = np.array(v)
    print(k, "n=%d ops=%d fwd %.1f ms rev %.1f ms" % (len(v), v[:, 0].sum(), v[:, 1].sum(), v[:, 2].sum()))
out = os.environ.get("BATCH_TIMES_JSON")
if out:
    json.dump({"nops": nops.tolist(), "fwd_ms": fwd.tolist(), "rev_ms": rev.tolist()}, open(out, "w"))
