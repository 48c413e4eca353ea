"""UCCSD.kernel() (L-BFGS-B over energy_and_grad) on literal H_n chains, STO-3G, 0.8 A -- the reference's quickstart
table (docs/source/quickstart.rst:75-116): MP2-screened and sorted ansatz as in the reference (default), or every
excitation in generator order (--unscreened).  Prints the published columns (parameters, wall time, error to FCI).
usage: hchain_kernel.py [--unscreened] [--no-fci] [--json out.json] n [n ...]"""
import json, os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tencirchem_b200 import UCCSD
from tencirchem_b200.hchain import h_chain_integrals, mp2_t2

PUBLISHED = {4: (11, 0.05, 0.01), 6: (39, 0.4, 0.27), 8: (108, 1.9, 0.72), 10: (246, 14, 1.37), 12: (495, 51, 2.13),
             14: (899, 2093, 2.92), 16: (1520, 50909, 3.72)}  # parameters, wall time (s), error (mH)
args = sys.argv[1:]
screened = "--unscreened" not in args
want_fci = "--no-fci" not in args
out_json = args[args.index("--json") + 1] if "--json" in args else None
sizes = [int(a) for a in args if a.isdigit()] or [4, 6, 8]
rows = []
for n in sizes:
    t0 = time.time(); int1e, int2e, e_nuc, e_hf, e_orb = h_chain_integrals(n, 0.8, True); t_int = time.time() - t0
    e_corr, t2 = mp2_t2(int2e, e_orb, n // 2)
    ucc = UCCSD.from_integral(int1e, int2e, n, e_nuc, t2=t2 if screened else None)
    t0 = time.time(); ucc.energy_and_grad(np.zeros(ucc.n_params)); t_stage = time.time() - t0  # plan + Hamiltonian upload
    t0 = time.time(); e = ucc.kernel(); t_opt = time.time() - t0
    row = {"n": n, "dets": ucc.civector_size, "ops": len(ucc.ex_ops), "params": ucc.n_params, "e_hf": e_hf, "e_mp2": e_hf + e_corr,
           "e_uccsd": e, "nfev": int(ucc.opt_res["nfev"]), "kernel_s": t_opt, "staging_s": t_stage, "integrals_rhf_s": t_int,
           "screened": screened}
    if want_fci:
        t0 = time.time(); row["e_fci"] = ucc.e_fci; row["fci_s"] = time.time() - t0
        row["error_mH"] = 1e3 * (e - row["e_fci"])
    if n in PUBLISHED:
        row["published_params"], row["published_wall_s"], row["published_error_mH"] = PUBLISHED[n]
    rows.append(row)
    print(json.dumps(row), flush=True)
if out_json:
    json.dump(rows, open(out_json, "w"), indent=1)
