#!/usr/bin/env python
"""Benchmark of the B200 civector UCC engine: UCCSD energy_and_grad evaluations/s.

Contract (one JSON line on stdout from rank 0):
  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload 16o|12o|8o|4o]
* a "step" is ONE energy_and_grad evaluation (forward sweep, H.c, energy, reverse sweep, all gradients)
  of unscreened UCCSD in generator order on synthetic inputs: random_integral(n, seed 2077) as MO
  integrals, theta = rand(P) - 0.5 with seed 2077 + step.
* default workload: (16e,16o), N = 165,636,900 determinants, M = 5792 excitations, P = 2928 parameters
  (BASELINE.json configs[3], the size its target is quoted on; a 1.3 GB vector >> the 126 MB L2, so no
  explicit L2 flush is needed between steps).
* N > 1 (default `--multi sharded`): ONE evaluation at a time with the CI vector split by alpha rows over the N
  ranks (tencirchem_b200/sharded.py; NCCL row moves between the row layouts of the batch segments), every rank
  evaluates the same theta, `value` = evaluations / max-over-ranks device time, scaling "strong".
  `--multi replicas` / `--batch B`: every rank evaluates its own theta sets (no data-path collective), scaling "weak".
* `--parity-golden`: before the timed region the engine evaluates the committed sparse-oracle case of the workload
  (tests/golden/sparse_parity_<n>o.json: full op list, sparse initial state, a few non-zero parameters) and the JSON
  line carries the absolute errors (`parity`); it doubles as a warm-up evaluation.
* `--impl reference`: the CPU oracle port of the reference numpy engine (oracle/civector_numpy.py; the
  Python reference itself cannot be imported here -- DESIGN.md) timed on a bounded sample, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {"20o": 20, "18o": 18, "16o": 16, "14o": 14, "12o": 12, "10o": 10, "8o": 8, "6o": 6, "4o": 4, "2o": 2}
METRIC = "uccsd_energy_and_grad_evals_per_s"


def workload_desc(n, ansatz="uccsd", integrals="random"):
    names = {"uccsd": "UCCSD (unscreened, generator order)", "kupccgsd": "kUpCCGSD (k=3)", "puccd": "pUCCD (hard-core bosons)"}
    src = "random_integral seed 2077" if integrals == "random" else f"H{n} chain STO-3G 0.8 A (RHF MO integrals)"
    return f"{names[ansatz]} energy_and_grad ({n}e,{n}o), {src}"


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------
def cpu_reference_sample(n, n_ops_sample=2, ansatz="uccsd"):
    """Times the oracle port of the reference numpy engine (nocache / 'civector-large' form, which is what
    the reference selects above 16 qubits, ucc.py:233-239) on a bounded sample and extrapolates to one
    energy_and_grad evaluation by operation counts."""
    from oracle import civector_numpy as O
    from oracle import pools

    no = n // 2
    hcb = ansatz == "puccd"
    ops, pids = {"uccsd": pools.uccsd_ex_ops, "kupccgsd": pools.kupccgsd_ex_ops, "puccd": pools.puccd_ex_ops}[ansatz](no, n - no)
    m = len(ops)
    nq = n if hcb else 2 * n
    t0 = time.perf_counter()
    ci_strings, strs2addr = O.get_ci_strings(nq, n, hcb, strs2addr=True)
    t_strings = time.perf_counter() - t0
    size = len(ci_strings)
    rng = np.random.default_rng(0)
    ket = rng.standard_normal(size)
    bra = rng.standard_normal(size)
    # one op of each cost class is enough for the reference: its [N] tables make every excitation cost the same
    picks = [ops[0], ops[len(ops) // 3], ops[-1]][:n_ops_sample] if n_ops_sample <= 3 else ops[:n_ops_sample]
    _ = bra[:1024] @ ket[:1024]  # spin up BLAS threads outside the timed sample
    t_tab = t_fwd = t_rev = 0.0
    for f_idx in picks:
        t0 = time.perf_counter()
        perm, phase, f2 = O.get_operators(nq, n, strs2addr, tuple(f_idx), ci_strings, hcb)
        t1 = time.perf_counter()
        ket = O.evolve_excitation(ket, perm, phase, f2, 1 - np.cos(0.3), np.sin(0.3))
        t2 = time.perf_counter()
        bra = O.evolve_excitation(bra, perm, phase, f2, 1 - np.cos(0.3), -np.sin(0.3))
        ket = O.evolve_excitation(ket, perm, phase, f2, 1 - np.cos(0.3), -np.sin(0.3))
        _ = bra @ (ket[perm] * phase)
        t3 = time.perf_counter()
        t_tab += t1 - t0
        t_fwd += t2 - t1
        t_rev += t3 - t2
        del perm, phase, f2
    k = len(picks)
    t_tab, t_fwd, t_rev = t_tab / k, t_fwd / k, t_rev / k
    del ket, bra, ci_strings
    # H.c: the oracle's Knowles-Handy port at a size that fits, scaled by the 2 n^4 N flop count
    from math import comb
    if hcb:
        ns = n
        i1, i2 = O.random_integral(ns)
        h = O.get_h_fcifunc_hcb_from_integral(i1, i2, ns)
        x = rng.standard_normal(comb(ns, ns // 2))
        scale = 1.0
    else:
        ns = min(n, 10)
        i1, i2 = O.random_integral(ns)
        h = O.get_h_fcifunc_from_integral(i1, i2, ns)
        x = rng.standard_normal(comb(ns, ns // 2) ** 2)
        scale = (n**4 * comb(n, n // 2) ** 2) / (ns**4 * comb(ns, ns // 2) ** 2)
    h(x)
    t0 = time.perf_counter()
    h(x)
    t_sig_small = time.perf_counter() - t0
    t_sigma = t_sig_small * scale
    # civector-large rebuilds the tables in both sweeps (evolve_civector.py:253-308); up to 16 qubits the
    # reference defaults to the cached-table engine (ucc.py:233-239) and the build is staging, not per call
    cached = nq <= 16
    t_eval = m * ((0.0 if cached else t_tab) + t_fwd) + t_sigma + m * ((0.0 if cached else t_tab) + t_rev)
    sample = (f"{k} of {m} excitations at full size N={size} (table build {t_tab:.2f}s, rotation {t_fwd:.2f}s, reverse step "
              f"{t_rev:.2f}s each; strings {t_strings:.1f}s) + H.c timed at ({ns}e,{ns}o) = {t_sig_small:.2f}s scaled x{scale:.0f} by 2n^4N; "
              f"extrapolated by operation counts to one evaluation = {t_eval:.0f}s")
    # the reference's numpy engine is single-threaded in the gathers / elementwise passes that make up >99.9 % of
    # the evaluation; only the dgemm inside H.c uses the BLAS thread pool
    return {"value": 1.0 / t_eval, "unit": "evals/s", "cores": 1, "kind": "port", "sample": sample,
            "host_cpus": os.cpu_count() or 1,
            "note": "numpy fancy-index gathers and elementwise passes are single-threaded, exactly as in the reference's numpy backend; the BLAS pool (all host cpus) is used only by the dgemm of H.c"}


def scaling_label(args, world):
    """the default N > 1 run shards ONE evaluation over the ranks (fixed total work): a strong-scaling curve, N = 1
    included; replicas / batched theta sets per rank are weak scaling"""
    sharded = args.multi == "sharded" and args.batch == 0 and args.ansatz != "puccd"
    return "strong" if sharded else "weak"


def workload_config(n, args):
    """the workload keys both arms print (same inputs, same sizes)"""
    from math import comb

    from oracle import pools  # integer op-pool generators only (bench legs may import oracle/; the product never does)

    no = n // 2
    hcb = args.ansatz == "puccd"
    ops, pids = {"uccsd": pools.uccsd_ex_ops, "kupccgsd": pools.kupccgsd_ex_ops, "puccd": pools.puccd_ex_ops}[args.ansatz](no, n - no)
    if args.max_ops > 0:
        ops, pids = ops[: args.max_ops], pids[: args.max_ops]
    size = comb(n, no) if hcb else comb(n, no) ** 2
    return {"workload": workload_desc(n, args.ansatz, args.integrals), "n_determinants": size, "n_excitations": len(ops),
            "n_params": len(set(pids))}


CONFIG_KEYS = ("workload", "n_determinants", "n_excitations", "n_params", "l2")


def l2_note(size, world=1, sharded=False):
    """how the timed region keeps L2 from serving the vectors (both arms print the same note for the same workload)"""
    if sharded and world > 1:
        return "inputs larger than L2 (ket + bra shards vs 126 MB)" if 2 * size * 8 / world > 1.26e8 else "vectors fit in L2; no flush"
    return "inputs larger than L2 (1.3 GB vectors vs 126 MB)" if size * 8 > 2.5e8 else "vectors fit in L2; no flush"


def emit(line):
    """print the JSON line; `config` keeps the workload keys both arms share (CONFIG_KEYS), everything else that
    describes how THIS engine ran the workload (sharding, batches, staging time) goes to `engine_config`"""
    cfg = line.get("config", {})
    extra = {k: v for k, v in cfg.items() if k not in CONFIG_KEYS}
    line["config"] = {k: cfg[k] for k in CONFIG_KEYS if k in cfg}
    if extra:
        line["engine_config"] = extra
    print(json.dumps(line), flush=True)


def run_reference(args):
    """`--impl reference`: the CPU port of the reference's numpy engine on this box's host cores, rank 0 only.  One
    full evaluation takes hours at (16e,16o), so a step is a bounded sample extrapolated by operation counts
    (cpu_reference_sample); the sample is timed ONCE (after one smaller warm-up sample) and every step reports it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = WORKLOADS[args.workload]
    if args.warmup > 0:
        cpu_reference_sample(n, n_ops_sample=1, ansatz=args.ansatz)
    res = cpu_reference_sample(n, n_ops_sample=2, ansatz=args.ansatz)
    t_eval = 1.0 / res["value"]
    cfg = workload_config(n, args)
    cfg["l2"] = l2_note(cfg["n_determinants"], args.gpus, args.multi == "sharded" and args.batch == 0 and args.ansatz != "puccd")
    cfg["note"] = ("CPU oracle port of the reference numpy engine; the bounded sample is timed once and extrapolated by "
                   "operation counts to one evaluation; every step reports that figure")
    line = {
        "impl": "reference", "metric": METRIC, "value": 1.0 / t_eval, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_eval, "higher_is_better": True,
        "scaling": scaling_label(args, args.gpus),
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {**res, "value": 1.0 / t_eval},
        "e2e": {"value": 1.0 / t_eval, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def progress(msg):
    """append a line to gpurun_out/bench_progress.log: a long multi-GPU run that dies leaves a trace of how far it got"""
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "bench_progress.log"), "a") as f:
            f.write(f"{time.strftime('%H:%M:%S')} {msg}\n")
    except OSError:
        pass


def golden_parity(n, evaluate, rank=0):
    """Evaluates the committed sparse-oracle case of the workload (tests/golden/sparse_parity_<n>o.json) through
    `evaluate(params, (alpha_ref_addr, beta_ref_addr, amps)) -> (energy, grad)` and returns the absolute errors."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_cases import SparseCase

    case = SparseCase(n)
    t0 = time.perf_counter()
    e, g = evaluate(case.params, case.init_sparse())
    res = case.check(e, g)
    res.update({"case": f"tests/golden/sparse_parity_{n}o.json (sparse oracle, oracle/sparse.py)", "energy": e,
                "energy_expected": case.energy, "seconds": round(time.perf_counter() - t0, 1)})
    return res


def driver_phase_table(summary, steps, step_ms):
    """per-step device / host ms of the sharded driver's phases (ShardedUCC.timer), with the NVLink GB/s of the
    collectives (bytes this rank sent, over the time the compute stream waited for the collective)"""
    out = {}
    for name, d in sorted(summary.items(), key=lambda kv: -kv[1]["device_ms"]):
        row = {"device_ms_per_step": round(d["device_ms"] / steps, 3), "host_ms_per_step": round(d["host_ms"] / steps, 3),
               "calls_per_step": d["calls"] // steps, "share_of_step": round(d["device_ms"] / steps / step_ms, 4) if step_ms else None}
        if d["bytes"]:
            row["GB_sent_per_step"] = round(d["bytes"] / steps / 1e9, 3)
            row["GBps"] = round(d["bytes"] / (d["device_ms"] * 1e-3) / 1e9, 1) if d["device_ms"] > 0 else None
        out[name] = row
    covered = sum(d["device_ms"] for d in summary.values()) / steps
    out["_unattributed_ms_per_step"] = round(step_ms - covered, 3)
    return out


# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from tencirchem_b200 import KUPCCGSD, PUCCD, UCCSD
    from tencirchem_b200.evolve_civector import _bind_hamiltonian, get_plan
    from tencirchem_b200.hamiltonian import random_integral

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    torch.cuda.init()
    if world > 1:
        # keep NCCL's version banner off stdout unless the caller asked for it (the driver reads NCCL_DEBUG=INFO lines)
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = WORKLOADS[args.workload]
    if args.integrals == "hchain":
        from tencirchem_b200.hchain import h_chain_integrals

        int1e, int2e, _, _ = h_chain_integrals(n)
    else:
        int1e, int2e = random_integral(n)
    np.random.seed(2077)  # kUpCCGSD draws its initial guess from the global stream
    ucc = {"uccsd": UCCSD, "kupccgsd": KUPCCGSD, "puccd": PUCCD}[args.ansatz].from_integral(int1e, int2e, n)
    if args.max_ops > 0:
        # truncated ansatz (NOT the headline workload): the first max_ops excitations of the generator order, for
        # demonstrations at sizes where a full evaluation exceeds the time budget; the config line says so
        ucc.ex_ops = ucc.ex_ops[: args.max_ops]
        ids = list(ucc.param_ids[: args.max_ops])
        remap = {v: i for i, v in enumerate(sorted(set(ids)))}
        ucc.param_ids = [remap[v] for v in ids]
    m, p_count = len(ucc.ex_ops), ucc.n_params
    size = ucc.civector_size

    if world > 1 and args.multi == "sharded" and not ucc.hcb and args.batch == 0:
        return run_b200_sharded(args, ucc, int1e, int2e, world, rank, local_rank, n)

    t0 = time.perf_counter()
    plan = get_plan(ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, ucc.hcb, device=local_rank)
    _bind_hamiltonian(plan, ucc.hamiltonian)
    staging_s = time.perf_counter() - t0

    def theta(step):
        np.random.seed((2077 + step + 1000003 * rank) % (2**32))
        return np.random.rand(p_count) - 0.5

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the public API call a user makes routes to the same cached plan (one process per GPU: torch's current device)
    if args.batch > 0:
        return run_b200_batched(args, plan, ucc, theta, barrier, world, rank, n, staging_s)

    parity = None
    if args.parity_golden:
        def evaluate(p_, init):
            ia, ib, iv = init
            x = torch.zeros(plan.size, dtype=torch.float64, device="cuda")
            x[torch.as_tensor(ia * plan.n_strings + ib, device="cuda")] = torch.as_tensor(iv, device="cuda")
            return plan.energy_and_grad_dev(p_, init=x)

        parity = golden_parity(n, evaluate, rank)
    for w in range(args.warmup):
        plan.energy_and_grad_dev(theta(-1 - w))
    barrier()

    # ---- device-timed region: K evaluations, vectors resident in HBM, CUDA events on the launch stream
    plan.profile_enable(True)
    launches0 = plan.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    last = None
    for k in range(args.steps):
        last = plan.energy_and_grad_dev(theta(k))
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    prof = plan.profile_read()
    plan.profile_enable(False)
    launches = plan.launch_count - launches0

    # ---- end-to-end region: the public API with host buffers (params in, energy + gradient out)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        e_host, g_host = ucc.energy_and_grad(theta(k))
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert abs(e_host - last[0]) < 1e-9 * max(1.0, abs(e_host)), (e_host, last[0])

    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    evals = args.steps * world
    value = evals / (dev_ms * 1e-3)
    e2e_value = evals / (e2e_ms * 1e-3)
    peak, peak_src = measured_peak_hbm()
    # dominant kernel = the slot with the largest device time
    dom = max(prof, key=lambda s: prof[s][0])
    d_ms, d_launch, d_bytes = prof[dom]
    achieved = d_bytes / (d_ms * 1e-3) / 1e9 if d_ms > 0 else 0.0
    kernel_names = {"rev_alphabeta": "reverse_kernel<false,false>", "rev_alpha": "reverse_kernel<false,true>",
                    "rev_beta": "reverse_kernel<true,false>", "fwd_alphabeta": "rotate_kernel<false,false>",
                    "fwd_alpha": "rotate_kernel<false,true>", "fwd_beta": "rotate_kernel<true,false>",
                    "sigma_ab": "sigma_ab_kernel", "sigma_ss": "spmm_rows_kernel+transpose_kernel", "dot": "dot_kernel",
                    "fwd_batch": "batch_kernel<1> (one launch per batch of excitations, whole vector read+written once)",
                    "rev_batch": "lean_persist_kernel<3> (+ batch_grad_reduce_kernel; interleaved (ket, bra) elements read+written once per batch)"}
    total_prof_ms = sum(v[0] for v in prof.values())
    phases = {s: {"ms_per_step": round(v[0] / args.steps, 3), "launches_per_step": v[1] // args.steps,
                  "algorithmic_GBps": round(v[2] / (v[0] * 1e-3) / 1e9, 1) if v[0] > 0 else None,
                  "share": round(v[0] / total_prof_ms, 4) if total_prof_ms else None} for s, v in prof.items()}
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(args.workload, {}).get(dom)
    except Exception:
        pass
    dense_bytes = (48 * m + 16) * size
    line = {
        "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak" if world > 1 else scaling_label(args, world), "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_desc(n, args.ansatz, args.integrals), "n_determinants": size, "n_excitations": m, "n_params": p_count,
                   "l2": l2_note(size),
                   "multi_gpu": "independent theta sets per rank, no data-path collective" if world > 1 else "single GPU",
                   "n_batches": plan.n_batches, "staging_s": round(staging_s, 2)},
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": 8 * p_count, "d2h_bytes_per_step": 8 * (m + 1),
                "note": "public API UCCSD.energy_and_grad(params): host params in, host (energy, gradient) out; the CI vectors never leave HBM"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": kernel_names.get(dom, dom), "phase": dom, "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
                     "traffic_source": "ncu --set full capture of this kernel, bytes per launch (profiles/ncu_traffic.json)" if traffic else None,
                     "avg_launch_us": 1e3 * d_ms / max(d_launch, 1), "launches": d_launch,
                     "algorithmic_bytes_per_launch": d_bytes / max(d_launch, 1)},
        "effective_dense_model_GBps": dense_bytes / (dev_ms / args.steps * 1e-3) / 1e9,
        "phases": phases,
        "clocks": clocks,
        "energy_last_step": last[0],
    }
    if parity is not None:
        line["parity"] = parity
    if not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_reference_sample(n, n_ops_sample=2, ansatz=args.ansatz)
        except Exception as exc:  # pragma: no cover
            line["cpu_baseline"] = {"error": repr(exc)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_b200_sharded(args, ucc, int1e, int2e, world, rank, local_rank, n):
    """N > 1: ONE evaluation at a time, the CI vector split by alpha rows over the ranks (strong scaling;
    tencirchem_b200/sharded.py).  Every rank evaluates the same theta; device time = MAX over ranks."""
    import torch
    import torch.distributed as dist

    from tencirchem_b200.evolve_civector import clear_cache
    from tencirchem_b200.sharded import ShardedUCC, TorchComm

    clear_cache()
    torch.cuda.empty_cache()
    m, p_count, size = len(ucc.ex_ops), ucc.n_params, ucc.civector_size
    t0 = time.perf_counter()
    eng = ShardedUCC(ucc.n_qubits, ucc.n_elec, ucc.ex_ops, ucc.param_ids, TorchComm(), device=local_rank, sigma_mode=args.sigma,
                     chunk_bytes=int(args.chunk_gib * 2**30))
    eng.set_hamiltonian(int1e, int2e)
    staging_s = time.perf_counter() - t0
    plan = eng.states[0].plan

    def theta(step):  # the same parameters on every rank, and the same as rank 0 of the single-GPU run
        np.random.seed((2077 + step) % (2**32))
        return np.random.rand(p_count) - 0.5

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    progress(f"rank {rank}: plans + Hamiltonian ready after {staging_s:.1f} s, {torch.cuda.memory_allocated() / 2**30:.1f} GiB allocated by torch")
    parity = None
    if args.parity_golden:
        parity = golden_parity(n, lambda p_, init: eng.energy_and_grad(p_, init_sparse=init), rank)
        progress(f"rank {rank}: golden parity {parity}")
    for w in range(args.warmup):
        eng.energy_and_grad(theta(-1 - w))
        progress(f"rank {rank}: warm-up {w} done")
    barrier()
    plan.profile_enable(True)
    eng.timer.enabled = True
    eng.timer.reset()
    torch.cuda.reset_peak_memory_stats()
    launches0 = plan.launch_count
    reshards0 = eng.reshards
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    last = None
    for k in range(args.steps):
        # host parameters in, host (energy, gradient) out: this call IS the public API of the sharded engine
        last = eng.energy_and_grad(theta(k))
    ev1.record()
    barrier()
    e2e_s = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    prof = plan.profile_read()
    plan.profile_enable(False)
    driver_phases = eng.timer.summary()
    eng.timer.enabled = False
    launches = plan.launch_count - launches0
    reshards = (eng.reshards - reshards0) // max(args.steps, 1)
    peak_gib = torch.cuda.max_memory_allocated() / 2**30
    free_b, total_b = torch.cuda.mem_get_info()
    t = torch.tensor([dev_ms, e2e_s * 1e3, peak_gib, (total_b - free_b) / 2**30], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, peak_gib, used_gib = (float(x) for x in t)
    progress(f"rank {rank}: {args.steps} timed step(s) done, {dev_ms / args.steps:.1f} ms per step (max over ranks)")
    if rank != 0:
        dist.destroy_process_group()
        return
    value = args.steps / (dev_ms * 1e-3)
    peak, peak_src = measured_peak_hbm()
    dom = max(prof, key=lambda s_: prof[s_][0])
    d_ms, d_launch, d_bytes = prof[dom]
    achieved = d_bytes / (d_ms * 1e-3) / 1e9 if d_ms > 0 else 0.0
    total_prof_ms = sum(v[0] for v in prof.values())
    phases = {s_: {"ms_per_step": round(v[0] / args.steps, 3), "launches_per_step": v[1] // args.steps,
                   "algorithmic_GBps": round(v[2] / (v[0] * 1e-3) / 1e9, 1) if v[0] > 0 else None,
                   "share": round(v[0] / total_prof_ms, 4) if total_prof_ms else None} for s_, v in prof.items()}
    line = {
        "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_desc(n, args.ansatz, args.integrals) + (f" TRUNCATED to the first {m} excitations" if args.max_ops > 0 else ""),
                   "n_determinants": size, "n_excitations": m,
                   "n_params": p_count, "l2": l2_note(size, world, True),
                   "multi_gpu": f"one evaluation, CI vector sharded by alpha rows over {world} ranks; {reshards} row re-shards "
                                + ("(rows written straight into the peers' shard pool over NVLink, tcc_push_rows) " if getattr(eng.pool, "p2p", False)
                                   else "(pack, NCCL all-to-all, unpack) ")
                                + f"per evaluation in {len(eng.segments)} segments; "
                                + ("H.c with the vector kept sharded (3 alpha-beta passes on per-layout link tables, alpha-alpha on "
                                   "column blocks, beta-beta local)" if eng.sigma_mode == "distributed" else
                                   "H.c by row blocks of the all-gathered vector, result reduce-scattered")
                                + "; gradient and energy all-reduced",
                   "n_batches": plan.n_batches, "staging_s": round(staging_s, 2)},
        "e2e": {"value": args.steps / (e2e_ms * 1e-3), "unit": "evals/s", "h2d_bytes_per_step": 8 * p_count * world,
                "d2h_bytes_per_step": 8 * (m + 1) * world,
                "note": "ShardedUCC.energy_and_grad(params): host params in, host (energy, gradient) out on every rank; wall clock"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "lean_persist_kernel<3> (rank 0's share of the rows)" if dom == "rev_batch" else dom,
                     "phase": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "traffic": None,
                     "avg_launch_us": 1e3 * d_ms / max(d_launch, 1), "launches": d_launch,
                     "algorithmic_bytes_per_launch": d_bytes / max(d_launch, 1)},
        "phases_rank0": phases,
        "driver_phases_rank0": driver_phase_table(driver_phases, args.steps, dev_ms / args.steps),
        "memory": {"peak_torch_allocated_GiB_max_over_ranks": round(peak_gib, 2), "device_used_GiB_max_over_ranks": round(used_gib, 2),
                   "shard_GiB": round(8 * size / world / 2**30, 2)},
        "clocks": clocks,
        "energy_last_step": last[0],
    }
    if parity is not None:
        line["parity"] = parity
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = {"note": "reported by the N=1 run (rank 0, bounded sample); not repeated at N>1"}
    emit(line)
    dist.destroy_process_group()


def run_b200_batched(args, plan, ucc, theta, barrier, world, rank, n, staging_s):
    """Small-active-space regime (BASELINE configs 1-2): B random theta per step through tcc_energy_and_grad_batch."""
    import torch
    import torch.distributed as dist

    bsz = args.batch
    p_count, m, size = ucc.n_params, len(ucc.ex_ops), ucc.civector_size

    def thetas(step):
        return np.stack([theta(step * bsz + i) for i in range(bsz)])

    for w in range(args.warmup):
        plan.energy_and_grad_batch(thetas(-1 - w))
    barrier()
    launches0 = plan.launch_count
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    if rank == 0:
        sampler.start()
    sets = [thetas(k) for k in range(args.steps)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for k in range(args.steps):
        e_b, g_b = plan.energy_and_grad_batch(sets[k])
    ev1.record()
    barrier()
    wall_s = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = plan.launch_count - launches0
    t = torch.tensor([dev_ms, wall_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    evals = args.steps * bsz * world
    peak, peak_src = measured_peak_hbm()
    # whole-call roofline on the dense model: (48 M + 16) N bytes per evaluation
    dense = (48 * m + 16) * size * evals
    line = {
        "metric": METRIC, "value": evals / (dev_ms * 1e-3), "unit": "evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_desc(n, args.ansatz, args.integrals) + f", {bsz} random theta per step (batched)", "n_determinants": size,
                   "n_excitations": m, "n_params": p_count, "n_batches": plan.n_batches, "staging_s": round(staging_s, 2),
                   "l2": "vectors of all sets of a step: %.1f MB" % (4 * 8 * size * bsz / 1e6)},
        "e2e": {"value": evals / (wall_ms * 1e-3), "unit": "evals/s", "h2d_bytes_per_step": 8 * p_count * bsz,
                "d2h_bytes_per_step": 8 * (m + 1) * bsz,
                "note": "Plan.energy_and_grad_batch(thetas[B,P]) -> (E[B], grad[B,P]) with host buffers; the device-timed value spans the same calls"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "batch_kernel<2> (all sets in one launch per batch)", "achieved": dense / (dev_ms * 1e-3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": dense / (dev_ms * 1e-3) / 1e9 / peak, "peak_source": peak_src, "traffic": None,
                     "note": "effective bytes of the reference's dense model, (48M+16)N per evaluation; the vectors of small active spaces live in L2/shared memory, so this is not an HBM measurement"},
        "clocks": clocks, "energy_last": float(e_b[-1]),
    }
    if not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_reference_sample(n, n_ops_sample=2, ansatz=args.ansatz)
        except Exception as exc:  # pragma: no cover
            line["cpu_baseline"] = {"error": repr(exc)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="16o", choices=sorted(WORKLOADS))
    ap.add_argument("--ansatz", default="uccsd", choices=["uccsd", "kupccgsd", "puccd"],
                    help="ansatz of the workload (default: unscreened UCCSD, the configuration the metric is quoted on)")
    ap.add_argument("--integrals", default="random", choices=["random", "hchain"],
                    help="random_integral(n, seed 2077) used as MO integrals (default, SURVEY.md 8d) or the literal linear H_n chain, "
                         "STO-3G, 0.8 A, RHF orbitals (tencirchem_b200/hchain.py; the H4/H8/H12 molecules of BASELINE.json)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--max-ops", type=int, default=0,
                    help="truncate the ansatz to its first K excitations (demonstration runs only; 0 = full ansatz)")
    ap.add_argument("--sigma", default="auto", choices=["auto", "distributed", "replicated"],
                    help="N>1 sharded runs: H.c with the vector kept sharded, or on the all-reduced full vector")
    ap.add_argument("--multi", default="sharded", choices=["sharded", "replicas"],
                    help="N>1: 'sharded' = one evaluation with the CI vector split by alpha rows (strong scaling); "
                         "'replicas' = independent theta sets per rank (weak scaling)")
    ap.add_argument("--parity-golden", action="store_true",
                    help="evaluate the committed sparse-oracle case of the workload before the timed region and report the errors")
    ap.add_argument("--chunk-gib", type=float, default=4.0, help="N>1 sharded runs: staging budget of the row moves")
    ap.add_argument("--batch", type=int, default=0,
                    help="batched-theta mode: each step evaluates this many parameter sets in the same launches (small active spaces)")
    args = ap.parse_args()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # before torch touches the allocator: the shards of the sharded engine change size with the row layout
        os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
