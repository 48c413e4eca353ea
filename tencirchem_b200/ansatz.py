"""Thin ansatz classes over the B200 engine with the call surface of the reference's
``UCC / UCCSD / KUPCCGSD / PUCCD`` for THIS path only: ``from_integral``, ``energy``,
``energy_and_grad``, ``civector``, ``apply_excitation``, ``kernel``, ``get_ex_ops`` and the size
properties (tencirchem/static/ucc.py:52-125, :337-378, :440-479, :608-737, :1239-1290).

The reference classes need PySCF for molecules, HF/MP2/CCSD references and RDMs; none of that is
on the hot path, so these classes start from MO integrals (the reference's own ``from_integral``
route, ucc.py:52-125) and leave the rest to TenCirChem itself -- when TenCirChem is installed, use
its classes with ``engine="civector-b200"`` after ``tencirchem_b200.register()``.
"""
from __future__ import annotations

import logging
import threading
from math import comb
from time import time
from typing import List, Sequence, Tuple

import numpy as np
from scipy.optimize import OptimizeResult as _OptResult

logger = logging.getLogger(__name__)

from . import engine_ucc
from .ci_utils import get_ci_strings
from .hamiltonian import get_h_from_integral


def _uccsd_singles(no: int, nv: int):
    """UCC.get_ex1_ops (ucc.py:889-936): alpha then beta copy of each i->a, one shared parameter."""
    for i in range(no):
        for a in range(nv):
            yield ((2 * no + nv + a, no + nv + i), (no + a, i))


def _uccsd_doubles(no: int, nv: int):
    """UCC.get_ex2_ops (ucc.py:938-1031): yields groups of tuples that share one parameter."""
    n = no + nv
    occ_b, vir_b = (lambda i: i), (lambda a: no + a)
    occ_a, vir_a = (lambda i: n + i), (lambda a: n + no + a)
    for i in range(no):
        for j in range(i):
            for a in range(nv):
                for b in range(a):
                    yield ((vir_a(b), vir_a(a), occ_a(i), occ_a(j)), (vir_b(b), vir_b(a), occ_b(i), occ_b(j)))
    for i in range(no):
        for j in range(i + 1):
            for a in range(nv):
                for b in range(a + 1):
                    if i == j and a == b:
                        yield ((vir_b(a), vir_a(a), occ_a(i), occ_b(i)),)
                        continue
                    yield ((vir_b(b), vir_a(a), occ_a(i), occ_b(j)), (vir_a(b), vir_b(a), occ_b(i), occ_a(j)))
                    if i != j and a != b:
                        yield ((vir_b(a), vir_a(b), occ_a(i), occ_b(j)), (vir_a(a), vir_b(b), occ_b(i), occ_a(j)))


def _flatten(groups, first_id=0):
    ops, ids = [], []
    pid = first_id
    for group in groups:
        for op in group:
            ops.append(op)
            ids.append(pid)
        pid += 1
    return ops, ids


class UCC:
    """Base class: integrals + an ordered excitation list (ucc.py:158-335 minus the PySCF parts)."""

    def __init__(self, int1e, int2e, n_elec: int, e_core: float = 0.0, hcb: bool = False, engine: str = None,
                 ex_ops: Sequence[Tuple] = None, param_ids: Sequence[int] = None, init_state=None):
        self.int1e = np.asarray(int1e, dtype=np.float64)
        self.int2e = np.asarray(int2e, dtype=np.float64)
        self.n_elec = int(n_elec)
        if self.n_elec % 2:
            raise ValueError("UCC class does not support open shell system")  # ucc.py:207-208
        self.e_core = float(e_core)
        self.hcb = bool(hcb)
        self._check_engine(engine)
        self.engine = engine or engine_ucc.ENGINE_NAME
        self.active = len(self.int1e)
        self.n_qubits = self.active if self.hcb else 2 * self.active
        self.hamiltonian = get_h_from_integral(self.int1e, self.int2e, self.n_elec, self.hcb, "fcifunc")
        self.ex_ops = list(ex_ops) if ex_ops is not None else None
        self._param_ids = list(param_ids) if param_ids is not None else None
        self.init_guess = None
        self.init_state = init_state
        self.scipy_minimize_options = None
        self.opt_res = None
        self.params = None
        self._sharded = None

    def shard(self, comm, device: int = 0, **kwargs):
        """Spread the CI vector of this ansatz over the ranks of `comm` (tencirchem_b200.sharded: alpha-row sharding,
        one process per GPU with `TorchComm()`, or `LocalComm(n)` for n virtual ranks on one GPU).  Afterwards
        `civector`, `energy`, `energy_and_grad` and `kernel` run on the sharded engine; every rank gets the same
        results, so the L-BFGS-B drivers of all ranks take the same path."""
        from .sharded import ShardedUCC

        self._sanity_check()
        if self.hcb:
            raise ValueError("hcb (pUCCD) vectors have a single alpha row and are not sharded")
        if self.init_state is not None:
            raise ValueError("the sharded engine starts from the HF determinant; init_state is not supported")
        self._sharded = ShardedUCC(self.n_qubits, self.n_elec, self.ex_ops, self.param_ids, comm, device=device, **kwargs)
        self._sharded.set_hamiltonian(self.int1e, self.int2e)
        return self

    @classmethod
    def from_integral(cls, int1e, int2e, n_elec, e_core: float = 0, **kwargs):
        """ucc.py:52-125."""
        return cls(int1e, int2e, n_elec, e_core, **kwargs)

    # ---- sizes (ucc.py:1239-1290) ----
    @property
    def no(self) -> int:
        return self.n_elec // 2

    @property
    def nv(self) -> int:
        return self.active - self.no

    @property
    def param_ids(self) -> List[int]:
        if self._param_ids is None and self.ex_ops is not None:
            return list(range(len(self.ex_ops)))
        return self._param_ids

    @param_ids.setter
    def param_ids(self, v):
        self._param_ids = v

    @property
    def n_params(self) -> int:
        if not self.param_ids:
            return 0
        return max(self.param_ids) + 1

    @property
    def civector_size(self) -> int:
        if not self.hcb:
            return comb(self.n_qubits // 2, self.n_elec // 2) ** 2
        return comb(self.n_qubits, self.n_elec // 2)

    @property
    def statevector_size(self) -> int:
        return 1 << self.n_qubits

    def get_ci_strings(self, strs2addr: bool = False):
        """ucc.py:481-510."""
        return get_ci_strings(self.n_qubits, self.n_elec, self.hcb, strs2addr=strs2addr)

    # ---- argument checks (ucc.py:410-438) ----
    def _check_engine(self, engine):
        if engine not in (None, engine_ucc.ENGINE_NAME):
            raise ValueError(f"Engine '{engine}' not supported")

    def _sanity_check(self):
        if self.ex_ops is None or self.param_ids is None:
            raise ValueError("`ex_ops` or `param_ids` not defined")
        if len(self.ex_ops) != len(self.param_ids):
            raise ValueError(
                f"Excitation operator size {len(self.ex_ops)} and parameter size {len(self.param_ids)} do not match"
            )

    def _check_params_argument(self, params, strict=True):
        if params is None:
            if self.params is not None:
                params = self.params
            elif strict:
                raise ValueError("Run the `.kernel` method to determine the parameters first")
            elif self.init_guess is not None:
                params = self.init_guess
            else:
                params = np.zeros(self.n_params)
        if len(params) != self.n_params:
            raise ValueError(f"Incompatible parameter shape. {self.n_params} is desired. Got {len(params)}")
        return np.asarray(params, dtype=np.float64)

    # ---- the hot path ----
    def civector(self, params=None, engine: str = None):
        """ucc.py:440-479."""
        self._sanity_check()
        params = self._check_params_argument(params)
        self._check_engine(engine)
        if self._sharded is not None:
            return self._sharded.civector(params)
        return engine_ucc.get_civector(params, self.n_qubits, self.n_elec, self.ex_ops, self.param_ids, self.hcb, self.init_state)

    def energy(self, params=None, engine: str = None) -> float:
        """ucc.py:608-655."""
        self._sanity_check()
        params = self._check_params_argument(params)
        self._check_engine(engine)
        if params is self.params and self.opt_res is not None:
            return self.opt_res["e"]
        if self._sharded is not None:
            return float(self._sharded.energy(params)) + self.e_core
        e = engine_ucc.get_energy(params, self.hamiltonian, self.n_qubits, self.n_elec, self.ex_ops, self.param_ids, self.hcb, self.init_state)
        return float(e) + self.e_core

    def energy_and_grad(self, params=None, engine: str = None):
        """ucc.py:657-707."""
        self._sanity_check()
        params = self._check_params_argument(params)
        self._check_engine(engine)
        if self._sharded is not None:
            e, g = self._sharded.energy_and_grad(params)
            return float(e + self.e_core), np.asarray(g)
        e, g = engine_ucc.get_energy_and_grad(
            params, self.hamiltonian, self.n_qubits, self.n_elec, self.ex_ops, self.param_ids, self.hcb, self.init_state
        )
        return float(e + self.e_core), np.asarray(g)

    def energy_and_grad_batch(self, params_batch, engine: str = None):
        """Additive API (gap i): many parameter vectors ``[n_sets, n_params]`` in the same
        kernel launches; returns ``(energies [n_sets], grads [n_sets, n_params])``.  Starts from the HF determinant."""
        from .evolve_civector import _bind_hamiltonian, get_plan

        self._sanity_check()
        self._check_engine(engine)
        if self.init_state is not None:
            raise ValueError("energy_and_grad_batch starts from the HF determinant; init_state is not supported")
        params_batch = np.asarray(params_batch, dtype=np.float64)
        if params_batch.ndim != 2 or params_batch.shape[1] != self.n_params:
            raise ValueError(f"Incompatible parameter shape. (n_sets, {self.n_params}) is desired. Got {params_batch.shape}")
        plan = get_plan(self.n_qubits, self.n_elec, self.ex_ops, self.param_ids, self.hcb)
        _bind_hamiltonian(plan, self.hamiltonian)
        e, g = plan.energy_and_grad_batch(params_batch)
        return e + self.e_core, g

    def apply_excitation(self, state, ex_op: Tuple, engine: str = None):
        """ucc.py:709-737."""
        self._check_engine(engine)
        return engine_ucc.apply_excitation(state, self.n_qubits, self.n_elec, ex_op, hcb=self.hcb)

    def pool_gradients(self, pool: Sequence[Sequence[Tuple[int, ...]]], params=None, engine: str = None) -> np.ndarray:
        """ADAPT-VQE screening (adapt_vqe.ipynb cell 9): for every entry of `pool` (a list of excitation tuples that
        share one parameter) the energy gradient 2 * sum_op <H psi| G_op |psi> at the current state psi = civector(params)
        -- one batched launch set for the WHOLE pool (tcc_pool_gradients) instead of one apply_excitation + dot product
        per operator.  psi and H psi never leave the GPU."""
        import torch

        from .evolve_civector import _bind_hamiltonian, get_plan

        self._sanity_check()
        self._check_engine(engine)
        flat = [tuple(int(i) for i in op) for entry in pool for op in entry]
        owner = [k for k, entry in enumerate(pool) for _ in entry]
        if not flat:
            return np.zeros(len(pool))
        params = self._check_params_argument(params)
        state_plan = get_plan(self.n_qubits, self.n_elec, self.ex_ops, self.param_ids, self.hcb)
        _bind_hamiltonian(state_plan, self.hamiltonian)
        psi = torch.empty(state_plan.size, dtype=torch.float64, device=f"cuda:{state_plan.device}")
        init = None
        if self.init_state is not None:
            init = torch.as_tensor(engine_ucc.translate_init_state(self.init_state, self.n_qubits, lambda: get_ci_strings(self.n_qubits, self.n_elec, self.hcb)),
                                   dtype=torch.float64, device=psi.device)
        state_plan.civector_dev(params, psi, init)
        hpsi = torch.empty_like(psi)
        state_plan.apply_h_dev(psi, hpsi)
        pool_plan = get_plan(self.n_qubits, self.n_elec, tuple(flat), tuple(range(len(flat))), self.hcb)
        per_op = pool_plan.pool_gradients_dev(psi, hpsi)
        out = np.zeros(len(pool))
        np.add.at(out, owner, 2.0 * per_op)
        return out

    # ---- reduced density matrices (ucc.py:755-852), MO basis of the integrals the ansatz was built from ----
    def _rdm12(self, statevector=None):
        import torch

        from .evolve_civector import default_device, get_plan

        if self.hcb:
            raise AssertionError("RDMs are defined for the fermionic CI vector (ucc.py:786 asserts `not self.hcb`)")
        civector = self.civector() if statevector is None else np.asarray(statevector, dtype=np.float64)
        if len(civector) == self.statevector_size and self.statevector_size != self.civector_size:
            civector = civector[self.get_ci_strings()]  # ucc.py:743-745
        if len(civector) != self.civector_size:
            raise ValueError(f"Incompatible statevector size: {len(civector)}")
        plan = get_plan(self.n_qubits, self.n_elec, self.ex_ops or (), self.param_ids or (), self.hcb)
        c_dev = torch.as_tensor(np.ascontiguousarray(civector), dtype=torch.float64, device=torch.device("cuda", default_device()))
        return plan.make_rdm12_dev(c_dev)

    def make_rdm1(self, statevector=None, basis: str = "MO") -> np.ndarray:
        """Spin-traced 1RDM, rdm1[p,q] = <p_a^+ q_a> + <p_b^+ q_b> (ucc.py:755-797), computed on the GPU.  Only the MO
        basis of the supplied integrals exists here (the reference's default "AO" needs its PySCF mean-field object)."""
        if basis != "MO":
            raise ValueError("this build works from MO integrals: only basis='MO' is available")
        return self._rdm12(statevector)[0]

    def make_rdm2(self, statevector=None, basis: str = "MO") -> np.ndarray:
        """Spin-traced 2RDM in PySCF's convention rdm2[p,q,r,s] = <p^+ r^+ s q> (ucc.py:799-852), on the GPU."""
        if basis != "MO":
            raise ValueError("this build works from MO integrals: only basis='MO' is available")
        return self._rdm12(statevector)[1]

    @property
    def e_ucc(self) -> float:
        """ucc.py:1033-1038."""
        return self.energy()

    @property
    def e_fci(self) -> float:
        """FCI energy of the active space (ucc.py:284-300 takes it from PySCF): Lanczos on the engine's H.c, cached."""
        if getattr(self, "_e_fci", None) is None:
            from .fci import fci_energy

            self._e_fci = fci_energy(self)
        return self._e_fci

    def _minimize_options(self) -> dict:
        """ucc.py:355-360: the reference's strict L-BFGS-B defaults (ftol = 10 eps, gtol = 100 eps of float64)
        unless ``scipy_minimize_options`` is set."""
        if self.scipy_minimize_options is None:
            eps = np.finfo(np.float64).eps
            return {"ftol": 1e1 * eps, "gtol": 1e2 * eps}
        return self.scipy_minimize_options

    def kernel(self) -> float:
        """ucc.py:337-378: L-BFGS-B over ``energy_and_grad``."""
        from scipy.optimize import minimize

        self._sanity_check()
        if self.init_guess is None:
            self.init_guess = np.zeros(self.n_params)
        if self.n_params == 0:
            e = self.energy(np.zeros(0))
            self.opt_res = _OptResult(e=e, fun=e, x=np.zeros(0), success=True, init_guess=self.init_guess)
            self.params = self.opt_res["x"]
            return e
        t0 = time()
        func = lambda x: tuple(np.asarray(v, dtype=np.float64) for v in self.energy_and_grad(x))
        res = minimize(func, x0=np.asarray(self.init_guess, dtype=np.float64), jac=True, method="L-BFGS-B",
                       options=self._minimize_options())
        if not res.success:
            logger.warning("Optimization failed. See `.opt_res` for details.")  # ucc.py:367-368
        opt_res = _OptResult(res)
        opt_res["opt_time"] = time() - t0
        opt_res["init_guess"] = self.init_guess
        opt_res["e"] = float(res.fun)
        self.opt_res = opt_res
        self.params = res.x.copy()
        return opt_res["e"]


DISCARD_EPS = 1e-12  # tencirchem/constants.py: two-body excitations with a smaller initial amplitude are dropped


def _spin_t2(t2, i, j, a, b, kind):
    """Element of ``spatial2spin(t2)`` (PySCF cc.addons, even = alpha, odd = beta) that ucc.py:938-1031 reads for an
    excitation of `kind`: "same" -> t2s[2i, 2j, 2a, 2b] (antisymmetrised), "ab" -> t2s[2i, 2j+1, 2a, 2b+1]."""
    if kind == "same":
        return t2[i, j, a, b] - t2[i, j, b, a]
    return t2[i, j, a, b]


class UCCSD(UCC):
    """uccsd.py:20-185.  Without amplitudes (``t2=None``, the default of :meth:`from_integral`, whose integrals need
    not come from a Hartree-Fock calculation) every excitation is kept in generator order with a zero initial guess.
    With MP2 amplitudes (``t2`` in PySCF's spatial ``[no, no, nv, nv]`` layout, e.g. from
    :func:`tencirchem_b200.hchain.mp2_t2`; ``init_method="mp2"`` in the reference) the two-body excitations are
    screened (``pick_ex2``: ``|t2| < epsilon`` dropped) and sorted by ``|t2|`` (``sort_ex2``) exactly as
    ``UCCSD.pick_and_sort`` does (uccsd.py:159-178), and the amplitudes are the initial guess."""

    def __init__(self, int1e, int2e, n_elec, e_core=0.0, t1=None, t2=None, pick_ex2: bool = True, sort_ex2: bool = True,
                 epsilon: float = DISCARD_EPS, **kwargs):
        super().__init__(int1e, int2e, n_elec, e_core, hcb=False, **kwargs)
        self.t1 = None if t1 is None else np.asarray(t1, dtype=np.float64)
        self.t2 = None if t2 is None else np.asarray(t2, dtype=np.float64)
        if self.t2 is not None and self.t2.shape != (self.no, self.no, self.nv, self.nv):
            raise ValueError(f"t2 must have the spatial shape {(self.no, self.no, self.nv, self.nv)}")
        self.pick_ex2 = bool(pick_ex2) and self.t2 is not None
        self.sort_ex2 = bool(sort_ex2) and self.t2 is not None
        self.t2_discard_eps = float(epsilon)
        self.ex_ops, self.param_ids, self.init_guess = self.get_ex_ops()

    def get_ex1_ops(self):
        """ucc.py:889-936; initial guess = t1[i, a] (zeros without CCSD amplitudes)."""
        ops, ids = _flatten(_uccsd_singles(self.no, self.nv))
        guess = [0.0] * (max(ids) + 1 if ids else 0)
        if self.t1 is not None:
            guess = [float(self.t1[i, a]) for i in range(self.no) for a in range(self.nv)]
        return ops, ids, guess

    def get_ex2_ops(self):
        """ucc.py:938-1031; the initial guess of each parameter is the spin-orbital amplitude the reference reads."""
        ops, ids = _flatten(_uccsd_doubles(self.no, self.nv))
        n_par = max(ids) + 1 if ids else 0
        if self.t2 is None:
            return ops, ids, [0.0] * n_par
        no, nv, t2 = self.no, self.nv, self.t2
        guess = []
        for i in range(no):
            for j in range(i):
                for a in range(nv):
                    for b in range(a):
                        guess.append(float(_spin_t2(t2, i, j, a, b, "same")))
        for i in range(no):
            for j in range(i + 1):
                for a in range(nv):
                    for b in range(a + 1):
                        guess.append(float(_spin_t2(t2, i, j, a, b, "ab")))
                        if not (i == j and a == b) and i != j and a != b:
                            guess.append(float(_spin_t2(t2, i, j, b, a, "ab")))
        assert len(guess) == n_par
        return ops, ids, guess

    def pick_and_sort(self, ex_ops, param_ids, init_guess, do_pick=True, do_sort=True):
        """uccsd.py:159-178 (stable sort by decreasing |amplitude|, drop below epsilon, renumber the parameters)."""
        pairs = list(zip(ex_ops, param_ids))
        if do_sort:
            pairs = sorted(pairs, key=lambda x: -abs(init_guess[x[1]]))
        kept = [(op, pid) for op, pid in pairs if not (do_pick and abs(init_guess[pid]) < self.t2_discard_eps)]
        unique_ids = sorted({pid for _, pid in kept})
        mapping = {old: new for new, old in enumerate(unique_ids)}
        return [op for op, _ in kept], [mapping[pid] for _, pid in kept], [init_guess[i] for i in unique_ids]

    def get_ex_ops(self):
        """uccsd.py:107-157."""
        o1, p1, g1 = self.get_ex1_ops()
        o2, p2, g2 = self.get_ex2_ops()
        if self.pick_ex2 or self.sort_ex2:
            o2, p2, g2 = self.pick_and_sort(o2, p2, g2, self.pick_ex2, self.sort_ex2)
        off = max(p1) + 1 if p1 else 0
        return o1 + o2, p1 + [p + off for p in p2], g1 + g2


class KUPCCGSD(UCC):
    """kupccgsd.py:20-267: k layers of generalised paired doubles followed by generalised singles."""

    def __init__(self, int1e, int2e, n_elec, e_core=0.0, k: int = 3, n_tries: int = 1, **kwargs):
        super().__init__(int1e, int2e, n_elec, e_core, hcb=False, **kwargs)
        self.k = k
        self.n_tries = n_tries
        self.ex_ops, self.param_ids, self.init_guess = self.get_ex_ops()
        # kupccgsd.py:91-96
        self.init_guess_list = [self.init_guess]
        for _ in range(self.n_tries - 1):
            self.init_guess_list.append(np.random.rand(self.n_params) - 0.5)
        self.e_tries_list = []
        self.opt_res_list = []
        self.staging_time = self.opt_time = None

    def get_ex1_ops(self):
        n = self.no + self.nv
        groups = (((n + a, n + i), (a, i)) for a in range(n) for i in range(a))
        ops, ids = _flatten(groups)
        return ops, ids, np.zeros(max(ids) + 1 if ids else 0)

    def get_ex2_ops(self):
        n = self.no + self.nv
        groups = (((a, n + a, n + i, i),) for a in range(n) for i in range(a))
        ops, ids = _flatten(groups)
        return ops, ids, np.zeros(max(ids) + 1 if ids else 0)

    def get_ex_ops(self):
        """kupccgsd.py:127-177."""
        o1, p1, _ = self.get_ex1_ops()
        o2, p2, _ = self.get_ex2_ops()
        ops, ids, nxt = [], [], 0
        for _ in range(self.k):
            for block_ops, block_ids in ((o2, p2), (o1, p1)):
                ops += block_ops
                ids += [i + nxt for i in block_ids]
                nxt = ids[-1] + 1 if ids else 0
        return ops, ids, np.random.rand(max(ids) + 1 if ids else 0) - 0.5

    def kernel(self):
        """kupccgsd.py:98-125: best of ``n_tries`` starts (``init_guess_list``).  With n_tries > 1 the L-BFGS-B runs
        advance in LOCK STEP: every iteration evaluates the current point of all unfinished tries in one batched call
        (``energy_and_grad_batch`` -> tcc_energy_and_grad_batch: all parameter sets share every kernel launch) instead
        of the reference's serial loop of restarts."""
        self._sanity_check()
        t0 = time()
        if self.n_tries == 1:
            if not np.allclose(self.init_guess, self.init_guess_list[0]):
                logger.info("Inconsistent `self.init_guess` and `self.init_guess_list`.  Use `self.init_guess`.")
            super().kernel()
            results = [self.opt_res]
        else:
            results = lockstep_minimize(self.energy_and_grad_batch, self.init_guess_list[: self.n_tries], self._minimize_options())
        self.opt_res_list = sorted(results, key=lambda r: r.fun)
        self.e_tries_list = [float(r.fun) for r in self.opt_res_list]
        self.opt_time = time() - t0
        self.opt_res = self.opt_res_list[0]
        self.opt_res["e"] = float(self.opt_res.fun)
        self.opt_res["e_tries"] = self.e_tries_list
        if not self.opt_res.success:
            logger.warning("Optimization failed. See `.opt_res` for details.")
        self.params = np.array(self.opt_res.x, dtype=np.float64)
        self.init_guess = self.opt_res.init_guess
        return float(self.opt_res.e)


def lockstep_minimize(fun_batch, x0_list, options=None, method="L-BFGS-B"):
    """Runs one SciPy L-BFGS-B per starting point, all in lock step: each optimiser lives in its own thread and blocks
    in its objective call until every unfinished optimiser has asked for a point; the points are then evaluated by ONE
    call ``fun_batch(thetas[B, P]) -> (energies[B], grads[B, P])``.  Every try follows exactly the iterates of a serial
    run (same algorithm, same options); only the evaluation is shared.  Returns the list of OptimizeResult (with
    ``init_guess`` and ``e``), in the order of ``x0_list``."""
    from scipy.optimize import minimize

    n = len(x0_list)
    cond = threading.Condition()
    pending, answers, errors = {}, {}, []
    state = {"active": n, "batches": 0, "evals": 0}

    def flush_locked():
        ids = sorted(pending)
        try:
            e, g = fun_batch(np.stack([pending[i] for i in ids]))
            for k, i in enumerate(ids):
                answers[i] = (float(e[k]), np.array(g[k], dtype=np.float64))
        except BaseException as exc:  # hand the failure to every waiting optimiser
            errors.append(exc)
            for i in ids:
                answers[i] = None
        state["batches"] += 1
        state["evals"] += len(ids)
        pending.clear()
        cond.notify_all()

    def objective(i, x):
        with cond:
            pending[i] = np.array(x, dtype=np.float64)
            if len(pending) == state["active"]:
                flush_locked()
            else:
                cond.wait_for(lambda: i in answers)
            ans = answers.pop(i)
        if ans is None:
            raise RuntimeError("batched evaluation failed") from errors[0]
        return ans

    results = [None] * n

    def run(i):
        try:
            res = minimize(lambda x: objective(i, x), x0=np.asarray(x0_list[i], dtype=np.float64), jac=True, method=method,
                           options=options)
            res["init_guess"] = x0_list[i]
            res["e"] = float(res.fun)
            results[i] = res
        except BaseException as exc:
            errors.append(exc)
        finally:
            with cond:
                state["active"] -= 1
                if pending and len(pending) == state["active"]:
                    flush_locked()

    threads = [threading.Thread(target=run, args=(i,), daemon=True) for i in range(n)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    if errors:
        raise errors[0]
    for r in results:
        r["lockstep_batches"] = state["batches"]
        r["lockstep_evals"] = state["evals"]
    return results


class PUCCD(UCC):
    """puccd.py:20-236: paired doubles on the hard-core-boson representation."""

    def __init__(self, int1e, int2e, n_elec, e_core=0.0, **kwargs):
        super().__init__(int1e, int2e, n_elec, e_core, hcb=True, **kwargs)
        self.ex_ops, self.param_ids, self.init_guess = self.get_ex_ops()

    def get_ex_ops(self):
        """puccd.py:101-146."""
        ops = [(self.no + a, i) for i in range(self.no) for a in range(self.nv - 1, -1, -1)]
        return ops, list(range(len(ops))), [0.0] * len(ops)

    # ---- reduced density matrices on the hard-core-boson vector (puccd.py:148-203), MO basis ----
    def _rdm_hcb(self, statevector=None):
        import torch

        from .evolve_civector import default_device, get_plan

        civector = self.civector() if statevector is None else np.asarray(statevector, dtype=np.float64)
        if len(civector) == self.statevector_size and self.statevector_size != self.civector_size:
            civector = civector[self.get_ci_strings()]  # ucc.py:743-745
        if len(civector) != self.civector_size:
            raise ValueError(f"Incompatible statevector size: {len(civector)}")
        plan = get_plan(self.n_qubits, self.n_elec, self.ex_ops or (), self.param_ids or (), self.hcb)
        c_dev = torch.as_tensor(np.ascontiguousarray(civector), dtype=torch.float64, device=torch.device("cuda", default_device()))
        return plan.make_rdm12_hcb_dev(c_dev)

    def make_rdm1(self, statevector=None, basis: str = "MO") -> np.ndarray:
        """puccd.py:148-165: diagonal, rdm1[i, i] = 2 <n_i>; the occupations are reduced on the GPU."""
        if basis != "MO":
            raise ValueError("this build works from MO integrals: only basis='MO' is available")
        occ, _, _ = self._rdm_hcb(statevector)
        return np.diag(2.0 * occ)

    def make_rdm2(self, statevector=None, basis: str = "MO") -> np.ndarray:
        """puccd.py:167-203: the pair-hop and pair-pair elements come from one GPU reduction per (p, q); the
        assembly into [n, n, n, n] (with the reference's final factor 2) is index bookkeeping on the host."""
        if basis != "MO":
            raise ValueError("this build works from MO integrals: only basis='MO' is available")
        occ, hop, both = self._rdm_hcb(statevector)
        n = self.n_qubits
        rdm2 = np.zeros((n, n, n, n))
        for p in range(n):
            rdm2[p, p, p, p] = occ[p]
            for q in range(p):
                rdm2[p, q, p, q] = rdm2[q, p, q, p] = hop[p, q]
                rdm2[p, p, q, q] = rdm2[q, q, p, p] = 2 * both[p, q]
                rdm2[p, q, q, p] = rdm2[q, p, p, q] = -both[p, q]
        return 2.0 * rdm2

    @property
    def e_puccd(self) -> float:
        """puccd.py:231-236."""
        return self.energy()
