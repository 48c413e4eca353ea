"""The three engine functions of the ``civector`` path with the reference's signatures
(tencirchem/static/evolve_civector.py:180-243, :311-313), executed by the CUDA library.

Results are host numpy arrays, like the reference's numpy backend; ``Plan`` exposes the
device-resident variants.  Module-level plan cache = the reference's ``CI_OPERATOR_BATCH_CACHE``
(:109-110); ``clear_cache()`` is what ``tencirchem.clear_cache()`` (``__init__.py``:55-62) should call.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .plan import Plan

from collections import OrderedDict

_PLAN_CACHE: "OrderedDict" = OrderedDict()  # least recently used first
_MAX_CACHED = 8
# device-memory bound of the cache: tables plus the lazily allocated workspaces (up to six full vectors per plan)
_MAX_CACHED_BYTES = 48 << 30


def clear_cache():
    _PLAN_CACHE.clear()


def _plan_footprint(plan: Plan) -> int:
    return int(plan.device_bytes)


def default_device() -> int:
    """The GPU this process computes on: TCC_B200_DEVICE, else torch's current device when torch has already
    initialised CUDA (one process per GPU under torchrun), else device 0."""
    import os
    import sys

    env = os.environ.get("TCC_B200_DEVICE")
    if env is not None:
        return int(env)
    torch = sys.modules.get("torch")
    if torch is not None and torch.cuda.is_available() and torch.cuda.is_initialized():
        return int(torch.cuda.current_device())
    return 0


def get_plan(n_qubits, n_elec, ex_ops, param_ids, hcb=False, device=None) -> Plan:
    if device is None:
        device = default_device()
    ex_ops = tuple(tuple(op) for op in ex_ops)
    if param_ids is None:
        param_ids = range(len(ex_ops))
    key = (n_qubits, n_elec, ex_ops, tuple(param_ids), bool(hcb), device)
    plan = _PLAN_CACHE.get(key)
    if plan is not None:
        _PLAN_CACHE.move_to_end(key)  # LRU: a hit makes the plan the most recently used
        return plan
    # evict least recently used plans while the cache is over its count or device-memory bound (ADAPT-style loops
    # create a new key per iteration; every plan pins its tables and workspaces until it is dropped)
    while _PLAN_CACHE and (len(_PLAN_CACHE) >= _MAX_CACHED or sum(_plan_footprint(q) for q in _PLAN_CACHE.values()) > _MAX_CACHED_BYTES):
        _PLAN_CACHE.popitem(last=False)
    plan = Plan(n_qubits, n_elec, ex_ops, param_ids, hcb=hcb, device=device)
    _PLAN_CACHE[key] = plan
    return plan


def _bind_hamiltonian(plan: Plan, hamiltonian):
    if not getattr(hamiltonian, "is_b200_handle", False):
        raise ValueError(
            "the B200 engine needs a Hamiltonian handle from tencirchem_b200.hamiltonian.get_h_from_integral "
            "(it carries the integrals the GPU needs); got " + repr(type(hamiltonian))
        )
    if plan._h_token is not hamiltonian:
        if bool(hamiltonian.hcb) != plan.hcb:
            raise ValueError("Hamiltonian handle and ansatz disagree on hcb")
        plan.set_hamiltonian(hamiltonian.int1e, hamiltonian.int2e, token=hamiltonian)


def get_civector(params, n_qubits, n_elec, ex_ops, param_ids, hcb=False, init_state=None):
    """evolve_civector.py:180-198."""
    return get_plan(n_qubits, n_elec, ex_ops, param_ids, hcb).civector(params, init_state)


def get_energy_and_grad_civector(
    params, hamiltonian, n_qubits, n_elec, ex_ops: Tuple, param_ids: Tuple, hcb: bool = False, init_state=None
):
    """evolve_civector.py:201-218: (energy without e_core, gradient over parameters)."""
    plan = get_plan(n_qubits, n_elec, ex_ops, param_ids, hcb)
    _bind_hamiltonian(plan, hamiltonian)
    return plan.energy_and_grad(params, init_state)


def get_energy_civector(params, hamiltonian, n_qubits, n_elec, ex_ops, param_ids, hcb=False, init_state=None):
    """engine_ucc.py:79-89 fused on the device (state preparation, H.c and the dot product)."""
    plan = get_plan(n_qubits, n_elec, ex_ops, param_ids, hcb)
    _bind_hamiltonian(plan, hamiltonian)
    return plan.energy(params, init_state)


def apply_excitation_civector(civector, n_qubits, n_elec, f_idx, hcb):
    """evolve_civector.py:311-313."""
    plan = get_plan(n_qubits, n_elec, (), (), hcb)
    return plan.apply_excitation(np.asarray(civector, dtype=np.float64), tuple(f_idx))
