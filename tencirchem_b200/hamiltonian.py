"""Hamiltonian handles of the B200 engine, mirroring tencirchem/static/hamiltonian.py for the
``fcifunc`` path (:146-196, :240-269).

The reference's handle is a plain closure and ``apply_op`` dispatches on ``inspect.isfunction``
(hamiltonian.py:249-250), so the handle here is also a real function; it carries the integrals as
attributes so the engine can upload the absorbed two-body array once per plan.
"""
from __future__ import annotations

from inspect import isfunction

import numpy as np

from .plan import Plan


def _make_handle(int1e, int2e, n_elec, hcb):
    int1e = np.ascontiguousarray(np.asarray(int1e, dtype=np.float64))
    n_orb = len(int1e)
    int2e = np.ascontiguousarray(np.asarray(int2e, dtype=np.float64))
    if int1e.shape != (n_orb, n_orb):
        raise ValueError(f"Invalid one-boby integral array shape: {int1e.shape}")
    if int2e.shape != (n_orb,) * 4:
        raise ValueError(f"Invalid two-boby integral array shape: {int2e.shape}; expected full s1 symmetry array")
    n_qubits = n_orb if hcb else 2 * n_orb
    state = {"plan": None}

    def fci_func(civector):
        if state["plan"] is None:
            from .evolve_civector import default_device

            plan = Plan(n_qubits, n_elec, (), (), hcb=hcb, device=default_device())
            plan.set_hamiltonian(int1e, int2e, token=fci_func)
            state["plan"] = plan
        return state["plan"].apply_h(civector)

    fci_func.int1e = int1e
    fci_func.int2e = int2e
    fci_func.n_elec = n_elec
    fci_func.hcb = hcb
    fci_func.n_qubit = n_qubits  # utils/misc.py:118-119 reads this attribute from function handles
    fci_func.is_b200_handle = True
    return fci_func


def get_h_fcifunc_from_integral(int1e, int2e, n_elec):
    """hamiltonian.py:146-155."""
    return _make_handle(int1e, int2e, n_elec, hcb=False)


def get_h_fcifunc_hcb_from_integral(int1e, int2e, n_elec):
    """hamiltonian.py:158-183."""
    return _make_handle(int1e, int2e, n_elec, hcb=True)


def get_h_from_integral(int1e, int2e, n_elec, hcb: bool, htype: str = "fcifunc"):
    """hamiltonian.py:186-196, ``fcifunc`` branch only (sparse / MPO forms belong to other engines)."""
    if htype.lower() != "fcifunc":
        raise ValueError(f"htype '{htype}' is not supported by the B200 civector engine; use 'fcifunc'")
    if hcb:
        return get_h_fcifunc_hcb_from_integral(int1e, int2e, n_elec)
    return get_h_fcifunc_from_integral(int1e, int2e, n_elec)


def apply_op(op, state):
    """hamiltonian.py:240-252 for function handles and matrices."""
    if isfunction(op):
        return op(state)
    return op @ state


def symmetrize_int2e(int2e):
    """hamiltonian.py:264-269."""
    int2e = 0.25 * (int2e + int2e.transpose((0, 1, 3, 2)) + int2e.transpose((1, 0, 2, 3)) + int2e.transpose((2, 3, 0, 1)))
    return 0.5 * (int2e + int2e.transpose(3, 2, 1, 0))


def random_integral(nao: int, seed: int = 2077):
    """hamiltonian.py:255-261: the PySCF-free input generator (legacy global RNG stream, as upstream)."""
    np.random.seed(seed)
    int1e = np.random.uniform(-1, 1, size=(nao, nao))
    int2e = np.random.uniform(-1, 1, size=(nao, nao, nao, nao))
    int1e = 0.5 * (int1e + int1e.T)
    return int1e, symmetrize_int2e(int2e)
