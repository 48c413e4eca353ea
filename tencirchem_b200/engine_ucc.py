"""Engine dispatch for the B200 engine, mirroring tencirchem/static/engine_ucc.py (:33-174).

``ENGINE_NAME`` starts with ``"civector"`` on purpose: the reference's two hard-coded gates
(ucc.py:591-592 ``engine.startswith("civector")`` and engine_ucc.py:62/:72) then treat it like the
stock civector engine and only the whitelist in ``UCC._check_engine`` (ucc.py:427-430) needs the
new name.  ``register()`` performs exactly that patch when TenCirChem is importable.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .ci_utils import civector_to_statevector, get_ci_strings, statevector_to_civector
from .evolve_civector import apply_excitation_civector, get_civector as _get_civector, get_energy_and_grad_civector, get_energy_civector

ENGINE_NAME = "civector-b200"

GETVECTOR_MAP = {ENGINE_NAME: _get_civector}
ENERGY_MAP = {ENGINE_NAME: get_energy_civector}
ENERGY_AND_GRAD_MAP = {ENGINE_NAME: get_energy_and_grad_civector}
APPLY_EXCITATION_MAP = {ENGINE_NAME: apply_excitation_civector}


def _check_engine(engine):
    if engine not in ENERGY_AND_GRAD_MAP:
        raise ValueError(f"Engine '{engine}' not supported")  # engine_ucc.py:120-121


def translate_init_state(init_state, n_qubits, ci_strings_fn):
    """engine_ucc.py:163-174 (array inputs; circuits belong to the gate-level engines)."""
    if init_state is None:
        return None
    if hasattr(init_state, "state") and callable(init_state.state):  # tc.Circuit duck type
        sv = np.asarray(init_state.state()).real
        return statevector_to_civector(sv, ci_strings_fn())
    init_state = np.asarray(init_state)
    if n_qubits <= 32 and len(init_state) == (1 << n_qubits):
        return statevector_to_civector(init_state, ci_strings_fn())
    return init_state


def get_civector(params, n_qubits, n_elec, ex_ops, param_ids, hcb, init_state, engine=ENGINE_NAME):
    """engine_ucc.py:59-66."""
    _check_engine(engine)
    if param_ids is None:
        param_ids = range(len(ex_ops))
    init_state = translate_init_state(init_state, n_qubits, lambda: get_ci_strings(n_qubits, n_elec, hcb))
    return GETVECTOR_MAP[engine](
        params, n_qubits=n_qubits, n_elec=n_elec, ex_ops=tuple(ex_ops), param_ids=tuple(param_ids), hcb=hcb, init_state=init_state
    )


def get_statevector(params, n_qubits, n_elec, ex_ops, param_ids, hcb, init_state, engine=ENGINE_NAME):
    """engine_ucc.py:69-76."""
    ket = get_civector(params, n_qubits, n_elec, ex_ops, param_ids, hcb, init_state, engine)
    return civector_to_statevector(ket, n_qubits, get_ci_strings(n_qubits, n_elec, hcb))


def get_energy(params, hamiltonian, n_qubits, n_elec, ex_ops: Tuple, param_ids: Tuple, hcb: bool, init_state, engine=ENGINE_NAME):
    """engine_ucc.py:79-89."""
    _check_engine(engine)
    if param_ids is None:
        param_ids = range(len(ex_ops))
    init_state = translate_init_state(init_state, n_qubits, lambda: get_ci_strings(n_qubits, n_elec, hcb))
    return ENERGY_MAP[engine](params, hamiltonian, n_qubits, n_elec, tuple(ex_ops), tuple(param_ids), hcb, init_state)


def get_energy_and_grad(params, hamiltonian, n_qubits, n_elec, ex_ops, param_ids, hcb, init_state, engine=ENGINE_NAME):
    """engine_ucc.py:119-126."""
    _check_engine(engine)
    if param_ids is None:
        param_ids = range(len(ex_ops))
    init_state = translate_init_state(init_state, n_qubits, lambda: get_ci_strings(n_qubits, n_elec, hcb))
    return ENERGY_AND_GRAD_MAP[engine](params, hamiltonian, n_qubits, n_elec, tuple(ex_ops), tuple(param_ids), hcb, init_state)


def apply_excitation(state, n_qubits, n_elec, ex_op, hcb, engine=ENGINE_NAME):
    """engine_ucc.py:139-160: accepts a CI vector or a full statevector and returns the same kind."""
    if engine not in APPLY_EXCITATION_MAP:
        raise ValueError(f"Engine '{engine}' not supported")
    state = np.asarray(state, dtype=np.float64)
    is_statevector_input = n_qubits <= 32 and len(state) == (1 << n_qubits)
    ci_strings = None
    if is_statevector_input:
        ci_strings = get_ci_strings(n_qubits, n_elec, hcb)
        state = statevector_to_civector(state, ci_strings)
    res = APPLY_EXCITATION_MAP[engine](state, n_qubits, n_elec, tuple(ex_op), hcb)
    if is_statevector_input:
        return civector_to_statevector(res, n_qubits, ci_strings)
    return res


def register():
    """Insert the engine into an importable TenCirChem (the maintainer-side patch of INTEGRATION.md).
    Returns True when the reference package was found and patched."""
    try:
        from tencirchem.static import engine_ucc as ref_engine, ucc as ref_ucc
        import tencirchem as ref_pkg
    except Exception:
        return False
    from . import hamiltonian as b200_h
    from .evolve_civector import clear_cache

    ref_engine.GETVECTOR_MAP[ENGINE_NAME] = _get_civector
    ref_engine.ENERGY_AND_GRAD_MAP[ENGINE_NAME] = get_energy_and_grad_civector
    ref_engine.APPLY_EXCITATION_MAP[ENGINE_NAME] = apply_excitation_civector

    orig_check = ref_ucc.UCC._check_engine

    def _check_engine_patched(self, engine):
        if engine == ENGINE_NAME:
            return
        return orig_check(self, engine)

    ref_ucc.UCC._check_engine = _check_engine_patched

    orig_get_h = ref_ucc.UCC._get_hamiltonian_and_core

    def _get_hamiltonian_and_core_patched(self, engine):
        # same contract as ucc.py:584-606: returns (hamiltonian, e_core, engine)
        self._check_engine(engine)
        effective = self.engine if engine is None else engine
        if effective != ENGINE_NAME:
            return orig_get_h(self, engine)
        key = "fcifunc-b200"
        hamiltonian = self.hamiltonian_lib.get(key)
        if hamiltonian is None:
            if self.int1e is None:
                from tencirchem.static.hamiltonian import get_integral_from_hf

                self.int1e, self.int2e, e_core = get_integral_from_hf(self.hf, self.active_space)
            else:
                e_core = self.e_core
            hamiltonian = b200_h.get_h_from_integral(self.int1e, self.int2e, self.n_elec, self.hcb, "fcifunc")
            self.hamiltonian_lib[key] = hamiltonian
        else:
            e_core = self.e_core
        return hamiltonian, e_core, effective

    ref_ucc.UCC._get_hamiltonian_and_core = _get_hamiltonian_and_core_patched

    orig_clear = ref_pkg.clear_cache

    def _clear_cache_patched():
        orig_clear()
        clear_cache()

    ref_pkg.clear_cache = _clear_cache_patched
    return True
