"""Boundary proof for ``tencirchem_b200.register()``: TenCirChem itself cannot be imported here
(PySCF / tensorcircuit are absent), so the test builds a STUB ``tencirchem`` package with the reference's dispatch
surface for this path -- the three engine maps (tencirchem/static/engine_ucc.py:33-39, :110-116, :129-136), the
dispatch functions that index them (:42-160), ``UCC._check_engine`` (ucc.py:427-430), ``UCC._get_hamiltonian_and_core``
with the control flow of ucc.py:584-606 INCLUDING the constructor's call (ucc.py:305-308: ``hamiltonian_lib = {}``,
``int1e = int2e = None``, then ``_get_hamiltonian_and_core(self.engine)``), ``apply_op`` (hamiltonian.py:240-252) and
``tencirchem.clear_cache`` (__init__.py:55-62) -- whose stock "civector" engine is the CPU oracle.  ``register()`` is
then executed against it and the patched maps are driven exactly as ``UCC.energy/energy_and_grad/civector/
apply_excitation`` drive them upstream.

CPU part: the patch lands, unknown names still raise the reference's ValueError, and a call with the new engine name
reaches the product code (which fails loudly without a GPU -- no CPU fallback).  GPU part: same class, same call,
``engine="civector"`` (oracle) vs ``engine="civector-b200"`` (CUDA) agree within the north_star tolerances.
"""
import importlib
import os
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

STUB = {
    "tencirchem/__init__.py": """
        CLEARED = []
        def clear_cache():
            from tencirchem.static import evolve_civector
            evolve_civector.clear()
            CLEARED.append(1)
        from tencirchem.static.ucc import UCC  # noqa: E402
    """,
    "tencirchem/static/__init__.py": "",
    "tencirchem/static/evolve_civector.py": """
        from oracle import civector_numpy as O
        def get_civector(params, n_qubits, n_elec, ex_ops, param_ids, hcb=False, init_state=None):
            return O.get_civector(params, n_qubits, n_elec, ex_ops, param_ids, hcb, init_state)
        def get_energy_and_grad_civector(params, hamiltonian, n_qubits, n_elec, ex_ops, param_ids, hcb=False, init_state=None):
            return O.get_energy_and_grad(params, hamiltonian, n_qubits, n_elec, ex_ops, param_ids, hcb, init_state)
        def apply_excitation_civector(civector, n_qubits, n_elec, f_idx, hcb):
            return O.apply_excitation(civector, n_qubits, n_elec, f_idx, hcb)
        def clear():
            O.clear_cache()
    """,
    "tencirchem/static/hamiltonian.py": """
        from inspect import isfunction
        from oracle import civector_numpy as O
        def get_integral_from_hf(hf, active_space=None):
            return hf["int1e"], hf["int2e"], hf["e_core"]
        def get_h_from_integral(int1e, int2e, n_elec, hcb, htype):
            assert htype == "fcifunc"
            return O.get_h_from_integral(int1e, int2e, n_elec, hcb)
        def apply_op(op, state):
            if isfunction(op):
                return op(state)
            return op @ state
    """,
    "tencirchem/static/engine_ucc.py": """
        import numpy as np
        from tencirchem.static.evolve_civector import get_civector as get_civector_, get_energy_and_grad_civector, apply_excitation_civector
        from tencirchem.static.hamiltonian import apply_op
        GETVECTOR_MAP = {"civector": get_civector_}
        ENERGY_AND_GRAD_MAP = {"civector": get_energy_and_grad_civector}
        APPLY_EXCITATION_MAP = {"civector": apply_excitation_civector}
        def get_civector(params, n_qubits, n_elec, ex_ops, param_ids, hcb, init_state, engine):
            if param_ids is None:
                param_ids = range(len(ex_ops))
            ket = GETVECTOR_MAP[engine](params, n_qubits=n_qubits, n_elec=n_elec, ex_ops=tuple(ex_ops), param_ids=tuple(param_ids),
                                        hcb=hcb, init_state=init_state)
            assert engine.startswith("civector")  # engine_ucc.py:62: civector engines return the ket as is
            return ket
        def get_energy(params, hamiltonian, n_qubits, n_elec, ex_ops, param_ids, hcb, init_state, engine):
            if param_ids is None:
                param_ids = range(len(ex_ops))
            ket = GETVECTOR_MAP[engine](params, n_qubits, n_elec, tuple(ex_ops), tuple(param_ids), hcb=hcb, init_state=init_state)
            hket = apply_op(hamiltonian, ket)
            return ket @ hket
        def get_energy_and_grad(params, hamiltonian, n_qubits, n_elec, ex_ops, param_ids, hcb, init_state, engine):
            if engine not in ENERGY_AND_GRAD_MAP:
                raise ValueError(f"Engine '{engine}' not supported")
            return ENERGY_AND_GRAD_MAP[engine](params, hamiltonian, n_qubits, n_elec, tuple(ex_ops), tuple(param_ids), hcb, init_state)
        def apply_excitation(state, n_qubits, n_elec, ex_op, hcb, engine):
            if engine not in APPLY_EXCITATION_MAP:
                raise ValueError(f"Engine '{engine}' not supported")
            return APPLY_EXCITATION_MAP[engine](np.asarray(state), n_qubits, n_elec, ex_op, hcb)
    """,
    "tencirchem/static/ucc.py": """
        import numpy as np
        from tencirchem.static.engine_ucc import get_civector, get_energy, get_energy_and_grad, apply_excitation
        from tencirchem.static.hamiltonian import get_integral_from_hf, get_h_from_integral
        class UCC:
            def __init__(self, int1e, int2e, n_elec, e_core, ex_ops, param_ids, engine=None, hcb=False):
                self.hf = {"int1e": int1e, "int2e": int2e, "e_core": e_core}
                self.active_space = None
                self.n_elec, self.hcb = n_elec, hcb
                self.n_qubits = len(int1e) if hcb else 2 * len(int1e)
                self._check_engine(engine)
                if engine is None:
                    engine = "civector"
                self.engine = engine
                # ucc.py:305-308
                self.hamiltonian_lib = {}
                self.int1e = self.int2e = None
                self.hamiltonian, self.e_core, _ = self._get_hamiltonian_and_core(self.engine)
                self.ex_ops, self.param_ids, self.init_state = ex_ops, param_ids, None
            def _check_engine(self, engine):
                supported_engine = [None, "tensornetwork", "statevector", "civector", "civector-large", "pyscf"]
                if engine not in supported_engine:
                    raise ValueError(f"Engine '{engine}' not supported")
            def _get_hamiltonian_and_core(self, engine):
                self._check_engine(engine)
                if engine is None:
                    engine, hamiltonian, e_core = self.engine, self.hamiltonian, self.e_core
                else:
                    assert engine.startswith("civector") or engine == "pyscf"
                    htype = "fcifunc"
                    hamiltonian = self.hamiltonian_lib.get(htype)
                    if hamiltonian is None:
                        if self.int1e is None:
                            self.int1e, self.int2e, e_core = get_integral_from_hf(self.hf, self.active_space)
                        else:
                            e_core = self.e_core
                        hamiltonian = get_h_from_integral(self.int1e, self.int2e, self.n_elec, self.hcb, htype)
                        self.hamiltonian_lib[htype] = hamiltonian
                    else:
                        e_core = self.e_core
                return hamiltonian, e_core, engine
            def civector(self, params, engine=None):
                self._check_engine(engine)
                if engine is None:
                    engine = self.engine
                return get_civector(params, self.n_qubits, self.n_elec, self.ex_ops, self.param_ids, self.hcb, self.init_state, engine)
            def energy(self, params, engine=None):
                hamiltonian, e_core, engine = self._get_hamiltonian_and_core(engine)
                e = get_energy(params, hamiltonian, self.n_qubits, self.n_elec, self.ex_ops, self.param_ids, self.hcb, self.init_state, engine)
                return float(e) + e_core
            def energy_and_grad(self, params, engine=None):
                hamiltonian, e_core, engine = self._get_hamiltonian_and_core(engine)
                e, g = get_energy_and_grad(params, hamiltonian, self.n_qubits, self.n_elec, self.ex_ops, self.param_ids, self.hcb, self.init_state, engine)
                return float(e) + e_core, np.asarray(g)
            def apply_excitation(self, state, ex_op, engine=None):
                self._check_engine(engine)
                if engine is None:
                    engine = self.engine
                return apply_excitation(state, self.n_qubits, self.n_elec, ex_op, hcb=self.hcb, engine=engine)
    """,
}


@pytest.fixture()
def stub_tencirchem(tmp_path, monkeypatch):
    for rel, src in STUB.items():
        path = tmp_path / rel
        path.parent.mkdir(parents=True, exist_ok=True)
        path.write_text(textwrap.dedent(src))
    monkeypatch.syspath_prepend(str(tmp_path))
    monkeypatch.syspath_prepend(ROOT)
    for name in [m for m in sys.modules if m == "tencirchem" or m.startswith("tencirchem.")]:
        monkeypatch.delitem(sys.modules, name)
    pkg = importlib.import_module("tencirchem")
    yield pkg
    for name in [m for m in sys.modules if m == "tencirchem" or m.startswith("tencirchem.")]:
        sys.modules.pop(name, None)


def _inputs(n=3):
    from oracle import civector_numpy as O
    from oracle import pools

    int1e, int2e = O.random_integral(n)
    ex_ops, param_ids = pools.uccsd_ex_ops(n // 2 + n % 2, n // 2) if n % 2 else pools.uccsd_ex_ops(n // 2, n // 2)
    n_elec = 2 * (n // 2 + n % 2) if n % 2 else n
    np.random.seed(2077)
    params = np.random.rand(max(param_ids) + 1) - 0.5
    return int1e, int2e, n_elec, tuple(ex_ops), tuple(param_ids), params


def test_register_patches_the_reference_maps(stub_tencirchem):
    import tencirchem_b200
    from tencirchem_b200 import engine_ucc as b200_engine
    from tencirchem.static import engine_ucc as ref_engine, ucc as ref_ucc

    name = b200_engine.ENGINE_NAME
    int1e, int2e, n_elec, ex_ops, param_ids, params = _inputs(2)
    with pytest.raises(ValueError, match="not supported"):
        ref_ucc.UCC(int1e, int2e, n_elec, 0.25, ex_ops, param_ids, engine=name)  # before the patch: unknown engine
    assert tencirchem_b200.register() is True
    for m in (ref_engine.GETVECTOR_MAP, ref_engine.ENERGY_AND_GRAD_MAP, ref_engine.APPLY_EXCITATION_MAP):
        assert name in m and "civector" in m
    # the whitelist accepts the new name and still rejects others with the reference's message (ucc.py:429-430)
    ucc = ref_ucc.UCC(int1e, int2e, n_elec, 0.25, ex_ops, param_ids)  # default engine: stock path untouched
    ucc._check_engine(name)
    with pytest.raises(ValueError, match="Engine 'nope' not supported"):
        ucc._check_engine("nope")
    e_ref, g_ref = ucc.energy_and_grad(params)  # stock engine still works through the patched class
    assert np.isfinite(e_ref) and g_ref.shape == params.shape
    # constructor path (ucc.py:305-308) with the new engine: the handle is the B200 one, cached under its own key
    ucc_b = ref_ucc.UCC(int1e, int2e, n_elec, 0.25, ex_ops, param_ids, engine=name)
    assert getattr(ucc_b.hamiltonian, "is_b200_handle", False) and ucc_b.e_core == 0.25
    assert "fcifunc-b200" in ucc_b.hamiltonian_lib and ucc_b.int1e is int1e
    h2, e_core2, eng2 = ucc_b._get_hamiltonian_and_core(None)
    assert h2 is ucc_b.hamiltonian and e_core2 == 0.25 and eng2 == name
    # per-call engine override on an instance built with the stock engine: second handle, same e_core
    h3, e_core3, eng3 = ucc._get_hamiltonian_and_core(name)
    assert getattr(h3, "is_b200_handle", False) and e_core3 == 0.25 and eng3 == name
    assert not getattr(ucc.hamiltonian, "is_b200_handle", False)  # the instance's own handle is untouched
    # clear_cache chains to the B200 plan cache
    from tencirchem_b200 import evolve_civector as b200_ev

    b200_ev._PLAN_CACHE["sentinel"] = object()
    stub_tencirchem.clear_cache()
    assert "sentinel" not in b200_ev._PLAN_CACHE and stub_tencirchem.CLEARED


def test_dispatch_reaches_the_product_and_fails_loudly_without_gpu(stub_tencirchem):
    """No CPU fallback: with the engine registered, a call through the reference-shaped class reaches the CUDA library
    and raises EngineUnavailable on a box without a GPU (on a GPU box it simply computes)."""
    import tencirchem_b200
    from tencirchem_b200._lib import EngineUnavailable
    from tencirchem.static import ucc as ref_ucc

    assert tencirchem_b200.register()
    int1e, int2e, n_elec, ex_ops, param_ids, params = _inputs(2)
    ucc = ref_ucc.UCC(int1e, int2e, n_elec, 0.0, ex_ops, param_ids, engine=tencirchem_b200.engine_ucc.ENGINE_NAME)
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        e, g = ucc.energy_and_grad(params)
        assert np.isfinite(e)
    else:
        with pytest.raises(EngineUnavailable):
            ucc.energy_and_grad(params)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 4, 5])
def test_registered_engine_matches_stock_engine_through_the_reference_class(stub_tencirchem, n):
    import tencirchem_b200
    from tencirchem.static import ucc as ref_ucc

    assert tencirchem_b200.register()
    name = tencirchem_b200.engine_ucc.ENGINE_NAME
    int1e, int2e, n_elec, ex_ops, param_ids, params = _inputs(n)
    ucc = ref_ucc.UCC(int1e, int2e, n_elec, 0.125, ex_ops, param_ids)  # stock engine; B200 per call
    e0, g0 = ucc.energy_and_grad(params)
    e1, g1 = ucc.energy_and_grad(params, engine=name)
    assert abs(e0 - e1) < 1e-10 and np.abs(g0 - g1).max() < 1e-8
    assert abs(ucc.energy(params) - ucc.energy(params, engine=name)) < 1e-10
    c0, c1 = ucc.civector(params), ucc.civector(params, engine=name)
    assert np.abs(c0 - c1).max() < 1e-12
    for op in (ex_ops[0], ex_ops[-1]):
        assert np.abs(ucc.apply_excitation(c0, op) - ucc.apply_excitation(c0, op, engine=name)).max() < 1e-12
    # an instance constructed with the new engine (constructor call of _get_hamiltonian_and_core)
    ucc_b = ref_ucc.UCC(int1e, int2e, n_elec, 0.125, ex_ops, param_ids, engine=name)
    e2, g2 = ucc_b.energy_and_grad(params)
    assert abs(e0 - e2) < 1e-10 and np.abs(g0 - g2).max() < 1e-8
    # the user-visible handle ucc.hamiltonian(civector) (ucc_functions.ipynb cell 41) runs on the GPU
    assert np.abs(ucc_b.hamiltonian(c0) - ucc.hamiltonian(c0)).max() < 1e-11
