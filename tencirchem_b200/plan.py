"""Python handle of a ``tcc_plan`` (include/tcc_b200.h): the cached excitation tables of one ansatz.

Plays the role of the reference's ``CI_OPERATOR_BATCH_CACHE`` entry
(tencirchem/static/evolve_civector.py:109-150) but holds compact per-spin hop lists on the GPU
instead of ``[M, N]`` tables.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _dptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


class Plan:
    def __init__(self, n_qubits: int, n_elec: int, ex_ops: Sequence[Sequence[int]], param_ids=None,
                 hcb: bool = False, device: Optional[int] = 0, shard: Optional[Sequence[int]] = None):
        """``device=None`` builds a host-only plan (strings / table introspection, no compute).
        ``shard=(rank, world)`` builds the plan of one rank of an alpha-row sharded engine (tcc_plan_create_sharded)."""
        self.lib = _lib.load()
        self.n_qubits, self.n_elec, self.hcb = int(n_qubits), int(n_elec), bool(hcb)
        self.ex_ops = tuple(tuple(int(i) for i in op) for op in ex_ops)
        if param_ids is None:  # engine_ucc.py:43-44
            param_ids = range(len(self.ex_ops))
        self.param_ids = tuple(int(i) for i in param_ids)
        if len(self.param_ids) != len(self.ex_ops):
            raise ValueError("param_ids and ex_ops have different lengths")
        flat = np.asarray([i for op in self.ex_ops for i in op], dtype=np.int32)
        lens = np.asarray([len(op) for op in self.ex_ops], dtype=np.int32)
        pids = np.asarray(self.param_ids, dtype=np.int32)
        handle = C.c_void_p()
        ip = C.POINTER(C.c_int32)
        if shard is None:
            _lib.check(
                self.lib.tcc_plan_create(
                    self.n_qubits, self.n_elec, int(self.hcb), len(self.ex_ops),
                    flat.ctypes.data_as(ip), lens.ctypes.data_as(ip), pids.ctypes.data_as(ip),
                    -1 if device is None else int(device), C.byref(handle),
                )
            )
        else:
            if self.hcb:
                raise ValueError("hcb plans have a single alpha row and are not sharded")
            _lib.check(
                self.lib.tcc_plan_create_sharded(
                    self.n_qubits, self.n_elec, len(self.ex_ops),
                    flat.ctypes.data_as(ip), lens.ctypes.data_as(ip), pids.ctypes.data_as(ip),
                    -1 if device is None else int(device), int(shard[0]), int(shard[1]), C.byref(handle),
                )
            )
        self.handle = handle
        self.device = device
        self.size = int(self.lib.tcc_civector_size(handle))
        self.n_strings = int(self.lib.tcc_num_strings(handle))
        self.n_params = int(self.lib.tcc_num_params(handle))
        self.n_ops = len(self.ex_ops)
        self._h_token = None

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            try:
                self.lib.tcc_plan_destroy(h)
            except Exception:
                pass
            self.handle = None

    # ---- introspection (host) ----
    def ci_strings_spin(self) -> np.ndarray:
        out = np.zeros(self.n_strings, dtype=np.uint64)
        _lib.check(self.lib.tcc_ci_strings_spin(self.handle, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    def excitation_tables(self, j: int):
        perm = np.zeros(self.size, dtype=np.uint64)
        phase = np.zeros(self.size, dtype=np.int8)
        f2 = np.zeros(self.size, dtype=np.int8)
        _lib.check(
            self.lib.tcc_excitation_tables(
                self.handle, int(j), perm.ctypes.data_as(C.POINTER(C.c_uint64)),
                phase.ctypes.data_as(C.POINTER(C.c_int8)), f2.ctypes.data_as(C.POINTER(C.c_int8)),
            )
        )
        return perm, phase, f2

    @property
    def n_batches(self) -> int:
        return int(self.lib.tcc_num_batches(self.handle))

    def batch_info(self, b: int) -> dict:
        """Schedule introspection: ops of batch ``b`` (application order), active orbital counts, tile geometry."""
        vals = [C.c_int32(0) for _ in range(6)]
        _lib.check(self.lib.tcc_batch_info(self.handle, int(b), *[C.byref(v) for v in vals], None))
        n_ops = vals[0].value
        idx = np.zeros(max(n_ops, 1), dtype=np.int32)
        _lib.check(self.lib.tcc_batch_info(self.handle, int(b), *[C.byref(v) for v in vals], idx.ctypes.data_as(C.POINTER(C.c_int32))))
        keys = ("n_ops", "active_alpha", "active_beta", "n_tiles", "max_tile_rows", "max_tile_cols")
        info = {k: v.value for k, v in zip(keys, vals)}
        info["op_indices"] = idx[:n_ops].tolist()
        return info

    def batch_pairs(self, b: int, pos: int, e_alpha: int, e_beta: int, which: int = 0, with_round_len: bool = False):
        """Pair list (which=0) or lean stream (1: 8-byte, 2: 16-byte elements) of one excitation of a batch for tiles
        of class pair (e_alpha, e_beta), as uint32 entries (include/tcc_b200.h, tcc_batch_pairs); optionally also the
        number of entries per round of a stream."""
        n, rl = C.c_int32(0), C.c_int32(0)
        _lib.check(self.lib.tcc_batch_pairs(self.handle, int(b), int(pos), int(e_alpha), int(e_beta), int(which), None, 0,
                                            C.byref(n), C.byref(rl)))
        out = np.zeros(max(n.value, 1), dtype=np.uint32)
        _lib.check(self.lib.tcc_batch_pairs(self.handle, int(b), int(pos), int(e_alpha), int(e_beta), int(which),
                                            out.ctypes.data_as(C.POINTER(C.c_uint32)), int(out.size), C.byref(n), C.byref(rl)))
        return (out[: n.value], rl.value) if with_round_len else out[: n.value]

    # ---- alpha-row sharding (include/tcc_b200.h, multi-GPU section) ----
    def shard_info(self) -> dict:
        v = [C.c_int32(0) for _ in range(4)]
        rows = C.c_int64(0)
        _lib.check(self.lib.tcc_shard_info(self.handle, *[C.byref(x) for x in v], C.byref(rows)))
        return dict(rank=v[0].value, world=v[1].value, n_layouts=v[2].value, n_segments=v[3].value,
                    max_local_rows=rows.value)

    def shard_segments(self) -> list:
        """[(layout, batch_lo, batch_hi)] in application order."""
        out = []
        for seg in range(self.shard_info()["n_segments"]):
            v = [C.c_int32(0) for _ in range(3)]
            _lib.check(self.lib.tcc_shard_segment(self.handle, seg, *[C.byref(x) for x in v]))
            out.append(tuple(x.value for x in v))
        return out

    def shard_layout(self, layout: int):
        """(owner[num_strings] of every storage row, orbital mask T) of a row layout."""
        owner = np.zeros(self.n_strings, dtype=np.int32)
        mask = C.c_uint32(0)
        _lib.check(self.lib.tcc_shard_layout(self.handle, int(layout), owner.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(mask)))
        return owner, mask.value

    def shard_set_params(self, params):
        params = _f64(params)
        if params.shape != (self.n_params,):
            raise ValueError(f"Incompatible parameter shape. {(self.n_params,)} is desired. Got {params.shape}")
        _lib.check(self.lib.tcc_shard_set_params(self.handle, _dptr(params), self._stream()))

    def shard_forward(self, seg: int, vec_local):
        _lib.check(self.lib.tcc_shard_forward(self.handle, int(seg), C.c_void_p(vec_local.data_ptr()), self._stream()))

    def shard_reverse(self, seg: int, ket_local, bra_local):
        _lib.check(self.lib.tcc_shard_reverse(self.handle, int(seg), C.c_void_p(ket_local.data_ptr()),
                                              C.c_void_p(bra_local.data_ptr()), self._stream()))

    def shard_reverse_strided(self, seg: int, ket_local, bra_local, row_stride: int):
        """ket_local / bra_local: views whose consecutive rows lie `row_stride` doubles apart (interleaved block)."""
        _lib.check(self.lib.tcc_shard_reverse_strided(self.handle, int(seg), C.c_void_p(ket_local.data_ptr()),
                                                      C.c_void_p(bra_local.data_ptr()), int(row_stride), self._stream()))

    def shard_reverse_interleaved(self, seg: int, kb_local):
        """kb_local: [local rows, nb, 2] block with (ket, bra) of every determinant side by side."""
        _lib.check(self.lib.tcc_shard_reverse_interleaved(self.handle, int(seg), C.c_void_p(kb_local.data_ptr()), self._stream()))

    def shard_can_interleave(self) -> bool:
        return bool(self.lib.tcc_shard_can_interleave(self.handle))

    def shard_gradient(self) -> np.ndarray:
        g = np.zeros(max(self.n_params, 1), dtype=np.float64)
        _lib.check(self.lib.tcc_shard_gradient(self.handle, _dptr(g), self._stream()))
        return g[: self.n_params]

    def apply_h_rows(self, c_full, sigma_full, row_lo: int, row_cnt: int):
        """sigma_full += contributions of storage rows [row_lo, row_lo + row_cnt) to H c (full storage-order matrices)."""
        _lib.check(self.lib.tcc_apply_h_rows(self.handle, C.c_void_p(c_full.data_ptr()), C.c_void_p(sigma_full.data_ptr()),
                                             int(row_lo), int(row_cnt), self._stream()))

    def shard_sigma_layouts(self):
        out = np.zeros(3, dtype=np.int32)
        _lib.check(self.lib.tcc_shard_sigma_layouts(self.handle, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return [int(x) for x in out]

    def shard_sigma_bb(self, c_local, sigma_local):
        _lib.check(self.lib.tcc_shard_sigma_bb(self.handle, self._ptr(c_local), self._ptr(sigma_local), int(c_local.shape[0]),
                                               self._stream()))

    def shard_sigma_aa(self, c_cols, sigma_cols):
        _lib.check(self.lib.tcc_shard_sigma_aa(self.handle, self._ptr(c_cols), self._ptr(sigma_cols), int(c_cols.shape[1]),
                                               self._stream()))

    def shard_sigma_ab(self, pass_id: int, c_local, sigma_local):
        _lib.check(self.lib.tcc_shard_sigma_ab(self.handle, int(pass_id), self._ptr(c_local), self._ptr(sigma_local),
                                               self._stream()))

    def pool_gradients_dev(self, c, sigma) -> np.ndarray:
        """<sigma| G_k |c> for every excitation k of this plan (the plan's excitation list is the operator pool);
        c, sigma: reference-order CI vectors on the device (tcc_pool_gradients)."""
        self._check_dev(c, "c")
        self._check_dev(sigma, "sigma")
        out = np.zeros(max(self.n_ops, 1), dtype=np.float64)
        _lib.check(self.lib.tcc_pool_gradients(self.handle, self._ptr(c), self._ptr(sigma), _dptr(out), self._stream()))
        return out[: self.n_ops]

    def make_rdm12_dev(self, c):
        """(rdm1 [n, n], rdm2 [n, n, n, n]) of a reference-order CI vector on the device (tcc_make_rdm12)."""
        self._check_dev(c, "c")
        n = self.n_qubits // 2
        rdm1 = np.zeros((n, n), dtype=np.float64)
        rdm2 = np.zeros((n, n, n, n), dtype=np.float64)
        _lib.check(self.lib.tcc_make_rdm12(self.handle, self._ptr(c), _dptr(rdm1), _dptr(rdm2), self._stream()))
        return rdm1, rdm2

    def make_rdm12_hcb_dev(self, c):
        """(occ [n], hop [n, n], both [n, n]) of a reference-order hard-core-boson CI vector on the device
        (tcc_make_rdm12_hcb, puccd.py:148-203)."""
        self._check_dev(c, "c")
        n = self.n_qubits
        occ = np.zeros(n, dtype=np.float64)
        hop = np.zeros((n, n), dtype=np.float64)
        both = np.zeros((n, n), dtype=np.float64)
        _lib.check(self.lib.tcc_make_rdm12_hcb(self.handle, self._ptr(c), _dptr(occ), _dptr(hop), _dptr(both), self._stream()))
        return occ, hop, both

    def to_storage_order_dev(self, ref, out):
        """out (storage order) <- ref (reference order); full CI vectors on the device."""
        _lib.check(self.lib.tcc_to_storage_order(self.handle, self._ptr(ref), self._ptr(out), self._stream()))

    def from_storage_order_dev(self, storage, out):
        _lib.check(self.lib.tcc_from_storage_order(self.handle, self._ptr(storage), self._ptr(out), self._stream()))

    def storage_order(self) -> np.ndarray:
        """ref2int: storage address of each per-spin string, indexed by its reference address."""
        out = np.zeros(self.n_strings, dtype=np.uint32)
        _lib.check(self.lib.tcc_storage_order(self.handle, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    def op_support_bytes(self, j: int) -> int:
        return int(self.lib.tcc_op_support_bytes(self.handle, int(j)))

    @property
    def launch_count(self) -> int:
        return int(self.lib.tcc_launch_count(self.handle))

    @property
    def device_bytes(self) -> int:
        return int(self.lib.tcc_plan_device_bytes(self.handle))

    def profile_enable(self, on=True):
        _lib.check(self.lib.tcc_profile_enable(self.handle, int(bool(on))))

    def profile_read(self):
        """{slot: (ms, launches, algorithmic_bytes)} accumulated since ``profile_enable``."""
        k = len(_lib.PROF_SLOTS)
        ms = np.zeros(k, dtype=np.float64)
        launches = np.zeros(k, dtype=np.int64)
        nbytes = np.zeros(k, dtype=np.int64)
        _lib.check(
            self.lib.tcc_profile_read(
                self.handle, _dptr(ms), launches.ctypes.data_as(C.POINTER(C.c_int64)), nbytes.ctypes.data_as(C.POINTER(C.c_int64))
            )
        )
        return {name: (float(ms[i]), int(launches[i]), int(nbytes[i])) for i, name in enumerate(_lib.PROF_SLOTS)}

    def profile_batches(self, on=True):
        """Returns (fwd_ms, rev_ms) per batch of the last energy_and_grad call (zeros if recording was off)."""
        fwd = np.zeros(max(self.n_batches, 1), dtype=np.float32)
        rev = np.zeros(max(self.n_batches, 1), dtype=np.float32)
        fp = C.POINTER(C.c_float)
        _lib.check(self.lib.tcc_profile_batches(self.handle, int(bool(on)), fwd.ctypes.data_as(fp), rev.ctypes.data_as(fp)))
        return fwd[: self.n_batches], rev[: self.n_batches]

    # ---- Hamiltonian ----
    def set_hamiltonian(self, int1e, int2e, token=None):
        n = self.n_qubits if self.hcb else self.n_qubits // 2
        int1e, int2e = _f64(int1e), _f64(int2e)
        if int1e.shape != (n, n):
            raise ValueError(f"Invalid one-boby integral array shape: {int1e.shape}")  # hamiltonian.py:53-54
        if int2e.shape != (n, n, n, n):
            raise ValueError(f"Invalid two-boby integral array shape: {int2e.shape}")
        _lib.check(self.lib.tcc_set_hamiltonian(self.handle, _dptr(int1e), _dptr(int2e)))
        self._h_token = token if token is not None else object()

    def _check_params(self, params) -> np.ndarray:
        params = _f64(params).reshape(-1)
        if len(params) < self.n_params:
            raise ValueError(f"Incompatible parameter shape. {self.n_params} is desired. Got {len(params)}")  # ucc.py:423-424
        return params

    def _check_state(self, state, name="init_state") -> Optional[np.ndarray]:
        if state is None:
            return None
        state = _f64(state).reshape(-1)
        if len(state) != self.size:
            raise ValueError(f"{name} has {len(state)} elements, the CI vector has {self.size}")
        return state

    # ---- host-buffer calls ----
    def civector(self, params, init_state=None) -> np.ndarray:
        params = self._check_params(params)
        init = self._check_state(init_state)
        out = np.empty(self.size, dtype=np.float64)
        _lib.check(self.lib.tcc_civector_host(self.handle, _dptr(params), _dptr(init), _dptr(out)))
        return out

    def apply_h(self, civector) -> np.ndarray:
        c = self._check_state(civector, "civector")
        out = np.empty(self.size, dtype=np.float64)
        _lib.check(self.lib.tcc_apply_h_host(self.handle, _dptr(c), _dptr(out)))
        return out

    def energy(self, params, init_state=None) -> float:
        params = self._check_params(params)
        init = self._check_state(init_state)
        e = C.c_double(0.0)
        _lib.check(self.lib.tcc_energy_and_grad_host(self.handle, _dptr(params), _dptr(init), C.byref(e), None))
        return float(e.value)

    def energy_and_grad(self, params, init_state=None):
        params = self._check_params(params)
        init = self._check_state(init_state)
        e = C.c_double(0.0)
        grad = np.zeros(max(self.n_params, 1), dtype=np.float64)
        _lib.check(self.lib.tcc_energy_and_grad_host(self.handle, _dptr(params), _dptr(init), C.byref(e), _dptr(grad)))
        g = np.zeros(params.shape)
        g[: self.n_params] = grad[: self.n_params]
        return float(e.value), g

    def energy_and_grad_batch(self, thetas, with_grad=True):
        """Evaluate many parameter sets in the same launches (thetas [n_sets, n_params]); returns
        (energies [n_sets], grads [n_sets, n_params] or None)."""
        thetas = np.ascontiguousarray(np.asarray(thetas, dtype=np.float64))
        if thetas.ndim != 2 or thetas.shape[1] < self.n_params:
            raise ValueError(f"Incompatible parameter shape. (n_sets, {self.n_params}) is desired. Got {thetas.shape}")
        thetas = np.ascontiguousarray(thetas[:, : self.n_params])
        n_sets = len(thetas)
        energies = np.zeros(n_sets, dtype=np.float64)
        grads = np.zeros((n_sets, max(self.n_params, 1)), dtype=np.float64) if with_grad else None
        _lib.check(
            self.lib.tcc_energy_and_grad_batch(self.handle, n_sets, _dptr(thetas), _dptr(energies), _dptr(grads) if with_grad else None, None)
        )
        return energies, (grads[:, : self.n_params] if with_grad else None)

    def apply_excitation(self, civector, ex_op) -> np.ndarray:
        c = self._check_state(civector, "civector")
        op = np.asarray(ex_op, dtype=np.int32)
        out = np.empty(self.size, dtype=np.float64)
        _lib.check(
            self.lib.tcc_apply_excitation_host(
                self.handle, _dptr(c), op.ctypes.data_as(C.POINTER(C.c_int32)), len(op), _dptr(out)
            )
        )
        return out

    # ---- device-pointer calls (torch tensors own the memory) ----
    @staticmethod
    def _ptr(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    @staticmethod
    def _stream():
        import torch

        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _check_dev(self, t, name):
        import torch

        if t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous() or t.numel() != self.size:
            raise ValueError(f"{name} must be a contiguous float64 CUDA tensor with {self.size} elements")

    def civector_dev(self, params, out, init=None):
        params = self._check_params(params)
        self._check_dev(out, "out")
        if init is not None:
            self._check_dev(init, "init")
        _lib.check(self.lib.tcc_civector(self.handle, _dptr(params), self._ptr(init), self._ptr(out), self._stream()))
        return out

    def apply_h_dev(self, c, sigma):
        self._check_dev(c, "c")
        self._check_dev(sigma, "sigma")
        _lib.check(self.lib.tcc_apply_h(self.handle, self._ptr(c), self._ptr(sigma), self._stream()))
        return sigma

    def energy_and_grad_dev(self, params, init=None):
        """Vectors stay in the plan's own device buffers; only (E, grad) come back."""
        params = self._check_params(params)
        if init is not None:
            self._check_dev(init, "init")
        e = C.c_double(0.0)
        grad = np.zeros(max(self.n_params, 1), dtype=np.float64)
        _lib.check(
            self.lib.tcc_energy_and_grad(self.handle, _dptr(params), self._ptr(init), C.byref(e), _dptr(grad), self._stream())
        )
        return float(e.value), grad[: self.n_params].copy()

    def rotate_dev(self, j, theta, c):
        self._check_dev(c, "c")
        _lib.check(self.lib.tcc_rotate(self.handle, int(j), float(theta), self._ptr(c), self._stream()))

    def reverse_step_dev(self, j, theta, ket, bra, grad_slot):
        self._check_dev(ket, "ket")
        self._check_dev(bra, "bra")
        _lib.check(
            self.lib.tcc_reverse_step(
                self.handle, int(j), float(theta), self._ptr(ket), self._ptr(bra), C.c_void_p(grad_slot.data_ptr()), self._stream()
            )
        )
