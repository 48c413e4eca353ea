"""ctypes binding of ``libtcc_b200.so`` (the C ABI declared in ``include/tcc_b200.h``).

This is the same stub a TenCirChem maintainer would add to bind the library.
The library is built in-tree by ``tencirchem_b200/csrc/build.sh`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TCC_B200_LIB_NAME: another build of the SAME library in tencirchem_b200/lib (kernel experiments with other compile flags)
LIB_PATH = os.path.join(_HERE, "lib", os.environ.get("TCC_B200_LIB_NAME", "libtcc_b200.so"))

TCC_OK = 0
TCC_ERR_TOO_MANY_QUBITS = -1
TCC_ERR_BAD_EXCITATION = -2
TCC_ERR_BAD_ARGUMENT = -3
TCC_ERR_NO_HAMILTONIAN = -4
TCC_ERR_CUDA = -5
TCC_ERR_NO_DEVICE = -6

_p = C.c_void_p
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

# name -> (restype, argtypes); must list every symbol of include/tcc_b200.h
SIGNATURES = {
    "tcc_last_error": (C.c_char_p, []),
    "tcc_version": (C.c_char_p, []),
    "tcc_device_count": (C.c_int, []),
    "tcc_plan_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, _ip, C.c_int, C.POINTER(_p)]),
    "tcc_plan_destroy": (None, [_p]),
    "tcc_civector_size": (C.c_int64, [_p]),
    "tcc_num_strings": (C.c_int64, [_p]),
    "tcc_num_ops": (C.c_int, [_p]),
    "tcc_num_params": (C.c_int, [_p]),
    "tcc_plan_device_bytes": (C.c_int64, [_p]),
    "tcc_num_batches": (C.c_int, [_p]),
    "tcc_batch_info": (C.c_int, [_p, C.c_int] + [_ip] * 7),
    "tcc_batch_pairs": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.c_int, _ip, _ip]),
    "tcc_storage_order": (C.c_int, [_p, C.POINTER(C.c_uint32)]),
    "tcc_to_storage_order": (C.c_int, [_p, _p, _p, _p]),
    "tcc_from_storage_order": (C.c_int, [_p, _p, _p, _p]),
    "tcc_ci_strings_spin": (C.c_int, [_p, C.POINTER(C.c_uint64)]),
    "tcc_excitation_tables": (C.c_int, [_p, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_int8), C.POINTER(C.c_int8)]),
    "tcc_set_hamiltonian": (C.c_int, [_p, _dp, _dp]),
    "tcc_apply_h": (C.c_int, [_p, _p, _p, _p]),
    "tcc_civector": (C.c_int, [_p, _dp, _p, _p, _p]),
    "tcc_energy": (C.c_int, [_p, _dp, _p, _dp, _p]),
    "tcc_energy_and_grad": (C.c_int, [_p, _dp, _p, _dp, _dp, _p]),
    "tcc_energy_and_grad_batch": (C.c_int, [_p, C.c_int, _dp, _dp, _dp, _p]),
    "tcc_apply_excitation": (C.c_int, [_p, _p, _ip, C.c_int, _p, _p]),
    "tcc_civector_host": (C.c_int, [_p, _dp, _dp, _dp]),
    "tcc_apply_h_host": (C.c_int, [_p, _dp, _dp]),
    "tcc_energy_and_grad_host": (C.c_int, [_p, _dp, _dp, _dp, _dp]),
    "tcc_apply_excitation_host": (C.c_int, [_p, _dp, _ip, C.c_int, _dp]),
    "tcc_rotate": (C.c_int, [_p, C.c_int, C.c_double, _p, _p]),
    "tcc_reverse_step": (C.c_int, [_p, C.c_int, C.c_double, _p, _p, _p, _p]),
    "tcc_op_support_bytes": (C.c_int64, [_p, C.c_int]),
    "tcc_launch_count": (C.c_int64, [_p]),
    "tcc_make_rdm12": (C.c_int, [_p, _p, _dp, _dp, _p]),
    "tcc_make_rdm12_hcb": (C.c_int, [_p, _p, _dp, _dp, _dp, _p]),
    "tcc_plan_create_sharded": (C.c_int, [C.c_int, C.c_int, C.c_int, _ip, _ip, _ip, C.c_int, C.c_int, C.c_int, C.POINTER(_p)]),
    "tcc_shard_info": (C.c_int, [_p, _ip, _ip, _ip, _ip, C.POINTER(C.c_int64)]),
    "tcc_shard_segment": (C.c_int, [_p, C.c_int, _ip, _ip, _ip]),
    "tcc_shard_layout": (C.c_int, [_p, C.c_int, _ip, C.POINTER(C.c_uint32)]),
    "tcc_shard_set_params": (C.c_int, [_p, _dp, _p]),
    "tcc_shard_forward": (C.c_int, [_p, C.c_int, _p, _p]),
    "tcc_shard_reverse": (C.c_int, [_p, C.c_int, _p, _p, _p]),
    "tcc_shard_reverse_strided": (C.c_int, [_p, C.c_int, _p, _p, C.c_int64, _p]),
    "tcc_shard_reverse_interleaved": (C.c_int, [_p, C.c_int, _p, _p]),
    "tcc_shard_can_interleave": (C.c_int, [_p]),
    "tcc_pool_gradients": (C.c_int, [_p, _p, _p, _dp, _p]),
    "tcc_push_rows": (C.c_int, [C.c_int, _p, C.c_int64, C.c_int64, _p, _p, _p, _p, C.c_int64, _p]),
    "tcc_shard_gradient": (C.c_int, [_p, _dp, _p]),
    "tcc_apply_h_rows": (C.c_int, [_p, _p, _p, C.c_int64, C.c_int64, _p]),
    "tcc_shard_sigma_layouts": (C.c_int, [_p, _ip]),
    "tcc_shard_sigma_bb": (C.c_int, [_p, _p, _p, C.c_int64, _p]),
    "tcc_shard_sigma_aa": (C.c_int, [_p, _p, _p, C.c_int64, _p]),
    "tcc_shard_sigma_ab": (C.c_int, [_p, C.c_int, _p, _p, _p]),
    "tcc_profile_enable": (C.c_int, [_p, C.c_int]),
    "tcc_profile_read": (C.c_int, [_p, _dp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "tcc_profile_batches": (C.c_int, [_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
}

PROF_SLOTS = ("fwd_beta", "fwd_alpha", "fwd_alphabeta", "rev_beta", "rev_alpha", "rev_alphabeta", "sigma_ss", "sigma_ab", "dot", "fwd_batch", "rev_batch")

_lib = None


class EngineUnavailable(RuntimeError):
    """The CUDA library or a GPU is missing.  There is deliberately no CPU fallback."""


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineUnavailable(
            f"{LIB_PATH} not found: build it with tencirchem_b200/csrc/build.sh "
            "(the B200 engine has no CPU fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().tcc_last_error().decode()


def check(status: int):
    """Map a status code back to the exception the reference raises at the same place."""
    if status == TCC_OK:
        return
    msg = last_error()
    if status in (TCC_ERR_TOO_MANY_QUBITS, TCC_ERR_BAD_EXCITATION, TCC_ERR_BAD_ARGUMENT, TCC_ERR_NO_HAMILTONIAN):
        raise ValueError(msg)
    if status == TCC_ERR_NO_DEVICE:
        raise EngineUnavailable(msg)
    raise RuntimeError(f"tcc_b200 status {status}: {msg}")
