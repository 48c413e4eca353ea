"""tencirchem_b200 -- B200-native (sm_100a) implementation of TenCirChem's ``civector`` UCC engine.

Only the hot path behind ``UCCSD/kUpCCGSD/pUCCD.energy()/.energy_and_grad()/.civector()`` lives
here: excitation rotations, H|psi> and the reverse-sweep gradient, as
hand-written CUDA behind the C ABI of ``include/tcc_b200.h``.  There is no CPU fallback: every
compute entry point raises if the CUDA library or a GPU is missing.
"""
from .engine_ucc import (  # noqa: F401
    ENGINE_NAME,
    apply_excitation,
    get_civector,
    get_energy,
    get_energy_and_grad,
    register,
)
from .evolve_civector import clear_cache  # noqa: F401
from .ansatz import KUPCCGSD, PUCCD, UCC, UCCSD  # noqa: F401

__version__ = "0.1.0"
