"""FCI ground-state energy from the engine's own H.c (``tcc_apply_h``): the reference value the UCC classes report as
``e_fci`` (ucc.py:284-300 takes it from PySCF's FCI solver; there is no PySCF here, and sigma = H.c on the GPU is
exactly what an FCI solver iterates).  Plain Lanczos from the Hartree-Fock determinant: only the lowest eigenvalue is
wanted, so three vectors are kept and the tridiagonal matrix is diagonalised on the host; loss of orthogonality only
produces copies of converged eigenvalues and does not move the lowest one.

Host-side reference tooling (like ``hchain``): torch is used for the vector updates and dot products, the Hamiltonian
action is the CUDA path of the engine."""
from __future__ import annotations

import numpy as np


def lanczos_ground_energy(apply_h, n: int, device, tol: float = 1e-10, max_iter: int = 400, start_index: int = 0) -> float:
    """Lowest eigenvalue of the symmetric operator ``apply_h(v, out)`` on float64 device vectors of length n."""
    import torch
    from scipy.linalg import eigvalsh_tridiagonal

    v = torch.zeros(n, dtype=torch.float64, device=device)
    v[start_index] = 1.0
    v_prev = torch.zeros_like(v)
    w = torch.empty_like(v)
    alphas, betas = [], []
    beta = 0.0
    e_old = None
    for it in range(max_iter):
        apply_h(v, w)
        alpha = float(torch.dot(v, w))
        alphas.append(alpha)
        w.add_(v, alpha=-alpha)
        if it:
            w.add_(v_prev, alpha=-beta)
        beta = float(torch.linalg.vector_norm(w))
        e = float(eigvalsh_tridiagonal(np.array(alphas), np.array(betas), select="i", select_range=(0, 0))[0]) if betas else alpha
        if beta < 1e-13 or (e_old is not None and abs(e - e_old) < tol and it > 2):
            return e
        e_old = e
        betas.append(beta)
        v_prev, v, w = v, w, v_prev
        v.mul_(1.0 / beta)
    return e_old


def fci_energy(ucc, tol: float = 1e-10, max_iter: int = 400) -> float:
    """E_FCI (electronic + e_core) of the active space a UCC object was built on."""
    import torch

    from .evolve_civector import _bind_hamiltonian, get_plan

    from .hamiltonian import get_h_from_integral

    # always the fermionic problem: a pUCCD object (hard-core bosons on n qubits) reports the molecular FCI energy too
    n_orb = ucc.n_qubits if ucc.hcb else ucc.n_qubits // 2
    plan = get_plan(2 * n_orb, ucc.n_elec, (), (), False)
    ham = get_h_from_integral(ucc.int1e, ucc.int2e, ucc.n_elec, False, "fcifunc") if ucc.hcb else ucc.hamiltonian
    _bind_hamiltonian(plan, ham)
    device = torch.device(f"cuda:{plan.device}")
    e = lanczos_ground_energy(lambda x, y: plan.apply_h_dev(x, y), plan.size, device, tol, max_iter)
    return e + float(ucc.e_core)
