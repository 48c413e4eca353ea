"""CPU oracle for the TenCirChem `civector` UCC hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker or as the timed CPU baseline.  The product (``tencirchem_b200``)
never imports this package and fails loudly if its CUDA library is missing.

Parity status: PINNED against the reference's own known-answer values that do
not need PySCF/OpenFermion to reproduce: the published
H2 civector / energies, the integer doctests for CI strings and addresses, the
sign doctest of ``apply_excitation`` and the published op-pool lists.  The
third-party pieces of the path that are absent from ``/root/reference``
(pyscf==2.7.0 ``cistring``/``direct_nosym``; openfermion==1.6.1
``jordan_wigner``) are restated from their published algorithms and
cross-checked against a brute-force second-quantised model
(``oracle/bruteforce.py``).
"""
