"""CPU oracle for the dtFFT GPU reshape path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy / plain-C restatement of the reference's algorithm for the
hot path (local permutes, pack/unpack, the pencil decomposition and the per-peer
block geometry of ``reshape_handle_generic``).  It is the *checker* for the CUDA
product in ``dtfft_b200/``; nothing in the product imports it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import or execute anything under ``oracle/``.

Pinning status (see DESIGN.md "Oracle"): the reference (Fortran 2018 + MPI) cannot be
built in this image and ships no stored golden vectors.  The oracle is pinned against
the reference's own known-answer *properties* (tests/test_oracle_pins.py restates
``src/tests/test_host_kernels.F90`` and ``src/tests/test_device_kernels.F90``) and by
cross-checking two independent restatements: the kernel-level index maps of
``src/include/_dtfft_kernel_host_routines.inc`` driven by the block geometry of
``src/dtfft_reshape_handle_generic.F90`` must reproduce, bit for bit, the global-array
redistribution that the host MPI-datatype path
(``src/dtfft_reshape_handle_datatype.F90``) performs by construction -- and that
construction is restated too (``datatype_path.py``: the derived datatypes, displacements and
all-to-all(w) of that file) and checked against the redistribution
(tests/test_oracle_datatype_path.py).  On the GPU box the oracle is additionally pinned by
golden vectors and digests produced by the reference's own regenerated device kernels
(tests/golden/, tests/test_oracle_golden.py, tests/test_vs_reference_kernels_gpu.py).
"""
