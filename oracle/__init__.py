"""CPU oracle for the Marlin spectral time-step hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, on libTorch's CPU kernels (Python
``torch`` 2.11 = the same ATen/MKL arithmetic the reference dispatches to), the algorithm
of the reference's hot path.  It is imported only by ``tests/``, by
``__graft_entry__.smoke()`` and by the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py``.  Nothing in the product (``marlin_b200/``) imports it and the product never
falls back to it.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks this restatement against the
reference's own gold files (extracted into ``tests/golden/*.npz`` by
``tests/golden/make_golden.py``):
  * test/tests/cahnhilliard/gold/cahnhilliard_out.e  (2-D CH, ABM order 2, 10x10 substeps)
  * test/tests/solvers/gold/diagonal_*.csv           (AB orders 1-4, AM corrector)
  * test/tests/mechanics/gold/mech3d.h5              (3-D de Geus finite-strain solve)
  * test/tests/gradient/gold/*.csv, tensor_compute/gold/backandforth_out.csv
  * unit/src/ParsedTensorTest.C known answers (parser / simplify / derivative strings)
Unpinned by any reference test (stated in DESIGN.md): AB order 5, fp32 results, FFTSemiImplicit
as a stand-alone operator, grids >= 150^2 for CH and > 16^3 for mechanics.

Every function cites the reference file:line it follows (paths relative to the reference
repository root).
"""
