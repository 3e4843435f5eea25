"""CPU oracle for the PyIMCOM per-postage-stamp coaddition hot path.

TEST INFRASTRUCTURE -- not product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this package; the product
(``pyimcom_b200``) never does and fails loudly when its CUDA library is missing.

Contents (each function cites the reference file:line it restates):

* ``oracle.routines``  -- ctypes front-end of ``croutines.c``: the furry_parakeet / routine.py
  primitives ``iD5512C, iD5512C_sym, gridD5512C, lakernel1, lsolve_sps, build_reduced_T_wrap``
  with the reference's in-place NumPy signatures (routine.py:125-588).
* ``oracle.lakernel``  -- ``CholKernel, EigenKernel, IterKernel`` (lakernel.py:50-744) on
  NumPy/SciPy LAPACK, exposing the f64 intermediates the parity tests compare.
* ``oracle.sysmat``    -- A / mBhalf assembly for one OutStamp from pixel positions and PSF-overlap
  tables (psfutil.py:1401-1732, coadd.py:886-1085) and the coaddition tail (coadd.py:1221-1363).

Parity pinning: ``tests/golden/make_golden.py`` imports the *reference itself* from
``/root/reference`` (behind ``oracle/refhost.py``'s stub finder) in the build container, runs the
reference's own OutStamp path on seeded synthetic blocks and commits the results as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this oracle against them, together
with the known-answer vectors of the reference's tests/pyimcom/test_routine.py and test_la.py.
"""
