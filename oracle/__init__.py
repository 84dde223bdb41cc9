"""CPU oracle for the echopype calibrate -> clean -> commongrid hot path.

TEST INFRASTRUCTURE ONLY.  This package is a float64 numpy/scipy/pandas restatement of the
reference algorithm (OSOceanAcoustics/echopype @ 13c0fa0).  It exists so that the CUDA product
path in ``echopype_b200`` can be checked against the reference arithmetic on the same inputs.

Rules (enforced by tests/test_no_oracle_in_product.py):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
    reference`` legs may import anything from here;
  * nothing under ``echopype_b200/`` imports it; the product fails loudly without its CUDA library.

Pinning status (see DESIGN.md "Oracle"):
  * seawater formulae, dB<->linear helpers, chirp replica / filter-decimate, per-channel
    convolution, dB-string and bin-string parsers: pinned against outputs of the reference's own
    functions executed from /root/reference (tests/golden/make_golden.py, committed vectors);
  * noise removal, MVBS, NASC, index binning, pulse-length lookup, env-param interpolation: pinned
    against the reference's own known-answer tests (restated in tests/test_oracle_*.py);
  * full compute_Sv / compute_TS numeric values for EK60 / EK80 / AZFP: PARITY UNPINNED offline -
    every value-pinning test of the reference needs pooch-downloaded raw files and EchoView /
    MATLAB goldens that are not available here.  Those functions are a
    formula restatement guarded by closed-form checks.

Each function cites the reference file:line it follows (paths relative to /root/reference).
"""
