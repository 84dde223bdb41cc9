"""CPU oracle for the echopype calibrate -> clean -> commongrid hot path.

TEST INFRASTRUCTURE ONLY.  This package is a float64 numpy/scipy/pandas restatement of the
reference algorithm (OSOceanAcoustics/echopype @ 13c0fa0).  It exists so that the CUDA product
path in ``echopype_b200`` can be checked against the reference arithmetic on the same inputs.

Rules (enforced by tests/test_abi_and_isolation.py::test_product_never_imports_oracle):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
    reference`` legs may import anything from here;
  * nothing under ``echopype_b200/`` imports it; the product fails loudly without its CUDA library.

Pinning status (see DESIGN.md "Oracle"): PINNED.
  * compute_Sv / compute_TS for EK60, EK80 CW power (incl. a GPT channel), EK80 CW complex, EK80 broadband (pulse
    compression) and AZFP, the env / cal parameter assembly behind them, echo_range / the TVG range, and
    estimate_ / remove_background_noise: checked against outputs of the REFERENCE'S OWN CODE.  The reference's modules
    (calibrate/range.py, calibrate_ek.py, calibrate_azfp.py, cal_params.py, env_params.py, ek80_complex.py, utils/uwa.py,
    utils/align.py, clean/api.py:362-511) are imported / lifted unmodified from /root/reference and executed over a
    small labelled-array stand-in for xarray (tests/golden/xrlite.py) by tests/golden/make_golden_calibrate.py; the
    results are committed as tests/golden/calibrate_vectors.npz and tests/test_reference_pinned.py asserts that the
    oracle reproduces them to 1e-9 dB / 1e-12 relative (and that the CUDA path matches them directly at 1e-4 dB);
  * seawater formulae, dB<->linear helpers, chirp replica / filter-decimate, per-channel convolution, dB-string and
    bin-string parsers: pinned against outputs of the reference's own functions (tests/golden/make_golden.py);
  * the impulse / transient / attenuated-signal noise masks (clean/api.py:30-359 and their workers in clean/utils.py),
    consolidate.add_depth (consolidate/api.py:66-247 with utils/align.py and consolidate/ek_depth_utils.py),
    mask.frequency_differencing / apply_mask (mask/api.py:41-676 with the equation parser and helpers) and compute_MVBS_index_binning
    (commongrid/api.py:195-266): executed the same way by tests/golden/make_golden_{masks,consolidate,mask,commongrid}.py;
    tests/test_reference_pinned_masks.py and tests/test_reference_pinned_consolidate.py assert that the oracle reproduces
    those outputs (masks exactly, values to 1e-9 dB or better) and compare the CUDA results with them directly;
  * MVBS, NASC, pulse-length lookup, env-param interpolation and impulse noise on depth VALUES (flox):
    pinned against the reference's own known-answer tests (restated in tests/test_oracle_*.py, tests/test_mask*.py); the
    bin reduction itself is flox's (a third-party dependency absent here), anchored on the reference's brute-force mock test.

Each function cites the reference file:line it follows (paths relative to /root/reference).
"""
