"""CPU tests of the host-side parameter assembly (echopype_b200/calibrate/*) against the independent
oracle-based restatement in tests/oracle_glue.py.  No CUDA needed: the calibrator constructors and the
parameter dictionaries they feed to the row-setup kernels are pure numpy."""

import numpy as np
import pytest

import oracle_glue as og
from echopype_b200 import synth
from echopype_b200.calibrate.calibrate_azfp import CalibrateAZFP
from echopype_b200.calibrate.calibrate_ek import CalibrateEK60, CalibrateEK80, _cp


def _bc(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.ndim == 1 and b.ndim == 2:
        a = a[:, None]
    if b.ndim == 1 and a.ndim == 2:
        b = b[:, None]
    return np.broadcast_arrays(a, b)


def _same(a, b, rtol=1e-13):
    a, b = _bc(a, b)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=0, equal_nan=True)


@pytest.mark.parametrize("time_varying", [False, True])
def test_ek60_params(time_varying):
    ed = synth.make_ek60(4, 12, 16, time_varying=time_varying)
    prm, te = CalibrateEK60(ed)._power_params("Sv")
    want = og.ek60(ed, "Sv")["params"]
    for k in ("sound_speed", "sound_absorption", "gain_correction", "sa_correction", "tau_effective"):
        _same(prm[k], want[k])
    assert te.dims == ("channel",)


def test_azfp_params():
    ed = synth.make_azfp(4, 6, 16)
    cal = CalibrateAZFP(ed, env_params={"salinity": 30.0, "pressure": 50.0})
    chan = ed["Sonar/Beam_group1"]["channel"].values
    want = og.azfp(ed, "Sv", 30.0, 50.0)["params"]
    _same(_cp(cal.env_params["sound_speed"], chan), want["sound_speed"])
    _same(_cp(cal.env_params["sound_absorption"], chan), want["sound_absorption"])


@pytest.mark.parametrize("mode,encode", [("CW", "power"), ("CW", "complex"), ("BB", "complex")])
def test_ek80_params(mode, encode):
    ed = synth.make_ek80(C=3, P=4, R=64, B=4, mode=mode, encode=encode, gpt_channel=1 if mode == "CW" else None)
    cal = CalibrateEK80(ed, waveform_mode=mode, encode_mode=encode)
    chan = ed["Sonar/Beam_group1"]["channel"].values
    want = og.ek80(ed, "Sv", mode, encode)["params"]
    _same(_cp(cal.env_params["sound_speed"], chan), want["sound_speed"])
    _same(_cp(cal.env_params["sound_absorption"], chan), want["sound_absorption"])
    _same(_cp(cal.cal_params["sa_correction"], chan), want["sa_correction"])
    te = cal._tau_effective(encode)
    _same(te.values, want["tau_effective"], rtol=1e-12)
    gain = _cp(cal.cal_params["gain_correction"], chan)
    if mode == "BB":
        g, Bm = _bc(gain, cal._get_B_theta_phi_m())
        gain = g - Bm
        _same(_cp(cal.cal_params["equivalent_beam_angle"], chan), want["equivalent_beam_angle"])
    _same(gain, want["gain_correction"])


def test_index_binning_ping_time_is_the_tile_mean():
    """compute_MVBS_index_binning: the ping_time of a tile is the mean time of its pings (DataArray.coarsen(...).mean() with
    coord_func="mean", commongrid/api.py:219-238), the short last tile averaged over its real members; the vectorised form
    equals the per-tile float64 mean truncated to nanoseconds."""
    from echopype_b200.commongrid.api import _coarsen_time_mean

    rng = np.random.default_rng(0)
    t = np.datetime64("2024-01-01", "ns") + np.cumsum(rng.integers(1, 3_000_000_000, size=1003)).astype("timedelta64[ns]")
    ti = t.astype(np.int64)
    for pn in (1, 3, 10, 100, 1003, 2000):
        nP = -(-len(t) // pn)
        want = np.array([ti[i * pn] + int(np.mean((ti[i * pn:(i + 1) * pn] - ti[i * pn]).astype(np.float64))) for i in range(nP)])
        got = _coarsen_time_mean(t, pn).astype(np.int64)
        np.testing.assert_array_equal(got, want)
