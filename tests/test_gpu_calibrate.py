"""GPU parity: compute_Sv / compute_TS through the public API (ctypes C-ABI -> sm_100a kernels) against the
float64 CPU oracle on the same synthetic inputs.  Tolerance (north_star): Sv within 1e-4 dB, identical NaN
masks; echo_range within 1 float32 ulp (relative 1.2e-7) of the float64 reference."""

import numpy as np
import pytest

import oracle_glue as og

pytestmark = pytest.mark.gpu

SV_ATOL = 1e-4  # dB, float32 device path vs float64 oracle
RANGE_RTOL = 1.3e-7


@pytest.fixture(scope="module")
def ep():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import echopype_b200 as ep

    return ep


def _check_range(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(want)), "echo_range NaN mask differs"
    ok = ~np.isnan(want)
    err = np.abs(got[ok] - want[ok])
    assert (err <= RANGE_RTOL * np.abs(want[ok]) + 1e-12).all(), f"echo_range max rel err {np.max(err / np.maximum(np.abs(want[ok]), 1e-30)):.3e}"


@pytest.mark.parametrize("cal_type", ["Sv", "TS"])
@pytest.mark.parametrize("shape,time_varying", [((4, 64, 1000), False), ((3, 37, 515), True), ((2, 5, 4096), True), ((1, 1, 7), False)])
def test_ek60_power(ep, cal_type, shape, time_varying):
    from echopype_b200 import synth

    C, P, R = shape
    ed = synth.make_ek60(C, P, R, seed=11 + R, nan_tail=0.2, time_varying=time_varying)
    fn = ep.calibrate.compute_Sv if cal_type == "Sv" else ep.calibrate.compute_TS
    ds = fn(ed)
    want = og.ek60(ed, cal_type)
    assert ds[cal_type].dims == ("channel", "ping_time", "range_sample")
    assert ds[cal_type].on_device and ds[cal_type].dtype.__str__().endswith("float32")
    og.compare_db(ds[cal_type].values, want["out"], SV_ATOL, cal_type)
    _check_range(ds["echo_range"].values, want["echo_range"])
    if R > 3:
        assert np.isnan(ds[cal_type].values[:, :, :3]).all()  # R' <= 0 for n in {0,1,2} (calibrate_ek.py:107)


@pytest.mark.parametrize("cal_type", ["Sv", "TS"])
@pytest.mark.parametrize("shape", [(4, 50, 2048), (2, 9, 333)])
def test_azfp(ep, cal_type, shape):
    from echopype_b200 import synth

    ed = synth.make_azfp(*shape, seed=5)
    env = {"salinity": 30.0, "pressure": 50.0}
    fn = ep.calibrate.compute_Sv if cal_type == "Sv" else ep.calibrate.compute_TS
    ds = fn(ed, env_params=env)
    want = og.azfp(ed, cal_type, 30.0, 50.0)
    og.compare_db(ds[cal_type].values, want["out"], SV_ATOL, cal_type)
    _check_range(ds["echo_range"].values, want["echo_range"])


@pytest.mark.parametrize("cal_type", ["Sv", "TS"])
@pytest.mark.parametrize("gpt", [None, 1])
def test_ek80_cw_power(ep, cal_type, gpt):
    from echopype_b200 import synth

    ed = synth.make_ek80(C=3, P=21, R=640, mode="CW", encode="power", gpt_channel=gpt, nan_tail=0.2, seed=77)
    fn = ep.calibrate.compute_Sv if cal_type == "Sv" else ep.calibrate.compute_TS
    ds = fn(ed, waveform_mode="CW", encode_mode="power")
    want = og.ek80(ed, cal_type, "CW", "power")
    og.compare_db(ds[cal_type].values, want["out"], SV_ATOL, cal_type)
    _check_range(ds["echo_range"].values, want["echo_range"])
    if cal_type == "Sv":
        np.testing.assert_allclose(ds["tau_effective"].values, want["tau_effective"], rtol=1e-12)


@pytest.mark.parametrize("cal_type", ["Sv", "TS"])
@pytest.mark.parametrize("B", [4, 3, 1])
def test_ek80_cw_complex(ep, cal_type, B):
    from echopype_b200 import synth

    ed = synth.make_ek80(C=2, P=13, R=500, B=B, mode="CW", encode="complex", gpt_channel=None, nan_tail=0.2, seed=78)
    fn = ep.calibrate.compute_Sv if cal_type == "Sv" else ep.calibrate.compute_TS
    ds = fn(ed, waveform_mode="CW", encode_mode="complex")
    want = og.ek80(ed, cal_type, "CW", "complex")
    og.compare_db(ds[cal_type].values, want["out"], SV_ATOL, cal_type)
    _check_range(ds["echo_range"].values, want["echo_range"])


def test_output_contract(ep):
    from echopype_b200 import synth

    ed = synth.make_ek60(2, 8, 64)
    ds = ep.calibrate.compute_Sv(ed)
    for name in ("Sv", "echo_range", "tau_effective", "frequency_nominal", "sound_speed", "sound_absorption",
                 "sa_correction", "gain_correction", "equivalent_beam_angle", "source_filenames", "water_level"):
        assert name in ds, name
    assert ds["Sv"].attrs["long_name"] == "Volume backscattering strength (Sv re 1 m-1)"
    assert ds["Sv"].attrs["units"] == "dB"
    assert ds["echo_range"].attrs == {"long_name": "Range distance", "units": "m"}
    assert ds.attrs["processing_function"] == "calibrate.compute_Sv"
    ts = ep.calibrate.compute_TS(ed)
    assert "TS" in ts and "tau_effective" not in ts


def test_argument_errors(ep):
    from echopype_b200 import synth

    ed80 = synth.make_ek80(C=1, P=2, R=16, mode="CW", encode="power")
    with pytest.raises(ValueError, match="waveform_mode and encode_mode must be specified"):
        ep.calibrate.compute_Sv(ed80)
    with pytest.raises(ValueError, match="must be recorded as complex samples"):
        ep.calibrate.compute_Sv(ed80, waveform_mode="BB", encode_mode="power")
    ed60 = synth.make_ek60(1, 2, 16)
    with pytest.raises(ValueError, match="assume_single_filter_time can only be used on complex EK80 data."):
        ep.calibrate.compute_Sv(ed60, assume_single_filter_time=True)
    with pytest.raises(ReferenceError):
        ep.calibrate.compute_Sv(synth.make_azfp(1, 2, 16))
