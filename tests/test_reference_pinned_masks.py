"""The noise-mask oracles AND the CUDA masks against outputs of the reference's own code.

tests/golden/mask_vectors.npz holds what echopype's unmodified mask_attenuated_signal / mask_impulse_noise /
mask_transient_noise (clean/api.py:30-359) and pool_Sv, index_binning_pool_Sv,
index_binning_downsample_upsample_along_depth, echopy_impulse_noise_mask, echopy_attenuated_signal_mask
(clean/utils.py) return on small synthetic volumes; they were executed from /root/reference by
tests/golden/make_golden_masks.py.

* CPU (not gpu): oracle/clean.py reproduces those masks exactly and the pooled / upsampled Sv to 1e-9 dB - it is a
  pinned oracle for these functions.
* GPU: the product's masks are compared DIRECTLY with the reference masks; the only samples set aside are those whose
  comparison sits within float32 rounding of the threshold (identified from the reference's own pooled values).
"""

import ast
import os

import numpy as np
import pytest

from oracle import clean as oclean
from oracle.commongrid import parse_x_bin

HERE = os.path.dirname(os.path.abspath(__file__))
DIMS = ("channel", "ping_time", "range_sample")
MARGIN_DB = 2e-3  # float32 Sv / float32 window sums against the reference's float64 pooled values


@pytest.fixture(scope="module")
def vec():
    return np.load(os.path.join(HERE, "golden", "mask_vectors.npz"))


def _kw(vec, key):
    return ast.literal_eval(str(vec[f"{key}__kw"]))


def _f64(a):
    return np.asarray(a, dtype=np.float64)


ATT_KEYS = ["att_a", "att_b", "att_default_thr", "att_outside", "att_echo_range"]


# ---- CPU: the oracle is pinned ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("key", ATT_KEYS)
def test_oracle_equals_reference_attenuated_mask(vec, key):
    kw = _kw(vec, key)
    got = oclean.mask_attenuated_signal(_f64(vec[f"{key}__Sv"]), _f64(vec[f"{key}__range"]), parse_x_bin(kw["upper_limit_sl"], "range_bin"),
                                        parse_x_bin(kw["lower_limit_sl"], "range_bin"), kw["num_side_pings"],
                                        oclean.extract_dB(kw["attenuation_signal_threshold"]))
    want = vec[f"{key}__mask"]
    assert got.shape == want.shape and np.array_equal(got, want)
    if key != "att_outside":
        assert want.any() and not want.all()
        assert np.all(want == want[:, :, :1])  # whole pings


def test_oracle_equals_reference_echopy_attenuated(vec):
    Sv, rv = _f64(vec["att_fn__Sv"])[0], _f64(vec["att_fn__range"])[0]
    for i in range(4):
        u, l, n, t = vec[f"att_fn{i}__args"]
        got = oclean.echopy_attenuated_signal_mask(Sv, rv, u, l, int(n), t)
        assert np.array_equal(got, vec[f"att_fn{i}__mask"]), i
    assert vec["att_fn0__mask"].any() and not vec["att_fn3__mask"].any()  # num_side_pings = 0: an empty block, never masked


def test_oracle_equals_reference_impulse(vec):
    Sv, rv = _f64(vec["imp__Sv"]), _f64(vec["imp__range"])
    for i in range(3):
        k, t = vec[f"imp_fn{i}__args"]
        assert np.array_equal(oclean.echopy_impulse_noise_mask(Sv[0].T, int(k), t), vec[f"imp_fn{i}__mask"])
    for i in range(2):
        kw = _kw(vec, f"imp_api{i}")
        mask, up = oclean.mask_impulse_noise_index_binning(Sv, rv, parse_x_bin(kw["depth_bin"], "range_bin"), kw["num_side_pings"],
                                                           oclean.extract_dB(kw["impulse_noise_threshold"]))
        np.testing.assert_allclose(up, vec[f"imp_api{i}__upsampled"], rtol=0, atol=1e-9, equal_nan=True)
        assert np.array_equal(mask, vec[f"imp_api{i}__mask"]) and mask.any()


@pytest.mark.parametrize("which,i", [("depth", 0), ("depth", 1), ("index", 0), ("index", 1)])
def test_oracle_equals_reference_transient(vec, which, i):
    key = f"tr_{which}{i}"
    kw = _kw(vec, key)
    src = "tr" if which == "depth" else "tri"
    Sv, rv = _f64(vec[f"{src}__Sv"]), _f64(vec[f"{src}__range"])
    fn = oclean.mask_transient_noise_depth_binning if which == "depth" else oclean.mask_transient_noise_index_binning
    mask, pooled = fn(Sv, rv, parse_x_bin(kw["depth_bin"], "range_bin"), kw["num_side_pings"], parse_x_bin(kw["exclude_above"], "range_bin"),
                      oclean.extract_dB(kw["transient_noise_threshold"]), func=np.nanmean if kw["func"] == "nanmean" else np.nanmedian)
    want_pooled = vec[f"{key}__pooled"]
    assert np.array_equal(np.isnan(pooled), np.isnan(want_pooled)) and np.isfinite(want_pooled).any()
    np.testing.assert_allclose(pooled, want_pooled, rtol=0, atol=1e-9, equal_nan=True)
    assert np.array_equal(mask, vec[f"{key}__mask"]) and mask.any()


# ---- GPU: the product against the reference outputs -------------------------------------------------------------------
@pytest.fixture(scope="module")
def ep():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import echopype_b200 as ep

    return ep


def _ds(Sv, rv, range_var="depth"):
    from echopype_b200.dataset import Dataset

    C, P, R = Sv.shape
    return Dataset({"Sv": (DIMS, np.asarray(Sv, np.float32)), range_var: (DIMS, np.asarray(rv, np.float32))},
                   coords={"channel": np.array([f"ch{i}" for i in range(C)], dtype=object),
                           "ping_time": np.datetime64("2024-01-01") + np.arange(P) * np.timedelta64(1, "s"), "range_sample": np.arange(R)})


@pytest.mark.gpu
@pytest.mark.parametrize("key", ATT_KEYS)
def test_cuda_attenuated_mask_equals_reference(ep, vec, key):
    kw = _kw(vec, key)
    Sv, rv = vec[f"{key}__Sv"], vec[f"{key}__range"]
    got = ep.clean.mask_attenuated_signal(_ds(Sv, rv, kw.get("range_var", "depth")), **kw)
    assert tuple(got.dims) == DIMS
    g = got.values.astype(bool)
    want = vec[f"{key}__mask"]
    thr = oclean.extract_dB(kw["attenuation_signal_threshold"])
    sure = np.ones(want.shape[:2], dtype=bool)
    if key != "att_outside":
        for c in range(Sv.shape[0]):
            mg = oclean.attenuated_signal_margin(_f64(Sv[c]), _f64(rv[c]), parse_x_bin(kw["upper_limit_sl"], "range_bin"),
                                                 parse_x_bin(kw["lower_limit_sl"], "range_bin"), kw["num_side_pings"], thr)
            sure[c] = ~(np.abs(mg) < 1e-9)  # float64 on both sides: only last-bit ties are set aside
    assert sure.mean() > 0.95
    assert np.array_equal(g[sure], want[sure])


@pytest.mark.gpu
@pytest.mark.parametrize("i", [0, 1])
def test_cuda_impulse_mask_equals_reference(ep, vec, i):
    kw = _kw(vec, f"imp_api{i}")
    Sv, rv = vec["imp__Sv"], vec["imp__range"]
    got = ep.clean.mask_impulse_noise(_ds(Sv, rv), use_index_binning=True, **kw).values.astype(bool)
    want, up = vec[f"imp_api{i}__mask"], vec[f"imp_api{i}__upsampled"]
    k, thr, P = kw["num_side_pings"], oclean.extract_dB(kw["impulse_noise_threshold"]), Sv.shape[1]
    fwd, bwd = np.full(up.shape, np.inf), np.full(up.shape, np.inf)
    fwd[:, : P - k], bwd[:, k:] = up[:, : P - k] - up[:, k:], up[:, k:] - up[:, : P - k]
    fwd[np.isnan(fwd)], bwd[np.isnan(bwd)] = np.inf, np.inf
    sure = (np.abs(fwd - thr) > MARGIN_DB) & (np.abs(bwd - thr) > MARGIN_DB)
    assert sure.mean() > 0.98 and np.array_equal(got[sure], want[sure])


@pytest.mark.gpu
@pytest.mark.parametrize("which,i", [("depth", 0), ("depth", 1), ("index", 0), ("index", 1)])
def test_cuda_transient_mask_equals_reference(ep, vec, which, i):
    key = f"tr_{which}{i}"
    kw = _kw(vec, key)
    src = "tr" if which == "depth" else "tri"
    Sv, rv = vec[f"{src}__Sv"], vec[f"{src}__range"]
    got = ep.clean.mask_transient_noise(_ds(Sv, rv), use_index_binning=(which == "index"), **kw).values.astype(bool)
    want, pooled = vec[f"{key}__mask"], vec[f"{key}__pooled"]
    thr = oclean.extract_dB(kw["transient_noise_threshold"])
    with np.errstate(invalid="ignore"):
        margin = _f64(Sv) - pooled - thr
    sure = ~(np.abs(margin) < MARGIN_DB)  # NaN margins (no pooled value / NaN Sv) compare False on both sides: kept
    assert sure.mean() > 0.98 and np.array_equal(got[sure], want[sure])


# ---- compute_MVBS_index_binning (K7) against the executed reference function (tests/golden/make_golden_commongrid.py) ---
IB_KEYS = ["ib_a", "ib_b", "ib_c", "ib_d"]


@pytest.fixture(scope="module")
def gvec():
    return np.load(os.path.join(HERE, "golden", "commongrid_vectors.npz"))


@pytest.mark.parametrize("key", IB_KEYS)
def test_oracle_equals_reference_index_binning(gvec, key):
    from oracle import commongrid as ogrid

    rsn, pn = (int(v) for v in gvec[f"{key}__args"])
    got = ogrid.compute_MVBS_index_binning(_f64(gvec[f"{key}__Sv_in"]), _f64(gvec[f"{key}__echo_range_in"]), rsn, pn)
    want = gvec[f"{key}__Sv"]
    assert got["Sv"].shape == want.shape and np.array_equal(np.isnan(got["Sv"]), np.isnan(want))
    np.testing.assert_allclose(got["Sv"], want, rtol=0, atol=1e-11, equal_nan=True)
    # the coarsened echo_range as the reference computes it (before xarray's index alignment, see the generator's docstring)
    np.testing.assert_allclose(got["echo_range"], gvec[f"{key}__echo_range_coarsened"], rtol=1e-15, equal_nan=True)
    np.testing.assert_array_equal(got["actual_range"], gvec[f"{key}__actual_range"])
    np.testing.assert_array_equal(gvec[f"{key}__range_sample"], np.arange(want.shape[2]))


@pytest.mark.parametrize("key", IB_KEYS)
def test_host_tile_mean_ping_time_equals_coarsened_coordinate(gvec, key):
    """The ping_time coordinate of the output is the coarsened one (coord_func="mean"): the tile's mean time, NaT padding
    of a short last tile skipped.  The stored values went through a float64 mean of absolute nanoseconds (256 ns grain)
    in the labelled-array stand-in, hence the 1 us tolerance; the product averages offsets from the tile's first ping."""
    from echopype_b200.commongrid.api import _coarsen_time_mean

    pn = int(gvec[f"{key}__args"][1])
    t_in = gvec[f"{key}__ping_time_in"].astype("datetime64[ns]")
    got = _coarsen_time_mean(t_in, pn).astype(np.int64)
    want = gvec[f"{key}__ping_time"]
    assert got.shape == want.shape and np.max(np.abs(got - want)) < 1000
    if len(t_in) >= 2 * pn:
        assert not np.array_equal(got, t_in[::pn].astype(np.int64))  # it is NOT the first ping of the tile


@pytest.mark.gpu
@pytest.mark.parametrize("key", IB_KEYS)
def test_cuda_index_binning_equals_reference(ep, gvec, key):
    from echopype_b200.dataset import Dataset

    rsn, pn = (int(v) for v in gvec[f"{key}__args"])
    Sv, er = gvec[f"{key}__Sv_in"], gvec[f"{key}__echo_range_in"]
    C, P, R = Sv.shape
    ds = Dataset({"Sv": (DIMS, Sv), "echo_range": (DIMS, er), "frequency_nominal": (("channel",), 38e3 * (1 + np.arange(C)))},
                 coords={"channel": np.array([f"ch{i}" for i in range(C)], dtype=object),
                         "ping_time": gvec[f"{key}__ping_time_in"].astype("datetime64[ns]"), "range_sample": np.arange(R)})
    got = ep.commongrid.compute_MVBS_index_binning(ds, range_sample_num=rsn, ping_num=pn)
    want = gvec[f"{key}__Sv"]
    g = np.asarray(got["Sv"].values, dtype=np.float64)
    assert g.shape == want.shape and np.array_equal(np.isnan(g), np.isnan(want))
    assert np.nanmax(np.abs(g - want)) < 1e-4  # dB (north_star)
    np.testing.assert_allclose(np.asarray(got["echo_range"].values), gvec[f"{key}__echo_range_coarsened"], rtol=2e-7, equal_nan=True)
    assert np.max(np.abs(np.asarray(got["ping_time"].values).astype("datetime64[ns]").astype(np.int64) - gvec[f"{key}__ping_time"])) < 1000
    np.testing.assert_array_equal(np.asarray(got["range_sample"].values), gvec[f"{key}__range_sample"])
    np.testing.assert_allclose(got["Sv"].attrs["actual_range"], gvec[f"{key}__actual_range"], atol=0.011)


# ---- compute_MVBS: the bin grids and argument checks around the flox group-by, from the executed reference ---------------
GRID_KEYS = ["grid_a", "grid_b", "grid_c", "grid_d", "grid_e", "grid_f"]


@pytest.mark.parametrize("key", GRID_KEYS)
def test_oracle_and_host_bin_grids_equal_reference(gvec, key):
    """compute_MVBS (commongrid/api.py:31-191) was executed with only compute_raw_MVBS (the flox group-by) replaced by a recorder:
    the range grid (from range_var.max(skipna=True) or from range_var_max + 1e-8), the ping grid (pandas resample plus one closing
    edge, whatever `closed` is) and the output coordinates (interval LEFT ends, also for closed="right") are the reference's."""
    from echopype_b200.commongrid import utils as gutils
    from oracle import commongrid as ogrid

    kw = ast.literal_eval(str(gvec[f"{key}__kw"]))
    rb = parse_x_bin(kw["range_bin"], "range_bin")
    if "range_var_max" in kw:
        rmax = parse_x_bin(kw["range_var_max"], "range_bin") + 1e-8
    else:
        rmax = float(np.nanmax(_f64(gvec[f"{key}__range_in"])))
    want_r, want_p = gvec[f"{key}__range_edges"], gvec[f"{key}__ping_edges"]
    t_in = gvec[f"{key}__ping_time_in"]
    for edges_r, edges_p in [(ogrid.range_edges(rmax, rb), ogrid.ping_edges(t_in, kw["ping_time_bin"])),
                             (gutils.range_edges(rmax, gutils._parse_x_bin(kw["range_bin"])),
                              gutils.ping_time_edges(t_in.astype("datetime64[ns]"), kw["ping_time_bin"]).astype(np.int64))]:
        np.testing.assert_array_equal(np.asarray(edges_r, dtype=np.float64), want_r)
        np.testing.assert_array_equal(np.asarray(edges_p), want_p)
    np.testing.assert_array_equal(gvec[f"{key}__out_range"], want_r[:-1])
    np.testing.assert_array_equal(gvec[f"{key}__out_ping_time"], want_p[:-1])
    assert list(gvec[f"{key}__closed"]) == [kw.get("closed", "left")] * 2
    # attribute strings of the reference for this setting
    val, unit = gutils.ping_time_bin_parsing_and_conversion(kw["ping_time_bin"])
    rv = kw.get("range_var", "echo_range")
    assert str(gvec[f"{key}__cell_methods"]) == (f"ping_time: mean (interval: {val} {unit} comment: ping_time is the interval start) "
                                                  f"{rv}: mean (interval: {rb} meter comment: {rv} is the interval start)")
    assert str(gvec[f"{key}__range_meter_interval"]) == str(rb) + "m" and str(gvec[f"{key}__ping_time_interval"]) == kw["ping_time_bin"]


def test_host_compute_MVBS_rejects_what_the_reference_rejects(gvec):
    import echopype_b200 as ep
    from echopype_b200.dataset import Dataset

    C, P, R = 2, 6, 8
    ds = Dataset({"Sv": (DIMS, np.zeros((C, P, R), np.float32)), "echo_range": (DIMS, np.ones((C, P, R), np.float32)),
                  "frequency_nominal": (("channel",), np.array([38e3, 120e3]))},
                 coords={"channel": np.array(["ch0", "ch1"], dtype=object), "ping_time": np.datetime64("2024-03-01") + np.arange(P) * np.timedelta64(1, "s"),
                         "range_sample": np.arange(R)})
    kws = {"range_var": dict(range_var="range"), "missing_depth": dict(range_var="depth"), "range_bin_type": dict(range_bin=10),
           "range_bin_unit": dict(range_bin="10km"), "range_bin_nounit": dict(range_bin="10"), "closed": dict(closed="both"),
           "ping_time_bin_type": dict(ping_time_bin=10), "reindex": dict(method="cohorts", reindex=True)}
    for label, etype, msg in gvec["grid_bad__cases"]:
        label, etype, msg = str(label), str(etype), str(msg)
        assert etype != "ok"
        with pytest.raises({"ValueError": ValueError, "TypeError": TypeError}[etype]) as ei:
            ep.commongrid.compute_MVBS(ds, **kws[label])
        assert str(ei.value) == msg, (label, str(ei.value), msg)
