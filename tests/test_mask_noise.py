"""clean.mask_impulse_noise / clean.mask_transient_noise with use_index_binning=True.

CPU part: the oracle against the properties the reference's own tests check (tests/clean/test_noise.py:342-447
reflection == symmetric padding; :616-682 block means; :696-768 the negated impulse condition), on mock data.
GPU part: the device masks against the oracle; samples whose comparison sits within float32 rounding of the threshold
are identified from the oracle's margin and excluded.
"""

import numpy as np
import pytest

from oracle import clean as oclean


def _mock(C=2, P=40, R=200, seed=3, nan_frac=0.02, spacing=(0.19, 0.25)):
    rng = np.random.default_rng(seed)
    Sv = rng.normal(-70.0, 6.0, size=(C, P, R))
    Sv[:, rng.integers(0, P, 6), :] += 18.0 * (rng.random((C, 6, R)) < 0.5)  # loud pings: impulse / transient candidates
    Sv[rng.random(Sv.shape) < nan_frac] = np.nan
    Sv[0, 5, R // 2 :] = np.nan  # a short ping
    depth = np.stack([np.broadcast_to(2.0 + s * np.arange(R), (P, R)) for s in np.resize(spacing, C)]).copy()
    return Sv, depth


def test_oracle_block_means_property():
    """tests/clean/test_noise.py:616-682: every upsampled value equals the manual linear-domain mean of its block."""
    Sv, depth = _mock()
    up = oclean.index_binning_downsample_upsample(Sv, depth, 2.0)
    for c, n in enumerate(oclean.samples_per_depth_bin(depth, 2.0)):
        assert n == int(np.ceil(2.0 / (depth[c, 0, 1] - depth[c, 0, 0])))
        for p in range(0, Sv.shape[1], 7):
            for b in range(-(-Sv.shape[2] // n)):
                blk = Sv[c, p, n * b : n * (b + 1)]
                want = np.nan if np.all(np.isnan(blk)) else oclean.lin2log(np.nanmean(oclean.log2lin(blk)))
                got = np.unique(up[c, p, n * b : n * (b + 1)])
                assert got.size == 1 and (np.isnan(want) and np.isnan(got[0]) or np.isclose(got[0], want, atol=1e-10, rtol=1e-10))


def test_oracle_impulse_condition_negation():
    """tests/clean/test_noise.py:696-768: where the mask is False and all three values exist, one side is <= threshold."""
    Sv, depth = _mock()
    k, thr = 2, 10.0
    mask, up = oclean.mask_impulse_noise_index_binning(Sv, depth, 2.0, k, thr)
    assert mask.any() and not mask.all()
    clean = np.where(mask, np.nan, up)
    for p in range(k, Sv.shape[1] - k):
        left, right = clean[:, p] - clean[:, p - k], clean[:, p] - clean[:, p + k]
        ok = ~(np.isnan(clean[:, p]) | np.isnan(left) | np.isnan(right))
        assert np.all((left[ok] <= thr) | (right[ok] <= thr))
    # edges: the missing side counts as +inf (np.c_ with the NaN dummy), so only the existing side decides
    assert np.array_equal(mask[:, 0], np.where(np.isnan(up[:, 0] - up[:, k]), True, up[:, 0] - up[:, k] > thr))


def test_oracle_pooled_Sv_is_symmetric_padding():
    """tests/clean/test_noise.py:342-447: generic_filter(mode="reflect") == windows over np.pad(mode="symmetric")."""
    Sv, depth = _mock(P=24, R=90)
    k, depth_bin, excl = 2, 1.0, 4.0
    pooled, m0 = oclean.index_binning_pool_Sv(Sv, depth, depth_bin, k, excl)
    assert m0 == int(np.argmin(depth[0, 0] <= excl)) and np.all(np.isnan(pooled[:, :, :m0]))
    for c, n in enumerate(oclean.samples_per_depth_bin(depth, depth_bin)):
        pad = np.pad(Sv[c][:, m0:], ((k, k), (n, n)), mode="symmetric")
        for p in range(0, Sv.shape[1], 5):
            for j in range(0, Sv.shape[2] - m0, 11):
                win = oclean.log2lin(pad[p : p + 2 * k + 1, j : j + 2 * n + 1])
                want = oclean.lin2log(np.nanmean(win))
                assert np.isclose(pooled[c, p, m0 + j], want, rtol=1e-10, atol=1e-10)


# ---- GPU --------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ep():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import echopype_b200 as ep

    return ep


def _ds(ep, Sv, depth, range_var="depth"):
    from echopype_b200.dataset import Dataset

    C, P, R = Sv.shape
    dims = ("channel", "ping_time", "range_sample")
    return Dataset({"Sv": (dims, Sv.astype(np.float32)), range_var: (dims, depth.astype(np.float32))},
                   coords={"channel": np.array([f"ch{i}" for i in range(C)], dtype=object),
                           "ping_time": np.datetime64("2024-01-01") + np.arange(P) * np.timedelta64(1, "s"), "range_sample": np.arange(R)})


@pytest.mark.gpu
@pytest.mark.parametrize("shape,depth_bin,k", [((2, 40, 200), "2m", 2), ((3, 33, 1001), "5m", 3), ((1, 7, 64), "1m", 1),
                                               # the single-pass kernel (rows in 16-sample units, blocks >= 16 samples): several
                                               # chunks per channel, k up to the chunk scale, a full-width row
                                               ((2, 300, 512), "5m", 2), ((1, 90, 4096), "7m", 5), ((3, 21, 1024), "4m", 10)])
def test_mask_impulse_noise_vs_oracle(ep, shape, depth_bin, k):
    Sv, depth = _mock(*shape)
    Sv32 = Sv.astype(np.float32).astype(np.float64)
    thr = 10.0
    want, up = oclean.mask_impulse_noise_index_binning(Sv32, depth, float(depth_bin[:-1]), k, thr)
    got = ep.clean.mask_impulse_noise(_ds(ep, Sv, depth), depth_bin, k, "10.0dB", "depth", use_index_binning=True)
    assert tuple(got.dims) == ("channel", "ping_time", "range_sample")
    g = got.values.astype(bool)
    # margin of the two comparisons (NaN difference = +inf)
    P = shape[1]
    fwd = np.full(Sv.shape, np.inf)
    bwd = np.full(Sv.shape, np.inf)
    fwd[:, : P - k] = up[:, : P - k] - up[:, k:]
    bwd[:, k:] = up[:, k:] - up[:, : P - k]
    fwd[np.isnan(fwd)] = np.inf
    bwd[np.isnan(bwd)] = np.inf
    sure = (np.abs(fwd - thr) > 1e-3) & (np.abs(bwd - thr) > 1e-3)
    assert sure.mean() > 0.99 and want.any()
    np.testing.assert_array_equal(g[sure], want[sure])


@pytest.mark.gpu
def test_impulse_block_means_within_tolerance(ep):
    import torch

    from echopype_b200 import kernels

    for P, R, db in ((25, 333, 2.0), (130, 512, 5.0)):  # two-kernel path; single-pass kernel
        Sv, depth = _mock(2, P, R)
        Sv32 = Sv.astype(np.float32)
        up = oclean.index_binning_downsample_upsample(Sv32.astype(np.float64), depth, db)
        nsamp = oclean.samples_per_depth_bin(depth, db)
        _, blocks = kernels.impulse_noise_mask(torch.from_numpy(Sv32).cuda(), nsamp, 2, P, R, 2, 10.0)
        blocks = blocks.cpu().numpy()
        for c, n in enumerate(nsamp):
            ref = up[c][:, ::n]
            got = blocks[c][:, : ref.shape[1]]
            np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
            assert np.nanmax(np.abs(got - ref)) < 1e-4  # dB, the tolerance of the path


@pytest.mark.gpu
@pytest.mark.parametrize("shape,depth_bin,k,excl", [((2, 40, 200), "2m", 3, "6m"), ((2, 30, 150), "1m", 25, "0m"), ((2, 300, 301), "2m", 25, "5m"),
                                                  ((1, 12, 90), "1m", 2, "4.0m"), ((2, 9, 64), "1m", 1, "500m")])
def test_mask_transient_noise_vs_oracle(ep, shape, depth_bin, k, excl):
    import torch

    from echopype_b200 import kernels

    Sv, depth = _mock(*shape)
    Sv32 = Sv.astype(np.float32).astype(np.float64)
    thr = 6.0
    want, pooled = oclean.mask_transient_noise_index_binning(Sv32, depth, float(depth_bin[:-1]), k, float(excl[:-1]), thr)
    got = ep.clean.mask_transient_noise(_ds(ep, Sv, depth), "nanmean", depth_bin, k, excl, "6.0dB", "depth", use_index_binning=True)
    g = got.values.astype(bool)
    with np.errstate(invalid="ignore"):
        margin = np.abs((Sv32 - pooled) - thr)
    sure = np.isnan(margin) | (margin > 1e-3)
    assert sure.mean() > 0.99
    np.testing.assert_array_equal(g[sure], want[sure])
    if float(excl[:-1]) < 100:
        assert want.any()
    # pooled Sv itself: within the tolerance of the path, identical NaN pattern
    C, P, R = shape
    nsamp = oclean.samples_per_depth_bin(depth, float(depth_bin[:-1]))
    m0 = int(np.argmin(depth <= float(excl[:-1])))
    _, pl = kernels.transient_noise_mask(torch.from_numpy(Sv.astype(np.float32)).cuda(), nsamp, C, P, R, min(m0, R), k, thr, want_pooled=True)
    pl = pl.cpu().numpy()
    np.testing.assert_array_equal(np.isnan(pl), np.isnan(pooled))
    if not np.all(np.isnan(pooled)):
        assert np.nanmax(np.abs(pl - pooled)) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("C,P,R,depth_bin,k,excl,nan_frac", [
    (2, 60, 4096, 5.0, 4, 100.0, 0.0),      # full-width strip (512 threads), no NaN: constant-count path only
    (2, 60, 4096, 5.0, 4, 100.0, 0.001),    # deficits in most windows
    (1, 70, 2048, 10.0, 25, 0.0, 0.0005),   # m0 = 0: the zero slot serves n - w - 1 = -1; k close to P / 2
    (3, 33, 1008, 3.0, 7, 31.3, 0.01),      # m0 not a multiple of 16; R / 16 not a multiple of 32 (inactive threads)
    (2, 33, 1008, 3.0, 7, 31.3, 0.0),       # the same without NaN: base correction on the constant-count path
    (1, 20, 4112, 2.0, 2, 5.0, 0.0),        # R > 4096 served because exclude_above leaves <= 4096 columns
    (1, 40, 512, 40.0, 3, 9.0, 0.003),      # long range windows (w > 128: mirror zones of several threads)
    (2, 15, 64, 4.0, 14, 2.5, 0.05),        # windows as long as the sliced axis allows, k ~ P
    (1, 400, 256, 2.0, 3, 0.0, 0.002),      # several chunks per channel (warm-up windows at chunk starts)
])
def test_transient_strip_kernel_stress(ep, C, P, R, depth_bin, k, excl, nan_frac):
    """The single-pass strip kernel (masknoise.cu transient_strip_kernel) against the oracle: conflict-free prefix layout,
    constant-count path vs deficit path, reflection at both ends of the sliced axis, unaligned exclude_above."""
    import torch

    from echopype_b200 import kernels

    rng = np.random.default_rng(R + P)
    Sv = rng.normal(-70.0, 6.0, size=(C, P, R))
    Sv[:, rng.integers(0, P, 3), :] += 18.0 * (rng.random((C, 3, R)) < 0.5)
    if nan_frac:
        Sv[rng.random(Sv.shape) < nan_frac] = np.nan
        Sv[0, P // 3, R // 2:] = np.nan  # a short ping
        Sv[-1, P // 2:P // 2 + 2, :] = np.nan  # all-NaN pings
    depth = np.stack([np.broadcast_to(1.0 + s * np.arange(R), (P, R)) for s in np.resize((0.19, 0.23, 0.31), C)]).copy()
    Sv32 = Sv.astype(np.float32).astype(np.float64)
    thr = 6.0
    want, pooled = oclean.mask_transient_noise_index_binning(Sv32, depth, depth_bin, k, excl, thr)
    nsamp = oclean.samples_per_depth_bin(depth, depth_bin)
    m0 = int(np.argmin(depth <= excl))
    assert R % 16 == 0 and R - (m0 & ~15) <= 4096  # the strip kernel's domain
    mask, pl = kernels.transient_noise_mask(torch.from_numpy(Sv.astype(np.float32)).cuda(), nsamp, C, P, R, m0, k, thr, want_pooled=True)
    mask2, _ = kernels.transient_noise_mask(torch.from_numpy(Sv.astype(np.float32)).cuda(), nsamp, C, P, R, m0, k, thr)
    assert torch.equal(mask, mask2)  # the instantiation without the pooled output
    pl, g = pl.cpu().numpy(), mask.cpu().numpy().astype(bool)
    np.testing.assert_array_equal(np.isnan(pl), np.isnan(pooled))
    assert np.nanmax(np.abs(pl - pooled)) < 1e-4
    with np.errstate(invalid="ignore"):
        margin = np.abs((Sv32 - pooled) - thr)
    sure = np.isnan(margin) | (margin > 1e-3)
    assert sure.mean() > 0.99
    np.testing.assert_array_equal(g[sure], want[sure])
    assert not g[:, :, :m0].any()


@pytest.mark.gpu
def test_mask_noise_argument_errors(ep):
    Sv, depth = _mock(1, 8, 40)
    ds = _ds(ep, Sv, depth)
    with pytest.raises(ValueError, match="`range_var` must be either `echo_range` or `depth`."):
        ep.clean.mask_impulse_noise(ds, range_var="range", use_index_binning=True)
    with pytest.raises(ValueError, match="must be `nanmean` or `nanmedian`"):
        ep.clean.mask_transient_noise(ds, func="mean", use_index_binning=True)
    with pytest.raises(ValueError, match="requires `echo_range` data variable"):
        ep.clean.mask_transient_noise(ds, range_var="echo_range", use_index_binning=True)


@pytest.mark.gpu
@pytest.mark.parametrize("index_binning", [True, False])
def test_mask_transient_noise_nanmedian_vs_oracle(ep, index_binning):
    """func="nanmedian" (clean/api.py:132-145): the pooled value is np.nanmedian of the window's linear values - a radix
    select per sample on the device - for index windows (generic_filter, reflected) and for depth-value windows."""
    import torch

    from echopype_b200 import kernels

    thr = 4.0
    if index_binning:
        Sv, depth = _mock(2, 21, 96, seed=5)
        k, depth_bin, excl = 3, "1m", "4m"
        Sv32 = Sv.astype(np.float32).astype(np.float64)
        want, pooled = oclean.mask_transient_noise_index_binning(Sv32, depth, 1.0, k, 4.0, thr, func=np.nanmedian)
    else:
        Sv, _ = _mock(2, 14, 70, seed=21)
        rng = np.random.default_rng(21)
        off = rng.choice([0.0, 0.07, 0.13], size=(2, 14))
        depth = (3.0 + off[:, :, None] + 0.19 * np.arange(70)[None, None, :]).astype(np.float32).astype(np.float64)
        depth[0, 5, 35:] = np.nan
        k, depth_bin, excl = 2, "2m", "4m"
        Sv32 = Sv.astype(np.float32).astype(np.float64)
        want, pooled = oclean.mask_transient_noise_depth_binning(Sv32, depth, 2.0, k, 4.0, thr, func=np.nanmedian)
    got = ep.clean.mask_transient_noise(_ds(ep, Sv, depth), "nanmedian", depth_bin, k, excl, "4.0dB", "depth", use_index_binning=index_binning)
    g = got.values.astype(bool)
    with np.errstate(invalid="ignore"):
        margin = np.abs((Sv32 - pooled) - thr)
    sure = np.isnan(margin) | (margin > 1e-3)
    assert sure.mean() > 0.99 and (~np.isnan(pooled)).any()
    np.testing.assert_array_equal(g[sure], want[sure])
    C, P, R = Sv.shape
    Svt = torch.from_numpy(Sv.astype(np.float32)).cuda()
    if index_binning:
        nsamp = oclean.samples_per_depth_bin(depth, 1.0)
        m0 = int(np.argmin(depth <= 4.0))
        _, pl = kernels.transient_noise_mask_median(Svt, nsamp, C, P, R, min(m0, R), k, thr, want_pooled=True)
    else:
        _, pl = kernels.transient_noise_mask_depth_median(Svt, torch.from_numpy(depth.astype(np.float32)).cuda(), C, P, R, np.nanmin(depth),
                                                          np.nanmax(depth), 2.0, 4.0, k, thr, want_pooled=True)
    pl = pl.cpu().numpy()
    np.testing.assert_array_equal(np.isnan(pl), np.isnan(pooled))
    assert np.nanmax(np.abs(pl - pooled)) < 1e-4


# ---- use_index_binning=False: intervals of depth VALUES (clean/utils.py:192-260) --------------------------------------
def _mock_depth_varying(C=2, P=30, R=220, seed=11, depth_bin=2.0):
    """Mock volume whose transducer depth changes from ping to ping (so the intervals start at different samples in
    neighbouring pings), sized so that every ping still reaches every depth interval as the reference requires."""
    Sv, _ = _mock(C, P, R, seed=seed)
    rng = np.random.default_rng(seed)
    off = rng.choice([0.0, 0.07, 0.13], size=(C, P))
    for Rr in range(R, R + 40):
        depth = (3.0 + off[:, :, None] + 0.19 * np.arange(Rr)[None, None, :]).astype(np.float32).astype(np.float64)
        edges = np.arange(depth.min(), depth.max() + depth_bin, depth_bin)
        if depth.max(axis=2).min() - edges[-2] > 0.3:  # every ping has samples in the last interval
            Sv2, _ = _mock(C, P, Rr, seed=seed)
            return Sv2, depth
    raise AssertionError("no suitable size")


def test_oracle_depth_binning_property():
    """tests/clean/test_noise.py:550-612 in spirit: every upsampled value is the linear mean of the samples of its own
    depth interval in its own ping."""
    Sv, depth = _mock_depth_varying()
    down, up = oclean.downsample_upsample_along_depth(Sv, depth, 2.0)
    edges = np.arange(depth.min(), depth.max() + 2.0, 2.0)
    for c in range(Sv.shape[0]):
        for p in range(0, Sv.shape[1], 4):
            b = np.digitize(depth[c, p], edges[:-1]) - 1
            for n in range(0, Sv.shape[2], 9):
                m = (depth[c, p] >= edges[b[n]]) & (depth[c, p] < edges[b[n] + 1]) & ~np.isnan(Sv[c, p])
                want = oclean.lin2log(oclean.log2lin(Sv[c, p][m]).mean()) if m.any() else np.nan
                assert (np.isnan(want) and np.isnan(up[c, p, n])) or np.isclose(up[c, p, n], want, rtol=1e-10, atol=1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("k,depth_bin", [(2, "2m"), (1, "5m")])
def test_mask_impulse_noise_depth_binning_vs_oracle(ep, k, depth_bin):
    import torch

    from echopype_b200 import kernels

    db = float(depth_bin[:-1])
    Sv, depth = _mock_depth_varying(depth_bin=db)
    Sv32 = Sv.astype(np.float32).astype(np.float64)
    thr = 10.0
    want, up = oclean.mask_impulse_noise_depth_binning(Sv32, depth, db, k, thr)
    got = ep.clean.mask_impulse_noise(_ds(ep, Sv, depth), depth_bin, k, "10.0dB", "depth")  # the reference's default path
    g = got.values.astype(bool)
    P = Sv.shape[1]
    fwd = np.full(Sv.shape, np.inf)
    bwd = np.full(Sv.shape, np.inf)
    fwd[:, : P - k] = up[:, : P - k] - up[:, k:]
    bwd[:, k:] = up[:, k:] - up[:, : P - k]
    fwd[np.isnan(fwd)] = np.inf
    bwd[np.isnan(bwd)] = np.inf
    sure = (np.abs(fwd - thr) > 1e-3) & (np.abs(bwd - thr) > 1e-3)
    assert sure.mean() > 0.99 and want.any() and not want.all()
    np.testing.assert_array_equal(g[sure], want[sure])
    # the interval means and first samples themselves
    C, P, R = Sv.shape
    edges = np.arange(depth.min(), depth.max() + db, db)
    _, means, first, upd = kernels.impulse_noise_mask_depth(torch.from_numpy(Sv.astype(np.float32)).cuda(),
                                                       torch.from_numpy(depth.astype(np.float32)).cuda(), edges, C, P, R, k, thr)
    down, _ = oclean.downsample_upsample_along_depth(Sv32, depth, db)
    means = means.cpu().numpy()
    np.testing.assert_array_equal(np.isnan(means), np.isnan(down))
    assert np.nanmax(np.abs(means - down)) < 1e-4
    want_first = np.stack([[np.searchsorted(depth[c, p], edges[:-1], side="left") for p in range(P)] for c in range(C)])
    np.testing.assert_array_equal(first.cpu().numpy(), want_first)
    upd = upd.cpu().numpy()
    np.testing.assert_array_equal(np.isnan(upd), np.isnan(up))
    assert np.nanmax(np.abs(upd - up)) < 1e-4  # upsampled_Sv of the reference, per sample


@pytest.mark.gpu
@pytest.mark.parametrize("k,depth_bin,excl", [(2, "2m", "4m"), (3, "1m", "0m")])
def test_mask_transient_noise_depth_windows_vs_oracle(ep, k, depth_bin, excl):
    """The reference's default path (use_index_binning=False): windows of depth values, clean/utils.py:28-105."""
    import torch

    from echopype_b200 import kernels

    db, ex = float(depth_bin[:-1]), float(excl[:-1])
    Sv, _ = _mock(2, 14, 70, seed=21)
    rng = np.random.default_rng(21)
    off = rng.choice([0.0, 0.07, 0.13], size=(2, 14))
    depth = (3.0 + off[:, :, None] + 0.19 * np.arange(70)[None, None, :]).astype(np.float32).astype(np.float64)
    depth[0, 5, 35:] = np.nan  # the short ping of the mock: NaN range where the samples are padding
    Sv32 = Sv.astype(np.float32).astype(np.float64)
    thr = 3.0
    want, pooled = oclean.mask_transient_noise_depth_binning(Sv32, depth, db, k, ex, thr)
    got = ep.clean.mask_transient_noise(_ds(ep, Sv, depth), "nanmean", depth_bin, k, excl, "3.0dB", "depth")  # defaults otherwise
    g = got.values.astype(bool)
    with np.errstate(invalid="ignore"):
        margin = np.abs((Sv32 - pooled) - thr)
    sure = np.isnan(margin) | (margin > 1e-3)
    assert sure.mean() > 0.99 and want.any() and (~np.isnan(pooled)).any() and np.isnan(pooled[:, :k]).all()
    np.testing.assert_array_equal(g[sure], want[sure])
    C, P, R = Sv.shape
    _, pl = kernels.transient_noise_mask_depth(torch.from_numpy(Sv.astype(np.float32)).cuda(), torch.from_numpy(depth.astype(np.float32)).cuda(),
                                               C, P, R, np.nanmin(depth), np.nanmax(depth), db, ex, k, thr, want_pooled=True)
    pl = pl.cpu().numpy()
    np.testing.assert_array_equal(np.isnan(pl), np.isnan(pooled))
    assert np.nanmax(np.abs(pl - pooled)) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("C,P,R,k,depth_bin,excl,nan_frac", [
    (2, 30, 256, 2, 2.0, 4.0, 0.02),     # NaN samples: deficit counts; pings near both ends of the axis are NaN
    (1, 12, 512, 5, 1.0, 0.0, 0.0),      # p + k == P: the window loses its last ping; no NaN: constant counts
    (2, 9, 64, 6, 3.0, 5.0, 0.05),       # 2 k + 1 > P: no ping qualifies, everything NaN / False
    (1, 40, 4096, 3, 10.0, 100.0, 0.001),  # full-width rows
])
def test_transient_depth_windows_uniform_rows(ep, C, P, R, k, depth_bin, excl, nan_frac):
    """Depth-value windows on volumes whose depth rows do not change with the ping (the single-pass strip kernel with
    per-column interval tables, masknoise.cu kDepth) against the oracle's pool_Sv (clean/utils.py:28-105), and against the
    per-sample bisection kernels on the same input."""
    import torch

    from echopype_b200 import kernels

    rng = np.random.default_rng(C * 1000 + R)
    Sv = rng.normal(-70.0, 6.0, size=(C, P, R))
    Sv[:, rng.integers(0, P, 3), :] += 12.0 * (rng.random((C, 3, R)) < 0.5)
    if nan_frac:
        Sv[rng.random(Sv.shape) < nan_frac] = np.nan
        Sv[0, P // 3, R // 2:] = np.nan
    spacing = np.resize((0.19, 0.23), C)
    depth = np.ascontiguousarray(np.stack([np.broadcast_to((2.0 + s * np.arange(R)).astype(np.float32), (P, R)) for s in spacing]), dtype=np.float64)
    Sv32 = Sv.astype(np.float32).astype(np.float64)
    thr = 3.0
    Svt, dt = torch.from_numpy(Sv.astype(np.float32)).cuda(), torch.from_numpy(np.ascontiguousarray(depth, dtype=np.float32)).cuda()
    assert dt.is_contiguous()
    if nan_frac:  # a short ping: no depth where the samples are padding (calibrate/range.py:143-148)
        depth[0, P // 3, R // 2:] = np.nan
        dt[0, P // 3, R // 2:] = float("nan")
    flag = torch.empty(1, dtype=torch.int32, device="cuda")
    ref = torch.empty((2, C, R), dtype=torch.float32, device="cuda")
    kernels._lib.call("epb_depth_rows_uniform", kernels.ptr(dt), kernels.ptr(Svt), kernels.ptr(ref), kernels.ptr(flag), C, P, R, kernels.stream())
    assert int(flag.item()) == 0
    mask, pl = kernels.transient_noise_mask_depth(Svt, dt, C, P, R, np.nanmin(depth), np.nanmax(depth), depth_bin, excl, k, thr, want_pooled=True)
    mask2, _ = kernels.transient_noise_mask_depth(Svt, dt, C, P, R, np.nanmin(depth), np.nanmax(depth), depth_bin, excl, k, thr)
    assert torch.equal(mask, mask2)
    # the general kernels on the same volume (a perturbed copy of one depth value makes the rows non-uniform)
    d2 = dt.clone()
    d2[0, P - 1, R - 1] = torch.nextafter(d2[0, P - 1, R - 1], torch.tensor(float("inf"), device="cuda"))
    kernels._lib.call("epb_depth_rows_uniform", kernels.ptr(d2), kernels.ptr(Svt), kernels.ptr(ref), kernels.ptr(flag), C, P, R, kernels.stream())
    assert int(flag.item()) != 0
    if R <= 512:
        want, pooled = oclean.mask_transient_noise_depth_binning(Sv32, depth, depth_bin, k, excl, thr)
        d3 = dt.clone()  # a sample with an Sv but no depth: not uniform
        d3[C - 1, 0, 3] = float("nan")
        Sv3 = Svt.clone()
        Sv3[C - 1, 0, 3] = -60.0
        kernels._lib.call("epb_depth_rows_uniform", kernels.ptr(d3), kernels.ptr(Sv3), kernels.ptr(ref), kernels.ptr(flag), C, P, R, kernels.stream())
        assert int(flag.item()) != 0
    else:  # the oracle's Python loops need minutes here: the per-sample bisection kernels stand in (themselves oracle-tested)
        mg, pg = kernels.transient_noise_mask_depth(Svt, d2, C, P, R, np.nanmin(depth), np.nanmax(depth), depth_bin, excl, k, thr, want_pooled=True)
        pooled = pg.cpu().numpy().astype(np.float64)
        pooled[0, P - 1 - k:, :] = np.nan  # rows whose windows see the perturbed value: not compared
        with np.errstate(invalid="ignore"):
            want = (Sv32 - pooled) > thr
    pl, g = pl.cpu().numpy(), mask.cpu().numpy().astype(bool)
    cmp = np.ones(pooled.shape, bool)
    if R > 512:
        cmp[0, P - 1 - k:, :] = False
    np.testing.assert_array_equal(np.isnan(pl)[cmp], np.isnan(pooled)[cmp])
    if (~np.isnan(pooled[cmp])).any():
        assert np.nanmax(np.abs(pl - pooled)[cmp]) < 1e-4
    with np.errstate(invalid="ignore"):
        margin = np.abs((Sv32 - pooled) - thr)
    sure = (np.isnan(margin) | (margin > 1e-3)) & cmp
    np.testing.assert_array_equal(g[sure], want[sure])


@pytest.mark.gpu
@pytest.mark.parametrize("C,P,R,k,db", [(2, 47, 256, 2, 2.0), (1, 120, 1024, 3, 5.0), (2, 12, 4096, 1, 5.0), (1, 9, 64, 4, 1.0)])
def test_impulse_depth_single_pass_equals_two_kernel_path(ep, C, P, R, k, db):
    """mask_impulse_noise on its default (depth-value) path without the upsampled array (impulse_depth_fused_kernel: one
    pass, interval starts and means of the window pings in a shared-memory ring) against the two-kernel path that
    materialises upsampled_Sv (itself checked against the oracle above): same interval starts, same mask, means to 1e-5 dB."""
    import torch

    from echopype_b200 import kernels

    rng = np.random.default_rng(R + k)
    Sv = rng.normal(-70.0, 6.0, size=(C, P, R))
    Sv[:, rng.integers(0, P, 4), :] += 18.0 * (rng.random((C, 4, R)) < 0.5)
    Sv[rng.random(Sv.shape) < 0.02] = np.nan
    off = rng.choice([0.0, 0.07, 0.13], size=(C, P))  # the transducer depth changes from ping to ping
    depth = (3.0 + off[:, :, None] + 0.19 * np.arange(R)[None, None, :]).astype(np.float32)
    Sv[0, P // 3, R // 2:] = np.nan
    depth[0, P // 3, R // 2:] = np.nan  # a short ping: no depth where the samples are padding
    edges = np.arange(np.nanmin(depth), np.nanmax(depth) + db, db, dtype=np.float64)
    Svt, dt = torch.from_numpy(Sv.astype(np.float32)).cuda(), torch.from_numpy(depth).cuda()
    m2, u2, f2, up = kernels.impulse_noise_mask_depth(Svt, dt, edges, C, P, R, k, 10.0)
    m1, u1, f1, none = kernels.impulse_noise_mask_depth(Svt, dt, edges, C, P, R, k, 10.0, want_upsampled=False)
    assert none is None and up is not None
    assert torch.equal(f1, f2)
    a, b = u1.cpu().numpy(), u2.cpu().numpy()
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
    assert np.nanmax(np.abs(a - b)) < 1e-5
    # masks: equal except where a comparison sits within float rounding of the threshold (the NaN-aware sums of the two
    # paths add in a different order)
    upn = up.cpu().numpy().astype(np.float64)
    fwd = np.full(upn.shape, np.inf)
    bwd = np.full(upn.shape, np.inf)
    fwd[:, : P - k] = upn[:, : P - k] - upn[:, k:]
    bwd[:, k:] = upn[:, k:] - upn[:, : P - k]
    fwd[np.isnan(fwd)] = np.inf
    bwd[np.isnan(bwd)] = np.inf
    sure = (np.abs(fwd - 10.0) > 1e-4) & (np.abs(bwd - 10.0) > 1e-4)
    assert sure.mean() > 0.999
    g1, g2 = m1.cpu().numpy().astype(bool), m2.cpu().numpy().astype(bool)
    np.testing.assert_array_equal(g1[sure], g2[sure])
    assert g2.any() and not g2.all()


# ---- clean.mask_attenuated_signal (clean/api.py:269-359) ---------------------------------------------------------------
def _host_ds(Sv, depth, range_var="depth"):
    from echopype_b200.dataset import Dataset

    C, P, R = Sv.shape
    dims = ("channel", "ping_time", "range_sample")
    return Dataset({"Sv": (dims, Sv.astype(np.float32)), range_var: (dims, depth.astype(np.float32))},
                   coords={"channel": np.array([f"ch{i}" for i in range(C)], dtype=object),
                           "ping_time": np.datetime64("2024-01-01") + np.arange(P) * np.timedelta64(1, "s"), "range_sample": np.arange(R)})


def test_mask_attenuated_signal_argument_errors():
    """tests/clean/test_noise.py:46-68 (range variable missing), :772-789 (upper limit below lower limit): ValueError
    before any device work."""
    import echopype_b200 as ep

    Sv, depth = _mock(1, 8, 32)
    with pytest.raises(ValueError, match="`range_var` must be either `echo_range` or `depth`."):
        ep.clean.mask_attenuated_signal(_host_ds(Sv, depth), range_var="range")
    with pytest.raises(ValueError, match="Masking attenuated signal requires `depth` data variable in `ds_Sv`."):
        ep.clean.mask_attenuated_signal(_host_ds(Sv, depth, "echo_range"))
    with pytest.raises(ValueError, match="Minimum range has to be shorter than maximum range"):
        ep.clean.mask_attenuated_signal(_host_ds(Sv, depth), upper_limit_sl="180m", lower_limit_sl="170m")
    with pytest.raises(ValueError, match="Decibal string"):
        ep.clean.mask_attenuated_signal(_host_ds(Sv, depth), attenuation_signal_threshold="8")


def test_oracle_attenuated_properties():
    """Whole pings only; nothing within num_side_pings of the ends; a ping 20 dB below its neighbours in the layer is
    masked with a negative threshold and its neighbours are not (the block median hardly moves)."""
    P, R, n = 30, 120, 4
    depth = np.broadcast_to(1.0 + 0.5 * np.arange(R), (1, P, R)).copy()
    Sv = np.full((1, P, R), -60.0) + np.random.default_rng(0).normal(0, 0.1, (1, P, R))
    Sv[0, 12] -= 20.0
    m = oclean.mask_attenuated_signal(Sv, depth, 20.0, 40.0, n, -6.0)
    assert m[0, 12].all() and m.sum() == R
    m8 = oclean.mask_attenuated_signal(Sv, depth, 20.0, 40.0, n, 8.0)  # the reference's default sign: everything assessable
    assert m8[0, n:P - n].all() and not m8[0, :n].any() and not m8[0, P - n:].any()
    assert not oclean.mask_attenuated_signal(Sv, depth, 400.0, 500.0, n, 8.0).any()  # outside the searching range


@pytest.mark.gpu
@pytest.mark.parametrize("C,P,R,n,upper,lower,thr", [
    (2, 80, 512, 5, 20.0, 60.0, -2.0),       # staged windows (10 x ~210 / ~160 keys)
    (1, 40, 2048, 10, 100.0, 250.0, -1.0),   # 20 x 789 keys = 8 staged steps: the candidates are compacted twice
    (1, 64, 4096, 15, 100.0, 400.0, -1.5),   # 30 x 1579 keys > staging buffer: selection from global memory
    (2, 33, 203, 2, 10.0, 30.0, -3.0),       # R not a multiple of 16: byte stores
    (3, 41, 256, 1, 10.0, 10.3, 0.5),        # one- or two-sample layers
    (1, 20, 128, 0, 15.0, 20.0, 8.0),        # num_side_pings = 0: empty block, nothing masked
    (1, 6, 16384, 1, 100.0, 2400.0, 0.2),    # block staged (24 210 keys), the ping's copy does not fit behind it: from global
    (1, 7, 32768, 1, 100.0, 6000.0, 0.2),    # layer longer than the staging buffer: both selections from global memory
    (1, 9, 64, 6, 5.0, 9.0, 8.0),            # no ping has num_side_pings neighbours on both sides
])
def test_mask_attenuated_signal_vs_oracle(ep, C, P, R, n, upper, lower, thr):
    Sv, depth = _mock(C, P, R, seed=R + n)
    rng = np.random.default_rng(P)
    Sv[:, rng.integers(0, P, max(2, P // 6)), :] -= 8.0  # attenuated pings
    depth[0, 5, R // 2:] = np.nan  # the short ping has no depth where it has no samples: np.argmin lands on the first NaN
    Sv32, d32 = Sv.astype(np.float32).astype(np.float64), depth.astype(np.float32).astype(np.float64)
    want = oclean.mask_attenuated_signal(Sv32, d32, upper, lower, n, thr)
    got = ep.clean.mask_attenuated_signal(_ds(ep, Sv, depth), f"{upper}m", f"{lower}m", n, f"{thr}dB")
    assert tuple(got.dims) == ("channel", "ping_time", "range_sample") and got.values.shape == (C, P, R)
    g = got.values.astype(bool)
    assert np.all(g == g[:, :, :1])
    sure = np.ones((C, P), dtype=bool)
    for c in range(C):
        sure[c] = ~(np.abs(oclean.attenuated_signal_margin(Sv32[c], d32[c], upper, lower, n, thr)) < 1e-9)
    assert sure.mean() > 0.98
    np.testing.assert_array_equal(g[sure], want[sure])
    if n in (2, 5, 10, 15, 1):
        assert want.any() and not want.all()
    else:
        assert not want.any()
    # the limits the kernel found are np.argmin's
    from echopype_b200 import kernels
    import torch

    _, lim = kernels.attenuated_signal_mask(torch.from_numpy(Sv.astype(np.float32)).cuda(), torch.from_numpy(depth.astype(np.float32)).cuda(),
                                            C, P, R, upper, lower, n, thr)
    lim = lim.cpu().numpy()
    with np.errstate(invalid="ignore"):
        np.testing.assert_array_equal(lim[..., 0], np.argmin(np.abs(d32 - upper), axis=2))
        np.testing.assert_array_equal(lim[..., 1], np.argmin(np.abs(d32 - lower), axis=2))


@pytest.mark.gpu
def test_mask_attenuated_signal_outside_searching_range(ep):
    """tests/clean/test_noise.py:792-815: limits beyond the echosounder range give an all-False mask."""
    Sv, depth = _mock(2, 12, 64)
    got = ep.clean.mask_attenuated_signal(_ds(ep, Sv, depth), "1800m", "2800m", 15, "-6dB")
    assert got.values.shape == (2, 12, 64) and not got.values.any()
    assert not ep.clean.mask_attenuated_signal(_ds(ep, Sv, depth)).values.any()  # the defaults (400 m .. 500 m) as well


# ---- the selection algorithm of attenuated_ping_kernel, emulated step by step (CPU) -------------------------------------
def _emulated_staged_median(keys, scratch_capacity, quantum=2048):
    """staged_median_keys of csrc/attenuation.cu in numpy, one array operation per CTA-wide pass: order-preserving keys with
    NaN = all ones, the bit loop starting at the highest bit in which the smallest and the largest valid key differ,
    candidates compacted (in arbitrary order) once they are a quarter of the scanned array, the even-count step looking
    for the smallest key above the median in the WHOLE array.  Returns (m, ka, kb, scanned keys / array length)."""
    NAN = np.uint32(0xFFFFFFFF)
    rup = lambda n: (n + quantum - 1) // quantum * quantum  # noqa: E731
    src = np.concatenate([keys, np.full(rup(len(keys)) - len(keys), NAN, np.uint32)])
    valid = src != NAN
    m = int(valid.sum())
    if m == 0:
        return 0, None, None, 1.0
    lo, hi = int(src.min()), int(np.where(valid, src, 0).max())
    k, prefix, bit = (m - 1) >> 1, lo, -1
    if lo != hi:
        bit = int(lo ^ hi).bit_length() - 1
        prefix = lo & ~(((2 << bit) - 1) & 0xFFFFFFFF) & 0xFFFFFFFF
    cur, cand, free, scanned = src, m, 0, len(src)
    rng = np.random.default_rng(len(keys))
    while bit >= 0:
        test = (0xFFFFFFFF << bit) & 0xFFFFFFFF
        match = ((cur.astype(np.uint64) ^ prefix) & test) == 0
        z = int(match.sum())
        scanned += len(cur)
        if k >= z:
            prefix |= 1 << bit
            k -= z
            cand -= z
        else:
            cand = z
        if bit > 0 and len(cur) // quantum > 1 and cand * 4 <= len(cur) and free + rup(cand) <= scratch_capacity:
            take = (((cur.astype(np.uint64) ^ prefix) & test) == 0) & (cur != NAN)
            got = cur[take]
            assert len(got) == cand  # the padding written behind the copies relies on this count
            rng.shuffle(got)
            scanned += len(cur)
            cur = np.concatenate([got, np.full(rup(cand) - cand, NAN, np.uint32)])
            free += len(cur)
        bit -= 1
    ka = kb = prefix
    if m % 2 == 0:
        eq = int((cur == prefix).sum())
        above = src[(src > prefix) & (src != NAN)]
        if k + 1 >= eq:
            kb = int(above.min())
    return m, ka, kb, scanned / len(src)


def _to_key(v):
    u = np.asarray(v, np.float32).view(np.uint32)
    k = np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint32)
    k[np.isnan(v)] = np.uint32(0xFFFFFFFF)
    return k


def _from_key(k):
    k = np.uint32(k)
    u = (k ^ np.uint32(0x80000000)) if (k & np.uint32(0x80000000)) else np.uint32(~k)
    return np.array([u], np.uint32).view(np.float32)[0]


def test_attenuated_selection_algorithm_emulation():
    """The radix selection with its two short cuts returns the two middle elements of the sorted valid values for clustered
    Sv, duplicates, a single repeated value, both signs, +-inf / +-0 and every NaN fraction; on a default-sized window of
    clustered Sv (30 pings x 526 samples) it scans about a third of what 34 full passes would."""
    rng = np.random.default_rng(0)
    work = []
    for trial in range(240):
        n = int(rng.choice([1, 2, 3, 5, 17, 100, 526, 2048, 2049, 5000, 15780, 16384]))
        kind = trial % 6
        v = [rng.normal(-70, 6, n), rng.normal(-70, 6, n).round(0), np.full(n, -66.5), rng.normal(0, 50, n),
             rng.choice([-np.inf, np.inf, -70.0, -60.0, 0.0, -0.0], n), rng.normal(-64, 3, n)][kind].astype(np.float32)
        v[rng.random(n) < rng.choice([0, 0.02, 0.5, 1.0])] = np.nan
        m, ka, kb, w = _emulated_staged_median(_to_key(v), 24576 - (n + 2047) // 2048 * 2048)
        ok = np.sort(v[~np.isnan(v)])
        assert m == len(ok)
        if m == 0:
            continue
        assert _from_key(ka) == ok[(m - 1) // 2] and _from_key(kb) == ok[m // 2], (trial, n, kind)
        if n == 15780 and kind in (0, 5) and m > 15000:
            work.append(w)
    assert work and max(work) < 16.0  # against 1 + 32 (+ 1) full passes of the plain loop
