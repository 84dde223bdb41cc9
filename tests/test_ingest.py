"""Raw power ingest: int16 counts -> float32 backscatter_r, and the fused pipeline on counts.

CPU part: the oracle against golden vectors made by executing the reference (tests/golden/make_golden_ingest.py).
GPU part: epb_ingest_power_i16 bit-exact against the oracle; the fused kernel on int16 counts against the same
kernel on the float32 image of the counts (identical member counts, sums within float64 atomic-order noise).
"""

import os

import numpy as np
import pytest

from oracle import convert as oconv

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ingest_vectors.npz"))


def _pings(name):
    lens, flat = GOLD[f"{name}_lens"], GOLD[f"{name}_counts"]
    offs = np.concatenate([[0], np.cumsum(lens)])
    return [flat[offs[i] : offs[i + 1]] for i in range(len(lens))]


@pytest.mark.parametrize("name", ["ragged", "equal"])
def test_oracle_matches_reference_padding_and_scaling(name):
    pings = _pings(name)
    assert oconv.INDEX2POWER == float(GOLD["INDEX2POWER"])
    ref = GOLD[f"{name}_power"]  # reference: pad_shorter_ping + astype(float32) * INDEX2POWER (float64)
    got = oconv.power_from_counts(oconv.pad_shorter_ping(pings))
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
    np.testing.assert_array_equal(got[~np.isnan(ref)], ref[~np.isnan(ref)])
    # the int16 ingest format carries the same information; its float32 image is the reference value rounded once
    img = oconv.ingest_power_i16(oconv.pack_counts(pings))
    np.testing.assert_array_equal(np.isnan(img), np.isnan(ref))
    np.testing.assert_array_equal(img[~np.isnan(ref)], ref[~np.isnan(ref)].astype(np.float32))


def test_hi_lo_split_reproduces_float64_product_for_every_count():
    """The device conversion fmaf(c, hi, c * lo) (epb_common.cuh count_to_db_f), emulated in float64."""
    hi = np.float32(oconv.INDEX2POWER)
    lo = np.float32(oconv.INDEX2POWER - np.float64(hi))
    c = np.arange(-32767, 32768).astype(np.float64)
    cl = (c.astype(np.float32) * lo).astype(np.float32)
    got = (c * np.float64(hi) + cl.astype(np.float64)).astype(np.float32)
    np.testing.assert_array_equal(got, (c * oconv.INDEX2POWER).astype(np.float32))


# ---- GPU --------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ep():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import echopype_b200 as ep

    return ep


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 7, 8, 4099, 65536 + 3])
def test_ingest_kernel_bit_exact(ep, n):
    import torch

    from echopype_b200 import kernels

    rng = np.random.default_rng(n)
    c = rng.integers(-32768, 32767, size=n, endpoint=True).astype(np.int16)
    if n >= 65536:
        c[:65536] = np.arange(-32768, 32768).astype(np.int16)  # every count once, the padding marker included
    got = kernels.ingest_power_i16(torch.from_numpy(c).cuda()).cpu().numpy()
    want = oconv.ingest_power_i16(c)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_array_equal(got[~np.isnan(want)], want[~np.isnan(want)])


@pytest.mark.gpu
def test_golden_pings_through_the_ingest_kernel(ep):
    import torch

    from echopype_b200 import kernels

    ref = GOLD["ragged_power"]
    got = kernels.ingest_power_i16(torch.from_numpy(oconv.pack_counts(_pings("ragged"))).cuda()).cpu().numpy()
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
    np.testing.assert_array_equal(got[~np.isnan(ref)], ref[~np.isnan(ref)].astype(np.float32))


@pytest.mark.gpu
def test_synthetic_counts_are_the_counts_of_the_synthetic_power(ep):
    from echopype_b200 import kernels

    a = kernels.synth_fill((2, 37, 1000), kind=0, seed=77, nan_tail=0.3).cpu().numpy()
    q = kernels.synth_fill_i16((2, 37, 1000), seed=77, nan_tail=0.3)
    b = kernels.ingest_power_i16(q).cpu().numpy()
    assert np.isnan(a).any() and (q.cpu().numpy() == -32768).sum() == np.isnan(a).sum()
    np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))


I16_CASES = [
    # (C, P, R), time_varying, ping_num, range_sample_num, range_bin, ping_time_bin
    ((2, 203, 4096), False, 5, 30, "20m", "20s"),   # benchmark tile shape, partial last tile
    ((3, 97, 2048), False, 5, 30, "10m", "7s"),     # one column group per thread, two CTAs per SM
    ((2, 64, 520), False, 8, 16, "5m", "1min"),     # R % 8 == 0 but not a warp multiple
    ((2, 45, 1000), False, None, None, "20m", "20s"),  # Sv -> MVBS without noise removal
    ((2, 61, 2048), True, 5, 30, "10m", "7s"),      # irregular volume: ingest + general kernel through the scratch
    ((2, 50, 1004), False, 5, 30, "10m", "7s"),     # R % 8 != 0: no direct path, ingest + float pipeline
    ((2, 123, 4096), False, 10, 20, "20m", "20s"),  # ping_num > 8: two sweeps over sub-tiles, counts read twice
    ((2, 90, 1024), False, 17, 30, "10m", "7s"),    # three sub-tiles of 6
]


@pytest.mark.gpu
@pytest.mark.parametrize("shape,tv,pn,rn,rb,tb", I16_CASES)
def test_fused_pipeline_on_counts_equals_float_path(ep, shape, tv, pn, rn, rb, tb):
    import torch

    from echopype_b200 import synth

    C, P, R = shape
    rng = np.random.default_rng(5)
    q = synth._power_host(rng, C, P, R, 0.2, raw_counts=True)
    ed_f = synth.make_ek60(C, P, R, seed=5, time_varying=tv, backscatter=oconv.ingest_power_i16(q))
    ed_q = synth.make_ek60(C, P, R, seed=5, time_varying=tv, backscatter=torch.from_numpy(q).cuda())
    args = dict(ping_num=pn, range_sample_num=rn, range_bin=rb, ping_time_bin=tb, finalize=False)
    a = ep.pipeline.compute_Sv_clean_MVBS(ed_q, **args)
    b = ep.pipeline.compute_Sv_clean_MVBS(ed_f, fast=(R % 8 == 0), **args)  # odd R: the counts take the general kernel
    fa, fb = a.attrs["acc"].cpu().numpy(), b.attrs["acc"].cpu().numpy()
    assert fa.shape == fb.shape and fb[..., 1].sum() > 0
    np.testing.assert_array_equal(fa[..., 1:3], fb[..., 1:3])  # survivors / NaN members per bin: exact
    # fast kernel: the same float32 terms, float64 atomics in any order; the general kernel (irregular volumes, odd R)
    # also adds float32 partial sums in scheduling order
    rtol = 1e-12 if (not tv and R % 8 == 0) else 2e-6
    np.testing.assert_allclose(fa[..., 0], fb[..., 0], rtol=rtol)
    if pn:
        np.testing.assert_array_equal(a.attrs["noise_estimate"].values, b.attrs["noise_estimate"].values)
    # streamed from the host (2 bytes per sample over PCIe) == resident
    ed_h = synth.make_ek60(C, P, R, seed=5, time_varying=tv, backscatter=q)
    c = ep.pipeline.compute_Sv_clean_MVBS(ed_h, chunk_pings=40, **args)
    fc = c.attrs["acc"].cpu().numpy()
    np.testing.assert_array_equal(fc[..., 1:3], fb[..., 1:3])
    np.testing.assert_allclose(fc[..., 0], fb[..., 0], rtol=rtol)


@pytest.mark.gpu
@pytest.mark.parametrize("R", [504, 500, 502])  # K1 reads the counts itself (R % 8 == 0, R % 4 == 0) / ingest kernel first
def test_compute_Sv_accepts_raw_counts(ep, R):
    from echopype_b200 import synth

    rng = np.random.default_rng(9)
    q = synth._power_host(rng, 2, 30, R, 0.2, raw_counts=True)
    for fn in (ep.calibrate.compute_Sv, ep.calibrate.compute_TS):
        ds_q = fn(synth.make_ek60(2, 30, R, seed=9, backscatter=q))
        ds_f = fn(synth.make_ek60(2, 30, R, seed=9, backscatter=oconv.ingest_power_i16(q)))
        name = "Sv" if "Sv" in ds_q else "TS"
        np.testing.assert_array_equal(ds_q[name].values.view(np.uint32), ds_f[name].values.view(np.uint32))
        np.testing.assert_array_equal(ds_q["echo_range"].values.view(np.uint32), ds_f["echo_range"].values.view(np.uint32))
