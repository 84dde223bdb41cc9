"""Thin typed wrappers around the C ABI (include/epb200.h): allocate outputs as torch CUDA tensors,
pass raw pointers + the current torch stream, raise on any error.  No arithmetic happens here."""

import ctypes

import numpy as np
import torch

from . import _lib
from .device import ParamPack, empty, ptr, require_cuda, stream

CAL = {"Sv": 0, "TS": 1}
SONAR_EX60, SONAR_EX80 = 0, 1


def alloc_rows(C, P, device=None):
    return torch.empty(int(C) * int(P) * _lib.ROW_BYTES, dtype=torch.uint8, device=device or require_cuda())


class RowBuilder:
    """Parameters of one row-setup launch, uploaded once; :meth:`build` (re)launches the epb_rows_* kernel."""

    def __init__(self, name, C, P, R, pack, args):
        self.name, self.shape, self.pack, self.args = name, (int(C), int(P), int(R)), pack, args

    def build(self, rows=None):
        C, P, R = self.shape
        rows = rows if rows is not None else alloc_rows(C, P, self.pack.device)
        _lib.call(self.name, ptr(rows), C, P, R, *self.args, stream())
        rows._keep = self.pack  # parameters must outlive the asynchronous launch
        return rows


def ek_power_row_builder(C, P, R, sonar, cal_type, prm, is_gpt=None):
    """prm: dict with sample_interval, sound_speed, sound_absorption, transmit_duration_nominal, transmit_power,
    gain_correction, sa_correction, equivalent_beam_angle, frequency_nominal, tau_effective (host arrays)."""
    dev = require_cuda()
    pk = ParamPack(C, P, dev)
    sv = cal_type == "Sv"
    null = ParamPack.null()
    gpt = pk.vec(np.asarray(is_gpt, dtype=np.uint8), torch.uint8) if is_gpt is not None else None
    args = (
        sonar, CAL[cal_type],
        pk.cp(prm["sample_interval"]), pk.cp(prm["sound_speed"]), pk.cp(prm["sound_absorption"]),
        pk.cp(prm["transmit_duration_nominal"]), pk.cp(prm["transmit_power"]), pk.cp(prm["gain_correction"]),
        pk.cp(prm["sa_correction"]) if sv else null, pk.cp(prm["equivalent_beam_angle"]) if sv else null,
        pk.cp(prm["frequency_nominal"]), pk.cp(prm["tau_effective"]) if sv else null, gpt,
    )
    return RowBuilder("epb_rows_ek_power", C, P, R, pk, args)


def rows_ek_power(C, P, R, sonar, cal_type, prm, is_gpt=None):
    return ek_power_row_builder(C, P, R, sonar, cal_type, prm, is_gpt).build()


def azfp_row_builder(C, P, R, cal_type, prm):
    dev = require_cuda()
    pk = ParamPack(C, P, dev)
    args = (
        CAL[cal_type], pk.cp(prm["sound_speed"]), pk.cp(prm["sound_absorption"]),
        pk.cp(prm["transmit_duration_nominal"]), pk.vec(prm["N"]), pk.vec(prm["f_dig"]), pk.vec(prm["L"]),
        pk.vec(prm["EL"]), pk.vec(prm["DS"]), pk.vec(prm["TVR"]), pk.vec(prm["VTX0"]),
        pk.vec(prm["equivalent_beam_angle"]), pk.vec(prm["Sv_offset"]),
    )
    return RowBuilder("epb_rows_azfp", C, P, R, pk, args)


def rows_azfp(C, P, R, cal_type, prm):
    return azfp_row_builder(C, P, R, cal_type, prm).build()


def rows_ek80_complex(C, P, R, cal_type, waveform_bb, n_beam, prm, is_gpt=None):
    dev = require_cuda()
    pk = ParamPack(C, P, dev)
    rows = alloc_rows(C, P, dev)
    sv = cal_type == "Sv"
    null = ParamPack.null()
    gpt = pk.vec(np.asarray(is_gpt, dtype=np.uint8), torch.uint8) if is_gpt is not None else None
    _lib.call(
        "epb_rows_ek80_complex", ptr(rows), C, P, R, CAL[cal_type], int(bool(waveform_bb)), int(n_beam),
        pk.cp(prm["sample_interval"]), pk.cp(prm["sound_speed"]), pk.cp(prm["sound_absorption"]),
        pk.cp(prm["transmit_duration_nominal"]), pk.cp(prm["transmit_power"]), pk.cp(prm["gain_correction"]),
        pk.cp(prm["sa_correction"]) if (sv and not waveform_bb) else null,
        pk.cp(prm["equivalent_beam_angle"]) if sv else null, pk.cp(prm["freq_center"]),
        pk.cp(prm["tau_effective"]) if sv else null, pk.cp(prm["impedance_transducer"]),
        pk.cp(prm["impedance_transceiver"]), gpt, stream(),
    )
    rows._keep = pk
    return rows


def new_minmax(device=None):
    mm = torch.empty(4, dtype=torch.float32, device=device or require_cuda())
    _lib.call("epb_minmax_init", ptr(mm), stream())
    return mm


def sv_power(x, rows, C, P, R, want_range=True, want_minmax=False, out=None, rng=None):
    out = out if out is not None else empty((C, P, R), device=x.device)
    rng = rng if rng is not None else (empty((C, P, R), device=x.device) if want_range else None)
    mm = new_minmax(x.device) if want_minmax else None
    # int16 raw power counts (ingest format) are read directly, 2 bytes per sample, when the row length allows it
    name = "epb_sv_power_i16" if x.dtype == torch.int16 else "epb_sv_power"
    _lib.call(name, ptr(x), ptr(rows), ptr(out), ptr(rng), ptr(mm), C, P, R, stream())
    return out, rng, mm


def sv_complex(re, im, rows, C, P, R, B, want_range=True, want_minmax=False):
    out = empty((C, P, R), device=re.device)
    rng = empty((C, P, R), device=re.device) if want_range else None
    mm = new_minmax(re.device) if want_minmax else None
    _lib.call("epb_sv_complex", ptr(re), ptr(im), ptr(rows), ptr(out), ptr(rng), ptr(mm), C, P, R, int(B), stream())
    return out, rng, mm


def pulse_compress_sv(re, im, replicas, rows, C, P, R, B, want_range=True, want_pc=False, want_minmax=False, method="auto"):
    """replicas: list of per-channel complex transmit replicas (host, complex128).
    method: "fft" (overlap-save FFT kernel), "direct" (tap loop) or "auto" (FFT whenever the replicas fit)."""
    dev = re.device
    offs = np.zeros(C + 1, dtype=np.int32)
    for c, tx in enumerate(replicas):
        offs[c + 1] = offs[c] + len(tx)
    cat = np.concatenate([np.asarray(tx, dtype=np.complex128) for tx in replicas])
    rep = torch.from_numpy(np.ascontiguousarray(np.stack([cat.real, cat.imag], axis=1).astype(np.float32))).to(dev)
    inv_norm = torch.from_numpy(np.asarray([1.0 / np.linalg.norm(tx) ** 2 for tx in replicas], dtype=np.float64)).to(dev)
    out = empty((C, P, R), device=dev)
    rng = empty((C, P, R), device=dev) if want_range else None
    pc = empty((C, P, R, 2), device=dev) if want_pc else None
    mm = new_minmax(dev) if want_minmax else None
    h_off = (ctypes.c_int * (C + 1))(*offs.tolist())
    lib = _lib.load()
    fits = max(len(tx) for tx in replicas) <= int(lib.epb_pulse_fft_max_taps())
    if method == "fft" and not fits:
        raise ValueError("replica too long for the FFT form")
    ws = None
    if method == "fft" or (method == "auto" and fits):
        nws = int(lib.epb_pulse_fft_workspace_bytes(int(C)))
        ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        _lib.call(
            "epb_pulse_compress_sv_fft", ptr(re), ptr(im), ptr(rep), h_off, ptr(inv_norm), ptr(rows), ptr(out), ptr(rng),
            ptr(pc), ptr(mm), C, P, R, int(B), ptr(ws), nws, stream(),
        )
    else:
        _lib.call(
            "epb_pulse_compress_sv", ptr(re), ptr(im), ptr(rep), h_off, ptr(inv_norm), ptr(rows), ptr(out), ptr(rng),
            ptr(pc), ptr(mm), C, P, R, int(B), stream(),
        )
    out._keep = (rep, inv_norm, ws)
    return out, rng, pc, mm


def noise_estimate(Sv, echo_range, alpha_cp, pack, C, P, R, ping_num, range_sample_num, noise_max=None):
    nP = -(-P // ping_num)
    noise = empty((C, nP), device=Sv.device)
    nm = float("nan") if noise_max is None else float(noise_max)
    _lib.call(
        "epb_noise_estimate", ptr(Sv), ptr(echo_range), alpha_cp, ptr(noise), C, P, R, int(ping_num),
        int(range_sample_num), ctypes.c_float(nm), stream(),
    )
    noise._keep = pack
    return noise


def noise_apply(Sv, echo_range, alpha_cp, pack, noise, C, P, R, ping_num, snr, want_noise=True, want_corr=True, want_minmax=True):
    sn = empty((C, P, R), device=Sv.device) if want_noise else None
    sc = empty((C, P, R), device=Sv.device) if want_corr else None
    mm = new_minmax(Sv.device) if want_minmax else None
    _lib.call(
        "epb_noise_apply", ptr(Sv), ptr(echo_range), alpha_cp, ptr(noise), ptr(sn), ptr(sc), ptr(mm), C, P, R,
        int(ping_num), ctypes.c_float(float(snr)), stream(),
    )
    if mm is not None:
        mm._keep = pack
    return sn, sc, mm


def new_acc(C, nX, nR, device=None):
    acc = torch.empty((int(C), int(nX), int(nR), 4), dtype=torch.float64, device=device or require_cuda())
    _lib.call("epb_zero", ptr(acc), acc.numel() * 8, stream())
    return acc


def bin_reduce(Sv, range_var, xbin, r_edges, acc, C, P, R, nX, closed_right=False, with_height=False):
    nR = int(r_edges.numel()) - 1
    is64 = 1 if range_var.dtype == torch.float64 else 0
    _lib.call(
        "epb_bin_reduce", ptr(Sv), ptr(range_var), is64, ptr(xbin), ptr(r_edges), nR, int(closed_right),
        int(with_height), ptr(acc), C, P, R, nX, stream(),
    )
    return acc


def bin_reduce_law(Sv, rows, xbin, r_edges, acc, C, P, R, nX, closed_right=False, depth_off=None, depth_scale=None, fast=True):
    """fast=False forces the warp-per-row kernel (no workspace -> no dispatch to the persistent kernel)."""
    nR = int(r_edges.numel()) - 1
    nws = int(_lib.load().epb_pipeline_workspace_bytes_r(int(C), int(P), int(R), 0)) if fast else 0
    ws = torch.empty(nws, dtype=torch.uint8, device=Sv.device) if fast else None
    _lib.call(
        "epb_bin_reduce_law", ptr(Sv), ptr(rows), ptr(depth_off), ptr(depth_scale), ptr(xbin), ptr(r_edges), nR,
        int(closed_right), ptr(acc), C, P, R, nX, ptr(ws), nws, stream(),
    )
    return acc


def bin_finalize(acc, skipna=True, fill_value=float("nan"), to_db=True, want_height=False):
    C, nX, nR, _ = acc.shape
    out = empty((C, nX, nR), device=acc.device)
    h = torch.empty((C, nX, nR), dtype=torch.float64, device=acc.device) if want_height else None
    _lib.call(
        "epb_bin_finalize", ptr(acc), ptr(out), ptr(h), C * nX * nR, int(bool(skipna)), ctypes.c_float(float(fill_value)),
        int(bool(to_db)), stream(),
    )
    return out, h


def coarsen(Sv, echo_range, C, P, R, ping_num, range_sample_num):
    nP, nR = -(-P // ping_num), -(-R // range_sample_num)
    out = empty((C, nP, nR), device=Sv.device)
    er = empty((C, nP, nR), device=Sv.device) if echo_range is not None else None
    _lib.call("epb_coarsen", ptr(Sv), ptr(echo_range), ptr(out), ptr(er), C, P, R, int(ping_num), int(range_sample_num), stream())
    return out, er


def to_device_f64(a, device=None):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64))).to(device or require_cuda())


def add_depth(echo_range, off_p, scale, C, P, R):
    """depth = off[p] + echo_range * scale (scale: (P,) per ping or (C, P)); float32 device tensor [C,P,R]."""
    pk = ParamPack(C, P, echo_range.device)
    out = empty((C, P, R), device=echo_range.device)
    off_p = np.asarray(off_p, dtype=np.float64)
    scale = np.asarray(scale, dtype=np.float64)
    off_p = off_p.reshape(1, -1) if off_p.ndim == 1 else off_p  # per ping: explicit (1, P), unambiguous when C == P
    scale = scale.reshape(1, -1) if scale.ndim == 1 else scale
    _lib.call("epb_add_depth", ptr(echo_range), pk.cp(off_p), pk.cp(scale),
              ptr(out), C, P, R, stream())
    out._keep = pk
    return out


def freq_diff_mask(Sv, a, b, op, diff, C, P, R):
    m = torch.empty((int(P), int(R)), dtype=torch.uint8, device=Sv.device)
    _lib.call("epb_freq_diff_mask", ptr(Sv), int(a), int(b), int(op), ctypes.c_float(float(diff)), ptr(m), C, P, R, stream())
    return m


def apply_mask(src, mask, has_channel, fill, C, P, R):
    """fill: a number, or a float32 device tensor of shape (P, R) broadcast over channel."""
    out = empty((C, P, R), device=src.device)
    if isinstance(fill, torch.Tensor):
        _lib.call("epb_apply_mask_fill_array", ptr(src), ptr(mask), int(bool(has_channel)), ptr(fill), ptr(out), C, P, R, stream())
    else:
        _lib.call("epb_apply_mask", ptr(src), ptr(mask), int(bool(has_channel)), ctypes.c_float(float(fill)), ptr(out), C, P, R, stream())
    return out


def minmax(a):
    """(min, max, has_nan) over the non-NaN elements of a float32 device tensor (one streaming pass)."""
    mm = new_minmax(a.device)
    mm[2:] = 0
    _lib.call("epb_minmax", ptr(a), a.numel(), ptr(mm), stream())
    lo, hi, has_nan, _ = mm.tolist()
    if lo == float("inf"):
        lo = hi = float("nan")
    return lo, hi, bool(has_nan)


def range_max_into(x, rows, C, P, R, out):
    """epb_range_max into a 1-element float64 device tensor (asynchronous; -inf when every range is NaN)."""
    _lib.call("epb_range_max", ptr(x), ptr(rows), C, P, R, ptr(out), stream())
    return out


def range_max(x, rows, C, P, R):
    """Exact float64 nanmax of the echo_range implied by `rows` (x: raw samples for the NaN rule, or None)."""
    out = torch.empty(1, dtype=torch.float64, device=rows.device)
    range_max_into(x, rows, C, P, R, out)
    v = float(out.item())
    return float("nan") if v == float("-inf") else v


def pipeline_power_mvbs(x, rows, xbin, r_edges, acc, C, P, R, nX, ping_num, range_sample_num, noise_max=None,
                        snr=3.0, closed_right=False, noise_out=None, Sv=None, echo_range=None, Sv_noise=None,
                        Sv_corrected=None, fast=True, rmax_out=None):
    """fast=False forces the general kernel (no workspace -> no fast-path dispatch)."""
    nm = float("nan") if noise_max is None else float(noise_max)
    nR = int(r_edges.numel()) - 1
    nws = int(_lib.load().epb_pipeline_workspace_bytes_r(int(C), int(P), int(R), int(ping_num))) if fast else 0
    ws = torch.empty(nws, dtype=torch.uint8, device=x.device) if fast else None
    _lib.call(
        "epb_pipeline_power_mvbs", ptr(x), ptr(rows), ptr(xbin), ptr(r_edges), nR, int(closed_right), ptr(acc),
        ptr(noise_out), ptr(Sv), ptr(echo_range), ptr(Sv_noise), ptr(Sv_corrected), C, P, R, nX, int(ping_num),
        int(range_sample_num), ctypes.c_float(nm), ctypes.c_float(float(snr)), ptr(rmax_out), ptr(ws), nws, stream(),
    )
    return acc


def pipeline_power_mvbs_i16(counts, scratch, rows, xbin, r_edges, acc, C, P, R, nX, ping_num, range_sample_num,
                            noise_max=None, snr=3.0, closed_right=False, noise_out=None, rmax_out=None):
    """Fused pipeline on int16 raw power counts (-32768 = padding); scratch: float32 buffer of C*P*R elements."""
    nm = float("nan") if noise_max is None else float(noise_max)
    nR = int(r_edges.numel()) - 1
    nws = int(_lib.load().epb_pipeline_workspace_bytes(int(C), int(P), int(ping_num)))
    ws = torch.empty(nws, dtype=torch.uint8, device=counts.device)
    assert counts.dtype == torch.int16 and scratch.dtype == torch.float32 and scratch.numel() >= C * P * R
    _lib.call(
        "epb_pipeline_power_mvbs_i16", ptr(counts), ptr(scratch), ptr(rows), ptr(xbin), ptr(r_edges), nR, int(closed_right),
        ptr(acc), ptr(noise_out), C, P, R, nX, int(ping_num), int(range_sample_num), ctypes.c_float(nm),
        ctypes.c_float(float(snr)), ptr(rmax_out), ptr(ws), nws, stream(),
    )
    return acc


def ingest_power_i16(counts, out=None):
    """int16 raw power counts -> float32 backscatter_r (dB), -32768 -> NaN (convert/parse_base.py:24,302,686-730)."""
    assert counts.dtype == torch.int16 and counts.is_cuda and counts.is_contiguous()
    out = out if out is not None else torch.empty(counts.shape, dtype=torch.float32, device=counts.device)
    _lib.call("epb_ingest_power_i16", ptr(counts), ptr(out), int(counts.numel()), stream())
    return out


def range_diff_mean(rng, C, P, R):
    """Per-channel nanmean of the forward differences of a [C,P,R] float32 range variable (host float64 array [C])."""
    s = torch.empty(C, dtype=torch.float64, device=rng.device)
    n = torch.empty(C, dtype=torch.int64, device=rng.device)
    _lib.call("epb_range_diff_mean", ptr(rng), ptr(s), ptr(n), C, P, R, stream())
    s, n = s.cpu().numpy(), n.cpu().numpy()
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(n > 0, s / np.maximum(n, 1), np.nan)


def first_not_le(a, threshold):
    """First flat index of ``a`` where ``a <= threshold`` is False (np.argmin of the <= mask); None if there is none."""
    out = torch.empty(1, dtype=torch.int64, device=a.device)
    _lib.call("epb_first_not_le", ptr(a), int(a.numel()), ctypes.c_float(float(threshold)), ptr(out), stream())
    v = int(out.item())
    return None if v < 0 else v  # UINT64_MAX reads back as -1


def impulse_noise_mask(Sv, nsamp, C, P, R, num_side_pings, threshold, out=None):
    """out: optional (mask, block_means) buffers to reuse."""
    ns = torch.from_numpy(np.ascontiguousarray(nsamp, dtype=np.int32)).to(Sv.device)
    nbmax = int(max(-(-R // int(n)) for n in nsamp))
    mask, blocks = out if out is not None else (torch.empty((C, P, R), dtype=torch.uint8, device=Sv.device),
                                                torch.empty((C, P, nbmax), dtype=torch.float32, device=Sv.device))
    _lib.call("epb_impulse_noise_mask", ptr(Sv), ptr(ns), ptr(blocks), ptr(mask), C, P, R, nbmax, int(num_side_pings),
              ctypes.c_float(float(threshold)), stream())
    return mask, blocks


def impulse_noise_mask_depth(Sv, depth, edges, C, P, R, num_side_pings, threshold, want_upsampled=True):
    """Depth-value binning variant (use_index_binning=False); edges: host float64 array of interval edges.
    want_upsampled=False: the reference's upsampled_Sv is not materialised (single-pass kernel when the shape allows it)."""
    nb = len(edges) - 1
    e = torch.from_numpy(np.ascontiguousarray(edges, dtype=np.float64)).to(Sv.device)
    means = torch.empty((C, P, nb), dtype=torch.float32, device=Sv.device)
    first = torch.empty((C, P, nb), dtype=torch.int32, device=Sv.device)
    mask = torch.empty((C, P, R), dtype=torch.uint8, device=Sv.device)
    fused = (not want_upsampled and R % 16 == 0 and R <= 4096 and P < (1 << 30) and Sv.data_ptr() % 16 == 0 and depth.data_ptr() % 16 == 0
             and R * 8 + (nb + 1) * 4 + (2 * int(num_side_pings) + 1) * (2 * nb + 1 + R // 16) * 4 <= 96 * 1024)
    up = None if fused else torch.empty((C, P, R), dtype=torch.float32, device=Sv.device)  # the reference's upsampled_Sv
    scratch = torch.empty(nb + 1, dtype=torch.float32, device=Sv.device)
    _lib.call("epb_impulse_noise_mask_depth", ptr(Sv), ptr(depth), ptr(e), nb, ptr(means), ptr(first), ptr(up), ptr(mask),
              C, P, R, int(num_side_pings), ctypes.c_float(float(threshold)), ptr(scratch), stream())
    return mask, means, first, up


def transient_noise_mask_depth(Sv, depth, C, P, R, dmin, dmax, depth_bin, exclude_above, num_side_pings, threshold, want_pooled=False):
    """Depth-value window variant (use_index_binning=False).  Volumes whose depth rows are the same for every ping of a
    channel (checked on the device) take the single-pass strip kernel; others the per-sample bisection kernels, which need
    12 bytes of scratch per sample."""
    k = int(num_side_pings)
    if R % 16 == 0 and P < (1 << 30) and (2 * k + 1) * R < (1 << 24) and depth.data_ptr() % 16 == 0 and Sv.data_ptr() % 16 == 0:
        flag = torch.empty(1, dtype=torch.int32, device=Sv.device)
        ref = torch.empty((2, C, R), dtype=torch.float32, device=Sv.device)  # [0]: reference rows, [1]: scratch
        _lib.call("epb_depth_rows_uniform", ptr(depth), ptr(Sv), ptr(ref), ptr(flag), C, P, R, stream())
        nonuniform = int(flag.item())
        col0 = 0
        if nonuniform == 0:
            # columns shallower than exclude_above in every channel belong to no window (d - bin >= exclude_above): the strip
            # starts at the 16-aligned column at or before the first deeper one (64 KB of reference rows read back)
            with np.errstate(invalid="ignore"):
                deep = ref[0].cpu().numpy() >= float(exclude_above)
            first = min((int(np.argmax(row)) if row.any() else R - 16) for row in deep)
            col0 = (min(first, R - 16) // 16) * 16
        if nonuniform == 0 and R - col0 <= 4096:
            tables = torch.empty(C * 3 * R, dtype=torch.int16, device=Sv.device)
            mask = torch.empty((C, P, R), dtype=torch.uint8, device=Sv.device)
            pooled = torch.empty((C, P, R), dtype=torch.float32, device=Sv.device) if want_pooled else None
            _lib.call("epb_transient_noise_mask_depth_uniform", ptr(Sv), ptr(depth), ptr(ref), ptr(tables), ptr(mask), ptr(pooled),
                      C, P, R, ctypes.c_double(float(dmin)), ctypes.c_double(float(dmax)), ctypes.c_double(float(depth_bin)),
                      ctypes.c_double(float(exclude_above)), k, ctypes.c_float(float(threshold)), col0, stream())
            return mask, pooled
    pre = torch.empty((C, P, R + 1), dtype=torch.float64, device=Sv.device)
    cnt = torch.empty((C, P, R + 1), dtype=torch.int32, device=Sv.device)
    mask = torch.empty((C, P, R), dtype=torch.uint8, device=Sv.device)
    pooled = torch.empty((C, P, R), dtype=torch.float32, device=Sv.device) if want_pooled else None
    _lib.call("epb_transient_noise_mask_depth", ptr(Sv), ptr(depth), ptr(pre), ptr(cnt), ptr(mask), ptr(pooled), C, P, R,
              ctypes.c_double(float(dmin)), ctypes.c_double(float(dmax)), ctypes.c_double(float(depth_bin)),
              ctypes.c_double(float(exclude_above)), int(num_side_pings), ctypes.c_float(float(threshold)), stream())
    return mask, pooled


def transient_noise_mask(Sv, nsamp, C, P, R, min_range_sample, num_side_pings, threshold, want_pooled=False, out=None):
    """out: optional (mask, window_sums) buffers to reuse."""
    ns = torch.from_numpy(np.ascontiguousarray(nsamp, dtype=np.int32)).to(Sv.device)
    # the single-pass strip kernel (up to 4096 samples past exclude_above) needs no (sum, count) scratch (8 bytes per sample);
    # same conditions as epb_transient_noise_mask (masknoise.cu)
    m0, wmax = int(min_range_sample), int(max(nsamp))
    threads = ((R - (m0 & ~15)) // 16 + 31) // 32 * 32
    strip = (R % 16 == 0 and 0 < R - (m0 & ~15) <= 4096 and wmax < R - m0 and wmax <= 480 and P < (1 << 30)
             and (2 * int(num_side_pings) + 1) * (2 * wmax + 1) < (1 << 24)
             and (wmax + 16) // 16 + 1 + threads + (wmax + 15) // 16 + 1 <= 320)
    if out is not None:
        mask, sums = out
    else:
        mask = torch.empty((C, P, R), dtype=torch.uint8, device=Sv.device)
        sums = None if strip else torch.empty((C, P, R, 2), dtype=torch.float32, device=Sv.device)
    pooled = torch.empty((C, P, R), dtype=torch.float32, device=Sv.device) if want_pooled else None
    _lib.call("epb_transient_noise_mask", ptr(Sv), ptr(ns), ptr(sums), ptr(mask), ptr(pooled), C, P, R, int(min_range_sample),
              int(max(nsamp)), int(num_side_pings), ctypes.c_float(float(threshold)), stream())
    return mask, pooled


def transient_noise_mask_median(Sv, nsamp, C, P, R, min_range_sample, num_side_pings, threshold, want_pooled=False):
    """func="nanmedian", index windows."""
    ns = torch.from_numpy(np.ascontiguousarray(nsamp, dtype=np.int32)).to(Sv.device)
    mask = torch.empty((C, P, R), dtype=torch.uint8, device=Sv.device)
    pooled = torch.empty((C, P, R), dtype=torch.float32, device=Sv.device) if want_pooled else None
    _lib.call("epb_transient_noise_mask_median", ptr(Sv), ptr(ns), ptr(mask), ptr(pooled), C, P, R, int(min_range_sample),
              int(max(nsamp)), int(num_side_pings), ctypes.c_float(float(threshold)), stream())
    return mask, pooled


def transient_noise_mask_depth_median(Sv, depth, C, P, R, dmin, dmax, depth_bin, exclude_above, num_side_pings, threshold, want_pooled=False):
    """func="nanmedian", depth-value windows."""
    mask = torch.empty((C, P, R), dtype=torch.uint8, device=Sv.device)
    pooled = torch.empty((C, P, R), dtype=torch.float32, device=Sv.device) if want_pooled else None
    _lib.call("epb_transient_noise_mask_depth_median", ptr(Sv), ptr(depth), ptr(mask), ptr(pooled), C, P, R,
              ctypes.c_double(float(dmin)), ctypes.c_double(float(dmax)), ctypes.c_double(float(depth_bin)),
              ctypes.c_double(float(exclude_above)), int(num_side_pings), ctypes.c_float(float(threshold)), stream())
    return mask, pooled


def attenuated_signal_mask(Sv, rng, C, P, R, upper_limit_sl, lower_limit_sl, num_side_pings, threshold):
    """(mask [C,P,R] uint8, limits [C,P,2] int32 = the up / lw sample indices of every ping)."""
    mask = torch.empty((C, P, R), dtype=torch.uint8, device=Sv.device)
    limits = torch.empty((C, P, 2), dtype=torch.int32, device=Sv.device)
    _lib.call("epb_attenuated_signal_mask", ptr(Sv), ptr(rng), ptr(limits), ptr(mask), C, P, R,
              ctypes.c_double(float(upper_limit_sl)), ctypes.c_double(float(lower_limit_sl)), int(num_side_pings),
              ctypes.c_double(float(threshold)), stream())
    return mask, limits


def is_raw_counts(x):
    """True for int16 raw power counts (the ingest format: -32768 = padding), host array or tensor."""
    return getattr(x, "dtype", None) in (torch.int16, np.dtype("int16"))


def power_to_device_f32(x, keep_counts=False):
    """Power samples -> float32 CUDA tensor.  int16 raw counts cross PCIe as they are (2 bytes per sample) and are
    converted on the device (epb_ingest_power_i16); float data goes through :func:`device.to_device_f32`."""
    from .device import to_device_f32

    data = x.data if hasattr(x, "dims") else x  # DataArray -> its array (ndarray.data would be a raw buffer)
    if not is_raw_counts(data):
        return to_device_f32(data)
    t = data if isinstance(data, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(data))
    t = t.to(require_cuda(), non_blocking=True).contiguous()
    if keep_counts and t.shape[-1] % 4 == 0:
        return t  # K1 (epb_sv_power_i16) reads the counts itself
    return ingest_power_i16(t)


def synth_fill_i16(shape_cpr, seed, ping_offset=0, nan_tail=0.005, device=None):
    C, P, R = (int(s) for s in shape_cpr)
    out = torch.empty((C, P, R), dtype=torch.int16, device=device or require_cuda())
    _lib.call("epb_synth_fill_i16", ptr(out), C, P, R, ctypes.c_ulonglong(int(seed)), int(ping_offset),
              ctypes.c_uint(int(round(nan_tail * 65536))), stream())
    return out


def synth_fill(shape_cpr, kind, seed, inner=1, ping_offset=0, nan_tail=0.005, scale=1.0, device=None, out=None):
    C, P, R = (int(s) for s in shape_cpr)
    dev = device or require_cuda()
    full = (C, P, R) if inner == 1 else (C, P, R, inner)
    out = out if out is not None else empty(full, device=dev)
    q16 = int(round(nan_tail * 65536))
    _lib.call(
        "epb_synth_fill", ptr(out), C, P, R, int(inner), int(kind), ctypes.c_ulonglong(int(seed)), int(ping_offset),
        ctypes.c_uint(q16), ctypes.c_float(float(scale)), stream(),
    )
    return out
