"""Fused Sv -> background-noise removal -> MVBS pipeline (one pass over HBM, 4 bytes per sample).

The reference runs this chain as three calls that each materialise (channel, ping_time, range_sample) arrays:
``compute_Sv`` (calibrate/api.py:249) -> ``remove_background_noise`` (clean/api.py:436) -> ``compute_MVBS``
(commongrid/api.py:31).  :func:`compute_Sv_clean_MVBS` gives the result of that chain (same arguments, same
output Dataset as ``compute_MVBS`` applied to ``Sv_corrected``) from the raw power samples with a single
kernel (epb_pipeline_power_mvbs); the full-size intermediates are produced only on request (``keep=``).

:class:`FusedPlan` splits the call into the host-side assembly (argument validation, env / cal parameters,
ping-bin grid; O(channel x ping_time), done once per dataset) and :meth:`FusedPlan.run`, the device work
(row setup, exact range maximum, the fused kernel, the straddling-bin reduce, mean -> dB).

Host-resident volumes are STREAMED: contiguous (channel, ping-chunk) slabs are copied to the device on a copy
stream into a ring of slab buffers while the fused kernel consumes the previous slab on the compute stream
(the kernel accumulates into the same (sum, count) grid, so the result is identical to the one-shot launch).

Multi-GPU: the volume shards over ``ping_time`` (one process per GPU, each holding a contiguous, time-ordered
ping range whose length is a multiple of ``ping_num`` on all but the last rank).  The ping-bin grid is global
(derived from the first / last ping time over all ranks).  Each rank accumulates (sum, count) only for the bins
its own pings touch; ONE all-reduce(sum) over the first / last local bin of every rank merges the bins that
straddle shard boundaries (:func:`straddle_reduce`).  Two scalar all-reduces (max) fix the global grid.
"""

import ctypes
import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import kernels
from .calibrate.api import CALIBRATOR, check_input_args_combination
from .clean.utils import extract_dB
from .commongrid.api import _set_MVBS_attrs
from .commongrid.utils import _parse_x_bin, assign_bins, ping_time_bin_parsing_and_conversion, ping_time_edges, range_edges
from .dataset import DataArray, Dataset, EchoData
from .device import empty, require_cuda

from .utils.prov import echopype_prov_attrs

DIMS = ("channel", "ping_time", "range_sample")
_KEEP = ("Sv", "echo_range", "Sv_noise", "Sv_corrected")


def _dist():
    import torch.distributed as dist

    return dist


def _comm_device(group):
    dist = _dist()
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def global_ping_edges(ping_time, ping_time_bin, group=None):
    """Ping-bin edges of commongrid/api.py:118-124 for the GLOBAL ping axis: the pandas resample grid depends
    only on the first and the last ping time, which are all-reduced (min / max) across the ranks."""
    pt = np.asarray(ping_time).astype("datetime64[ns]").astype(np.int64)
    lo, hi = int(pt.min()), int(pt.max())
    if group is not None:
        dist = _dist()
        t = torch.tensor([-lo, hi], dtype=torch.int64, device=_comm_device(group))
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        lo, hi = -int(t[0].item()), int(t[1].item())
    return ping_time_edges(np.array([lo, hi], dtype="datetime64[ns]"), ping_time_bin)


def straddle_plan(lo, hi, group, shard=None):
    """Gather the (first, last) global ping bin of every rank once (host ints): the bins a rank shares with its
    neighbours depend only on the ping times, not on the data.

    ``shard`` = (n_pings, first_ping_ns, last_ping_ns, ping_num) of this rank: the same collective carries it so
    that the preconditions of ping-sharded execution are CHECKED instead of assumed - every rank but the last must
    hold a multiple of ``ping_num`` pings (otherwise the coarsen(ping_time=ping_num, boundary="pad") noise tiles of
    clean/api.py:403-407 would differ from the single-process tiling near every shard boundary) and the shards must
    be time-ordered across ranks.  Violations raise ValueError on every rank."""
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    idx = torch.full((world, 5), -1, dtype=torch.int64, device=_comm_device(group))
    idx[rank, 0], idx[rank, 1] = int(lo), int(hi)
    if shard is not None:
        idx[rank, 2], idx[rank, 3], idx[rank, 4] = int(shard[0]), int(shard[1]), int(shard[2])
    # values are >= -1 except timestamps before 1970, which ping-sharded execution does not support
    dist.all_reduce(idx, op=dist.ReduceOp.MAX, group=group)
    tab = idx.cpu().tolist()
    if shard is not None:
        ping_num = int(shard[3])
        held = [r for r in range(world) if tab[r][2] > 0]
        for a, b in zip(held[:-1], held[1:]):
            if ping_num > 1 and tab[a][2] % ping_num != 0:
                raise ValueError(
                    f"ping-sharded noise removal: rank {a} holds {tab[a][2]} pings, not a multiple of ping_num={ping_num}; "
                    "only the last shard may be short (its last tile is padded like coarsen(boundary='pad'))"
                )
            if tab[a][4] > tab[b][3]:
                raise ValueError(f"ping-sharded execution needs time-ordered shards: rank {a} ends after rank {b} starts")
    windows = [t[:2] for t in tab]
    plan = {"world": world, "rank": rank, "lo": int(lo), "hi": int(hi), "slots": straddle_slots(windows, lo, hi)}
    plan.update(straddle_neighbours(windows, rank))
    return plan


def straddle_neighbours(windows, rank):
    """Who shares this rank's first / last ping bin.  When every shared bin is held by exactly two (adjacent) ranks the
    exchange is PAIRWISE: ``left`` / ``right`` are the ranks to swap one bin slice with (None: that edge is not shared)
    and ``pairwise`` is True on every rank.  A bin held by three or more ranks (a shard that lies inside one ping bin)
    makes ``pairwise`` False everywhere and the all-reduce path is used."""
    holders = {}
    for r, (lo, hi) in enumerate(windows):
        if lo < 0 or hi < lo:
            continue
        for b in {int(lo), int(hi)}:
            holders.setdefault(b, []).append(r)
    pairwise = all(len(v) <= 2 for v in holders.values())
    lo, hi = windows[rank]
    left = right = None
    if pairwise and lo >= 0 and hi >= lo:
        for b, v in holders.items():
            if len(v) == 2 and rank in v:
                other = v[0] if v[1] == rank else v[1]
                if b == int(lo) and other < rank:
                    left = other
                if b == int(hi) and other > rank:
                    right = other
    return {"pairwise": pairwise, "left": left, "right": right}


def straddle_slots(idx, lo, hi):
    """For the first (slot 0) and the last (slot hi - lo) local ping bin: which slices of the exchange buffer hold the
    same global bin (entry 2 r: first bin of rank r, 2 r + 1: its last bin).  idx: [(lo, hi)] of every rank, -1 = none."""
    slots = []
    if hi < lo:
        return slots
    for slot, b in ((0, int(lo)), (int(hi) - int(lo), int(hi))):
        if slot != 0 and hi == lo:
            break
        src = []
        for r, (r_lo, r_hi) in enumerate(idx):
            if r_lo < 0:
                continue
            if r_lo == b:
                src.append(2 * r)
            elif r_hi == b:  # r_hi != r_lo here
                src.append(2 * r + 1)
        slots.append((slot, src))
    return slots


def straddle_reduce(acc, lo, hi, group, plan=None):
    """Merge the ping bins shared between ping-sharded ranks.

    acc : this rank's accumulators [C, hi - lo + 1, nR, 4] for the GLOBAL ping bins lo..hi (inclusive).
    Shards are contiguous and time-ordered, so only the first and the last local bin of a rank can also
    receive samples on another rank.  Every rank contributes those two bin slices (C x nR x 4 float64 each) to
    ONE all-reduce(sum) of a [world, 2, C, nR, 4] buffer (a few KB); afterwards every rank holds the complete
    sums of its own edge bins.  Interior bins never leave the rank.  No host synchronisation happens here when a
    ``plan`` (:func:`straddle_plan`) is supplied.  Returns acc (updated in place).
    """
    dist = _dist()
    world = dist.get_world_size(group)
    if world == 1:
        return acc
    plan = plan or straddle_plan(lo, hi, group)
    rank = plan["rank"]
    C, nXl, nR, _ = acc.shape
    cdev = _comm_device(group)
    buf = torch.zeros((world * 2, C, nR, 4), dtype=torch.float64, device=cdev)
    if hi >= lo:
        buf[2 * rank] = acc[:, 0].to(cdev)
        if hi != lo:
            buf[2 * rank + 1] = acc[:, hi - lo].to(cdev)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)  # the data-path collective
    for slot, src in plan["slots"]:
        acc[:, slot] = buf[src].sum(dim=0).to(acc.device)
    return acc


def straddle_exchange_cuda(acc, rmax, plan, group):
    """NCCL path of :func:`straddle_reduce` plus the global range maximum in THREE launches: epb_straddle_pack (own
    first / last ping bin and max(rmax) into this rank's slot of a [world, 2 S + 1] buffer, zeros elsewhere), ONE
    all-reduce(sum) of that buffer (the data-path collective), epb_straddle_unpack (sums of the shared slices into the
    own edge bins, maximum of the range maxima).  ``rmax``: device float64 tensor of local maxima or None.
    Returns the 1-element global range maximum (or None).  No host synchronisation."""
    from . import _lib
    from .device import ptr, stream

    dist = _dist()
    world, rank = plan["world"], plan["rank"]
    C, nXl, nR, _ = acc.shape
    W = 2 * C * nR * 4 + 1
    st = plan.setdefault("_cuda", {})
    if st.get("W") != W:
        src = np.zeros((2, world), dtype=np.int32)
        n0 = n1 = 0
        for slot, lst in plan["slots"]:
            if slot == 0:
                src[0, : len(lst)], n0 = lst, len(lst)
            else:
                src[1, : len(lst)], n1 = lst, len(lst)
        st.update(W=W, buf=torch.empty(world * W, dtype=torch.float64, device=acc.device),
                  src=torch.from_numpy(src).to(acc.device), n0=n0, n1=n1)
    buf = st["buf"]
    has_last = int(plan["hi"] > plan["lo"])
    _lib.call("epb_straddle_pack", ptr(acc), C, nXl, nR, ptr(rmax), 0 if rmax is None else int(rmax.numel()), ptr(buf),
              rank, world, has_last, stream())
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)  # the data-path collective
    out = torch.empty(1, dtype=torch.float64, device=acc.device) if rmax is not None else None
    _lib.call("epb_straddle_unpack", ptr(buf), ptr(st["src"]), st["n0"], st["n1"], C, nXl, nR, world, ptr(acc), ptr(out), stream())
    return out


def straddle_exchange_p2p(acc, plan, group):
    """Pairwise form of the straddling-bin exchange over NCCL point-to-point (ONE grouped send/recv launch): the own
    first / last ping-bin slices are packed (epb_straddle_pack), the first goes to ``left``, the last to ``right``, the
    neighbours' slices land in the inbox half of the same buffer and epb_straddle_unpack adds them to the own edge bins.
    Only neighbours synchronise - an N-rank all-reduce would expose the skew of the slowest rank to every rank."""
    from . import _lib
    from .device import ptr, stream

    dist = _dist()
    C, nXl, nR, _ = acc.shape
    S = C * nR * 4
    W = 2 * S + 1
    st = plan.setdefault("_p2p", {})
    left, right = plan["left"], plan["right"]
    single = plan["hi"] == plan["lo"]
    if st.get("W") != W:
        src = np.zeros((2, 2), dtype=np.int32)  # entries: 0 own first, 1 own last, 2 inbox from left, 3 inbox from right
        n0 = n1 = 0
        first = [0] + ([2] if left is not None else []) + ([3] if (single and right is not None) else [])
        if len(first) > 1:
            src[0, : len(first)], n0 = first[:2], min(len(first), 2)
        if not single and right is not None:
            src[1, :2], n1 = [1, 3], 2
        if len(first) > 2:  # a one-bin shard shared with both neighbours would need three sources: not pairwise by construction
            raise RuntimeError("internal: pairwise straddle plan with a three-way bin")
        st.update(W=W, buf=torch.empty(2 * W, dtype=torch.float64, device=acc.device), src=torch.from_numpy(src).to(acc.device), n0=n0, n1=n1)
    buf = st["buf"]
    _lib.call("epb_straddle_pack", ptr(acc), C, nXl, nR, None, 0, ptr(buf), 0, 2, int(not single), stream())
    ops = []
    if left is not None:
        ops.append(dist.P2POp(dist.isend, buf[0:S], left, group))
        ops.append(dist.P2POp(dist.irecv, buf[W : W + S], left, group))
    if right is not None:
        ops.append(dist.P2POp(dist.isend, buf[0:S] if single else buf[S : 2 * S], right, group))
        ops.append(dist.P2POp(dist.irecv, buf[W + S : W + 2 * S], right, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        _lib.call("epb_straddle_unpack", ptr(buf), ptr(st["src"]), st["n0"], st["n1"], C, nXl, nR, 2, ptr(acc), None, stream())
    return 2 if ops else 1


class FusedPlan:
    """Host-side assembly of one fused Sv -> noise -> MVBS job; :meth:`run` launches the device work.

    Arguments as :func:`compute_Sv_clean_MVBS`.  ``chunk_pings``: slab length (pings) for streaming a
    host-resident volume (rounded to a multiple of ``ping_num``); device-resident volumes run in one launch.
    """

    def __init__(self, echodata: EchoData, ping_num=None, range_sample_num=None, background_noise_max=None,
                 SNR_threshold="3.0dB", range_bin="20m", ping_time_bin="20s", skipna=True, fill_value=np.nan,
                 closed="left", range_var_max=None, keep: Sequence[str] = (), group=None, chunk_pings=8192,
                 fast=True, **cal_kwargs):
        waveform_mode = cal_kwargs.pop("waveform_mode", None)
        encode_mode = cal_kwargs.pop("encode_mode", None)
        waveform_mode = "BB" if waveform_mode == "FM" else waveform_mode
        if echodata.sonar_model in ("EK80", "ES80", "EA640"):
            if waveform_mode is None or encode_mode is None:
                raise ValueError("waveform_mode and encode_mode must be specified for EK80 calibration")
            check_input_args_combination(waveform_mode, encode_mode)
            if encode_mode != "power":
                raise ValueError("The fused pipeline handles power samples; use compute_Sv for complex samples")
        if echodata.sonar_model not in CALIBRATOR:
            raise ValueError(f"Unsupported sonar_model {echodata.sonar_model!r}")
        for k in keep:
            if k not in _KEEP:
                raise ValueError(f"keep entries must be among {_KEEP}")
        if not isinstance(range_bin, str):
            raise TypeError("range_bin must be a string")
        self.rb = _parse_x_bin(range_bin, "range_bin")
        if closed not in ["right", "left"]:
            raise ValueError(f"{closed} is not a valid option. Options are 'left' or 'right'.")
        if not isinstance(ping_time_bin, str):
            raise TypeError("ping_time_bin must be a string")
        self.do_noise = ping_num is not None
        self.snr = extract_dB(SNR_threshold) if self.do_noise else 0.0
        self.noise_max = extract_dB(background_noise_max) if (self.do_noise and background_noise_max is not None) else None
        if self.do_noise and (range_sample_num is None or int(ping_num) <= 0 or int(range_sample_num) <= 0):
            raise ValueError("ping_num and range_sample_num must be positive integers")
        self.ping_num = int(ping_num) if self.do_noise else 0
        self.range_sample_num = int(range_sample_num) if self.do_noise else 0
        self.range_bin, self.ping_time_bin, self.closed = range_bin, ping_time_bin, closed
        self.skipna, self.fill_value, self.keep, self.group = skipna, fill_value, tuple(keep), group
        self.range_var_max = range_var_max
        self.fast = bool(fast)  # False: force the general kernel (tests compare the two)
        self.p2p = True  # ping-sharded: pairwise neighbour exchange when the plan allows it (False: always all-reduce)
        # Optional: pairwise exchange + re-finalisation of the two edge ping bins on a side stream, so that the main stream
        # never waits for a neighbour.  Measured SLOWER on B200 (N = 2: 2.24 ms per step against 1.70 ms): the fused kernel
        # is persistent and holds every SM (one 216 KB CTA each), so the NCCL send/recv kernels of step s cannot start
        # before the fused kernel of step s + 1 retires its CTAs, and the small kernels of both streams interleave.  Off by
        # default; the result is identical either way (bench.py verification, N > 1).
        self.async_exchange = os.environ.get("EPB_ASYNC_EXCHANGE", "0") == "1" and group is not None
        if self.async_exchange:  # experiment knob: leave SMs to the NCCL kernels of the side stream
            from . import _lib as _l

            _l.call("epb_set_grid_reserve", int(os.environ.get("EPB_GRID_RESERVE", "4")))
        self._comm_stream = None
        self._pending = None  # event: the side-stream work of the last run() is complete

        self.dev = require_cuda()
        self.chunk_pings = max(1, int(chunk_pings))
        self._ring = None
        self._scratch = None  # float32 image of an int16 volume / slab, written only when the general kernel runs
        self._streams = None
        # Host-resident samples: the first slabs start moving NOW, while the plan (host parameter assembly, row records,
        # range grid with its one host synchronisation, accumulators) is still being set up - the copy engine is the
        # bottleneck of the streamed run, so every millisecond it idles at the start is lost.  (EK80: the beam group is only
        # known after the calibration object exists; its prefetch starts then.)
        self._pref = None
        if echodata.sonar_model not in ("EK80", "ES80", "EA640"):
            self._maybe_prefetch(echodata["Sonar/Beam_group1"])
        self.cal_obj = CALIBRATOR[echodata.sonar_model](
            echodata, env_params=cal_kwargs.pop("env_params", None), cal_params=cal_kwargs.pop("cal_params", None),
            ecs_file=cal_kwargs.pop("ecs_file", None), waveform_mode=waveform_mode, encode_mode=encode_mode,
            drop_last_hanning_zero=cal_kwargs.pop("drop_last_hanning_zero", False), slice_dict={},
        )
        if cal_kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(cal_kwargs)}")
        self.beam = echodata[getattr(self.cal_obj, "ed_beam_group", None) or "Sonar/Beam_group1"]
        self.row_builder = self.cal_obj._power_row_builder("Sv")  # host params are assembled here, once
        self.C, self.P, self.R = self.row_builder.shape

        # ping-bin grid (global when sharded) and this rank's window lo..hi of it
        pt = np.asarray(self.beam["ping_time"].values)
        self.p_edges = global_ping_edges(pt, ping_time_bin, group)
        xb = assign_bins(pt, self.p_edges, closed)
        inside = xb[xb >= 0]
        self.x_lo = int(inside.min()) if inside.size else 0
        self.x_hi = int(inside.max()) if inside.size else -1
        self.nX = self.x_hi - self.x_lo + 1
        self.sorted_pings = bool(np.all(np.diff(pt.astype("datetime64[ns]").astype(np.int64)) >= 0))
        if group is not None and not self.sorted_pings:
            raise ValueError("ping-sharded execution needs time-ordered ping_time on every rank")
        self.xbin = torch.from_numpy(np.where(xb >= 0, xb - self.x_lo, -1).astype(np.int32)).to(self.dev)
        self.launches = 0  # kernels of libepb200 launched by run() so far (bench.py "gpu_launches")
        self.record_events = False  # bench.py: CUDA events around every fused-kernel launch -> kernel_events
        self.kernel_events = []
        self.comm_events = []  # bench.py: events around the straddling-bin / range-maximum collectives of a step
        if group is not None:
            pt_ns = pt.astype("datetime64[ns]").astype(np.int64)
            shard = (self.P, int(pt_ns[0]) if self.P else -1, int(pt_ns[-1]) if self.P else -1, self.ping_num if self.do_noise else 1)
            self._splan = straddle_plan(self.x_lo, self.x_hi, group, shard)
        else:
            self._splan = None
        self._ub = None  # cached upper bound of the range grid (depends on the parameters only, not on the samples)
        if self._pref is None:
            self._maybe_prefetch(self.beam)

    # ---- device work ------------------------------------------------------------------------------------------
    def _group_max(self, v):
        if self.group is None:
            return v
        dist = _dist()
        t = torch.tensor([v if v == v else -np.inf], dtype=torch.float64, device=_comm_device(self.group))
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def _upper_edges(self, rows):
        """Range-bin edges against which the kernel bins.  With ``range_var_max`` they are final.  Otherwise the
        reference uses nanmax(echo_range) (commongrid/api.py:108-114), which depends on the NaN tails of the samples;
        the kernel bins against an upper-bound grid (range law at the last sample of every row, from the parameters
        alone, cached per plan) and the grid is cut back to the exact maximum when the result is read: bins above
        the exact maximum cannot have members.  This keeps the device sequence free of host synchronisation."""
        if self._ub is None:
            C, P, R = self.C, self.P, self.R
            if self.range_var_max is not None:
                ub = _parse_x_bin(self.range_var_max) + 1e-8
            else:
                ub = self._group_max(kernels.range_max(None, rows, C, P, R))
            e = range_edges(ub, self.rb)
            self._ub = (e, torch.from_numpy(np.ascontiguousarray(e, dtype=np.float64)).to(self.dev))
        return self._ub

    def run(self, x=None, finalize=True):
        """Launch the device work (no host synchronisation for a device-resident volume after the first call).
        x: raw samples (device tensor or host array, default: the EchoData's own).
        Returns (mvbs [C, nX, nR_upper] float32 device tensor or None, acc, rmax, outs, noise); ``rmax`` is a 1-element
        float64 device tensor holding the exact nanmax(echo_range) of this rank (or None with ``range_var_max``);
        :meth:`wrap` reads it and trims the range grid."""
        C, P, R = self.C, self.P, self.R
        x = self.beam["backscatter_r"].data if x is None else x
        rows = self.row_builder.build()
        self.launches += 1
        self.rows = rows
        raw = kernels.is_raw_counts(x)  # int16 raw power counts (ingest format, -32768 = padding)
        if raw and (self.keep or not self.fast):  # full-size outputs / forced general kernel: float image first
            x, raw = kernels.power_to_device_f32(x), False
            self.launches += 1
        on_device = isinstance(x, torch.Tensor) and x.is_cuda
        outs = {k: (empty((C, P, R), device=self.dev) if k in self.keep else None) for k in _KEEP}
        noise = empty((C, -(-P // self.ping_num)), device=self.dev) if self.do_noise else None
        e_ub, edges_t = self._upper_edges(rows)
        nX = max(self.nX, 1)
        acc = kernels.new_acc(C, nX, len(e_ub) - 1, self.dev)
        self.launches += 1
        rmax = None
        if on_device:
            if self.range_var_max is None:
                rmax = torch.empty(1, dtype=torch.float64, device=self.dev)  # filled by the fused launch
                self.launches += 2
            if self.record_events:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            if raw:
                if self._scratch is None or self._scratch.numel() < C * P * R:
                    self._scratch = torch.empty(C * P * R, dtype=torch.float32, device=self.dev)
                kernels.pipeline_power_mvbs_i16(
                    x, self._scratch, rows, self.xbin, edges_t, acc, C, P, R, nX, self.ping_num, self.range_sample_num,
                    noise_max=self.noise_max, snr=self.snr, closed_right=(self.closed == "right"), noise_out=noise, rmax_out=rmax,
                )
                self.launches += 1  # the (gated) ingest kernel
            else:
                kernels.pipeline_power_mvbs(
                    x, rows, self.xbin, edges_t, acc, C, P, R, nX, self.ping_num, self.range_sample_num,
                    noise_max=self.noise_max, snr=self.snr, closed_right=(self.closed == "right"), noise_out=noise,
                    fast=self.fast, Sv=outs["Sv"], echo_range=outs["echo_range"], Sv_noise=outs["Sv_noise"], Sv_corrected=outs["Sv_corrected"],
                    rmax_out=rmax,
                )
            if self.record_events:
                ev[1].record()
                self.kernel_events.append(ev)
            self.launches += 3 if self.fast else 1
        else:
            rmax = self._run_streamed(x, rows, outs, noise, acc, edges_t)
        if (self.group is not None and finalize and self.async_exchange and self.p2p and acc.is_cuda and self._splan["pairwise"]
                and _dist().get_backend(self.group) == "nccl"):
            mvbs, rmax = self._exchange_async(acc, rmax)
            return mvbs, acc, rmax, outs, noise
        if self.group is not None:
            if self.record_events:
                cev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                cev[0].record()
            if acc.is_cuda and _dist().get_backend(self.group) == "nccl" and self._splan["pairwise"] and self.p2p:
                self.launches += straddle_exchange_p2p(acc, self._splan, self.group)
                if rmax is not None:  # global nanmax(echo_range): needed when the grid is read (wrap), not by the bins;
                    # the 1-double max all-reduce runs asynchronously on NCCL's stream, off the critical path
                    if rmax.numel() > 1:
                        rmax = rmax.max().reshape(1)
                    rmax._work = _dist().all_reduce(rmax, op=_dist().ReduceOp.MAX, group=self.group, async_op=True)
            elif acc.is_cuda and _dist().get_backend(self.group) == "nccl":
                rmax = straddle_exchange_cuda(acc, rmax, self._splan, self.group)
                self.launches += 2
            else:
                straddle_reduce(acc, self.x_lo, self.x_hi, self.group, self._splan)
                if rmax is not None:  # global nanmax(echo_range): a 1-double max all-reduce, read only in wrap()
                    rmax = rmax.max().reshape(1).to(_comm_device(self.group))
                    _dist().all_reduce(rmax, op=_dist().ReduceOp.MAX, group=self.group)
            if self.record_events:
                cev[1].record()
                self.comm_events.append(cev)
        mvbs = None
        if finalize:
            mvbs, _ = kernels.bin_finalize(acc, skipna=self.skipna, fill_value=self.fill_value, to_db=True)
            self.launches += 1
        return mvbs, acc, rmax, outs, noise

    def _exchange_async(self, acc, rmax):
        """Finalise every bin on the main stream (the edge ping bins provisionally), then, on the side stream: swap the
        edge slices with the neighbours, add them, finalise the two edge ping bins again.  The main stream does not
        wait; :meth:`sync_pending` (called by :meth:`wrap`) makes the side-stream results visible."""
        from . import _lib
        from .device import ptr, stream

        main = torch.cuda.current_stream()
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.dev)
        comm = self._comm_stream
        mvbs, _ = kernels.bin_finalize(acc, skipna=self.skipna, fill_value=self.fill_value, to_db=True)
        self.launches += 1
        ready = torch.cuda.Event()
        ready.record(main)
        for t in (acc, mvbs, rmax):
            if t is not None:
                t.record_stream(comm)
        C, nXl, nR, _ = acc.shape
        with torch.cuda.stream(comm):
            comm.wait_event(ready)
            if self.record_events:
                cev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                cev[0].record(comm)
            self.launches += straddle_exchange_p2p(acc, self._splan, self.group)
            slots = ([0] if self._splan["left"] is not None or (nXl == 1 and self._splan["right"] is not None) else [])
            if nXl > 1 and self._splan["right"] is not None:
                slots.append(nXl - 1)
            for slot in slots:  # the merged edge ping bins: mean -> dB again (one tiny launch per channel)
                for c in range(C):
                    _lib.call("epb_bin_finalize", ptr(acc[c, slot]), ptr(mvbs[c, slot]), None, nR, int(bool(self.skipna)),
                              ctypes.c_float(float(self.fill_value)), 1, stream())
                    self.launches += 1
            if rmax is not None:
                if rmax.numel() > 1:
                    rmax = rmax.max().reshape(1)
                    rmax.record_stream(comm)
                rmax._work = _dist().all_reduce(rmax, op=_dist().ReduceOp.MAX, group=self.group, async_op=True)
            if self.record_events:
                cev[1].record(comm)
                self.comm_events.append(cev)
            done = torch.cuda.Event()
            done.record(comm)
        self._pending = done
        return mvbs, rmax

    def sync_pending(self):
        """make the side-stream part of the last :meth:`run` (edge-bin exchange) visible to the current stream"""
        if self._pending is not None:
            torch.cuda.current_stream().wait_event(self._pending)
            self._pending = None

    def _host_slabs(self, x):
        """Host tensor of the samples (float32, or int16 raw counts), slab length, and the 3-slab device ring + streams."""
        R, P = self.R, self.P
        raw = kernels.is_raw_counts(x)
        if raw:
            xh = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
        else:
            xh = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))
            if xh.dtype != torch.float32:
                xh = xh.float()
        xh = xh.contiguous()
        pn = self.ping_num if self.do_noise else 1
        chunk = min(P, max(pn, (self.chunk_pings // pn) * pn))
        if self._ring is None or self._ring[0].numel() < chunk * R or self._ring[0].dtype != xh.dtype:
            self._ring = [torch.empty(chunk * R, dtype=xh.dtype, device=self.dev) for _ in range(3)]
            self._scratch = torch.empty(chunk * R, dtype=torch.float32, device=self.dev) if raw else None
            self._streams = (torch.cuda.Stream(device=self.dev), [torch.cuda.Event() for _ in range(3)],
                             [torch.cuda.Event() for _ in range(3)])
        return xh, raw, chunk

    def _maybe_prefetch(self, beam):
        x0 = beam["backscatter_r"].data if "backscatter_r" in beam else None
        if x0 is None or (isinstance(x0, torch.Tensor) and x0.is_cuda) or self.keep or not self.fast or len(x0.shape) != 3:
            return
        self.C, self.P, self.R = (int(v) for v in x0.shape)
        self._prefetch(x0)

    def _prefetch(self, x):
        """Start the H2D copies of the first ring slabs (see __init__); :meth:`_run_streamed` picks them up."""
        xh, _, chunk = self._host_slabs(x)
        copy_s, filled, _ = self._streams
        copy_s.wait_stream(torch.cuda.current_stream())  # a reused ring: earlier readers are done
        n = 0
        with torch.cuda.stream(copy_s):
            for c in range(self.C):
                for p0 in range(0, self.P, chunk):
                    if n == 3:
                        break
                    pc = min(chunk, self.P - p0)
                    self._ring[n][: pc * self.R].copy_(xh[c, p0 : p0 + pc].reshape(-1), non_blocking=True)
                    filled[n].record(copy_s)
                    n += 1
                if n == 3:
                    break
        self._pref = (x, n, chunk, xh)  # (xh: keeps a converted host copy alive until the copies ran)

    def _run_streamed(self, x, rows, outs, noise, acc, edges_t):
        """Host-resident volume: slabs of (1 channel, chunk pings) move H2D on a copy stream into a 3-slab ring
        while the fused kernel runs on the previous slab and accumulates into the same grid.  Returns the device
        tensor of per-slab exact range maxima (or None with ``range_var_max``)."""
        C, P, R = self.C, self.P, self.R
        pref, self._pref = self._pref, None
        if pref is not None and x is pref[0]:  # the samples whose first slabs __init__ put on their way
            xh, raw, chunk, npre = pref[3], kernels.is_raw_counts(pref[3]), pref[2], pref[1]
        else:
            xh, raw, chunk = self._host_slabs(x)
            npre = 0
        pn = self.ping_num if self.do_noise else 1
        copy_s, filled, freed = self._streams
        main = torch.cuda.current_stream()
        nX = max(self.nX, 1)
        want_rmax = self.range_var_max is None
        rmax_t = torch.full((C * (-(-P // chunk)),), -np.inf, dtype=torch.float64, device=self.dev) if want_rmax else None
        copy_s.wait_stream(main)
        i = 0
        rows_b = rows.view(C * P, -1)
        for c in range(C):
            for p0 in range(0, P, chunk):
                pc = min(chunk, P - p0)
                slot = i % 3
                buf = self._ring[slot][: pc * R]
                if i >= npre:  # (the first slabs may already be on their way: _prefetch)
                    with torch.cuda.stream(copy_s):
                        if i >= 3:
                            copy_s.wait_event(freed[slot])
                        buf.copy_(xh[c, p0 : p0 + pc].reshape(-1), non_blocking=True)
                        filled[slot].record(copy_s)
                main.wait_event(filled[slot])
                rsub = rows_b[c * P + p0 : c * P + p0 + pc]
                sub = lambda t: None if t is None else t[c, p0 : p0 + pc]  # noqa: E731
                nz = None if noise is None else noise[c, p0 // pn : p0 // pn + -(-pc // pn)]
                if raw:
                    kernels.pipeline_power_mvbs_i16(
                        buf, self._scratch, rsub, self.xbin[p0 : p0 + pc], edges_t, acc[c : c + 1], 1, pc, R, nX, self.ping_num,
                        self.range_sample_num, noise_max=self.noise_max, snr=self.snr, closed_right=(self.closed == "right"),
                        noise_out=nz, rmax_out=rmax_t[i : i + 1] if want_rmax else None,
                    )
                    self.launches += 1
                else:
                    kernels.pipeline_power_mvbs(
                        buf, rsub, self.xbin[p0 : p0 + pc], edges_t, acc[c : c + 1], 1, pc, R, nX, self.ping_num,
                        self.range_sample_num, noise_max=self.noise_max, snr=self.snr, closed_right=(self.closed == "right"),
                        fast=self.fast, noise_out=nz,
                        Sv=sub(outs["Sv"]), echo_range=sub(outs["echo_range"]), Sv_noise=sub(outs["Sv_noise"]),
                        Sv_corrected=sub(outs["Sv_corrected"]), rmax_out=rmax_t[i : i + 1] if want_rmax else None,
                    )
                self.launches += (3 if self.fast else 1) + (2 if want_rmax else 0)
                freed[slot].record(main)
                i += 1
        return rmax_t

    @staticmethod
    def wait_rmax(rmax):
        """make the asynchronous global range maximum of a ping-sharded run visible to the current stream"""
        work = getattr(rmax, "_work", None)
        if work is not None:
            work.wait()
            rmax._work = None

    def wrap(self, mvbs, acc, rmax, outs, noise):
        """Dataset with the variables / coords / attrs of compute_MVBS (commongrid/api.py:130-189).  Reads the exact
        range maximum (one host synchronisation) and cuts the upper-bound range grid back to it."""
        beam = self.beam
        e_ub = self._ub[0]
        self.sync_pending()
        if rmax is not None:
            self.wait_rmax(rmax)
            v = float(rmax.max().item())
            r_edges = range_edges(float("nan") if v == float("-inf") else v, self.rb)
        else:
            r_edges = e_ub
        nR = len(r_edges) - 1
        if nR < len(e_ub) - 1:
            acc = acc[:, :, :nR]
            mvbs = None if mvbs is None else mvbs[:, :, :nR]
        pe = self.p_edges[self.x_lo : self.x_hi + 1] if self.nX > 0 else self.p_edges[:0]
        ds = Dataset(coords={"ping_time": pe, "channel": beam["channel"].values, "echo_range": r_edges[:-1]})
        if mvbs is not None:
            v = mvbs.cpu().numpy().astype(np.float64)
            ds["Sv"] = (("channel", "ping_time", "echo_range"), v[:, : self.nX] if self.nX > 0 else v[:, :0])
            _set_MVBS_attrs(ds)
            ds["echo_range"].attrs.update({"long_name": "Range distance", "units": "m"})
            resvalue, reslabel = ping_time_bin_parsing_and_conversion(self.ping_time_bin)
            ds["Sv"].attrs.update(
                {
                    "cell_methods": (
                        f"ping_time: mean (interval: {resvalue} {reslabel} "
                        "comment: ping_time is the interval start) "
                        f"echo_range: mean (interval: {self.rb} meter "
                        "comment: echo_range is the interval start)"
                    ),
                    "binning_mode": "physical units",
                    "range_meter_interval": str(self.rb) + "m",
                    "ping_time_interval": self.ping_time_bin,
                }
            )
        else:
            ds.attrs["acc"] = acc.contiguous()
        ds["frequency_nominal"] = beam["frequency_nominal"]
        prov = echopype_prov_attrs(process_type="processing")
        prov["processing_function"] = "pipeline.compute_Sv_clean_MVBS"
        ds.attrs.update(prov)
        if self.do_noise:
            ds.attrs["noise_estimate"] = DataArray(noise, ("channel", "ping_tile"), name="noise")
        if self.keep:
            kept = Dataset(coords={d: beam[d].values for d in DIMS})
            for k in self.keep:
                da = DataArray(outs[k], DIMS, name=k, law={"kind": "derived"} if k in ("Sv", "Sv_corrected") else None)
                if k == "echo_range":
                    da.law = {"rows": self.rows, "kind": "echo_range", "minmax": None}
                kept[k] = da
            ds.attrs["kept"] = kept
        ds.attrs["_rows"] = self.rows  # keeps the row table alive for callers that re-run the kernel
        return ds


def compute_Sv_clean_MVBS(
    echodata: EchoData,
    ping_num: Optional[int] = None,
    range_sample_num: Optional[int] = None,
    background_noise_max: Optional[str] = None,
    SNR_threshold: str = "3.0dB",
    range_bin: str = "20m",
    ping_time_bin: str = "20s",
    skipna: bool = True,
    fill_value: float = np.nan,
    closed: str = "left",
    range_var_max: Optional[str] = None,
    keep: Sequence[str] = (),
    group=None,
    finalize: bool = True,
    chunk_pings: int = 8192,
    fast: bool = True,
    **cal_kwargs,
):
    """
    Calibrate, remove background noise and bin-average in one device pass.

    Equivalent to ``compute_MVBS(remove_background_noise(compute_Sv(echodata, **cal_kwargs), ping_num,
    range_sample_num, background_noise_max, SNR_threshold), "echo_range", range_bin, ping_time_bin, ...)`` with
    ``Sv_corrected`` taking the place of ``Sv`` (``ping_num=None``: no noise removal, MVBS of ``Sv``).
    Power-sample data only (EK60 / ES70, AZFP, EK80 ``encode_mode="power"``).

    keep : names among {"Sv", "echo_range", "Sv_noise", "Sv_corrected"} to materialise as device arrays; they
        are returned in ``ds.attrs["kept"]`` as a Dataset.
    group : torch.distributed process group for ping-sharded execution (each rank passes its own shard and gets
        the MVBS of the ping bins its shard touches; bins straddling two shards are complete on both ranks).
    finalize : when False the raw accumulators (device tensor [C, nX, nR, 4]) are returned in ``attrs["acc"]``.
    chunk_pings : slab length for streaming a host-resident volume to the device.
    """
    plan = FusedPlan(
        echodata, ping_num=ping_num, range_sample_num=range_sample_num, background_noise_max=background_noise_max,
        SNR_threshold=SNR_threshold, range_bin=range_bin, ping_time_bin=ping_time_bin, skipna=skipna, fill_value=fill_value,
        closed=closed, range_var_max=range_var_max, keep=keep, group=group, chunk_pings=chunk_pings, fast=fast, **cal_kwargs,
    )
    mvbs, acc, rmax, outs, noise = plan.run(finalize=finalize)
    return plan.wrap(mvbs, acc, rmax, outs, noise)
