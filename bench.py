#!/usr/bin/env python
"""bench.py - benchmark of echopype_b200 (contract: task statement (4)).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--config cfg2|cfg3|cfg4|cfg5] [--scaling weak|strong] [--no-extra] [--no-verify] [--no-cpu]

Workloads (BASELINE.json configs; 1 sample = one (channel, ping, range_sample) element; samples/s, whole job):
  cfg2 (default, configs[1])  EK60 power 4 ch x 100 000 ping x 4096: compute_Sv -> remove_background_noise(5, 30, 3 dB)
                              -> compute_MVBS("20m", "20s"), the fused kernel epb_pipeline_power_mvbs (4 B / sample)
  cfg3 (configs[2])           EK80 broadband 6 ch x 50 000 ping x 8192 complex samples x 4 beams: compute_Sv with pulse
                              compression (K3, epb_pulse_compress_sv; 36 B / sample, 79 GB of input)
  cfg4 (configs[3])           AZFP 4 ch x 200 000 ping x 2048: compute_Sv + compute_MVBS (fused kernel, no noise removal)
  cfg5 (configs[4])           EK80 CW power 6 ch x 1 000 000 ping x 4096: Sv -> noise -> MVBS (fused kernel, 98 GB)
One "step" = one pass of the chain over the whole volume (inputs far exceed the 126 MB L2: no flush needed).

* own arm `value`: inputs resident in HBM; CUDA events on the launching stream, barrier + synchronize on both sides,
  max over ranks.  `roofline`: the dominant kernel alone (events around each launch) at its algorithmic bytes per
  sample against MEASURED_PEAKS.json hbm_gbs.
* own arm `e2e`: the public call on an EchoData whose samples live in PINNED HOST memory (host parameter assembly +
  streamed H2D + kernels + D2H of the result inside the timed region; wall clock, max over ranks).
* `extra` (default config only): `kernels` = K1 (compute_Sv alone, with / without echo_range) on the same volume;
  `configs` = the resident value + roofline of cfg3 / cfg4 / cfg5 measured in the same run (N = 1: full size on one
  GPU; N > 1: STRONG scaling, the config's volume split over the N ranks by ping_time); `sustained` = the
  headline step repeated back to back for >= 2 s with clocks sampled.
* N > 1 (torchrun, one rank per GPU): by default WEAK scaling (every rank holds its own 100 000-ping shard of one
  global, time-ordered volume).  The ping axis starts 10 s after a 20 s bin edge, so EVERY shard boundary falls inside
  a ping bin: the straddling-bin exchange merges real partial sums.  `config.verified`: on a reduced volume the
  N-rank grid is compared on rank 0 with the single-GPU grid of the concatenated volume (member counts bit-equal,
  linear sums to a few float32 ulp, MVBS within 2 float32 ulp of the dB value).
* `cpu_baseline` (rank 0, N = 1) and `--impl reference`: the numpy float64 oracle (a port of the reference's operation
  sequence, pinned to outputs of the reference's own code by tests/test_reference_pinned.py; echopype itself cannot
  be imported in this image) on one core / ping-sharded over all host cores, bounded samples.
"""

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "samples/sec Sv->MVBS pipeline on (chan,ping,range) volume"
UNIT = "samples/s"
PING_NUM, RS_NUM, SNR, RANGE_BIN, PING_BIN = 5, 30, "3.0dB", "20m", "20s"
# the ping axis starts 10 pings (= 10 s) after a 20 s bin edge: every shard boundary (multiples of 100 000 pings) falls INSIDE
# a ping bin, so the straddling-bin exchange merges real partial sums; 10 is a multiple of ping_num, so noise tiles do not
# straddle ping bins (the unaligned case, origin 7 s, is timed as extra.unaligned_origin_7s)
PING_ORIGIN = 10

CONFIGS = {
    "cfg2": dict(kind="pipeline", sonar="EK60", C=4, P=100_000, R=4096, noise=True, seed=2000, bytes_per_sample=4,
                 chain=f"EK60 Sv->remove_noise(ping_num={PING_NUM},range_sample_num={RS_NUM},SNR={SNR})->compute_MVBS({RANGE_BIN},{PING_BIN})"),
    "cfg3": dict(kind="bb", sonar="EK80", C=6, P=50_000, R=8192, B=4, seed=3000, bytes_per_sample=36,
                 chain="EK80 BB pulse-compressed compute_Sv (complex samples, 4 beams)"),
    "cfg4": dict(kind="pipeline", sonar="AZFP", C=4, P=200_000, R=2048, noise=False, seed=4000, bytes_per_sample=4,
                 chain=f"AZFP compute_Sv->compute_MVBS({RANGE_BIN},{PING_BIN})"),
    "cfg5": dict(kind="pipeline", sonar="EK80", C=6, P=1_000_000, R=4096, noise=True, seed=5000, bytes_per_sample=4,
                 chain=f"EK80 CW power Sv->remove_noise(ping_num={PING_NUM},range_sample_num={RS_NUM},SNR={SNR})->compute_MVBS({RANGE_BIN},{PING_BIN})"),
}


def workload_name(name, P_total, world, scaling):
    c = CONFIGS[name]
    shape = f"{c['C']}ch x {P_total} ping x {c['R']} range" + (f" x {c['B']} beams" if "B" in c else "")
    if world == 1:
        return f"{name}: {c['chain']}, {shape}"
    if scaling == "weak":
        return f"{name}: {c['chain']}, {shape} per GPU"
    return f"{name}: {c['chain']}, {shape} split over {world} GPUs by ping_time"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------------------
# clocks: sample NVML during the timed region
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
        0x100: "display_clock_setting",
    }

    def __init__(self, index):
        self.samples, self.power, self.reasons, self.max_mhz, self._stop, self._t = [], [], set(), None, threading.Event(), None
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                index = int(vis.split(",")[index])
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, repr(e)

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "power_w_max": (round(max(self.power), 1) if self.power else None)}


# ---------------------------------------------------------------------------------------------------------------
# synthetic EchoData of a config (device or host)
# ---------------------------------------------------------------------------------------------------------------
def make_echodata(name, P, ping_offset, device, backscatter=None, seed=None, nan_tail=0.005):
    from echopype_b200 import synth

    c = CONFIGS[name]
    seed = c["seed"] if seed is None else seed
    if name == "cfg2":
        return synth.make_ek60(c["C"], P, c["R"], seed=seed, device=device, nan_tail=nan_tail, ping_offset=ping_offset, backscatter=backscatter)
    if name == "cfg4":
        return synth.make_azfp(c["C"], P, c["R"], seed=seed, device=device, ping_offset=ping_offset, backscatter=backscatter)
    if name == "cfg5":
        return synth.make_ek80(C=c["C"], P=P, R=c["R"], mode="CW", encode="power", device=device, gpt_channel=1, nan_tail=nan_tail,
                               seed=seed, ping_offset=ping_offset, backscatter=backscatter)
    return synth.make_ek80(C=c["C"], P=P, R=c["R"], B=c["B"], mode="BB", encode="complex", device=device, nan_tail=nan_tail, seed=seed,
                           ping_offset=ping_offset, backscatter=backscatter)


def chain_kwargs(name):
    c = CONFIGS[name]
    kw = dict(range_bin=RANGE_BIN, ping_time_bin=PING_BIN)
    if c.get("noise"):
        kw.update(ping_num=PING_NUM, range_sample_num=RS_NUM, SNR_threshold=SNR)
    if name == "cfg4":
        kw["env_params"] = {"salinity": 30.0, "pressure": 50.0}
    if name == "cfg5":
        kw.update(waveform_mode="CW", encode_mode="power")
    return kw


# ---------------------------------------------------------------------------------------------------------------
# CPU oracle chain (cpu_baseline and --impl reference); the only place bench.py executes oracle/
# ---------------------------------------------------------------------------------------------------------------
_W = {}


def _oracle_chain(name, ed):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_glue as og
    from oracle import clean as oclean
    from oracle import commongrid as ogrid

    if name == "cfg3":
        return og.ek80(ed, "Sv", "BB", "complex")["out"]
    ref = {"cfg2": lambda: og.ek60(ed, "Sv"), "cfg4": lambda: og.azfp(ed, "Sv", 30.0, 50.0), "cfg5": lambda: og.ek80(ed, "Sv", "CW", "power")}[name]()
    Sv = ref["out"]
    if CONFIGS[name].get("noise"):
        Sv = oclean.remove_background_noise(Sv, ref["echo_range"], ref["sound_absorption"], PING_NUM, RS_NUM, None, SNR)["Sv_corrected"]
    pt = np.asarray(ed["Sonar/Beam_group1"]["ping_time"].values).astype("datetime64[ns]").astype(np.int64)
    return ogrid.compute_MVBS(Sv, ref["echo_range"], pt, range_bin=RANGE_BIN, ping_time_bin=PING_BIN)["Sv"]


def _worker_init(name, P):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    _W["name"], _W["P"] = name, P


def _worker_prepare(i):
    _W["ed"] = make_echodata(_W["name"], _W["P"], PING_ORIGIN + i * _W["P"], device=False, seed=CONFIGS[_W["name"]]["seed"] + i)
    return i


def _worker_step(_):
    if "ed" not in _W:  # a worker that was not handed a prepare task
        _worker_prepare(os.getpid() % 1000)
    return float(_oracle_chain(_W["name"], _W["ed"]).shape[1])


def cpu_baseline_single(name, P_sample):
    _oracle_chain(name, make_echodata(name, min(100, P_sample), PING_ORIGIN, device=False))  # warm imports / allocator
    ed = make_echodata(name, P_sample, PING_ORIGIN, device=False)
    t0 = time.perf_counter()
    _oracle_chain(name, ed)
    dt = time.perf_counter() - t0
    c = CONFIGS[name]
    return c["C"] * P_sample * c["R"] / dt, dt


def run_reference(args):
    """--impl reference: the oracle chain on all host cores, ping-sharded; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import multiprocessing as mp

    name = args.config
    c = CONFIGS[name]
    cores = os.cpu_count() or 1
    P_w = args.ref_pings_per_core if name != "cfg3" else max(2, args.ref_pings_per_core // 50)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_worker_init, initargs=(name, P_w)) as pool:
        list(pool.imap(_worker_prepare, range(cores), chunksize=1))  # one resident shard per worker process
        for _ in range(args.warmup):
            pool.map(_worker_step, range(cores), chunksize=1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_worker_step, range(cores), chunksize=1)
        dt = time.perf_counter() - t0
    n_step = c["C"] * P_w * c["R"] * cores
    value = n_step * args.steps / dt
    P_total = c["P"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(name, P_total, world, args.scaling)},
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"each step = {cores} processes x ({c['C']}ch x {P_w} ping x {c['R']} range) = {n_step} samples of the {name} volume, "
                      "normalised to samples/s",
            "what": "numpy/scipy/pandas float64 port of the reference chain (oracle/, pinned to outputs of echopype's own code by "
                    "tests/test_reference_pinned.py; echopype itself is not importable in this image), ping-sharded over all host "
                    "cores with multiprocessing - the stand-in for the reference's dask-chunked path",
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(local_rank):
    """Pin this process (and therefore its pinned host allocations: first touch) to the CPUs of the NUMA node the GPU
    hangs off.  With 4-8 ranks streaming from pinned memory, host buffers on the wrong socket push every H2D copy over
    the inter-socket link.  Returns a short description for the bench line."""
    try:
        import pynvml

        pynvml.nvmlInit()
        idx = local_rank
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis and all(v.strip().isdigit() for v in vis.split(",")):
            idx = int(vis.split(",")[local_rank])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # nvml pads the domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return {"numa_node": None, "note": "no NUMA information for the GPU"}
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed)}
    except Exception as e:  # noqa: BLE001
        return {"numa_node": None, "note": repr(e)[:120]}


class Ctx:
    def __init__(self):
        self.numa = bind_to_gpu_numa(int(os.environ.get("LOCAL_RANK", "0")))
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
        torch.cuda.set_device(self.local)
        self.group = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.group = dist.group.WORLD

    def sync_all(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def free(self):
        import gc

        gc.collect()
        self.torch.cuda.empty_cache()


def shard_of(P_total, world, rank, multiple):
    """contiguous ping range of a rank for STRONG scaling: boundaries at multiples of `multiple` (ping_num), chosen so
    that they do not coincide with 20-ping bin edges where possible"""
    per = -(-P_total // world)
    per = -(-per // multiple) * multiple
    if per % 20 == 0 and world > 1 and per + multiple <= P_total:
        per += multiple  # move the boundaries off the 20 s bin edges: the exchange merges real partial bins
    lo = min(P_total, rank * per)
    hi = min(P_total, lo + per)
    return lo, hi


def time_pipeline(ctx, name, P_local, ping_offset, steps, warmup, seed=None, sustained_s=0.0, keep=True):
    """resident fused-pipeline value of one config on this rank's shard; returns a dict of timings"""
    torch = ctx.torch
    from echopype_b200 import pipeline

    c = CONFIGS[name]
    ed = make_echodata(name, P_local, ping_offset, device=True, seed=seed)
    plan = pipeline.FusedPlan(ed, group=ctx.group, **chain_kwargs(name))
    plan.record_events = True
    for _ in range(max(warmup, 3)):
        plan.run()
    ctx.sync_all()
    plan.kernel_events.clear()
    plan.comm_events.clear()
    l0 = plan.launches
    clocks = ClockSampler(ctx.local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        mvbs = plan.run()[0]
    plan.sync_pending()  # N > 1: the side-stream exchange of the last step belongs to the timed region
    e1.record()
    ctx.sync_all()
    clk = clocks.stop()
    ms_total = e0.elapsed_time(e1)
    k_ms = [a.elapsed_time(b) for a, b in plan.kernel_events]
    c_ms = [a.elapsed_time(b) for a, b in plan.comm_events]
    res = {"launches": plan.launches - l0, "clocks": clk, "grid": list(mvbs.shape), "nan_frac": float(torch.isnan(mvbs).float().mean()),
           "n_local": c["C"] * P_local * c["R"]}
    ms_total, k_avg, c_avg = ctx.max_over_ranks([ms_total, sum(k_ms) / len(k_ms), (sum(c_ms) / len(c_ms)) if c_ms else 0.0])
    res.update(ms_total=ms_total, k_avg=k_avg, c_avg=c_avg)
    if sustained_s > 0:  # the same step back to back for >= sustained_s seconds
        n = max(steps, int(sustained_s * 1e3 / max(ms_total / steps, 1e-3)))
        plan.record_events = False
        clocks = ClockSampler(ctx.local)
        clocks.start()
        e0.record()
        for _ in range(n):
            plan.run()
        plan.sync_pending()
        e1.record()
        ctx.sync_all()
        clk2 = clocks.stop()
        (ms_s,) = ctx.max_over_ranks([e0.elapsed_time(e1)])
        res["sustained"] = {"steps": n, "seconds": round(ms_s * 1e-3, 3), "ms_per_step": ms_s / n, "clocks": clk2}
    if keep:
        res["ed"], res["plan"], res["mvbs"] = ed, plan, mvbs
    return res


def roofline_of(name, n_local, k_avg_ms, share, kernel):
    peak, peak_src = measured_peak()
    bps = CONFIGS[name]["bytes_per_sample"]
    achieved = bps * n_local / (k_avg_ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "kernel": kernel, "kernel_ms": k_avg_ms, "algorithmic_bytes_per_sample": bps, "peak_source": peak_src, "share_of_step": share}


def time_bb(ctx, P_local, ping_offset, steps, warmup):
    """cfg3: compute_Sv on a device-resident broadband volume (value: the public call; roofline: K3 alone)."""
    torch = ctx.torch
    import numpy as np

    import echopype_b200 as ep
    from echopype_b200 import kernels
    from echopype_b200.calibrate.calibrate_ek import CalibrateEK80

    c = CONFIGS["cfg3"]
    C, R, B = c["C"], c["R"], c["B"]
    ed = make_echodata("cfg3", P_local, ping_offset, device=True)
    kw = dict(waveform_mode="BB", encode_mode="complex")
    for _ in range(max(2, min(warmup, 3))):
        ds = ep.calibrate.compute_Sv(ed, **kw)
    del ds
    ctx.sync_all()
    clocks = ClockSampler(ctx.local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ds = ep.calibrate.compute_Sv(ed, **kw)
    e1.record()
    ctx.sync_all()
    clk = clocks.stop()
    ms_total = e0.elapsed_time(e1)
    nan_frac = float(torch.isnan(ds["Sv"].data[:, :: max(1, P_local // 64)]).float().mean())
    del ds
    ctx.free()
    # K3 alone through the C ABI
    cal = CalibrateEK80(ed, **kw)
    cal._cal_complex_samples("Sv")
    beam = ed["Sonar/Beam_group1"]
    re, im = beam["backscatter_r"].data, beam["backscatter_i"].data
    tx = [cal._tx[ch] for ch in np.asarray(beam["channel"].values)]
    M = max(len(t) for t in tx)
    ctx.free()
    ev = []
    for i in range(2 + steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = kernels.pulse_compress_sv(re, im, tx, cal.rows, C, P_local, R, B, want_range=True)
        b.record()
        del out
        if i >= 2:
            ev.append((a, b))
    ctx.sync_all()
    k_ms = sorted(a.elapsed_time(b) for a, b in ev)
    ms_total, k_avg = ctx.max_over_ranks([ms_total, sum(k_ms) / len(k_ms)])
    del re, im, ed, cal
    ctx.free()
    return {"ms_total": ms_total, "k_avg": k_avg, "taps": M, "clocks": clk, "nan_frac": nan_frac, "n_local": C * P_local * R}


def verify_sharded(ctx, name, P_v):
    """N ranks on P_v pings each vs ONE GPU on the concatenated volume (rank 0).  Returns a dict for config.verified."""
    torch, dist = ctx.torch, ctx.dist
    from echopype_b200 import kernels, pipeline

    c = CONFIGS[name]
    C, R = c["C"], c["R"]
    kw = chain_kwargs(name)
    off = PING_ORIGIN + ctx.rank * P_v
    ed = make_echodata(name, P_v, off, device=True, seed=c["seed"] + 100 + ctx.rank, nan_tail=0.05)
    plan = pipeline.FusedPlan(ed, group=ctx.group, **kw)
    _, acc, rmax, _, _ = plan.run(finalize=False)
    mv, _ = kernels.bin_finalize(acc, to_db=True)
    # the optional side-stream form of a step (exchange + edge-bin finalisation off the main stream) must give the same grid
    plan.async_exchange = True
    mv_async = plan.run()[0]
    plan.sync_pending()
    plan.async_exchange = False
    torch.cuda.synchronize()
    same = torch.tensor([int(torch.equal(torch.nan_to_num(mv_async, nan=1.0), torch.nan_to_num(mv, nan=1.0))
                             and torch.equal(torch.isnan(mv_async), torch.isnan(mv)))], dtype=torch.int64, device="cuda")
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    async_equal = bool(int(same.item()))
    del mv_async
    x = ed["Sonar/Beam_group1"]["backscatter_r"].data
    parts = [torch.empty_like(x) for _ in range(ctx.world)] if ctx.rank == 0 else None
    dist.gather(x, parts, dst=0)
    # every rank sends its window [x_lo, x_hi], accumulators and finalised bins to rank 0
    meta = torch.tensor([plan.x_lo, plan.x_hi, acc.shape[2]], dtype=torch.int64, device="cuda")
    metas = [torch.empty_like(meta) for _ in range(ctx.world)] if ctx.rank == 0 else None
    dist.gather(meta, metas, dst=0)
    nX_g = len(plan.p_edges) - 1
    pad_acc = torch.zeros((C, nX_g, acc.shape[2], 4), dtype=torch.float64, device="cuda")
    pad_mv = torch.full((C, nX_g, acc.shape[2]), float("nan"), dtype=torch.float32, device="cuda")
    if plan.nX > 0:
        pad_acc[:, plan.x_lo : plan.x_hi + 1] = acc
        pad_mv[:, plan.x_lo : plan.x_hi + 1] = mv
    accs = [torch.empty_like(pad_acc) for _ in range(ctx.world)] if ctx.rank == 0 else None
    mvs = [torch.empty_like(pad_mv) for _ in range(ctx.world)] if ctx.rank == 0 else None
    dist.gather(pad_acc, accs, dst=0)
    dist.gather(pad_mv, mvs, dst=0)
    if rmax is not None:
        plan.wait_rmax(rmax)
    rm = rmax.reshape(-1)[:1].clone() if rmax is not None else torch.zeros(1, dtype=torch.float64, device="cuda")
    result = None
    if ctx.rank == 0:
        xcat = torch.cat(parts, dim=1).contiguous()
        del parts
        ed1 = make_echodata(name, P_v * ctx.world, PING_ORIGIN, device=True, backscatter=xcat, seed=c["seed"] + 100)
        plan1 = pipeline.FusedPlan(ed1, group=None, **kw)
        _, acc1, rmax1, _, _ = plan1.run(finalize=False)
        mv1, _ = kernels.bin_finalize(acc1, to_db=True)
        ok_counts, max_rel, max_db, shared, nan_same = True, 0.0, 0.0, 0, True
        owners = torch.zeros(nX_g, dtype=torch.int64, device="cuda")
        nR = min(acc1.shape[2], accs[0].shape[2])
        for r in range(ctx.world):
            lo, hi, _ = (int(v) for v in metas[r])
            if hi < lo:
                continue
            owners[lo : hi + 1] += 1
            a, b = accs[r][:, lo : hi + 1, :nR], acc1[:, lo : hi + 1, :nR]
            ok_counts &= bool(torch.equal(a[..., 1], b[..., 1]) and torch.equal(a[..., 2], b[..., 2]))
            rel = ((a[..., 0] - b[..., 0]).abs() / b[..., 0].abs().clamp_min(1e-300))[b[..., 1] > 0]
            max_rel = max(max_rel, float(rel.max()) if rel.numel() else 0.0)
            g, w = mvs[r][:, lo : hi + 1, :nR], mv1[:, lo : hi + 1, :nR]
            nan_same &= bool(torch.equal(torch.isnan(g), torch.isnan(w)))
            d = (g - w).abs()
            d = d[~torch.isnan(d)]
            max_db = max(max_db, float(d.max()) if d.numel() else 0.0)
        shared = int((owners > 1).sum())
        rmax_ok = bool(rmax is None or abs(float(rm) - float(rmax1.max())) == 0.0)
        # member counts are integers: bit-equal.  Linear sums: float32 partial sums per CTA (their grouping depends on the
        # tile -> CTA assignment, which differs between the two runs) added into float64 cells: equal to a few float32 ulp.
        # The MVBS itself is a float32 dB value: 1 ulp at -70 dB is 7.6e-6 dB.
        result = {"ok": bool(ok_counts and max_rel <= 4e-6 and nan_same and max_db <= 1.6e-5 and shared >= ctx.world - 1 and rmax_ok and async_equal),
                  "async_side_stream_step_equals_synchronous_step": async_equal,
                  "counts_bit_equal": ok_counts, "max_rel_diff_of_linear_sums": max_rel, "nan_masks_equal": nan_same, "max_abs_dB": max_db,
                  "tolerance": "counts bit-equal; sums 4e-6 relative (float32 partials); MVBS 1.6e-5 dB (2 float32 ulp of the dB value)",
                  "ping_bins_shared_by_two_ranks": shared, "range_max_equal": rmax_ok,
                  "volume": f"{ctx.world} ranks x ({C}ch x {P_v} ping x {R} range) vs one GPU on the concatenated {P_v * ctx.world} pings"}
        del xcat, ed1, plan1, acc1, mv1
    del ed, plan, acc, mv, pad_acc, pad_mv
    ctx.free()
    flag = torch.tensor([1 if (result is None or result["ok"]) else 0], dtype=torch.int64, device="cuda")
    dist.broadcast(flag, src=0)
    return result


def e2e_legs(ctx, name, P_local, ping_offset, steps, seed):
    """public API from pinned host memory: float32 samples, then (EK60) the int16 raw counts"""
    torch = ctx.torch
    import numpy as np

    from echopype_b200 import kernels, pipeline

    c = CONFIGS[name]
    C, R = c["C"], c["R"]
    kw = chain_kwargs(name)
    n_local = C * P_local * R
    out = {}
    x_dev = make_echodata(name, P_local, ping_offset, device=True, seed=seed)["Sonar/Beam_group1"]["backscatter_r"].data
    x_pin = torch.empty(x_dev.shape, dtype=torch.float32, pin_memory=True)
    x_pin.copy_(x_dev)
    torch.cuda.synchronize()
    del x_dev
    ctx.free()
    # the copies alone (same slabs, no kernels): what the host side of this rank can deliver while all ranks copy
    slab = torch.empty((8192, R), dtype=torch.float32, device="cuda")
    ctx.sync_all()
    t0 = time.perf_counter()
    for c_ in range(C):
        for p0 in range(0, P_local, 8192):
            pc = min(8192, P_local - p0)
            slab[:pc].copy_(x_pin[c_, p0 : p0 + pc], non_blocking=True)
    torch.cuda.synchronize()
    (dt_copy,) = ctx.max_over_ranks([time.perf_counter() - t0])
    out["h2d_only"] = {"GBps_per_rank": n_local * 4 / dt_copy / 1e9, "GBps_all_ranks": n_local * 4 * ctx.world / dt_copy / 1e9, "numa": ctx.numa}
    del slab
    ed_host = make_echodata(name, P_local, ping_offset, device=False, backscatter=x_pin.numpy(), seed=seed)
    for _ in range(2):
        ds = pipeline.compute_Sv_clean_MVBS(ed_host, group=ctx.group, **kw)
    ctx.sync_all()
    t0 = time.perf_counter()
    for _ in range(steps):
        ds = pipeline.compute_Sv_clean_MVBS(ed_host, group=ctx.group, **kw)
        host_mvbs = ds["Sv"].values  # already a host array (D2H happened inside the call)
    torch.cuda.synchronize()
    (dt,) = ctx.max_over_ranks([time.perf_counter() - t0])
    d2h = int(host_mvbs.size * 4)
    out["e2e"] = {"value": n_local * ctx.world * steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(n_local * 4), "d2h_bytes_per_step": d2h,
                  "steps": steps, "ms_per_step": dt / steps * 1e3,
                  "api": "echopype_b200.pipeline.compute_Sv_clean_MVBS(echodata with pinned-host backscatter_r)"}
    out["host_mvbs"] = host_mvbs
    del ed_host, x_pin
    if name != "cfg2":
        return out
    # the same call on RAW POWER COUNTS (int16): 2 bytes per sample over PCIe
    q_dev = kernels.synth_fill_i16((C, P_local, R), seed=seed, nan_tail=0.005, ping_offset=ping_offset)  # the same volume
    q_pin = torch.empty(q_dev.shape, dtype=torch.int16, pin_memory=True)
    q_pin.copy_(q_dev)
    torch.cuda.synchronize()
    del q_dev
    ctx.free()
    ed_raw = make_echodata(name, P_local, ping_offset, device=False, backscatter=q_pin.numpy(), seed=seed)
    for _ in range(2):
        ds = pipeline.compute_Sv_clean_MVBS(ed_raw, group=ctx.group, **kw)
    ctx.sync_all()
    t0 = time.perf_counter()
    for _ in range(steps):
        ds = pipeline.compute_Sv_clean_MVBS(ed_raw, group=ctx.group, **kw)
        raw_mvbs = ds["Sv"].values
    torch.cuda.synchronize()
    (dt_raw,) = ctx.max_over_ranks([time.perf_counter() - t0])
    raw_same = bool(raw_mvbs.shape == host_mvbs.shape and np.array_equal(np.isnan(raw_mvbs), np.isnan(host_mvbs))
                    and np.allclose(raw_mvbs, host_mvbs, rtol=0, atol=1e-6, equal_nan=True))
    out["e2e_raw_counts"] = {
        "value": n_local * ctx.world * steps / dt_raw, "unit": UNIT, "h2d_bytes_per_step": int(n_local * 2), "d2h_bytes_per_step": d2h,
        "steps": steps, "ms_per_step": dt_raw / steps * 1e3, "matches_float32_e2e_within_1e-6_dB": raw_same,
        "api": "the same call on an EchoData holding the int16 raw power counts of the datagrams (pinned host; -32768 = NaN "
               "padding), converted in registers by the fused kernel"}
    del ed_raw, q_pin
    ctx.free()
    return out


def extra_kernels(ctx, ed, steps):
    """K1 (compute_Sv alone) on the headline volume: CUDA events per launch, same run as the headline numbers."""
    torch = ctx.torch
    from echopype_b200 import kernels
    from echopype_b200.calibrate.calibrate_ek import CalibrateEK60

    peak, _ = measured_peak()
    beam = ed["Sonar/Beam_group1"]
    x = beam["backscatter_r"].data
    C, P, R = (int(s) for s in x.shape)
    n = C * P * R
    cal = CalibrateEK60(ed)
    rows = cal._power_row_builder("Sv").build()
    out, rng = torch.empty_like(x), torch.empty_like(x)

    def timed(fn, nbytes):
        for _ in range(3):
            fn()
        ev = []
        for _ in range(max(5, steps)):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            ev.append((a, b))
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)
        avg = sum(ms) / len(ms)
        return {"ms": avg, "ms_min": ms[0], "GBps": nbytes / avg / 1e6, "frac_of_measured_hbm": nbytes / avg / 1e6 / peak,
                "Gsamples_s": n / avg / 1e6, "launches_timed": len(ms)}

    res = {
        "K1 sv_power (compute_Sv, 8 B/sample: power in, Sv out)": timed(lambda: kernels.sv_power(x, rows, C, P, R, want_range=False, out=out), 8 * n),
        "K1 sv_power + echo_range (12 B/sample)": timed(lambda: kernels.sv_power(x, rows, C, P, R, want_range=True, out=out, rng=rng), 12 * n),
    }
    del out, rng, rows, cal
    ctx.free()
    # the reference's own test setting of the noise removal (tests/utils/test_processinglevels_integration.py:111):
    # ping_num = 10 > 8 takes the two-sweep form of the fused kernel (sub-tiles of 5 rows, second read from L2)
    from echopype_b200 import pipeline

    def fused(label, nbytes, **kw):
        plan = pipeline.FusedPlan(ed, SNR_threshold=SNR, range_bin=RANGE_BIN, ping_time_bin=PING_BIN, **kw)
        plan.record_events = True
        for _ in range(3):
            plan.run()
        torch.cuda.synchronize()
        plan.kernel_events.clear()
        for _ in range(max(5, steps)):
            plan.run()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in plan.kernel_events)
        avg = sum(ms) / len(ms)
        res[label] = {"ms": avg, "ms_min": ms[0], "GBps": nbytes * n / avg / 1e6, "frac_of_measured_hbm": nbytes * n / avg / 1e6 / peak,
                      "Gsamples_s": n / avg / 1e6, "launches_timed": len(ms)}
        del plan
        ctx.free()

    fused("fused pipeline, remove_background_noise(ping_num=10, range_sample_num=20) (4 B/sample)", 4, ping_num=10, range_sample_num=20)
    # full-size outputs streamed out of the same kernel (keep=): 4 B in + 4 B out per kept array
    fused("fused pipeline, keep=('Sv_corrected',) (8 B/sample)", 8, ping_num=PING_NUM, range_sample_num=RS_NUM, keep=("Sv_corrected",))
    fused("fused pipeline, keep=('Sv','echo_range','Sv_noise','Sv_corrected') (20 B/sample)", 20, ping_num=PING_NUM,
          range_sample_num=RS_NUM, keep=("Sv", "echo_range", "Sv_noise", "Sv_corrected"))
    return res


def run_b200(args):
    ctx = Ctx()
    torch = ctx.torch
    world, rank = ctx.world, ctx.rank
    if args.gpus != world and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    name = args.config
    c = CONFIGS[name]
    mult = PING_NUM if c.get("noise") else 1
    if args.pings:
        c["P"] = args.pings
    if args.scaling == "strong" and world > 1:
        lo, hi = shard_of(c["P"], world, rank, mult)
        P_local, ping_offset, P_total = hi - lo, PING_ORIGIN + lo, c["P"]
    else:
        P_local, ping_offset, P_total = c["P"], PING_ORIGIN + rank * c["P"], c["P"]
    seed = c["seed"] + rank
    # N > 1: the first point-to-point exchanges after start-up run slower (NCCL channel warm-up, rank skew): 10 warm-up steps
    steps, warmup = args.steps, max(args.warmup, 3 if world == 1 else 10)

    line = {"metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32"}
    config = {"workload": workload_name(name, P_total, world, args.scaling),
              "l2": "inputs (GBs per GPU) far exceed the 126 MB L2; no flush needed",
              "ping_axis": f"1 s ping interval starting {PING_ORIGIN} s after a {PING_BIN} bin edge: shard boundaries fall inside ping bins"}
    if c["kind"] == "pipeline":
        res = time_pipeline(ctx, name, P_local, ping_offset, steps, warmup, seed=seed, sustained_s=(0.0 if args.no_extra else 2.0))
        n_total = float(c["C"] * P_total * c["R"]) * (world if args.scaling == "weak" else 1)
        line.update(value=n_total * steps / (res["ms_total"] * 1e-3), ms_per_step=res["ms_total"] / steps,
                    data="synthetic (Philox4x32-10 on device, int16-quantised power, 0.5% NaN-padded pings)")
        config.update(parallelism=f"ping_time sharded over {world} GPU(s); neighbour exchange of the straddling ping bins only",
                      collectives_ms_per_step=round(res["c_avg"], 4), mvbs_grid=res["grid"], mvbs_nan_frac=round(res["nan_frac"], 4))
        kern = ("pipeline_fast_kernel (epb_pipeline_power_mvbs: power -> Sv -> " + ("noise removal -> " if c.get("noise") else "") + "MVBS accumulators)")
        line["roofline"] = roofline_of(name, res["n_local"], res["k_avg"], res["k_avg"] * steps / res["ms_total"], kern)
        line["gpu_launches"] = res["launches"]
        line["clocks"] = res["clocks"]
        extra = {}
        if "sustained" in res:
            s = res["sustained"]
            s["value"] = n_total / (s["ms_per_step"] * 1e-3)
            extra["sustained"] = s
        resident_mvbs = res["mvbs"]
        if not args.no_extra and name == "cfg2":
            extra["kernels"] = extra_kernels(ctx, res["ed"], steps)
            res.pop("plan"), res.pop("ed")
            ctx.free()
            # noise tiles that straddle ping bins (time origin 7 s after a bin edge): the general tile path of the fused kernel
            u = time_pipeline(ctx, name, P_local, ping_offset - PING_ORIGIN + 7, steps, 3, seed=seed, keep=False)
            extra["unaligned_origin_7s"] = {"ms_per_step": u["ms_total"] / steps, "kernel_ms": u["k_avg"], "value": n_total * steps / (u["ms_total"] * 1e-3),
                                            "roofline_frac": roofline_of(name, u["n_local"], u["k_avg"], 1.0, "")["frac"]}
        res.pop("plan", None), res.pop("ed", None)
        res.pop("mvbs")
        ctx.free()
        # ---- end to end through the public API from pinned host memory -------------------------------------------------
        e2e_P = P_local if name != "cfg5" else min(P_local, 100_000)
        legs = e2e_legs(ctx, name, e2e_P, ping_offset, max(1, min(steps, args.e2e_steps)), seed)
        line["e2e"] = legs["e2e"]
        line["e2e"]["h2d_only"] = legs["h2d_only"]
        if e2e_P != P_local:
            line["e2e"]["sample"] = f"first {e2e_P} pings of the rank's shard (pinned host memory for the full 98 GB volume is not assumed)"
        if "e2e_raw_counts" in legs:
            line["e2e_raw_counts"] = legs["e2e_raw_counts"]
        if e2e_P == P_local:
            import numpy as np

            hm = legs["host_mvbs"]
            config["e2e_matches_resident_nan_mask"] = bool(np.array_equal(np.isnan(hm), np.isnan(resident_mvbs.cpu().numpy()[:, : hm.shape[1], : hm.shape[2]])))
        del resident_mvbs, legs
        ctx.free()
    else:  # cfg3
        res = time_bb(ctx, P_local, ping_offset, max(2, min(steps, 5)), warmup)
        n_total = float(c["C"] * P_total * c["R"]) * (world if args.scaling == "weak" else 1)
        st = max(2, min(steps, 5))
        line.update(value=n_total * st / (res["ms_total"] * 1e-3), ms_per_step=res["ms_total"] / st, steps=st,
                    data="synthetic (Philox4x32-10 on device, N(0,1) x 1e-3 complex samples, 0.5% NaN-padded pings)")
        config.update(parallelism=f"ping_time sharded over {world} GPU(s); no collective", replica_taps=res["taps"], sv_nan_frac=round(res["nan_frac"], 4))
        line["roofline"] = roofline_of(name, res["n_local"], res["k_avg"], res["k_avg"] * st / res["ms_total"], "pulse compression + Sv epilogue (epb_pulse_compress_sv)")
        line["gpu_launches"] = 2 * st
        line["clocks"] = res["clocks"]
        line["e2e"] = None
        extra = {}

    # ---- the other configs, measured in the same run (resident, strong scaling over the N ranks) -------------------------
    if not args.no_extra and name == "cfg2":
        others = {}
        for other in ("cfg4", "cfg5", "cfg3"):
            oc = CONFIGS[other]
            m = PING_NUM if oc.get("noise") else 1
            lo, hi = shard_of(oc["P"], world, rank, m)
            try:
                if oc["kind"] == "pipeline":
                    r = time_pipeline(ctx, other, hi - lo, PING_ORIGIN + lo, max(3, min(steps, 10)), 3, seed=oc["seed"] + rank, keep=False)
                    st = max(3, min(steps, 10))
                    n_tot = float(oc["C"] * oc["P"] * oc["R"])
                    others[other] = {
                        "workload": workload_name(other, oc["P"], world, "strong"), "scaling": "strong", "value": n_tot * st / (r["ms_total"] * 1e-3),
                        "unit": UNIT, "ms_per_step": r["ms_total"] / st, "steps": st, "collectives_ms_per_step": round(r["c_avg"], 4),
                        "roofline": roofline_of(other, r["n_local"], r["k_avg"], r["k_avg"] * st / r["ms_total"], "pipeline_fast_kernel"),
                        "mvbs_grid": r["grid"], "clocks": r["clocks"]}
                    del r
                else:
                    st = 3
                    r = time_bb(ctx, hi - lo, PING_ORIGIN + lo, st, 2)
                    n_tot = float(oc["C"] * oc["P"] * oc["R"])
                    others[other] = {
                        "workload": workload_name(other, oc["P"], world, "strong"), "scaling": "strong", "value": n_tot * st / (r["ms_total"] * 1e-3),
                        "unit": UNIT, "ms_per_step": r["ms_total"] / st, "steps": st, "replica_taps": r["taps"],
                        "roofline": roofline_of(other, r["n_local"], r["k_avg"], r["k_avg"] * st / r["ms_total"], "epb_pulse_compress_sv (K3)"),
                        "complex_input_values_per_s": n_tot * oc["B"] * st / (r["ms_total"] * 1e-3), "clocks": r["clocks"]}
            except Exception as e:  # noqa: BLE001 - an extra leg must not take the headline line down
                others[other] = {"error": repr(e)[:300]}
            ctx.free()
        extra["configs"] = others

    # ---- N > 1: the sharded result against the single-GPU result of the concatenated volume ---------------------------------
    if world > 1 and not args.no_verify and c["kind"] == "pipeline":
        try:
            v = verify_sharded(ctx, name, args.verify_pings)
        except Exception as e:  # noqa: BLE001
            v = {"ok": False, "error": repr(e)[:300]}
        if rank == 0:
            config["verified"] = bool(v["ok"])
            config["verification"] = v

    if rank == 0:
        line["config"] = config
        if extra:
            line["extra"] = extra
        traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_file) and name in ("cfg2",):
            tr = json.load(open(traffic_file))
            per_sample = tr.get("pipeline_kernel_dram_bytes_per_sample")
            if per_sample:
                line["roofline"]["traffic"] = per_sample * c["C"] * P_local * c["R"]
                line["roofline"]["traffic_source"] = tr.get("source")
        if world == 1 and not args.no_cpu:
            Ps = args.cpu_sample_pings if name != "cfg3" else 40
            v, secs = cpu_baseline_single(name, Ps)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"{c['C']}ch x {Ps} ping x {c['R']} range of the {name} volume, {secs:.1f} s, numpy float64 oracle",
                "host_cores_available": os.cpu_count(),
            }
        order = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                 "config", "roofline", "e2e", "e2e_raw_counts", "gpu_launches", "clocks", "cpu_baseline", "extra"]
        print(json.dumps({k: line[k] for k in order if k in line}), flush=True)
    if world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()
    _ = torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--pings", type=int, default=0, help="override the config's ping count (per GPU for weak scaling)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample-pings", type=int, default=4000)
    ap.add_argument("--ref-pings-per-core", type=int, default=1000)
    ap.add_argument("--verify-pings", type=int, default=5000, help="pings per rank of the N-rank vs 1-rank verification volume")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip extra.kernels / extra.configs / extra.sustained")
    ap.add_argument("--no-verify", action="store_true", help="skip the N-rank vs 1-rank verification (N > 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
