#!/usr/bin/env python
"""bench.py - headline benchmark of echopype_b200 (contract: task statement (4)).

Workload (BASELINE.json configs[1], "cfg2"): EK60 power volume 4 channels x 100 000 pings x 4096 range samples
per GPU (6.55 GB float32, far above the 126 MB L2 so no flush is needed between steps) through
    compute_Sv -> remove_background_noise(ping_num=5, range_sample_num=30, SNR 3 dB) -> compute_MVBS("20m", "20s").
One "step" = one pass of that chain over the whole volume.  Metric: samples/s (1 sample = one (channel, ping,
range_sample) element), whole job over all GPUs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--pings P]

* own arm, `value`: inputs resident in HBM; a step launches the row-setup kernel, the exact range-maximum
  kernels, the accumulator memset, the fused pipeline kernel (epb_pipeline_power_mvbs), [N>1: the straddling-bin
  all-reduce] and the mean->dB kernel.  Timed with CUDA events on the launching stream, barrier + synchronize on
  both sides, max over ranks.  `roofline` is the fused kernel alone (events around each launch), 4 algorithmic
  bytes per sample, against MEASURED_PEAKS.json hbm_gbs.
* own arm, `e2e`: the public call echopype_b200.pipeline.compute_Sv_clean_MVBS(echodata) on an EchoData whose
  backscatter_r lives in PINNED HOST memory: host parameter assembly + streamed H2D of the volume + kernels + D2H
  of the MVBS grid are all inside the timed region (wall clock with synchronize on both sides, max over ranks).
* `cpu_baseline` (rank 0, N=1): the numpy float64 oracle (a port of the reference's operation sequence; the
  reference itself cannot be imported in this image) on one core over a bounded ping sample.
* `--impl reference`: the same oracle chain, ping-sharded over all host cores with multiprocessing (the
  emulation of the reference's dask-chunked path), each step a bounded sample of the workload.

Multi-GPU (torchrun, one rank per GPU): weak scaling - every rank holds its own 100 000-ping shard of one
global time-ordered volume (ping offsets rank*P); the only data-path collective is the straddling-bin reduce.
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
C, R = 4, 4096
PING_NUM, RS_NUM, SNR, RANGE_BIN, PING_BIN = 5, 30, "3.0dB", "20m", "20s"
SEED = 2000


def workload_name(P):
    return (f"cfg2: EK60 Sv->remove_noise(ping_num={PING_NUM},range_sample_num={RS_NUM},SNR={SNR})->compute_MVBS({RANGE_BIN},{PING_BIN}), "
            f"{C}ch x {P} ping x {R} range per GPU")


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
        self.samples, self.reasons, self.max_mhz, self._stop, self._t = [], set(), None, threading.Event(), None
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
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.02)

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
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------------------------
# CPU oracle chain (cpu_baseline and --impl reference); the only place bench.py executes oracle/
# ---------------------------------------------------------------------------------------------------------------
_W = {}


def _oracle_make(P, ping_offset, seed):
    from echopype_b200 import synth

    return synth.make_ek60(C, P, R, seed=seed, nan_tail=0.005, ping_offset=ping_offset)


def _oracle_chain(ed):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_glue as og
    from oracle import clean as oclean
    from oracle import commongrid as ogrid

    ref = og.ek60(ed, "Sv")
    nz = oclean.remove_background_noise(ref["out"], ref["echo_range"], ref["sound_absorption"], PING_NUM, RS_NUM, None, SNR)
    pt = np.asarray(ed["Sonar/Beam_group1"]["ping_time"].values).astype("datetime64[ns]").astype(np.int64)
    mv = ogrid.compute_MVBS(nz["Sv_corrected"], ref["echo_range"], pt, range_bin=RANGE_BIN, ping_time_bin=PING_BIN)
    return mv["Sv"]


def _worker_init(P, seed):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    _W["P"], _W["seed"] = P, seed


def _worker_prepare(i):
    _W["ed"] = _oracle_make(_W["P"], i * _W["P"], _W["seed"] + i)
    return i


def _worker_step(_):
    if "ed" not in _W:  # a worker that was not handed a prepare task
        _worker_prepare(os.getpid() % 1000)
    return float(_oracle_chain(_W["ed"]).shape[1])


def cpu_baseline_single(P_sample):
    ed = _oracle_make(P_sample, 0, SEED)
    _oracle_chain(_oracle_make(100, 0, SEED))  # warm imports / allocator
    t0 = time.perf_counter()
    _oracle_chain(ed)
    dt = time.perf_counter() - t0
    return C * P_sample * R / dt, dt


def run_reference(args):
    """--impl reference: the oracle chain on all host cores, ping-sharded; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    cores = os.cpu_count() or 1
    P_w = args.ref_pings_per_core
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_worker_init, initargs=(P_w, SEED)) as pool:
        # one resident shard per worker process (imap with chunksize 1 over `cores` idle workers)
        list(pool.imap(_worker_prepare, range(cores), chunksize=1))
        for _ in range(args.warmup):
            pool.map(_worker_step, range(cores), chunksize=1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_worker_step, range(cores), chunksize=1)
        dt = time.perf_counter() - t0
    n_step = C * P_w * R * cores
    value = n_step * args.steps / dt
    sample = f"{cores} processes x ({C}ch x {P_w} ping x {R} range) per step = {n_step} samples/step of the cfg2 volume"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.pings), "reference": "numpy/scipy/pandas float64 port of the reference chain "
                   "(oracle/; echopype itself is not importable in this image), ping-sharded over all host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        group = dist.group.WORLD
    if args.gpus != world and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    import echopype_b200 as ep
    from echopype_b200 import pipeline, synth

    P = args.pings
    n_local = C * P * R
    kw = dict(ping_num=PING_NUM, range_sample_num=RS_NUM, SNR_threshold=SNR, range_bin=RANGE_BIN, ping_time_bin=PING_BIN)
    ed = synth.make_ek60(C, P, R, seed=SEED + rank, device=True, nan_tail=0.005, ping_offset=rank * P)
    plan = pipeline.FusedPlan(ed, group=group, **kw)
    plan.record_events = True

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident value ---------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        plan.run()
    sync_all()
    plan.kernel_events.clear()
    plan.comm_events.clear()
    l0 = plan.launches
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        mvbs = plan.run()[0]
    e1.record()
    sync_all()
    clk = clocks.stop()
    ms_total = e0.elapsed_time(e1)
    launches = plan.launches - l0
    k_ms = [a.elapsed_time(b) for a, b in plan.kernel_events]
    k_avg = sum(k_ms) / len(k_ms)
    c_ms = [a.elapsed_time(b) for a, b in plan.comm_events]
    c_avg = sum(c_ms) / len(c_ms) if c_ms else 0.0
    t = torch.tensor([ms_total, k_avg, c_avg], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, k_avg, c_avg = float(t[0]), float(t[1]), float(t[2])
    value = n_local * world * args.steps / (ms_total * 1e-3)
    nan_frac = float(torch.isnan(mvbs).float().mean())
    plan.record_events = False

    # ---- end to end through the public API from pinned host memory -----------------------------------------------
    x_dev = ed["Sonar/Beam_group1"]["backscatter_r"].data
    x_pin = torch.empty(x_dev.shape, dtype=torch.float32, pin_memory=True)
    x_pin.copy_(x_dev)
    torch.cuda.synchronize()
    ed_host = synth.make_ek60(C, P, R, seed=SEED + rank, ping_offset=rank * P, backscatter=x_pin.numpy())
    del ed, plan, x_dev
    torch.cuda.empty_cache()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(2):
        ds = pipeline.compute_Sv_clean_MVBS(ed_host, group=group, **kw)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ds = pipeline.compute_Sv_clean_MVBS(ed_host, group=group, **kw)
        host_mvbs = ds["Sv"].values  # already a host array (D2H happened inside the call)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    d2h = int(host_mvbs.size * 4)
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t[0])
    e2e_value = n_local * world * e2e_steps / dt
    same = bool(np.array_equal(np.isnan(host_mvbs), np.isnan(mvbs.cpu().numpy()[:, : host_mvbs.shape[1]])))

    # ---- the same call on RAW POWER COUNTS (int16): 2 bytes per sample over PCIe ----------------
    from echopype_b200 import kernels

    del ed_host, x_pin
    q_dev = kernels.synth_fill_i16((C, P, R), seed=SEED + rank, nan_tail=0.005, ping_offset=rank * P)  # the same volume
    q_pin = torch.empty(q_dev.shape, dtype=torch.int16, pin_memory=True)
    q_pin.copy_(q_dev)
    torch.cuda.synchronize()
    del q_dev
    torch.cuda.empty_cache()
    ed_raw = synth.make_ek60(C, P, R, seed=SEED + rank, ping_offset=rank * P, backscatter=q_pin.numpy())
    for _ in range(2):
        ds = pipeline.compute_Sv_clean_MVBS(ed_raw, group=group, **kw)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ds = pipeline.compute_Sv_clean_MVBS(ed_raw, group=group, **kw)
        raw_mvbs = ds["Sv"].values
    torch.cuda.synchronize()
    dt_raw = time.perf_counter() - t0
    t = torch.tensor([dt_raw], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt_raw = float(t[0])
    raw_same = bool(raw_mvbs.shape == host_mvbs.shape and np.array_equal(np.isnan(raw_mvbs), np.isnan(host_mvbs))
                    and np.allclose(raw_mvbs, host_mvbs, rtol=0, atol=1e-6, equal_nan=True))

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = 4.0 * n_local / (k_avg * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (Philox4x32-10 on device, int16-quantised EK60 power, 0.5% NaN-padded pings)",
            "config": {
                "workload": workload_name(P), "l2": "inputs (6.55 GB/GPU) far exceed the 126 MB L2; no flush needed",
                "parallelism": f"ping_time sharded over {world} GPU(s); straddling-bin all-reduce only",
                "collectives_ms_per_step": round(c_avg, 4),
                "mvbs_grid": list(mvbs.shape), "mvbs_nan_frac": round(nan_frac, 4), "e2e_matches_resident_nan_mask": same,
            },
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": "pipeline_fast_kernel<ping_num=5, 2 column groups, noise> (epb_pipeline_power_mvbs)", "kernel_ms": k_avg,
                "algorithmic_bytes_per_sample": 4, "peak_source": peak_src,
                "share_of_step": k_avg * args.steps / ms_total,
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n_local * 4), "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
                    "api": "echopype_b200.pipeline.compute_Sv_clean_MVBS(echodata with pinned-host backscatter_r)"},
            "e2e_raw_counts": {"value": n_local * world * e2e_steps / dt_raw, "unit": UNIT, "h2d_bytes_per_step": int(n_local * 2),
                               "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": dt_raw / e2e_steps * 1e3,
                               "matches_float32_e2e_within_1e-6_dB": raw_same,
                               "api": "the same call on an EchoData holding the int16 raw power counts of the datagrams "
                                      "(pinned host; -32768 = NaN padding), converted in registers by the fused kernel"},
            "gpu_launches": launches,
            "clocks": clk,
        }
        traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_file):
            tr = json.load(open(traffic_file))
            per_sample = tr.get("pipeline_kernel_dram_bytes_per_sample")
            if per_sample:
                line["roofline"]["traffic"] = per_sample * n_local
                line["roofline"]["traffic_source"] = tr.get("source")
        if world == 1 and not args.no_cpu:
            v, secs = cpu_baseline_single(args.cpu_sample_pings)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"{C}ch x {args.cpu_sample_pings} ping x {R} range of the cfg2 volume, {secs:.1f} s, numpy float64 oracle",
                "host_cores_available": os.cpu_count(),
            }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    _ = ep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pings", type=int, default=100000, help="pings per GPU (cfg2: 100000)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample-pings", type=int, default=4000)
    ap.add_argument("--ref-pings-per-core", type=int, default=1000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
