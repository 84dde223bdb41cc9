"""Fused Sv -> background-noise removal -> MVBS pipeline (one pass over HBM, 4 bytes per sample).

The reference runs this chain as three calls that each materialise (channel, ping_time, range_sample) arrays:
``compute_Sv`` (calibrate/api.py:249) -> ``remove_background_noise`` (clean/api.py:436) -> ``compute_MVBS``
(commongrid/api.py:31).  :func:`compute_Sv_clean_MVBS` gives the result of that chain (same arguments, same
output Dataset as ``compute_MVBS`` applied to ``Sv_corrected``) from the raw power samples with a single
kernel (epb_pipeline_power_mvbs); the full-size intermediates are produced only on request (``keep=``).

Multi-GPU: the volume shards over ``ping_time`` (one process per GPU, each holding a contiguous ping range
whose length is a multiple of ``ping_num`` on all but the last rank).  The ping-bin grid is global (derived from
the first / last ping time over all ranks), every rank accumulates (sum, count) for its own pings and ONE
all-reduce(sum) over the small accumulator grid merges the bins that straddle shard boundaries.
"""

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


def _all_reduce(t, op, group):
    import torch.distributed as dist

    dist.all_reduce(t, op=op, group=group)
    return t


def global_ping_edges(ping_time, ping_time_bin, group=None):
    """Ping-bin edges of commongrid/api.py:118-124 for the GLOBAL ping axis: the pandas resample grid depends
    only on the first and the last ping time, which are all-reduced (min / max) across the ranks."""
    pt = np.asarray(ping_time).astype("datetime64[ns]").astype(np.int64)
    lo, hi = int(pt.min()), int(pt.max())
    if group is not None:
        import torch.distributed as dist

        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        t = torch.tensor([-lo, hi], dtype=torch.int64, device=dev)
        _all_reduce(t, dist.ReduceOp.MAX, group)
        lo, hi = -int(t[0].item()), int(t[1].item())
    return ping_time_edges(np.array([lo, hi], dtype="datetime64[ns]"), ping_time_bin)


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
    group : torch.distributed process group for ping-sharded execution (each rank passes its own shard).
    finalize : when False the raw accumulators (device tensor [C, nX, nR, 4]) are returned in ``attrs["acc"]``.
    """
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
    rb = _parse_x_bin(range_bin, "range_bin")
    if closed not in ["right", "left"]:
        raise ValueError(f"{closed} is not a valid option. Options are 'left' or 'right'.")
    if not isinstance(ping_time_bin, str):
        raise TypeError("ping_time_bin must be a string")
    do_noise = ping_num is not None
    snr = extract_dB(SNR_threshold) if do_noise else 0.0
    noise_max = extract_dB(background_noise_max) if (do_noise and background_noise_max is not None) else None
    if do_noise and (range_sample_num is None or int(ping_num) <= 0 or int(range_sample_num) <= 0):
        raise ValueError("ping_num and range_sample_num must be positive integers")

    dev = require_cuda()
    cal_obj = CALIBRATOR[echodata.sonar_model](
        echodata, env_params=cal_kwargs.pop("env_params", None), cal_params=cal_kwargs.pop("cal_params", None),
        ecs_file=cal_kwargs.pop("ecs_file", None), waveform_mode=waveform_mode, encode_mode=encode_mode,
        drop_last_hanning_zero=cal_kwargs.pop("drop_last_hanning_zero", False), slice_dict={},
    )
    if cal_kwargs:
        raise TypeError(f"unexpected keyword arguments {sorted(cal_kwargs)}")
    rows, x, C, P, R, tau_effective = cal_obj._power_rows("Sv")
    beam = echodata[getattr(cal_obj, "ed_beam_group", None) or "Sonar/Beam_group1"]

    # bin edges: range from the exact nanmax of echo_range (commongrid/api.py:108-115), pings from the global grid
    if range_var_max is None:
        rmax = kernels.range_max(x, rows, C, P, R)
        if group is not None:
            import torch.distributed as dist

            t = torch.tensor([rmax if rmax == rmax else -np.inf], dtype=torch.float64, device=dev)
            _all_reduce(t, dist.ReduceOp.MAX, group)
            rmax = float(t.item())
    else:
        rmax = _parse_x_bin(range_var_max) + 1e-8
    r_edges = range_edges(rmax, rb)
    pt = np.asarray(beam["ping_time"].values)
    p_edges = global_ping_edges(pt, ping_time_bin, group)
    xbin_np = assign_bins(pt, p_edges, closed)
    nX, nR = len(p_edges) - 1, len(r_edges) - 1
    xbin = torch.from_numpy(xbin_np).to(dev)
    edges_t = torch.from_numpy(np.ascontiguousarray(r_edges, dtype=np.float64)).to(dev)
    acc = kernels.new_acc(C, nX, nR, dev)
    outs = {k: (empty((C, P, R), device=dev) if k in keep else None) for k in _KEEP}
    noise = empty((C, -(-P // int(ping_num))), device=dev) if do_noise else None
    kernels.pipeline_power_mvbs(
        x, rows, xbin, edges_t, acc, C, P, R, nX, int(ping_num) if do_noise else 0, int(range_sample_num) if do_noise else 0,
        noise_max=noise_max, snr=snr, closed_right=(closed == "right"), noise_out=noise, Sv=outs["Sv"],
        echo_range=outs["echo_range"], Sv_noise=outs["Sv_noise"], Sv_corrected=outs["Sv_corrected"],
    )
    if group is not None:
        import torch.distributed as dist

        _all_reduce(acc, dist.ReduceOp.SUM, group)  # the single data-path collective: straddling bins merge here
    ds = Dataset(coords={"ping_time": p_edges[:-1], "channel": beam["channel"].values, "echo_range": r_edges[:-1]})
    if finalize:
        mvbs, _ = kernels.bin_finalize(acc, skipna=skipna, fill_value=fill_value, to_db=True)
        ds["Sv"] = (("channel", "ping_time", "echo_range"), mvbs.cpu().numpy().astype(np.float64))
        _set_MVBS_attrs(ds)
        ds["echo_range"].attrs.update({"long_name": "Range distance", "units": "m"})
        resvalue, reslabel = ping_time_bin_parsing_and_conversion(ping_time_bin)
        ds["Sv"].attrs.update(
            {
                "cell_methods": (
                    f"ping_time: mean (interval: {resvalue} {reslabel} "
                    "comment: ping_time is the interval start) "
                    f"echo_range: mean (interval: {rb} meter "
                    "comment: echo_range is the interval start)"
                ),
                "binning_mode": "physical units",
                "range_meter_interval": str(rb) + "m",
                "ping_time_interval": ping_time_bin,
            }
        )
    else:
        ds.attrs["acc"] = acc
    ds["frequency_nominal"] = beam["frequency_nominal"]
    prov = echopype_prov_attrs(process_type="processing")
    prov["processing_function"] = "pipeline.compute_Sv_clean_MVBS"
    ds.attrs.update(prov)
    if do_noise:
        ds.attrs["noise_estimate"] = DataArray(noise, ("channel", "ping_tile"), name="noise")
    if keep:
        kept = Dataset(coords={d: beam[d].values for d in DIMS})
        for k in keep:
            da = DataArray(outs[k], DIMS, name=k)
            if k == "echo_range":
                da.law = {"rows": rows, "kind": "echo_range", "minmax": None}
            kept[k] = da
        ds.attrs["kept"] = kept
    ds.attrs["_rows"] = rows  # keeps the row table alive for callers that re-run the kernel
    return ds
