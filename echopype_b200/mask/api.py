"""mask.frequency_differencing / mask.apply_mask (echopype/mask/api.py:467-676, :307-464) with the
O(channel x ping x range) comparisons and selections on the device (epb_freq_diff_mask, epb_apply_mask).  Masks are
(ping_time, range_sample) or (channel, ping_time, range_sample) boolean arrays; on the device they are uint8 tensors."""

import datetime
from typing import List, Union

import numpy as np
import torch

from .. import kernels
from ..dataset import DataArray, Dataset, as_dataset
from ..device import require_cuda, to_device_f32
from ..utils.prov import echopype_prov_attrs, insert_input_processing_level
from .freq_diff import _check_freq_diff_source_Sv, _parse_freq_diff_eq

_OPS = {">": 0, "<": 1, "<=": 2, ">=": 3, "==": 4}


def _history():
    return f"{datetime.datetime.now(datetime.UTC)}. `depth` calculated using:"


def frequency_differencing(source_Sv, storage_options: dict = {}, freqABEq: str = None, chanABEq: str = None) -> DataArray:
    """Mask where ``Sv[chanA] - Sv[chanB] <operator> diff`` (arguments, errors and attrs as the reference).  Returns a
    (ping_time, range_sample) boolean DataArray named ``mask`` held on the device as uint8."""
    freqAB, chanAB, operator, diff = _parse_freq_diff_eq(freqABEq, chanABEq)
    if isinstance(source_Sv, str):
        raise NotImplementedError("file inputs are outside the accelerated path; pass a Dataset")
    source_Sv = as_dataset(source_Sv)
    _check_freq_diff_source_Sv(source_Sv, freqAB, chanAB)
    chan = [str(c) for c in np.asarray(source_Sv["channel"].values).tolist()]
    if freqAB is not None:
        f = np.asarray(source_Sv["frequency_nominal"].values, dtype=np.float64)
        chanA = chan[int(np.argwhere(f == freqAB[0]).flatten()[0])]
        chanB = chan[int(np.argwhere(f == freqAB[1]).flatten()[0])]
    else:
        chanA, chanB = chanAB
    a, b = chan.index(chanA), chan.index(chanB)
    sv = source_Sv["Sv"]
    if tuple(sv.dims) != ("channel", "ping_time", "range_sample"):
        raise ValueError("Sv must have dims ('channel', 'ping_time', 'range_sample')")
    dev = require_cuda()
    C, P, R = sv.shape
    sv_t = to_device_f32(sv.data, dev)
    m = kernels.freq_diff_mask(sv_t, a, b, _OPS[operator], float(diff), C, P, R)
    da = DataArray(m, ("ping_time", "range_sample"), name="mask",
                   coords={k: source_Sv[k].values for k in ("ping_time", "range_sample") if k in source_Sv.coords})
    da.attrs.update({
        "mask_type": "frequency differencing",
        "history": f"{_history()}. Mask created by mask.frequency_differencing. Operation: Sv['{chanA}'] - Sv['{chanB}'] {operator} {diff}",
    })
    return da


def _mask_tensor(m, dev):
    """DataArray / ndarray / tensor -> (uint8 device tensor, has_channel).  Masks arrive validated (_validate_mask: no NaN,
    0 / 1 only); the NaN -> False rule of api.py:431-435 is kept for callers of this helper that skip the validation."""
    if isinstance(m, DataArray):  # align by dimension NAME like the reference (its own maskers return
        # (channel, range_sample, ping_time)): transpose to ([channel,] ping_time, range_sample)
        unknown = [d for d in m.dims if d not in ("channel", "ping_time", "range_sample")]
        if unknown:
            raise ValueError(f"mask has dimensions {unknown} that the source variable does not have")
        want = tuple(d for d in ("channel", "ping_time", "range_sample") if d in m.dims)
        if tuple(m.dims) != want and set(want) >= {"ping_time", "range_sample"}:
            m = m.transpose(*want)
    data = m.data if isinstance(m, DataArray) else m
    dims = tuple(m.dims) if isinstance(m, DataArray) else None
    if isinstance(data, torch.Tensor):
        t = data
        if t.dtype != torch.uint8:
            t = (torch.nan_to_num(t.float(), nan=0.0) != 0).to(torch.uint8)
        t = t.to(dev)
    else:
        a = np.asarray(data)
        if a.dtype.kind == "f":
            a = np.where(np.isnan(a), 0.0, a)
        t = torch.from_numpy(np.ascontiguousarray(a.astype(bool).astype(np.uint8))).to(dev)
    has_channel = (dims is not None and "channel" in dims) or (dims is None and t.ndim == 3)
    if dims is not None and has_channel and dims[0] != "channel":
        raise ValueError("masks with a channel dimension must have it first")
    return t.contiguous(), has_channel


_ALLOWED_MASK_DIMS = [
    {"ping_time", "range_sample"}, {"ping_time", "depth"}, {"ping_time", "echo_range"},
    {"channel", "ping_time", "range_sample"}, {"channel", "ping_time", "depth"}, {"channel", "ping_time", "echo_range"},
]


def _validate_mask(m, target_dims):
    """The checks of mask/api.py:131-160 (_validate_and_collect_mask_input) and :41-71 (_check_mask_dim_alignment) for one
    mask, with the reference's exception types and messages: allowed dimension sets, no NaN, only 0 / 1 / True / False,
    dimensions equal to those of the source variable when 'channel' is set aside.  uint8 / bool DEVICE tensors (the masks
    this package produces) hold 0 / 1 by construction and are not read back."""
    if isinstance(m, DataArray):
        if set(m.dims) not in _ALLOWED_MASK_DIMS:
            raise ValueError(
                "Masks must have one of the following dimensions: "
                "{'ping_time', 'range_sample'}, "
                "{'ping_time', 'depth'}, "
                "{'ping_time', 'echo_range'}, "
                "{'channel', 'ping_time', 'range_sample'}, "
                "{'channel', 'ping_time', 'depth'}"
                "{'channel', 'ping_time', 'echo_range'}"
            )
    data = m.data if isinstance(m, DataArray) else m
    if isinstance(data, torch.Tensor):
        if data.dtype not in (torch.uint8, torch.bool):
            if data.is_floating_point() and bool(torch.isnan(data).any()):
                raise TypeError("Mask cannot contain NaN")
            if not bool(((data == 0) | (data == 1)).all()):
                raise TypeError("Mask must be boolean (True/False or 1/0)")
    else:
        a = np.asarray(data)
        if a.dtype.kind == "f" and np.any(np.isnan(a)):
            raise TypeError("Mask cannot contain NaN")
        if a.dtype.kind != "b" and not np.all(np.isin(np.unique(a), [0, 1, True, False])):
            raise TypeError("Mask must be boolean (True/False or 1/0)")
    if isinstance(m, DataArray):
        mask_dims, want = set(m.dims) - {"channel"}, set(target_dims) - {"channel"}
        if "channel" in m.dims and "channel" not in target_dims:
            raise ValueError("'channel' is a dimension in mask but not a dimension in source.")
        if mask_dims != want:
            raise ValueError(
                f"The dimensions of mask: ({mask_dims}) do not match "
                f"the dimensions of source ({want}) "
                "when not considering 'channel'."
            )


def apply_mask(source_ds, mask: Union[DataArray, List[DataArray]], var_name: str = "Sv",
               fill_value: Union[int, float] = np.nan, storage_options_ds: dict = {}, storage_options_mask=None) -> Dataset:
    """``source_ds[var_name]`` where the mask(s) hold, ``fill_value`` elsewhere (echopype.mask.apply_mask).  Several
    masks are combined with logical AND.  Returns a shallow copy of the Dataset with the masked variable (device
    resident) and the reference's provenance attributes."""
    source_ds = as_dataset(source_ds)
    if var_name not in source_ds.variables:
        raise ValueError("The Dataset source_ds does not contain the variable var_name!")
    if not isinstance(fill_value, (int, float, DataArray)):
        raise TypeError("The input fill_value must be of type int, float, or xr.DataArray!")
    masks = list(mask) if isinstance(mask, (list, tuple)) else [mask]
    if not masks:
        raise ValueError("mask must contain at least one mask")
    src = source_ds[var_name]
    for m in masks:  # the reference validates every mask before it touches the data (mask/api.py:380-381)
        _validate_mask(m, tuple(src.dims))
    if tuple(src.dims) != ("channel", "ping_time", "range_sample"):
        raise ValueError(f"source_ds[{var_name}] must have dims ('channel', 'ping_time', 'range_sample')")
    dev = require_cuda()
    C, P, R = src.shape
    if isinstance(fill_value, DataArray):  # mask/api.py:233-246: squeeze a length-1 channel, shape of one channel of var_name
        fdata = fill_value.data
        fshape = tuple(int(n) for n in fdata.shape if int(n) != 1) if fdata.ndim > 2 else tuple(fdata.shape)
        if fshape != (P, R):
            raise ValueError(f"If fill_value is an array it must be of the same shape as {var_name}!")
        fill = to_device_f32(fdata, dev).reshape(P, R).contiguous()
    else:
        fill = float(fill_value)
    out = to_device_f32(src.data, dev)
    for m in masks:
        mt, has_c = _mask_tensor(m, dev)
        if tuple(mt.shape[-2:]) != (P, R):
            raise ValueError(
                f"The final constructed mask is not of the same shape as source_ds[{var_name}] "
                "along the ping_time, and range_sample dimensions!"
            )
        if has_c and mt.shape[0] != C:
            raise ValueError(
                f"If both the final constructed mask and source_ds[{var_name}] "
                "have the channel dimension, that dimension should match between the two."
            )
        out = kernels.apply_mask(out, mt, has_c, fill, C, P, R)
    lo, hi, _ = kernels.minmax(out)
    attrs = dict(src.attrs)
    attrs.update({
        "long_name": "Volume backscattering strength, masked (Sv re 1 m-1)",
        "actual_range": [round(float(lo), 2), round(float(hi), 2)],
        "history": f"{_history()}. Created masked Sv dataarray.",
    })
    m0 = masks[0]
    if isinstance(m0, DataArray) and len(m0.attrs) > 0:
        ma = dict(m0.attrs)
        if "history" in ma:
            attrs["history"] += f"\n{ma.pop('history')}"
        attrs.update(ma)
    output_ds = source_ds.copy()
    # a NaN fill keeps "NaN wherever the range variable is NaN"; a finite fill (or a fill array) does not, and the
    # index-space binning of compute_MVBS must not be used on the result (Dataset.__setitem__ drops the range laws)
    nan_fill = isinstance(fill, float) and fill != fill
    trusted = isinstance(src.law, dict) and src.law.get("kind") == "derived"
    output_ds[var_name] = DataArray(out, src.dims, name=var_name, attrs=attrs, law={"kind": "derived"} if (nan_fill and trusted) else None)
    prov = echopype_prov_attrs(process_type="mask")
    prov["mask_function"] = "mask.apply_mask"
    output_ds.attrs.update(prov)
    output_ds = insert_input_processing_level(output_ds, input_ds=source_ds)
    return output_ds
