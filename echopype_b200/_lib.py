"""ctypes binding of libepb200.so (include/epb200.h).  The product path has NO fallback: if the
library is missing or a call fails, an exception is raised."""

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_longlong, c_ubyte, c_uint, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# EPB200_LIB: developer knob for A/B builds of the same library (tools/ab_build.py); never a different backend
LIB_PATH = os.environ.get("EPB200_LIB") or os.path.join(_HERE, "libepb200.so")


class EpbError(RuntimeError):
    pass


class epb_cp(Structure):
    _fields_ = [("ptr", c_void_p), ("sc", c_longlong), ("sp", c_longlong)]


class epb_row(Structure):
    _fields_ = (
        [(n, c_double) for n in ("p0", "p1", "p2", "p3", "p4", "off1", "off2", "r0", "a", "two_alpha", "K", "fscale", "foff", "slog")]
        + [("n_start", c_int), ("law", c_int), ("azfp_N", c_int), ("reserved", c_int)]
        + [(n, c_float) for n in ("a_h", "a_l", "r0_h", "r0_l", "bp_h", "bp_l", "two_alpha_f", "slog2", "fscale_f", "foffK",
                                  "c0", "c1", "c2", "spow")]
        + [("range_last", c_double)]
    )


ROW_BYTES = ctypes.sizeof(epb_row)
assert ROW_BYTES == 192

i64, vp = c_longlong, c_void_p

# name -> (restype, argtypes); mirrors include/epb200.h one to one
SIGNATURES = {
    "epb_last_error": (c_char_p, []),
    "epb_version": (c_int, []),
    "epb_rows_ek_power": (c_int, [vp, i64, i64, i64, c_int, c_int] + [epb_cp] * 10 + [vp, vp]),
    "epb_rows_azfp": (c_int, [vp, i64, i64, i64, c_int, epb_cp, epb_cp, epb_cp] + [vp] * 9 + [vp]),
    "epb_rows_ek80_complex": (c_int, [vp, i64, i64, i64, c_int, c_int, c_int] + [epb_cp] * 12 + [vp, vp]),
    "epb_sv_power": (c_int, [vp, vp, vp, vp, vp, i64, i64, i64, vp]),
    "epb_sv_power_i16": (c_int, [vp, vp, vp, vp, vp, i64, i64, i64, vp]),
    "epb_sv_complex": (c_int, [vp, vp, vp, vp, vp, vp, i64, i64, i64, c_int, vp]),
    "epb_pulse_compress_sv": (c_int, [vp, vp, vp, POINTER(c_int), vp, vp, vp, vp, vp, vp, i64, i64, i64, c_int, vp]),
    "epb_pulse_fft_workspace_bytes": (i64, [i64]),
    "epb_pulse_fft_max_taps": (c_int, []),
    "epb_pulse_compress_sv_fft": (c_int, [vp, vp, vp, POINTER(c_int), vp, vp, vp, vp, vp, vp, i64, i64, i64, c_int, vp, i64, vp]),
    "epb_noise_estimate": (c_int, [vp, vp, epb_cp, vp, i64, i64, i64, c_int, c_int, c_float, vp]),
    "epb_noise_apply": (c_int, [vp, vp, epb_cp, vp, vp, vp, vp, i64, i64, i64, c_int, c_float, vp]),
    "epb_bin_reduce": (c_int, [vp, vp, c_int, vp, vp, c_int, c_int, c_int, vp, i64, i64, i64, i64, vp]),
    "epb_bin_reduce_law": (c_int, [vp, vp, vp, vp, vp, vp, c_int, c_int, vp, i64, i64, i64, i64, vp, i64, vp]),
    "epb_bin_finalize": (c_int, [vp, vp, vp, i64, c_int, c_float, c_int, vp]),
    "epb_coarsen": (c_int, [vp, vp, vp, vp, i64, i64, i64, c_int, c_int, vp]),
    "epb_pipeline_power_mvbs": (
        c_int,
        [vp, vp, vp, vp, c_int, c_int, vp, vp, vp, vp, vp, vp, i64, i64, i64, i64, c_int, c_int, c_float, c_float, vp, vp, i64, vp],
    ),
    "epb_pipeline_power_mvbs_i16": (
        c_int,
        [vp, vp, vp, vp, vp, c_int, c_int, vp, vp, i64, i64, i64, i64, c_int, c_int, c_float, c_float, vp, vp, i64, vp],
    ),
    "epb_ingest_power_i16": (c_int, [vp, vp, i64, vp]),
    "epb_synth_fill_i16": (c_int, [vp, i64, i64, i64, c_ulonglong, i64, c_uint, vp]),
    "epb_pipeline_workspace_bytes": (i64, [i64, i64, c_int]),
    "epb_pipeline_workspace_bytes_r": (i64, [i64, i64, i64, c_int]),
    "epb_set_grid_reserve": (c_int, [c_int]),
    "epb_pipeline_smem_bytes": (i64, [i64, c_int, c_int, c_int, c_int]),
    "epb_add_depth": (c_int, [vp, epb_cp, epb_cp, vp, i64, i64, i64, vp]),
    "epb_freq_diff_mask": (c_int, [vp, c_int, c_int, c_int, c_float, vp, i64, i64, i64, vp]),
    "epb_apply_mask": (c_int, [vp, vp, c_int, c_float, vp, i64, i64, i64, vp]),
    "epb_apply_mask_fill_array": (c_int, [vp, vp, c_int, vp, vp, i64, i64, i64, vp]),
    "epb_straddle_pack": (c_int, [vp, i64, i64, i64, vp, c_int, vp, c_int, c_int, c_int, vp]),
    "epb_straddle_unpack": (c_int, [vp, vp, c_int, c_int, i64, i64, i64, c_int, vp, vp, vp]),
    "epb_range_diff_mean": (c_int, [vp, vp, vp, i64, i64, i64, vp]),
    "epb_first_not_le": (c_int, [vp, i64, c_float, vp, vp]),
    "epb_impulse_noise_mask": (c_int, [vp, vp, vp, vp, i64, i64, i64, c_int, c_int, c_float, vp]),
    "epb_transient_noise_mask_depth": (c_int, [vp, vp, vp, vp, vp, vp, i64, i64, i64, c_double, c_double, c_double, c_double, c_int, c_float, vp]),
    "epb_depth_rows_uniform": (c_int, [vp, vp, vp, vp, i64, i64, i64, vp]),
    "epb_transient_noise_mask_depth_uniform": (c_int, [vp, vp, vp, vp, vp, vp, i64, i64, i64, c_double, c_double, c_double, c_double, c_int, c_float, i64, vp]),
    "epb_impulse_noise_mask_depth": (c_int, [vp, vp, vp, c_int, vp, vp, vp, vp, i64, i64, i64, c_int, c_float, vp, vp]),
    "epb_transient_noise_mask": (c_int, [vp, vp, vp, vp, vp, i64, i64, i64, c_int, c_int, c_int, c_float, vp]),
    "epb_transient_noise_mask_median": (c_int, [vp, vp, vp, vp, i64, i64, i64, c_int, c_int, c_int, c_float, vp]),
    "epb_transient_noise_mask_depth_median": (c_int, [vp, vp, vp, vp, i64, i64, i64, c_double, c_double, c_double, c_double, c_int, c_float, vp]),
    "epb_attenuated_signal_mask": (c_int, [vp, vp, vp, vp, i64, i64, i64, c_double, c_double, c_int, c_double, vp]),
    "epb_zero": (c_int, [vp, i64, vp]),
    "epb_minmax_init": (c_int, [vp, vp]),
    "epb_minmax": (c_int, [vp, i64, vp, vp]),
    "epb_range_max": (c_int, [vp, vp, i64, i64, i64, vp, vp]),
    "epb_synth_fill": (c_int, [vp, i64, i64, i64, i64, c_int, c_ulonglong, i64, c_uint, c_float, vp]),
}

_lib = None


def load():
    """Load libepb200.so (once).  Raises EpbError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EpbError(
            f"{LIB_PATH} not found: the CUDA library has not been built. "
            "Run `python -m echopype_b200.build` (needs nvcc). There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if os.environ.get("EPB200_LIB") and not hasattr(lib, name):
            continue  # A/B build of an older revision (developer knob only)
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().epb_last_error()
        raise EpbError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    check(rc, name)
