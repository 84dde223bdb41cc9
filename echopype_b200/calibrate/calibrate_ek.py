"""EK60 / EK80 calibrators: the reference's CALIBRATOR classes (calibrate/calibrate_ek.py:56-710,
calibrate_base.py:10-128) with the array arithmetic moved to the device.

Host side (this file): select the beam group, assemble env / cal parameters as small float64
(channel[, ping_time]) arrays, synthesize the transmit replica (EK80).  Device side: one row-setup
launch (epb_rows_*) folds those parameters into per-row records, then ONE fused sample kernel
(epb_sv_power / epb_sv_complex / epb_pulse_compress_sv) produces Sv|TS and echo_range.
"""

import numpy as np

from .. import kernels
from ..dataset import DataArray, Dataset, EchoData
from ..device import require_cuda, to_device_f32
from ..utils.log import _init_logger
from .cal_params import get_cal_params_EK
from .ek80_complex import get_filter_coeff, get_tau_effective, get_transmit_signal
from .env_params import get_env_params_EK

logger = _init_logger(__name__)
DIMENSION_ORDER = ("channel", "ping_time", "range_sample")


def retrieve_correct_beam_group(echodata: EchoData, waveform_mode: str, encode_mode: str) -> str:
    """echodata/simrad.py:146-179 (+ :44-140)."""
    if echodata.sonar_model in ["EK60", "ES70"]:
        if waveform_mode != "CW":
            raise RuntimeError("Incorrect waveform_mode input provided!")
        if encode_mode != "power":
            raise RuntimeError("Incorrect encode_mode input provided!")
        if "backscatter_i" in echodata["Sonar/Beam_group1"].variables:
            raise RuntimeError(
                "Provided echodata object does not correspond to an EK60-like sensor, but is labeled as data from an EK60-like sensor!"
            )
        return "Sonar/Beam_group1"
    elif echodata.sonar_model in ["EK80", "ES80", "EA640"]:
        if "waveform_encode_descr" not in echodata["Sonar"]:
            raise ValueError("Echodata missing `waveform_encode_descr`. Reconvert using the latest Echopype version.")
        descr = np.asarray(echodata["Sonar"]["waveform_encode_descr"].values)
        match_str = "power" if encode_mode == "power" else ("complex_CW" if waveform_mode == "CW" else "complex_FM")
        idx = np.flatnonzero(descr == match_str)
        if idx.size == 0:
            raise RuntimeError(
                f"No beam group with the specified encode_mode {encode_mode} and waveform_mode {waveform_mode} "
                "found in the provided echodata!"
            )
        return f"Sonar/Beam_group{int(idx[0]) + 1}"
    raise RuntimeError("EchoData was produced by a non-Simrad or unknown Simrad echo sounder!")


def _cp(da, channel, ping_time=None):
    """DataArray | scalar -> float64 numpy broadcastable to (C,P): scalar, (C,), or (C,P)."""
    if da is None:
        return None
    if not isinstance(da, DataArray):
        return np.asarray(da, dtype=np.float64)
    v = np.asarray(da.values, dtype=np.float64)
    dims = tuple(d for d in da.dims)
    extra = [d for d in dims if d not in ("channel", "ping_time")]
    for d in extra:  # e.g. a leftover length-1 'time1' or 'beam' dimension
        ax = dims.index(d)
        v = np.take(v, 0, axis=ax)
        dims = dims[:ax] + dims[ax + 1 :]
    if dims == ("ping_time", "channel"):
        v, dims = v.T, ("channel", "ping_time")
    if dims == ("ping_time",):  # explicit (1, P): a bare (P,) vector would be taken for per-channel values when P == C
        return v.reshape(1, -1)
    if dims == ("channel",) and "channel" in da.coords:
        have = np.asarray(da.coords["channel"])
        want = np.asarray(channel)
        if have.shape == want.shape and not np.array_equal(have, want) and set(have.tolist()) == set(want.tolist()):
            v = v[[int(np.flatnonzero(have == c)[0]) for c in want]]
    if dims == ("channel", "ping_time") and v.shape[1] > 1 and bool(np.all(v == v[:, :1])):
        v = np.ascontiguousarray(v[:, :1])  # constant along ping_time: (C,1) broadcasts the same way and ships C values
    return v


class CalibrateBase:
    """calibrate_base.py:10-128."""

    def __init__(self, echodata: EchoData, env_params=None, cal_params=None, ecs_file=None):
        self.echodata = echodata
        self.sonar_type = None
        self.ecs_file = ecs_file
        if ecs_file is not None:
            raise NotImplementedError(
                "ECS calibration files are outside the accelerated path (SURVEY.md section 2 #8); "
                "pass the values through env_params / cal_params instead."
            )
        if env_params is None:
            self.env_params = {}
        elif isinstance(env_params, dict):
            self.env_params = env_params
        else:
            raise ValueError("'env_params' has to be None or a dict")
        if cal_params is None:
            self.cal_params = {}
        elif isinstance(cal_params, dict):
            self.cal_params = cal_params
        else:
            raise ValueError("'cal_params' has to be None or a dict")
        self.range_meter = None

    def _add_params_to_output(self, ds_out):
        for key, val in {**self.env_params, **self.cal_params}.items():
            if val is not None:
                ds_out[key] = val
        return ds_out

    def _check_echodata_backscatter_size(self):
        """calibrate_base.py:95-128: advisory warning above 2 GiB of backscatter."""
        beam = self.echodata[getattr(self, "ed_beam_group", None) or "Sonar/Beam_group1"]
        total = beam["backscatter_r"].nbytes
        if "backscatter_i" in beam and getattr(self, "encode_mode", "power") == "complex":
            total += beam["backscatter_i"].nbytes
        if total / (1024**3) > 2.0:
            logger.warning(
                "The Echodata backscatter variables are large and can cause memory issues. "
                "Consider modifying the workflow that uses compute_Sv as below: "
                "Prior to `compute_Sv` run `echodata.chunk(CHUNK_DICTIONARY) "
                "and after `compute_Sv` run `ds_Sv.to_zarr(ZARR_STORE, compute=True)`. "
                "This will ensure that the computation is lazily evaluated, "
                "with the results stored directly in a Zarr store on disk, rather then in memory."
            )


class CalibrateEK(CalibrateBase):
    def __init__(self, echodata, env_params, cal_params, ecs_file, **kwargs):
        super().__init__(echodata, env_params, cal_params, ecs_file)
        self.ed_beam_group = None
        self.beam = None
        self.vend = None

    # ---- shared helpers ---------------------------------------------------------------------------
    def _shape(self):
        bs = self.beam["backscatter_r"]
        order = [bs.dims.index(d) for d in DIMENSION_ORDER] + ([bs.dims.index("beam")] if "beam" in bs.dims else [])
        if order != list(range(bs.ndim)):
            raise ValueError(f"backscatter_r must have dims {DIMENSION_ORDER}[+beam], got {bs.dims}")
        return bs.shape

    def _is_gpt(self):
        if self.sonar_type == "EK60":
            return np.ones(self.beam.sizes["channel"], dtype=bool)
        if "transceiver_type" not in self.vend:
            return np.zeros(self.beam.sizes["channel"], dtype=bool)
        return np.asarray(self.vend["transceiver_type"].values).astype(str) == "GPT"

    def _sonar_code(self):
        m = self.echodata.sonar_model
        if m in ("EK60", "ES70"):
            return kernels.SONAR_EX60
        if m in ("EK80", "ES80", "EA640"):
            return kernels.SONAR_EX80
        raise ValueError("The specified sonar_model is not supported!")

    def _transmit_signal(self, tx_coeff, fs):
        """One replica build per calibration object: tau_effective and the matched filter use the same signal."""
        memo = getattr(self, "_tx_memo", None)
        if memo is None:
            memo = get_transmit_signal(self.beam, tx_coeff, self.waveform_mode, fs, getattr(self, "drop_last_hanning_zero", False))
            self._tx_memo = memo
            self._tx = memo[0]
        return memo

    def _tau_effective(self, what):
        """calibrate_ek.py:112-151 / :586-607: transmit-signal tau_eff with GPT channels (all of EK60)
        overwritten by transmit_duration_nominal of ping 0."""
        beam, vend = self.beam, self.vend
        chan = np.asarray(beam["channel"].values)
        tau_nom0 = np.asarray(beam["transmit_duration_nominal"].transpose("channel", "ping_time").values, dtype=np.float64)[:, 0]
        attrs = {
            "long_name": "Effective pulse length", "units": "s",
            "description": "Effective pulse length used for Sv. GPT uses transmit_duration_nominal.",
        }
        if self.sonar_type == "EK60":
            # every EK60 channel is GPT: the reference's replica attempt always fails (no filter
            # coefficients) and is overwritten by the nominal duration of ping 0 (calibrate_ek.py:134-151)
            return DataArray(tau_nom0.copy(), dims=["channel"], coords={"channel": chan}, attrs=attrs)
        try:
            tx_coeff = get_filter_coeff(vend)
            fs = self.cal_params["receiver_sampling_frequency"]
            tx, tx_time = self._transmit_signal(tx_coeff, fs)
            tau_eff = get_tau_effective(
                tx, {k: 1 / np.diff(v[:2]) for k, v in tx_time.items()}, self.waveform_mode, chan, beam["ping_time"]
            ).values.astype(np.float64)
            self._tx = tx
        except Exception as e:  # noqa: BLE001 - same broad fallback as the reference
            logger.warning(
                f"Could not compute tau_effective from transmit signal in {what} encoding mode; "
                "falling back to transmit_duration_nominal. Error: %s",
                repr(e),
            )
            tau_eff = tau_nom0.copy()
        gpt = self._is_gpt()
        tau_eff = np.where(gpt, tau_nom0, tau_eff)
        return DataArray(tau_eff, dims=["channel"], coords={"channel": chan}, attrs=attrs)

    def _finish(self, cal_type, out_t, rng_t, tau_effective, rows, minmax=None):
        beam = self.beam
        coords = {d: beam[d].values for d in DIMENSION_ORDER}
        ds = Dataset(coords=coords)
        # law {"kind": "derived"}: NaN wherever echo_range is NaN - what index-space binning relies on (dataset.py)
        ds[cal_type] = DataArray(out_t, DIMENSION_ORDER, name=cal_type, law={"kind": "derived"})
        er = DataArray(rng_t, DIMENSION_ORDER, name="echo_range")
        er.law = {"rows": rows, "kind": "echo_range", "minmax": minmax}
        ds["echo_range"] = er
        if cal_type == "Sv" and tau_effective is not None:
            ds["tau_effective"] = tau_effective
        ds["frequency_nominal"] = beam["frequency_nominal"]
        return self._add_params_to_output(ds)

    def _power_params(self, cal_type):
        beam = self.beam
        chan = np.asarray(beam["channel"].values)
        prm = {
            "sample_interval": _cp(beam["sample_interval"], chan),
            "sound_speed": _cp(self.env_params["sound_speed"], chan),
            "sound_absorption": _cp(self.env_params["sound_absorption"], chan),
            "transmit_duration_nominal": _cp(beam["transmit_duration_nominal"], chan),
            "transmit_power": _cp(beam["transmit_power"], chan),
            "gain_correction": _cp(self.cal_params["gain_correction"], chan),
            "frequency_nominal": _cp(beam["frequency_nominal"], chan),
        }
        tau_effective = None
        if cal_type == "Sv":
            tau_effective = self._tau_effective("power")
            prm["sa_correction"] = _cp(self.cal_params["sa_correction"], chan)
            prm["equivalent_beam_angle"] = _cp(self.cal_params["equivalent_beam_angle"], chan)
            prm["tau_effective"] = tau_effective.values
        return prm, tau_effective

    def _power_row_builder(self, cal_type: str):
        """Host assembly of the row-setup launch (parameters uploaded once; .build() launches epb_rows_ek_power)."""
        if cal_type not in ("Sv", "TS"):
            raise ValueError("cal_type must be 'Sv' or 'TS'")
        C, P, R = self._shape()[:3]
        require_cuda()
        prm, tau_effective = self._power_params(cal_type)
        rb = kernels.ek_power_row_builder(C, P, R, self._sonar_code(), cal_type, prm, self._is_gpt())
        rb.tau_effective = tau_effective
        return rb

    def _power_rows(self, cal_type: str):
        """Row records + device-resident raw samples for the power-sample path: (rows, x, C, P, R, tau_effective)."""
        rb = self._power_row_builder(cal_type)
        C, P, R = rb.shape
        tau_effective = rb.tau_effective
        rows = rb.build()
        x = kernels.power_to_device_f32(self.beam["backscatter_r"].data, keep_counts=True)  # float32 dB, or int16 raw counts
        self.rows = rows
        return rows, x, C, P, R, tau_effective

    def _cal_power_samples(self, cal_type: str) -> Dataset:
        """Device version of CalibrateEK._cal_power_samples (calibrate_ek.py:79-206)."""
        rows, x, C, P, R, tau_effective = self._power_rows(cal_type)
        out, rng, mm = kernels.sv_power(x, rows, C, P, R, want_range=True, want_minmax=True)
        return self._finish(cal_type, out, rng, tau_effective, rows, mm)


class CalibrateEK60(CalibrateEK):
    def __init__(self, echodata, env_params=None, cal_params=None, ecs_file=None, **kwargs):
        super().__init__(echodata, env_params, cal_params, ecs_file)
        self.sonar_type = "EK60"
        self.waveform_mode = "CW"
        self.encode_mode = "power"
        self.ed_beam_group = retrieve_correct_beam_group(echodata, self.waveform_mode, self.encode_mode)
        self.beam = echodata[self.ed_beam_group]
        self.vend = echodata["Vendor_specific"]
        self.env_params = get_env_params_EK(sonar_type=self.sonar_type, beam=self.beam, env=echodata["Environment"], user_dict=self.env_params)
        self.cal_params = get_cal_params_EK(
            waveform_mode=self.waveform_mode, freq_center=self.beam["frequency_nominal"], beam=self.beam, vend=self.vend,
            user_dict=self.cal_params, sonar_type=self.sonar_type,
        )

    def compute_Sv(self, **kwargs):
        return self._cal_power_samples(cal_type="Sv")

    def compute_TS(self, **kwargs):
        return self._cal_power_samples(cal_type="TS")


class CalibrateEK80(CalibrateEK):
    def __init__(self, echodata, env_params=None, cal_params=None, waveform_mode=None, encode_mode=None, ecs_file=None,
                 slice_dict=None, drop_last_hanning_zero=False, **kwargs):
        super().__init__(echodata, env_params, cal_params, ecs_file)
        self.sonar_type = "EK80"
        self.waveform_mode = waveform_mode
        self.encode_mode = encode_mode
        self.drop_last_hanning_zero = drop_last_hanning_zero
        self.slice_dict = slice_dict or {}
        self.ed_beam_group = retrieve_correct_beam_group(echodata, waveform_mode, encode_mode)
        self.beam = echodata[self.ed_beam_group]
        vend = echodata["Vendor_specific"]
        # align Vendor_specific rows with the beam group's channels (calibrate_ek.py:333)
        bch = np.asarray(self.beam["channel"].values)
        vch = np.asarray(vend["channel"].values)
        if not np.array_equal(bch, vch):
            vend = vend.isel(channel=[int(np.flatnonzero(vch == c)[0]) for c in bch])
        self.vend = vend
        if self.waveform_mode == "BB":
            self.freq_center = (self.beam["transmit_frequency_start"] + self.beam["transmit_frequency_stop"]) / 2
        else:
            self.freq_center = self.beam["frequency_nominal"]
        self.env_params = get_env_params_EK(
            sonar_type=self.sonar_type, beam=self.beam, env=echodata["Environment"], user_dict=self.env_params, freq=self.freq_center
        )
        self.cal_params = get_cal_params_EK(
            waveform_mode=self.waveform_mode, freq_center=self.freq_center, beam=self.beam, vend=self.vend,
            user_dict=self.cal_params, sonar_type="EK80",
        )

    def _get_B_theta_phi_m(self):
        """calibrate_ek.py:507-530."""
        cp = self.cal_params
        need = ["angle_offset_alongship", "angle_offset_athwartship", "beamwidth_alongship", "beamwidth_athwartship"]
        if any(cp.get(k) is None for k in need):
            return 0.0
        chan = np.asarray(self.beam["channel"].values)
        fa = (np.abs(-_cp(cp[need[0]], chan)) / (_cp(cp[need[2]], chan) / 2)) ** 2
        ft = (np.abs(-_cp(cp[need[1]], chan)) / (_cp(cp[need[3]], chan) / 2)) ** 2
        B = 0.5 * 6.0206 * (fa + ft - 0.18 * fa * ft)
        return np.where(np.isnan(B), 0.0, B)

    def _cal_complex_samples(self, cal_type: str) -> Dataset:
        """Device version of CalibrateEK80._cal_complex_samples (calibrate_ek.py:532-659)."""
        beam = self.beam
        chan = np.asarray(beam["channel"].values)
        C, P, R, B = self._shape()
        require_cuda()
        bb = self.waveform_mode == "BB"
        tx_coeff = get_filter_coeff(self.vend)
        fs = self.cal_params["receiver_sampling_frequency"]
        tx, tx_time = self._transmit_signal(tx_coeff, fs)
        gain = _cp(self.cal_params["gain_correction"], chan)
        if bb:  # transceiver gain compensation, calibrate_ek.py:561-562
            g, Bm = np.asarray(gain, dtype=np.float64), np.asarray(self._get_B_theta_phi_m(), dtype=np.float64)
            if g.ndim == 1 and Bm.ndim == 2:
                g = g[:, None]
            if Bm.ndim == 1 and g.ndim == 2:
                Bm = Bm[:, None]
            gain = g - Bm
        tau_effective = self._tau_effective("complex") if cal_type == "Sv" else None
        prm = {
            "sample_interval": _cp(beam["sample_interval"], chan),
            "sound_speed": _cp(self.env_params["sound_speed"], chan),
            "sound_absorption": _cp(self.env_params["sound_absorption"], chan),
            "transmit_duration_nominal": _cp(beam["transmit_duration_nominal"], chan),
            "transmit_power": _cp(beam["transmit_power"], chan),
            "gain_correction": gain,
            "freq_center": _cp(self.freq_center, chan),
            "impedance_transducer": _cp(self.cal_params["impedance_transducer"], chan),
            "impedance_transceiver": _cp(self.cal_params["impedance_transceiver"], chan),
        }
        if cal_type == "Sv":
            prm["equivalent_beam_angle"] = _cp(self.cal_params["equivalent_beam_angle"], chan)
            prm["tau_effective"] = tau_effective.values
            if not bb:
                prm["sa_correction"] = _cp(self.cal_params["sa_correction"], chan)
        rows = kernels.rows_ek80_complex(C, P, R, cal_type, bb, B, prm, self._is_gpt())
        re = to_device_f32(beam["backscatter_r"].data)
        im = to_device_f32(beam["backscatter_i"].data)
        if bb:
            out, rng, _, mm = kernels.pulse_compress_sv(re, im, [tx[c] for c in chan], rows, C, P, R, B, want_minmax=True)
        else:
            out, rng, mm = kernels.sv_complex(re, im, rows, C, P, R, B, want_minmax=True)
        self.rows = rows
        return self._finish(cal_type, out, rng, tau_effective, rows, mm)

    def _compute_cal(self, cal_type) -> Dataset:
        if self.waveform_mode == "BB" or self.encode_mode == "complex":
            return self._cal_complex_samples(cal_type=cal_type)
        return self._cal_power_samples(cal_type=cal_type)

    def compute_Sv(self):
        return self._compute_cal(cal_type="Sv")

    def compute_TS(self):
        return self._compute_cal(cal_type="TS")
