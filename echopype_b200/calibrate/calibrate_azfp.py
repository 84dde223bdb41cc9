"""AZFP calibrator: echopype/calibrate/calibrate_azfp.py:10-117 with the array math on the device."""

import numpy as np

from .. import kernels
from ..dataset import DataArray, Dataset, EchoData
from ..device import require_cuda, to_device_f32
from .cal_params import get_cal_params_AZFP
from .calibrate_ek import DIMENSION_ORDER, CalibrateBase, _cp
from .env_params import get_env_params_AZFP


class CalibrateAZFP(CalibrateBase):
    def __init__(self, echodata: EchoData, env_params=None, cal_params=None, ecs_file=None, **kwargs):
        if ecs_file is not None:
            raise ValueError("Using ECS file for calibration is not currently supported for AZFP!")
        super().__init__(echodata, env_params, cal_params, None)
        self.sonar_type = "AZFP"
        self.env_params = get_env_params_AZFP(echodata=self.echodata, user_dict=self.env_params)
        self.cal_params = get_cal_params_AZFP(
            beam=self.echodata["Sonar/Beam_group1"], vend=self.echodata["Vendor_specific"], user_dict=self.cal_params
        )

    def _power_row_builder(self, cal_type):
        """Host assembly of the row-setup launch (parameters uploaded once; .build() launches epb_rows_azfp)."""
        if cal_type not in ("Sv", "TS"):
            raise ValueError("cal_type not recognized!")
        if "sound_speed" not in self.env_params:
            raise RuntimeError(
                "sounds_speed not included in env_params, "
                "use echopype.calibrate.env_params.get_env_params_AZFP() to compute env_params "
                "by supplying temperature, salinity, and pressure."
            )
        beam = self.echodata["Sonar/Beam_group1"]
        vend = self.echodata["Vendor_specific"]
        bs = beam["backscatter_r"]
        if tuple(bs.dims) != DIMENSION_ORDER:
            raise ValueError(f"backscatter_r must have dims {DIMENSION_ORDER}, got {bs.dims}")
        C, P, R = bs.shape
        chan = np.asarray(beam["channel"].values)
        require_cuda()
        cp = self.cal_params
        prm = {
            "sound_speed": _cp(self.env_params["sound_speed"], chan),
            "sound_absorption": _cp(self.env_params["sound_absorption"], chan),
            "transmit_duration_nominal": _cp(beam["transmit_duration_nominal"], chan),
            "N": _cp(vend["number_of_samples_per_average_bin"], chan),
            "f_dig": _cp(vend["digitization_rate"], chan),
            "L": _cp(vend["lock_out_index"], chan),
            "EL": _cp(cp["EL"], chan), "DS": _cp(cp["DS"], chan), "TVR": _cp(cp["TVR"], chan), "VTX0": _cp(cp["VTX0"], chan),
            "equivalent_beam_angle": _cp(cp["equivalent_beam_angle"], chan), "Sv_offset": _cp(cp["Sv_offset"], chan),
        }
        for k in ("N", "f_dig", "L", "EL", "DS", "TVR", "VTX0", "equivalent_beam_angle", "Sv_offset"):
            v = np.asarray(prm[k], dtype=np.float64)
            if v.ndim == 2:  # (channel, ping_time) duplicates of a per-channel constant
                v = v[:, 0]
            prm[k] = np.ascontiguousarray(np.broadcast_to(v, (C,)))
        rb = kernels.azfp_row_builder(C, P, R, cal_type, prm)
        rb.tau_effective = None
        return rb

    def _power_rows(self, cal_type):
        """Row records + device-resident counts: (rows, x, C, P, R, None)."""
        rb = self._power_row_builder(cal_type)
        C, P, R = rb.shape
        rows = rb.build()
        x = to_device_f32(self.echodata["Sonar/Beam_group1"]["backscatter_r"].data)
        self.rows = rows
        return rows, x, C, P, R, None

    def _cal_power_samples(self, cal_type, **kwargs):
        """Device version of CalibrateAZFP._cal_power_samples (calibrate_azfp.py:49-111) including
        compute_range_AZFP (calibrate/range.py:11-95)."""
        rows, x, C, P, R, _ = self._power_rows(cal_type)
        beam = self.echodata["Sonar/Beam_group1"]
        out, rng, mm = kernels.sv_power(x, rows, C, P, R, want_range=True, want_minmax=True)
        ds = Dataset(coords={d: beam[d].values for d in DIMENSION_ORDER})
        ds[cal_type] = DataArray(out, DIMENSION_ORDER, name=cal_type, law={"kind": "derived"})
        er = DataArray(rng, DIMENSION_ORDER, name="echo_range")
        er.law = {"rows": rows, "kind": "echo_range", "minmax": mm}
        ds["echo_range"] = er
        ds["frequency_nominal"] = beam["frequency_nominal"]
        return self._add_params_to_output(ds)

    def _check_echodata_backscatter_size(self):
        self.ed_beam_group = "Sonar/Beam_group1"
        super()._check_echodata_backscatter_size()

    def compute_Sv(self, **kwargs):
        return self._cal_power_samples(cal_type="Sv")

    def compute_TS(self, **kwargs):
        return self._cal_power_samples(cal_type="TS")
