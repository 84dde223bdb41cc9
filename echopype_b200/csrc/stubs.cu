// Entry points not implemented yet: they fail loudly (never a silent fallback).
#include "epb_common.cuh"
#define EPB_STUB(name, ...)                                         \
  extern "C" int name(__VA_ARGS__) {                                \
    epb_set_error(#name ": not implemented in this build");         \
    return EPB_E_UNSUPPORTED;                                       \
  }
EPB_STUB(epb_pulse_compress_sv, const float*, const float*, const float*, const int*, const double*, const epb_row*,
         float*, float*, float*, float*, epb_i64, epb_i64, epb_i64, int, void*)
EPB_STUB(epb_noise_estimate, const float*, const float*, epb_cp, float*, epb_i64, epb_i64, epb_i64, int, int, float, void*)
EPB_STUB(epb_noise_apply, const float*, const float*, epb_cp, const float*, float*, float*, float*, epb_i64, epb_i64,
         epb_i64, int, float, void*)
EPB_STUB(epb_bin_reduce, const float*, const void*, int, const int*, const double*, int, int, int, double*, epb_i64,
         epb_i64, epb_i64, epb_i64, void*)
EPB_STUB(epb_bin_finalize, const double*, float*, double*, epb_i64, int, float, int, void*)
EPB_STUB(epb_coarsen, const float*, const float*, float*, float*, epb_i64, epb_i64, epb_i64, int, int, void*)
EPB_STUB(epb_bin_bounds, const epb_row*, const double*, int, int, const double*, const double*, double, int*, epb_i64,
         epb_i64, epb_i64, void*)
EPB_STUB(epb_pipeline_power_mvbs, const float*, const epb_row*, const int*, const int*, int, double*, float*, float*,
         float*, float*, float*, epb_i64, epb_i64, epb_i64, epb_i64, int, int, float, float, void*)
extern "C" epb_i64 epb_pipeline_smem_bytes(epb_i64 R, int ping_num) { return R * 4 * (ping_num > 0 ? ping_num : 1); }
