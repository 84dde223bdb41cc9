// Entry points not implemented yet: they fail loudly (never a silent fallback).
#include "epb_common.cuh"
#define EPB_STUB(name, ...)                                         \
  extern "C" int name(__VA_ARGS__) {                                \
    epb_set_error(#name ": not implemented in this build");         \
    return EPB_E_UNSUPPORTED;                                       \
  }
EPB_STUB(epb_pulse_compress_sv, const float*, const float*, const float*, const int*, const double*, const epb_row*,
         float*, float*, float*, float*, epb_i64, epb_i64, epb_i64, int, void*)
