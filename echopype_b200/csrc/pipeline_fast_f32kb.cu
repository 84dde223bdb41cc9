// Instantiations of the persistent fused kernel (pipeline_fast_impl.cuh) with full-size outputs (keep=): float32 samples,
// ping_num in {5, 6, 7, 8}.
#include "pipeline_fast_impl.cuh"

EPB_DEFINE_FAST_KEEP_LAUNCHER(epb_fast_launch_f32kb, false, 5, 6, 7, 8, false)
