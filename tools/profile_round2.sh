#!/bin/bash
# Round-2 evidence run (on the GPU box, under gpurun): every kernel alone (CUDA events), every public call on a cfg2 volume,
# ncu --set full captures of the kernels this round changed, and the ncu launch list of the default bench command.
# Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
python tools/bench_kernels.py --which copy,sv,svr,svrm,noise,bins,pipe,pipe16,masks,k2,pulse --bb-pings 8000 > gpurun_out/kbench_r2.log 2>&1
python tools/bench_kernels.py --which pipe --R 8192 --P 50000 > gpurun_out/kbench_r2_wide.log 2>&1
python tools/bench_api.py 2>/dev/null | grep '^{' > gpurun_out/api_r2.log
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:pipeline_fast_kernel -c 1 -o gpurun_out/r2_fast python tools/bench_kernels.py --which pipe --iters 1 > /dev/null 2>&1
timeout 300 $NCU -k regex:bin_reduce_staged -c 1 -o gpurun_out/r2_bins python tools/bench_kernels.py --which bins --iters 1 > /dev/null 2>&1
timeout 300 $NCU -k regex:transient_strip -c 1 -o gpurun_out/r2_strip python tools/bench_kernels.py --which masks --iters 1 > /dev/null 2>&1
timeout 300 $NCU -k regex:impulse_fused -c 1 -o gpurun_out/r2_impulse python tools/bench_kernels.py --which masks --iters 1 > /dev/null 2>&1
timeout 300 $NCU -k regex:pulse_fft_kernel -c 2 -o gpurun_out/r2_fft python tools/bench_kernels.py --which pulse --bb-pings 4000 --iters 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_r2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_under_ncu_r2.log 2>&1
ls -la gpurun_out | tail -12
