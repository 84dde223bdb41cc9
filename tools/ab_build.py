"""Build variants of libepb200.so for A/B timing on the GPU box in one gpurun call.

    python tools/ab_build.py NAME=-DFLAG[,-DFLAG2] ...   ->  gpurun_ab/libepb200_NAME.so  (pipeline_fast.cu recompiled)
    EPB200_LIB=gpurun_ab/libepb200_NAME.so python tools/bench_kernels.py --which pipe
"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from echopype_b200.build import CSRC, FLAGS, NVCC, OBJ, build  # noqa: E402

build()
out = os.path.join(ROOT, "gpurun_ab")
os.makedirs(out, exist_ok=True)
for spec in sys.argv[1:]:
    name, _, defs = spec.partition("=")
    defs = [d for d in defs.split(",") if d]
    o = os.path.join(out, f"pipeline_fast_{name}.o")
    subprocess.check_call([NVCC] + FLAGS + defs + ["-I", os.path.join(ROOT, "include"), "-c", os.path.join(CSRC, "pipeline_fast.cu"), "-o", o])
    objs = [x for x in sorted(glob.glob(os.path.join(OBJ, "*.o"))) if not x.endswith("pipeline_fast.o")] + [o]
    lib = os.path.join(out, f"libepb200_{name}.so")
    subprocess.check_call([NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    print("built", lib)
