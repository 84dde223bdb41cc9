"""Build variants of libepb200.so for A/B timing on the GPU box in one gpurun call.

    python tools/ab_build.py NAME=-DFLAG[,-DFLAG2] ...   ->  gpurun_ab/libepb200_NAME.so  (the fused fast-path units
                                                              recompiled; AB_SRC=a,b selects other source files)
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
SRCS = os.environ.get("AB_SRC", "pipeline_fast,pipeline_fast_f32a,pipeline_fast_f32b,pipeline_fast_i16a,pipeline_fast_i16b").split(",")
for spec in sys.argv[1:]:
    name, _, defs = spec.partition("=")
    defs = [d for d in defs.split(",") if d]
    new = []
    for src in SRCS:
        o = os.path.join(out, f"{src}_{name}.o")
        subprocess.check_call([NVCC] + FLAGS + defs + ["-I", os.path.join(ROOT, "include"), "-c", os.path.join(CSRC, src + ".cu"), "-o", o])
        new.append(o)
    skip = {s_ + ".o" for s_ in SRCS}
    objs = [x for x in sorted(glob.glob(os.path.join(OBJ, "*.o"))) if os.path.basename(x) not in skip] + new
    lib = os.path.join(out, f"libepb200_{name}.so")
    subprocess.check_call([NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    print("built", lib)
