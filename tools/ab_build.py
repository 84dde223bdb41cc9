"""Build variants of libepb200.so for A/B timing on the GPU box in one gpurun call.

    python tools/ab_build.py NAME=-DFLAG[,-DFLAG2] ...   ->  gpurun_ab/libepb200_NAME.so  (pipeline_fast.cu recompiled;
                                                              AB_SRC=masknoise selects another source file)
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
SRC = os.environ.get("AB_SRC", "pipeline_fast")
for spec in sys.argv[1:]:
    name, _, defs = spec.partition("=")
    defs = [d for d in defs.split(",") if d]
    o = os.path.join(out, f"{SRC}_{name}.o")
    subprocess.check_call([NVCC] + FLAGS + defs + ["-I", os.path.join(ROOT, "include"), "-c", os.path.join(CSRC, SRC + ".cu"), "-o", o])
    objs = [x for x in sorted(glob.glob(os.path.join(OBJ, "*.o"))) if not x.endswith(SRC + ".o")] + [o]
    lib = os.path.join(out, f"libepb200_{name}.so")
    subprocess.check_call([NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    print("built", lib)
