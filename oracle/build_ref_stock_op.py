"""TEST/BENCH INFRASTRUCTURE — builds the REFERENCE's stock CUDA op as the on-GPU comparator.

Compiles the unmodified kernels of /root/reference/projects/IDOL/idol/models/ops/src (where they
lie; nothing is copied into this repo) for sm_100a into oracle/_ref/msda_stock_ref.so.  The only
change is the 2-token torch>=2 compat patch SURVEY §8c describes (`value.type()` ->
`value.scalar_type()` inside AT_DISPATCH_FLOATING_TYPES at src/cuda/ms_deform_attn_cuda.cu:64,134),
applied with sed to a scratch copy under /tmp.  Kernels (ms_deform_im2col_cuda.cuh) are untouched.
Only runs where /root/reference exists (the build container); the GPU box uses the prebuilt .so.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/projects/IDOL/idol/models/ops/src"
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "msda_stock_ref.so")


def main():
    if not os.path.isdir(SRC):
        print("reference sources not present; nothing to do")
        return 0
    if os.path.exists(OUT):
        print("already built:", OUT)
        return 0
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="msda_stock_")
    work = os.path.join(tmp, "src")
    shutil.copytree(SRC, work)
    cu = os.path.join(work, "cuda", "ms_deform_attn_cuda.cu")
    subprocess.check_call(["sed", "-i", "s/AT_DISPATCH_FLOATING_TYPES(value.type(),/AT_DISPATCH_FLOATING_TYPES(value.scalar_type(),/", cu])
    inc = ce.include_paths(device_type="cuda") if "device_type" in ce.include_paths.__code__.co_varnames else ce.include_paths(cuda=True)
    incs = sum((["-I", p] for p in inc + [work]), [])
    import sysconfig
    incs += ["-I", sysconfig.get_paths()["include"]]
    defs = ["-DWITH_CUDA", "-DTORCH_EXTENSION_NAME=msda_stock_ref", "-DTORCH_API_INCLUDE_EXTENSION_H",
            "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    objs = []
    for rel in ("vision.cpp", "cpu/ms_deform_attn_cpu.cpp"):
        o = os.path.join(tmp, rel.replace("/", "_") + ".o")
        subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-c", os.path.join(work, rel), "-o", o] + incs + defs)
        objs.append(o)
    o = os.path.join(tmp, "cuda.o")
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                           "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
                           "-D__CUDA_NO_HALF2_OPERATORS__", "-c", cu, "-o", o] + incs + defs)
    objs.append(o)
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    subprocess.check_call(["g++", "-shared", "-o", OUT] + objs +
                          ["-L", libdir, "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda", "-ltorch_python",
                           "-L", "/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + libdir])
    shutil.rmtree(tmp, ignore_errors=True)
    print("built", OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
