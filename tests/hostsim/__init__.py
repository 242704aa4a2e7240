"""TEST INFRASTRUCTURE: the product's BVH traversal source compiled for the host (g++), see hostsim.cpp."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libhostsim.so")


def build(force=False):
    src = os.path.join(HERE, "hostsim.cpp")
    csrc = os.path.join(os.path.dirname(os.path.dirname(HERE)), "zetaray_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("zr_scene.cuh", "zr_common.cuh", "zr_bvh.h", "zr_bsdf.cuh", "zr_rt.cuh", "zr_rpt.cuh", "zr_pixel.cuh", "zr_rgi.cuh")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return SO
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I" + cuda_inc,
           "-D__forceinline__=inline __attribute__((always_inline))", src, "-o", SO, "-lpthread"]
    subprocess.check_call(cmd)
    return SO


def load():
    lib = C.CDLL(build())
    lib.hostsim_validate.restype = C.c_uint64
    lib.hostsim_bsdf_sampler_pdf.restype = C.c_float
    lib.hostsim_bsdf_sampler_pdf_nodiffuse.restype = C.c_float
    return lib
