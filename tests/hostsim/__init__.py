"""TEST INFRASTRUCTURE: the product's device headers compiled for the host (g++) -- hostsim.cpp (traversal, BSDF, ray queries, ReSTIR PT
shift, ReSTIR GI reuse, probes mirrored by the oracle) and hostsim_di.cpp (ReSTIR DI; a translation unit of its own)."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libhostsim.so")
SO_DI = os.path.join(HERE, "libhostsim_di.so")
SO_IO = os.path.join(HERE, "libhostsim_io.so")
_CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "zetaray_b200", "csrc")
_INC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "include")


def _deps(src):
    hdrs = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cuh", ".h"))]
    hdrs += [os.path.join(_INC, f) for f in os.listdir(_INC)]
    return [src, os.path.join(HERE, "prelude.h")] + hdrs


def _build(src, so, force, defines=()):
    if not force and os.path.exists(so) and all(os.path.getmtime(d) <= os.path.getmtime(so) for d in _deps(src)):
        return so
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I" + cuda_inc,
           "-D__forceinline__=inline __attribute__((always_inline))"] + ["-D" + d for d in defines] + [src, "-o", so, "-lpthread"]
    subprocess.check_call(cmd)
    return so


def build(force=False):
    _build(os.path.join(HERE, "hostsim_di.cpp"), SO_DI, force)
    _build(os.path.join(HERE, "hostsim_io.cpp"), SO_IO, force)
    return _build(os.path.join(HERE, "hostsim.cpp"), SO, force)


def load():
    lib = C.CDLL(build())
    lib.hostsim_validate.restype = C.c_uint64
    lib.hostsim_bsdf_sampler_pdf.restype = C.c_float
    lib.hostsim_bsdf_sampler_pdf_nodiffuse.restype = C.c_float
    return lib


def load_di():
    build()
    return C.CDLL(SO_DI)


def load_io():
    build()
    lib = C.CDLL(SO_IO)
    lib.hostsim_oct32_round_trip.restype = C.c_uint64
    return lib
