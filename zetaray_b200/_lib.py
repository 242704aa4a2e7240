"""ctypes binding of include/zr_abi.h. No compute happens here."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("ZETARAY_B200_LIB") or os.path.join(HERE, "libzetaray_b200.so")


class ZRError(RuntimeError):
    pass


def _load():
    if not os.path.exists(SO_PATH):
        raise ZRError(
            "libzetaray_b200.so is missing -- run `python -m zetaray_b200.build` (there is no CPU fallback)")
    try:
        import torch  # noqa: F401  (makes the process share torch's libcudart.so.12)
    except Exception:
        pass
    return C.CDLL(SO_PATH)


class _NoLibrary:
    """ZETARAY_B200_STRUCTS_ONLY=1 (bench.py's CPU reference arm): only the ctypes mirrors of the ABI structs are needed, the
    shared library is NOT mapped into the process; any call through `lib` fails loudly."""

    def __getattr__(self, name):
        if name in ("zr_last_error", "zr_abi_version", "zr_kernel_launch_count"):
            return type("_Stub", (), {"restype": None, "argtypes": None})()
        raise ZRError("zetaray_b200 was imported with ZETARAY_B200_STRUCTS_ONLY=1: %s is not available in this process" % name)


lib = _NoLibrary() if os.environ.get("ZETARAY_B200_STRUCTS_ONLY") == "1" else _load()

u32, u64, f32, vp, i32 = C.c_uint32, C.c_uint64, C.c_float, C.c_void_p, C.c_int32


class FrameConstants(C.Structure):
    _fields_ = [
        ("CurrView", f32 * 12), ("PrevView", f32 * 12), ("CurrViewInv", f32 * 12), ("PrevViewInv", f32 * 12),
        ("CurrViewProj", f32 * 16), ("PrevViewProj", f32 * 16),
        ("CameraPos", f32 * 3), ("CameraNear", f32),
        ("AspectRatio", f32), ("PixelSpreadAngle", f32), ("TanHalfFOV", f32), ("dt", f32),
        ("FrameNum", u32), ("CurrGBufferDescHeapOffset", u32), ("PrevGBufferDescHeapOffset", u32),
        ("BaseColorMapsDescHeapOffset", u32),
        ("NormalMapsDescHeapOffset", u32), ("MetallicRoughnessMapsDescHeapOffset", u32),
        ("EmissiveMapsDescHeapOffset", u32), ("EnvMapDescHeapOffset", u32),
        ("RenderWidth", u32), ("RenderHeight", u32), ("DisplayWidth", u32), ("DisplayHeight", u32),
        ("CurrCameraJitter", f32 * 2), ("PrevCameraJitter", f32 * 2),
        ("PlanetRadius", f32), ("SunCosAngularRadius", f32), ("SunSinAngularRadius", f32), ("pad", f32),
        ("SunDir", f32 * 3), ("SunIlluminance", f32),
        ("RayleighSigmaSColor", f32 * 3), ("RayleighSigmaSScale", f32),
        ("OzoneSigmaAColor", f32 * 3), ("OzoneSigmaAScale", f32),
        ("MieSigmaS", f32), ("MieSigmaA", f32), ("AtmosphereAltitude", f32), ("g", f32),
        ("NumFramesCameraStatic", u32), ("CameraStatic", u32), ("Accumulate", u32), ("SunMoved", u32),
        ("CameraRayUVGradsScale", f32), ("MipBias", f32), ("OneDivNumEmissiveTriangles", f32),
        ("NumEmissiveTriangles", u32),
        ("FocusDepth", f32), ("LensRadius", f32), ("DoF", u32), ("pad2", u32),
    ]


class GBuffer(C.Structure):
    _fields_ = [("d_core", vp), ("d_depth", vp), ("d_motion_emissive", vp), ("d_coat", vp), ("d_tridiff", vp)]


class FrameInputs(C.Structure):
    _fields_ = [("frame", FrameConstants), ("curr", GBuffer), ("prev", GBuffer), ("scene", vp)]


class Image2D(C.Structure):
    _fields_ = [("d_ptr", vp), ("width", u32), ("height", u32), ("pitch_bytes", u32), ("texel_bytes", u32)]


class SceneDesc(C.Structure):
    _fields_ = [
        ("h_vertices", vp), ("num_vertices", u32),
        ("h_indices", vp), ("num_indices", u32),
        ("h_instances", vp), ("num_instances", u32),
        ("h_instance_num_tris", vp),
        ("h_materials", vp), ("num_materials", u32),
        ("h_emissives", vp), ("num_emissives", u32),
    ]


class DirectParams(C.Structure):
    _fields_ = [("temporal_resample", u32), ("spatial_resample", u32), ("stochastic_spatial", u32),
                ("extra_disocclusion_sampling", u32), ("M_max", u32), ("alpha_min", f32)]


class SvgfParams(C.Structure):
    _fields_ = [("sigma_z", f32), ("k_n", f32), ("sigma_l", f32), ("radius", u32), ("num_passes", u32)]


class IndirectParams(C.Structure):
    _fields_ = [("max_non_tr_bounces", u32), ("max_glossy_tr_bounces", u32), ("russian_roulette", u32),
                ("temporal_resample", u32), ("num_spatial_passes", u32), ("M_max_temporal", u32),
                ("M_max_spatial", u32), ("boiling_suppression", u32), ("sort_temporal", u32),
                ("sort_spatial", u32), ("alpha_min", f32)]


class CompositingParams(C.Structure):
    _fields_ = [("emissive_di", u32), ("indirect", u32), ("firefly_filter", u32)]


lib.zr_last_error.restype = C.c_char_p
lib.zr_abi_version.restype = u32
lib.zr_kernel_launch_count.restype = u64

class GIParams(C.Structure):
    _fields_ = [("max_non_tr_bounces", u32), ("max_glossy_tr_bounces", u32), ("russian_roulette", u32), ("stochastic_multi_bounce", u32),
                ("boiling_suppression", u32), ("M_max", u32), ("temporal_resample", u32)]


class RendererDesc(C.Structure):
    _fields_ = [("width", u32), ("height", u32), ("with_tridiff", C.c_int), ("two_streams", C.c_int)]


# zr_halo_exchange_fn (include/zr_abi.h "Strip-sharded frames")
HALO_EXCHANGE_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(Image2D), C.c_int, C.c_void_p)

# every other entry point returns zr_status (int32)
EXPORTS = [
    "zr_last_error", "zr_abi_version", "zr_kernel_launch_count",
    "zr_device_malloc", "zr_device_free", "zr_memcpy_h2d", "zr_memcpy_d2h", "zr_memset_d", "zr_stream_synchronize",
    "zr_alias_table_build", "zr_alias_table_sample",
]


def check(status):
    if status != 0:
        raise ZRError("zr_status %d: %s" % (status, lib.zr_last_error().decode()))


def declared_symbols():
    """All ZR_API functions declared in include/zr_abi.h (parsed from the header)."""
    import re
    hdr = os.path.join(os.path.dirname(HERE), "include", "zr_abi.h")
    txt = open(hdr).read()
    return sorted(set(re.findall(r"ZR_API\s+[\w\s\*]+?\b(zr_\w+)\s*\(", txt)))
