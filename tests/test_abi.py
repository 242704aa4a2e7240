"""The C-ABI library loads and exports every symbol include/zr_abi.h declares (no compute, no GPU)."""
import ctypes as C


def test_library_exports_all_declared_symbols():
    from zetaray_b200 import _lib
    declared = _lib.declared_symbols()
    assert len(declared) > 40
    missing = [s for s in declared if not hasattr(_lib.lib, s)]
    assert not missing, missing


def test_abi_version_and_struct_sizes():
    from zetaray_b200 import _lib
    assert _lib.lib.zr_abi_version() >> 16 == 1
    # cbFrameConstants layout (FrameConstants.h:84-97 offset asserts)
    FC = _lib.FrameConstants
    for f in ("CameraPos", "AspectRatio", "FrameNum", "NormalMapsDescHeapOffset", "RenderWidth", "CurrCameraJitter",
              "PlanetRadius", "SunDir", "RayleighSigmaSColor", "OzoneSigmaAColor", "MieSigmaS",
              "NumFramesCameraStatic", "CameraRayUVGradsScale"):
        assert getattr(FC, f).offset % 16 == 0, f
    assert C.sizeof(FC) == 544
