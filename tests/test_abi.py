"""The C-ABI library loads and exports every symbol include/zr_abi.h declares (no compute, no GPU)."""
import ctypes as C

import numpy as np


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


def test_scene_create_rejects_malformed_descriptions_on_the_host():
    """zr_scene_create validates what the kernels will index (index ranges, vertex indices, material and emissive offsets, 64-bit
    arithmetic) before anything is allocated, so a malformed description is an error code even without a GPU -- never an
    out-of-bounds device read."""
    import ctypes as C
    from zetaray_b200 import _lib
    from zetaray_b200.passes import _vp
    from tests import scene_util
    flat = scene_util.SCENES["cornell"]()

    def try_create(mutate):
        arrs = [np.ascontiguousarray(x).copy() for x in (flat.vertices, flat.indices, flat.instances, flat.instance_num_tris,
                                                         flat.materials, flat.emissives)]
        mutate(*arrs)
        v, i, inst, nt, m, e = arrs
        d = _lib.SceneDesc()
        d.h_vertices, d.num_vertices = _vp(v), len(v)
        d.h_indices, d.num_indices = _vp(i), len(i)
        d.h_instances, d.num_instances = _vp(inst), len(inst)
        d.h_instance_num_tris = _vp(nt)
        d.h_materials, d.num_materials = _vp(m), len(m)
        d.h_emissives, d.num_emissives = (_vp(e) if len(e) else None), len(e)
        h = C.c_void_p()
        rc = _lib.lib.zr_scene_create(C.byref(d), C.byref(h))
        return rc, _lib.lib.zr_last_error().decode()

    def bad_index(v, i, inst, nt, m, e): i[5] = len(v) + 7
    def bad_material(v, i, inst, nt, m, e): inst["MatIdx"][0] = len(m)
    def bad_range(v, i, inst, nt, m, e): nt[-1] = 0x7fffffff           # BaseIdxOffset + 3 * n wraps in 32 bits
    def bad_emissive(v, i, inst, nt, m, e):
        k = int(np.argmax(inst["BaseEmissiveTriOffset"] != 0xffffffff))
        inst["BaseEmissiveTriOffset"][k] = len(e)
    for mutate, word in ((bad_index, "vertex buffer"), (bad_material, "material"), (bad_range, "index buffer"), (bad_emissive, "emissive")):
        rc, msg = try_create(mutate)
        assert rc == 1 and word in msg, (mutate.__name__, rc, msg)       # ZR_ERR_INVALID_ARG
