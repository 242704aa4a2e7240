"""BVH traversal + G-buffer + pre-lighting on the device vs the oracle's brute force (bit-exact)."""
import ctypes as C
import numpy as np
import pytest

from tests.orc import ptr


def _rays(n, seed, flat_bounds=2.5):
    rng = np.random.default_rng(seed)
    o = (rng.random((n, 3), dtype=np.float32) * 2 - 1) * np.float32(flat_bounds) + np.array([0, 1, 0], dtype=np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    # axis-aligned and degenerate directions too
    d[: n // 20] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, size=n // 20)] * np.where(rng.random((n // 20, 1)) < 0.5, -1, 1).astype(np.float32)
    r = np.zeros((n, 8), dtype=np.float32)
    r[:, 0:3] = o
    r[:, 3] = np.where(rng.random(n) < 0.5, 0.0, 1e-6).astype(np.float32)
    r[:, 4:7] = d
    r[:, 7] = np.where(rng.random(n) < 0.7, np.float32(3.402823466e+38), rng.random(n).astype(np.float32) * 3).astype(np.float32)
    return r


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["cornell", "glossy"])
def test_trace_vs_bruteforce(which):
    import torch
    from zetaray_b200 import lib, check
    from zetaray_b200.passes import Scene
    from tests import scene_util
    from tests.gpu_util import dev, dptr, host, stream
    flat = scene_util.cornell() if which == "cornell" else scene_util.glossy_cornell()
    osc = scene_util.OracleScene(flat)
    sc = Scene(flat)
    st = sc.bvh_stats()
    assert st["tris"] == flat.num_triangles and st["nodes"] >= 1
    n = 200000
    rays = _rays(n, 5)
    ref = np.zeros((n, 4), dtype=np.float32)
    osc.o.orc_trace_closest(osc.h, ptr(rays), n, ptr(ref))
    ref_any = np.zeros(n, dtype=np.uint32)
    osc.o.orc_trace_any(osc.h, ptr(rays), n, ptr(ref_any))
    d_r = dev(rays)
    d_h = torch.zeros(n * 4, dtype=torch.float32, device="cuda")
    d_f = torch.zeros(n, dtype=torch.int32, device="cuda")
    check(lib.zr_scene_trace_closest(sc.handle, dptr(d_r), n, dptr(d_h), stream()))
    check(lib.zr_scene_trace_any(sc.handle, dptr(d_r), n, dptr(d_f), stream()))
    torch.cuda.synchronize()
    got = d_h.cpu().numpy().reshape(n, 4)
    assert (ref[:, 0] < 1e30).mean() > 0.3
    assert got.view(np.uint32).tobytes() == ref.view(np.uint32).tobytes()
    assert (d_f.cpu().numpy().view(np.uint32) == ref_any).all()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,jitter,dof", [(320, 180, (0.0, 0.0), 0), (481, 271, (0.25, -0.3333), 0), (320, 180, (0.1, 0.2), 1)])
def test_gbuffer_and_prelighting(w, h, jitter, dof):
    import torch
    from zetaray_b200 import lib, check, _lib
    from zetaray_b200.passes import Scene, GBuffers, GBufferRT
    from tests import scene_util, synth
    flat = scene_util.glossy_cornell()
    osc = scene_util.OracleScene(flat)
    sc = Scene(flat)
    sc.prelighting()
    alias = sc.alias_table()
    assert alias.tobytes() == osc.alias.tobytes()
    fc = synth.look_at_frame_constants(w, h, frame=7, jitter=jitter, prev_jitter=(0.1, -0.1))
    if dof:
        fc.DoF, fc.FocusDepth, fc.LensRadius = 1, 4.0, 0.05
    ref = osc.gbuffer(fc, tridiff=True)
    gb = GBuffers(w, h, with_tridiff=True)
    fi = _lib.FrameInputs()
    fi.frame = fc
    gb.fill_inputs(fi)
    fi.scene = sc.handle
    p = GBufferRT()
    p.Render(fi, None)
    check(lib.zr_stream_synchronize(None))
    got = gb.download()
    names = ["core", "depth", "motion_emissive", "coat", "tridiff"]
    for name, a, b in zip(names, got, ref):
        assert a.tobytes() == b.tobytes(), name
    gb.close()
