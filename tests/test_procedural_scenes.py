"""CPU tier: the procedural C4 / C5 stand-in scenes are what they claim to be -- sizes SURVEY 8d asks for, the reference's
buffer conventions, and geometry that faces the camera (single-sided materials are invisible from behind, so a patch
with the wrong handedness would silently render black)."""
import numpy as np
import pytest

from tests import hostsim
from tests.orc import ptr
from tests.test_bvh_host import world_tris, build, THREADS
from zetaray_b200 import procedural, scene as zscene


def _camera_rays(cam, w, h):
    tan = np.tan(0.5 * np.pi / 3)
    xs = ((np.arange(w) + 0.5) / w * 2 - 1) * tan * (w / h)
    ys = (1 - (np.arange(h) + 0.5) / h * 2) * tan
    X, Y = np.meshgrid(xs, ys)
    d = np.stack([X, Y, np.ones_like(X)], -1).reshape(-1, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((w * h, 8), dtype=np.float32)
    rays[:, 0:3] = cam; rays[:, 3] = 1e-4; rays[:, 4:7] = d; rays[:, 7] = 3.0e38
    return rays


@pytest.mark.parametrize("name,detail", [("atrium", 0.12), ("tunnel", 0.1), ("atrium", 0.5)])
def test_camera_sees_front_faces(name, detail):
    make, cam = procedural.SCENES[name]
    flat = make(detail)
    wt, tri_mesh, first = world_tris(flat)
    nodes, order, leaf, info = build(wt)
    hs = hostsim.load()
    rays = _camera_rays(np.array(cam, dtype=np.float32), 96, 54)
    hits = np.zeros((len(rays), 4), dtype=np.float32)
    hs.hostsim_trace(ptr(nodes), ptr(leaf), ptr(tri_mesh), ptr(first), ptr(rays), len(rays), ptr(hits), None, None, THREADS)
    hit = hits[:, 0] < 3.0e38
    assert hit.mean() > 0.995, "closed scene: every camera ray must hit something"
    tri = hits[hit, 3].view(np.uint32)
    ng = np.cross(wt[tri, 3:6], wt[tri, 6:9])               # clockwise front faces: cross(e1, e2) is the front normal
    facing = -(ng * rays[hit, 4:7]).sum(axis=1)
    mat = flat.instances["MatIdx"][tri_mesh[tri]]
    double_sided = (flat.materials["CoatColor_Flags"][mat] & (1 << 25)) != 0
    back = (facing < 0) & ~double_sided
    assert back.mean() < 0.01, "%d of %d primary hits see the back of a single-sided surface" % (back.sum(), len(back))


def test_benchmark_sizes_and_conventions():
    a = procedural.atrium(1.0)
    assert 250_000 <= a.num_triangles <= 350_000 and len(a.materials) == 25
    assert len(a.emissives) >= 13107            # the reference's presampling threshold (DefaultRendererImpl.h:37-41)
    t = procedural.tunnel(1.0)
    assert 1_000_000 <= t.num_triangles <= 2_000_000
    assert len(t.emissives) > 10_000
    for s in (a, t):
        assert s.vertices.dtype == zscene.VERTEX and s.instances.dtype == zscene.MESH_INSTANCE and s.emissives.dtype == zscene.EMISSIVE_TRI
        assert len(s.instances) < 65535 and int(s.instances["MatIdx"].max()) < len(s.materials)
        em = s.instances["BaseEmissiveTriOffset"] != 0xffffffff
        assert int(s.instance_num_tris[em].sum()) == len(s.emissives)
        # emissive ranges are consecutive in instance order
        off = s.instances["BaseEmissiveTriOffset"][em].astype(np.int64)
        assert np.array_equal(off, np.concatenate([[0], np.cumsum(s.instance_num_tris[em])[:-1]]))
        # instancing: several instances share one vertex / index range
        assert len(np.unique(s.instances["BaseIdxOffset"])) < len(s.instances)
        # transmissive and metallic materials are present in the tunnel / atrium (the glossy + glass budget)
        flags = s.materials["CoatColor_Flags"]
        assert ((flags >> 24) & 1).any() and ((flags >> 26) & 1).any()


def test_batch_emissive_constructor_matches_the_scalar_one():
    rng = np.random.default_rng(3)
    n = 200
    v0 = rng.normal(size=(n, 3)) * 5; v1 = v0 + rng.normal(size=(n, 3)); v2 = v0 + rng.normal(size=(n, 3))
    uv = rng.random((3, n, 2))
    ids = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    batch = zscene.emissive_triangles(v0, v1, v2, uv[0], uv[1], uv[2], 0x8040ff, 0x4a00, ids, True)
    for k in range(n):
        one = zscene.emissive_triangle(v0[k], v1[k], v2[k], uv[0][k], uv[1][k], uv[2][k], 0x8040ff, 0x4a00, int(ids[k]), True)
        assert batch[k].tobytes() == one.tobytes(), k
    x, y, z = zscene.pcg3d_np([1, 7, 0xffffffff], [0, 0, 5], [3, 9, 2])
    for k, args in enumerate([(1, 0, 3), (7, 0, 9), (0xffffffff, 5, 2)]):
        assert (int(x[k]), int(y[k]), int(z[k])) == zscene.pcg3d(*args)
