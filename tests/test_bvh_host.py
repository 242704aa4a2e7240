"""CPU tier: the product's BVH builder (zr_bvh_build_host, plain C++ inside the C-ABI library) and the product's traversal
source compiled for the host (tests/hostsim) against brute force -- on the Cornell fixture and on the procedural
"Sponza-class" / "Subway-class" scenes at the sizes the benchmark uses (3 x 10^5 and 10^6 triangles), where the oracle's
brute-force renderer cannot follow. The GPU tier repeats the ray comparison through the real kernels (test_scene_gpu.py)."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import hostsim
from tests.orc import ptr
from tests.scene_util import OracleScene, cornell
from zetaray_b200 import lib, check, procedural

NODE_BYTES = 80
STACK_ENTRIES = 96          # zr_bvh.h BVH_STACK_ENTRIES
THREADS = min(os.cpu_count() or 8, 32)


def world_tris(flat):
    """World-space triangles exactly as the traversal sees them (the oracle applies the shaders' TransformTRS arithmetic)."""
    osc = OracleScene(flat)
    n = osc.o.orc_scene_num_tris(osc.h)
    wt = np.zeros((n, 9), dtype=np.float32)
    osc.o.orc_scene_get_tris(osc.h, ptr(wt))
    tri_mesh = np.repeat(np.arange(len(flat.instances), dtype=np.uint32), flat.instance_num_tris)
    first = np.concatenate([[0], np.cumsum(flat.instance_num_tris)[:-1]]).astype(np.uint32)
    return wt, tri_mesh, first


def build(wt):
    n = len(wt)
    info = (C.c_uint32 * 4)()
    check(lib.zr_bvh_build_host(ptr(wt), n, None, 0, None, info))
    nodes = np.zeros(info[0] * NODE_BYTES, dtype=np.uint8)
    order = np.zeros(n, dtype=np.uint32)
    check(lib.zr_bvh_build_host(ptr(wt), n, ptr(nodes), info[0], ptr(order), info))
    leaf = np.zeros((n, 12), dtype=np.float32)
    leaf[:, 0:3] = wt[order, 0:3]
    leaf[:, 3] = order.view(np.float32)
    leaf[:, 4:7] = wt[order, 3:6]
    leaf[:, 8:11] = wt[order, 6:9]
    return nodes, order, leaf, tuple(info)


def make_rays(wt, n, seed, cam):
    """Half camera-like rays from the eye, half rays between random points of the scene's bounding box (tmax at the far
    point, like shadow segments); a few axis-parallel ones (zero direction components -> infinite reciprocals)."""
    rng = np.random.default_rng(seed)
    lo = np.minimum(wt[:, 0:3], np.minimum(wt[:, 0:3] + wt[:, 3:6], wt[:, 0:3] + wt[:, 6:9])).min(axis=0)
    hi = np.maximum(wt[:, 0:3], np.maximum(wt[:, 0:3] + wt[:, 3:6], wt[:, 0:3] + wt[:, 6:9])).max(axis=0)
    rays = np.zeros((n, 8), dtype=np.float32)
    h = n // 2
    d = rng.normal(size=(h, 3)); d[:, 2] = np.abs(d[:, 2]) + 0.5
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[:h, 0:3] = cam; rays[:h, 3] = 1e-6; rays[:h, 4:7] = d; rays[:h, 7] = 3.0e38
    a = lo + (hi - lo) * (0.02 + 0.96 * rng.random((n - h, 3)))
    b = lo + (hi - lo) * (0.02 + 0.96 * rng.random((n - h, 3)))
    dd = b - a
    ln = np.linalg.norm(dd, axis=1, keepdims=True)
    rays[h:, 0:3] = a; rays[h:, 3] = 3e-6; rays[h:, 4:7] = dd / ln; rays[h:, 7] = ln[:, 0]
    k = min(16, n - h)
    for j in range(k):          # axis-parallel
        ax = j % 3
        v = np.zeros(3); v[ax] = 1.0 if j % 2 else -1.0
        rays[h + j, 4:7] = v; rays[h + j, 7] = 3.0e38
    # degenerate rays: zero direction, NaN direction, NaN origin -- no hit under the hit rule, and the traversal must say so
    # without sweeping the tree (zr_scene.cuh::Traverse's early-out)
    if n - h > k + 3:
        rays[h + k, 4:7] = 0.0; rays[h + k, 7] = 3.0e38
        rays[h + k + 1, 4] = np.nan; rays[h + k + 1, 7] = 3.0e38
        rays[h + k + 2, 0] = np.nan; rays[h + k + 2, 7] = 3.0e38
    return rays


CASES = [
    ("cornell", lambda: cornell(), (0.0, 1.2, -4.043), 4000),
    ("atrium-small", lambda: procedural.atrium(0.12), procedural.ATRIUM_CAMERA, 4000),
    ("tunnel-small", lambda: procedural.tunnel(0.1), procedural.TUNNEL_CAMERA, 4000),
    ("atrium-C4", lambda: procedural.atrium(1.0), procedural.ATRIUM_CAMERA, 1500),
    ("tunnel-C5", lambda: procedural.tunnel(1.0), procedural.TUNNEL_CAMERA, 600),
]


@pytest.mark.parametrize("name,make,cam,nrays", CASES, ids=[c[0] for c in CASES])
def test_builder_and_traversal_against_brute_force(name, make, cam, nrays):
    hs = hostsim.load()
    flat = make()
    wt, tri_mesh, first = world_tris(flat)
    assert len(wt) == flat.num_triangles
    nodes, order, leaf, info = build(wt)
    num_nodes, num_tris, max_depth, max_stack = info
    assert num_tris == len(wt)
    assert sorted(order.tolist()) == list(range(len(wt))) if len(wt) < 50000 else np.array_equal(np.sort(order), np.arange(len(wt)))
    # the kernels' traversal stack holds every tree the benchmark scenes produce (scene creation refuses others)
    assert max_stack <= STACK_ENTRIES, (name, max_stack)
    stats = (C.c_uint64 * 4)()
    bad = hs.hostsim_validate(ptr(nodes), num_nodes, ptr(leaf), num_tris, stats)
    assert bad == 0, (name, bad)
    assert stats[0] == num_nodes and stats[3] == num_tris and stats[2] == max_depth
    # product traversal (host build of the same source) == brute force, bit for bit: t, barycentrics, triangle
    rays = make_rays(wt, nrays, 11, np.array(cam, dtype=np.float32))
    got = np.zeros((nrays, 4), dtype=np.float32)
    anyf = np.zeros(nrays, dtype=np.uint32)
    hs.hostsim_trace(ptr(nodes), ptr(leaf), ptr(tri_mesh), ptr(first), ptr(rays), nrays, ptr(got), ptr(anyf), None, THREADS)
    ref = np.zeros((nrays, 4), dtype=np.float32)
    hs.hostsim_brute(ptr(wt), num_tris, ptr(rays), nrays, ptr(ref), THREADS)
    assert got.tobytes() == ref.tobytes(), (name, int((got.view(np.uint32) != ref.view(np.uint32)).any(axis=1).sum()))
    hit = ref[:, 0] < 3.0e38
    assert np.array_equal(anyf != 0, hit), name
    deg = nrays // 2 + 16
    assert not hit[deg:deg + 3].any()           # the three degenerate rays
    assert hit.mean() > 0.2       # not a degenerate comparison: a good share of the rays hit something


def test_build_host_argument_errors():
    info = (C.c_uint32 * 4)()
    assert lib.zr_bvh_build_host(None, 1, None, 0, None, info) != 0
    wt = np.zeros((4, 9), dtype=np.float32)
    wt[:, 3] = 1.0; wt[:, 7] = 1.0; wt[:, 0] = np.arange(4)
    nodes = np.zeros(NODE_BYTES, dtype=np.uint8)
    check(lib.zr_bvh_build_host(ptr(wt), 4, None, 0, None, info))
    assert info[0] >= 1 and info[1] == 4
    assert lib.zr_bvh_build_host(ptr(wt), 4, ptr(nodes), 0, None, info) != 0      # capacity too small
