"""CPU tier: traversal cost of the product's BVH on the benchmark-size scenes, counted on the host build of the traversal source
(node visits and triangle tests per ray). Guards the builder against quality regressions and the traversal against rays that sweep
the tree (degenerate rays must cost nothing); the numbers are the ones DESIGN.md section 10 quotes."""

import numpy as np
import pytest

from tests import hostsim
from tests.orc import ptr
from tests.test_bvh_host import world_tris, build
from tests.test_procedural_scenes import _camera_rays
from zetaray_b200 import procedural


def ray_sets(flat, wt, cam, trace_closest):
    """Primary rays from the scene's camera, cosine-distributed secondary rays from the primary hits, shadow segments from the
    primary hits to random emissive triangles."""
    W, H = 160, 90
    prim = _camera_rays(np.array(cam, dtype=np.float32), W, H)
    hits = trace_closest(prim)
    ok = hits[:, 0] < 3e38
    tri = hits[ok, 3].view(np.uint32)
    P = prim[ok, 0:3] + prim[ok, 4:7] * hits[ok, 0:1]
    ng = np.cross(wt[tri, 3:6], wt[tri, 6:9])
    ng /= np.linalg.norm(ng, axis=1, keepdims=True) + 1e-30
    ng = np.where(((ng * prim[ok, 4:7]).sum(1) > 0)[:, None], -ng, ng)
    rng = np.random.default_rng(1)
    m = len(P)
    u = rng.random((m, 2))
    r = np.sqrt(u[:, 0]); ph = 2 * np.pi * u[:, 1]
    tmp = np.where(np.abs(ng[:, 0:1]) < 0.9, [[1.0, 0, 0]], [[0, 1.0, 0]])
    t1 = np.cross(ng, tmp); t1 /= np.linalg.norm(t1, axis=1, keepdims=True)
    t2 = np.cross(ng, t1)
    d = (r * np.cos(ph))[:, None] * t1 + (r * np.sin(ph))[:, None] * t2 + np.sqrt(1 - u[:, 0])[:, None] * ng
    sec = np.zeros((m, 8), dtype=np.float32)
    sec[:, 0:3] = P + 1e-3 * ng; sec[:, 3] = 1e-6; sec[:, 4:7] = d; sec[:, 7] = 3e38
    L = flat.emissives["Vtx0"].astype(np.float64)[rng.integers(0, len(flat.emissives), m)]
    dd = L - P
    ln = np.linalg.norm(dd, axis=1, keepdims=True)
    sh = np.zeros((m, 8), dtype=np.float32)
    sh[:, 0:3] = P + 1e-3 * ng; sh[:, 3] = 3e-6; sh[:, 4:7] = dd / ln; sh[:, 7] = ln[:, 0] * 0.999
    return prim, sec, sh


@pytest.mark.parametrize("name", ["atrium", "tunnel"])
def test_node_visits_per_ray_on_benchmark_scenes(name):
    hs = hostsim.load()
    make, cam = procedural.SCENES[name]
    flat = make(1.0)
    wt, tri_mesh, first = world_tris(flat)
    nodes, order, leaf, info = build(wt)

    def closest(rays):
        out = np.zeros((len(rays), 4), dtype=np.float32)
        hs.hostsim_trace(ptr(nodes), ptr(leaf), ptr(tri_mesh), ptr(first), ptr(rays), len(rays), ptr(out), None, None, 8)
        return out

    def stats(rays, anyhit):
        a = np.zeros(len(rays), dtype=np.uint32); b = np.zeros(len(rays), dtype=np.uint32)
        hs.hostsim_trace_stats(ptr(nodes), ptr(leaf), ptr(tri_mesh), ptr(first), ptr(rays), len(rays), ptr(a), ptr(b), anyhit)
        return a, b
    prim, sec, sh = ray_sets(flat, wt, cam, closest)
    for label, rays, anyhit in (("primary", prim, 0), ("secondary", sec, 0), ("shadow", sh, 1)):
        n, t = stats(rays, anyhit)
        print("%s %-9s node visits mean %.1f p99 %d max %d | triangle tests mean %.1f max %d" % (name, label, n.mean(), np.percentile(n, 99), n.max(), t.mean(), t.max()))
        assert n.mean() < 20, (name, label, n.mean())           # measured: 9.6 - 14.3
        assert np.percentile(n, 99) < 60, (name, label)         # measured: <= 33
        assert n.max() < 2000 and t.mean() < 8, (name, label, n.max(), t.mean())
    # degenerate rays: zero direction, NaN direction, NaN origin -> no node is visited at all
    deg = prim[:3].copy()
    deg[0, 4:7] = 0.0; deg[1, 4] = np.nan; deg[2, 0] = np.nan
    for anyhit in (0, 1):
        n, t = stats(deg, anyhit)
        assert n.max() == 0 and t.max() == 0, (name, n, t)
