"""CPU tier: the BVH builder + the product's traversal source (host build) against brute force on random triangle soups, including
the inputs a builder tends to get wrong: duplicate triangles (tie rule), all centroids coincident, axis-aligned planes (zero-
thickness boxes), coordinates far from the origin, one huge triangle among tiny ones, a single triangle -- and zero-area
triangles, where only a weaker property holds (see test_zero_area_triangles)."""
import ctypes as C

import numpy as np
import pytest

from tests import hostsim
from tests.orc import ptr
from tests.test_bvh_host import build, THREADS


def soup(kind, n, rng):
    v0 = rng.normal(size=(n, 3)) * 3
    e1 = rng.normal(size=(n, 3)) * 0.4
    e2 = rng.normal(size=(n, 3)) * 0.4
    if kind == "duplicates":
        v0[1::7] = v0[0]; e1[1::7] = e1[0]; e2[1::7] = e2[0]        # copies of triangle 0: equal t -> the lowest index must win
        if n > 2:
            v0[2::5] = v0[2]; e1[2::5] = e1[2]; e2[2::5] = e2[2]
    elif kind == "zero_area":
        e2[::3] = e1[::3] * 2.0             # needles: e2 parallel to e1
    elif kind == "coincident":
        c = rng.normal(size=3)
        v0 = c - (e1 + e2) / 3.0            # every centroid at c
    elif kind == "planar":
        v0[:, 1] = 0.25; e1[:, 1] = 0.0; e2[:, 1] = 0.0             # everything in the plane y = 0.25
    elif kind == "far":
        v0 += np.array([8192.0, -4096.0, 16384.0])
    elif kind == "mixed_scale":
        v0[0] = [-50, -1, -50]; e1[0] = [100, 0, 0]; e2[0] = [0, 0, 100]        # a floor under pebbles
        e1[1:] *= 0.02; e2[1:] *= 0.02
    wt = np.concatenate([v0, e1, e2], axis=1).astype(np.float32)
    return np.ascontiguousarray(wt)


def rays_for(wt, m, rng):
    p = np.concatenate([wt[:, 0:3], wt[:, 0:3] + wt[:, 3:6], wt[:, 0:3] + wt[:, 6:9]])
    lo, hi = p.min(axis=0), p.max(axis=0)
    ext = np.maximum(hi - lo, 1e-3)
    o = lo - 0.2 * ext + 1.4 * ext * rng.random((m, 3))
    # aim at points ON triangles so that most rays hit something
    t = rng.integers(0, len(wt), m)
    b = rng.random((m, 2)); b[b.sum(axis=1) > 1] = 1 - b[b.sum(axis=1) > 1]
    tgt = wt[t, 0:3] + b[:, 0:1] * wt[t, 3:6] + b[:, 1:2] * wt[t, 6:9]
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True) + 1e-30
    rays = np.zeros((m, 8), dtype=np.float32)
    rays[:, 0:3] = o; rays[:, 3] = 1e-6; rays[:, 4:7] = d; rays[:, 7] = 3.0e38
    rays[::5, 7] = np.linalg.norm(tgt - o, axis=1)[::5] * 0.5        # segments that stop short
    k = m // 10
    rays[:k, 4:7] = np.eye(3)[rng.integers(0, 3, k)] * rng.choice([-1.0, 1.0], (k, 1))      # axis-parallel
    return rays


KINDS = ["plain", "duplicates", "coincident", "planar", "far", "mixed_scale"]


@pytest.mark.parametrize("kind", KINDS)
def test_random_soups(kind):
    hs = hostsim.load()
    rng = np.random.default_rng(KINDS.index(kind) + 100)
    for n in (1, 2, 3, 4, 9, 33, 100, 257, 1500):
        wt = soup(kind, n, rng)
        nodes, order, leaf, info = build(wt)
        assert np.array_equal(np.sort(order), np.arange(n))
        assert info[3] <= 96
        stats = (C.c_uint64 * 4)()
        assert hs.hostsim_validate(ptr(nodes), info[0], ptr(leaf), n, stats) == 0, (kind, n)
        m = 600
        rays = rays_for(wt, m, rng)
        tri_mesh = np.zeros(n, dtype=np.uint32); first = np.zeros(1, dtype=np.uint32)
        got = np.zeros((m, 4), dtype=np.float32); ref = np.zeros((m, 4), dtype=np.float32)
        anyf = np.zeros(m, dtype=np.uint32)
        hs.hostsim_trace(ptr(nodes), ptr(leaf), ptr(tri_mesh), ptr(first), ptr(rays), m, ptr(got), ptr(anyf), None, THREADS)
        hs.hostsim_brute(ptr(wt), n, ptr(rays), m, ptr(ref), THREADS)
        assert got.tobytes() == ref.tobytes(), (kind, n, int((got.view(np.uint32) != ref.view(np.uint32)).any(axis=1).sum()))
        assert np.array_equal(anyf != 0, ref[:, 0] < 3.0e38), (kind, n)


def test_zero_area_triangles():
    """Known limit of the hit rule (Moller-Trumbore, accept unless det == 0): on a zero-area triangle the determinant is rounding
    noise instead of 0, so a ray can 'hit' it at a numerically meaningless (t, u, v). Brute force reports such a hit wherever the
    ray is; the BVH only if the ray also crosses the needle's box -- the two can differ, but ONLY in rays whose brute-force answer
    is a zero-area triangle (a copy of which, or nothing nearer, is what the BVH returns). DXR never reports degenerate triangles
    at all; scenes are expected not to contain them (the reference's assets and the procedural scenes do not)."""
    hs = hostsim.load()
    total = bad = 0
    for seed in range(6):
        rng = np.random.default_rng(500 + seed)
        n = [7, 50, 300, 1200, 50, 300][seed]
        wt = soup("zero_area", n, rng)
        needle = np.zeros(n, dtype=bool); needle[::3] = True
        nodes, order, leaf, info = build(wt)
        m = 3000
        rays = rays_for(wt, m, rng)
        tri_mesh = np.zeros(n, dtype=np.uint32); first = np.zeros(1, dtype=np.uint32)
        got = np.zeros((m, 4), dtype=np.float32); ref = np.zeros((m, 4), dtype=np.float32)
        hs.hostsim_trace(ptr(nodes), ptr(leaf), ptr(tri_mesh), ptr(first), ptr(rays), m, ptr(got), None, None, THREADS)
        hs.hostsim_brute(ptr(wt), n, ptr(rays), m, ptr(ref), THREADS)
        diff = np.nonzero((got.view(np.uint32) != ref.view(np.uint32)).any(axis=1))[0]
        total += m; bad += len(diff)
        for i in diff:
            k = int(ref[i, 3:].view(np.uint32)[0])
            assert k < n and needle[k], (seed, i, k)          # the brute-force answer of every differing ray is a needle
            assert not (got[i, 0] < ref[i, 0]), (seed, i)     # and the BVH never reports something nearer than brute force
    assert bad < total // 500
