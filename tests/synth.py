"""Seeded synthetic per-pixel state for kernel-level parity tests (SURVEY 8d 'synthetic-frame mode')."""
import ctypes as C
import numpy as np

from zetaray_b200._lib import FrameConstants

FLT_MAX = np.float32(3.402823466e+38)


from zetaray_b200.camera import look_at_frame_constants  # noqa: E402,F401  (moved into the package)


def synth_gbuffer(w, h, seed, miss_frac=0.1, emissive_frac=0.03):
    """Random but plausible G-buffer planes: core uint4, depth f32, motion_emissive uint2."""
    rng = np.random.default_rng(seed)
    n = w * h
    depth = (1.0 + 9.0 * rng.random(n, dtype=np.float32)).astype(np.float32)
    miss = rng.random(n) < miss_frac
    depth[miss] = FLT_MAX
    flags = np.zeros(n, dtype=np.uint32)
    flags[miss] = 4
    em = (~miss) & (rng.random(n) < emissive_frac)
    flags[em] |= 2
    rough = rng.choice(np.array([26, 77, 153, 255], dtype=np.uint32), size=n)
    normal = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    base = rng.integers(0, 1 << 24, size=n, dtype=np.uint64).astype(np.uint32)
    core = np.zeros((n, 4), dtype=np.uint32)
    core[:, 0] = depth.view(np.uint32)
    core[:, 1] = normal
    core[:, 2] = base
    core[:, 3] = flags | (rough << 8)
    # small motion vectors (snorm16 x2) -- a few px
    mv = rng.integers(-40, 41, size=(n, 2)).astype(np.int16).view(np.uint16).astype(np.uint32)
    me = np.zeros((n, 2), dtype=np.uint32)
    me[:, 0] = mv[:, 0] | (mv[:, 1] << 16)
    return core, depth.copy(), me


def synth_hdr(w, h, seed, fireflies=True):
    rng = np.random.default_rng(seed)
    img = np.zeros((w * h, 4), dtype=np.float32)
    img[:, :3] = rng.random((w * h, 3), dtype=np.float32) * np.float32(2.0)
    if fireflies:
        k = rng.integers(0, w * h, size=max(1, w * h // 50))
        img[k, :3] *= np.float32(200.0)
    return img
