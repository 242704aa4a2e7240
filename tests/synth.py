"""Seeded synthetic per-pixel state for kernel-level parity tests (SURVEY 8d 'synthetic-frame mode')."""
import ctypes as C
import numpy as np

from zetaray_b200._lib import FrameConstants

FLT_MAX = np.float32(3.402823466e+38)


def look_at_frame_constants(w, h, frame=1, jitter=(0.0, 0.0), prev_jitter=(0.0, 0.0), cam=(0.0, 1.2, -4.043), prev_cam=None):
    """cbFrameConstants for the default camera: left-handed, +Z forward, vfov 60 deg.
    prev_cam: last frame's camera position (a translating camera); defaults to cam (static)."""
    fc = FrameConstants()
    pc = cam if prev_cam is None else prev_cam
    view = np.array([[1, 0, 0, -cam[0]], [0, 1, 0, -cam[1]], [0, 0, 1, -cam[2]]], dtype=np.float32)
    inv = np.array([[1, 0, 0, cam[0]], [0, 1, 0, cam[1]], [0, 0, 1, cam[2]]], dtype=np.float32)
    pview = np.array([[1, 0, 0, -pc[0]], [0, 1, 0, -pc[1]], [0, 0, 1, -pc[2]]], dtype=np.float32)
    pinv = np.array([[1, 0, 0, pc[0]], [0, 1, 0, pc[1]], [0, 0, 1, pc[2]]], dtype=np.float32)
    for name, m in (("CurrView", view), ("PrevView", pview), ("CurrViewInv", inv), ("PrevViewInv", pinv)):
        arr = getattr(fc, name)
        for i, v in enumerate(m.reshape(-1)):
            arr[i] = float(v)
    fc.CameraPos[0], fc.CameraPos[1], fc.CameraPos[2] = cam
    fc.CameraNear = 0.2
    fc.AspectRatio = np.float32(w) / np.float32(h)
    fc.TanHalfFOV = float(np.tan(np.float32(0.5) * np.float32(np.pi / 3)).astype(np.float32))
    fc.PixelSpreadAngle = float(np.arctan(np.float32(2 * fc.TanHalfFOV / h)))
    fc.FrameNum = frame
    fc.RenderWidth, fc.RenderHeight, fc.DisplayWidth, fc.DisplayHeight = w, h, w, h
    fc.CurrCameraJitter[0], fc.CurrCameraJitter[1] = jitter
    fc.PrevCameraJitter[0], fc.PrevCameraJitter[1] = prev_jitter
    fc.CameraRayUVGradsScale = 1.0
    fc.NumFramesCameraStatic = 0
    fc.CameraStatic = 0
    fc.Accumulate = 0
    return fc


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
