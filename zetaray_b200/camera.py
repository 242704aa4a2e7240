"""The default camera of the reference and its per-frame constants (cbFrameConstants) for headless runs: bench.py, smoke and the
tests all render with this producer (the reference's is Common::UpdateFrameConstants, DefaultRenderer.cpp:31-125)."""
import numpy as np

from ._lib import FrameConstants


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



def halton(i, b):
    f = np.float32(1.0); r = np.float32(0.0); bf = np.float32(b)
    while i > 0:
        f = np.float32(f / bf)
        r = np.float32(r + f * np.float32(i % b))
        i = int(np.float32(i) / bf)
    return r


class FrameSequence:
    """cbFrameConstants for consecutive frames of a static camera: jitter = Halton(2,3) - 0.5
    over an 8-phase cycle, prev* = last frame's curr*."""

    def __init__(self, w, h, jitter=True, first_frame=1, cam_path=None, accumulate=False):
        """cam_path(frame) -> camera position (a translating camera); None = the static default camera.
        accumulate: Accumulate + CameraStatic with NumFramesCameraStatic counting up (the reference's accumulation mode)."""
        self.w, self.h, self.jitter = w, h, jitter
        self.frame = first_frame - 1
        self.prev_jitter = (0.0, 0.0)
        self.cam_path, self.accumulate = cam_path, accumulate
        self.prev_cam = None
        self.static_frames = 0

    def next(self):
        self.frame += 1
        j = (0.0, 0.0)
        if self.jitter:
            ph = self.frame % 8
            j = (float(halton(ph + 1, 2) - np.float32(0.5)), float(halton(ph + 1, 3) - np.float32(0.5)))
        if self.cam_path is None:
            fc = look_at_frame_constants(self.w, self.h, frame=self.frame, jitter=j, prev_jitter=self.prev_jitter)
        else:
            cam = tuple(float(np.float32(c)) for c in self.cam_path(self.frame))
            fc = look_at_frame_constants(self.w, self.h, frame=self.frame, jitter=j, prev_jitter=self.prev_jitter, cam=cam,
                                               prev_cam=self.prev_cam or cam)
            self.prev_cam = cam
        if self.accumulate:
            self.static_frames += 1
            fc.Accumulate, fc.CameraStatic, fc.NumFramesCameraStatic = 1, 1, self.static_frames
        self.prev_jitter = j
        return fc
