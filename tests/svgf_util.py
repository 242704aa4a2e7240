"""Frame loop of the SVGF oracle (oracle/orc_svgf.cpp), shaped like zr_svgf_pass: temporal stage + a-trous passes with ping-pong planes."""
import ctypes as C
import os
import numpy as np

from tests import orc
from tests.orc import ptr


class OracleSVGF:
    def __init__(self, w, h, sigma_z=0.02, k_n=16.0, sigma_l=4.0, radius=2, num_passes=5, nthreads=None):
        self.o = orc.load()
        self.w, self.h = w, h
        self.prm = np.array([sigma_z, k_n, sigma_l, float(radius)], dtype=np.float32)
        self.num_passes = num_passes
        self.nthreads = nthreads or (8 if w * h <= 512 * 512 else max(8, min(os.cpu_count() or 8, 128)))
        n = w * h
        self.guide = [np.zeros((n, 2), dtype=np.uint32), np.zeros((n, 2), dtype=np.uint32)]
        self.hist = [np.zeros((n, 4), dtype=np.uint32), np.zeros((n, 4), dtype=np.uint32)]
        self.cur = 0
        self.valid = False

    def render(self, fc, core, me, signal):
        """core: uint32[n,4], me: uint32[n,2], signal: float32[n,4]. Returns (denoised float32[n,4], accumulated cv uint32[n,2])."""
        n = self.w * self.h
        self.cur = 1 - self.cur
        cv = [np.zeros((n, 2), dtype=np.uint32), np.zeros((n, 2), dtype=np.uint32)]
        core = np.ascontiguousarray(core); me = np.ascontiguousarray(me); signal = np.ascontiguousarray(signal, dtype=np.float32)
        self.o.orc_svgf_temporal(C.byref(fc), ptr(core), ptr(me), ptr(signal), ptr(self.guide[1 - self.cur]), ptr(self.hist[1 - self.cur]),
                                 1 if self.valid else 0, ptr(self.hist[self.cur]), ptr(cv[0]), ptr(self.guide[self.cur]), self.nthreads)
        accumulated = cv[0].copy()
        out = np.zeros((n, 4), dtype=np.float32)
        plane = 0
        for k in range(self.num_passes):
            last = k + 1 == self.num_passes
            self.o.orc_svgf_atrous(self.w, self.h, ptr(self.guide[self.cur]), ptr(cv[plane]), ptr(cv[1 - plane]), ptr(out) if last else None,
                                   1 << k, ptr(self.prm), self.nthreads)
            plane = 1 - plane
        self.valid = True
        return out, accumulated
