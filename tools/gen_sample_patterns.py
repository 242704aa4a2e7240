"""Generates the rotated-disk sample tables used by the spatial reuse passes:
  zetaray_b200/assets/disk512.bin  512 x float2 in the unit disk   (role of k_samples in
                                   IndirectLighting/ReSTIR_PT/SampleSet.hlsli:6-523)
  zetaray_b200/assets/disk32.bin   32 x float2 in [0,1]^2          (role of k_samples in
                                   DirectLighting/Emissive/Resampling.hlsli:352-386)
The reference's literal tables are data in its shader sources and are not copied; these are
re-generated low-discrepancy sets with the same role, storage precision (binary16) and value range.
Values are written as float32 already rounded to binary16."""
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "zetaray_b200", "assets")


def r2(n):
    g = 1.32471795724474602596
    a = np.array([1.0 / g, 1.0 / (g * g)])
    return (0.5 + np.arange(1, n + 1)[:, None] * a[None, :]) % 1.0


def concentric(u):
    a = 2 * u[:, 0] - 1
    b = 2 * u[:, 1] - 1
    r = np.where(np.abs(a) > np.abs(b), a, b)
    phi = np.where(np.abs(a) > np.abs(b), (np.pi / 4) * (b / np.where(a == 0, 1, a)),
                   np.pi / 2 - (np.pi / 4) * (a / np.where(b == 0, 1, b)))
    return np.stack([r * np.cos(phi), r * np.sin(phi)], axis=1)


def main():
    os.makedirs(OUT, exist_ok=True)
    d = concentric(r2(512))
    # keep samples off the centre so a rotated tap rarely lands on the pixel itself
    rad = np.linalg.norm(d, axis=1, keepdims=True)
    d = d / np.maximum(rad, 1e-6) * (0.15 + 0.85 * rad)
    d.astype(np.float16).astype(np.float32).tofile(os.path.join(OUT, "disk512.bin"))
    s = r2(32)
    s.astype(np.float16).astype(np.float32).tofile(os.path.join(OUT, "disk32.bin"))
    print("wrote disk512.bin, disk32.bin")


if __name__ == "__main__":
    main()
