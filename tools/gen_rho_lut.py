"""Generates zetaray_b200/assets/rho_lut.bin: directional albedo of the single-scattering GGX
dielectric reflection lobe, the table GGXReflectance_Dielectric looks up
(ZetaRenderPass/Common/BSDF.hlsli:279-296; the reference ships it as Assets/LUT/rho.dds,
64 (n.wo) x 32 (alpha) x 16 (eta), R16_UNORM). This script does NOT read the reference asset: the
table is re-derived by numerical integration

    rho(mu_o, alpha, eta) = E_{wh ~ VNDF}[ F_dielectric(wh.wo, eta) * G2(wi, wo) / G1(wo) ]

with the visible-normal sampler of BSDF.hlsli:418-438 on a 64 x 64 stratified grid. When the
reference tree is present, `--compare` prints the max abs difference to the reference's table."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "zetaray_b200", "assets", "rho_lut.bin")
NX, NY, NZ = 64, 32, 16


def fresnel(cos_i, eta):
    # eta = n_t / n_i (ShadingData.eta); BSDF.hlsli:735-760
    er = 1.0 / eta
    sin2 = np.clip(1.0 - cos_i * cos_i, 0.0, 1.0)
    cos_t2 = 1.0 - er * er * sin2
    tir = cos_t2 <= 0
    cos_t = np.sqrt(np.where(tir, 1.0, cos_t2))
    rp = (cos_i - er * cos_t) / (cos_i + er * cos_t)
    rs = (er * cos_i - cos_t) / (er * cos_i + cos_t)
    return np.where(tir, 1.0, 0.5 * (rp * rp + rs * rs))


def g1(a2, c):
    c2 = c * c
    t2 = (1.0 - c2) / c2
    return 2.0 / (np.sqrt(1.0 + a2 * t2) + 1.0)


def rho(mu, alpha, eta, n=64):
    u1, u2 = np.meshgrid((np.arange(n) + 0.5) / n, (np.arange(n) + 0.5) / n)
    u1 = u1.ravel(); u2 = u2.ravel()
    wo = np.array([np.sqrt(max(0.0, 1 - mu * mu)), 0.0, mu])
    vh = np.array([alpha * wo[0], alpha * wo[1], wo[2]]); vh /= np.linalg.norm(vh)
    phi = 2 * np.pi * u1
    z = (1 - u2) * (1 + vh[2]) - vh[2]
    st = np.sqrt(np.clip(1 - z * z, 0, 1))
    c = np.stack([st * np.cos(phi), st * np.sin(phi), z], axis=1)
    nh = c + vh
    ne = np.stack([alpha * nh[:, 0], alpha * nh[:, 1], np.maximum(0.0, nh[:, 2])], axis=1)
    ne /= np.linalg.norm(ne, axis=1, keepdims=True)
    wh_wo = np.clip(ne @ wo, 0, 1)
    wi = 2 * wh_wo[:, None] * ne - wo
    ci = wi[:, 2]
    valid = ci > 0
    a2 = alpha * alpha
    ci_s = np.where(valid, ci, 1.0)
    g1i = g1(a2, ci_s); g1o = g1(a2, max(mu, 1e-5))
    g2_over_g1 = g1i / (g1i + g1o - g1i * g1o)
    f = fresnel(wh_wo, eta)
    return float(np.mean(np.where(valid, f * g2_over_g1, 0.0)))


def main():
    lut = np.zeros((NZ, NY, NX), dtype=np.float64)
    for k in range(NZ):
        eta = 0.5 + k / (NZ - 1) * 1.49
        for j in range(NY):
            # axis conventions fitted against the reference table's values (not its bytes): n.wo at
            # texel centres, roughness (alpha = r^2) and eta at texel edges
            r = 0.045 + j / (NY - 1) * (1.0 - 0.045)
            alpha = r * r
            for i in range(NX):
                lut[k, j, i] = rho((i + 0.5) / NX, alpha, eta)
    q = np.floor(np.clip(lut, 0, 1) * 65535 + 0.5).astype(np.uint16)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    q.tofile(OUT)
    print("wrote", OUT, q.shape)
    if "--compare" in sys.argv:
        ref = "/root/reference/Assets/LUT/rho.dds"
        if os.path.exists(ref):
            r = np.frombuffer(open(ref, "rb").read()[128:], dtype=np.uint16).reshape(NZ, NY, NX).astype(np.float64) / 65535
            d = np.abs(r - q / 65535.0)
            print("vs reference table: max abs diff %.4f, mean %.5f" % (d.max(), d.mean()))
            print("per-eta-slice max:", np.round(d.reshape(NZ, -1).max(axis=1), 4))


if __name__ == "__main__":
    main()
