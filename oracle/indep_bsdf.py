"""ORACLE -- test infrastructure, not product code.

Independent float64 evaluator of the reference's layered surface shader (`BSDF::Unified`, ZetaRenderPass/Common/BSDF.hlsli
:1176-1266 with ShadingData::Init :584-638 and SetWi :695-718). It was written from the MODEL the reference implements --
a clear-coat slab over {metal | translucent dielectric | glossy dielectric over an energy-preserving Oren-Nayar diffuse
base} -- using the text-book forms of the ingredients (Walter et al. 2007 microfacet BRDF / BTDF, Heitz 2014 Smith Lambda,
unpolarised Fresnel from Snell's law, Portsmouth et al. 2024 "EON" with its analytic constants), NOT by transcribing
`oracle/orc_bsdf.h` or the HLSL statement by statement: the algebra differs (Lambda form of G2 instead of the reference's
folded `G2_Opt`, n_i / n_t Fresnel instead of the relative-index mad chains, analytic EON constants instead of the decimal
literals), the precision differs (float64, vectorised numpy), and the only shared numbers are model parameters
(thresholds, the E_FON fit, the directional-albedo table). A transcription error common to `orc_bsdf.h` and
`zr_bsdf.cuh` (which were produced from the same restatement) shows up here as a disagreement.

`unified(desc arrays..., wi)` returns (f, tolerance_scale, near_threshold): f with n.wi folded in like the reference,
and a per-sample flag for inputs within float32 rounding of a discrete decision (delta lobes, validity tests, TIR),
where a float32 and a float64 evaluation may legitimately take different branches."""
import numpy as np

PI = np.pi
MIN_N_DOT_H_SPECULAR = 0.99998          # BSDF.hlsli:31
MAX_ALPHA_SPECULAR = 0.0016             # BSDF.hlsli:35
ETA_AIR = 1.0
# EON constants, analytic (Portsmouth, Kutz, Hill 2024, eqs. for A, E_avg): the reference carries them as decimals
EON_A_COEFF = 0.5 - 2.0 / (3.0 * PI)            # 0.287793398
EON_AVG_COEFF = 2.0 / 3.0 - 28.0 / (15.0 * PI)  # 0.0724882111


def _dot(a, b):
    return np.sum(a * b, axis=-1)


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def _sat(x):
    return np.clip(x, 0.0, 1.0)


def ggx_d(cos_h, a2):
    return a2 / (PI * ((cos_h * cos_h) * (a2 - 1.0) + 1.0) ** 2)


def smith_lambda(cos_t, a2):
    c2 = cos_t * cos_t
    return 0.5 * (np.sqrt(1.0 + a2 * (1.0 - c2) / c2) - 1.0)


def fresnel_unpolarised(cos_i, n_t):
    """Air-normalised interface n_i = 1 -> n_t; returns (F, tir, cos_t)."""
    sin2_t = (1.0 - cos_i * cos_i) / (n_t * n_t)
    tir = sin2_t >= 1.0
    cos_t = np.sqrt(np.where(tir, 0.0, 1.0 - sin2_t))
    with np.errstate(divide="ignore", invalid="ignore"):
        r_s = (cos_i - n_t * cos_t) / (cos_i + n_t * cos_t)
        r_p = (n_t * cos_i - cos_t) / (n_t * cos_i + cos_t)
    return np.where(tir, 1.0, 0.5 * (r_s * r_s + r_p * r_p)), tir, cos_t


def schlick(f0, c):
    return f0 + (1.0 - f0) * (1.0 - c) ** 5


def e_fon(mu, r):
    """Directional albedo of the single-scatter FON lobe, the paper's rational fit (model parameters)."""
    m = 1.0 - mu
    g_over_pi = m * (0.0571085289 + m * (0.491881867 + m * (-0.332181442 + m * 0.0714429953)))
    return (1.0 + r * g_over_pi) / (1.0 + EON_A_COEFF * r)


class RhoTable:
    """Assets/LUT/rho.dds: 64 (n.wo) x 32 (alpha) x 16 (eta) UNORM16, sampled trilinear with clamp addressing."""
    def __init__(self, u16):
        self.v = np.asarray(u16, dtype=np.float64).reshape(16, 32, 64) / 65535.0

    def sample(self, alpha, ndotwo, eta):
        u = ndotwo
        v = (alpha - 0.002025) / (1.0 - 0.002025)
        w = (eta - 0.5) / (1.99 - 0.5)
        out = 0.0
        px, py, pz = u * 64 - 0.5, v * 32 - 0.5, w * 16 - 0.5
        x0, y0, z0 = np.floor(px), np.floor(py), np.floor(pz)
        fx, fy, fz = px - x0, py - y0, pz - z0
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    xi = np.clip(x0 + dx, 0, 63).astype(int); yi = np.clip(y0 + dy, 0, 31).astype(int); zi = np.clip(z0 + dz, 0, 15).astype(int)
                    wgt = (fx if dx else 1 - fx) * (fy if dy else 1 - fy) * (fz if dz else 1 - fz)
                    out = out + wgt * self.v[zi, yi, xi]
        return _sat(out)


def unified(rho, n, wo, wi, metallic, roughness, base, eta_curr, eta_next, spec_tr, tr_depth, subsurface, coat_w, coat_col,
            coat_rough, coat_ior):
    """All arguments are arrays over samples (vectors as [N,3]); returns f [N,3] (n.wi included) and near_threshold [N]."""
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    n, wo, wi, base, coat_col = map(f64, (n, wo, wi, base, coat_col))
    roughness, eta_curr, eta_next, tr_depth, subsurface, coat_w, coat_rough, coat_ior = map(f64, (roughness, eta_curr, eta_next,
                                                                                                   tr_depth, subsurface, coat_w, coat_rough, coat_ior))
    metallic = np.asarray(metallic, dtype=bool); spec_tr = np.asarray(spec_tr, dtype=bool)
    # storage precision of the two `half` material parameters (ShadingData::Init signature, BSDF.hlsli:586-587)
    tr_depth = tr_depth.astype(np.float16).astype(np.float64); subsurface = subsurface.astype(np.float16).astype(np.float64)
    N = n.shape[0]
    near = np.zeros(N, dtype=bool)

    def decide(x, thr, ge=True, eps=4e-6):
        """x >= thr (or x > thr), flagging samples a float32 evaluation could decide differently."""
        nonlocal near
        near |= np.abs(x - thr) <= eps * np.maximum(1.0, np.abs(thr))
        return (x >= thr) if ge else (x > thr)

    # ---- surface set-up -------------------------------------------------------------------------------------------
    coated = coat_w != 0
    rough_in = roughness.copy()
    roughen = (coat_w > 0) & (coat_rough > 0)
    r_coated = np.minimum(rough_in ** 4 + 2.0 * coat_rough ** 4, 1.0) ** 0.25
    roughness = np.where(roughen, rough_in + (r_coated - rough_in) * coat_w, rough_in)
    alpha = roughness * roughness
    coat_alpha = coat_rough * coat_rough
    in_air = eta_curr == ETA_AIR
    ior_base = np.where(in_air, eta_next, eta_curr)
    eta_plain = eta_next / eta_curr
    eta_under_coat = np.maximum(ior_base, coat_ior) / np.minimum(ior_base, coat_ior)
    eta = eta_plain + (eta_under_coat - eta_plain) * coat_w
    coat_eta = np.where(in_air, coat_ior, 1.0 / coat_ior)
    gloss_delta = alpha <= MAX_ALPHA_SPECULAR
    coat_delta = coat_alpha <= MAX_ALPHA_SPECULAR
    near |= np.abs(alpha - MAX_ALPHA_SPECULAR) < 1e-8
    thin = subsurface > 0
    transmissive = spec_tr | thin

    cos_o_raw = _dot(n, wo)
    back_o = ~decide(cos_o_raw, 0.0, ge=False)
    cos_o = np.maximum(cos_o_raw, 1e-5)
    cos_i_raw = _dot(n, wi)
    refl = decide(cos_i_raw, 0.0, ge=True)
    s = np.where(refl, 1.0, eta)
    h = _unit(wo + wi * s[:, None])
    h = np.where((~refl & (eta > 1.0))[:, None], -h, h)
    cos_h = _sat(_dot(n, h))
    h_o = _sat(_dot(h, wo))
    h_i = np.abs(_dot(h, wi))
    cos_i = np.maximum(np.abs(cos_i_raw), 1e-5)
    o_i = _dot(wo, wi)
    bad_h = spec_tr & ((cos_h == 0) | (h_o == 0))
    near |= spec_tr & ((np.abs(_dot(n, h)) < 4e-6) | (np.abs(_dot(h, wo)) < 4e-6))
    invalid = back_o | bad_h | (refl & (cos_i_raw <= 0)) | (~refl & (~transmissive | metallic))

    f = np.zeros((N, 3))
    live = ~invalid

    # ---- coat slab ----------------------------------------------------------------------------------------------------
    F0_coat = ((coat_eta - 1.0) / (coat_eta + 1.0)) ** 2
    sin2_t_coat = (1.0 - h_o * h_o) / (coat_eta * coat_eta)
    coat_tir = sin2_t_coat >= 1.0
    near |= coated & (np.abs(sin2_t_coat - 1.0) < 4e-6)
    cos_t_coat = np.sqrt(np.where(coat_tir, 0.0, 1.0 - sin2_t_coat))
    F_coat = np.where(coat_tir, 1.0, schlick(F0_coat, np.where(coat_eta > 1.0, h_o, cos_t_coat)))

    def microfacet_refl(a, delta, F):
        """D F G2 / (4 cos_o) (x cos_i folded in), or the delta-lobe convention F * [n.h >= threshold]."""
        a2 = a * a
        with np.errstate(divide="ignore", invalid="ignore"):
            g2 = 1.0 / (1.0 + smith_lambda(cos_i, a2) + smith_lambda(cos_o, a2))
            rough = ggx_d(cos_h, a2) * g2 / (4.0 * cos_o)
        on_peak = decide(cos_h, MIN_N_DOT_H_SPECULAR) if np.any(delta) else np.zeros(N, dtype=bool)
        val = np.where(delta, on_peak.astype(np.float64), rough)
        return val[:, None] * (F if F.ndim == 2 else F[:, None])

    coat_lobe = coat_w[:, None] * microfacet_refl(coat_alpha, coat_delta, F_coat)[:, :1] * np.ones((1, 3))
    dead_coat = coated & ~refl & coat_tir
    live &= ~dead_coat
    use_coat = coated & refl & live
    f = np.where(use_coat[:, None], coat_lobe, f)
    live &= ~(coated & coat_tir)              # everything under a totally reflecting coat is dark
    refl_coat = np.where(coat_delta, F_coat, rho.sample(coat_alpha, cos_o, coat_eta))
    with np.errstate(divide="ignore", invalid="ignore"):
        coat_tr = np.where(coat_col > 0, np.power(np.maximum(coat_col, 1e-300), (1.0 / np.maximum(cos_t_coat, 1e-300))[:, None]), 0.0)
    w_base = np.where(coated[:, None], 1.0 + ((1.0 - refl_coat)[:, None] * coat_tr - 1.0) * coat_w[:, None], 1.0)

    # ---- base: Fresnel of the glossy interface ------------------------------------------------------------------------------
    F_diel, tir, _ = fresnel_unpolarised(h_o, eta)
    near |= ~metallic & (np.abs((1.0 - h_o * h_o) / (eta * eta) - 1.0) < 4e-6)
    F_g = np.where(metallic[:, None], schlick(base, h_o[:, None]), F_diel[:, None])
    tir = tir & ~metallic
    gloss = microfacet_refl(alpha, gloss_delta, F_g)

    metal_like = metallic | tir
    f = np.where((live & metal_like)[:, None], f + w_base * gloss, f)
    live &= ~metal_like

    refl_g = np.where(gloss_delta, F_g[:, 0], rho.sample(alpha, cos_o, eta))

    # opaque (possibly thin-walled) dielectric: EON diffuse under the glossy interface
    sigma = np.sqrt(alpha)
    A = 1.0 / (1.0 + EON_A_COEFF * sigma)
    B = sigma * A
    s_t = o_i - cos_i * cos_o
    s_t = np.where(s_t > 0, s_t / np.maximum(cos_i, cos_o), s_t)
    f_ss = (A + B * s_t) / PI
    E_avg = A + EON_AVG_COEFF * B
    E_o = e_fon(np.maximum(cos_o_raw, 1e-4), roughness)
    E_i = e_fon(cos_i, sigma)
    rho_ms = base * base * E_avg[:, None] / (1.0 - base * (1.0 - E_avg)[:, None])
    f_ms = rho_ms / PI * ((1.0 - E_o) * (1.0 - E_i) / (1.0 - E_avg))[:, None]
    eon = cos_i[:, None] * (base * f_ss[:, None] + f_ms)
    lambert = cos_i[:, None] * base / PI
    diffuse = np.where((sigma == 0)[:, None], lambert, eon)
    diffuse = diffuse * np.where(subsurface == 0, 1.0, 0.5 * subsurface)[:, None]
    opaque = live & ~spec_tr
    f = np.where(opaque[:, None], f + w_base * ((1.0 - refl_g)[:, None] * diffuse + gloss * refl[:, None]), f)
    live &= spec_tr

    # translucent base: reflection lobe, or Walter's BTDF
    f = np.where((live & refl)[:, None], f + gloss * w_base, f)
    live &= ~refl
    a2 = alpha * alpha
    with np.errstate(divide="ignore", invalid="ignore"):
        g2 = 1.0 / (1.0 + smith_lambda(cos_i, a2) + smith_lambda(cos_o, a2))
        denom = (h_i + h_o / eta) ** 2
        btdf_cos = ggx_d(cos_h, a2) * g2 * h_i * h_o / (cos_o * denom)       # (x cos_i) / cos_i cancelled
        btdf_cos = np.where(denom > 0, btdf_cos, 0.0)
    on_peak_t = decide(cos_h, MIN_N_DOT_H_SPECULAR) if np.any(gloss_delta & live) else np.zeros(N, dtype=bool)
    tr_lobe = np.where(gloss_delta, on_peak_t.astype(np.float64), btdf_cos) * (1.0 - F_g[:, 0])
    tint = np.where((tr_depth > 0)[:, None], 1.0, base)
    refl_g_t = np.where(gloss_delta, 0.0, refl_g)
    f = np.where(live[:, None], ((1.0 - refl_g_t) * tr_lobe)[:, None] * tint * w_base, f)
    return f, near
