// ORACLE -- test infrastructure, not product code (see orc_math.h header).
// Per-pixel reconstruction from the packed G-buffer shared by the lighting restatements (what every
// reference kernel does in its first ~60 lines, e.g. ReSTIR_PT_PathTrace.hlsl:441-524).
#pragma once
#include "orc_scene.h"
#include <functional>
#include <thread>
#include <vector>

namespace orc
{
struct Frame
{
    const Scene* sc;
    const zr_frame_constants* fc;
    const uint4* core; const uint2* me; const uint2* coat;          // current G-buffer
    const uint4* pcore; const uint2* pcoat;                          // previous G-buffer
    uint32_t W, H;
};

// Everything the kernels reconstruct per pixel from the G-buffer
struct Pixel
{
    GFlags flags; float roughness; float z; float3 pos, normal, origin; float2 lensSample;
    BSDF::ShadingData surface; float eta_next;
};

inline float3 row3(const float m[3][4], int r) { return f3(m[r][0], m[r][1], m[r][2]); }

// prev == false: current frame camera / jitter; true: previous frame's
inline Pixel LoadPixel(const Frame& f, const uint4* core, const uint2* coat, int px, int py, bool prev, int coatX, int coatY)
{
    const zr_frame_constants& fc = *f.fc;
    Pixel p;
    const size_t idx = (size_t)py * f.W + px;
    GCore g = LoadCore(core, idx);
    p.flags = DecodeFlags(g.flagsByte);
    p.roughness = g.roughness;
    p.z = g.depth;
    p.lensSample = f2(0, 0);
    p.origin = prev ? f3(fc.PrevViewInv[0][3], fc.PrevViewInv[1][3], fc.PrevViewInv[2][3]) : f3(fc.CameraPos[0], fc.CameraPos[1], fc.CameraPos[2]);
    if (fc.DoF)
    {
        uint3 h = RNG::PCG3d(uint3{ (uint32_t)px, (uint32_t)py, (uint32_t)px });
        RNG rngDoF = RNG::Init(h.z, h.y, prev ? fc.FrameNum - 1 : fc.FrameNum);
        p.lensSample = Sampling::UniformSampleDiskConcentric(rngDoF.Uniform2D());
        p.lensSample = p.lensSample * fc.LensRadius;
    }
    const float2 renderDim = f2((float)f.W, (float)f.H);
    const float (*V)[4] = prev ? fc.PrevView : fc.CurrView;
    const float2 jitter = prev ? f2(fc.PrevCameraJitter[0], fc.PrevCameraJitter[1]) : f2(fc.CurrCameraJitter[0], fc.CurrCameraJitter[1]);
    p.pos = Math::WorldPosFromScreenSpace2(f2((float)px, (float)py), renderDim, p.z, fc.TanHalfFOV, fc.AspectRatio, jitter,
        row3(V, 0), row3(V, 1), row3(V, 2), fc.DoF != 0, p.lensSample, fc.FocusDepth, p.origin);
    p.normal = Math::DecodeUnitVector(g.normalEnc);
    const float4 baseColor = p.flags.subsurface ? g.baseColor : f4(g.baseColor.x, g.baseColor.y, g.baseColor.z, 0);
    p.eta_next = BSDF::DEFAULT_ETA_MAT;
    if (p.flags.transmissive)
        p.eta_next = DecodeIOR(g.iorEnc);
    float coat_weight = 0; float3 coat_color = f3(0.0f); float coat_roughness = 0; float coat_ior = BSDF::DEFAULT_ETA_COAT;
    if (p.flags.coated)
    {
        Coat c = UnpackCoat(LoadCoat(coat, (size_t)coatY * f.W + coatX));
        coat_weight = c.weight; coat_color = c.color; coat_roughness = c.roughness; coat_ior = c.ior;
    }
    const float3 wo = normalize(p.origin - p.pos);
    p.surface = BSDF::ShadingData::Init(p.normal, wo, p.flags.metallic, p.roughness, f3(baseColor.x, baseColor.y, baseColor.z),
        BSDF::ETA_AIR, p.eta_next, p.flags.transmissive, p.flags.trDepthGt0 ? 1.0f : 0.0f, to_half(baseColor.w),
        coat_weight, coat_color, coat_roughness, coat_ior);
    return p;
}

inline GFlags FlagsAt(const uint4* core, uint32_t W, int x, int y, float* roughness = nullptr)
{
    uint32_t w = core[(size_t)y * W + x].w;
    if (roughness) *roughness = (float)((w >> 8) & 0xff) / 255.0f;
    return DecodeFlags(w & 0xff);
}

inline void parallel_for(uint32_t n, int nthreads, const std::function<void(uint32_t, uint32_t)>& fn)
{
    if (nthreads <= 1 || n < 2) { fn(0, n); return; }
    std::vector<std::thread> th;
    uint32_t per = (n + nthreads - 1) / nthreads;
    for (int i = 0; i < nthreads; i++)
    {
        uint32_t a = i * per, b = a + per > n ? n : a + per;
        if (a >= b) break;
        th.emplace_back(fn, a, b);
    }
    for (auto& t : th) t.join();
}

} // namespace orc
