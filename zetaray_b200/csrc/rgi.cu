// rgi.cu -- IndirectLighting with INTEGRATOR::ReSTIR_GI (emissive NEE): one kernel per frame.
//
// Replaces IndirectLighting/ReSTIR_GI/ReSTIR_GI.hlsl (+ Resampling.hlsli, PathTracing.hlsli, ReSTIR_GI_NEE.hlsli,
// Reservoir.hlsli, ../NEE.hlsli) and the host sequencing of IndirectLighting.cpp:277-368 for the ReSTIR_GI_WoPS / _WPS
// variants: a path-traced initial candidate (second path vertex + its outgoing radiance), temporal reuse with one or two
// reprojected candidates weighted by the reconnection Jacobian, wave-level outlier suppression; NEE after the first indirect
// vertex optionally from the light voxel grid (ReSTIR_GI_LVG). Not built: sun/sky NEE, ray differentials (they only feed texture LOD; no textures in this build), the spatial
// pass (commented out upstream, Resampling.hlsli:603-608).
// A block is 16 consecutive 8x8 groups of the reference's swizzled dispatch, one warp per reference wave, so the
// Russian-roulette WaveActiveMax (PathTracing.hlsli:64-67, evaluated over the lanes at the same loop iteration) is a warp max
// and SuppressOutlierReservoirs (Resampling.hlsli:533-539) a warp sum. Block barriers keep the warps of a block at the
// same stage (zr_rpt.cuh "block-synchronous phases"); this first version synchronises per stage, not inside NEE or the
// temporal resampling.
// D3D semantics kept: FindTemporalCandidate does not reject negative tap coordinates and out-of-bounds texture reads
// return 0, so LoadPrev / the reservoir readers return zeros there.
// Measured on B200 (profiles/r1e_ab_degenerate_ray.json): with the degenerate-ray cut of zr_scene.cuh::Traverse compiled into THIS
// translation unit -- in either formulation, empty stack or early return -- ptxas lays k_rgi out so that warps run the traversal with
// 6.9 instead of 12.4 active lanes (30.0 G instead of 16.7 G warp instructions, 116 ms instead of 60 ms on the 300 k-triangle atrium),
// although the atrium's ReSTIR GI frame contains no degenerate ray at all. The other kernels are unaffected and the path tracer needs
// the cut (3.5 s -> 0.12 s on the tunnel), so only this unit opts out until the reconvergence difference is understood with
// ncu's source view (next round). Cost: a zero-direction ray (2-3 per 14 400 pixels on scenes with glass) sweeps the tree here.
#define ZR_NO_DEGENERATE_RAY_EARLY_OUT
#include "zr_pixel.cuh"
#include "zr_rpt.cuh"       // ZR_PHASE
#include "zr_schedule.h"

namespace zr
{
namespace
{
    struct GIParams
    {
        uint32_t maxNonTrBounces, maxGlossyTrBounces, russianRoulette, stochasticMultiBounce, boilingSuppression, M_max;
        uint32_t temporalResample, resetTemporal;
        uint32_t rowBegin, rowEnd;
    };

    struct GIReservoir
    {
        float3 pos, Lo, normal, target_z;
        float W, w_sum; uint32_t ID; float M;       // M is a half in the reference; small integers are exact
        static ZR_D GIReservoir Init()
        {
            GIReservoir r;
            r.pos = f3(FLT_MAX_); r.normal = f3(0); r.Lo = f3(0); r.M = 0; r.w_sum = 0; r.W = 0; r.ID = UINT32_MAX_; r.target_z = f3(0);
            return r;
        }
        ZR_D bool Update(float weight, float3 vtxPos, float3 vtxNormal, uint32_t vtxID, float3 vtxLo, float3 target, RNG& rng)
        {
            if (weight != weight)
                return false;
            w_sum += weight;
            M += 1;
            if (rng.Uniform() < (weight / fmaxf(1e-6f, w_sum)))
            {
                pos = vtxPos; normal = vtxNormal; ID = vtxID; Lo = vtxLo; target_z = target;
                return true;
            }
            return false;
        }
    };

    // zr_rgi_reservoir (48 bytes): A = {pos.xyz, ID}, B = {Lo.rg (half2), Lo.b | M (half2), w_sum, W}, C = {oct32 normal, 0, 0, 0}
    ZR_D void WriteReservoir(zr_rgi_reservoir& s, const GIReservoir& r, float M_max)
    {
        s.pos[0] = r.pos.x; s.pos[1] = r.pos.y; s.pos[2] = r.pos.z; s.ID = r.ID;
        const float M_clamped = fminf(r.M, M_max);
        s.Lo_rg = (uint32_t)zr_f32_to_f16(r.Lo.x) | ((uint32_t)zr_f32_to_f16(r.Lo.y) << 16);
        s.Lo_b_M = (uint32_t)zr_f32_to_f16(r.Lo.z) | ((uint32_t)zr_f32_to_f16(M_clamped) << 16);
        s.w_sum = r.w_sum; s.W = r.W;
        s.normal = Math::EncodeOct32u(r.normal);
        s.pad[0] = s.pad[1] = s.pad[2] = 0;
    }
    ZR_D const zr_rgi_reservoir* TexelOrNull(const zr_rgi_reservoir* res, uint32_t W, uint32_t H, int x, int y)
    {
        return (x < 0 || y < 0 || x >= (int)W || y >= (int)H) ? nullptr : &res[(size_t)y * W + x];
    }
    ZR_D GIReservoir PartialReadReservoir_Reuse(const zr_rgi_reservoir* res, uint32_t W, uint32_t H, int x, int y)
    {
        const zr_rgi_reservoir* s = TexelOrNull(res, W, H, x, y);
        GIReservoir r;
        r.normal = f3(0); r.w_sum = 0; r.W = 0; r.target_z = f3(0);
        if (!s) { r.pos = f3(0); r.ID = 0; r.Lo = f3(0); r.M = 0; return r; }      // out-of-bounds texture read
        r.pos = f3(s->pos[0], s->pos[1], s->pos[2]); r.ID = s->ID;
        r.Lo = f3(zr_f16_to_f32((uint16_t)(s->Lo_rg & 0xffff)), zr_f16_to_f32((uint16_t)(s->Lo_rg >> 16)), zr_f16_to_f32((uint16_t)(s->Lo_b_M & 0xffff)));
        r.M = (float)(uint16_t)zr_f16_to_f32((uint16_t)(s->Lo_b_M >> 16));
        return r;
    }
    ZR_D void PartialReadReservoir_ReuseRest(const zr_rgi_reservoir* res, uint32_t W, uint32_t H, int x, int y, GIReservoir& r)
    {
        const zr_rgi_reservoir* s = TexelOrNull(res, W, H, x, y);
        r.w_sum = s ? s->w_sum : 0; r.W = s ? s->W : 0;
        r.normal = Math::DecodeOct32(s ? s->normal : 0u);
    }

    // ---- NEE ----
    ZR_D bool IsSpecular(const BSDF::ShadingData& surface)
    {
        return surface.GlossSpecular() && (surface.metallic || surface.specTr) && (!surface.Coated() || surface.CoatSpecular());
    }

    // ReSTIR_GI_NEE.hlsli:8-121 with NumLightSamples = 1. ReSTIR GI: skipDiffuse = true (MIS_NON_DIFFUSE_BSDF_SAMPLING 1),
    // approximate shadow rays; the plain path tracer: skipDiffuse = false, APPROXIMATE_EMISSIVE_SHADOW_RAY 0
    // (PathTracer/Params.hlsli:19-27).
    template<bool SkipDiffuse, bool PreciseShadow>
    ZR_D float3 NEE_Emissive_MIS(const SceneDev& sc, float3 pos, float3 normal, BSDF::ShadingData surface, uint32_t sampleSetIdx, RNG& rng)
    {
        float3 ld = f3(0);
        const bool specular = IsSpecular(surface);
        const int numLightSamples = specular ? 0 : 1;
        {
            BSDF::BSDFSample bsdfSample = SkipDiffuse ? BSDF::SampleBSDF_NoDiffuse(normal, surface, rng) : BSDF::SampleBSDF(normal, surface, rng);
            float3 wi = bsdfSample.wi;
            float3 f = bsdfSample.f;
            float wiPdf = bsdfSample.pdf;
            HitEmissive hitInfo = FindClosestEmissive(sc, pos, normal, wi, surface.Transmissive());
            if (hitInfo.HitWasEmissive())
            {
                const zr_emissive_tri& emissive = sc.emissives[hitInfo.emissiveTriIdx];
                float3 le = Light::Le_EmissiveTriangle(emissive);
                const float3 vtx0 = Light::Vtx0(emissive);
                const float3 vtx1 = Light::DecodeEmissiveTriV1(emissive);
                const float3 vtx2 = Light::DecodeEmissiveTriV2(emissive);
                float3 lightNormal = cross(vtx1 - vtx0, vtx2 - vtx0);
                float twoArea = length(lightNormal);
                twoArea = fmaxf(twoArea, 1e-6f);
                lightNormal = dot(lightNormal, lightNormal) == 0 ? f3(1.0f) : lightNormal / twoArea;
                lightNormal = Light::IsDoubleSided(emissive) && dot(-wi, lightNormal) < 0 ? -lightNormal : lightNormal;
                const float lightSourcePdf = numLightSamples > 0 ? sc.aliasTable[hitInfo.emissiveTriIdx].CachedP_Orig : 0;
                const float lightPdf = lightSourcePdf * (2.0f / twoArea);
                float dwdA = hitInfo.t > 0 ? saturate(dot(lightNormal, -wi)) / (hitInfo.t * hitInfo.t) : 0;
                wiPdf *= dwdA;
                le *= f * dwdA;
                ld = RT::PowerHeuristic(wiPdf, lightPdf, le, 1, (float)numLightSamples);
            }
        }
        for (int s_l = 0; s_l < numLightSamples; s_l++)
        {
            const Light::LightSample lightSample = Light::SampleLight(sc, pos, sampleSetIdx, rng, false);
            float3 le = lightSample.le;
            const float lightPdf = lightSample.pdf;
            const uint32_t lightID = lightSample.ID;
            const float t = length(lightSample.pos - pos);
            const float3 wi = (lightSample.pos - pos) / t;
            if (dot(lightSample.normal, -wi) > 0)
            {
                const float dwdA = saturate(dot(lightSample.normal, -wi)) / (t * t);
                surface.SetWi(wi, normal);
                le *= BSDF::Unified(surface).f * dwdA;
                if (dot(le, le) > 0)
                    le *= (PreciseShadow ? Visibility_Segment_Precise(sc, pos, wi, t, normal, lightID, surface.Transmissive())
                                         : Visibility_Segment(sc, pos, wi, t, normal, lightID, surface.Transmissive())) ? 1.0f : 0.0f;
                float bsdfPdf = SkipDiffuse ? BSDF::BSDFSamplerPdf_NoDiffuse(normal, surface, wi) : BSDF::BSDFSamplerPdf(normal, surface, wi, rng);
                bsdfPdf *= dwdA;
                ld += RT::PowerHeuristic(lightPdf, bsdfPdf, le, (float)numLightSamples);
            }
        }
        return ld;
    }

    // NEE.hlsli:152-221 with NumSamples = 1 (only .ld is consumed by the GI path tracer)
    ZR_D float3 NEE_Emissive_1(const SceneDev& sc, float3 pos, float3 normal, BSDF::ShadingData surface, uint32_t sampleSetIdx, RNG& rng)
    {
        float3 ret = f3(0);
        const Light::LightSample lightSample = Light::SampleLight(sc, pos, sampleSetIdx, rng, false);
        const float3 le = lightSample.le;
        const float lightPdf = lightSample.pdf;
        const float t = length(lightSample.pos - pos);
        const float3 wi = (lightSample.pos - pos) / t;
        if (dot(lightSample.normal, -wi) > 0)
        {
            const float dwdA = saturate(dot(lightSample.normal, -wi)) / (t * t);
            surface.SetWi(wi, normal);
            float3 ld = le * BSDF::Unified(surface).f * dwdA;
            if (Math::Luminance(ld) > 1e-6f)
                ld *= Visibility_Segment(sc, pos, wi, t, normal, lightSample.ID, surface.Transmissive()) ? 1.0f : 0.0f;
            ret += ld / lightPdf;
        }
        ret = ret / 1.0f;
        return ret;
    }

    // ReSTIR_GI_NEE.hlsli:123-193 with numSamples = 1 (the ReSTIR_GI_LVG variant); extents / offset arrive as halves (ReSTIR_GI.hlsl:52-55)
    ZR_D float3 NEE_Emissive_LVG(const SceneDev& sc, const zr_frame_constants& fc, float3 pos, float3 normal, BSDF::ShadingData surface, uint32_t sampleSetIdx, RNG& rng)
    {
        float3 ret = f3(0);
        const float3 extents = f3(to_half(sc.lvgExtents[0]), to_half(sc.lvgExtents[1]), to_half(sc.lvgExtents[2]));
        const float offset_y = to_half(sc.lvgOffsetY);
        LVG::VoxelLight s;
        float3 lightPos, lightNormal, le; float lightPdf; uint32_t lightID;
        if (LVG::Sample(sc, pos, extents, offset_y, fc.CurrView, s, rng))
        {
            lightPos = s.pos; lightNormal = s.normal; le = s.le; lightPdf = s.pdf; lightID = s.ID;
            if (s.twoSided && dot(lightNormal, pos - lightPos) < 0)
                lightNormal = -lightNormal;
        }
        else
        {
            const Light::LightSample ls = Light::SampleLight(sc, pos, sampleSetIdx, rng, false);
            lightPos = ls.pos; lightNormal = ls.normal; le = ls.le; lightPdf = ls.pdf; lightID = ls.ID;
        }
        const float t = length(lightPos - pos);
        const float3 wi = (lightPos - pos) / t;
        if (lightID != UINT32_MAX_ && dot(lightNormal, -wi) > 0)
        {
            const float dwdA = saturate(dot(lightNormal, -wi)) / (t * t);
            surface.SetWi(wi, normal);
            le *= BSDF::Unified(surface).f * dwdA;
            if (Math::Luminance(le) > 1e-6f)
                le *= Visibility_Segment(sc, pos, wi, t, normal, lightID, surface.Transmissive()) ? 1.0f : 0.0f;
            ret += le / fmaxf(lightPdf, 1e-6f);
        }
        ret = ret / 1.0f;
        return ret;
    }

    // PlainPT: the macro set of IndirectLighting/PathTracer/Params.hlsli -- MIS_ALL_BOUNCES 1, MIS_NON_DIFFUSE_BSDF_SAMPLING 0,
    // APPROXIMATE_EMISSIVE_SHADOW_RAY 0 (ReSTIR_GI_NEE.hlsli:225-238)
    template<bool PlainPT>
    ZR_D float3 NEE(const SceneDev& sc, const zr_frame_constants& fc, float3 pos, float3 normal, const BSDF::ShadingData& surface, uint32_t sampleSetIdx, int bounce, RNG& rng)
    {
        if (PlainPT)
            return NEE_Emissive_MIS<false, true>(sc, pos, normal, surface, sampleSetIdx, rng);
        if (bounce == 0)
            return NEE_Emissive_MIS<true, false>(sc, pos, normal, surface, sampleSetIdx, rng);
        if (sc.lvg && sc.sampleSetSize)
            return NEE_Emissive_LVG(sc, fc, pos, normal, surface, sampleSetIdx, rng);
        return NEE_Emissive_1(sc, pos, normal, surface, sampleSetIdx, rng);
    }

    // ---- temporal reuse ----
    struct PrevTexel { float depth; GFlags flags; float roughness; float2 normalEnc; float iorEnc; float3 baseColor; };
    ZR_D PrevTexel LoadPrev(const FrameView& f, int x, int y)
    {
        PrevTexel t;
        if (x < 0 || y < 0 || x >= (int)f.W || y >= (int)f.H)
        {
            t.depth = 0; t.flags = DecodeFlags(0); t.roughness = 0; t.normalEnc = f2(0, 0); t.iorEnc = 0; t.baseColor = f3(0);
            return t;
        }
        const uint4 g = ld128(&f.pcore[(size_t)y * f.W + x]);
        t.depth = asfloat(g.x); t.flags = DecodeFlags(g.w & 0xff); t.roughness = (float)((g.w >> 8) & 0xff) / 255.0f;
        t.normalEnc = Math::DecodeUNorm2(g.y); t.iorEnc = (float)((g.w >> 16) & 0xff) / 255.0f;
        t.baseColor = f3((float)(g.z & 0xff) / 255.0f, (float)((g.z >> 8) & 0xff) / 255.0f, (float)((g.z >> 16) & 0xff) / 255.0f);
        return t;
    }

    struct TemporalSampleData { float3 posW, normal; float roughness; int sx, sy; bool metallic, transmissive; float eta_next; };

    ZR_D bool PlaneHeuristic(float3 samplePos, float3 currNormal, float3 currPos, float linearDepth, float th)
    {
        return fabsf(dot(currNormal, samplePos - currPos)) <= th * linearDepth;
    }

    ZR_D float3 PrevCamPos(const zr_frame_constants& fc) { return f3(fc.PrevViewInv[0][3], fc.PrevViewInv[1][3], fc.PrevViewInv[2][3]); }

    ZR_D int FindTemporalCandidate(const FrameView& f, const SceneDev& sc, int x, int y, float3 posW, float3 normal, float viewZ, float roughness, bool transmissive,
        float2 prevUV, RNG& rng, TemporalSampleData data[2], bool valid[2])
    {
        const zr_frame_constants& fc = f.fc;
        valid[0] = valid[1] = false;
        if (prevUV.x < 0.0f || prevUV.y < 0.0f || prevUV.x > 1.0f || prevUV.y > 1.0f)
            return 0;
        const float2 renderDim = f2((float)f.W, (float)f.H);
        const float2 pp = prevUV * renderDim;
        const int prevPixelX = (int)pp.x, prevPixelY = (int)pp.y;
        int curr = 0;
        const float3 prevCamPos = PrevCamPos(fc);
        for (int i = 0; i < 3; i++)
        {
            const float theta = rng.Uniform() * TWO_PI;
            float sinTheta, cosTheta;
            zr_sincosf(theta, &sinTheta, &cosTheta);
            const float2 offset = f2(16.0f * sinTheta, 16.0f * cosTheta);
            const float m = i > 0 ? 1.0f : 0.0f;
            const int sx = (int)((float)prevPixelX + m * offset.x), sy = (int)((float)prevPixelY + m * offset.y);
            if ((float)sx >= renderDim.x || (float)sy >= renderDim.y)
                continue;
            if (i > 0 && (uint32_t)sx == (uint32_t)x && (uint32_t)sy == (uint32_t)y)
                continue;
            const PrevTexel t = LoadPrev(f, sx, sy);
            if (t.flags.emissive)
                continue;
            float2 lensSample = f2(0, 0);
            float3 origin = prevCamPos;
            if (fc.DoF)
            {
                uint3 h = RNG::PCG3d(make_uint3((uint32_t)sx, (uint32_t)sy, (uint32_t)sx));
                RNG rngDoF = RNG::Init(h.z, h.y, fc.FrameNum - 1);
                lensSample = Sampling::UniformSampleDiskConcentric(rngDoF.Uniform2D());
                lensSample = lensSample * fc.LensRadius;
            }
            const float3 prevPos = Math::WorldPosFromScreenSpace2(f2((float)sx, (float)sy), renderDim, t.depth, fc.TanHalfFOV, fc.AspectRatio,
                f2(fc.PrevCameraJitter[0], fc.PrevCameraJitter[1]), row3(fc.PrevView, 0), row3(fc.PrevView, 1), row3(fc.PrevView, 2),
                fc.DoF != 0, lensSample, fc.FocusDepth, origin);
            const float tolerance = 0.005f * (fc.DoF ? 10.0f : 1.0f);
            if (!PlaneHeuristic(prevPos, normal, posW, viewZ, tolerance))
                continue;
            const float3 prevNormal = Math::DecodeUnitVector(t.normalEnc);
            bool ok = dot(prevNormal, normal) > 0.1f;
            if (roughness < 0.5f)
                ok = ok && (fabsf(t.roughness - roughness) < 0.15f);
            float prevEta_mat = BSDF::DEFAULT_ETA_MAT;
            if (t.flags.transmissive)
                prevEta_mat = DecodeIOR(t.iorEnc);
            ok = ok && (t.flags.transmissive == transmissive);
            ok = fc.DoF ? true : ok;
            valid[curr] = ok;
            if (ok)
            {
                TemporalSampleData& d = data[curr];
                d.sx = (int)(int16_t)sx; d.sy = (int)(int16_t)sy;
                d.posW = prevPos; d.normal = prevNormal; d.metallic = t.flags.metallic; d.roughness = t.roughness;
                d.transmissive = t.flags.transmissive; d.eta_next = prevEta_mat;
                curr++;
                if (curr == 2)
                    break;
            }
        }
        return curr;
    }

    ZR_D float TargetLumAtTemporalPixel(const FrameView& f, const SceneDev& sc, const GIReservoir& r_curr, const TemporalSampleData& c, bool testVisibility)
    {
        const zr_frame_constants& fc = f.fc;
        float3 wi = r_curr.pos - c.posW;
        if (dot(wi, wi) == 0)
            return 0;
        const float t = length(wi);
        wi = wi / fmaxf(t, 1e-6f);
        const float3 baseColor_prev = LoadPrev(f, c.sx, c.sy).baseColor;
        float3 camPos_prev = PrevCamPos(fc);
        if (fc.DoF)
        {
            uint3 h = RNG::PCG3d(make_uint3((uint32_t)c.sx, (uint32_t)c.sy, (uint32_t)c.sx));
            RNG rngDoF = RNG::Init(h.z, h.y, fc.FrameNum - 1);
            float2 lensSample = Sampling::UniformSampleDiskConcentric(rngDoF.Uniform2D());
            lensSample = lensSample * fc.LensRadius;
            camPos_prev += mad(lensSample.x, row3(fc.PrevView, 0), lensSample.y * row3(fc.PrevView, 1));
        }
        const float3 wo_prev = normalize(camPos_prev - c.posW);
        BSDF::ShadingData surface_prev = BSDF::ShadingData::Init(c.normal, wo_prev, c.metallic, c.roughness, baseColor_prev, BSDF::ETA_AIR,
            c.eta_next, c.transmissive, 0.0f, 0.0f, 0.0f, f3(0.0f), 0.0f, BSDF::DEFAULT_ETA_COAT, sc.rho);
        surface_prev.SetWi(wi, c.normal);
        const float3 target_prev = r_curr.Lo * BSDF::Unified(surface_prev).f;
        const float targetLum_prev = Math::Luminance(target_prev);
        if (testVisibility && targetLum_prev > 1e-5f)
        {
            if (!Visibility_Segment(sc, c.posW, wi, t, c.normal, r_curr.ID, surface_prev.Transmissive()))
                return 0;
        }
        return targetLum_prev;
    }

    ZR_D float JacobianReconnectionShift(float3 x2_normal, float3 x1_r, float3 x1_q, float3 x2_q)
    {
        float3 v_r = x1_r - x2_q;
        const float t_r2 = dot(v_r, v_r);
        v_r = dot(v_r, v_r) == 0 ? v_r : v_r / fmaxf(sqrtf(t_r2), 1e-6f);
        float3 v_q = x1_q - x2_q;
        const float t_q2 = dot(v_q, v_q);
        v_q = dot(v_q, v_q) == 0 ? v_q : v_q / fmaxf(sqrtf(t_q2), 1e-6f);
        const float cosPhi_r = dot(v_r, x2_normal);
        const float cosPhi_q = dot(v_q, x2_normal);
        return (fabsf(cosPhi_r) * t_q2) / fmaxf(fabsf(cosPhi_q) * t_r2, 1e-6f);
    }

    ZR_D void TemporalResample1(const FrameView& f, const SceneDev& sc, const zr_rgi_reservoir* prevRes, float3 posW, float3 normal, BSDF::ShadingData surface,
        const TemporalSampleData& c, GIReservoir& r, RNG& rng)
    {
        GIReservoir r_prev = PartialReadReservoir_Reuse(prevRes, f.W, f.H, c.sx, c.sy);
        const float M_new = (float)(uint16_t)(r.M + r_prev.M);
        if (r.w_sum != 0)
        {
            float targetLum_prev = 0.0f;
            if (r_prev.M > 0 && Math::Luminance(r.Lo) > 1e-6f)
                targetLum_prev = TargetLumAtTemporalPixel(f, sc, r, c, true);
            const float p_curr = Math::Luminance(r.target_z);
            const float J_curr_to_temporal = JacobianReconnectionShift(r.normal, c.posW, posW, r.pos);
            const float m_curr = p_curr / fmaxf(p_curr + r_prev.M * targetLum_prev * J_curr_to_temporal, 1e-6f);
            r.w_sum *= m_curr;
        }
        if (r_prev.ID == UINT32_MAX_ || (r_prev.Lo.x + r_prev.Lo.y + r_prev.Lo.z) == 0)
        {
            const float targetLum = Math::Luminance(r.target_z);
            r.W = targetLum > 0.0f ? r.w_sum / targetLum : 0.0f;
            r.M = M_new;
            return;
        }
        float3 wi = r_prev.pos - posW;
        const float t = length(wi);
        wi = wi / t;
        surface.SetWi(wi, normal);
        const float3 target_curr = r_prev.Lo * BSDF::Unified(surface).f;
        const float targetLum_curr = Math::Luminance(target_curr);
        if (targetLum_curr > 1e-6f)
        {
            if (Visibility_Segment(sc, posW, wi, t, normal, r_prev.ID, surface.Transmissive()))
            {
                PartialReadReservoir_ReuseRest(prevRes, f.W, f.H, c.sx, c.sy, r_prev);
                const float targetLum_prev = r_prev.W > 0 ? r_prev.w_sum / r_prev.W : 0;
                const float J_temporal_to_curr = JacobianReconnectionShift(r_prev.normal, posW, c.posW, r_prev.pos);
                const float numerator = r_prev.M * targetLum_prev;
                const float denom = numerator / fmaxf(J_temporal_to_curr, 1e-6f) + targetLum_curr;
                const float m_prev = numerator / fmaxf(denom, 1e-6f);
                const float w_prev = m_prev * targetLum_curr * r_prev.W;
                r.Update(w_prev, r_prev.pos, r_prev.normal, r_prev.ID, r_prev.Lo, target_curr, rng);
            }
        }
        const float targetLum = Math::Luminance(r.target_z);
        r.W = targetLum > 0.0f ? r.w_sum / targetLum : 0.0f;
        r.M = M_new;
    }

    ZR_D void TemporalResample2(const FrameView& f, const SceneDev& sc, const zr_rgi_reservoir* prevRes, float3 posW, float3 normal, BSDF::ShadingData surface,
        const TemporalSampleData c[2], GIReservoir& r, RNG& rng)
    {
        uint16_t M_new = (uint16_t)r.M;
        GIReservoir r_prev[2];
        for (int k = 0; k < 2; k++)
        {
            r_prev[k] = PartialReadReservoir_Reuse(prevRes, f.W, f.H, c[k].sx, c[k].sy);
            M_new = (uint16_t)(M_new + (uint16_t)r_prev[k].M);
        }
        {
            const float p_curr = Math::Luminance(r.target_z);
            float denom = p_curr;
            if (Math::Luminance(r.Lo) > 1e-5f)
            {
                for (int p = 0; p < 2; p++)
                {
                    if (r_prev[p].M == 0)
                        continue;
                    const float targetLum_prev = TargetLumAtTemporalPixel(f, sc, r, c[p], p != 0);
                    const float J_curr_to_temporal = JacobianReconnectionShift(r.normal, c[p].posW, posW, r.pos);
                    denom += r_prev[p].M * J_curr_to_temporal * targetLum_prev;
                }
            }
            const float m_curr = denom == 0 ? 0 : p_curr / denom;
            r.w_sum *= m_curr;
        }
        for (int i = 0; i < 2; i++)
        {
            float3 wi = r_prev[i].pos - posW;
            const float t = (wi.x == 0 && wi.y == 0 && wi.z == 0) ? 0 : length(wi);
            wi = wi / fmaxf(t, 1e-6f);
            surface.SetWi(wi, normal);
            const float3 target_curr = r_prev[i].Lo * BSDF::Unified(surface).f;
            const float targetLum_curr = Math::Luminance(target_curr);
            if (targetLum_curr < 1e-5f)
                continue;
            if (Visibility_Segment(sc, posW, wi, t, normal, r_prev[i].ID, surface.Transmissive()))
            {
                PartialReadReservoir_ReuseRest(prevRes, f.W, f.H, c[i].sx, c[i].sy, r_prev[i]);
                const float targetLum_prev = r_prev[i].W > 0 ? r_prev[i].w_sum / r_prev[i].W : 0;
                const float J_temporal_to_curr = JacobianReconnectionShift(r_prev[i].normal, posW, c[i].posW, r_prev[i].pos);
                const float numerator = r_prev[i].M * targetLum_prev;
                float denom = (numerator / J_temporal_to_curr) + targetLum_curr;
                if (r_prev[1 - i].M > 0 && targetLum_prev > 0)
                {
                    const float J_temporal_to_temporal = JacobianReconnectionShift(r_prev[i].normal, c[1 - i].posW, c[i].posW, r_prev[i].pos);
                    const float targetLum_other = TargetLumAtTemporalPixel(f, sc, r_prev[i], c[1 - i], true);
                    denom += r_prev[1 - i].M * targetLum_other / fmaxf(J_temporal_to_temporal, 1e-6f);
                }
                denom = J_temporal_to_curr == 0 ? 0 : denom;
                const float m_prev = denom == 0 ? 0 : numerator / denom;
                const float w_prev = m_prev * targetLum_curr * r_prev[i].W;
                r.Update(w_prev, r_prev[i].pos, r_prev[i].normal, r_prev[i].ID, r_prev[i].Lo, target_curr, rng);
            }
        }
        const float targetLum = Math::Luminance(r.target_z);
        r.W = targetLum > 0.0f ? r.w_sum / targetLum : 0.0f;
        r.M = (float)M_new;
    }


#ifndef ZR_RGI_THREADS
#define ZR_RGI_THREADS 1024
#endif
    // PlainPT = true: IndirectLighting/PathTracer/PathTracer.hlsl (INTEGRATOR::PATH_TRACING) -- the same dispatch shape, RNG
    // seeds and PathTrace loop without reservoirs or reuse (main :98-212, EstimateIndirectLighting :58-104).
    template<bool PlainPT>
    __global__ void ZR_LB(ZR_RGI_THREADS) k_rgi(SceneDev sc, FrameView f, GIParams prm, zr_rgi_reservoir* __restrict__ resCurr,
        const zr_rgi_reservoir* __restrict__ resPrev, float4* __restrict__ finalImg, uint32_t dispX, uint32_t dispY,
        const uint32_t* __restrict__ order)
    {
        const zr_frame_constants& fc = f.fc;
        uint2 sg = make_uint2(0, 0);
        const uint32_t groupFlat = order[blockIdx.x] * (ZR_RGI_THREADS / 64) + (threadIdx.x >> 6);
        const uint32_t tInGroup = threadIdx.x & 63;
        uint2 px = make_uint2(0xffffffffu, 0xffffffffu);
        if (groupFlat < dispX * dispY)
            px = SwizzleThreadGroup(groupFlat, 0, tInGroup & 7, tInGroup >> 3, 8, 8, dispX, 16, 4, 16 * dispY, sg);
        bool active = px.x < f.W && px.y < f.H && px.y >= prm.rowBegin && px.y < prm.rowEnd;
        const size_t idx = active ? (size_t)px.y * f.W + px.x : 0;
        GFlags flags = DecodeFlags(0);
        float roughness = 0;
        float3 baseColor = f3(0);
        if (active)
        {
            const uint4 g = ld128(&f.core[idx]);
            flags = DecodeFlags(g.w & 0xff);
            roughness = (float)((g.w >> 8) & 0xff) / 255.0f;
            baseColor = f3((float)(g.z & 0xff) / 255.0f, (float)((g.z >> 8) & 0xff) / 255.0f, (float)((g.z >> 16) & 0xff) / 255.0f);
            if (flags.invalid || flags.emissive)
            {
                if (!fc.Accumulate || !fc.CameraStatic)
                {
                    const float4 prev = finalImg[idx];
                    finalImg[idx] = f4(0, 0, 0, prev.w);
                }
                active = false;
            }
        }
        // ---- main :96-150 + EstimateIndirectLighting :547-556 ----
        Pixel p;
        BSDF::ShadingData surface0 = BSDF::ShadingData::InitEmpty();
        RNG rngThread, rngGroup;
        rngThread.State = rngGroup.State = 0;
        int maxNumBounces = 0;
        uint32_t sampleSetIdx = 0;
        GIReservoir r = GIReservoir::Init();
        BSDF::BSDFSample bsdfSample0 = BSDF::BSDFSample::Init();
        if (active)
        {
            p = LoadPixel(f, sc, f.core, f.coat, (int)px.x, (int)px.y, false, (int)px.x, (int)px.y);
            const float3 wo = normalize(p.origin - p.pos);
            // PathTracer.hlsl:184-185 also passes flags.trDepthGt0 as the transmission depth
            surface0 = BSDF::ShadingData::Init(p.normal, wo, flags.metallic, roughness, baseColor, BSDF::ETA_AIR, p.eta_next, flags.transmissive,
                (PlainPT && flags.trDepthGt0) ? 1.0f : 0.0f, 0.0f, 0.0f, f3(0.0f), 0.0f, BSDF::DEFAULT_ETA_COAT, sc.rho);
            rngGroup = RNG::Init(sg.x ^ 61u, sg.y ^ 61u, fc.FrameNum);
            rngThread = RNG::Init(px.x ^ 511u, px.y ^ 31u, fc.FrameNum);
            maxNumBounces = (int)(flags.transmissive ? prm.maxGlossyTrBounces : prm.maxNonTrBounces);
            if (!PlainPT && prm.stochasticMultiBounce && (roughness >= 0.1f || fc.CameraStatic))
                maxNumBounces = rngGroup.Uniform() < 0.5f ? 1 : maxNumBounces;
            sampleSetIdx = rngGroup.UniformUintBounded_Faster(sc.numSampleSets);
        }
        ZR_PHASE();
        // ---- RIS_InitialCandidates :39-81 ----
        if (active)
            bsdfSample0 = BSDF::SampleBSDF(p.normal, surface0, rngThread);
        bool traced = active && (bsdfSample0.pdf != 0);
        ZR_PHASE();
        Hit hit0 = RPT::MissHit();
        if (traced)
        {
            hit0 = FindClosest(sc, p.pos, p.normal, bsdfSample0.wi, surface0.Transmissive());
            traced = hit0.hit;
        }
        float3 hitPos0 = f3(0);
        // ---- PathTrace (PathTracing.hlsli:10-98), lock-step ----
        float3 pos = f3(0), normal = f3(0), li = f3(0), throughput = f3(1.0f);
        float eta_curr = BSDF::ETA_AIR, eta_next = BSDF::DEFAULT_ETA_MAT;
        bool inTranslucentMedium = false;
        int bounce = 0;
        BSDF::BSDFSample bsdfSample = bsdfSample0;
        Hit hitInfo = hit0;
        BSDF::ShadingData surface = BSDF::ShadingData::InitEmpty();
        bool tracing = traced;
        if (traced)
        {
            hitPos0 = p.pos + hit0.t * bsdfSample0.wi;
            pos = p.pos; normal = p.normal;
            eta_curr = dot(normal, bsdfSample0.wi) < 0 ? p.eta_next : BSDF::ETA_AIR;
            inTranslucentMedium = dot(normal, bsdfSample0.wi) < 0;
        }
        while (__syncthreads_or(tracing))
        {
            bool atRR = false;
            if (tracing)
            {
                const float3 hitPos = mad(hitInfo.t, bsdfSample.wi, pos);
                if (!GetMaterialData(sc, -bsdfSample.wi, eta_curr, hitInfo, surface, eta_next))
                    tracing = false;
                else
                {
                    li += throughput * NEE<PlainPT>(sc, fc, hitPos, hitInfo.normal, surface, sampleSetIdx, bounce, rngThread);
                    // ACCOUNT_FOR_TRANSMITTANCE == 1 (PathTracing.hlsli:41-48): Beer's law inside a translucent medium
                    if (PlainPT && inTranslucentMedium && (surface.trDepth > 0))
                    {
                        const float3 c = surface.baseColor_Fr0_TrCol;
                        const float3 extCoeff = f3(-zr_logf(c.x), -zr_logf(c.y), -zr_logf(c.z)) / surface.trDepth;
                        throughput *= f3(zr_expf(-hitInfo.t * extCoeff.x), zr_expf(-hitInfo.t * extCoeff.y), zr_expf(-hitInfo.t * extCoeff.z));
                    }
                    if (bounce >= (maxNumBounces - 1))
                        tracing = false;
                    else
                    {
                        pos = hitPos;
                        normal = hitInfo.normal;
                        bounce++;
                        atRR = true;
                    }
                }
            }
            // Russian roulette against the wave's maximum throughput
            const uint32_t rrMask = __ballot_sync(0xffffffffu, atRR);
            bool doRR = false;
            float waveThroughput = 0.0f;
            if (rrMask)
            {
                const int rrBounce = __shfl_sync(0xffffffffu, bounce, __ffs(rrMask) - 1);
                doRR = prm.russianRoulette && (rrBounce >= 3);
                if (doRR)
                    waveThroughput = WaveMax32(atRR ? Math::Luminance(throughput) : -FLT_MAX_);
            }
            ZR_PHASE();
            if (atRR)
            {
                bool go = true;
                if (doRR)
                {
                    const float p_terminate = fmaxf(0.05f, 1 - waveThroughput);
                    if (rngGroup.Uniform() < p_terminate)
                        go = false;
                    else
                        throughput /= (1 - p_terminate);
                }
                if (go)
                {
                    bsdfSample = BSDF::BSDFSample::Init();
                    if (bounce < maxNumBounces)
                        bsdfSample = BSDF::SampleBSDF(normal, surface, rngThread);
                    if (Math::Luminance(bsdfSample.bsdfOverPdf) == 0)
                        go = false;
                }
                if (go)
                {
                    hitInfo = FindClosest(sc, pos, normal, bsdfSample.wi, surface.Transmissive());
                    if (!hitInfo.hit)
                        go = false;
                }
                if (go)
                {
                    throughput *= bsdfSample.bsdfOverPdf;
                    const bool transmitted = dot(normal, bsdfSample.wi) < 0;
                    eta_curr = transmitted ? (eta_curr == BSDF::ETA_AIR ? eta_next : BSDF::ETA_AIR) : eta_curr;
                    inTranslucentMedium = transmitted ? !inTranslucentMedium : inTranslucentMedium;
                }
                tracing = go;
            }
        }
        if (PlainPT)
        {
            // EstimateIndirectLighting :97-103, main :199-211
            if (!active)
                return;
            float3 liOut = f3(0);
            if (traced)
            {
                liOut = li;
                if (dot(liOut, liOut) > 0)
                    liOut *= bsdfSample0.bsdfOverPdf;
            }
            liOut = isnan3(liOut) ? f3(0) : liOut;
            const float4 prev = finalImg[idx];
            if (fc.Accumulate && fc.CameraStatic)
                finalImg[idx] = f4(prev.x + liOut.x, prev.y + liOut.y, prev.z + liOut.z, prev.w);
            else
                finalImg[idx] = f4(liOut.x, liOut.y, liOut.z, prev.w);
            return;
        }
        // ---- rest of RIS_InitialCandidates :83-113 ----
        if (traced)
        {
            const float3 lo = li;
            float3 target = lo;
            if (dot(lo, lo) > 0)
            {
                surface0.SetWi(bsdfSample0.wi, p.normal);
                target *= BSDF::Unified(surface0).f;
            }
            const float targetLum = Math::Luminance(target);
            const float w = targetLum / fmaxf(bsdfSample0.pdf, 1e-6f);
            r.Update(w, hitPos0, hit0.normal, hit0.ID, lo, target, rngThread);
            r.W = targetLum > 0 ? 1.0f / bsdfSample0.pdf : 0.0f;
        }
        ZR_PHASE();
        // ---- temporal reuse (EstimateIndirectLighting :564-596) ----
        if (prm.temporalResample)
        {
            if (active)
            {
                const float2 renderDim = f2((float)f.W, (float)f.H);
                const float2 motionVec = unpack_snorm16x2(__ldg(&f.me[idx].x));
                const float2 currUV = f2((float)px.x + 0.5f, (float)px.y + 0.5f) / renderDim;
                const float2 prevUV = currUV - motionVec;
                TemporalSampleData data[2]; bool valid[2];
                FindTemporalCandidate(f, sc, (int)px.x, (int)px.y, p.pos, p.normal, p.z, roughness, surface0.specTr, prevUV, rngThread, data, valid);
                if (valid[1] && roughness > 0.05f)
                    TemporalResample2(f, sc, resPrev, p.pos, p.normal, surface0, data, r, rngThread);
                else if (valid[0])
                    TemporalResample1(f, sc, resPrev, p.pos, p.normal, surface0, data[0], r, rngThread);
            }
            if (prm.boilingSuppression)
            {
                const float waveSum = WaveSum32(active ? r.w_sum : 0.0f);
                if (active)
                {
                    const float waveAvg = (waveSum - r.w_sum) / 31.0f;
                    if (r.w_sum > 25 * waveAvg)
                        r.M = 1;
                }
            }
        }
        if (!active)
            return;
        if (prm.temporalResample || prm.resetTemporal)
        {
            zr_rgi_reservoir rec;
            WriteReservoir(rec, r, (float)prm.M_max);
            uint4 v[3];
            memcpy(v, &rec, 48);
            uint4* q = reinterpret_cast<uint4*>(&resCurr[idx]);
            q[0] = v[0]; q[1] = v[1]; q[2] = v[2];
        }
        float3 liOut = r.target_z * r.W;
        liOut = isnan3(liOut) ? f3(0) : liOut;
        const float4 prev = finalImg[idx];
        if (fc.Accumulate && fc.CameraStatic)
            finalImg[idx] = f4(prev.x + liOut.x, prev.y + liOut.y, prev.z + liOut.z, prev.w);
        else
            finalImg[idx] = f4(liOut.x, liOut.y, liOut.z, prev.w);
    }
}
} // namespace zr

// ------------------------------------------------------------------------------------------------
// IndirectLighting pass object for INTEGRATOR::ReSTIR_GI (IndirectLighting.cpp:277-368, :1016-1024)
// ------------------------------------------------------------------------------------------------
struct zr_gi_pass
{
    uint32_t width = 0, height = 0;
    zr_rgi_reservoir* d_res[2] = { nullptr, nullptr };
    float4* d_final = nullptr;
    int currTemporalIdx = 0;
    bool isTemporalReservoirValid = false;
    bool resetTemporalTextures = true;
    zr_gi_params params{};
    bool plainPathTracer = false;       // INTEGRATOR::PATH_TRACING instead of ReSTIR_GI (both read cb_ReSTIR_GI in the reference)
    uint32_t rowBegin = 0, rowEnd = 0xffffffffu;
    zr::TileCosts tileCosts;
    zr::BlockSchedule sched;

    static void Defaults(zr_gi_params* p)
    {
        // IndirectLighting.h:231-244, IndirectLighting.cpp:143-160
        p->max_non_tr_bounces = 3; p->max_glossy_tr_bounces = 4; p->russian_roulette = 1; p->stochastic_multi_bounce = 1;
        p->boiling_suppression = 1; p->M_max = 10; p->temporal_resample = 1;
    }
    void Release()
    {
        for (int i = 0; i < 2; i++) { if (d_res[i]) cudaFree(d_res[i]); d_res[i] = nullptr; }
        if (d_final) cudaFree(d_final);
        d_final = nullptr;
        sched.Release();
    }
    zr_status OnWindowResized(uint32_t w, uint32_t h)
    {
        Release();
        width = w; height = h;
        const size_t n = (size_t)w * h;
        for (int i = 0; i < 2; i++) ZR_CUDA(cudaMalloc(&d_res[i], n * sizeof(zr_rgi_reservoir)));
        ZR_CUDA(cudaMalloc(&d_final, n * 16));
        return ResetTemporal();
    }
    zr_status ResetTemporal()
    {
        const size_t n = (size_t)width * height;
        for (int i = 0; i < 2; i++) ZR_CUDA(cudaMemset(d_res[i], 0, n * sizeof(zr_rgi_reservoir)));
        ZR_CUDA(cudaMemset(d_final, 0, n * 16));
        currTemporalIdx = 0; isTemporalReservoirValid = false; resetTemporalTextures = true;
        return ZR_OK;
    }
    zr_status Render(const zr_frame_inputs* in, cudaStream_t stream)
    {
        using namespace zr;
        if (!in || !in->scene || !in->curr.d_core || !in->curr.d_motion_emissive || !in->curr.d_coat)
        {
            set_error("zr_gi_pass_render: missing scene or G-buffer");
            return ZR_ERR_INVALID_ARG;
        }
        if (in->frame.RenderWidth != width || in->frame.RenderHeight != height)
        {
            set_error("zr_gi_pass_render: frame is %ux%u but the pass was sized %ux%u", in->frame.RenderWidth, in->frame.RenderHeight, width, height);
            return ZR_ERR_INVALID_ARG;
        }
        if (in->scene->dev.numEmissives == 0 || !in->scene->aliasBuilt)
        {
            set_error("zr_gi_pass_render: the emissive variant needs emissive triangles and zr_prelighting_render first "
                "(the sun/sky variant is not part of this build)");
            return ZR_ERR_UNSUPPORTED;
        }
        if (in->scene->dev.sampleSetSize && !in->scene->samplesValid)
        {
            set_error("zr_gi_pass_render: presampling is enabled but zr_presample_emissives has not run");
            return ZR_ERR_NOT_INITIALIZED;
        }
        if (in->scene->dev.lvg && in->scene->dev.sampleSetSize && !in->scene->lvgValid)
        {
            set_error("zr_gi_pass_render: the light voxel grid is enabled but zr_build_light_voxel_grid has not run");
            return ZR_ERR_NOT_INITIALIZED;
        }
        const bool doTemporal = !plainPathTracer && params.temporal_resample && isTemporalReservoirValid;
        if (doTemporal && !in->prev.d_core)
        {
            set_error("zr_gi_pass_render: temporal reuse needs the previous G-buffer");
            return ZR_ERR_INVALID_ARG;
        }
        FrameView f;
        f.fc = in->frame;
        f.core = (const uint4*)in->curr.d_core; f.depth = (const float*)in->curr.d_depth;
        f.me = (const uint2*)in->curr.d_motion_emissive; f.coat = (const uint2*)in->curr.d_coat;
        f.pcore = (const uint4*)in->prev.d_core; f.pcoat = (const uint2*)in->prev.d_coat;
        f.W = width; f.H = height;
        GIParams prm{ params.max_non_tr_bounces, params.max_glossy_tr_bounces, params.russian_roulette, params.stochastic_multi_bounce,
            params.boiling_suppression, params.M_max, doTemporal ? 1u : 0u, resetTemporalTextures ? 1u : 0u, rowBegin,
            rowEnd < height ? rowEnd : height };
        const uint32_t dispX = (width + 7) / 8, dispY = (height + 7) / 8;
        if (!sched.UpToDate(prm.rowBegin, prm.rowEnd, tileCosts.version))
            ZR_CUDA(sched.Upload(ScheduleSwizzled(dispX, dispY, 8, 8, ZR_RGI_THREADS / 64, prm.rowBegin, prm.rowEnd, tileCosts), prm.rowBegin, prm.rowEnd,
                tileCosts.version));
        const int cur = currTemporalIdx;
        if (plainPathTracer)
        {
            ZR_PROF("k_pathtracer", stream);
            k_rgi<true><<<sched.count, ZR_RGI_THREADS, 0, stream>>>(in->scene->dev, f, prm, d_res[cur], d_res[1 - cur], d_final, dispX, dispY, sched.d_order);
            ZR_LAUNCH_CHECK();
            return ZR_OK;       // no reservoirs: the ReSTIR GI history is left as it is (and is dropped by SetMethod)
        }
        ZR_PROF("k_rgi", stream);
        k_rgi<false><<<sched.count, ZR_RGI_THREADS, 0, stream>>>(in->scene->dev, f, prm, d_res[cur], d_res[1 - cur], d_final, dispX, dispY, sched.d_order);
        ZR_LAUNCH_CHECK();
        isTemporalReservoirValid = true;
        currTemporalIdx = 1 - cur;
        resetTemporalTextures = false;
        return ZR_OK;
    }
};

extern "C"
{
    zr_status zr_gi_pass_create(uint32_t width, uint32_t height, zr_gi_pass** out)
    {
        if (!out || !width || !height) { zr::set_error("zr_gi_pass_create: bad args"); return ZR_ERR_INVALID_ARG; }
        zr_gi_pass* p = new zr_gi_pass();
        zr_gi_pass::Defaults(&p->params);
        zr_status s = p->OnWindowResized(width, height);
        if (s != ZR_OK) { p->Release(); delete p; return s; }
        *out = p;
        return ZR_OK;
    }
    zr_status zr_gi_pass_resize(zr_gi_pass* p, uint32_t width, uint32_t height)
    {
        if (!p || !width || !height) return ZR_ERR_INVALID_ARG;
        return p->OnWindowResized(width, height);
    }
    zr_status zr_gi_pass_reset_temporal(zr_gi_pass* p) { return p ? p->ResetTemporal() : ZR_ERR_INVALID_ARG; }
    zr_status zr_gi_pass_default_params(zr_gi_params* out)
    {
        if (!out) return ZR_ERR_INVALID_ARG;
        zr_gi_pass::Defaults(out);
        return ZR_OK;
    }
    zr_status zr_gi_pass_set_params(zr_gi_pass* p, const zr_gi_params* params)
    {
        if (!p || !params) return ZR_ERR_INVALID_ARG;
        if (params->max_non_tr_bounces < 1 || params->max_non_tr_bounces > 8 || params->max_glossy_tr_bounces < 1 ||
            params->max_glossy_tr_bounces > 8 || params->M_max == 0 || params->M_max > 2047)
        {
            zr::set_error("zr_gi_pass_set_params: value out of range (bounces 1..8, M_max 1..2047)");
            return ZR_ERR_INVALID_ARG;
        }
        p->params = *params;
        return ZR_OK;
    }
    // IndirectLighting::SetMethod for the two integrators that share this pass object (IndirectLighting.cpp:203-235):
    // ZR_INTEGRATOR_PATH_TRACING or ZR_INTEGRATOR_RESTIR_GI; a change drops the temporal history.
    zr_status zr_gi_pass_set_method(zr_gi_pass* p, zr_integrator method)
    {
        if (!p || (method != ZR_INTEGRATOR_PATH_TRACING && method != ZR_INTEGRATOR_RESTIR_GI))
        {
            zr::set_error("zr_gi_pass_set_method: PATH_TRACING (0) or RESTIR_GI (1); ReSTIR PT is zr_indirect_pass");
            return ZR_ERR_INVALID_ARG;
        }
        const bool plain = method == ZR_INTEGRATOR_PATH_TRACING;
        if (plain == p->plainPathTracer) return ZR_OK;
        p->plainPathTracer = plain;
        return p->ResetTemporal();
    }
    zr_status zr_gi_pass_render(zr_gi_pass* p, const zr_frame_inputs* in, void* stream)
    {
        if (!p) return ZR_ERR_INVALID_ARG;
        return p->Render(in, (cudaStream_t)stream);
    }
    zr_status zr_gi_pass_get_output(zr_gi_pass* p, zr_gi_output id, zr_image2d* out)
    {
        if (!p || !out) return ZR_ERR_INVALID_ARG;
        const uint32_t w = p->width, h = p->height;
        switch (id)
        {
        case ZR_GI_FINAL: *out = zr_image2d{ p->d_final, w, h, w * 16u, 16u }; break;
        case ZR_GI_RESERVOIR_CURR: *out = zr_image2d{ p->d_res[1 - p->currTemporalIdx], w, h, w * 48u, 48u }; break;
        case ZR_GI_RESERVOIR_PREV: *out = zr_image2d{ p->d_res[p->currTemporalIdx], w, h, w * 48u, 48u }; break;
        default: zr::set_error("zr_gi_pass_get_output: unknown output id"); return ZR_ERR_INVALID_ARG;
        }
        return ZR_OK;
    }
    void zr_gi_pass_destroy(zr_gi_pass* p) { if (p) { p->Release(); delete p; } }
}
