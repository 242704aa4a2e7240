// zr_rgi.cuh -- the device functions of ReSTIR GI and of the plain path tracer (reservoir record, next-event estimation variants,
// temporal candidate search, reconnection Jacobian, temporal resampling with one and two candidates), kept in a header so that the
// host build of the device source (tests/hostsim) can hold them to the oracle without a GPU; rgi.cu adds the kernel and the pass.
//   IndirectLighting/ReSTIR_GI/Resampling.hlsli      FindTemporalCandidate :124-233, TargetLumAtTemporalPixel :235-284,
//                                                    JacobianReconnectionShift :288-308, TemporalResample1 :310-371, TemporalResample2 :373-454
//   IndirectLighting/ReSTIR_GI/ReSTIR_GI_NEE.hlsli   NEE_Emissive_MIS :8-121, NEE_Emissive_LVG :123-193, NEE :195-270
//   IndirectLighting/NEE.hlsli                       NEE_Emissive<1> :152-221
//   IndirectLighting/ReSTIR_GI/Reservoir.hlsli       Reservoir, read / write :9-131
#pragma once
#include "zr_pixel.cuh"

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
}
} // namespace zr
