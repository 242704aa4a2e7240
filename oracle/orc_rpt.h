// ORACLE -- test infrastructure, not product code (see orc_math.h header).
//
// ReSTIR PT state containers and the hybrid shift, restated on the CPU:
//   IndirectLighting/ReSTIR_PT/Shift.hlsli       Reconnection :16-172, OffsetPathContext :191-358,
//                                                CanReconnect :360-375, Replay :377-474, StepPath :476-546,
//                                                Shift2 :662-816, Replay_kGt2 :818-859
//   IndirectLighting/ReSTIR_PT/Reservoir.hlsli   Reservoir :9-463 (7 planes -> one 64-byte record)
//   IndirectLighting/ReSTIR_PT/ReSTIR_PT_NEE.hlsli  NEE_Bsdf :145-222, NEE_Emissive :224-302,
//                                                EvalDirect_Emissive_Case2/3 :306-391
// Emissive-light variant only (NEE_EMISSIVE == 1, no presampled sets); sun/sky is out of scope.
// Parity unpinned: the reference has no tests for any of this.
#pragma once
#include "orc_scene.h"

namespace orc
{
namespace RPT
{
    using BSDF::LOBE;
    using BSDF::ShadingData;
    using Light::TYPE;

    struct Reconnection
    {
        static constexpr uint32_t EMPTY = 0xf;
        float3 x_k;
        uint32_t ID;
        uint32_t meshIdx;
        float partialJacobian;
        float3 w_k_lightNormal_w_sky;
        float lightPdf;
        uint32_t seed_replay;
        uint32_t seed_nee;
        float dwdA;
        float3 L;           // half3: values are rounded to binary16 on assignment
        uint32_t k;
        LOBE lobe_k_min_1;
        LOBE lobe_k;
        TYPE lt_k;
        TYPE lt_k_plus_1;
        bool x_k_in_motion;

        static Reconnection Init()
        {
            Reconnection ret;
            ret.k = EMPTY;
            ret.lt_k = Light::NONE;
            ret.lt_k_plus_1 = Light::NONE;
            ret.partialJacobian = 0;
            ret.x_k = f3(FLT_MAX_);
            ret.seed_replay = 0;
            ret.w_k_lightNormal_w_sky = f3(0);
            ret.L = f3(0);
            ret.lightPdf = 0;
            ret.seed_nee = 0;
            ret.dwdA = 0;
            ret.ID = 0; ret.meshIdx = 0; ret.lobe_k_min_1 = BSDF::DIFFUSE_R; ret.lobe_k = BSDF::DIFFUSE_R; ret.x_k_in_motion = false;
            return ret;
        }
        bool Empty() const { return k == EMPTY; }
        bool IsCase2() const { return lt_k_plus_1 != Light::NONE; }
        bool IsCase3() const { return lt_k != Light::NONE; }
        bool IsCase1() const { return !IsCase2() && !IsCase3(); }
        void Clear() { k = EMPTY; lt_k = Light::NONE; lt_k_plus_1 = Light::NONE; }
        static float3 half3(float3 v) { return f3(to_half(v.x), to_half(v.y), to_half(v.z)); }

        void SetCase1(int k_, float3 x_k_, float t, float3 normal_k, uint32_t hitID, uint32_t meshIdx_, float3 w_k_min_1,
            LOBE l_k_min_1, float pdf_w_k_min_1, float3 w_k, LOBE l_k, float pdf_w_k)
        {
            lobe_k_min_1 = l_k_min_1;
            k = (uint32_t)k_; x_k = x_k_; ID = hitID; meshIdx = meshIdx_;
            lt_k = Light::NONE; lobe_k = l_k; w_k_lightNormal_w_sky = w_k;
            lt_k_plus_1 = Light::NONE;
            partialJacobian = pdf_w_k_min_1;
            float cos_theta_k = fabsf(dot(-w_k_min_1, normal_k));
            partialJacobian *= cos_theta_k / (t * t);
            partialJacobian *= pdf_w_k;
        }
        void SetCase2(int k_, float3 x_k_, float t, float3 normal_k, uint32_t hitID, uint32_t meshIdx_, float3 w_k_min_1,
            LOBE l_k_min_1, float pdf_w_k_min_1, float3 w_k, LOBE l_k, float pdf_w_k, TYPE t_k_plus_1, float pdf_light,
            float3 le, uint32_t seed, float dwdA_)
        {
            lobe_k_min_1 = l_k_min_1;
            k = (uint32_t)k_; x_k = x_k_; ID = hitID; meshIdx = meshIdx_;
            lt_k = Light::NONE; lobe_k = l_k; w_k_lightNormal_w_sky = w_k;
            lt_k_plus_1 = t_k_plus_1;
            lightPdf = pdf_light; dwdA = dwdA_; seed_nee = seed; L = half3(le);
            partialJacobian = pdf_w_k_min_1;
            float cos_theta_k = fabsf(dot(-w_k_min_1, normal_k));
            partialJacobian *= cos_theta_k / (t * t);
            if (lobe_k != BSDF::ALL)
                partialJacobian *= pdf_w_k;
        }
        void SetCase3(int k_, float3 x_k_, TYPE t, LOBE l_k_min_1, uint32_t lightID, float3 le, float3 lightNormal,
            float pdf_solidAngle, float pdf_light, float dwdA_, float3 w_sky, bool twoSided, uint32_t seed)
        {
            lobe_k_min_1 = l_k_min_1;
            k = (uint32_t)k_; x_k = x_k_; ID = lightID; lt_k = t; seed_nee = seed;
            partialJacobian = l_k_min_1 == BSDF::ALL ? 1.0f : pdf_solidAngle * dwdA_;
            lightPdf = twoSided ? pdf_light : -pdf_light;
            L = half3(le);
            lt_k_plus_1 = Light::NONE;
            if (t == Light::EMISSIVE) w_k_lightNormal_w_sky = lightNormal;
            else if (t == Light::SKY) w_k_lightNormal_w_sky = w_sky;
        }
    };

    struct Reservoir
    {
        float w_sum, W;
        float3 target;
        Reconnection rc;
        uint32_t M;

        static Reservoir Init()
        {
            Reservoir r;
            r.rc = Reconnection::Init();
            r.w_sum = 0; r.W = 0; r.M = 0; r.target = f3(0);
            return r;
        }
        bool Update(float weight, float3 target_, const Reconnection& rc_, RNG& rng)
        {
            if ((weight != weight) || isinf1(weight))
                return false;
            M += 1;
            if (weight == 0)
                return false;
            w_sum += weight;
            if (rng.Uniform() < (weight / w_sum))
            {
                rc = rc_;
                target = target_;
                return true;
            }
            return false;
        }

        // ---- 64-byte record <-> reservoir (Reservoir.hlsli:52-463, Emissive == true) ----
        void UnpackMetadata(uint32_t meta)
        {
            const uint32_t x = meta & 0xff, y = (meta >> 8) & 0xff, z = (meta >> 16) & 0xff;
            uint32_t k = x & 0xf;
            rc.k = k == Reconnection::EMPTY ? k : k + 2;
            rc.lobe_k_min_1 = BSDF::LobeFromValue(y & 0x7);
            rc.lobe_k = BSDF::LobeFromValue((y >> 3) & 0x7);
            rc.lt_k = Light::TypeFromValue((y >> 6) & 0x3);
            rc.lt_k_plus_1 = Light::TypeFromValue(z & 0x3);
            rc.x_k_in_motion = (z >> 2) != 0;
            M = x >> 4;
        }
        static Reservoir Load_NonReconnection(const zr_rpt_reservoir& s)
        {
            Reservoir r = Init();
            r.UnpackMetadata(s.meta);
            r.w_sum = s.w_sum; r.W = s.W;
            return r;
        }
        void Load_Reconnection(const zr_rpt_reservoir& s)
        {
            const float3 L_ = f3(half_lo(s.L_rg), half_hi(s.L_rg), half_lo(s.L_b));
            if (rc.IsCase1())
            {
                rc.partialJacobian = asfloat(s.jacobian_or_seed_nee);
                rc.seed_replay = s.seed_replay; rc.ID = s.ID;
                rc.w_k_lightNormal_w_sky = Math::DecodeOct32(s.w_k);
                rc.x_k = f3(asfloat(s.x_k_x), asfloat(s.x_k_y), asfloat(s.x_k_z));
                rc.meshIdx = s.meshIdx;
                rc.L = L_;
            }
            else if (rc.IsCase2())
            {
                rc.partialJacobian = asfloat(s.jacobian_or_seed_nee);
                rc.seed_replay = s.seed_replay; rc.ID = s.ID;
                rc.x_k = f3(asfloat(s.x_k_x), asfloat(s.x_k_y), asfloat(s.x_k_z));
                rc.L = L_;
                rc.w_k_lightNormal_w_sky = Math::DecodeOct32(s.w_k);
                rc.lightPdf = s.lightPdf; rc.dwdA = s.dwdA; rc.seed_nee = s.seed_nee; rc.meshIdx = s.meshIdx;
            }
            else
            {
                rc.seed_replay = s.seed_replay; rc.ID = s.ID;
                rc.partialJacobian = rc.lobe_k_min_1 == BSDF::ALL ? 1.0f : asfloat(s.jacobian_or_seed_nee);
                rc.x_k = f3(asfloat(s.x_k_x), asfloat(s.x_k_y), asfloat(s.x_k_z));
                rc.L = L_;
                rc.lightPdf = s.lightPdf;
                rc.seed_nee = s.jacobian_or_seed_nee;
                rc.w_k_lightNormal_w_sky = Math::DecodeOct32(s.w_k);
            }
        }
        static Reservoir Load(const zr_rpt_reservoir& s)
        {
            Reservoir r = Load_NonReconnection(s);
            if (r.rc.Empty())
                return r;
            r.Load_Reconnection(s);
            return r;
        }
        uint32_t PackMeta(uint32_t M_max) const
        {
            uint32_t m = M_max == 0 ? M : (M < M_max ? M : M_max);
            if (m > 15) m = 15;
            uint32_t k = rc.Empty() ? rc.k : (rc.k > 2 ? rc.k : 2) - 2;
            uint32_t x = (k | (m << 4)) & 0xff;
            uint32_t y = (uint32_t)rc.lobe_k_min_1 | ((uint32_t)rc.lobe_k << 3) | ((uint32_t)rc.lt_k << 6);
            uint32_t z = (uint32_t)rc.lt_k_plus_1 | ((rc.x_k_in_motion ? 1u : 0u) << 2);
            return x | ((y & 0xff) << 8) | ((z & 0xff) << 16);
        }
        // Reservoir::Write<true>: the whole record; fields a case does not store are zero
        void Write(zr_rpt_reservoir& s, uint32_t M_max = 0)
        {
            memset(&s, 0, sizeof(s));
            s.meta = PackMeta(M_max);
            w_sum = Math::Sanitize(w_sum);
            W = Math::Sanitize(W);
            s.w_sum = w_sum; s.W = W;
            if (rc.Empty())
                return;
            const uint32_t w_k_encoded = Math::EncodeOct32u(rc.w_k_lightNormal_w_sky);
            s.seed_replay = rc.seed_replay; s.ID = rc.ID;
            s.x_k_x = asuint(rc.x_k.x); s.x_k_y = asuint(rc.x_k.y); s.x_k_z = asuint(rc.x_k.z);
            s.w_k = w_k_encoded;
            s.L_rg = pack_half2(rc.L.x, rc.L.y);
            s.L_b = zr_f32_to_f16(rc.L.z);
            if (rc.IsCase1())
            {
                s.jacobian_or_seed_nee = asuint(rc.partialJacobian);
                s.meshIdx = rc.meshIdx;
            }
            else if (rc.IsCase2())
            {
                s.jacobian_or_seed_nee = asuint(rc.partialJacobian);
                s.lightPdf = rc.lightPdf; s.dwdA = rc.dwdA; s.seed_nee = rc.seed_nee; s.meshIdx = rc.meshIdx;
            }
            else
            {
                s.jacobian_or_seed_nee = rc.lobe_k_min_1 == BSDF::ALL ? rc.seed_nee : asuint(rc.partialJacobian);
                s.lightPdf = rc.lightPdf;
            }
        }
        // WriteReservoirData: A.x and B only
        void WriteReservoirData(zr_rpt_reservoir& s, uint32_t M_max) const
        {
            uint32_t k = rc.Empty() ? rc.k : (rc.k > 2 ? rc.k : 2) - 2;
            uint32_t m = M < M_max ? M : M_max;
            s.meta = (s.meta & 0xffffff00u) | ((k | (m << 4)) & 0xff);
            s.w_sum = w_sum; s.W = W;
        }
    };

    // Shift.hlsli:360-375
    inline bool CanReconnect(float alpha_lobe_k_min_1, float alpha_lobe_k, LOBE lobe_k_min_1, LOBE lobe_k, float alpha_min)
    {
        if ((alpha_lobe_k_min_1 < alpha_min) || (alpha_lobe_k < alpha_min)) return false;
        if ((lobe_k_min_1 == BSDF::GLOSSY_T) && (lobe_k == BSDF::GLOSSY_T)) return false;
        return true;
    }

    struct DirectLightingEstimate
    {
        float3 ld, le, wi, pos, normal;
        float pdf_solidAngle, dwdA;
        TYPE lt; LOBE lobe; uint32_t ID; float pdf_light; bool twoSided;
        static DirectLightingEstimate Init()
        {
            DirectLightingEstimate r;
            r.ld = f3(0); r.le = f3(0); r.wi = f3(0); r.pdf_solidAngle = 0; r.dwdA = 1; r.lt = Light::NONE;
            r.ID = UINT32_MAX_; r.pos = f3(0); r.pdf_light = 0; r.twoSided = true; r.normal = f3(0); r.lobe = BSDF::DIFFUSE_R;
            return r;
        }
    };

    inline bool IsSpecularSurface(const ShadingData& surface)
    {
        return surface.GlossSpecular() && (surface.metallic || surface.specTr) && (!surface.Coated() || surface.CoatSpecular());
    }

    // ReSTIR_PT_NEE.hlsli:145-222
    inline DirectLightingEstimate NEE_Bsdf(const Scene& sc, float3 pos, float3 normal, const ShadingData& surface, int nextBounce,
        int maxNumBounces, BSDF::BSDFSample& bsdfSample, HitEmissive& hitInfo, RNG& rng)
    {
        DirectLightingEstimate ret = DirectLightingEstimate::Init();
        const bool specular = IsSpecularSurface(surface);
        const int numLightSamples = specular ? 0 : 1;
        if (nextBounce <= maxNumBounces)
            bsdfSample = BSDF::SampleBSDF(normal, surface, rng);
        const float wiPdf = bsdfSample.pdf;
        const float3 wi = bsdfSample.wi;
        const float3 f = bsdfSample.f;
        hitInfo = FindClosestEmissive(sc, pos, normal, wi, surface.Transmissive());
        if (hitInfo.HitWasEmissive())
        {
            const zr_emissive_tri& emissive = sc.emissives[hitInfo.emissiveTriIdx];
            const float3 le = Light::Le_EmissiveTriangle(emissive);
            const float3 vtx0 = Light::Vtx0(emissive);
            const float3 vtx1 = Light::DecodeEmissiveTriV1(emissive);
            const float3 vtx2 = Light::DecodeEmissiveTriV2(emissive);
            float3 lightNormal = cross(vtx1 - vtx0, vtx2 - vtx0);
            float twoArea = length(lightNormal);
            lightNormal = dot(lightNormal, lightNormal) == 0 ? f3(0.0f) : lightNormal / twoArea;
            lightNormal = Light::IsDoubleSided(emissive) && (dot(-wi, lightNormal) < 0) ? -lightNormal : lightNormal;
            float lightPdf = 0;
            if (!specular)
            {
                const float lightSourcePdf = numLightSamples > 0 ? sc.aliasTable[hitInfo.emissiveTriIdx].CachedP_Orig : 0;
                lightPdf = twoArea > 0 ? lightSourcePdf * (2.0f / twoArea) : 0;
            }
            float dwdA = saturate(dot(lightNormal, -wi)) / (hitInfo.t * hitInfo.t);
            float wiPdf_area = wiPdf * dwdA;
            float3 ld = le * f * dwdA;
            ret.ld = specular ? (wiPdf_area > 0 ? ld / wiPdf_area : f3(0)) : RT::PowerHeuristic(wiPdf_area, lightPdf, ld);
            ret.le = le; ret.wi = wi; ret.pdf_solidAngle = wiPdf; ret.dwdA = dwdA; ret.ID = emissive.ID;
            ret.pos = mad(hitInfo.t, wi, pos);
            ret.normal = lightNormal; ret.pdf_light = lightPdf; ret.lobe = bsdfSample.lobe; ret.lt = Light::EMISSIVE;
            ret.twoSided = Light::IsDoubleSided(emissive);
        }
        if (nextBounce >= maxNumBounces)
            bsdfSample.bsdfOverPdf = f3(0);
        return ret;
    }

    // ReSTIR_PT_NEE.hlsli:224-302 (alias-table path)
    inline DirectLightingEstimate NEE_Emissive(const Scene& sc, float3 pos, float3 normal, ShadingData surface, uint32_t sampleSetIdx, RNG& rng)
    {
        DirectLightingEstimate ret = DirectLightingEstimate::Init();
        ret.lt = Light::EMISSIVE;
        ret.lobe = BSDF::ALL;
        const Light::LightSample lightSample = Light::SampleLight(sc, pos, sampleSetIdx, rng, true);
        float3 le = lightSample.le;
        const float lightPdf = lightSample.pdf;
        const uint32_t lightID = lightSample.ID;
        const bool twoSided = lightSample.twoSided;
        const float t = length(lightSample.pos - pos);
        const float3 wi = (lightSample.pos - pos) / t;
        if ((dot(lightSample.normal, -wi) > 0) && (t > 0))
        {
            const float dwdA = saturate(dot(lightSample.normal, -wi)) / (t * t);
            surface.SetWi(wi, normal);
            float3 ld = le * BSDF::Unified(surface).f * dwdA;
            if (dot(ld, ld) > 0)
                ld *= Visibility_Segment(sc, pos, wi, t, normal, lightID, surface.Transmissive()) ? 1.0f : 0.0f;
            float bsdfPdf = 0;
            if (dot(ld, ld) > 0)
            {
                bsdfPdf = BSDF::BSDFSamplerPdf(normal, surface, wi, rng);
                bsdfPdf *= dwdA;
            }
            ret.ld = RT::PowerHeuristic(lightPdf, bsdfPdf, ld);
            ret.le = le; ret.wi = wi; ret.pdf_solidAngle = lightPdf / dwdA; ret.dwdA = dwdA; ret.ID = lightID;
            ret.pos = lightSample.pos; ret.normal = lightSample.normal; ret.pdf_light = lightPdf; ret.twoSided = twoSided;
        }
        return ret;
    }

    // ReSTIR_PT_NEE.hlsli:306-343
    inline DirectLightingEstimate EvalDirect_Emissive_Case2(float3 normal, ShadingData surface, float3 wi, float3 le, float dwdA,
        float lightPdf, LOBE lobe, RNG& rngReplay, RNG& rngNEE)
    {
        surface.SetWi(wi, normal);
        float3 ld = le * BSDF::Unified(surface).f * dwdA;
        DirectLightingEstimate ret = DirectLightingEstimate::Init();
        if (dot(ld, ld) == 0)
            return ret;
        if (lobe == BSDF::ALL)
        {
            rngNEE.Uniform4D();
            float bsdfPdf = BSDF::BSDFSamplerPdf(normal, surface, wi, rngNEE);
            float bsdfPdf_area = bsdfPdf * dwdA;
            ret.ld = RT::PowerHeuristic(lightPdf, bsdfPdf_area, ld);
            ret.pdf_solidAngle = 1.0f;
        }
        else
        {
            BSDF::BSDFSamplerEval eval = BSDF::EvalBSDFSampler(normal, surface, wi, lobe, rngReplay);
            const bool specular = IsSpecularSurface(surface);
            float bsdfPdf_area = eval.pdf * dwdA;
            ret.ld = specular ? (bsdfPdf_area > 0 ? ld / bsdfPdf_area : f3(0)) : RT::PowerHeuristic(bsdfPdf_area, lightPdf, ld);
            ret.pdf_solidAngle = eval.pdf;
        }
        return ret;
    }

    // ReSTIR_PT_NEE.hlsli:345-391
    inline DirectLightingEstimate EvalDirect_Emissive_Case3(const Scene& sc, float3 pos, float3 normal, ShadingData surface, float3 wi,
        float t, float3 le, float3 lightNormal, float lightPdf, uint32_t lightID, bool twoSided, LOBE lobe, RNG& rngReplay, RNG& rngNEE)
    {
        float wiDotLightNormal = dot(lightNormal, -wi);
        float dwdA = fabsf(wiDotLightNormal) / (t * t);
        surface.SetWi(wi, normal);
        float3 ld = (wiDotLightNormal > 0) || twoSided ? le * BSDF::Unified(surface).f * dwdA : f3(0);
        if (dot(ld, ld) > 0)
            ld *= Visibility_Segment(sc, pos, wi, t, normal, lightID, surface.Transmissive()) ? 1.0f : 0.0f;
        DirectLightingEstimate ret = DirectLightingEstimate::Init();
        if (dot(ld, ld) == 0)
            return ret;
        if (lobe == BSDF::ALL)
        {
            rngNEE.Uniform4D();
            float bsdfPdf = BSDF::BSDFSamplerPdf(normal, surface, wi, rngNEE);
            float bsdfPdf_area = bsdfPdf * dwdA;
            ret.ld = RT::PowerHeuristic(lightPdf, bsdfPdf_area, ld);
            ret.pdf_solidAngle = 1.0f;
        }
        else
        {
            BSDF::BSDFSamplerEval eval = BSDF::EvalBSDFSampler(normal, surface, wi, lobe, rngReplay);
            const bool specular = IsSpecularSurface(surface);
            float bsdfPdf_area = eval.pdf * dwdA;
            ret.ld = specular ? (bsdfPdf_area > 0 ? ld / bsdfPdf_area : f3(0)) : RT::PowerHeuristic(bsdfPdf_area, lightPdf, ld);
            ret.pdf_solidAngle = bsdfPdf_area;
        }
        return ret;
    }

    struct OffsetPath { float3 target; float partialJacobian; bool surfKMin1Tramsmissive; };

    // Path context carried from replay to the reconnection step. The reference round-trips it
    // through the r-buffers (RGBA16F + 2 x RGBA32UI + R16UI, Shift.hlsli:191-358); Quantize() applies
    // that storage precision so keeping the context on chip gives the same numbers.
    struct OffsetPathContext
    {
        float3 throughput, pos, normal;
        ShadingData surface;
        float eta_curr, eta_next;
        RNG rngReplay;

        static OffsetPathContext Init()
        {
            OffsetPathContext c;
            c.throughput = f3(0); c.pos = f3(0); c.normal = f3(0);
            c.surface = ShadingData::InitEmpty();
            c.eta_curr = BSDF::ETA_AIR; c.eta_next = BSDF::DEFAULT_ETA_MAT; c.rngReplay.State = 0;
            return c;
        }
        OffsetPathContext Quantize() const
        {
            OffsetPathContext ctx = Init();
            ctx.throughput = f3(to_half(throughput.x), to_half(throughput.y), to_half(throughput.z));
            if (dot(ctx.throughput, ctx.throughput) == 0)
                return ctx;
            ctx.pos = pos;
            ctx.normal = Math::DecodeOct32(Math::EncodeOct32u(normal));
            ctx.eta_curr = mad(Math::UNorm8ToFloat(Math::FloatToUNorm8((eta_curr - 1.0f) / 1.5f)), 1.5f, 1.0f);
            ctx.eta_next = mad(Math::UNorm8ToFloat(Math::FloatToUNorm8((eta_next - 1.0f) / 1.5f)), 1.5f, 1.0f);
            float3 wo = Math::DecodeOct32(Math::EncodeOct32u(surface.wo));
            float roughness = Math::UNorm8ToFloat(Math::FloatToUNorm8(!surface.GlossSpecular() ? sqrtf(surface.alpha) : 0));
            float3 baseColor = Math::UnpackRGB8(Math::Float3ToRGB8(surface.baseColor_Fr0_TrCol));
            bool metallic = surface.metallic;
            bool specTr = surface.specTr;
            float trDepth = surface.trDepth > 0 ? 1.0f : 0.0f;
            bool coated = surface.Coated();
            float subsurface = Math::UNorm8ToFloat(Math::FloatToUNorm8(surface.subsurface));
            float eta_next_ = ctx.eta_curr == BSDF::ETA_AIR ? ctx.eta_next : BSDF::ETA_AIR;
            float coat_weight = 0; float3 coat_color = f3(0.0f); float coat_roughness = 0; float coat_ior = BSDF::DEFAULT_ETA_COAT;
            if (coated)
            {
                coat_weight = Math::UNorm8ToFloat(Math::FloatToUNorm8(surface.coat_weight));
                coat_color = Math::UnpackRGB8(Math::Float3ToRGB8(surface.coat_color));
                coat_roughness = Math::UNorm8ToFloat(Math::FloatToUNorm8(!surface.CoatSpecular() ? sqrtf(surface.coat_alpha) : 0));
                float coat_eta = surface.coat_eta >= 1.0f ? surface.coat_eta : 1.0f / surface.coat_eta;
                coat_ior = mad(Math::UNorm8ToFloat(Math::FloatToUNorm8((coat_eta - 1.0f) / 1.5f)), 1.5f, 1.0f);
            }
            ctx.surface = ShadingData::Init(ctx.normal, wo, metallic, roughness, baseColor, ctx.eta_curr, eta_next_, specTr,
                trDepth, to_half(subsurface), coat_weight, coat_color, coat_roughness, coat_ior);
            return ctx;
        }
    };

    // Shift.hlsli:377-474
    inline void Replay(const Scene& sc, int numBounces, BSDF::BSDFSample bsdfSample, float alpha_min, OffsetPathContext& ctx)
    {
        ctx.throughput = bsdfSample.bsdfOverPdf;
        int bounce = 0;
        ctx.eta_curr = dot(ctx.normal, bsdfSample.wi) < 0 ? ctx.eta_next : BSDF::ETA_AIR;
        bool inTranslucentMedium = ctx.eta_curr != BSDF::ETA_AIR;
        float alpha_lobe_prev = BSDF::LobeAlpha(ctx.surface, bsdfSample.lobe);
        LOBE lobe_prev = bsdfSample.lobe;
        while (true)
        {
            Hit hitInfo = FindClosest(sc, ctx.pos, ctx.normal, bsdfSample.wi, ctx.surface.Transmissive());
            if (!hitInfo.hit) { ctx.throughput = f3(0); return; }
            if (!GetMaterialData(sc, -bsdfSample.wi, ctx.eta_curr, hitInfo, ctx.surface, ctx.eta_next)) { ctx.throughput = f3(0); return; }
            ctx.pos = mad(hitInfo.t, bsdfSample.wi, ctx.pos);
            ctx.normal = hitInfo.normal;
            bounce++;
            if (inTranslucentMedium && (ctx.surface.trDepth > 0))
            {
                float3 c = ctx.surface.baseColor_Fr0_TrCol;
                float3 extCoeff = f3(-zr_logf(c.x), -zr_logf(c.y), -zr_logf(c.z)) / ctx.surface.trDepth;
                ctx.throughput *= f3(zr_expf(-hitInfo.t * extCoeff.x), zr_expf(-hitInfo.t * extCoeff.y), zr_expf(-hitInfo.t * extCoeff.z));
            }
            if (bounce >= numBounces)
                break;
            bsdfSample = BSDF::SampleBSDF(ctx.normal, ctx.surface, ctx.rngReplay);
            if (dot(bsdfSample.bsdfOverPdf, bsdfSample.bsdfOverPdf) == 0) { ctx.throughput = f3(0); return; }
            const float alpha_lobe = BSDF::LobeAlpha(ctx.surface, bsdfSample.lobe);
            if (CanReconnect(alpha_lobe_prev, alpha_lobe, lobe_prev, bsdfSample.lobe, alpha_min)) { ctx.throughput = f3(0); return; }
            const bool transmitted = dot(ctx.normal, bsdfSample.wi) < 0;
            ctx.eta_curr = transmitted ? (ctx.eta_curr == BSDF::ETA_AIR ? ctx.eta_next : BSDF::ETA_AIR) : ctx.eta_curr;
            ctx.throughput *= bsdfSample.bsdfOverPdf;
            inTranslucentMedium = ctx.eta_curr != BSDF::ETA_AIR;
            alpha_lobe_prev = alpha_lobe;
            lobe_prev = bsdfSample.lobe;
        }
    }

    // Shift.hlsli:818-859
    inline OffsetPathContext Replay_kGt2(const Scene& sc, float3 pos, float3 normal, float ior, const ShadingData& surface,
        const Reconnection& rc, float alpha_min)
    {
        OffsetPathContext ctx = OffsetPathContext::Init();
        ctx.pos = pos; ctx.normal = normal; ctx.surface = surface;
        ctx.rngReplay = RNG::InitSeed(rc.seed_replay);
        ctx.eta_curr = BSDF::ETA_AIR; ctx.eta_next = ior;
        ctx.throughput = f3(1);
        const int numBounces = (int)rc.k - 2;
        BSDF::BSDFSample bsdfSample = BSDF::SampleBSDF(ctx.normal, ctx.surface, ctx.rngReplay);
        if (dot(bsdfSample.bsdfOverPdf, bsdfSample.bsdfOverPdf) == 0) { ctx.throughput = f3(0); return ctx; }
        Replay(sc, numBounces, bsdfSample, alpha_min, ctx);
        return ctx;
    }

    // Shift.hlsli:476-546
    inline float StepPath(const Scene& sc, OffsetPathContext& ctx, float alpha_min, const Reconnection& rc)
    {
        if (!BSDF::IsLobeValid(ctx.surface, rc.lobe_k_min_1))
            return 0;
        float alpha_lobe_k_min_1 = BSDF::LobeAlpha(ctx.surface, rc.lobe_k_min_1);
        if (!CanReconnect(alpha_lobe_k_min_1, 1, rc.lobe_k_min_1, rc.lobe_k, alpha_min))
            return 0;
        float3 w_k_min_1 = normalize(rc.x_k - ctx.pos);
        BSDF::BSDFSamplerEval eval = BSDF::EvalBSDFSampler(ctx.normal, ctx.surface, w_k_min_1, rc.lobe_k_min_1, ctx.rngReplay);
        if (dot(eval.bsdfOverPdf, eval.bsdfOverPdf) == 0)
            return 0;
        Hit hitInfo = FindClosest(sc, ctx.pos, ctx.normal, w_k_min_1, ctx.surface.Transmissive());
        if (!hitInfo.hit || (hitInfo.ID != rc.ID))
            return 0;
        const float3 y_k = mad(hitInfo.t, w_k_min_1, ctx.pos);
        const bool transmitted = dot(ctx.normal, w_k_min_1) < 0;
        ctx.eta_curr = transmitted ? (ctx.eta_curr == BSDF::ETA_AIR ? ctx.eta_next : BSDF::ETA_AIR) : ctx.eta_curr;
        const bool inTranslucentMedium = ctx.eta_curr != BSDF::ETA_AIR;
        if (!GetMaterialData(sc, -w_k_min_1, ctx.eta_curr, hitInfo, ctx.surface, ctx.eta_next))
            return 0;
        if (inTranslucentMedium && (ctx.surface.trDepth > 0))
        {
            float3 c = ctx.surface.baseColor_Fr0_TrCol;
            float3 extCoeff = f3(-zr_logf(c.x), -zr_logf(c.y), -zr_logf(c.z)) / ctx.surface.trDepth;
            ctx.throughput *= f3(zr_expf(-hitInfo.t * extCoeff.x), zr_expf(-hitInfo.t * extCoeff.y), zr_expf(-hitInfo.t * extCoeff.z));
        }
        float partialJacobian = eval.pdf;
        partialJacobian *= fabsf(dot(-w_k_min_1, hitInfo.normal));
        partialJacobian /= (hitInfo.t * hitInfo.t);
        ctx.pos = y_k;
        ctx.normal = hitInfo.normal;
        ctx.throughput *= eval.bsdfOverPdf;
        return partialJacobian;
    }

    // Shift.hlsli:662-816 (Emissive == true). `replayed` = context from Replay_kGt2 (already quantised) when k > 2.
    inline OffsetPath Shift2(const Scene& sc, float3 pos, float3 normal, float ior, const ShadingData& surface, const Reconnection& rc,
        const OffsetPathContext* replayed, float alpha_min)
    {
        OffsetPathContext ctx = OffsetPathContext::Init();
        ctx.pos = pos; ctx.normal = normal; ctx.surface = surface;
        ctx.rngReplay = RNG::InitSeed(rc.seed_replay);
        ctx.eta_curr = BSDF::ETA_AIR; ctx.eta_next = ior;
        ctx.throughput = f3(1);
        OffsetPath ret; ret.target = f3(0); ret.partialJacobian = 0; ret.surfKMin1Tramsmissive = false;
        const int numBounces = (int)rc.k - 2;
        if (numBounces != 0)
        {
            ctx = *replayed;
            if (dot(ctx.throughput, ctx.throughput) == 0)
                return ret;
            // OffsetPathContext::Load leaves rngReplay at 0; the reference then advances it (Shift.hlsli:705-713)
            ctx.rngReplay.State = 0;
            for (int bounce = 0; bounce < numBounces; bounce++)
            {
                ctx.rngReplay.Uniform4D();
                ctx.rngReplay.Uniform4D();
                ctx.rngReplay.Uniform();
            }
        }
        ret.surfKMin1Tramsmissive = ctx.surface.specTr;
        if (!rc.IsCase3())
        {
            ret.partialJacobian = StepPath(sc, ctx, alpha_min, rc);
            if (ret.partialJacobian == 0)
                return ret;
            if (rc.IsCase1())
            {
                float3 w_k = rc.w_k_lightNormal_w_sky;
                BSDF::BSDFSamplerEval eval = BSDF::EvalBSDFSampler(ctx.normal, ctx.surface, w_k, rc.lobe_k, ctx.rngReplay);
                ctx.throughput *= eval.bsdfOverPdf;
                ret.target = ctx.throughput * rc.L;
                ret.partialJacobian *= eval.pdf;
                return ret;
            }
        }
        else
        {
            if (!BSDF::IsLobeValid(ctx.surface, rc.lobe_k_min_1))
                return ret;
            float alpha_lobe_k_min_1 = BSDF::LobeAlpha(ctx.surface, rc.lobe_k_min_1);
            if (alpha_lobe_k_min_1 < alpha_min)
                return ret;
        }
        RNG rngNEE = RNG::InitSeed(rc.seed_nee);
        if (rc.IsCase2())
        {
            float3 w_k = rc.w_k_lightNormal_w_sky;
            DirectLightingEstimate ls = EvalDirect_Emissive_Case2(ctx.normal, ctx.surface, w_k, rc.L, rc.dwdA, rc.lightPdf,
                rc.lobe_k, ctx.rngReplay, rngNEE);
            ret.target = ctx.throughput * ls.ld;
            ret.partialJacobian *= ls.pdf_solidAngle;
        }
        else
        {
            float3 wi_k_min_1 = rc.x_k - ctx.pos;
            float t = length(wi_k_min_1);
            wi_k_min_1 /= t;
            float3 lightNormal = rc.w_k_lightNormal_w_sky;
            bool twoSided = rc.lightPdf > 0;
            // note: the reference passes ctx.pos for the normal argument (Shift.hlsli:765-767)
            DirectLightingEstimate ls = EvalDirect_Emissive_Case3(sc, ctx.pos, ctx.pos, ctx.surface, wi_k_min_1, t, rc.L, lightNormal,
                fabsf(rc.lightPdf), rc.ID, twoSided, rc.lobe_k_min_1, ctx.rngReplay, rngNEE);
            ret.target = ctx.throughput * ls.ld;
            ret.partialJacobian = ls.pdf_solidAngle;
        }
        return ret;
    }
}
} // namespace orc
