// ORACLE -- test infrastructure, not product code (see orc_math.h header).
//
// Scene access + ray queries for the CPU restatement:
//   ZetaCore/Core/Material.h:296-427          material getters
//   ZetaRenderPass/Common/RayQuery.hlsli      Hit::FindClosest :15-144, Hit_Emissive :146-299,
//                                             Visibility_Segment :337-406, GetMaterialData :452-524
//   ZetaRenderPass/Common/RT.hlsli:245-262    OffsetRayRTG
//   ZetaRenderPass/Common/LightSource.hlsli   emissive triangle decode/sample/Le :46-137, 202-224
// The reference traverses a driver-opaque DXR TLAS; here closest/any hits are found by BRUTE FORCE
// over all world-space triangles with the same ray/triangle test and tie rule the product's BVH
// kernel uses (closest t, then lowest global triangle index), so the BVH is checked independently.
#pragma once
#include "orc_bsdf.h"
#include "orc_gbuffer.h"
#include "../include/zr_abi.h"
#include <vector>

namespace orc
{
constexpr uint32_t UINT32_MAX_ = 0xffffffffu;
constexpr uint32_t MAT_INVALID_ID = 0xffff;

namespace Mat
{
    inline bool DoubleSided(const zr_material& m) { return m.CoatColor_Flags & (1u << 25); }
    inline bool Metallic(const zr_material& m) { return m.CoatColor_Flags & (1u << 24); }
    inline bool Transmissive(const zr_material& m) { return m.CoatColor_Flags & (1u << 26); }
    inline bool ThinWalled(const zr_material& m) { return m.CoatColor_Flags & (1u << 29); }
    inline float3 GetBaseColorFactor(const zr_material& m) { return Math::UnpackRGB8(m.BaseColorFactor); }
    inline float3 GetCoatColor(const zr_material& m) { return Math::UnpackRGB8(m.CoatColor_Flags); }
    inline float3 GetEmissiveFactor(const zr_material& m) { return Math::UnpackRGB8(m.EmissiveFactor_NormalScale); }
    inline float GetCoatIOR(const zr_material& m)
    {
        uint32_t b = (m.EmissiveTex_AlphaCutoff_CoatIOR >> 24) & 0xff;
        return mad(1.5f / 255.0f, (float)b, 1.0f);
    }
    inline float GetSpecularRoughness(const zr_material& m) { return Math::UNorm8ToFloat((m.MRTex_SpecRoughness_CoatRoughness >> 16) & 0xff); }
    inline float GetCoatRoughness(const zr_material& m) { return Math::UNorm8ToFloat((m.MRTex_SpecRoughness_CoatRoughness >> 24) & 0xff); }
    inline float GetEmissiveStrength(const zr_material& m) { return zr_f16_to_f32((uint16_t)(m.EmissiveStrength_IOR & 0xffff)); }
    inline float GetSpecularIOR(const zr_material& m)
    {
        uint16_t encoded = (uint16_t)(m.EmissiveStrength_IOR >> 16);
        return mad(1.5f / 65535.0f, (float)encoded, 1.0f);
    }
    inline float GetTransmissionDepth(const zr_material& m) { return zr_f16_to_f32((uint16_t)(m.NormalTex_TrDepth >> 16)); }
    inline float GetSubsurface(const zr_material& m) { return Math::UNorm8ToFloat((m.BaseColorTex_Subsurf_CoatWeight >> 16) & 0xff); }
    inline float GetCoatWeight(const zr_material& m) { return Math::UNorm8ToFloat((m.BaseColorTex_Subsurf_CoatWeight >> 24) & 0xff); }
}

struct RayHit { bool hit; float t; float2 bary; uint32_t tri; };

struct Scene
{
    const zr_vertex* vertices = nullptr;
    const uint32_t* indices = nullptr;
    const zr_mesh_instance* instances = nullptr;
    uint32_t numInstances = 0;
    const zr_material* materials = nullptr;
    const zr_emissive_tri* emissives = nullptr;
    uint32_t numEmissives = 0;
    const zr_alias_entry* aliasTable = nullptr;
    // presampled emissive sets (PresampleEmissives.hlsl); sampleSetSize == 0: alias-table sampling
    const zr_presampled_tri* sampleSets = nullptr;
    uint32_t numSampleSets = 0, sampleSetSize = 0;
    // light voxel grid (BuildLightVoxelGrid.hlsl); lvg == nullptr: off
    const zr_voxel_sample* lvg = nullptr;
    uint32_t lvgDim[3] = { 0, 0, 0 };
    float lvgExtents[3] = { 0, 0, 0 };
    float lvgOffsetY = 0;
    // derived
    std::vector<float3> v0, e1, e2;
    std::vector<uint32_t> triMesh, triPrim, meshFirstTri;

    static float3 h3(const uint16_t h[3]) { return f3(zr_f16_to_f32(h[0]), zr_f16_to_f32(h[1]), zr_f16_to_f32(h[2])); }

    void Build(const uint32_t* instNumTris)
    {
        uint32_t total = 0;
        for (uint32_t m = 0; m < numInstances; m++) total += instNumTris[m];
        v0.resize(total); e1.resize(total); e2.resize(total); triMesh.resize(total); triPrim.resize(total);
        meshFirstTri.resize(numInstances);
        uint32_t g = 0;
        for (uint32_t m = 0; m < numInstances; m++)
        {
            const zr_mesh_instance& md = instances[m];
            meshFirstTri[m] = g;
            float4 q = normalize(Math::DecodeNormalized4(md.Rotation));
            float3 s = h3(md.Scale);
            float3 t = f3(md.Translation[0], md.Translation[1], md.Translation[2]);
            for (uint32_t p = 0; p < instNumTris[m]; p++, g++)
            {
                uint32_t tri = p * 3 + md.BaseIdxOffset;
                float3 pw[3];
                for (int k = 0; k < 3; k++)
                {
                    const zr_vertex& V = vertices[indices[tri + k] + md.BaseVtxOffset];
                    pw[k] = Math::TransformTRS(f3(V.pos[0], V.pos[1], V.pos[2]), t, q, s);
                }
                v0[g] = pw[0]; e1[g] = pw[1] - pw[0]; e2[g] = pw[2] - pw[0];
                triMesh[g] = m; triPrim[g] = p;
            }
        }
    }

    static bool TriHit(float3 o, float3 d, float3 v0, float3 e1, float3 e2, float tmin, float tmax, float& t, float& u, float& v)
    {
        float3 pvec = cross(d, e2);
        float det = dot(e1, pvec);
        if (det == 0.0f) return false;
        float inv = 1.0f / det;
        float3 tvec = o - v0;
        u = dot(tvec, pvec) * inv;
        if (!(u >= 0.0f) || u > 1.0f) return false;
        float3 qvec = cross(tvec, e1);
        v = dot(d, qvec) * inv;
        if (!(v >= 0.0f) || u + v > 1.0f) return false;
        t = dot(e2, qvec) * inv;
        return t > tmin && t < tmax;
    }

    RayHit Closest(float3 o, float3 d, float tmin, float tmax) const
    {
        RayHit r{ false, tmax, f2(0, 0), UINT32_MAX_ };
        for (uint32_t i = 0; i < (uint32_t)v0.size(); i++)
        {
            float t, u, v;
            if (TriHit(o, d, v0[i], e1[i], e2[i], tmin, tmax, t, u, v))
            {
                if (!r.hit || t < r.t)      // ascending i: ties keep the lowest index
                {
                    r.hit = true; r.t = t; r.bary = f2(u, v); r.tri = i;
                }
            }
        }
        return r;
    }

    uint32_t TriID(uint32_t tri) const { return RNG::PCG3d(uint3{ triMesh[tri], 0u, triPrim[tri] }).x; }

    // true if some triangle other than `ignoreID` is hit in (tmin, tmax)
    bool AnyHitExcept(float3 o, float3 d, float tmin, float tmax, uint32_t ignoreID) const
    {
        for (uint32_t i = 0; i < (uint32_t)v0.size(); i++)
        {
            float t, u, v;
            if (TriHit(o, d, v0[i], e1[i], e2[i], tmin, tmax, t, u, v))
            {
                if (ignoreID == UINT32_MAX_ || TriID(i) != ignoreID)
                    return true;
            }
        }
        return false;
    }
};

namespace RTU
{
    // RT.hlsli:245-262
    inline float3 OffsetRayRTG(float3 pos, float3 geometricNormal)
    {
        const float origin = 1.0f / 32.0f;
        const float float_scale = 1.0f / 65536.0f;
        const float int_scale = 256.0f;
        int ofx = (int)(int_scale * geometricNormal.x), ofy = (int)(int_scale * geometricNormal.y), ofz = (int)(int_scale * geometricNormal.z);
        auto adj = [](float p, int of) { return asfloat((uint32_t)((int)asuint(p) + ((p < 0) ? -of : of))); };
        float3 p_i = f3(adj(pos.x, ofx), adj(pos.y, ofy), adj(pos.z, ofz));
        return f3(fabsf(pos.x) < origin ? pos.x + float_scale * geometricNormal.x : p_i.x,
                  fabsf(pos.y) < origin ? pos.y + float_scale * geometricNormal.y : p_i.y,
                  fabsf(pos.z) < origin ? pos.z + float_scale * geometricNormal.z : p_i.z);
    }
}

constexpr float T_MIN_REFL_RAY = 1e-6f;
constexpr float T_MIN_TR_RAY = 5e-5f;

struct Hit
{
    bool hit; float t; float2 uv; float3 normal; uint32_t ID; uint32_t meshIdx; uint32_t matIdx;
};

struct HitEmissive
{
    bool hit; float t; uint32_t geoIdx, primIdx, emissiveTriIdx; float2 bary; float3 lightPos;
    bool HitWasEmissive() const { return emissiveTriIdx != UINT32_MAX_; }
};

// RayQuery.hlsli:213-299 (ToHitInfo) == the attribute part of Hit::FindClosest
inline Hit HitAttributes(const Scene& sc, uint32_t meshIdx, uint32_t primIdx, float2 bary, float t)
{
    Hit ret;
    const zr_mesh_instance& meshData = sc.instances[meshIdx];
    ret.hit = true;
    ret.t = t;
    ret.matIdx = meshData.MatIdx;
    ret.meshIdx = meshIdx;
    uint32_t tri = primIdx * 3 + meshData.BaseIdxOffset;
    const zr_vertex& V0 = sc.vertices[sc.indices[tri] + meshData.BaseVtxOffset];
    const zr_vertex& V1 = sc.vertices[sc.indices[tri + 1] + meshData.BaseVtxOffset];
    const zr_vertex& V2 = sc.vertices[sc.indices[tri + 2] + meshData.BaseVtxOffset];
    float4 q = normalize(Math::DecodeNormalized4(meshData.Rotation));
    float3 s = Scene::h3(meshData.Scale);
    float tmp = 1 - bary.x - bary.y;
    float2 uv = f2(mad(bary.y, V2.uv[0], tmp * V0.uv[0]), mad(bary.y, V2.uv[1], tmp * V0.uv[1]));
    uv = f2(mad(bary.x, V1.uv[0], uv.x), mad(bary.x, V1.uv[1], uv.y));
    ret.uv = uv;
    float3 v0_n = Math::DecodeOct32((uint32_t)V0.normal[0] | ((uint32_t)V0.normal[1] << 16));
    float3 v1_n = Math::DecodeOct32((uint32_t)V1.normal[0] | ((uint32_t)V1.normal[1] << 16));
    float3 v2_n = Math::DecodeOct32((uint32_t)V2.normal[0] | ((uint32_t)V2.normal[1] << 16));
    float3 hitNormal = mad(bary.y, v2_n, tmp * v0_n);
    hitNormal = mad(bary.x, v1_n, hitNormal);
    const float3 scaleInv = 1.0f / s;
    hitNormal *= scaleInv;
    hitNormal = Math::RotateVector(hitNormal, q);
    hitNormal = normalize(hitNormal);
    ret.normal = hitNormal;
    ret.ID = RNG::PCG3d(uint3{ meshIdx, 0u, primIdx }).x;
    return ret;
}

// Hit_Emissive::FindClosest (RayQuery.hlsli:148-205)
inline HitEmissive FindClosestEmissive(const Scene& sc, float3 pos, float3 normal, float3 wi, bool transmissive)
{
    HitEmissive ret;
    ret.hit = false;
    ret.emissiveTriIdx = UINT32_MAX_;
    ret.t = 0; ret.geoIdx = 0; ret.primIdx = 0; ret.bary = f2(0, 0); ret.lightPos = f3(0);
    bool wiBackface = dot(normal, wi) <= 0;
    if (wiBackface)
    {
        if (transmissive) normal = -normal;
        else return ret;
    }
    const float3 adjustedOrigin = RTU::OffsetRayRTG(pos, normal);
    RayHit h = sc.Closest(adjustedOrigin, wi, wiBackface ? T_MIN_TR_RAY : T_MIN_REFL_RAY, FLT_MAX_);
    if (h.hit)
    {
        ret.hit = true;
        ret.bary = h.bary;
        ret.t = h.t;
        ret.geoIdx = sc.triMesh[h.tri];
        ret.primIdx = sc.triPrim[h.tri];
        const zr_mesh_instance& meshData = sc.instances[ret.geoIdx];
        if (meshData.BaseEmissiveTriOffset == UINT32_MAX_)
            return ret;
        ret.emissiveTriIdx = meshData.BaseEmissiveTriOffset + ret.primIdx;
        ret.lightPos = mad(h.t, wi, adjustedOrigin);
    }
    return ret;
}

// Hit::FindClosest<ID, Curr> (RayQuery.hlsli:17-131)
inline Hit FindClosest(const Scene& sc, float3 pos, float3 normal, float3 wi, bool transmissive)
{
    Hit ret;
    ret.hit = false;
    ret.ID = UINT32_MAX_;
    ret.t = 0; ret.uv = f2(0, 0); ret.normal = f3(0); ret.meshIdx = 0; ret.matIdx = 0;
    float ndotwi = dot(normal, wi);
    if (ndotwi == 0)
        return ret;
    bool wiBackface = ndotwi < 0;
    if (wiBackface)
    {
        if (!transmissive) return ret;
        normal = -normal;
    }
    const float3 adjustedOrigin = RTU::OffsetRayRTG(pos, normal);
    RayHit h = sc.Closest(adjustedOrigin, wi, wiBackface ? T_MIN_TR_RAY : T_MIN_REFL_RAY, FLT_MAX_);
    if (h.hit)
        ret = HitAttributes(sc, sc.triMesh[h.tri], sc.triPrim[h.tri], h.bary, h.t);
    return ret;
}

// RayQuery.hlsli:337-406 with APPROXIMATE_EMISSIVE_SHADOW_RAY == 1
inline bool Visibility_Segment(const Scene& sc, float3 origin, float3 wi, float rayT, float3 normal, uint32_t triID,
    bool transmissive)
{
    if (triID == UINT32_MAX_) return false;
    if (rayT < 1e-6f) return false;
    float ndotwi = dot(normal, wi);
    if (ndotwi == 0) return false;
    bool wiBackface = ndotwi < 0;
    if (wiBackface)
    {
        if (transmissive) normal = -normal;
        else return false;
    }
    const float3 adjustedOrigin = RTU::OffsetRayRTG(origin, normal);
    const float tMin = 3e-6f;
    const float tMax = Math::PrevFloat32(rayT * 0.999f - Math::NextFloat32(tMin));
    return !sc.AnyHitExcept(adjustedOrigin, wi, tMin, tMax, triID);
}

// RayQuery.hlsli:337-406 with APPROXIMATE_EMISSIVE_SHADOW_RAY == 0 (the plain path tracer's setting,
// IndirectLighting/PathTracer/Params.hlsli:27): tMax = rayT, the committed hit is the closest one, and the light is visible
// iff nothing is hit or the closest hit is the light itself.
inline bool Visibility_Segment_Precise(const Scene& sc, float3 origin, float3 wi, float rayT, float3 normal, uint32_t triID,
    bool transmissive)
{
    if (triID == UINT32_MAX_) return false;
    if (rayT < 1e-6f) return false;
    float ndotwi = dot(normal, wi);
    if (ndotwi == 0) return false;
    bool wiBackface = ndotwi < 0;
    if (wiBackface)
    {
        if (transmissive) normal = -normal;
        else return false;
    }
    const float3 adjustedOrigin = RTU::OffsetRayRTG(origin, normal);
    const RayHit h = sc.Closest(adjustedOrigin, wi, 3e-6f, rayT);
    if (h.hit)
        return triID == sc.TriID(h.tri);
    return true;
}

// GetMaterialData (RayQuery.hlsli:452-510), textures unsupported (factors only)
inline bool GetMaterialData(const Scene& sc, float3 wo, float eta_curr, Hit& hitInfo, BSDF::ShadingData& surface, float& eta)
{
    const zr_material& mat = sc.materials[hitInfo.matIdx];
    const bool hitBackface = dot(wo, hitInfo.normal) < 0;
    eta = BSDF::DEFAULT_ETA_MAT;
    if (!Mat::DoubleSided(mat) && hitBackface)
        return false;
    if (Mat::DoubleSided(mat) && hitBackface)
        hitInfo.normal = -hitInfo.normal;
    float3 baseColor = Mat::GetBaseColorFactor(mat);
    float metallic = Mat::Metallic(mat) ? 1.0f : 0.0f;
    float roughness = Mat::GetSpecularRoughness(mat);
    bool tr = Mat::Transmissive(mat);
    eta = Mat::GetSpecularIOR(mat);
    float trDepth = tr ? Mat::GetTransmissionDepth(mat) : 0;
    float eta_next = eta_curr == BSDF::ETA_AIR ? eta : BSDF::ETA_AIR;
    float subsurface = Mat::ThinWalled(mat) ? to_half(Mat::GetSubsurface(mat)) : 0;
    float coat_weight = Mat::GetCoatWeight(mat);
    float3 coat_color = Mat::GetCoatColor(mat);
    float coat_roughness = Mat::GetCoatRoughness(mat);
    float coat_ior = Mat::GetCoatIOR(mat);
    surface = BSDF::ShadingData::Init(hitInfo.normal, wo, metallic >= 0.9f, roughness, baseColor, eta_curr, eta_next, tr,
        trDepth, subsurface, coat_weight, coat_color, coat_roughness, coat_ior);
    return true;
}

namespace Light
{
    enum TYPE : uint32_t { NONE = 0, SUN = 1, SKY = 2, EMISSIVE = 3 };
    inline TYPE TypeFromValue(uint32_t x) { return x <= 2 ? (TYPE)x : EMISSIVE; }

    inline float3 DecodeEmissiveTriV1(const zr_emissive_tri& tri)
    {
        float2 v = f2((float)tri.V0V1[0] / 65535.0f, (float)tri.V0V1[1] / 65535.0f);
        float3 decoded = Math::DecodeUnitVector(v);
        return mad(decoded, zr_f16_to_f32(tri.EdgeLengths[0]), f3(tri.Vtx0[0], tri.Vtx0[1], tri.Vtx0[2]));
    }
    inline float3 DecodeEmissiveTriV2(const zr_emissive_tri& tri)
    {
        float2 v = f2((float)tri.V0V2[0] / 65535.0f, (float)tri.V0V2[1] / 65535.0f);
        float3 decoded = Math::DecodeUnitVector(v);
        return mad(decoded, zr_f16_to_f32(tri.EdgeLengths[1]), f3(tri.Vtx0[0], tri.Vtx0[1], tri.Vtx0[2]));
    }
    inline bool IsDoubleSided(const zr_emissive_tri& tri) { return tri.PackedA & (1u << 25); }
    inline float3 Vtx0(const zr_emissive_tri& tri) { return f3(tri.Vtx0[0], tri.Vtx0[1], tri.Vtx0[2]); }

    // Le_EmissiveTriangle (LightSource.hlsli:202-224), no emissive textures
    inline float3 Le_EmissiveTriangle(const zr_emissive_tri& tri)
    {
        const float3 emissiveFactor = Math::UnpackRGB8(tri.PackedA);
        const float emissiveStrength = zr_f16_to_f32((uint16_t)(tri.PackedB >> 16));
        float3 le = emissiveFactor * emissiveStrength;
        if (Math::Luminance(le) == 0)
            return f3(0.0f);
        return le;
    }

    struct AliasTableSample { uint32_t idx; float pdf; };
    inline AliasTableSample SampleAlias(const zr_alias_entry* table, uint32_t numEmissiveTriangles, RNG& rng)
    {
        AliasTableSample ret;
        uint32_t u0 = rng.UniformUintBounded(numEmissiveTriangles);
        zr_alias_entry s = table[u0];
        if (rng.Uniform() < s.P_Curr) { ret.pdf = s.CachedP_Orig; ret.idx = u0; return ret; }
        ret.pdf = s.CachedP_Alias;
        ret.idx = s.Alias;
        return ret;
    }

    struct EmissiveTriSample { float3 pos, normal; float2 bary; float pdf; };
    inline EmissiveTriSample SampleEmissiveTri(float3 pos, const zr_emissive_tri& tri, RNG& rng, bool reverseNormalIfTwoSided = true)
    {
        EmissiveTriSample ret;
        float2 u = rng.Uniform2D();
        ret.bary = Sampling::UniformSampleTriangle(u);
        const float3 vtx0 = Vtx0(tri);
        const float3 vtx1 = DecodeEmissiveTriV1(tri);
        const float3 vtx2 = DecodeEmissiveTriV2(tri);
        ret.pos = (1.0f - ret.bary.x - ret.bary.y) * vtx0 + ret.bary.x * vtx1 + ret.bary.y * vtx2;
        ret.normal = cross(vtx1 - vtx0, vtx2 - vtx0);
        bool normalIs0 = dot(ret.normal, ret.normal) == 0;
        float twoArea = length(ret.normal);
        ret.pdf = normalIs0 ? 0.0f : 2.0f / twoArea;
        ret.normal = normalIs0 ? ret.normal : ret.normal / twoArea;
        ret.normal = reverseNormalIfTwoSided && IsDoubleSided(tri) && dot(pos - ret.pos, ret.normal) < 0 ? -ret.normal : ret.normal;
        return ret;
    }

    // One NEE light sample from a presampled set (ReSTIR_PT_NEE.hlsli:217-236, ReSTIR_DI_Temporal.hlsl:119-136,
    // LightSource.hlsli:99-106) or from alias table + uniform point on the triangle (the #else branches there).
    struct LightSample { float3 pos, normal, le; float2 bary; float pdf; uint32_t idx, ID; bool twoSided; };
    inline LightSample SampleLight(const Scene& sc, float3 pos, uint32_t sampleSetIdx, RNG& rng, bool advanceRng)
    {
        LightSample ls;
        if (sc.sampleSetSize > 0)
        {
            const uint32_t u = rng.UniformUintBounded_Faster(sc.sampleSetSize);
            const zr_presampled_tri& t = sc.sampleSets[(size_t)sampleSetIdx * sc.sampleSetSize + u];
            ls.pos = f3(t.pos[0], t.pos[1], t.pos[2]);
            ls.normal = Math::DecodeOct32(t.normal);
            ls.bary = Math::DecodeUNorm2(t.bary);
            ls.le = f3(zr_f16_to_f32(t.le[0]), zr_f16_to_f32(t.le[1]), zr_f16_to_f32(t.le[2]));
            ls.pdf = t.pdf; ls.idx = t.idx; ls.ID = t.ID; ls.twoSided = t.twoSided != 0;
            if (ls.twoSided && dot(pos - ls.pos, ls.normal) < 0)
                ls.normal = -ls.normal;
            if (advanceRng)
                rng.Uniform3D();
        }
        else
        {
            AliasTableSample entry = SampleAlias(sc.aliasTable, sc.numEmissives, rng);
            const zr_emissive_tri& tri = sc.emissives[entry.idx];
            EmissiveTriSample ts = SampleEmissiveTri(pos, tri, rng);
            ls.pos = ts.pos; ls.normal = ts.normal; ls.bary = ts.bary;
            ls.le = Le_EmissiveTriangle(tri);
            ls.pdf = entry.pdf * ts.pdf; ls.idx = entry.idx; ls.ID = tri.ID; ls.twoSided = IsDoubleSided(tri);
        }
        return ls;
    }
}
// Common/LightVoxelGrid.hlsli:8-69
namespace LVG
{
    inline uint32_t FlattenVoxelIndex(uint32_t x, uint32_t y, uint32_t z, uint32_t dx, uint32_t dy) { return z * dx * dy + y * dx + x; }

    inline float3 VoxelCenter(int vx, int vy, int vz, int dx, int dy, int dz, float3 voxelExtents, const float viewInv[3][4], float offset_y)
    {
        const int hx = dx >> 1, hy = dy >> 1, hz = dz >> 1;
        int cx = vx - hx, cy = vy - hy, cz = vz - hz;
        cx += vx < hx ? 1 : 0; cy += vy < hy ? 1 : 0; cz += vz < hz ? 1 : 0;
        cy *= -1;       // voxel space Y points in the opposite direction of camera space Y
        const float3 corner = f3((float)(cx * 2), (float)(cy * 2), (float)(cz * 2)) * voxelExtents;
        const float3 s = f3(Math::SignNotZero((float)cx), Math::SignNotZero((float)cy), Math::SignNotZero((float)cz));
        float3 centerV = corner + voxelExtents * s;
        centerV.y += offset_y;
        return Math::mul3x4(viewInv, centerV);
    }

    inline bool MapPosToVoxel(float3 pos, int dx, int dy, int dz, float3 voxelExtents, const float view[3][4], int& ox, int& oy, int& oz, float offset_y)
    {
        float3 posV = Math::mul3x4(view, pos);
        posV.y -= offset_y;
        const int hx = dx >> 1, hy = dy >> 1, hz = dz >> 1;
        float3 voxel = f3(floorf(fabsf(posV.x) / (2 * voxelExtents.x)), floorf(fabsf(posV.y) / (2 * voxelExtents.y)), floorf(fabsf(posV.z) / (2 * voxelExtents.z)));
        if (voxel.x >= (float)hx || voxel.y >= (float)hy || voxel.z >= (float)hz)
            return false;
        voxel = voxel * f3(Math::SignNotZero(posV.x), Math::SignNotZero(posV.y), Math::SignNotZero(posV.z));
        voxel.y *= -1;
        ox = (int)voxel.x + hx - (posV.x < 0 ? 1 : 0);
        oy = (int)voxel.y + hy - (posV.y >= 0 ? 1 : 0);
        oz = (int)voxel.z + hz - (posV.z < 0 ? 1 : 0);
        return true;
    }
}

namespace LVG
{
    // LightVoxelGrid.hlsli:54-68
    inline bool Sample(const Scene& sc, float3 pos, float3 voxelExtents, float offset_y, const float view[3][4], zr_voxel_sample& s, RNG& rng)
    {
        const float3 u = rng.Uniform3D();
        const float3 posJittered = pos + (u * 2.0f - 1.0f) * voxelExtents;
        int vx, vy, vz;
        if (!MapPosToVoxel(posJittered, (int)sc.lvgDim[0], (int)sc.lvgDim[1], (int)sc.lvgDim[2], voxelExtents, view, vx, vy, vz, offset_y))
            return false;
        const uint32_t start = FlattenVoxelIndex((uint32_t)vx, (uint32_t)vy, (uint32_t)vz, sc.lvgDim[0], sc.lvgDim[1]) * 64u;
        const uint32_t k = rng.UniformUintBounded_Faster(64u);
        s = sc.lvg[start + k];
        return true;
    }
}
} // namespace orc
