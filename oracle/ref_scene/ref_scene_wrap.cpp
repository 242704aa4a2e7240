// TEST INFRASTRUCTURE. C entry points around the REFERENCE's own scene-data constructors, compiled from the reference tree where it
// lies (oracle/ref_scene/build.sh -> oracle/_ref/libref_scene.so). They pin this repository's scene-ingest packers
// (zetaray_b200/scene.py) -- the flat-buffer formats on the caller's side of the drop-in boundary:
//   ZetaCore/Core/Material.h           Material ctor + setters (32 B)
//   ZetaCore/RayTracing/RtCommon.h     RT::EmissiveTriangle ctor / StoreVertices (48 B), RT::MeshInstance field types
//   ZetaCore/Math/OctahedralVector.h   oct32
//   ZetaCore/Math/Vector.h             half, half3, unorm4
//   ZetaCore/Math/Color.h              Float3ToRGB8
// Nothing here is product code and no reference source is copied.
#include "ZetaCore/RayTracing/RtCommon.h"
#include "ZetaCore/Core/Vertex.h"
#include <cstring>

using namespace ZetaRay;
using namespace ZetaRay::Math;

extern "C"
{
    int ref_scene_sizes(int* out)
    {
        out[0] = (int)sizeof(Material); out[1] = (int)sizeof(RT::MeshInstance); out[2] = (int)sizeof(RT::EmissiveTriangle);
        out[3] = (int)sizeof(Core::Vertex);
        return 4;
    }

    // p: base rgba, metallic, roughness, ior, transmission, emissive rgb, emissive strength, coat weight, coat rgb, coat roughness,
    //    coat ior, subsurface, transmission depth (19 floats); flags: bit0 double sided, bit1 thin walled
    void ref_material(const float* p, uint32_t flags, void* out32)
    {
        Material m;
        m.SetBaseColorFactor(float4(p[0], p[1], p[2], p[3]));
        m.SetMetallic(p[4]);
        m.SetSpecularRoughness(p[5]);
        m.SetSpecularIOR(p[6]);
        m.SetTransmission(p[7]);
        m.SetEmissiveFactor(float3(p[8], p[9], p[10]));
        m.SetEmissiveStrength(p[11]);
        m.SetCoatWeight(p[12]);
        m.SetCoatColor(float3(p[13], p[14], p[15]));
        m.SetCoatRoughness(p[16]);
        m.SetCoatIOR(p[17]);
        m.SetSubsurface(p[18]);
        m.SetTransmissionDepth(p[19]);
        m.SetDoubleSided((flags & 1) != 0);
        m.SetThinWalled((flags & 2) != 0);
        memcpy(out32, &m, sizeof(m));
    }

    // v: 9 floats (three vertices), uv: 6 floats
    void ref_emissive_triangle(const float* v, const float* uv, uint32_t factorRGB8, uint32_t tex, uint16_t strengthHalfBits,
        uint32_t triIdx, int doubleSided, void* out48)
    {
        RT::EmissiveTriangle t(float3(v[0], v[1], v[2]), float3(v[3], v[4], v[5]), float3(v[6], v[7], v[8]),
            float2(uv[0], uv[1]), float2(uv[2], uv[3]), float2(uv[4], uv[5]), factorRGB8, tex, half::asfloat16(strengthHalfBits),
            triIdx, doubleSided != 0);
        memcpy(out48, &t, sizeof(t));
    }

    uint32_t ref_oct32(float x, float y, float z)
    {
        oct32 o(x, y, z);
        uint32_t r;
        memcpy(&r, &o.v, 4);
        return r;
    }
    void ref_oct32_decode(uint32_t enc, float* out3)
    {
        oct32 o;
        memcpy(&o.v, &enc, 4);
        float3 d = o.decode();
        out3[0] = d.x; out3[1] = d.y; out3[2] = d.z;
    }
    uint16_t ref_half(float f) { return half(f).x; }
    uint32_t ref_rgb8(float r, float g, float b) { return Float3ToRGB8(float3(r, g, b)); }
    // quaternion (normalised floats in [-1, 1]) -> the unorm4 RT::MeshInstance::Rotation stores; scale -> half3
    void ref_instance_rotation_scale(const float* q4, const float* s3, uint16_t* outRot4, uint16_t* outScale3)
    {
        float4a q(q4[0], q4[1], q4[2], q4[3]);
        unorm4 u = unorm4::FromNormalized(q);       // RtAccelerationStructure.cpp:345 (maps [-1, 1] -> [0, 1] itself)
        memcpy(outRot4, &u, 8);
        half3 h(float3(s3[0], s3[1], s3[2]));
        memcpy(outScale3, &h, 6);
    }
}
