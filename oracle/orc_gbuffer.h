// ORACLE -- test infrastructure, not product code (see orc_math.h header).
// G-buffer record codecs: the reference's texture formats (DefaultRendererImpl.h:97-109) and
// GBuffers.hlsli:52-121 in the repo's packed layout (include/zr_abi.h, zr_gbuffer).
#pragma once
#include "orc_math.h"
#include "../include/zr_abi.h"

namespace orc
{
struct GFlags
{
    bool metallic, transmissive, emissive, invalid, trDepthGt0, subsurface, coated;
};

inline GFlags DecodeFlags(uint32_t v)
{
    GFlags r;
    r.transmissive = (v & 0x1) != 0;
    r.emissive = (v & (1 << 1)) != 0;
    r.invalid = (v & (1 << 2)) != 0;
    r.trDepthGt0 = (v & (1 << 3)) != 0;
    r.subsurface = (v & (1 << 4)) != 0;
    r.coated = (v & (1 << 5)) != 0;
    r.metallic = (v & (1 << 7)) != 0;
    return r;
}

struct GCore
{
    float depth;
    float2 normalEnc;    // as read from R16G16_UNORM
    float4 baseColor;    // as read from R8G8B8A8_UNORM
    uint32_t flagsByte;
    float roughness;     // as read from R8_UNORM
    float iorEnc;        // as read from R8_UNORM
};

inline GCore LoadCore(const uint4* core, size_t idx)
{
    uint4 c = core[idx];
    GCore g;
    g.depth = asfloat(c.x);
    g.normalEnc = Math::DecodeUNorm2(c.y);
    g.baseColor = f4((float)(c.z & 0xff) / 255.0f, (float)((c.z >> 8) & 0xff) / 255.0f,
        (float)((c.z >> 16) & 0xff) / 255.0f, (float)(c.z >> 24) / 255.0f);
    g.flagsByte = c.w & 0xff;
    g.roughness = (float)((c.w >> 8) & 0xff) / 255.0f;
    g.iorEnc = (float)((c.w >> 16) & 0xff) / 255.0f;
    return g;
}

inline uint32_t unorm8(float f) { return (uint32_t)mad(saturate(f), 255.0f, 0.5f); }

// GBuffer::EncodeIOR / DecodeIOR (GBuffers.hlsli:98-106), MIN_IOR 1, MAX_IOR 2.5
inline float EncodeIOR(float ior) { return (ior - 1.0f) / (2.5f - 1.0f); }
inline float DecodeIOR(float e) { return mad(e, 2.5f - 1.0f, 1.0f); }

struct Coat { float weight; float3 color; float roughness; float ior; };
inline Coat UnpackCoat(uint3 packed)
{
    Coat ret;
    ret.weight = Math::UNorm8ToFloat((packed.y >> 8) & 0xff);
    ret.roughness = Math::UNorm8ToFloat(packed.z & 0xff);
    uint32_t c = packed.x | ((packed.y & 0xff) << 16);
    ret.color = Math::UnpackRGB8(c);
    float normalized = Math::UNorm8ToFloat(packed.z >> 8);
    ret.ior = DecodeIOR(normalized);
    return ret;
}
inline uint3 LoadCoat(const uint2* coat, size_t idx)
{
    uint2 c = coat[idx];
    return uint3{ c.x & 0xffff, c.x >> 16, c.y & 0xffff };
}
} // namespace orc
