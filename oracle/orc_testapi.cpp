// ORACLE -- test infrastructure, not product code (see orc_math.h header).
// Small C entry points so the CPU-only tests can probe codecs and the BSDF restatement directly.
#include "orc_scene.h"

using namespace orc;

extern "C"
{
    // n unit vectors -> oct32 -> unit vectors (Math::EncodeOct32 / DecodeOct32, Math.hlsli:636-676)
    void orc_oct32_roundtrip(const float* in, int n, float* out)
    {
        for (int i = 0; i < n; i++)
        {
            float3 d = Math::DecodeOct32(Math::EncodeOct32u(f3(in[3 * i], in[3 * i + 1], in[3 * i + 2])));
            out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z;
        }
    }
    uint32_t orc_pack_r11g11b10(float r, float g, float b) { return pack_r11g11b10(f3(r, g, b)); }
    void orc_unpack_r11g11b10(uint32_t p, float* out) { float3 c = unpack_r11g11b10(p); out[0] = c.x; out[1] = c.y; out[2] = c.z; }
    uint32_t orc_pack_snorm16x2(float x, float y) { return pack_snorm16x2(f2(x, y)); }
    void orc_unpack_snorm16x2(uint32_t p, float* out) { float2 c = unpack_snorm16x2(p); out[0] = c.x; out[1] = c.y; }

    struct orc_surface_desc
    {
        float normal[3], wo[3]; uint32_t metallic; float roughness; float baseColor[3]; float eta_curr, eta_next; uint32_t specTr;
        float trDepth, subsurface, coat_weight; float coat_color[3]; float coat_roughness, coat_ior;
    };
    static BSDF::ShadingData mk(const orc_surface_desc* d)
    {
        return BSDF::ShadingData::Init(f3(d->normal[0], d->normal[1], d->normal[2]), f3(d->wo[0], d->wo[1], d->wo[2]), d->metallic != 0,
            d->roughness, f3(d->baseColor[0], d->baseColor[1], d->baseColor[2]), d->eta_curr, d->eta_next, d->specTr != 0, d->trDepth,
            d->subsurface, d->coat_weight, f3(d->coat_color[0], d->coat_color[1], d->coat_color[2]), d->coat_roughness, d->coat_ior);
    }
    // out: wi(3), lobe, pdf, bsdfOverPdf(3), f(3), rngStateAfter
    void orc_bsdf_sample(const orc_surface_desc* d, uint32_t seed, float* out)
    {
        BSDF::ShadingData s = mk(d);
        RNG rng = RNG::InitSeed(seed);
        BSDF::BSDFSample b = BSDF::SampleBSDF(f3(d->normal[0], d->normal[1], d->normal[2]), s, rng);
        out[0] = b.wi.x; out[1] = b.wi.y; out[2] = b.wi.z; out[3] = (float)b.lobe; out[4] = b.pdf;
        out[5] = b.bsdfOverPdf.x; out[6] = b.bsdfOverPdf.y; out[7] = b.bsdfOverPdf.z; out[8] = b.f.x; out[9] = b.f.y; out[10] = b.f.z;
        out[11] = asfloat(rng.State);
    }
    // out: pdf, bsdfOverPdf(3), f(3), rngStateAfter
    void orc_bsdf_eval_sampler(const orc_surface_desc* d, const float* wi, uint32_t lobe, uint32_t seed, float* out)
    {
        BSDF::ShadingData s = mk(d);
        RNG rng = RNG::InitSeed(seed);
        BSDF::BSDFSamplerEval e = BSDF::EvalBSDFSampler(f3(d->normal[0], d->normal[1], d->normal[2]), s, f3(wi[0], wi[1], wi[2]), (BSDF::LOBE)lobe, rng);
        out[0] = e.pdf; out[1] = e.bsdfOverPdf.x; out[2] = e.bsdfOverPdf.y; out[3] = e.bsdfOverPdf.z; out[4] = e.f.x; out[5] = e.f.y; out[6] = e.f.z;
        out[7] = asfloat(rng.State);
    }
    float orc_bsdf_sampler_pdf(const orc_surface_desc* d, const float* wi, uint32_t seed)
    {
        BSDF::ShadingData s = mk(d);
        RNG rng = RNG::InitSeed(seed);
        return BSDF::BSDFSamplerPdf(f3(d->normal[0], d->normal[1], d->normal[2]), s, f3(wi[0], wi[1], wi[2]), rng);
    }
    void orc_bsdf_sample_nodiffuse(const orc_surface_desc* d, uint32_t seed, float* out)
    {
        BSDF::ShadingData s = mk(d);
        RNG rng = RNG::InitSeed(seed);
        BSDF::BSDFSample b = BSDF::SampleBSDF_NoDiffuse(f3(d->normal[0], d->normal[1], d->normal[2]), s, rng);
        out[0] = b.wi.x; out[1] = b.wi.y; out[2] = b.wi.z; out[3] = (float)b.lobe; out[4] = b.pdf;
        out[5] = b.bsdfOverPdf.x; out[6] = b.bsdfOverPdf.y; out[7] = b.bsdfOverPdf.z; out[8] = b.f.x; out[9] = b.f.y; out[10] = b.f.z;
        out[11] = asfloat(rng.State);
    }
    float orc_bsdf_sampler_pdf_nodiffuse(const orc_surface_desc* d, const float* wi)
    {
        BSDF::ShadingData s = mk(d);
        return BSDF::BSDFSamplerPdf_NoDiffuse(f3(d->normal[0], d->normal[1], d->normal[2]), s, f3(wi[0], wi[1], wi[2]));
    }
    // f(wi) * |cos| as BSDF::Unified returns it
    void orc_bsdf_unified(const orc_surface_desc* d, const float* wi, float* out)
    {
        BSDF::ShadingData s = mk(d);
        s.SetWi(f3(wi[0], wi[1], wi[2]), f3(d->normal[0], d->normal[1], d->normal[2]));
        float3 f = BSDF::Unified(s).f;
        out[0] = f.x; out[1] = f.y; out[2] = f.z;
    }
}
