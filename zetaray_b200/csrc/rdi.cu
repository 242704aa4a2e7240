// rdi.cu -- ReSTIR DI for emissive triangles and the DirectLighting pass.
//
// Replaces DirectLighting/Emissive/ReSTIR_DI_Temporal.hlsl:29-390, ReSTIR_DI_Spatial.hlsl:27-192,
// Resampling.hlsli:36-519, PairwiseMIS.hlsli:10-232, Reservoir.hlsli:10-213, Util.hlsli:9-120 and the host
// sequencing of DirectLighting.cpp:166-284 (compiled configuration: USE_HALF_VECTOR_COPY_SHIFT 0, alias-table
// candidates, no presampled sets).
//
// Layout: the reservoir's two textures (RGBA32UI + RG32F, 24 B/px) are one 32-byte record = 2 x 128-bit;
// the RGBA16F target plane is a uint2 (8 B/px). k_di_spatial keeps the reference's 8x8 group / swizzle so one
// warp is one reference wave (the disocclusion vote is a wave op) and the group RNG is seeded by the block id.
#include "zr_pixel.cuh"
#include "zr_schedule.h"

#include "zr_rdi.cuh"

namespace zr
{
namespace
{
#ifndef ZR_RDI_THREADS
#define ZR_RDI_THREADS 1024
#endif
    // ReSTIR_DI_Temporal.hlsl main + EstimateDirectLighting. A block is ZR_RDI_THREADS/64 consecutive 8x8 groups of the
    // reference's swizzled dispatch, walking the resampling phases together (no thread leaves before the last barrier).
    __global__ void ZR_LB(ZR_RDI_THREADS) k_di_temporal(SceneDev sc, FrameView f, DIParams prm, zr_rdi_reservoir* __restrict__ resCurr,
        const zr_rdi_reservoir* __restrict__ resPrev, uint2* __restrict__ target, float4* __restrict__ finalImg, uint32_t dispX, uint32_t dispY,
        const uint32_t* __restrict__ order)
    {
        const zr_frame_constants& fc = f.fc;
        uint2 sg = make_uint2(0, 0);
        const uint32_t groupFlat = order[blockIdx.x] * (ZR_RDI_THREADS / 64) + (threadIdx.x >> 6);
        const uint32_t tInGroup = threadIdx.x & 63;
        uint2 px = make_uint2(0xffffffffu, 0xffffffffu);
        if (groupFlat < dispX * dispY)
            px = SwizzleThreadGroup(groupFlat, 0, tInGroup & 7, tInGroup >> 3, 8, 8, dispX, 16, 4, 16 * dispY, sg);
        const long long t0 = clock64();
        bool act = !(px.x >= f.W || px.y >= f.H || px.y < prm.rowBegin || px.y >= prm.rowEnd);
        const uint32_t x = px.x, y = px.y;
        const size_t idx = act ? (size_t)y * f.W + x : 0;
        if (act)
        {
            const GFlags flags = FlagsAt(f.core, f.W, x, y);
            if (flags.invalid)
            {
                // the sky / sun-disk background belongs to the sun-sky path (not in this build)
                finalImg[idx] = f4(0, 0, 0, 0);
                act = false;
            }
            else if (flags.emissive && !prm.spatial)
            {
                WriteEmissive(fc, f, finalImg, idx);
                act = false;
            }
        }
        Pixel p;
        p.surface = BSDF::ShadingData::InitEmpty();
        p.pos = f3(0); p.normal = f3(0); p.roughness = 0;
        RNG rng_thread; rng_thread.State = 0;
        int numBsdfSamples = 0;
        if (act)
        {
            p = LoadPixel(f, sc, f.core, f.coat, x, y, false, x, y);
            rng_thread = RNG::Init(x, y, fc.FrameNum);
            numBsdfSamples = (!p.surface.GlossSpecular() && p.roughness < 0.3f) ? 2 : 1;
        }
        // group-uniform index so that every thread of an 8x8 group uses the same presampled set (ReSTIR_DI_Temporal.hlsl:368-370)
        RNG rng_group = RNG::Init(groupFlat % dispX, groupFlat / dispX, fc.FrameNum);
        const uint32_t sampleSetIdx = rng_group.UniformUintBounded_Faster(sc.numSampleSets);
        Reservoir r = RIS_InitialCandidates_Sync(act, sc, p.pos, p.normal, p.roughness, p.surface, sampleSetIdx, numBsdfSamples, rng_thread);
        if (prm.temporal)
        {
            float2 motionVec = f2(0, 0);
            TemporalCandidate tc; tc.valid = false; tc.px = tc.py = 0; tc.pos = tc.normal = f3(0);
            tc.surface = BSDF::ShadingData::InitEmpty();
            ZR_PHASE();
            if (act)
            {
                motionVec = unpack_snorm16x2(__ldg(&f.me[idx].x));
                const float2 currUV = f2((float)x + 0.5f, (float)y + 0.5f) / f2((float)f.W, (float)f.H);
                const float2 prevUV = currUV - motionVec;
                tc = FindTemporalCandidate(f, sc, p.pos, p.normal, p.roughness, p.surface, prevUV);
            }
            TemporalResample1_Sync(act && tc.valid, sc, p.pos, p.normal, p.surface, tc, resPrev, f.W, r, rng_thread);
            if (act && prm.spatial)
            {
                const bool disoccluded = !tc.valid && (dot(motionVec, motionVec) > 0);
                r.target = disoccluded ? -r.target : r.target;
                WriteTarget(target, idx, r.target);
                r.target = Math::Sanitize(r.target);
            }
        }
        AccountCost(prm.costMap, f.W, f.H, px.x, px.y, t0);
        if (!act)
            return;
        if (prm.temporal || prm.reset)
        {
            zr_rdi_reservoir rec;
            r.Write(rec, prm.M_max);
            StoreRdi(&resCurr[idx], rec);
        }
        if (!prm.spatial || !prm.temporal)
            WriteFinal(fc, finalImg, idx, r.target * r.W);
    }

    // ReSTIR_DI_Spatial.hlsl main + SpatialResample
    __global__ void ZR_LB(ZR_RDI_THREADS) k_di_spatial(SceneDev sc, FrameView f, DIParams prm, const zr_rdi_reservoir* __restrict__ resCurr,
        const uint2* __restrict__ target, float4* __restrict__ finalImg, uint32_t dispX, uint32_t dispY,
        const uint32_t* __restrict__ order)
    {
        const zr_frame_constants& fc = f.fc;
        uint2 sg = make_uint2(0, 0);
        const uint32_t groupFlat = order[blockIdx.x] * (ZR_RDI_THREADS / 64) + (threadIdx.x >> 6);
        const uint32_t tInGroup = threadIdx.x & 63;
        uint2 px = make_uint2(0xffffffffu, 0xffffffffu);
        if (groupFlat < dispX * dispY)
            px = SwizzleThreadGroup(groupFlat, 0, tInGroup & 7, tInGroup >> 3, 8, 8, dispX, 16, 4, 16 * dispY, sg);
        const long long t0 = clock64();
        bool active = px.x < f.W && px.y < f.H && px.y >= prm.rowBegin && px.y < prm.rowEnd;
        const int x = (int)px.x, y = (int)px.y;
        const size_t idx = active ? (size_t)y * f.W + x : 0;
        if (active)
        {
            const GFlags flags = FlagsAt(f.core, f.W, x, y);
            if (flags.invalid) active = false;
            else if (flags.emissive)
            {
                WriteEmissive(fc, f, finalImg, idx);
                active = false;
            }
        }
        Pixel p;
        p.surface = BSDF::ShadingData::InitEmpty();
        p.pos = f3(0); p.normal = f3(0); p.roughness = 0; p.z = 0;
        Reservoir r = Reservoir::Init();
        bool disoccluded = false;
        if (active)
        {
            p = LoadPixel(f, sc, f.core, f.coat, x, y, false, x, y);
            zr_rdi_reservoir rec;
            LoadRdi(&resCurr[idx], rec);
            r = Reservoir::Load(rec);
            if (r.lightIdx != UINT32_MAX_)
            {
                const zr_emissive_tri& tri = sc.emissives[r.lightIdx];
                r.lightID = tri.ID;
                const float3 vtx0 = Light::Vtx0(tri);
                const float3 vtx1 = Light::DecodeEmissiveTriV1(tri);
                const float3 vtx2 = Light::DecodeEmissiveTriV2(tri);
                r.lightPos = (1.0f - r.bary.x - r.bary.y) * vtx0 + r.bary.x * vtx1 + r.bary.y * vtx2;
                r.lightNormal = cross(vtx1 - vtx0, vtx2 - vtx0);
                r.lightNormal = dot(r.lightNormal, r.lightNormal) == 0 ? r.lightNormal : normalize(r.lightNormal);
                r.doubleSided = Light::IsDoubleSided(tri);
                r.target = LoadTarget(target, idx);
                disoccluded = r.target.x < 0 || r.target.y < 0 || r.target.z < 0;
                r.target = abs3(r.target);
            }
        }
        const uint32_t waveDisoccluded = __popc(__ballot_sync(0xffffffffu, active && disoccluded));
        RNG rng_group = RNG::Init(groupFlat % dispX, groupFlat / dispX, fc.FrameNum);
        rng_group.UniformUintBounded_Faster(sc.numSampleSets);    // sample-set index: drawn, not used by the spatial pass (:137)
        const bool extra = !prm.stochasticSpatial || (rng_group.Uniform() < 0.6f);
        if (prm.extraDisocclusion)
            disoccluded = disoccluded && (waveDisoccluded > 3);
        int numSamples = extra ? 2 : 1;
        numSamples = !disoccluded ? numSamples : 4;
        RNG rng = RNG::Init((uint32_t)x, (uint32_t)y, fc.FrameNum);
        const float u0 = rng.Uniform();
        const int offset = (int)rng.UniformUintBounded_Faster(8);
        const float theta = u0 * TWO_PI;
        float sinTheta, cosTheta;
        zr_sincosf(theta, &sinTheta, &cosTheta);
        PairwiseMIS pairwiseMIS = PairwiseMIS::Init((uint32_t)numSamples, r);
        float3 samplePos[4]; int spx[4], spy[4]; uint32_t k = 0;
        ZR_PHASE();
        for (int i = 0; i < 4; i++)
        {
            if (!(active && i < numSamples)) continue;
            const float2 sampleUV = f2(c_disk32[((offset + i) & 31) * 2], c_disk32[((offset + i) & 31) * 2 + 1]);
            float2 rotated;
            rotated.x = dot(sampleUV, f2(cosTheta, -sinTheta));
            rotated.y = dot(sampleUV, f2(sinTheta, cosTheta));
            rotated = rotated * 16.0f;
            const float fx = fmaxf(rintf((float)x + rotated.x), 0.0f), fy = fmaxf(rintf((float)y + rotated.y), 0.0f);
            if (fx >= (float)f.W || fy >= (float)f.H) continue;
            const int qx = (int)fx, qy = (int)fy;
            float rough_i;
            const GFlags flags_i = FlagsAt(f.core, f.W, qx, qy, &rough_i);
            if (flags_i.invalid || flags_i.emissive) continue;
            const Pixel pi = LoadPixel(f, sc, f.core, f.coat, qx, qy, false, qx, qy);
            bool valid = PlaneHeuristicDI(pi.pos, p.normal, p.pos, p.z);
            valid = valid && (fabsf(rough_i - p.roughness) < 0.15f);
            if (!valid) continue;
            samplePos[k] = pi.pos; spx[k] = qx; spy[k] = qy;
            k++;
        }
        pairwiseMIS.k = k;
        for (uint32_t i = 0; i < 4; i++)
        {
            const bool go = active && (i < k);
            if (!__syncthreads_or(go))
                break;
            Pixel pi;
            pi.normal = f3(0);
            BSDF::ShadingData surface_i = BSDF::ShadingData::InitEmpty();
            Reservoir r_spatial = Reservoir::Init();
            float3 pos_i = f3(0);
            if (go)
            {
                pi = LoadPixel(f, sc, f.core, f.coat, spx[i], spy[i], false, spx[i], spy[i]);
                // the neighbour surface is rebuilt with transmission depth = 0 (Resampling.hlsli:507-510)
                const uint4 c = ld128(&f.core[(size_t)spy[i] * f.W + spx[i]]);
                const float3 bc = f3((float)(c.z & 0xff) / 255.0f, (float)((c.z >> 8) & 0xff) / 255.0f, (float)((c.z >> 16) & 0xff) / 255.0f);
                const float bw = pi.flags.subsurface ? (float)(c.z >> 24) / 255.0f : 0.0f;
                pos_i = samplePos[i];
                const float3 wo_i = normalize(pi.origin - pos_i);
                surface_i = BSDF::ShadingData::Init(pi.normal, wo_i, pi.flags.metallic, pi.roughness, bc, BSDF::ETA_AIR,
                    pi.eta_next, pi.flags.transmissive, 0.0f, to_half(bw), pi.surface.coat_weight, pi.surface.coat_color,
                    pi.coatRoughness, pi.coatIor, sc.rho);
                zr_rdi_reservoir recN;
                LoadRdi(&resCurr[(size_t)spy[i] * f.W + spx[i]], recN);
                r_spatial = Reservoir::Load(recN);
            }
            pairwiseMIS.Stream_Sync(go, sc, r, p.pos, p.normal, p.surface, r_spatial, pos_i, pi.normal, surface_i, rng);
        }
        AccountCost(prm.costMap, f.W, f.H, px.x, px.y, t0);
        if (!active)
            return;
        pairwiseMIS.End(r, rng);
        const Reservoir rs = pairwiseMIS.r_s;
        WriteFinal(fc, finalImg, idx, rs.target * rs.W);
    }
}
} // namespace zr

// ------------------------------------------------------------------------------------------------
// DirectLighting pass object (DirectLighting/Emissive/DirectLighting.h:36-57)
// ------------------------------------------------------------------------------------------------
#include <cstdio>
#include <string>
#include <vector>
#include <dlfcn.h>

struct zr_direct_pass
{
    uint32_t width = 0, height = 0;
    zr_rdi_reservoir* d_res[2] = { nullptr, nullptr };
    uint2* d_target = nullptr;      // RGBA16F
    float4* d_final = nullptr;
    int currTemporalIdx = 0;
    bool isTemporalReservoirValid = false;
    bool resetTemporalTextures = true;
    bool patternLoaded = false;
    zr_direct_params params{};
    // strip-sharded frames: owned rows, halo-exchange hook, optional cost map
    uint32_t rowBegin = 0, rowEnd = 0xffffffffu;
    zr_halo_exchange_fn exchange = nullptr;
    void* exchangeUser = nullptr;
    unsigned long long* d_costMap = nullptr;
    zr::TileCosts tileCosts;
    zr::BlockSchedule sched;        // both kernels share the 8x8-group geometry

    static void Defaults(zr_direct_params* p)
    {
        // DirectLighting.cpp:99-107, DirectLighting.h:93-98
        p->temporal_resample = 1; p->spatial_resample = 1; p->stochastic_spatial = 1; p->extra_disocclusion_sampling = 1;
        p->M_max = 20; p->alpha_min = 0.05f * 0.05f;
    }
    void Release()
    {
        for (int i = 0; i < 2; i++) { if (d_res[i]) cudaFree(d_res[i]); d_res[i] = nullptr; }
        sched.Release();
        if (d_target) cudaFree(d_target); if (d_final) cudaFree(d_final);
        d_target = nullptr; d_final = nullptr;
    }
    zr_status OnWindowResized(uint32_t w, uint32_t h)
    {
        Release();
        width = w; height = h;
        const size_t n = (size_t)w * h;
        for (int i = 0; i < 2; i++) ZR_CUDA(cudaMalloc(&d_res[i], n * sizeof(zr_rdi_reservoir)));
        ZR_CUDA(cudaMalloc(&d_target, n * 8));
        ZR_CUDA(cudaMalloc(&d_final, n * 16));
        return ResetTemporal();
    }
    zr_status ResetTemporal()
    {
        const size_t n = (size_t)width * height;
        ZR_CLEAR_BEGIN();
        for (int i = 0; i < 2; i++) ZR_CUDA(cudaMemset(d_res[i], 0, n * sizeof(zr_rdi_reservoir)));
        ZR_CUDA(cudaMemset(d_target, 0, n * 8));
        ZR_CUDA(cudaMemset(d_final, 0, n * 16));
        ZR_CLEAR_END();
        currTemporalIdx = 0; isTemporalReservoirValid = false; resetTemporalTextures = true;
        return ZR_OK;
    }
    zr_status LoadPattern()
    {
        if (patternLoaded) return ZR_OK;
        Dl_info info;
        std::string dir = ".";
        if (dladdr((void*)&zr_direct_pass::Defaults, &info) && info.dli_fname)
        {
            std::string p = info.dli_fname;
            size_t s = p.find_last_of('/');
            if (s != std::string::npos) dir = p.substr(0, s);
        }
        const std::string path = dir + "/assets/disk32.bin";
        float pat[64];
        FILE* fp = fopen(path.c_str(), "rb");
        if (!fp || fread(pat, 4, 64, fp) != 64)
        {
            if (fp) fclose(fp);
            zr::set_error("zr_direct_pass: cannot read %s (tools/extract_reference_tables.py writes it)", path.c_str());
            return ZR_ERR_NOT_INITIALIZED;
        }
        fclose(fp);
        ZR_CUDA(cudaMemcpyToSymbol(zr::c_disk32, pat, 256));
        patternLoaded = true;
        return ZR_OK;
    }
    zr_status Render(const zr_frame_inputs* in, cudaStream_t stream)
    {
        using namespace zr;
        if (!in || !in->scene || !in->curr.d_core || !in->curr.d_motion_emissive || !in->curr.d_coat)
        {
            set_error("zr_direct_pass_render: missing scene or G-buffer");
            return ZR_ERR_INVALID_ARG;
        }
        if (in->frame.RenderWidth != width || in->frame.RenderHeight != height)
        {
            set_error("zr_direct_pass_render: frame/pass size mismatch");
            return ZR_ERR_INVALID_ARG;
        }
        if (in->scene->dev.numEmissives == 0 || !in->scene->aliasBuilt)
        {
            // PathTracer.cpp:274-284: ReSTIR DI (emissive) only runs when the scene has emissive triangles
            set_error("zr_direct_pass_render: needs emissive triangles and zr_prelighting_render first (SkyDI is not part of this build)");
            return ZR_ERR_UNSUPPORTED;
        }
        if (in->scene->dev.sampleSetSize && !in->scene->samplesValid)
        {
            set_error("zr_direct_pass_render: presampling is enabled but zr_presample_emissives has not run");
            return ZR_ERR_NOT_INITIALIZED;
        }
        zr_status st = LoadPattern();
        if (st != ZR_OK) return st;
        const bool doTemporal = isTemporalReservoirValid && params.temporal_resample;
        const bool doSpatial = doTemporal && params.spatial_resample;
        if (doTemporal && (!in->prev.d_core || !in->prev.d_coat))
        {
            set_error("zr_direct_pass_render: temporal reuse needs the previous G-buffer");
            return ZR_ERR_INVALID_ARG;
        }
        FrameView f;
        f.fc = in->frame;
        f.core = (const uint4*)in->curr.d_core; f.depth = (const float*)in->curr.d_depth;
        f.me = (const uint2*)in->curr.d_motion_emissive; f.coat = (const uint2*)in->curr.d_coat;
        f.pcore = (const uint4*)in->prev.d_core; f.pcoat = (const uint2*)in->prev.d_coat;
        f.W = width; f.H = height;
        DIParams prm{ doTemporal, doSpatial, params.stochastic_spatial, params.extra_disocclusion_sampling, params.M_max,
            params.alpha_min, resetTemporalTextures, rowBegin, rowEnd < height ? rowEnd : height, d_costMap };
        const uint32_t dispX = (width + 7) / 8, dispY = (height + 7) / 8;
        const int cur = currTemporalIdx;
        if (!sched.UpToDate(prm.rowBegin, prm.rowEnd, tileCosts.version))
            ZR_CUDA(sched.Upload(zr::ScheduleSwizzled(dispX, dispY, 8, 8, ZR_RDI_THREADS / 64, prm.rowBegin, prm.rowEnd, tileCosts),
                prm.rowBegin, prm.rowEnd, tileCosts.version));
        ZR_PROF("k_di_temporal", stream);
        k_di_temporal<<<sched.count, ZR_RDI_THREADS, 0, stream>>>(in->scene->dev, f, prm, d_res[cur], d_res[1 - cur], d_target, d_final, dispX, dispY, sched.d_order);
        ZR_LAUNCH_CHECK();
        // the temporal output is what neighbours read in the spatial pass and what the next frame reprojects into
        if (exchange)
        {
            const zr_image2d plane{ d_res[cur], width, height, width * 32u, 32u };
            exchange(exchangeUser, &plane, 1, stream);
        }
        if (doSpatial)
        {
            ZR_PROF("k_di_spatial", stream);
            k_di_spatial<<<sched.count, ZR_RDI_THREADS, 0, stream>>>(in->scene->dev, f, prm, d_res[cur], d_target, d_final, dispX, dispY, sched.d_order);
            ZR_LAUNCH_CHECK();
        }
        isTemporalReservoirValid = true;
        currTemporalIdx = 1 - cur;
        resetTemporalTextures = false;
        return ZR_OK;
    }
};

extern "C"
{
    zr_status zr_direct_pass_create(uint32_t width, uint32_t height, zr_direct_pass** out)
    {
        if (!out || !width || !height) { zr::set_error("zr_direct_pass_create: bad args"); return ZR_ERR_INVALID_ARG; }
        zr_direct_pass* p = new zr_direct_pass();
        zr_direct_pass::Defaults(&p->params);
        zr_status s = p->OnWindowResized(width, height);
        if (s != ZR_OK) { p->Release(); delete p; return s; }
        *out = p;
        return ZR_OK;
    }
    zr_status zr_direct_pass_resize(zr_direct_pass* p, uint32_t width, uint32_t height)
    {
        if (!p || !width || !height) return ZR_ERR_INVALID_ARG;
        return p->OnWindowResized(width, height);
    }
    zr_status zr_direct_pass_reset_temporal(zr_direct_pass* p) { return p ? p->ResetTemporal() : ZR_ERR_INVALID_ARG; }
    zr_status zr_direct_pass_default_params(zr_direct_params* out)
    {
        if (!out) return ZR_ERR_INVALID_ARG;
        zr_direct_pass::Defaults(out);
        return ZR_OK;
    }
    zr_status zr_direct_pass_set_params(zr_direct_pass* p, const zr_direct_params* params)
    {
        if (!p || !params) return ZR_ERR_INVALID_ARG;
        if (params->M_max == 0 || params->M_max > 31) { zr::set_error("zr_direct_pass_set_params: M_max must be in 1..31 (5-bit field)"); return ZR_ERR_INVALID_ARG; }
        p->params = *params;
        return ZR_OK;
    }
    zr_status zr_direct_pass_render(zr_direct_pass* p, const zr_frame_inputs* in, void* stream)
    {
        if (!p) return ZR_ERR_INVALID_ARG;
        return p->Render(in, (cudaStream_t)stream);
    }
    zr_status zr_direct_pass_set_rows(zr_direct_pass* p, uint32_t y0, uint32_t y1)
    {
        if (!p || y0 >= y1 || y0 >= p->height) { zr::set_error("zr_direct_pass_set_rows: empty row range"); return ZR_ERR_INVALID_ARG; }
        p->rowBegin = y0; p->rowEnd = y1;
        return ZR_OK;
    }
    zr_status zr_direct_pass_set_halo_exchange(zr_direct_pass* p, zr_halo_exchange_fn fn, void* user)
    {
        if (!p) return ZR_ERR_INVALID_ARG;
        p->exchange = fn; p->exchangeUser = user;
        return ZR_OK;
    }
    zr_status zr_direct_pass_set_schedule_costs(zr_direct_pass* p, const double* h_tile_cost, uint32_t tiles_x, uint32_t tiles_y)
    {
        if (!p) return ZR_ERR_INVALID_ARG;
        if (h_tile_cost && (tiles_x != (p->width + 31) / 32 || tiles_y != (p->height + 31) / 32))
        {
            zr::set_error("zr_direct_pass_set_schedule_costs: expected %u x %u tiles", (p->width + 31) / 32, (p->height + 31) / 32);
            return ZR_ERR_INVALID_ARG;
        }
        p->tileCosts.cost.assign(h_tile_cost ? h_tile_cost : nullptr, h_tile_cost ? h_tile_cost + (size_t)tiles_x * tiles_y : nullptr);
        p->tileCosts.tilesX = h_tile_cost ? tiles_x : 0;
        p->tileCosts.version++;
        return ZR_OK;
    }
    zr_status zr_direct_pass_set_cost_map(zr_direct_pass* p, void* d_cycles)
    {
        if (!p) return ZR_ERR_INVALID_ARG;
        p->d_costMap = (unsigned long long*)d_cycles;
        return ZR_OK;
    }
    zr_status zr_direct_pass_get_output(zr_direct_pass* p, zr_direct_output id, zr_image2d* out)
    {
        if (!p || !out) return ZR_ERR_INVALID_ARG;
        const uint32_t w = p->width, h = p->height;
        switch (id)
        {
        case ZR_DIRECT_FINAL: *out = zr_image2d{ p->d_final, w, h, w * 16u, 16u }; break;
        case ZR_DIRECT_RESERVOIR_CURR: *out = zr_image2d{ p->d_res[1 - p->currTemporalIdx], w, h, w * 32u, 32u }; break;
        case ZR_DIRECT_TARGET: *out = zr_image2d{ p->d_target, w, h, w * 8u, 8u }; break;
        default: zr::set_error("zr_direct_pass_get_output: unknown output id"); return ZR_ERR_INVALID_ARG;
        }
        return ZR_OK;
    }
    zr_status zr_direct_pass_describe_io(zr_direct_pass* p, zr_resource_use* uses, int* n)
    {
        if (!p || !uses || !n) return ZR_ERR_INVALID_ARG;
        uses[0] = zr_resource_use{ ZR_RES_GBUFFER_CURR, 0 };
        uses[1] = zr_resource_use{ ZR_RES_GBUFFER_PREV, 0 };
        uses[2] = zr_resource_use{ ZR_RES_SCENE_BVH, 0 };
        uses[3] = zr_resource_use{ ZR_RES_ALIAS_TABLE, 0 };
        uses[4] = zr_resource_use{ ZR_RES_DI_FINAL, 1 };
        *n = 5;
        return ZR_OK;
    }
    void zr_direct_pass_destroy(zr_direct_pass* p) { if (p) { p->Release(); delete p; } }
}
