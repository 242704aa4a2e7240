// svgf.cu -- SVGF denoiser: temporal accumulation + variance estimate, then a-trous wavelet passes as TMA-tiled stencils.
//
// No counterpart in the reference (ZetaRay ships no SVGF; BASELINE.json's north_star and config 3 name it): the algorithm is
// defined by oracle/orc_svgf.cpp and this file is held to it bit for bit (tests/test_svgf_gpu.py).
//
// Execution model of an a-trous pass with step s. A tap pattern with stride s never leaves its sub-lattice
// {(x, y) : x = u s + px, y = v s + py}, so the image is VIEWED as a 4-D tensor {px, u, py, v} (strides 8 B, 8 s B, pitch, s pitch) and a
// block filters a dense 32 x 16 box of lattice points for two adjacent x-phases: one cp.async.bulk.tensor.4d brings the (32 + 2R) x
// (16 + 2R) x 2 lattice box of colour + variance, a second one the guide box {depth, normal}, both land as dense tiles in shared memory
// whatever the step (SASS: UTMALDG.4D), are expanded once to float (normal decode and luminance once per staged pixel, not once per
// tap), filtered from shared memory, and the result tile leaves through one cp.async.bulk.tensor.4d store (UTMASTG.4D). Every pass
// therefore reads 1.2-1.4x and writes 1.0x its algorithmic bytes from L2 / HBM at any step, with no strided global access.
// Step 1 is the same kernel over a plain 2-D map (64 x 16 pixel tiles).
#include "zr_common.cuh"
#include "zr_tma.cuh"
#include <cstdlib>
#include <cstring>

namespace zr
{
namespace
{
    struct SvgfParamsDev { float sigma_z, k_n, sigma_l; };

    ZR_D void UnpackCV(uint2 p, float3& c, float& var)
    {
        c = f3(half_lo(p.x), half_hi(p.x), half_lo(p.y));
        var = half_hi(p.y);
    }
    ZR_D uint2 PackCV(float3 c, float var) { return make_uint2(pack_half2(c.x, c.y), pack_half2(c.z, var)); }

    // -----------------------------------------------------------------------------------------------------------------
    // temporal accumulation + variance (orc_svgf_temporal). Internal planes are padded to `pitch` pixels per row.
    // -----------------------------------------------------------------------------------------------------------------
    __global__ void __launch_bounds__(256) k_svgf_temporal(zr_frame_constants fc, const uint4* __restrict__ core, const uint2* __restrict__ me,
        const float4* __restrict__ color, const uint2* __restrict__ prevGuide, const uint4* __restrict__ histPrev, int historyValid,
        uint4* __restrict__ histCurr, uint2* __restrict__ cv, uint2* __restrict__ guide, uint32_t pitch)
    {
        const int W = (int)fc.RenderWidth, H = (int)fc.RenderHeight;
        const int x = (int)(blockIdx.x * 32 + (threadIdx.x & 31)), y = (int)(blockIdx.y * 8 + (threadIdx.x >> 5));
        if (x >= W || y >= H) return;
        const size_t idx = (size_t)y * W + x, pidxOut = (size_t)y * pitch + x;
        const uint4 cr = ld128(&core[idx]);
        const float z = asfloat(cr.x);
        const float4 c4 = __ldg(&color[idx]);
        const float3 c = f3(c4.x, c4.y, c4.z);
        if (z == FLT_MAX_)
        {
            cv[pidxOut] = PackCV(c, 0.0f);
            guide[pidxOut] = make_uint2(asuint(FLT_MAX_), 0u);
            histCurr[pidxOut] = make_uint4(0u, 0u, 0u, 0u);
            return;
        }
        const float3 n = Math::DecodeUnitVector(Math::DecodeUNorm2(cr.y));
        const float l = Math::Luminance(c);
        float3 col = c; float m1 = l, m2 = l * l, N = 1.0f;
        if (historyValid)
        {
            const float2 renderDim = f2((float)W, (float)H);
            const float2 motionVec = unpack_snorm16x2(__ldg(&me[idx].x));
            const float2 currUV = f2((float)x + 0.5f, (float)y + 0.5f) / renderDim;
            const float2 prevUV = currUV - motionVec;
            const float2 pp = prevUV * renderDim;
            const int ppx = (int)pp.x, ppy = (int)pp.y;
            if (!(prevUV.x < 0.0f || prevUV.y < 0.0f || prevUV.x > 1.0f || prevUV.y > 1.0f) && ppx < W && ppy < H)
            {
                const size_t pidx = (size_t)ppy * pitch + ppx;
                const uint2 pg = __ldg(&prevGuide[pidx]);
                const float zp = asfloat(pg.x);
                if (zp != FLT_MAX_ && fabsf(zp - z) <= 0.1f * z)
                {
                    const float3 np = Math::DecodeUnitVector(Math::DecodeUNorm2(pg.y));
                    if (dot(np, n) >= 0.9f)
                    {
                        const uint4 h = ld128(&histPrev[pidx]);
                        const float3 hc = f3(half_lo(h.x), half_hi(h.x), half_lo(h.y));
                        const float hm1 = half_lo(h.z), hm2 = half_hi(h.z), hN = half_lo(h.w);
                        N = fminf(hN + 1.0f, 32.0f);
                        const float alpha = fmaxf(1.0f / N, 0.2f);
                        col = f3(fmaf(alpha, c.x - hc.x, hc.x), fmaf(alpha, c.y - hc.y, hc.y), fmaf(alpha, c.z - hc.z, hc.z));
                        m1 = fmaf(alpha, l - hm1, hm1);
                        m2 = fmaf(alpha, l * l - hm2, hm2);
                    }
                }
            }
        }
        float var = fmaxf(0.0f, m2 - m1 * m1);
        if (N < 4.0f)
        {
            float s1 = 0, s2 = 0, cnt = 0;
            for (int j = -1; j <= 1; j++)
                for (int i = -1; i <= 1; i++)
                {
                    const int tx = x + i, ty = y + j;
                    if (tx < 0 || ty < 0 || tx >= W || ty >= H) continue;
                    const size_t t = (size_t)ty * W + tx;
                    if (asfloat(__ldg(&core[t].x)) == FLT_MAX_) continue;
                    const float4 ct = __ldg(&color[t]);
                    const float lt = Math::Luminance(f3(ct.x, ct.y, ct.z));
                    s1 += lt; s2 = fmaf(lt, lt, s2); cnt += 1.0f;
                }
            const float mean = s1 / cnt;
            var = fmaxf(var, fmaxf(0.0f, s2 / cnt - mean * mean));
        }
        cv[pidxOut] = PackCV(col, var);
        guide[pidxOut] = make_uint2(cr.x, cr.y);
        st128(&histCurr[pidxOut], make_uint4(pack_half2(col.x, col.y), pack_half2(col.z, 0.0f), pack_half2(m1, m2), pack_half2(N, 0.0f)));
    }

    // -----------------------------------------------------------------------------------------------------------------
    // one a-trous pass. R = tap radius (1: 3x3, 2: 5x5); P = 2: strided lattice through 4-D maps (step >= 2), P = 1: step 1
    // -----------------------------------------------------------------------------------------------------------------
    template<int R, int P>
    struct AtrousTile
    {
        static constexpr int TU = P == 2 ? 32 : 64;                 // lattice columns of the output tile (x P phases each)
        static constexpr int TV = 16;
        // apron columns left and right. A TMA box must start on a 16-byte boundary in global memory: with 8-byte pixels the step-1
        // map needs an EVEN start column, so its apron is 2 even for the 3x3 taps (an odd start is an illegal-instruction fault,
        // r2i); in the lattice maps the innermost dimension is the 16-byte phase pair, so any lattice column will do.
        static constexpr int AX = P == 2 ? R : 2;
        static constexpr int ROW = (TU + 2 * AX) * P;               // staged elements per lattice row
        static constexpr int SV = TV + 2 * R;
        static constexpr int NS = ROW * SV;
        static constexpr int OUT = TU * P * TV;                     // 1024
        // shared memory layout (bytes)
        static constexpr int OFF_RAWC = 0;
        static constexpr int OFF_RAWG = OFF_RAWC + ((NS * 8 + 127) / 128) * 128;
        static constexpr int OFF_OUT = OFF_RAWG + ((NS * 8 + 127) / 128) * 128;
        static constexpr int OFF_F = OFF_OUT + OUT * 8;             // expanded box: 2 x float4 + 1 float per staged pixel
        static constexpr int OFF_BAR = OFF_F + 9 * NS * 4;
        static constexpr int BYTES = ((OFF_BAR + 8 + 127) / 128) * 128;
    };

    template<int R, int P, bool LAST>
    // The tensor maps are read from global memory (one address per (pass, plane), uploaded once when the pass is sized) instead of
    // travelling as 3 x 128 bytes of __grid_constant__ parameters with each of the five launches.
    __global__ void __launch_bounds__(512, 2) k_svgf_atrous(const CUtensorMap* __restrict__ pMapIn, const CUtensorMap* __restrict__ pMapGuide,
        const CUtensorMap* __restrict__ pMapOut, float4* __restrict__ outF, uint32_t W, uint32_t H, uint32_t step, uint32_t tilesU,
        SvgfParamsDev prm)
    {
        using T = AtrousTile<R, P>;
        extern __shared__ __align__(128) unsigned char smem[];
        uint2* rawC = reinterpret_cast<uint2*>(smem + T::OFF_RAWC);
        uint2* rawG = reinterpret_cast<uint2*>(smem + T::OFF_RAWG);
        uint2* outT = reinterpret_cast<uint2*>(smem + T::OFF_OUT);
        // expanded box: {r, g, b, variance}, {depth, normal}, luminance -- two 128-bit and one 32-bit shared-memory load per tap
        float4* s_cv = reinterpret_cast<float4*>(smem + T::OFF_F);
        float4* s_zn = s_cv + T::NS;
        float* s_lum = reinterpret_cast<float*>(s_zn + T::NS);
        uint64_t* bar = reinterpret_cast<uint64_t*>(smem + T::OFF_BAR);
        const uint32_t t = threadIdx.x;
        const int tu = (int)(blockIdx.x % tilesU), tv = (int)(blockIdx.x / tilesU);
        // phase of the sub-lattice: x-phases 2 * pair, 2 * pair + 1; y-phase py
        const int pair = P == 2 ? (int)(blockIdx.y % (step / 2)) : 0, py = P == 2 ? (int)(blockIdx.y / (step / 2)) : 0;
        const int u0 = tu * T::TU, v0 = tv * T::TV;
        if (t == 0)
        {
            tma::MbarInit(bar, 1);
            tma::FenceBarrierInit();
        }
        __syncthreads();
        if (t == 0)
        {
            tma::MbarArriveExpectTx(bar, 2u * T::NS * 8u);
            if (P == 2)
            {
                tma::Load4D(rawC, pMapIn, bar, 2 * pair, u0 - T::AX, py, v0 - R);
                tma::Load4D(rawG, pMapGuide, bar, 2 * pair, u0 - T::AX, py, v0 - R);
            }
            else
            {
                tma::Load2D(rawC, pMapIn, bar, u0 - T::AX, v0 - R);
                tma::Load2D(rawG, pMapGuide, bar, u0 - T::AX, v0 - R);
            }
        }
        tma::MbarWait(bar, 0);
        // expand the staged box once: halves -> float, luminance, normal decode; outside the image -> depth = FLT_MAX (weight 0)
        for (int e = (int)t; e < T::NS; e += 512)
        {
            const int sv = e / T::ROW, col = e % T::ROW;
            int x, y;
            if (P == 2) { x = (u0 - T::AX + (col >> 1)) * (int)step + 2 * pair + (col & 1); y = (v0 - R + sv) * (int)step + py; }
            else { x = u0 - T::AX + col; y = v0 - R + sv; }
            const bool inImg = x >= 0 && y >= 0 && x < (int)W && y < (int)H;
            float3 c = f3(0); float var = 0, z = FLT_MAX_; float3 n = f3(0);
            if (inImg)
            {
                UnpackCV(rawC[e], c, var);
                const uint2 g = rawG[e];
                z = asfloat(g.x);
                n = Math::DecodeUnitVector(Math::DecodeUNorm2(g.y));
            }
            s_cv[e] = f4(c.x, c.y, c.z, var);
            s_zn[e] = f4(z, n.x, n.y, n.z);
            s_lum[e] = Math::Luminance(c);
        }
        __syncthreads();
        constexpr float h5[5] = { 1.0f / 16, 1.0f / 4, 3.0f / 8, 1.0f / 4, 1.0f / 16 };
        constexpr float h3[3] = { 1.0f / 4, 1.0f / 2, 1.0f / 4 };
#pragma unroll
        for (int k = 0; k < 2; k++)
        {
            const int o = (int)t + k * 512;
            const int ov = o / (T::TU * P), ocol = o % (T::TU * P);
            const int ci = (ov + R) * T::ROW + ocol + T::AX * P;
            const float4 cvc = s_cv[ci], znc = s_zn[ci];
            const float zc = znc.x;
            uint2 packed = rawC[ci];
            float4 res = cvc;
            if (zc != FLT_MAX_)
            {
                const float3 cc = f3(cvc.x, cvc.y, cvc.z);
                const float varc = cvc.w, lc = s_lum[ci];
                const float3 nc = f3(znc.y, znc.z, znc.w);
                const float invZ = 1.0f / (prm.sigma_z * zc * (float)step);
                const float invL = 1.0f / fmaf(prm.sigma_l, sqrtf(fmaxf(varc, 0.0f)), 1e-4f);
                const float oneMinusK = 1.0f - prm.k_n;
                const float w0 = (R == 2 ? h5[2] : h3[1]) * (R == 2 ? h5[2] : h3[1]);
                float3 sumC = cc * w0;
                float sumV = (w0 * w0) * varc, sumW = w0;
#pragma unroll
                for (int j = -R; j <= R; j++)
#pragma unroll
                    for (int i = -R; i <= R; i++)
                    {
                        if (i == 0 && j == 0) continue;
                        const int ti = ci + j * T::ROW + i * P;
                        const float hw = (R == 2 ? h5[i + R] : h3[i + R]) * (R == 2 ? h5[j + R] : h3[j + R]);
                        const float4 cvt = s_cv[ti], znt = s_zn[ti];
                        const float wz = fmaf(-fabsf(znt.x - zc), invZ, 1.0f);
                        const float ndot = fmaf(znt.y, nc.x, fmaf(znt.z, nc.y, znt.w * nc.z));
                        const float wn = fmaf(ndot, prm.k_n, oneMinusK);
                        const float wl = fmaf(-fabsf(s_lum[ti] - lc), invL, 1.0f);
                        // wz, wl <= 1 by construction (1 - |d| * inv, one rounding), so max(., 0) is a saturate: one FFMA.SAT each
                        float w = hw * __saturatef(wz);
                        w = w * fmaxf(wn, 0.0f);
                        w = w * __saturatef(wl);
                        sumC = f3(fmaf(w, cvt.x, sumC.x), fmaf(w, cvt.y, sumC.y), fmaf(w, cvt.z, sumC.z));
                        sumV = fmaf(w * w, cvt.w, sumV);
                        sumW = sumW + w;
                    }
                const float3 oc = sumC / sumW;
                const float ovar = sumV / (sumW * sumW);
                packed = PackCV(oc, ovar);
                res = f4(oc.x, oc.y, oc.z, ovar);
            }
            if (LAST)
            {
                int x, y;
                if (P == 2) { x = (u0 + (ocol >> 1)) * (int)step + 2 * pair + (ocol & 1); y = (v0 + ov) * (int)step + py; }
                else { x = u0 + ocol; y = v0 + ov; }
                if (x < (int)W && y < (int)H)
                    outF[(size_t)y * W + x] = res;
            }
            else
                outT[o] = packed;
        }
        if (!LAST)
        {
            tma::FenceProxyAsync();
            __syncthreads();
            if (t == 0)
            {
                if (P == 2) tma::Store4D(pMapOut, outT, 2 * pair, u0, py, v0);
                else tma::Store2D(pMapOut, outT, u0, v0);
                tma::StoreCommit();
                tma::StoreWaitAll();
            }
        }
    }
}
} // namespace zr

// ---------------------------------------------------------------------------------------------------------------------
// pass object
// ---------------------------------------------------------------------------------------------------------------------
struct zr_svgf_pass
{
    static constexpr int MAX_PASSES = 5;
    uint32_t width = 0, height = 0, pitch = 0, rows = 0;      // pitch / rows: padded plane size in pixels (OnWindowResized)
    uint2* d_cv[2] = { nullptr, nullptr };
    uint2* d_guide[2] = { nullptr, nullptr };
    uint4* d_hist[2] = { nullptr, nullptr };
    float4* d_out = nullptr;
    int cur = 0;
    bool historyValid = false;
    zr_svgf_params params{};
    // tensor maps: [pass][plane]; load boxes depend on the radius, store boxes do not
    CUtensorMap mapCvLoad[MAX_PASSES][2], mapCvStore[MAX_PASSES][2], mapGuideLoad[MAX_PASSES][2];
    CUtensorMap* d_maps = nullptr;      // device copy: [kind 0 = cv load, 1 = cv store, 2 = guide load][pass][plane]
    const CUtensorMap* DevMap(int kind, int k, int plane) const { return d_maps + ((kind * MAX_PASSES + k) * 2 + plane); }
    bool mapsReady = false;

    static void Defaults(zr_svgf_params* p) { p->sigma_z = 0.02f; p->k_n = 16.0f; p->sigma_l = 4.0f; p->radius = 2; p->num_passes = 5; }

    void Release()
    {
        for (int i = 0; i < 2; i++)
        {
            if (d_cv[i]) cudaFree(d_cv[i]); if (d_guide[i]) cudaFree(d_guide[i]); if (d_hist[i]) cudaFree(d_hist[i]);
            d_cv[i] = nullptr; d_guide[i] = nullptr; d_hist[i] = nullptr;
        }
        if (d_out) cudaFree(d_out);
        if (d_maps) cudaFree(d_maps);
        d_out = nullptr; d_maps = nullptr; mapsReady = false;
    }

    zr_status EncodeMaps()
    {
        using namespace zr;
        const uint32_t R = params.radius;
        for (int k = 0; k < MAX_PASSES; k++)
        {
            const uint64_t s = 1ull << k;
            for (int pl = 0; pl < 2; pl++)
            {
                bool ok = true;
                if (s == 1)
                {
                    const uint64_t dims[2] = { pitch, rows };
                    const uint64_t strides[1] = { (uint64_t)pitch * 8 };
                    const uint32_t boxL[2] = { 64 + 2 * 2, 16 + 2 * R }, boxS[2] = { 64, 16 };       // AtrousTile<R, 1>::AX == 2
                    ok = ok && tma::EncodeWords(&mapCvLoad[k][pl], d_cv[pl], 2, dims, strides, boxL);
                    ok = ok && tma::EncodeWords(&mapGuideLoad[k][pl], d_guide[pl], 2, dims, strides, boxL);
                    ok = ok && tma::EncodeWords(&mapCvStore[k][pl], d_cv[pl], 2, dims, strides, boxS);
                }
                else
                {
                    // {phase_x, u, phase_y, v}: pixel (u s + phase_x, v s + phase_y)
                    const uint64_t dims[4] = { s, pitch / s, s, rows / s };
                    const uint64_t strides[3] = { s * 8, (uint64_t)pitch * 8, s * (uint64_t)pitch * 8 };
                    const uint32_t boxL[4] = { 2, 32 + 2 * R, 1, 16 + 2 * R }, boxS[4] = { 2, 32, 1, 16 };
                    ok = ok && tma::EncodeWords(&mapCvLoad[k][pl], d_cv[pl], 4, dims, strides, boxL);
                    ok = ok && tma::EncodeWords(&mapGuideLoad[k][pl], d_guide[pl], 4, dims, strides, boxL);
                    ok = ok && tma::EncodeWords(&mapCvStore[k][pl], d_cv[pl], 4, dims, strides, boxS);
                }
                if (!ok)
                {
                    set_error("zr_svgf_pass: cuTensorMapEncodeTiled failed (step %u)", (unsigned)s);
                    return ZR_ERR_CUDA;
                }
            }
        }
        // no kernel may still be using the old maps, and the new ones are in place before the next launch
        ZR_CUDA(cudaDeviceSynchronize());
        if (!d_maps) ZR_CUDA(cudaMalloc(&d_maps, sizeof(CUtensorMap) * 3 * MAX_PASSES * 2));
        CUtensorMap host[3][MAX_PASSES][2];
        memcpy(host[0], mapCvLoad, sizeof(mapCvLoad)); memcpy(host[1], mapCvStore, sizeof(mapCvStore)); memcpy(host[2], mapGuideLoad, sizeof(mapGuideLoad));
        ZR_CUDA(cudaMemcpy(d_maps, host, sizeof(host), cudaMemcpyHostToDevice));
        ZR_CUDA(cudaDeviceSynchronize());
        mapsReady = true;
        return ZR_OK;
    }

    zr_status OnWindowResized(uint32_t w, uint32_t h)
    {
        Release();
        width = w; height = h;
        // padded so that (a) every lattice view divides evenly (multiples of 32 >= 2 * 16) and (b) no TMA box is larger than the
        // tensor it is cut from, even for the coarsest lattice (step 16: 36 x 20 lattice points) of a small image
        pitch = (w + 31) / 32 * 32; rows = (h + 31) / 32 * 32;
        if (pitch < 16u * 36u) pitch = 16u * 36u;
        if (rows < 16u * 20u) rows = 16u * 20u;
        const size_t n = (size_t)pitch * rows;
        for (int i = 0; i < 2; i++)
        {
            ZR_CUDA(cudaMalloc(&d_cv[i], n * 8));
            ZR_CUDA(cudaMalloc(&d_guide[i], n * 8));
            ZR_CUDA(cudaMalloc(&d_hist[i], n * 16));
        }
        ZR_CUDA(cudaMalloc(&d_out, (size_t)w * h * 16));
        zr_status st = ResetTemporal();
        if (st != ZR_OK) return st;
        return EncodeMaps();
    }

    zr_status ResetTemporal()
    {
        const size_t n = (size_t)pitch * rows;
        ZR_CLEAR_BEGIN();
        for (int i = 0; i < 2; i++)
        {
            ZR_CUDA(cudaMemset(d_cv[i], 0, n * 8));
            ZR_CUDA(cudaMemset(d_guide[i], 0, n * 8));
            ZR_CUDA(cudaMemset(d_hist[i], 0, n * 16));
        }
        ZR_CUDA(cudaMemset(d_out, 0, (size_t)width * height * 16));
        ZR_CLEAR_END();
        historyValid = false;
        cur = 0;
        return ZR_OK;
    }

    template<int R>
    zr_status LaunchAtrous(int k, int inPlane, bool last, cudaStream_t stream)
    {
        using namespace zr;
        const uint32_t s = 1u << k;
        const SvgfParamsDev prm{ params.sigma_z, params.k_n, params.sigma_l };
        const CUtensorMap* mIn = DevMap(0, k, inPlane);
        const CUtensorMap* mG = DevMap(2, k, cur);
        const CUtensorMap* mOut = DevMap(1, k, 1 - inPlane);
        ZR_PROF("k_svgf_atrous", stream);
        if (s == 1)
        {
            using T = AtrousTile<R, 1>;
            const uint32_t tilesU = (width + T::TU - 1) / T::TU, tilesV = (height + T::TV - 1) / T::TV;
            if (last) k_svgf_atrous<R, 1, true><<<dim3(tilesU * tilesV, 1), 512, T::BYTES, stream>>>(mIn, mG, mOut, d_out, width, height, s, tilesU, prm);
            else k_svgf_atrous<R, 1, false><<<dim3(tilesU * tilesV, 1), 512, T::BYTES, stream>>>(mIn, mG, mOut, d_out, width, height, s, tilesU, prm);
        }
        else
        {
            using T = AtrousTile<R, 2>;
            const uint32_t latW = (width + s - 1) / s, latH = (height + s - 1) / s;
            const uint32_t tilesU = (latW + T::TU - 1) / T::TU, tilesV = (latH + T::TV - 1) / T::TV;
            const dim3 grid(tilesU * tilesV, (s / 2) * s);
            if (last) k_svgf_atrous<R, 2, true><<<grid, 512, T::BYTES, stream>>>(mIn, mG, mOut, d_out, width, height, s, tilesU, prm);
            else k_svgf_atrous<R, 2, false><<<grid, 512, T::BYTES, stream>>>(mIn, mG, mOut, d_out, width, height, s, tilesU, prm);
        }
        ZR_LAUNCH_CHECK();
        if (getenv("ZR_SVGF_DEBUG"))
        {
            cudaError_t e = cudaStreamSynchronize(stream);
            if (e != cudaSuccess) { set_error("k_svgf_atrous pass %d (step %u, radius %d, last %d): %s", k, s, R, (int)last, cudaGetErrorString(e)); return ZR_ERR_CUDA; }
        }
        return ZR_OK;
    }

    zr_status Render(const zr_frame_inputs* in, const void* d_signal, cudaStream_t stream)
    {
        using namespace zr;
        if (!in || !in->curr.d_core || !in->curr.d_motion_emissive || !d_signal)
        {
            set_error("zr_svgf_pass_render: missing input");
            return ZR_ERR_INVALID_ARG;
        }
        if (in->frame.RenderWidth != width || in->frame.RenderHeight != height)
        {
            set_error("zr_svgf_pass_render: frame/pass size mismatch");
            return ZR_ERR_INVALID_ARG;
        }
        if (!mapsReady) { set_error("zr_svgf_pass_render: tensor maps are not initialised"); return ZR_ERR_NOT_INITIALIZED; }
        cur = 1 - cur;
        {
            ZR_PROF("k_svgf_temporal", stream);
            k_svgf_temporal<<<dim3((width + 31) / 32, (height + 7) / 8), 256, 0, stream>>>(in->frame, (const uint4*)in->curr.d_core,
                (const uint2*)in->curr.d_motion_emissive, (const float4*)d_signal, d_guide[1 - cur], d_hist[1 - cur], historyValid ? 1 : 0,
                d_hist[cur], d_cv[0], d_guide[cur], pitch);
            ZR_LAUNCH_CHECK();
        }
        int plane = 0;
        for (uint32_t k = 0; k < params.num_passes; k++)
        {
            const bool last = k + 1 == params.num_passes;
            zr_status st = params.radius == 2 ? LaunchAtrous<2>((int)k, plane, last, stream) : LaunchAtrous<1>((int)k, plane, last, stream);
            if (st != ZR_OK) return st;
            plane = 1 - plane;
        }
        historyValid = true;
        return ZR_OK;
    }
};

extern "C"
{
    zr_status zr_svgf_pass_create(uint32_t width, uint32_t height, zr_svgf_pass** out)
    {
        if (!out || !width || !height) { zr::set_error("zr_svgf_pass_create: bad args"); return ZR_ERR_INVALID_ARG; }
        zr_svgf_pass* p = new zr_svgf_pass();
        zr_svgf_pass::Defaults(&p->params);
        using namespace zr;
        cudaError_t e = cudaSuccess;
#define ZR_SVGF_ATTR(R, P, LAST) \
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_svgf_atrous<R, P, LAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtrousTile<R, P>::BYTES)
        ZR_SVGF_ATTR(1, 1, false); ZR_SVGF_ATTR(1, 1, true); ZR_SVGF_ATTR(1, 2, false); ZR_SVGF_ATTR(1, 2, true);
        ZR_SVGF_ATTR(2, 1, false); ZR_SVGF_ATTR(2, 1, true); ZR_SVGF_ATTR(2, 2, false); ZR_SVGF_ATTR(2, 2, true);
#undef ZR_SVGF_ATTR
        if (e != cudaSuccess) { delete p; zr::cuda_fail(e, "cudaFuncSetAttribute(k_svgf_atrous)"); return ZR_ERR_CUDA; }
        zr_status s = p->OnWindowResized(width, height);
        if (s != ZR_OK) { p->Release(); delete p; return s; }
        *out = p;
        return ZR_OK;
    }
    zr_status zr_svgf_pass_resize(zr_svgf_pass* p, uint32_t width, uint32_t height)
    {
        if (!p || !width || !height) return ZR_ERR_INVALID_ARG;
        return p->OnWindowResized(width, height);
    }
    zr_status zr_svgf_pass_reset_temporal(zr_svgf_pass* p) { return p ? p->ResetTemporal() : ZR_ERR_INVALID_ARG; }
    zr_status zr_svgf_pass_default_params(zr_svgf_params* out)
    {
        if (!out) return ZR_ERR_INVALID_ARG;
        zr_svgf_pass::Defaults(out);
        return ZR_OK;
    }
    zr_status zr_svgf_pass_set_params(zr_svgf_pass* p, const zr_svgf_params* params)
    {
        if (!p || !params) return ZR_ERR_INVALID_ARG;
        if ((params->radius != 1 && params->radius != 2) || params->num_passes < 1 || params->num_passes > zr_svgf_pass::MAX_PASSES ||
            !(params->sigma_z > 0) || !(params->sigma_l > 0) || !(params->k_n >= 0))
        {
            zr::set_error("zr_svgf_pass_set_params: radius must be 1 or 2, 1..5 passes, positive sigmas");
            return ZR_ERR_INVALID_ARG;
        }
        const bool remap = params->radius != p->params.radius;
        p->params = *params;
        return remap ? p->EncodeMaps() : ZR_OK;
    }
    zr_status zr_svgf_pass_render(zr_svgf_pass* p, const zr_frame_inputs* in, const void* d_signal, void* stream)
    {
        if (!p) return ZR_ERR_INVALID_ARG;
        return p->Render(in, d_signal, (cudaStream_t)stream);
    }
    zr_status zr_svgf_pass_get_output(zr_svgf_pass* p, zr_svgf_output id, zr_image2d* out)
    {
        if (!p || !out) return ZR_ERR_INVALID_ARG;
        const uint32_t w = p->width, h = p->height;
        switch (id)
        {
        case ZR_SVGF_DENOISED: *out = zr_image2d{ p->d_out, w, h, w * 16u, 16u }; break;
        case ZR_SVGF_ACCUMULATED: *out = zr_image2d{ p->d_cv[0], w, h, p->pitch * 8u, 8u }; break;       // only valid with num_passes == 1 .. see header
        case ZR_SVGF_GUIDE: *out = zr_image2d{ p->d_guide[p->cur], w, h, p->pitch * 8u, 8u }; break;
        case ZR_SVGF_HISTORY: *out = zr_image2d{ p->d_hist[p->cur], w, h, p->pitch * 16u, 16u }; break;
        default: zr::set_error("zr_svgf_pass_get_output: unknown output id"); return ZR_ERR_INVALID_ARG;
        }
        return ZR_OK;
    }
    void zr_svgf_pass_destroy(zr_svgf_pass* p) { if (p) { p->Release(); delete p; } }
}
