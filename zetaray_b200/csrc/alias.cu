// alias.cu -- emissive-power alias table built on the device.
//
// Replaces the reference's GPU -> CPU -> GPU round trip (readback, CPU BuildAliasTable, upload:
// ZetaRenderPass/PreLighting/PreLighting.cpp:27-158, 512-585). Output is bit-identical to the CPU
// algorithm for 32-byte aligned input (the production case: mapped readback memory):
//   * Math::KahanSum        (ZetaCore/Math/Common.cpp:72-140): the 8 AVX lanes are 8 CUDA lanes, each
//                            running its own Kahan chain over (V1 + V2) pairs, then the scalar fold
//   * AliasTable_Normalize  (ZetaCore/Math/Sampling.cpp:13-50)
//   * Vose with LIFO stacks (PreLighting.cpp:70-137): the stacks are built by an order-preserving
//     partition; the inherently serial pairing loop runs on lane 0 of one warp while all 32 lanes
//     stream the next window of both stacks into shared memory, so the serial chain only touches
//     shared memory and registers.
#include "zr_common.cuh"

namespace zr
{
namespace
{
    constexpr int SUM_THREADS = 32;

    // One warp. Lanes 0..7 own the 8 SIMD accumulators; every lane helps loading.
    __global__ void k_kahan_sum(const float* __restrict__ w, uint32_t N, float* __restrict__ out /* [0]=sum, [1]=N/sum */)
    {
        const uint32_t lane = threadIdx.x;
        const uint32_t numSIMD = N - (N & 15u);
        float vSum = 0.0f, vComp = 0.0f;
        // 32 lanes load 32 consecutive floats = two SIMD iterations
        for (uint32_t base = 0; base < numSIMD; base += 32)
        {
            float v = (base + lane < numSIMD) ? w[base + lane] : 0.0f;
            // iteration A: elements [base, base+16): lane l (<8) needs v[l] + v[l+8]
            float a0 = __shfl_sync(0xffffffffu, v, lane & 7);
            float a1 = __shfl_sync(0xffffffffu, v, (lane & 7) + 8);
            float b0 = __shfl_sync(0xffffffffu, v, (lane & 7) + 16);
            float b1 = __shfl_sync(0xffffffffu, v, (lane & 7) + 24);
            if (lane < 8)
            {
                {
                    float vCurr = a0 + a1;
                    float vCorrected = vCurr - vComp;
                    float vNewSum = vSum + vCorrected;
                    float c = vNewSum - vSum;
                    vComp = c - vCorrected;
                    vSum = vNewSum;
                }
                if (base + 16 < numSIMD)
                {
                    float vCurr = b0 + b1;
                    float vCorrected = vCurr - vComp;
                    float vNewSum = vSum + vCorrected;
                    float c = vNewSum - vSum;
                    vComp = c - vCorrected;
                    vSum = vNewSum;
                }
            }
        }
        // scalar fold of the 8 lanes, then the tail (lane 0)
        float sum = 0.0f, compensation = 0.0f;
        for (int l = 0; l < 8; l++)
        {
            float s = __shfl_sync(0xffffffffu, vSum, l);
            float c = __shfl_sync(0xffffffffu, vComp, l);
            float corrected = s - compensation - c;
            float newSum = sum + corrected;
            compensation = (newSum - sum) - corrected;
            sum = newSum;
        }
        if (lane == 0)
        {
            for (uint32_t i = numSIMD; i < N; i++)
            {
                float corrected = w[i] - compensation;
                float newSum = sum + corrected;
                compensation = (newSum - sum) - corrected;
                sum = newSum;
            }
            out[0] = sum;
            out[1] = (float)N / sum;
        }
    }

    // weights *= N / sum; CachedP_Orig = w * (1/N)
    __global__ void k_normalize(float* __restrict__ w, uint32_t N, const float* __restrict__ sums,
        zr_alias_entry* __restrict__ table)
    {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= N) return;
        const float sumRcp = sums[1];
        const float oneDivN = 1.0f / (float)N;
        const float p = w[i] * sumRcp;
        w[i] = p;
        table[i].CachedP_Orig = p * oneDivN;
    }

    // Order-preserving partition into smaller (< 1) and larger (>= 1) stacks. One block, chunked scan.
    constexpr int PART_THREADS = 1024;
    __global__ void k_partition(const float* __restrict__ w, uint32_t N, uint32_t* __restrict__ smaller,
        uint32_t* __restrict__ larger, uint32_t* __restrict__ counts)
    {
        __shared__ uint32_t s_warp[32];
        __shared__ uint32_t s_base[2];
        const uint32_t tid = threadIdx.x;
        const uint32_t lane = tid & 31, warp = tid >> 5;
        if (tid == 0) { s_base[0] = 0; s_base[1] = 0; }
        __syncthreads();
        for (uint32_t base = 0; base < N; base += PART_THREADS)
        {
            const uint32_t i = base + tid;
            const bool valid = i < N;
            const bool isSmall = valid && (w[i] < 1.0f);
            const uint32_t ballot = __ballot_sync(0xffffffffu, isSmall);
            const uint32_t prefix = __popc(ballot & ((1u << lane) - 1u));
            if (lane == 0) s_warp[warp] = __popc(ballot);
            __syncthreads();
            uint32_t warpOff = 0, total = 0;
            for (uint32_t k = 0; k < PART_THREADS / 32; k++)
            {
                uint32_t c = s_warp[k];
                if (k < warp) warpOff += c;
                total += c;
            }
            const uint32_t smallRank = s_base[0] + warpOff + prefix;
            const uint32_t largeRank = s_base[1] + (tid - (warpOff + prefix));
            if (valid)
            {
                if (isSmall) smaller[smallRank] = i;
                else larger[largeRank] = i;
            }
            __syncthreads();
            if (tid == 0)
            {
                const uint32_t nValid = min((uint32_t)PART_THREADS, N - base);
                s_base[0] += total;
                s_base[1] += nValid - total;
            }
            __syncthreads();
        }
        if (tid == 0) { counts[0] = s_base[0]; counts[1] = s_base[1]; }
    }

    // Vose pairing. One warp; lane 0 runs the serial chain, all lanes refill the stack windows.
    constexpr int WIN = 1024;
    __global__ void k_vose(const float* __restrict__ w, const uint32_t* __restrict__ smaller,
        const uint32_t* __restrict__ larger, const uint32_t* __restrict__ counts,
        zr_alias_entry* __restrict__ table)
    {
        __shared__ uint32_t s_si[WIN];
        __shared__ float s_sp[WIN];
        __shared__ uint32_t s_li[WIN];
        __shared__ float s_lp[WIN];
        __shared__ int s_state[4];   // 0: small window count, 1: large window count, 2: done
        const uint32_t lane = threadIdx.x;

        // stack tops (number of elements still in the global part of each stack)
        int smallTop = (int)counts[0];
        int largeTop = (int)counts[1];
        // window cursors (lane 0 only)
        int sw = 0, swN = 0, lw = 0, lwN = 0;
        // a larger entry that dropped below 1 is the top of the smaller stack
        bool pending = false;
        uint32_t pendIdx = 0;
        float pendP = 0.0f;
        // current larger
        bool haveLarge = false;
        uint32_t largeIdx = 0;
        float largeP = 0.0f;

        for (;;)
        {
            // refill windows when lane 0 has run dry
            if (lane == 0)
            {
                s_state[0] = (sw >= swN) ? 1 : 0;
                s_state[1] = (lw >= lwN) ? 1 : 0;
            }
            __syncwarp();
            const bool refillS = s_state[0] != 0;
            const bool refillL = s_state[1] != 0;
            if (refillS)
            {
                const int n = min(WIN, smallTop);
                for (int k = (int)lane; k < n; k += 32)
                {
                    const uint32_t idx = smaller[smallTop - 1 - k];     // pop order = from the back
                    s_si[k] = idx;
                    s_sp[k] = w[idx];
                }
                smallTop -= n;
                sw = 0; swN = n;
            }
            if (refillL)
            {
                const int n = min(WIN, largeTop);
                for (int k = (int)lane; k < n; k += 32)
                {
                    const uint32_t idx = larger[largeTop - 1 - k];
                    s_li[k] = idx;
                    s_lp[k] = w[idx];
                }
                largeTop -= n;
                lw = 0; lwN = n;
            }
            __syncwarp();

            if (lane == 0)
            {
                bool done = false;
                for (;;)
                {
                    const bool smallEmpty = !pending && (sw >= swN);
                    const bool largeEmpty = !haveLarge && (lw >= lwN);
                    // a window ran dry but its global stack still has entries -> refill
                    if ((smallEmpty && smallTop > 0) || (largeEmpty && largeTop > 0))
                        break;
                    if (smallEmpty || largeEmpty)
                    {
                        done = true;
                        break;
                    }
                    uint32_t smallerIdx;
                    float smallerProb;
                    if (pending) { smallerIdx = pendIdx; smallerProb = pendP; pending = false; }
                    else { smallerIdx = s_si[sw]; smallerProb = s_sp[sw]; sw++; }
                    if (!haveLarge) { largeIdx = s_li[lw]; largeP = s_lp[lw]; lw++; haveLarge = true; }

                    table[smallerIdx].Alias = largeIdx;
                    table[smallerIdx].P_Curr = smallerProb;
                    largeP = (smallerProb + largeP) - 1.0f;
                    if (largeP < 1.0f)
                    {
                        haveLarge = false;
                        pending = true;
                        pendIdx = largeIdx;
                        pendP = largeP;
                    }
                }
                s_state[2] = done ? 1 : 0;
            }
            __syncwarp();
            if (s_state[2])
                break;
        }

        // leftovers alias to themselves with P_Curr = 1 (PreLighting.cpp:108-134)
        // lane 0 state is broadcast so all lanes can help
        int r_sw = __shfl_sync(0xffffffffu, sw, 0), r_swN = __shfl_sync(0xffffffffu, swN, 0);
        int r_lw = __shfl_sync(0xffffffffu, lw, 0), r_lwN = __shfl_sync(0xffffffffu, lwN, 0);
        const int r_pending = __shfl_sync(0xffffffffu, (int)pending, 0);
        const uint32_t r_pendIdx = __shfl_sync(0xffffffffu, pendIdx, 0);
        const int r_haveLarge = __shfl_sync(0xffffffffu, (int)haveLarge, 0);
        const uint32_t r_largeIdx = __shfl_sync(0xffffffffu, largeIdx, 0);
        if (lane == 0)
        {
            if (r_pending) { table[r_pendIdx].Alias = r_pendIdx; table[r_pendIdx].P_Curr = 1.0f; }
            if (r_haveLarge) { table[r_largeIdx].Alias = r_largeIdx; table[r_largeIdx].P_Curr = 1.0f; }
        }
        for (int k = r_sw + (int)lane; k < r_swN; k += 32) { uint32_t i = s_si[k]; table[i].Alias = i; table[i].P_Curr = 1.0f; }
        for (int k = r_lw + (int)lane; k < r_lwN; k += 32) { uint32_t i = s_li[k]; table[i].Alias = i; table[i].P_Curr = 1.0f; }
        for (int k = (int)lane; k < smallTop; k += 32) { uint32_t i = smaller[k]; table[i].Alias = i; table[i].P_Curr = 1.0f; }
        for (int k = (int)lane; k < largeTop; k += 32) { uint32_t i = larger[k]; table[i].Alias = i; table[i].P_Curr = 1.0f; }
    }

    __global__ void k_cache_alias_p(zr_alias_entry* __restrict__ table, uint32_t N)
    {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= N) return;
        table[i].CachedP_Alias = table[table[i].Alias].CachedP_Orig;
    }

    // Light::AliasTableSample::get, one RNG stream
    __global__ void k_sample(const zr_alias_entry* __restrict__ table, uint32_t N, uint32_t seed,
        uint32_t numDraws, uint32_t* __restrict__ outIdx, float* __restrict__ outPdf)
    {
        if (blockIdx.x != 0 || threadIdx.x != 0) return;
        RNG rng = RNG::InitSeed(seed);
        for (uint32_t d = 0; d < numDraws; d++)
        {
            const uint32_t u0 = rng.UniformUintBounded(N);
            const zr_alias_entry s = table[u0];
            if (rng.Uniform() < s.P_Curr) { outPdf[d] = s.CachedP_Orig; outIdx[d] = u0; }
            else { outPdf[d] = s.CachedP_Alias; outIdx[d] = s.Alias; }
        }
    }
}

zr_status alias_table_build(float* d_weights, uint32_t n, zr_alias_entry* d_table, uint32_t* d_scratch,
    cudaStream_t stream)
{
    if (!d_weights || !d_table || !d_scratch || n == 0)
    {
        set_error("zr_alias_table_build: null pointer or n == 0");
        return ZR_ERR_INVALID_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(d_weights) & 31) != 0)
    {
        set_error("zr_alias_table_build: d_weights must be 32-byte aligned (reference parity, SURVEY 8a-1)");
        return ZR_ERR_INVALID_ARG;
    }
    // scratch: [0, n) smaller stack, [n, 2n) larger stack, [2n, 2n + 16) Kahan sums and partition counts -- all in the caller's buffer, so
    // builds on different streams, devices or threads share nothing
    float* d_sums = reinterpret_cast<float*>(d_scratch + 2 * (size_t)n);
    uint32_t* d_counts = d_scratch + 2 * (size_t)n + 4;

    ZR_PROF("k_kahan_sum", stream);
    k_kahan_sum<<<1, SUM_THREADS, 0, stream>>>(d_weights, n, d_sums);
    ZR_LAUNCH_CHECK();
    ZR_PROF("k_normalize", stream);
    k_normalize<<<(n + 255) / 256, 256, 0, stream>>>(d_weights, n, d_sums, d_table);
    ZR_LAUNCH_CHECK();
    ZR_PROF("k_partition", stream);
    k_partition<<<1, PART_THREADS, 0, stream>>>(d_weights, n, d_scratch, d_scratch + n, d_counts);
    ZR_LAUNCH_CHECK();
    ZR_PROF("k_vose", stream);
    k_vose<<<1, 32, 0, stream>>>(d_weights, d_scratch, d_scratch + n, d_counts, d_table);
    ZR_LAUNCH_CHECK();
    ZR_PROF("k_cache_alias_p", stream);
    k_cache_alias_p<<<(n + 255) / 256, 256, 0, stream>>>(d_table, n);
    ZR_LAUNCH_CHECK();
    return ZR_OK;
}

zr_status alias_table_sample(const zr_alias_entry* d_table, uint32_t n, uint32_t seed, uint32_t num_draws,
    uint32_t* d_out_idx, float* d_out_pdf, cudaStream_t stream)
{
    if (!d_table || !d_out_idx || !d_out_pdf || n == 0)
    {
        set_error("zr_alias_table_sample: null pointer or n == 0");
        return ZR_ERR_INVALID_ARG;
    }
    ZR_PROF("k_sample", stream);
    k_sample<<<1, 32, 0, stream>>>(d_table, n, seed, num_draws, d_out_idx, d_out_pdf);
    ZR_LAUNCH_CHECK();
    return ZR_OK;
}
} // namespace zr

extern "C"
{
    zr_status zr_alias_table_build(float* d_weights, uint32_t n, zr_alias_entry* d_table, uint32_t* d_scratch, void* stream)
    {
        return zr::alias_table_build(d_weights, n, d_table, d_scratch, (cudaStream_t)stream);
    }
    zr_status zr_alias_table_sample(const zr_alias_entry* d_table, uint32_t n, uint32_t seed, uint32_t num_draws,
        uint32_t* d_out_idx, float* d_out_pdf, void* stream)
    {
        return zr::alias_table_sample(d_table, n, seed, num_draws, d_out_idx, d_out_pdf, (cudaStream_t)stream);
    }
}
