// hostsim.cpp -- TEST INFRASTRUCTURE. Compiles the product's own traversal source (zetaray_b200/csrc/zr_scene.cuh:
// Traverse<0/1>, TriHit, the node decode) for the host with g++, so that the BVH builder + traversal pair can be checked
// against brute force on scenes of the benchmark's size (10^5 - 10^6 triangles) in the CPU test tier, where no GPU exists.
// Nothing here is shipped or used by the product; the product path is the same source compiled by nvcc for sm_100a.
#include <cstring>
#include <cstdint>
#include <cmath>
#include <vector>
#include <thread>
#include <atomic>
#include <cuda_runtime.h>      // vector types + make_float3 (host-usable)

// host stand-ins for the few device intrinsics the headers use
template<typename T> static inline T __ldg(const T* p) { return *p; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
template<typename T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }

struct ZrTraverseStats { unsigned long long nodes, tris; };
static thread_local ZrTraverseStats g_stats;
#define ZR_TRAVERSE_STATS g_stats
#include "../../zetaray_b200/csrc/zr_scene.cuh"

namespace zr
{
void set_error(const char*, ...) {}
}

extern "C"
{
// rays: n x {origin xyz, tmin, dir xyz, tmax}; hits: n x {t, u, v, triGlobal-as-float-bits}; flags (any-hit) may be null
int hostsim_trace(const void* nodes, const float* leafTris /* 12 floats per tri, leaf order */, const uint32_t* triMesh,
    const uint32_t* meshFirstTri, const float* rays, uint32_t n, float* hits, uint32_t* anyFlags, const uint32_t* ignoreIDs, int threads)
{
    zr::SceneDev sc{};
    sc.nodes = reinterpret_cast<const uint4*>(nodes);
    sc.tris = reinterpret_cast<const float4*>(leafTris);
    sc.triMesh = triMesh;
    sc.meshFirstTri = meshFirstTri;
    std::atomic<uint32_t> next{0};
    auto work = [&]() {
        for (;;)
        {
            const uint32_t b = next.fetch_add(256);
            if (b >= n) break;
            for (uint32_t i = b; i < std::min(n, b + 256); i++)
            {
                const float* r = rays + (size_t)i * 8;
                const float3 o = make_float3(r[0], r[1], r[2]), d = make_float3(r[4], r[5], r[6]);
                if (hits)
                {
                    zr::RayHit h = zr::TraceClosest(sc, o, d, r[3], r[7]);
                    float* out = hits + (size_t)i * 4;
                    out[0] = h.hit ? h.t : zr::FLT_MAX_; out[1] = h.bary.x; out[2] = h.bary.y; out[3] = zr::asfloat(h.tri);
                }
                if (anyFlags)
                    anyFlags[i] = zr::TraceAnyExcept(sc, o, d, r[3], r[7], ignoreIDs ? ignoreIDs[i] : 0xffffffffu) ? 1u : 0u;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < std::max(threads, 1); t++) pool.emplace_back(work);
    for (auto& t : pool) t.join();
    return 0;
}

// brute force with the same hit rule (closest t, ties to the lowest global triangle index); worldTris 9 floats per tri
int hostsim_brute(const float* worldTris, uint32_t numTris, const float* rays, uint32_t n, float* hits, int threads)
{
    std::atomic<uint32_t> next{0};
    auto work = [&]() {
        for (;;)
        {
            const uint32_t i = next.fetch_add(1);
            if (i >= n) break;
            const float* r = rays + (size_t)i * 8;
            const float3 o = make_float3(r[0], r[1], r[2]), d = make_float3(r[4], r[5], r[6]);
            bool hit = false; float bt = r[7], bu = 0, bv = 0; uint32_t btri = 0xffffffffu;
            for (uint32_t k = 0; k < numTris; k++)
            {
                const float* w = worldTris + (size_t)k * 9;
                float t, u, v;
                if (zr::TriHit(o, d, make_float3(w[0], w[1], w[2]), make_float3(w[3], w[4], w[5]), make_float3(w[6], w[7], w[8]), r[3], r[7], t, u, v))
                    if (!hit || t < bt) { hit = true; bt = t; bu = u; bv = v; btri = k; }
            }
            float* out = hits + (size_t)i * 4;
            out[0] = hit ? bt : zr::FLT_MAX_; out[1] = bu; out[2] = bv; out[3] = zr::asfloat(btri);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < std::max(threads, 1); t++) pool.emplace_back(work);
    for (auto& t : pool) t.join();
    return 0;
}
}

// Structural check of a built tree, decoding the quantised boxes exactly as Traverse does. Returns the number of
// violations: a triangle vertex outside the box of a child it hangs under, a triangle referenced != 1 times, a child
// offset out of range. stats = {nodes visited, leaves, max depth, triangles seen}
extern "C" uint64_t hostsim_validate(const void* nodes_, uint32_t numNodes, const float* leafTris, uint32_t numTris, uint64_t stats[4])
{
    const zr::BVH8Node* nodes = reinterpret_cast<const zr::BVH8Node*>(nodes_);
    std::vector<uint8_t> seen(numTris, 0);
    std::vector<uint8_t> nodeSeen(numNodes, 0);
    uint64_t bad = 0;
    stats[0] = stats[1] = stats[2] = stats[3] = 0;
    struct Item { uint32_t node; float lo[3], hi[3]; uint32_t depth; };
    std::vector<Item> stack;
    Item root; root.node = 0; root.depth = 1;
    for (int a = 0; a < 3; a++) { root.lo[a] = -INFINITY; root.hi[a] = INFINITY; }
    stack.push_back(root);
    while (!stack.empty())
    {
        Item it = stack.back(); stack.pop_back();
        if (it.node >= numNodes) { bad++; continue; }
        if (nodeSeen[it.node]++) { bad++; continue; }
        stats[0]++; stats[2] = std::max<uint64_t>(stats[2], it.depth);
        const zr::BVH8Node& n = nodes[it.node];
        const float p[3] = { n.px, n.py, n.pz };
        const uint8_t ex[3] = { n.ex, n.ey, n.ez };
        float s[3];
        for (int a = 0; a < 3; a++) s[a] = __uint_as_float((uint32_t)ex[a] << 23);
        for (int c = 0; c < 8; c++)
        {
            const uint32_t meta = n.meta[c];
            if (meta == 0) continue;
            float lo[3], hi[3];
            for (int a = 0; a < 3; a++)
            {
                lo[a] = fmaf((float)n.qlo[a][c], s[a], p[a]);
                hi[a] = fmaf((float)n.qhi[a][c], s[a], p[a]);
                // a child box must also lie inside what its ancestors promised (else a ray clipped by the ancestor could miss it)
                if (lo[a] < it.lo[a] - 0.0f || hi[a] > it.hi[a] + 0.0f) { /* allowed: quantisation grids differ; only triangles matter */ }
            }
            if (meta & 0x20u)
            {
                Item ch; ch.node = n.childBase + (meta & 0x1fu); ch.depth = it.depth + 1;
                for (int a = 0; a < 3; a++) { ch.lo[a] = std::max(lo[a], it.lo[a]); ch.hi[a] = std::min(hi[a], it.hi[a]); }
                stack.push_back(ch);
            }
            else
            {
                stats[1]++;
                const uint32_t nt = meta >> 6, first = n.triBase + (meta & 0x1fu);
                for (uint32_t k = 0; k < nt; k++)
                {
                    if (first + k >= numTris) { bad++; continue; }
                    const float* t = leafTris + (size_t)(first + k) * 12;
                    uint32_t g; memcpy(&g, &t[3], 4);
                    if (g >= numTris || seen[g]++) bad++;
                    stats[3]++;
                    for (int v = 0; v < 3; v++)
                        for (int a = 0; a < 3; a++)
                        {
                            const float x = v == 0 ? t[a] : t[a] + t[4 * v + a];
                            // inside this child's box and inside every ancestor's (the intersection carried down)
                            if (x < std::max(lo[a], it.lo[a]) || x > std::min(hi[a], it.hi[a])) bad++;
                        }
                }
            }
        }
    }
    for (uint32_t i = 0; i < numTris; i++) if (seen[i] != 1) bad++;
    return bad;
}

// node visits and triangle tests of every ray (closest hit, or any hit when anyhit != 0), single-threaded
extern "C" int hostsim_trace_stats(const void* nodes, const float* leafTris, const uint32_t* triMesh, const uint32_t* meshFirstTri,
    const float* rays, uint32_t n, uint32_t* outNodes, uint32_t* outTris, int anyhit)
{
    zr::SceneDev sc{};
    sc.nodes = reinterpret_cast<const uint4*>(nodes); sc.tris = reinterpret_cast<const float4*>(leafTris);
    sc.triMesh = triMesh; sc.meshFirstTri = meshFirstTri;
    for (uint32_t i = 0; i < n; i++)
    {
        const float* r = rays + (size_t)i * 8;
        const float3 o = make_float3(r[0], r[1], r[2]), d = make_float3(r[4], r[5], r[6]);
        g_stats.nodes = g_stats.tris = 0;
        if (anyhit) zr::TraceAnyExcept(sc, o, d, r[3], r[7], 0xffffffffu); else zr::TraceClosest(sc, o, d, r[3], r[7]);
        outNodes[i] = (uint32_t)g_stats.nodes; outTris[i] = (uint32_t)g_stats.tris;
    }
    return 0;
}
