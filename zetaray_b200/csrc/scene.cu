// scene.cu -- scene upload, host BVH build and the ray-query test entry points.
//
// Replaces what ZetaCore/RayTracing/RtAccelerationStructure.cpp gets from the DXR driver (BLAS/TLAS
// build) with an own builder: binned-SAH binary BVH over world-space triangles -> collapsed to
// 8-wide nodes -> child boxes quantised to 8 bits (conservatively rounded outwards).
// World-space triangles are produced on the device with the same TransformTRS arithmetic the
// shading code uses (quantised rotation / half scale of RT::MeshInstance), so traversal geometry and
// shading geometry agree bit for bit.
#include "zr_scene.cuh"
#include <vector>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <dlfcn.h>

namespace zr
{
namespace
{
    __global__ void k_world_tris(SceneDev sc, const uint32_t* __restrict__ triMesh, const uint32_t* __restrict__ meshFirstTri,
        uint32_t numTris, float* __restrict__ out /* 9 floats per tri */)
    {
        const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
        if (g >= numTris) return;
        const uint32_t m = triMesh[g];
        const uint32_t p = g - meshFirstTri[m];
        const zr_mesh_instance md = LoadInstance(sc, m);
        const float4 q = normalize(Math::DecodeNormalized4(md.Rotation));
        const float3 s = h3(md.Scale);
        const float3 t = f3(md.Translation[0], md.Translation[1], md.Translation[2]);
        const uint32_t tri = p * 3 + md.BaseIdxOffset;
        float3 pw[3];
        for (int k = 0; k < 3; k++)
        {
            const VertexD V = LoadVertex(sc, sc.indices[tri + k] + md.BaseVtxOffset);
            pw[k] = Math::TransformTRS(V.pos, t, q, s);
        }
        const float3 e1 = pw[1] - pw[0], e2 = pw[2] - pw[0];
        float* o = out + (size_t)g * 9;
        o[0] = pw[0].x; o[1] = pw[0].y; o[2] = pw[0].z;
        o[3] = e1.x; o[4] = e1.y; o[5] = e1.z;
        o[6] = e2.x; o[7] = e2.y; o[8] = e2.z;
    }

    __global__ void k_trace_closest(SceneDev sc, const float* __restrict__ rays, uint32_t n, float* __restrict__ hits)
    {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= n) return;
        const float* r = rays + (size_t)i * 8;
        RayHit h = TraceClosest(sc, f3(r[0], r[1], r[2]), f3(r[4], r[5], r[6]), r[3], r[7]);
        float* o = hits + (size_t)i * 4;
        o[0] = h.hit ? h.t : FLT_MAX_; o[1] = h.bary.x; o[2] = h.bary.y; o[3] = asfloat(h.tri);
    }

    __global__ void k_trace_any(SceneDev sc, const float* __restrict__ rays, uint32_t n, uint32_t* __restrict__ flags)
    {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= n) return;
        const float* r = rays + (size_t)i * 8;
        flags[i] = TraceAnyExcept(sc, f3(r[0], r[1], r[2]), f3(r[4], r[5], r[6]), r[3], r[7], 0xffffffffu) ? 1u : 0u;
    }

    // ------------------------------------------------------------------------------------------
    // host BVH builder
    // ------------------------------------------------------------------------------------------
    struct AABB
    {
        float lo[3], hi[3];
        void reset() { for (int a = 0; a < 3; a++) { lo[a] = INFINITY; hi[a] = -INFINITY; } }
        void grow(const AABB& b) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
        void grow(const float p[3]) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
        float area() const
        {
            float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
            if (dx < 0) return 0;
            return 2.0f * (dx * dy + dy * dz + dz * dx);
        }
    };

    struct BNode { AABB box; int left = -1, right = -1; uint32_t first = 0, count = 0; };

    struct Builder
    {
        std::vector<AABB> triBox;
        std::vector<float> centroid;    // 3 per tri
        std::vector<uint32_t> order;
        std::vector<BNode> nodes;

        int build(uint32_t first, uint32_t count)
        {
            BNode n;
            n.box.reset();
            AABB cb; cb.reset();
            for (uint32_t i = first; i < first + count; i++)
            {
                n.box.grow(triBox[order[i]]);
                cb.grow(&centroid[order[i] * 3]);
            }
            n.first = first; n.count = count;
            const int idx = (int)nodes.size();
            nodes.push_back(n);
            if (count <= 3)
                return idx;
            // binned SAH over the widest centroid axis (try all 3)
            const int NB = 16;
            float bestCost = INFINITY; int bestAxis = -1; int bestSplit = -1;
            for (int a = 0; a < 3; a++)
            {
                const float ext = cb.hi[a] - cb.lo[a];
                if (!(ext > 0)) continue;
                AABB bb[NB]; uint32_t bc[NB];
                for (int b = 0; b < NB; b++) { bb[b].reset(); bc[b] = 0; }
                for (uint32_t i = first; i < first + count; i++)
                {
                    int b = (int)((centroid[order[i] * 3 + a] - cb.lo[a]) / ext * NB);
                    b = std::min(std::max(b, 0), NB - 1);
                    bb[b].grow(triBox[order[i]]); bc[b]++;
                }
                AABB r; r.reset();
                float rArea[NB]; uint32_t rCnt[NB]; uint32_t c = 0;
                for (int b = NB - 1; b > 0; b--) { r.grow(bb[b]); c += bc[b]; rArea[b] = r.area(); rCnt[b] = c; }
                AABB l; l.reset(); c = 0;
                for (int b = 0; b < NB - 1; b++)
                {
                    l.grow(bb[b]); c += bc[b];
                    if (c == 0 || rCnt[b + 1] == 0) continue;
                    const float cost = l.area() * (float)c + rArea[b + 1] * (float)rCnt[b + 1];
                    if (cost < bestCost) { bestCost = cost; bestAxis = a; bestSplit = b; }
                }
            }
            uint32_t mid;
            if (bestAxis < 0)
                mid = first + count / 2;        // all centroids coincide: split by index
            else
            {
                const float ext = cb.hi[bestAxis] - cb.lo[bestAxis];
                auto it = std::partition(order.begin() + first, order.begin() + first + count, [&](uint32_t t) {
                    int b = (int)((centroid[t * 3 + bestAxis] - cb.lo[bestAxis]) / ext * NB);
                    b = std::min(std::max(b, 0), NB - 1);
                    return b <= bestSplit;
                });
                mid = (uint32_t)(it - order.begin());
                if (mid == first || mid == first + count)
                    mid = first + count / 2;
            }
            const int l = build(first, mid - first);
            const int r = build(mid, first + count - mid);
            nodes[idx].left = l; nodes[idx].right = r;
            return idx;
        }
    };

    struct WideOut
    {
        std::vector<BVH8Node> nodes;
        std::vector<uint32_t> leafOrder;     // global tri index per slot in leaf order
        uint32_t maxDepth = 0;
    };

    void emit_wide(const Builder& b, int binIdx, uint32_t outIdx, WideOut& out, uint32_t depth)
    {
        out.maxDepth = std::max(out.maxDepth, depth);
        // gather up to 8 children by repeatedly opening the child with the largest area
        std::vector<int> kids;
        const BNode& root = b.nodes[binIdx];
        if (root.left < 0) kids.push_back(binIdx);
        else { kids.push_back(root.left); kids.push_back(root.right); }
        while (kids.size() < 8)
        {
            int bestK = -1; float bestA = -1;
            for (size_t k = 0; k < kids.size(); k++)
            {
                const BNode& c = b.nodes[kids[k]];
                if (c.left < 0) continue;
                const float a = c.box.area();
                if (a > bestA) { bestA = a; bestK = (int)k; }
            }
            if (bestK < 0) break;
            const BNode c = b.nodes[kids[bestK]];
            kids[bestK] = c.left;
            kids.push_back(c.right);
        }
        BVH8Node n;
        memset(&n, 0, sizeof(n));
        const AABB& box = root.box;
        n.px = box.lo[0]; n.py = box.lo[1]; n.pz = box.lo[2];
        uint8_t* ex[3] = { &n.ex, &n.ey, &n.ez };
        float scale[3];
        for (int a = 0; a < 3; a++)
        {
            const float ext = std::max(box.hi[a] - box.lo[a], 1e-30f);
            int e = (int)std::ceil(std::log2(ext / 255.0f));
            // make sure 255 * 2^e covers the extent even after rounding
            while (std::ldexp(255.0f, e) < ext) e++;
            e = std::min(std::max(e, -126), 127);
            *ex[a] = (uint8_t)(e + 127);
            scale[a] = std::ldexp(1.0f, e);
        }
        // internal children first get contiguous node slots
        std::vector<int> internalKids, leafKids;
        for (int k : kids) (b.nodes[k].left < 0 ? leafKids : internalKids).push_back(k);
        n.childBase = (uint32_t)out.nodes.size();
        n.triBase = (uint32_t)out.leafOrder.size();
        const uint32_t childBase = n.childBase;
        out.nodes.resize(out.nodes.size() + internalKids.size());
        int slot = 0;
        uint32_t triOff = 0, intOff = 0;
        std::vector<std::pair<int, uint32_t>> recurse;
        auto quant = [&](const AABB& cb, int c) {
            const float org[3] = { n.px, n.py, n.pz };
            for (int a = 0; a < 3; a++)
            {
                float lo = std::floor((cb.lo[a] - org[a]) / scale[a]);
                float hi = std::ceil((cb.hi[a] - org[a]) / scale[a]);
                // guard against rounding of the division itself
                while (lo > 0 && org[a] + lo * scale[a] > cb.lo[a]) lo -= 1;
                while (hi < 255 && org[a] + hi * scale[a] < cb.hi[a]) hi += 1;
                lo = std::min(std::max(lo, 0.0f), 255.0f);
                hi = std::min(std::max(hi, 0.0f), 255.0f);
                n.qlo[a][c] = (uint8_t)lo;
                n.qhi[a][c] = (uint8_t)hi;
            }
        };
        for (int k : internalKids)
        {
            n.meta[slot] = (uint8_t)(0x20u | intOff);
            quant(b.nodes[k].box, slot);
            recurse.push_back({ k, childBase + intOff });
            intOff++; slot++;
        }
        for (int k : leafKids)
        {
            const BNode& c = b.nodes[k];
            n.meta[slot] = (uint8_t)((c.count << 6) | triOff);
            quant(c.box, slot);
            for (uint32_t i = 0; i < c.count; i++)
                out.leafOrder.push_back(b.order[c.first + i]);
            triOff += c.count; slot++;
        }
        out.nodes[outIdx] = n;
        for (auto& r : recurse)
            emit_wide(b, r.first, r.second, out, depth + 1);
    }

    std::string asset_path(const char* name)
    {
        Dl_info info;
        std::string dir = ".";
        if (dladdr((void*)&asset_path, &info) && info.dli_fname)
        {
            std::string p = info.dli_fname;
            size_t s = p.find_last_of('/');
            if (s != std::string::npos) dir = p.substr(0, s);
        }
        return dir + "/assets/" + name;
    }
}

template<typename T>
static zr_status upload(zr_scene* sc, const T* h, size_t n, const T** d)
{
    void* p = nullptr;
    ZR_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    if (n) ZR_CUDA(cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice));
    sc->allocs[sc->numAllocs++] = p;
    *d = (const T*)p;
    return ZR_OK;
}

zr_status scene_create(const zr_scene_desc* desc, zr_scene** out)
{
    if (!desc || !out || !desc->h_vertices || !desc->h_indices || !desc->h_instances || !desc->h_instance_num_tris ||
        !desc->h_materials || desc->num_instances == 0)
    {
        set_error("zr_scene_create: null input");
        return ZR_ERR_INVALID_ARG;
    }
    zr_scene* sc = new zr_scene();
    zr_status st;
#define UP(field, ptr, n) if ((st = upload(sc, ptr, n, &sc->dev.field)) != ZR_OK) { zr_scene_destroy(sc); return st; }
    UP(vertices, desc->h_vertices, desc->num_vertices);
    UP(indices, desc->h_indices, desc->num_indices);
    UP(instances, desc->h_instances, desc->num_instances);
    UP(materials, desc->h_materials, desc->num_materials);
    UP(emissives, desc->h_emissives, desc->num_emissives);
    sc->dev.numInstances = desc->num_instances;
    sc->dev.numEmissives = desc->num_emissives;

    // triangle -> mesh maps
    std::vector<uint32_t> triMesh, meshFirst(desc->num_instances);
    uint32_t total = 0;
    for (uint32_t m = 0; m < desc->num_instances; m++)
    {
        meshFirst[m] = total;
        if (desc->h_instances[m].BaseIdxOffset + 3 * desc->h_instance_num_tris[m] > desc->num_indices)
        {
            set_error("zr_scene_create: instance %u indexes past the index buffer", m);
            zr_scene_destroy(sc);
            return ZR_ERR_INVALID_ARG;
        }
        for (uint32_t p = 0; p < desc->h_instance_num_tris[m]; p++) triMesh.push_back(m);
        total += desc->h_instance_num_tris[m];
    }
    if (total == 0) { set_error("zr_scene_create: scene has no triangles"); zr_scene_destroy(sc); return ZR_ERR_INVALID_ARG; }
    sc->dev.numTris = total;
    UP(triMesh, triMesh.data(), triMesh.size());
    UP(meshFirstTri, meshFirst.data(), meshFirst.size());

    // directional-albedo table
    {
        std::vector<uint16_t> rho(64 * 32 * 16);
        const std::string path = asset_path("rho_lut.bin");
        FILE* f = fopen(path.c_str(), "rb");
        if (!f || fread(rho.data(), 2, rho.size(), f) != rho.size())
        {
            if (f) fclose(f);
            set_error("zr_scene_create: cannot read %s (run tools/gen_rho_lut.py)", path.c_str());
            zr_scene_destroy(sc);
            return ZR_ERR_NOT_INITIALIZED;
        }
        fclose(f);
        UP(rho, rho.data(), rho.size());
    }

    // world-space triangles on the device, then BVH on the host
    float* d_wt = nullptr;
    ZR_CUDA(cudaMalloc(&d_wt, (size_t)total * 9 * sizeof(float)));
    k_world_tris<<<(total + 127) / 128, 128>>>(sc->dev, sc->dev.triMesh, sc->dev.meshFirstTri, total, d_wt);
    count_launch();
    std::vector<float> wt((size_t)total * 9);
    cudaError_t e = cudaMemcpy(wt.data(), d_wt, wt.size() * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_wt);
    if (e != cudaSuccess) { zr_scene_destroy(sc); return cuda_fail(e, "world triangles"); }

    Builder b;
    b.triBox.resize(total); b.centroid.resize((size_t)total * 3); b.order.resize(total);
    for (uint32_t i = 0; i < total; i++)
    {
        const float* t = &wt[(size_t)i * 9];
        float p[3][3];
        for (int a = 0; a < 3; a++) { p[0][a] = t[a]; p[1][a] = t[a] + t[3 + a]; p[2][a] = t[a] + t[6 + a]; }
        AABB bx; bx.reset();
        for (int k = 0; k < 3; k++) bx.grow(p[k]);
        for (int a = 0; a < 3; a++)
        {
            const float pad = 4e-7f * std::max(std::max(std::fabs(bx.lo[a]), std::fabs(bx.hi[a])), 1.0f);
            bx.lo[a] -= pad; bx.hi[a] += pad;
            b.centroid[(size_t)i * 3 + a] = 0.5f * (bx.lo[a] + bx.hi[a]);
        }
        b.triBox[i] = bx;
        b.order[i] = i;
    }
    b.nodes.reserve((size_t)total * 2);
    b.build(0, total);
    WideOut w;
    w.nodes.resize(1);
    emit_wide(b, 0, 0, w, 1);
    std::vector<float4> tris((size_t)total * 3);
    for (uint32_t s = 0; s < total; s++)
    {
        const uint32_t g = w.leafOrder[s];
        const float* t = &wt[(size_t)g * 9];
        uint32_t gb = g; float gf; memcpy(&gf, &gb, 4);
        tris[(size_t)s * 3 + 0] = make_float4(t[0], t[1], t[2], gf);
        tris[(size_t)s * 3 + 1] = make_float4(t[3], t[4], t[5], 0.0f);
        tris[(size_t)s * 3 + 2] = make_float4(t[6], t[7], t[8], 0.0f);
    }
    {
        const uint4* d_nodes = nullptr;
        if ((st = upload(sc, reinterpret_cast<const uint4*>(w.nodes.data()), w.nodes.size() * 5, &d_nodes)) != ZR_OK) { zr_scene_destroy(sc); return st; }
        sc->dev.nodes = d_nodes;
    }
    UP(tris, tris.data(), tris.size());
#undef UP
    sc->info.numNodes = (uint32_t)w.nodes.size();
    sc->info.numTris = total;
    sc->info.maxDepth = w.maxDepth;
    sc->info.bytes = (uint32_t)(w.nodes.size() * sizeof(BVH8Node) + tris.size() * sizeof(float4));

    // alias table storage (built by zr_prelighting_render)
    if (desc->num_emissives)
    {
        cudaMalloc(&sc->d_alias, (size_t)desc->num_emissives * sizeof(zr_alias_entry));
        cudaMalloc(&sc->d_power, (size_t)(desc->num_emissives + 8) * sizeof(float));
        cudaMalloc(&sc->d_aliasScratch, (size_t)desc->num_emissives * 2 * sizeof(uint32_t));
        sc->dev.aliasTable = sc->d_alias;
    }
    *out = sc;
    return ZR_OK;
}
} // namespace zr

extern "C"
{
    zr_status zr_scene_create(const zr_scene_desc* desc, zr_scene** out) { return zr::scene_create(desc, out); }
    void zr_scene_destroy(zr_scene* sc)
    {
        if (!sc) return;
        for (int i = 0; i < sc->numAllocs; i++) cudaFree(sc->allocs[i]);
        if (sc->d_alias) cudaFree(sc->d_alias);
        if (sc->d_power) cudaFree(sc->d_power);
        if (sc->d_aliasScratch) cudaFree(sc->d_aliasScratch);
        if (sc->d_sampleSets) cudaFree(sc->d_sampleSets);
        if (sc->d_lvg) cudaFree(sc->d_lvg);
        delete sc;
    }
    zr_status zr_scene_bvh_stats(const zr_scene* sc, uint32_t out[4])
    {
        if (!sc || !out) return ZR_ERR_INVALID_ARG;
        out[0] = sc->info.numNodes; out[1] = sc->info.numTris; out[2] = sc->info.maxDepth; out[3] = sc->info.bytes;
        return ZR_OK;
    }
    zr_status zr_scene_trace_closest(const zr_scene* sc, const float* d_rays, uint32_t n, float* d_hits, void* stream)
    {
        if (!sc || !d_rays || !d_hits) { zr::set_error("zr_scene_trace_closest: null argument"); return ZR_ERR_INVALID_ARG; }
        if (n == 0) return ZR_OK;
        ZR_PROF("k_trace_closest", (cudaStream_t)stream);
        zr::k_trace_closest<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sc->dev, d_rays, n, d_hits);
        ZR_LAUNCH_CHECK();
        return ZR_OK;
    }
    zr_status zr_scene_trace_any(const zr_scene* sc, const float* d_rays, uint32_t n, uint32_t* d_flags, void* stream)
    {
        if (!sc || !d_rays || !d_flags) { zr::set_error("zr_scene_trace_any: null argument"); return ZR_ERR_INVALID_ARG; }
        if (n == 0) return ZR_OK;
        ZR_PROF("k_trace_any", (cudaStream_t)stream);
        zr::k_trace_any<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sc->dev, d_rays, n, d_flags);
        ZR_LAUNCH_CHECK();
        return ZR_OK;
    }
    zr_status zr_scene_get_alias_table(const zr_scene* sc, const zr_alias_entry** d_table, uint32_t* n)
    {
        if (!sc || !d_table || !n) return ZR_ERR_INVALID_ARG;
        *d_table = sc->d_alias; *n = sc->dev.numEmissives;
        return ZR_OK;
    }
}
