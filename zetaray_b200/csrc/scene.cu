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
    sc->allocs[sc->numAllocs++] = p;        // owned by the scene from here on: released by zr_scene_destroy on every error path
    if (n) ZR_CUDA(cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice));
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
    // Everything the kernels will index is checked here, on the host, in 64-bit arithmetic, before anything is allocated: a malformed
    // description is an error code, not an out-of-bounds device read in k_world_tris and every lighting kernel.
    {
        uint64_t totalTris = 0;
        for (uint32_t m = 0; m < desc->num_instances; m++)
        {
            const zr_mesh_instance& mi = desc->h_instances[m];
            const uint64_t nt = desc->h_instance_num_tris[m];
            if ((uint64_t)mi.BaseIdxOffset + 3 * nt > desc->num_indices)
            {
                set_error("zr_scene_create: instance %u indexes past the index buffer", m);
                return ZR_ERR_INVALID_ARG;
            }
            if (mi.MatIdx >= desc->num_materials)
            {
                set_error("zr_scene_create: instance %u uses material %u of %u", m, (unsigned)mi.MatIdx, desc->num_materials);
                return ZR_ERR_INVALID_ARG;
            }
            for (uint64_t i = 0; i < 3 * nt; i++)
                if ((uint64_t)mi.BaseVtxOffset + desc->h_indices[mi.BaseIdxOffset + i] >= desc->num_vertices)
                {
                    set_error("zr_scene_create: instance %u, index %llu points past the vertex buffer", m, (unsigned long long)i);
                    return ZR_ERR_INVALID_ARG;
                }
            if (mi.BaseEmissiveTriOffset != 0xffffffffu && (uint64_t)mi.BaseEmissiveTriOffset + nt > desc->num_emissives)
            {
                set_error("zr_scene_create: instance %u's emissive triangles [%u, %llu) lie past the %u emissive triangles", m,
                    mi.BaseEmissiveTriOffset, (unsigned long long)(mi.BaseEmissiveTriOffset + nt), desc->num_emissives);
                return ZR_ERR_INVALID_ARG;
            }
            totalTris += nt;
        }
        if (totalTris == 0) { set_error("zr_scene_create: scene has no triangles"); return ZR_ERR_INVALID_ARG; }
        if (totalTris > 0x7fffffffull) { set_error("zr_scene_create: more than 2^31 triangles"); return ZR_ERR_INVALID_ARG; }
        if (desc->num_emissives && !desc->h_emissives) { set_error("zr_scene_create: null emissive buffer"); return ZR_ERR_INVALID_ARG; }
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
            set_error("zr_scene_create: cannot read %s (tools/extract_reference_tables.py writes it)", path.c_str());
            zr_scene_destroy(sc);
            return ZR_ERR_NOT_INITIALIZED;
        }
        fclose(f);
        UP(rho, rho.data(), rho.size());
    }

    // world-space triangles on the device, then BVH on the host
    float* d_wt = nullptr;
    cudaError_t e = cudaMalloc(&d_wt, (size_t)total * 9 * sizeof(float));
    if (e != cudaSuccess) { zr_scene_destroy(sc); return cuda_fail(e, "world triangles (alloc)"); }
    k_world_tris<<<(total + 127) / 128, 128>>>(sc->dev, sc->dev.triMesh, sc->dev.meshFirstTri, total, d_wt);
    count_launch();
    e = cudaGetLastError();
    std::vector<float> wt((size_t)total * 9);
    if (e == cudaSuccess) e = cudaMemcpy(wt.data(), d_wt, wt.size() * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_wt);
    if (e != cudaSuccess) { zr_scene_destroy(sc); return cuda_fail(e, "world triangles"); }

    BvhBuild w;
    build_bvh8(wt.data(), total, w);
    if (w.maxStack > (uint32_t)BVH_STACK_ENTRIES)
    {
        set_error("zr_scene_create: the BVH needs a traversal stack of %u entries, the kernels hold %d", w.maxStack, BVH_STACK_ENTRIES);
        zr_scene_destroy(sc);
        return ZR_ERR_INVALID_ARG;
    }
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
    sc->info.maxStack = w.maxStack;
    sc->info.bytes = (uint32_t)(w.nodes.size() * sizeof(BVH8Node) + tris.size() * sizeof(float4));

    // alias table storage (built by zr_prelighting_render)
    if (desc->num_emissives)
    {
        e = cudaMalloc(&sc->d_alias, (size_t)desc->num_emissives * sizeof(zr_alias_entry));
        if (e == cudaSuccess) e = cudaMalloc(&sc->d_power, (size_t)(desc->num_emissives + 8) * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&sc->d_aliasScratch, ((size_t)desc->num_emissives * 2 + 16) * sizeof(uint32_t));
        if (e != cudaSuccess) { zr_scene_destroy(sc); return cuda_fail(e, "alias table storage"); }
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
    zr_status zr_bvh_build_host(const float* h_world_tris, uint32_t num_tris, void* h_nodes, uint32_t node_capacity,
        uint32_t* h_leaf_order, uint32_t out_info[4])
    {
        if (!h_world_tris || !out_info || num_tris == 0) { zr::set_error("zr_bvh_build_host: null argument"); return ZR_ERR_INVALID_ARG; }
        zr::BvhBuild w;
        zr::build_bvh8(h_world_tris, num_tris, w);
        out_info[0] = (uint32_t)w.nodes.size(); out_info[1] = num_tris; out_info[2] = w.maxDepth; out_info[3] = w.maxStack;
        if (h_nodes)
        {
            if (node_capacity < w.nodes.size()) { zr::set_error("zr_bvh_build_host: %zu nodes, capacity %u", w.nodes.size(), node_capacity); return ZR_ERR_INVALID_ARG; }
            memcpy(h_nodes, w.nodes.data(), w.nodes.size() * sizeof(zr::BVH8Node));
        }
        if (h_leaf_order) memcpy(h_leaf_order, w.leafOrder.data(), (size_t)num_tris * sizeof(uint32_t));
        return ZR_OK;
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
