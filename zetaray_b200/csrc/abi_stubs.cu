// abi_stubs.cu -- entry points declared in include/zr_abi.h that are not implemented yet.
// They fail loudly (ZR_ERR_UNSUPPORTED + message); nothing here computes anything.
#include "zr_common.cuh"

extern "C"
{
    zr_status zr_scene_create(const zr_scene_desc* desc, zr_scene** out) { zr::set_error("zr_scene_create: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    void zr_scene_destroy(zr_scene* scene) {}
    zr_status zr_scene_bvh_stats(const zr_scene* scene, uint32_t out[4]) { zr::set_error("zr_scene_bvh_stats: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_scene_trace_closest(const zr_scene* scene, const float* d_rays, uint32_t n, float* d_hits, void* stream) { zr::set_error("zr_scene_trace_closest: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_scene_trace_any(const zr_scene* scene, const float* d_rays, uint32_t n, uint32_t* d_hit_flags, void* stream) { zr::set_error("zr_scene_trace_any: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_estimate_emissive_power(const zr_scene* scene, float* d_power, void* stream) { zr::set_error("zr_estimate_emissive_power: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_prelighting_render(zr_scene* scene, void* stream) { zr::set_error("zr_prelighting_render: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_scene_get_alias_table(const zr_scene* scene, const zr_alias_entry** d_table, uint32_t* n) { zr::set_error("zr_scene_get_alias_table: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_gbuffer_alloc(uint32_t width, uint32_t height, int with_tridiff, zr_gbuffer* out) { zr::set_error("zr_gbuffer_alloc: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    void zr_gbuffer_free(zr_gbuffer* g) {}
    zr_status zr_gbuffer_pass_create(zr_gbuffer_pass** out) { zr::set_error("zr_gbuffer_pass_create: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_gbuffer_pass_render(zr_gbuffer_pass* p, const zr_frame_inputs* in, void* stream) { zr::set_error("zr_gbuffer_pass_render: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_gbuffer_pass_describe_io(zr_gbuffer_pass* p, zr_resource_use* uses, int* n) { zr::set_error("zr_gbuffer_pass_describe_io: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    void zr_gbuffer_pass_destroy(zr_gbuffer_pass* p) {}
    zr_status zr_direct_pass_create(uint32_t width, uint32_t height, zr_direct_pass** out) { zr::set_error("zr_direct_pass_create: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_direct_pass_resize(zr_direct_pass* p, uint32_t width, uint32_t height) { zr::set_error("zr_direct_pass_resize: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_direct_pass_reset_temporal(zr_direct_pass* p) { zr::set_error("zr_direct_pass_reset_temporal: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_direct_pass_default_params(zr_direct_params* out) { zr::set_error("zr_direct_pass_default_params: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_direct_pass_set_params(zr_direct_pass* p, const zr_direct_params* params) { zr::set_error("zr_direct_pass_set_params: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_direct_pass_render(zr_direct_pass* p, const zr_frame_inputs* in, void* stream) { zr::set_error("zr_direct_pass_render: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_direct_pass_get_output(zr_direct_pass* p, zr_direct_output id, zr_image2d* out) { zr::set_error("zr_direct_pass_get_output: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_direct_pass_describe_io(zr_direct_pass* p, zr_resource_use* uses, int* n) { zr::set_error("zr_direct_pass_describe_io: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    void zr_direct_pass_destroy(zr_direct_pass* p) {}
    zr_status zr_indirect_pass_create(uint32_t width, uint32_t height, zr_indirect_pass** out) { zr::set_error("zr_indirect_pass_create: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_indirect_pass_resize(zr_indirect_pass* p, uint32_t width, uint32_t height) { zr::set_error("zr_indirect_pass_resize: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_indirect_pass_reset_temporal(zr_indirect_pass* p) { zr::set_error("zr_indirect_pass_reset_temporal: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_indirect_pass_default_params(zr_indirect_params* out) { zr::set_error("zr_indirect_pass_default_params: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_indirect_pass_set_params(zr_indirect_pass* p, const zr_indirect_params* params) { zr::set_error("zr_indirect_pass_set_params: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_indirect_pass_render(zr_indirect_pass* p, const zr_frame_inputs* in, void* stream) { zr::set_error("zr_indirect_pass_render: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_indirect_pass_render_until(zr_indirect_pass* p, const zr_frame_inputs* in, zr_indirect_stage last_stage, void* stream) { zr::set_error("zr_indirect_pass_render_until: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_indirect_pass_get_output(zr_indirect_pass* p, zr_indirect_output id, zr_image2d* out) { zr::set_error("zr_indirect_pass_get_output: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_indirect_pass_describe_io(zr_indirect_pass* p, zr_resource_use* uses, int* n) { zr::set_error("zr_indirect_pass_describe_io: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_indirect_pass_set_rows(zr_indirect_pass* p, uint32_t y0, uint32_t y1) { zr::set_error("zr_indirect_pass_set_rows: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    void zr_indirect_pass_destroy(zr_indirect_pass* p) {}
    zr_status zr_compositing_pass_create(uint32_t width, uint32_t height, zr_compositing_pass** out) { zr::set_error("zr_compositing_pass_create: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_compositing_pass_resize(zr_compositing_pass* p, uint32_t width, uint32_t height) { zr::set_error("zr_compositing_pass_resize: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_compositing_pass_set_params(zr_compositing_pass* p, const zr_compositing_params* params) { zr::set_error("zr_compositing_pass_set_params: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_compositing_pass_render(zr_compositing_pass* p, const zr_frame_inputs* in, const void* d_direct, const void* d_indirect, void* stream) { zr::set_error("zr_compositing_pass_render: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_compositing_pass_get_output(zr_compositing_pass* p, zr_image2d* out) { zr::set_error("zr_compositing_pass_get_output: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    void zr_compositing_pass_destroy(zr_compositing_pass* p) {}
    zr_status zr_taa_pass_create(uint32_t width, uint32_t height, zr_taa_pass** out) { zr::set_error("zr_taa_pass_create: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_taa_pass_resize(zr_taa_pass* p, uint32_t width, uint32_t height) { zr::set_error("zr_taa_pass_resize: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_taa_pass_set_blend_weight(zr_taa_pass* p, float w) { zr::set_error("zr_taa_pass_set_blend_weight: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_taa_pass_render(zr_taa_pass* p, const zr_frame_inputs* in, const void* d_signal, void* stream) { zr::set_error("zr_taa_pass_render: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    zr_status zr_taa_pass_get_output(zr_taa_pass* p, zr_image2d* out) { zr::set_error("zr_taa_pass_get_output: not implemented in this build"); return ZR_ERR_UNSUPPORTED; }
    void zr_taa_pass_destroy(zr_taa_pass* p) {}
}
