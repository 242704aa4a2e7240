"""Times the frame on the procedural stand-ins of BASELINE.json's configs C4 / C5 at their full sizes, one GPU, steady state.
Informational: bench.py's headline stays the metric's own configuration (Cornell 1080p ReSTIR PT); these scenes have
3 x 10^5 / 10^6 triangles and thousands of lights, so their frames are bound by BVH traversal instead of shading.

    python tools/bench_scenes.py atrium      # C4: 2560x1440, ReSTIR GI + light voxel grid + presampled sets, ReSTIR DI
    python tools/bench_scenes.py tunnel      # C5: 3840x2160, ReSTIR PT 5 bounces, 2 spatial passes, ReSTIR DI
    python tools/bench_scenes.py cornell     # the headline workload through the same code, for comparison

Under torchrun (WORLD_SIZE > 1) the frame is strip-sharded across the ranks through the native renderer (zr_renderer_set_shard + zr_comm,
uniform strips), timed as the maximum over ranks, and rank 0 prints the line:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 tools/bench_scenes.py atrium 10

Prints one JSON line: whole-frame ms (CUDA events, two streams), Mpaths/s, per-kernel ms (single-stream profiling pass),
scene / BVH statistics and the host decisions the renderer made (presampling, LVG)."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from zetaray_b200 import lib, check, procedural  # noqa: E402
from zetaray_b200.passes import Scene, Renderer, Comm  # noqa: E402
from zetaray_b200.camera import FrameSequence  # noqa: E402
from zetaray_b200.scene import FlatScene  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIGS = {
    "cornell": dict(res=(1920, 1080), integrator="pt", lvg=False, pt=dict(), cam=None,
                    what="Cornell Box 1080p, ReSTIR PT 3 bounces + ReSTIR DI + firefly/TAA"),
    "atrium": dict(res=(2560, 1440), integrator="gi", lvg=True, pt=dict(), cam=procedural.ATRIUM_CAMERA,
                   what="C4 stand-in: procedural atrium 1440p, ReSTIR GI (LVG NEE variant) + ReSTIR DI (presampled sets) + firefly/TAA"),
    "tunnel": dict(res=(3840, 2160), integrator="pt", lvg=False, pt=dict(max_non_tr_bounces=5, max_glossy_tr_bounces=5, num_spatial_passes=2),
                   cam=procedural.TUNNEL_CAMERA,
                   what="C5 stand-in: procedural station tunnel 4K, ReSTIR PT 5 bounces, 2 spatial passes + ReSTIR DI + firefly/TAA"),
}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "atrium"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    detail = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    cfg = CONFIGS[name]
    w, h = cfg["res"]
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t0 = time.perf_counter()
    if name == "cornell":
        flat = FlatScene.load(os.path.join(ROOT, "tests", "golden", "cornell_emissive.npz"))
    else:
        flat = procedural.SCENES[name][0](detail)
    t_gen = time.perf_counter() - t0
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = C.c_void_p(stream.cuda_stream)
    t0 = time.perf_counter()
    sc = Scene(flat)
    t_scene = time.perf_counter() - t0
    cam = cfg["cam"]
    out = {"workload": cfg["what"], "resolution": [w, h], "triangles": flat.num_triangles, "instances": len(flat.instances),
           "materials": len(flat.materials), "emissive_triangles": len(flat.emissives), "bvh": sc.bvh_stats(),
           "scene_generation_s": round(t_gen, 2), "scene_upload_and_bvh_build_s": round(t_scene, 2)}

    def make_renderer(two_streams):
        r = Renderer(sc, w, h, two_streams=two_streams)
        pres, lvg = r.ApplySceneSettings(use_lvg=cfg["lvg"])
        if cfg["integrator"] == "gi":
            r.SetMethod(Renderer.RESTIR_GI)
        if cfg["pt"]:
            r.indirect.SetParams(**cfg["pt"])
        if os.environ.get("ZR_DENOISE"):            # SVGF between Compositing and TAA; ZR_DENOISE=<radius> (1: 3x3 taps, 2: 5x5 taps)
            r.SetDenoiser(True)
            r.svgf.SetParams(radius=int(os.environ["ZR_DENOISE"]))
        return r, pres, lvg

    r, pres, lvg = make_renderer(True)
    out["presampled_sets"] = pres; out["light_voxel_grid"] = lvg
    seq = FrameSequence(w, h, cam_path=(lambda f: cam) if cam else None)
    for _ in range(5):
        r.Render(seq.next(), st)
    torch.cuda.synchronize()
    if world > 1:
        # every rank holds a complete history now; cut the frame into uniform strips and bring the sharded frame to steady state
        from zetaray_b200.sharding import StripPlan
        comm = Comm.from_torch()
        plan = StripPlan.uniform(h, world)
        r.SetShard(comm, plan.bounds, gather_output=True)
        for _ in range(4):
            r.Render(seq.next(), st)
        torch.cuda.synchronize()
        fcs = [seq.next() for _ in range(steps)]
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for fc in fcs:
            r.Render(fc, st)
        e1.record(stream)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        check(lib.zr_profile_enable(1))
        n = 5
        for _ in range(n):
            r.Render(seq.next(), st)
        torch.cuda.synchronize()
        buf = C.create_string_buffer(8192)
        check(lib.zr_profile_collect(buf, 8192)); check(lib.zr_profile_enable(0))
        mine = {k: round(float(tt) / n, 4) for k, c, tt in (x.split(":") for x in buf.value.decode().split(";") if x)}
        sums = [None] * world
        dist.all_gather_object(sums, round(sum(mine.values()), 3))
        sent, calls = comm.stats()
        if rank == 0:
            out.update({"n_gpus": world, "strips": plan.bounds, "ms_per_frame": round(ms, 3), "mpaths_per_s": round(w * h / (ms * 1e-3) / 1e6, 2),
                        "kernels_ms_per_frame": dict(sorted(mine.items(), key=lambda kv: -kv[1])), "kernel_ms_per_frame_by_rank": sums,
                        "halo_bytes_per_exchange_per_rank": sent // max(1, calls),
                        "parallelism": "1 frame / %d uniform horizontal strips, 32-row halos by grouped NCCL send/recv from C++, image gathered on rank 0" % world})
            print(json.dumps(out))
        dist.destroy_process_group()
        return
    l0 = lib.zr_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fcs = [seq.next() for _ in range(steps)]
    e0.record(stream)
    for fc in fcs:
        r.Render(fc, st)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out["ms_per_frame"] = round(ms, 3)
    out["mpaths_per_s"] = round(w * h / (ms * 1e-3) / 1e6, 2)
    out["gpu_launches_per_frame"] = (lib.zr_kernel_launch_count() - l0) / steps
    r.close()
    # per-kernel times: single stream, so a kernel's events do not include waiting for the other stream
    r, _, _ = make_renderer(False)
    for _ in range(4):
        r.Render(seq.next(), st)
    torch.cuda.synchronize()
    check(lib.zr_profile_enable(1))
    n = 5
    for _ in range(n):
        r.Render(seq.next(), st)
    buf = C.create_string_buffer(8192)
    check(lib.zr_profile_collect(buf, 8192)); check(lib.zr_profile_enable(0))
    out["kernels_ms_per_frame"] = {k: round(float(t) / n, 4) for k, c, t in (x.split(":") for x in buf.value.decode().split(";") if x)}
    out["kernels_ms_per_frame"] = dict(sorted(out["kernels_ms_per_frame"].items(), key=lambda kv: -kv[1]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
