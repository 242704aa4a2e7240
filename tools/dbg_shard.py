"""Debug aid: run the sharded-vs-unsharded comparison of tests/test_sharded_gpu.py and print, per plane, which rows differ."""
import ctypes as C, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
from zetaray_b200 import _lib
from zetaray_b200.passes import Scene, GBuffers, GBufferRT, DirectLighting, IndirectLighting, Compositing, TAA, download_image
from zetaray_b200.sharding import ShardedFrame, StripPlan
from tests import scene_util, rpt_util
W, H = 416, 296
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = C.c_void_p(stream.cuda_stream)
scene = Scene(scene_util.glossy_cornell()); scene.prelighting(st)
def pipeline():
    passes = dict(gbuffer=GBufferRT(), direct=DirectLighting(W, H), indirect=IndirectLighting(W, H), compositing=Compositing(W, H), taa=TAA(W, H))
    fi = _lib.FrameInputs(); fi.scene = scene.handle
    return ShardedFrame(passes, GBuffers(W, H), W, H, rank, world), fi
A, fiA = pipeline(); B, fiB = pipeline()
seq = rpt_util.FrameSequence(W, H)
B.begin_cost_measurement()
for _ in range(3):
    fc = seq.next(); A.render(fiA, fc, stream, st); B.render(fiB, fc, stream, st)
costs = B.end_cost_measurement()
plan = StripPlan.balanced(H, world, costs); B.shard(plan); B.debug_sync = bool(int(os.environ.get('ZR_DEBUG_SYNC', '0')))
y0, y1 = plan.rows(rank)
planes = (("di_final", lambda s: s.p["direct"].GetOutput(0), np.float32, 4), ("ind_final", lambda s: s.p["indirect"].GetOutput(0), np.float32, 4),
          ("di_res", lambda s: s.p["direct"].GetOutput(1), np.uint32, 8), ("ind_res", lambda s: s.p["indirect"].GetOutput(1), np.uint32, 16),
          ("comp", lambda s: s.p["compositing"].GetOutput(), np.float32, 4), ("taa", lambda s: s.p["taa"].GetOutput(), np.uint16, 4))
from zetaray_b200.sharding import plane_tensor
orig = B.halo.exchange
def checked(tensors):
    torch.cuda.synchronize()
    ref = None
    if len(tensors) == 1 and tensors[0].shape[1] == 32 * W:
        ref = plane_tensor(A.p["direct"].GetOutput(1))
        nm = "DI"
    own_ok = None if ref is None else bool((tensors[0][y0:y1] == ref[y0:y1]).all())
    orig(tensors)
    torch.cuda.synchronize()
    if ref is not None:
        lo, hi = max(0, y0 - 32), min(H, y1 + 32)
        print("rank %d HOOK %s ptr %x (A ptr %x) own rows ok before: %s; halo ok after: %s; calls %d" % (rank, nm, tensors[0].data_ptr(), ref.data_ptr(), own_ok,
              bool((tensors[0][lo:hi] == ref[lo:hi]).all()), B.halo.calls), flush=True)
B.halo.exchange = checked if int(os.environ.get('ZR_DEBUG_SYNC','0')) else orig
for f in range(2):
    fc = seq.next(); A.render(fiA, fc, stream, st); B.render(fiB, fc, stream, st); torch.cuda.synchronize()
    ga, gb_ = A.gb.download("curr"), B.gb.download("curr")
    lo, hi = max(0, y0 - 32), min(H, y1 + 32)
    for nm, x, y in zip(("core", "depth", "me", "coat"), ga, gb_):
        x = x.reshape(H, -1); y = y.reshape(H, -1)
        print("rank %d frame %d gbuffer %-5s rows [%d,%d) equal: %s" % (rank, f, nm, lo, hi, np.array_equal(x[lo:hi].view(np.uint8), y[lo:hi].view(np.uint8))), flush=True)
    for name, get, dt, comps in planes:
        a = download_image(get(A), dt, comps).reshape(H, W, comps); b = download_image(get(B), dt, comps).reshape(H, W, comps)
        d = (a.view(np.uint8).reshape(H, -1) != b.view(np.uint8).reshape(H, -1)).any(axis=1)
        own = d[y0:y1]; rows = np.nonzero(own)[0] + y0
        lo, hi = max(0, y0 - 32), min(H, y1 + 32)
        halo_bad = np.nonzero(np.concatenate([d[lo:y0], d[y1:hi]]))[0].size
        print("rank %d strip [%d,%d) frame %d %-9s own-row mismatches %4d %s | halo rows bad %d" % (rank, y0, y1, f, name, rows.size,
              ("rows %d..%d" % (rows.min(), rows.max())) if rows.size else "", halo_bad), flush=True)
dist.destroy_process_group()
