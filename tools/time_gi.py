"""Times the ReSTIR GI integrator (k_rgi) on the bench workload: Cornell 1080p, steady state. Informational (bench.py's metric
is the ReSTIR PT frame)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zetaray_b200 import lib, check, _lib
from zetaray_b200.passes import Scene, GBuffers, GBufferRT, IndirectLightingGI
from tests import scene_util, rpt_util
W, H = 1920, 1080
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = C.c_void_p(stream.cuda_stream)
sc = Scene(scene_util.cornell()); sc.prelighting(st)
gb, g, gi = GBuffers(W, H), GBufferRT(), IndirectLightingGI(W, H)
seq = rpt_util.FrameSequence(W, H)
fi = _lib.FrameInputs(); fi.scene = sc.handle
def frame():
    gb.flip(); fi.frame = seq.next(); gb.fill_inputs(fi); g.Render(fi, st); gi.Render(fi, st)
for _ in range(5): frame()
check(lib.zr_profile_enable(1))
n = 10
for _ in range(n): frame()
buf = C.create_string_buffer(4096); check(lib.zr_profile_collect(buf, 4096)); check(lib.zr_profile_enable(0))
out = {k: float(t) / n for k, c, t in (x.split(":") for x in buf.value.decode().split(";") if x)}
print(json.dumps({"workload": "Cornell 1080p, ReSTIR GI (emissive NEE, temporal reuse), ms per frame", "kernels_ms": out}))
