"""Writes tests/golden/oracle_hashes.json: SHA-256 of the CPU oracle's outputs for short frame sequences. The oracle has no
external pin for the lighting passes (the reference has no tests or CPU implementation for them); these hashes at least
freeze its behaviour, so an unintended change to the oracle shows up in the CPU test suite of every later round."""
import hashlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import scene_util, rpt_util


def run(which, w=96, h=54, nframes=3, presample=None, lvg=None):
    R = rpt_util.OracleRenderer(scene_util.SCENES[which](), w, h, nthreads=8)
    if presample:
        R.osc.set_presampling(*presample)
    if lvg:
        R.osc.set_light_voxel_grid(*lvg)
    cam = scene_util.CAMERAS.get(which, (0.0, 1.2, -4.043))
    seq = rpt_util.FrameSequence(w, h, cam_path=lambda f: (cam[0] + 0.02 * f, cam[1], cam[2]))
    taa_prev = np.zeros((w * h, 2), dtype=np.uint32)
    out = {}
    for i in range(nframes):
        fc = seq.next()
        R.gbuffer(fc); R.rdi(fc); R.rpt(fc); R.rgi(fc)
        comp, taa_prev = R.post(fc, taa_prev, i > 0)
    for name, arr in (("gbuffer_core", R.gb[R.cur][0]), ("di_reservoirs", R.di_curr_reservoirs()), ("di_final", R.di_final),
                      ("pt_reservoir_headers", np.stack([R.curr_reservoirs()["meta"], R.curr_reservoirs()["w_sum"].view(np.uint32),
                                                                 R.curr_reservoirs()["W"].view(np.uint32)])), ("pt_final", R.final),
                      ("gi_reservoirs", R.gi_curr_reservoirs()), ("gi_final", R.gi_final), ("taa", taa_prev)):
        out[name] = hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()
    R.pt(fc)        # the plain path tracer on the last frame's G-buffer (overwrites gi_final, hashed above)
    out["path_tracer_final"] = hashlib.sha256(np.ascontiguousarray(R.gi_final).tobytes()).hexdigest()
    if presample:
        out["sample_sets"] = hashlib.sha256(R.osc.sample_sets[:presample[0] * presample[1] * 10].tobytes()).hexdigest()
    if lvg:
        out["light_voxel_grid"] = hashlib.sha256(R.osc.lvg.tobytes()).hexdigest()
    return out


CASES = {
    "cornell": dict(which="cornell"),
    "glossy": dict(which="glossy"),
    "glass_presampled_lvg": dict(which="glass", presample=(16, 64), lvg=((8, 4, 8), (0.6, 0.45, 0.6), 0.1)),
    # the procedural C4 / C5 stand-ins (also freezes the scene generators)
    "atrium": dict(which="atrium", w=80, h=45),
    "tunnel": dict(which="tunnel", w=80, h=45),
}

if __name__ == "__main__":
    res = {k: run(**v) for k, v in CASES.items()}
    path = os.path.join(ROOT, "tests", "golden", "oracle_hashes.json")
    json.dump(res, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)
