"""ReSTIR DI (emissive) and the whole reference-shaped frame on the device vs the CPU oracle (bit-exact)."""
import ctypes as C
import numpy as np
import pytest

from tests.test_rpt_gpu import _diff_report


def _frame_loop(which, w, h, nframes, full=False, di_params=None, cam_path=None, accumulate=False, presample=None):
    from zetaray_b200 import lib, check, _lib
    from zetaray_b200.passes import Scene, GBuffers, GBufferRT, DirectLighting, IndirectLighting, Compositing, TAA, download_image
    from tests import scene_util, rpt_util
    flat = scene_util.SCENES[which]()
    R = rpt_util.OracleRenderer(flat, w, h)
    sc = Scene(flat)
    sc.prelighting()
    if presample:
        R.osc.set_presampling(*presample)
        sc.set_presampling(*presample)
    gb = GBuffers(w, h)
    gpass, di = GBufferRT(), DirectLighting(w, h)
    if di_params:
        for k, v in di_params.items():
            setattr(R.di_params, k, v)
        di.SetParams(**di_params)
    ind = IndirectLighting(w, h) if full else None
    comp = Compositing(w, h) if full else None
    taa = TAA(w, h) if full else None
    seq = rpt_util.FrameSequence(w, h, cam_path=cam_path, accumulate=accumulate)
    taa_prev = np.zeros((w * h, 2), dtype=np.uint32)
    problems = []
    for fr in range(nframes):
        fc = seq.next()
        R.gbuffer(fc)
        R.rdi(fc)
        gb.flip()
        fi = _lib.FrameInputs()
        fi.frame = fc
        gb.fill_inputs(fi)
        fi.scene = sc.handle
        gpass.Render(fi)
        sc.presample(fc.FrameNum)
        di.Render(fi)
        check(lib.zr_stream_synchronize(None))
        checks = [("di_reservoir", download_image(di.GetOutput(1), np.uint8, 32).view(rpt_util.RDI).reshape(-1), R.di_curr_reservoirs()),
                  ("di_final", download_image(di.GetOutput(0), np.float32, 4).view(np.uint32), R.di_final.view(np.uint32))]
        if fr >= 1 and R.di_params.temporal_resample and R.di_params.spatial_resample:
            checks.append(("di_target", download_image(di.GetOutput(2), np.uint32, 2), R.di_target))
        if full:
            R.rpt(fc)
            ind.Render(fi)
            comp.Render(fi, di.GetOutput(0).d_ptr, ind.GetOutput(0).d_ptr)
            taa.Render(fi, comp.GetOutput().d_ptr)
            check(lib.zr_stream_synchronize(None))
            ref_comp, ref_taa = R.post(fc, taa_prev, fr > 0)
            taa_prev = ref_taa
            checks.append(("indirect_final", download_image(ind.GetOutput(0), np.float32, 4).view(np.uint32), R.final.view(np.uint32)))
            checks.append(("composited", download_image(comp.GetOutput(), np.float32, 4).view(np.uint32), ref_comp.view(np.uint32)))
            checks.append(("taa", download_image(taa.GetOutput(), np.uint32, 2), ref_taa))
        for name, a, b in checks:
            msg = _diff_report(name, a, b)
            if msg:
                problems.append("frame %d: %s" % (fc.FrameNum, msg))
        if problems:
            break
    gb.close()
    return problems, R


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["cornell", "glossy"])
def test_rdi_frames(which):
    problems, R = _frame_loop(which, 320, 180, 4)
    assert not problems, "\n".join(problems)
    assert (R.di_curr_reservoirs()["lightIdx"] != 0xffffffff).sum() > 1000


@pytest.mark.gpu
def test_rdi_variants():
    problems, _ = _frame_loop("glossy", 256, 144, 3, di_params=dict(spatial_resample=0))
    assert not problems, "\n".join(problems)
    problems, _ = _frame_loop("glossy", 256, 144, 3, di_params=dict(stochastic_spatial=0, extra_disocclusion_sampling=0, M_max=8))
    assert not problems, "\n".join(problems)


@pytest.mark.gpu
@pytest.mark.parametrize("which,w,h", [("cornell", 480, 270), ("glossy", 333, 187)])
def test_full_frame_pipeline(which, w, h):
    # G-buffer -> pre-lighting -> ReSTIR DI -> ReSTIR PT -> compositing + firefly -> TAA, 4 frames
    problems, R = _frame_loop(which, w, h, 4, full=True)
    assert not problems, "\n".join(problems)


@pytest.mark.gpu
def test_full_frame_pipeline_glass_moving_camera():
    # transmissive scene + translating camera through the whole frame (DI disocclusion vote, TAA history reprojection)
    path = lambda f: (0.04 * f, 1.2, -4.043 + 0.03 * f)
    problems, _ = _frame_loop("glass", 320, 180, 5, full=True, cam_path=path)
    assert not problems, "\n".join(problems)


@pytest.mark.gpu
def test_full_frame_pipeline_accumulate():
    problems, _ = _frame_loop("glossy", 256, 144, 4, full=True, accumulate=True)
    assert not problems, "\n".join(problems)


@pytest.mark.gpu
def test_presampled_sets_di_and_full_frame():
    problems, _ = _frame_loop("glossy", 320, 180, 4, presample=(16, 64))
    assert not problems, "\n".join(problems)
    problems, _ = _frame_loop("cornell", 333, 187, 4, full=True, presample=(128, 512))
    assert not problems, "\n".join(problems)


@pytest.mark.gpu
def test_presampling_needs_the_presample_pass():
    from zetaray_b200 import lib, _lib
    from zetaray_b200.passes import Scene, GBuffers, GBufferRT, DirectLighting
    from tests import scene_util, rpt_util
    sc = Scene(scene_util.cornell())
    sc.prelighting()
    sc.set_presampling(4, 32)
    gb, g, di = GBuffers(64, 64), GBufferRT(), DirectLighting(64, 64)
    fi = _lib.FrameInputs()
    fi.scene = sc.handle
    fi.frame = rpt_util.FrameSequence(64, 64).next()
    gb.flip(); gb.fill_inputs(fi)
    g.Render(fi)
    assert lib.zr_direct_pass_render(di.handle, C.byref(fi), None) != 0      # presample pass has not run
    assert b"zr_presample_emissives" in lib.zr_last_error()
    sc.presample(1)
    di.Render(fi)
    assert lib.zr_scene_set_presampling(sc.handle, 4, 0) != 0
