"""Properties of the SVGF definition (oracle/orc_svgf.cpp) on synthetic inputs -- it has no reference counterpart to be pinned to, so
the CPU tier checks that it behaves like the filter it claims to be: constants are fixed points, noise on a flat surface shrinks by the
expected factor, depth / normal edges stop the kernel, invalid pixels pass through, the temporal stage integrates."""
import numpy as np

from tests import rpt_util
from tests.svgf_util import OracleSVGF

W, H = 96, 64


def _gbuffer(depth, normal_oct=0x7fff7fff):
    core = np.zeros((W * H, 4), dtype=np.uint32)
    core[:, 0] = np.asarray(depth, dtype=np.float32).reshape(-1).view(np.uint32)
    core[:, 1] = normal_oct
    me = np.zeros((W * H, 2), dtype=np.uint32)
    return core, me


def _fc():
    return rpt_util.FrameSequence(W, H, jitter=False).next()


def test_constant_image_is_a_fixed_point():
    core, me = _gbuffer(np.full((H, W), 3.0))
    sig = np.tile(np.array([0.5, 0.25, 0.125, 0.0], dtype=np.float32), (W * H, 1))
    for radius in (1, 2):
        out, _ = OracleSVGF(W, H, radius=radius).render(_fc(), core, me, sig)
        assert np.allclose(out[:, :3], sig[:, :3], rtol=2e-3)        # binary16 round trips per pass
        assert (out[:, 3] <= 1e-6).all()


def test_noise_on_a_flat_surface_is_filtered():
    rng = np.random.default_rng(3)
    core, me = _gbuffer(np.full((H, W), 3.0))
    sig = np.zeros((W * H, 4), dtype=np.float32)
    sig[:, :3] = (0.5 + 0.2 * rng.standard_normal((W * H, 1))).astype(np.float32)
    out, acc = OracleSVGF(W, H).render(_fc(), core, me, sig)
    inner = np.zeros((H, W), dtype=bool); inner[16:-16, 16:-16] = True
    s_in = sig[:, 0].reshape(H, W)[inner].std(); s_out = out[:, 0].reshape(H, W)[inner].std()
    assert s_out < 0.3 * s_in, (s_in, s_out)          # the luminance stop (4 sigma of a 3x3 estimate) keeps some
    assert abs(out[:, 0].reshape(H, W)[inner].mean() - 0.5) < 0.02
    # the variance estimate of the first frame comes from the 3x3 spatial fallback and is of the right size
    var = (acc[:, 1] >> 16).astype(np.uint16).view(np.float16).astype(np.float32).reshape(H, W)[inner]
    lum_var = (0.2 * (0.2126 + 0.7152 + 0.0722)) ** 2
    assert 0.3 * lum_var < var.mean() < 2.0 * lum_var, (var.mean(), lum_var)


def test_depth_and_normal_edges_stop_the_filter():
    depth = np.full((H, W), 2.0, dtype=np.float32); depth[:, W // 2:] = 4.0
    core, me = _gbuffer(depth)
    sig = np.zeros((W * H, 4), dtype=np.float32)
    img = np.zeros((H, W), dtype=np.float32); img[:, W // 2:] = 1.0
    sig[:, :3] = img.reshape(-1, 1)
    out, _ = OracleSVGF(W, H).render(_fc(), core, me, sig)
    o = out[:, 0].reshape(H, W)
    assert np.abs(o[:, :W // 2]).max() < 1e-3 and np.abs(o[:, W // 2:] - 1.0).max() < 1e-3
    # same depth, opposite normals: encode +z and -z octahedral (0.5, 0.5) is +z; flipped hemisphere for the right half
    core, me = _gbuffer(np.full((H, W), 3.0))
    n = core[:, 1].reshape(H, W).copy(); n[:, W // 2:] = 0x00000000
    core[:, 1] = n.reshape(-1)
    out, _ = OracleSVGF(W, H).render(_fc(), core, me, sig)
    o = out[:, 0].reshape(H, W)
    assert np.abs(o[:, :W // 2]).max() < 1e-3 and np.abs(o[:, W // 2:] - 1.0).max() < 1e-3


def test_invalid_pixels_pass_through_and_do_not_bleed():
    depth = np.full((H, W), 3.0, dtype=np.float32); depth[:20, :] = np.finfo(np.float32).max
    core, me = _gbuffer(depth)
    rng = np.random.default_rng(5)
    sig = np.zeros((W * H, 4), dtype=np.float32); sig[:, :3] = rng.random((W * H, 3)).astype(np.float32)
    sky = np.zeros((H, W), dtype=bool); sky[:20] = True
    sig[sky.reshape(-1), :3] = 100.0
    out, _ = OracleSVGF(W, H).render(_fc(), core, me, sig)
    assert np.allclose(out[sky.reshape(-1), :3], 100.0, rtol=1e-3)
    assert out[~sky.reshape(-1), :3].max() < 1.01


def test_temporal_stage_integrates():
    core, me = _gbuffer(np.full((H, W), 3.0))
    rng = np.random.default_rng(9)
    f = OracleSVGF(W, H, num_passes=1)
    seq = rpt_util.FrameSequence(W, H, jitter=False)
    means = []
    for i in range(12):
        sig = np.zeros((W * H, 4), dtype=np.float32)
        sig[:, :3] = (1.0 + 0.3 * rng.standard_normal((W * H, 1))).astype(np.float32)
        _, acc = f.render(seq.next(), core, me, sig)
        col = (acc[:, 0] & 0xffff).astype(np.uint16).view(np.float16).astype(np.float32)
        means.append(col.std())
    hist = f.hist[f.cur]
    N = (hist[:, 3] & 0xffff).astype(np.uint16).view(np.float16).astype(np.float32)
    assert (N == 12).all()
    assert means[-1] < 0.45 * means[0]          # alpha floors at 0.2: std -> sqrt(0.2 / 1.8) = 0.33 of the input
