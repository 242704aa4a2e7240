"""smoke(): three small frames of the whole hot path on cuda:0, checked bit-exactly against the CPU oracle."""


def run():
    from tests.test_rdi_gpu import _frame_loop
    problems, _ = _frame_loop("glossy", 160, 90, 3, full=True)
    assert not problems, "\n".join(problems)
    print("smoke frames ok (G-buffer -> ReSTIR DI -> ReSTIR PT -> compositing/firefly -> TAA, 160x90 x 3 frames, bit-exact vs oracle)")
