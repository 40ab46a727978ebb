"""derive-huss: oracle (CPU) and kernel (GPU) against outputs of the reference's own calc_huss_weedon2010."""
import os

import numpy as np
import pytest

from oracle import huss_oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "huss.npz")


def test_oracle_matches_reference_function():
    g = np.load(GOLDEN)
    out = huss_oracle.calc_huss_weedon2010(g["hurs"], g["ps"], g["tas"])
    assert np.array_equal(out, g["huss"], equal_nan=True)           # the same NumPy operations: bit-equal
    assert np.isnan(g["huss"][4:7]).all() and g["huss"][2] == 1e-7 and g["huss"][3] == 1e-1   # edge cases present


@pytest.mark.gpu
def test_kernel_matches_reference_function():
    import torch

    from attrici_b200.postprocess import calc_huss_weedon2010

    g = np.load(GOLDEN)
    out = calc_huss_weedon2010(g["hurs"], g["ps"], g["tas"])
    # same operation order with IEEE roundings; only exp() may differ by an ulp between libm and CUDA
    np.testing.assert_allclose(out, g["huss"], rtol=1e-14, equal_nan=True)
    assert np.array_equal(np.isnan(out), np.isnan(g["huss"]))
    assert np.array_equal(out == 1e-7, g["huss"] == 1e-7) and np.array_equal(out == 1e-1, g["huss"] == 1e-1)
    # odd length / unaligned views take the scalar kernel
    h, p, t = (torch.as_tensor(g[k], device="cuda")[1:1002] for k in ("hurs", "ps", "tas"))
    out2 = calc_huss_weedon2010(h.clone()[1:], p.clone()[1:], t.clone()[1:])
    np.testing.assert_allclose(out2.cpu().numpy(), g["huss"][2:1002], rtol=1e-14, equal_nan=True)


@pytest.mark.gpu
def test_kernel_streams_near_hbm_rate():
    import torch

    from attrici_b200.postprocess import calc_huss_weedon2010

    n = 1 << 27   # 1 GiB per array
    h = torch.rand(n, device="cuda", dtype=torch.float64) * 100
    p = torch.rand(n, device="cuda", dtype=torch.float64) * 5e4 + 5e4
    t = torch.rand(n, device="cuda", dtype=torch.float64) * 100 + 220
    calc_huss_weedon2010(h, p, t)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        calc_huss_weedon2010(h, p, t)
    e1.record()
    torch.cuda.synchronize()
    gbs = 5 * 32.0 * n / (e0.elapsed_time(e1) * 1e6)
    print(f"derive-huss: {gbs:.0f} GB/s algorithmic (24 B in + 8 B out per value)")
    assert gbs > 2000
