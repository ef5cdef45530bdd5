"""Lossless float codec of the P2P compression hook: bit-exact round trips (incl. NaN / Inf /
denormals / odd sizes) and a real size reduction on normally distributed data."""
import pytest
import torch

from uccl_b200.p2p.compress import Compressor

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("numel", [1, 31, 4096, 4097, 100003, (1 << 22) + 5])
def test_roundtrip_bit_exact(dtype, numel):
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(numel)
    x = (torch.randn(numel, generator=g) * 3).to(dtype)
    if numel > 64:  # special values and wildly different magnitudes in one block
        x[3] = float("nan")
        x[5] = float("inf")
        x[7] = -float("inf")
        x[11] = 0.0
        x[13] = -0.0
        x[17] = 1e-38 if dtype == torch.float32 else 1e-30
        x[19] = 3e38 if dtype == torch.float32 else 1e30
    x = x.to(dev)
    comp = Compressor("for")
    buf, nbytes = comp.compress(x)
    assert 64 < nbytes <= Compressor.bound(numel, dtype)
    y = comp.decompress(buf, numel, dtype)
    torch.cuda.synchronize()
    view = torch.int16 if dtype == torch.bfloat16 else torch.int32
    assert torch.equal(x.view(view).cpu(), y.view(view).cpu())


def test_compression_ratio_on_gaussian_data():
    dev = torch.device("cuda", 0)
    comp = Compressor("for")
    for dtype, limit in ((torch.bfloat16, 0.88), (torch.float32, 0.94)):
        x = torch.randn(1 << 22, device=dev).to(dtype)
        _, nbytes = comp.compress(x)
        ratio = nbytes / (x.numel() * x.element_size())
        print(f"{dtype}: compressed to {ratio:.3f}x")
        assert ratio < limit
    const = torch.full((1 << 20,), 1.5, device=dev, dtype=torch.bfloat16)  # zero exponent bits needed
    _, nbytes = comp.compress(const)
    assert nbytes / (const.numel() * 2) < 0.52
    assert not Compressor("none").wants(x) and Compressor("for").wants(x)
