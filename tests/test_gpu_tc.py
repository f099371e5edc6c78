"""GPU tests of the tcgen05 building blocks and the tensor-core engine."""
import pytest
import torch

import a3gc_ip_b200 as A
from a3gc_ip_b200 import _lib

pytestmark = pytest.mark.gpu


def to_image(m: torch.Tensor) -> torch.Tensor:
    """[rows, K] 16-bit -> K-major no-swizzle operand image [K/8][rows][8]."""
    rows, K = m.shape
    return m.view(rows, K // 8, 8).permute(1, 0, 2).contiguous()


@pytest.mark.parametrize("dtype,K,N", [(torch.float16, 64, 256), (torch.float16, 16, 64), (torch.bfloat16, 128, 128), (torch.float16, 256, 16)])
def test_umma_selftest(dtype, K, N):
    g = torch.Generator().manual_seed(K + N)
    a = torch.randn(128, K, generator=g).to(dtype)
    b = torch.randn(N, K, generator=g).to(dtype)
    want = a.float() @ b.float().t()
    ai, bi = to_image(a).cuda(), to_image(b).cuda()
    d = torch.full((128, N), float("nan"), device="cuda")
    flags = (1 if dtype == torch.bfloat16 else 0)
    rc = _lib.lib().a3gc_tc_selftest(ai.data_ptr(), bi.data_ptr(), d.data_ptr(), K, N, flags, _lib.stream_ptr(d.device))
    _lib.check(rc, "a3gc_tc_selftest")
    torch.cuda.synchronize()
    err = float((d.cpu() - want).abs().max() / want.abs().max())
    print(f"selftest {dtype} K={K} N={N}: max rel err = {err:.3e}")
    assert err < 1e-5


@pytest.mark.parametrize("K,N", [(64, 64), (256, 64)])
def test_umma_vector_operand(K, N):
    """A operand of 8 rows per K chunk with an 8-row-group stride of 0 (flag 4): D row r = A row (r & 7)."""
    g = torch.Generator().manual_seed(K * 3 + N)
    a = torch.randn(8, K, generator=g).to(torch.float16)
    b = torch.randn(N, K, generator=g).to(torch.float16)
    want = (a.float() @ b.float().t()).repeat(16, 1)
    ai, bi = to_image(a).cuda(), to_image(b).cuda()
    d = torch.full((128, N), float("nan"), device="cuda")
    rc = _lib.lib().a3gc_tc_selftest(ai.data_ptr(), bi.data_ptr(), d.data_ptr(), K, N, 4, _lib.stream_ptr(d.device))
    _lib.check(rc, "a3gc_tc_selftest")
    torch.cuda.synchronize()
    err = float((d.cpu() - want).abs().max() / want.abs().max())
    print(f"vector-operand selftest K={K} N={N}: max rel err = {err:.3e}")
    assert err < 1e-5
