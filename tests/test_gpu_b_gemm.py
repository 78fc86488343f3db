"""-m gpu: the tcgen05 3xTF32 GEMM vs an fp64 reference.  Tolerance: |err| <= 1e-4 * max(1, |ref|)
(measured ~4e-5 at K=1024: the tensor core's fp32 accumulator truncates on every MMA), i.e. 10x inside
the path's 1e-3 fp32 contract; a single TF32 pass would fail it."""
import pytest
import torch

from mp_former_b200 import native

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel_err(y, r):
    return ((y.double() - r).abs() / r.abs().clamp(min=1.0)).max().item()


def ref64(a, b, bias=None, relu=False):
    y = a.double() @ b.double().transpose(-1, -2)
    if bias is not None:
        y = y + bias.double()
    return y.relu() if relu else y


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 64, 256), (256, 128, 256), (1000, 100, 256),
                                   (4096, 288, 256), (300, 1024, 256), (777, 256, 1024), (128, 2048, 256),
                                   (65536, 100, 256)])
def test_gemm_matches_fp64(M, N, K):
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    a = torch.randn(M, K, device=DEV, generator=g)
    b = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    bias = torch.randn(N, device=DEV, generator=g)
    bh, bl = native.split_tf32(b)
    assert torch.equal((bh.double() + bl.double()).float(), b) or (bh + bl - b).abs().max() < 1e-6 * b.abs().max()
    for relu in (False, True):
        y = native.gemm_tf32x3(a, bh, bl, bias, relu=relu)
        r = ref64(a, b, bias, relu)
        err = rel_err(y, r)
        assert err < 1e-4, (M, N, K, relu, err)
    yt = native.gemm_tf32x3(a, bh, bl, None, transpose_c=True)
    assert yt.shape == (N, M)
    assert rel_err(yt.t(), ref64(a, b)) < 1e-4


def test_batched_transposed_mask_logit_shape():
    """mask logits: out[b,q,hw] = sum_c E[b,q,c] F[b,hw,c] (ref decoder :1865), F channels-last."""
    B, Q, C, H, W = 3, 100, 256, 64, 64
    g = torch.Generator(device=DEV).manual_seed(1)
    E = torch.randn(B, Q, C, device=DEV, generator=g)
    F_ = torch.randn(B, C, H, W, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    eh, el = native.split_tf32(E)
    a = F_.permute(0, 2, 3, 1).reshape(B, H * W, C)
    assert a.is_contiguous()
    out = native.gemm_tf32x3(a, eh, el, None, transpose_c=True).view(B, Q, H, W)
    ref = torch.einsum("bqc,bchw->bqhw", E.double(), F_.double())
    assert rel_err(out, ref) < 1e-4
    # a single-pass TF32 product of the same operands is >10x less accurate (why 3xTF32 is used)
    torch.backends.cuda.matmul.allow_tf32 = True
    tf32 = torch.einsum("bqc,bchw->bqhw", E, F_)
    torch.backends.cuda.matmul.allow_tf32 = False
    assert (tf32.double() - ref).abs().max().item() > 10 * (out.double() - ref).abs().max().item()


def test_strided_a_and_bad_args():
    g = torch.Generator(device=DEV).manual_seed(2)
    big = torch.randn(512, 512, device=DEV, generator=g)
    a = big[:, :256]                                   # row stride 512, K = 256
    b = torch.randn(96, 256, device=DEV, generator=g)
    bh, bl = native.split_tf32(b)
    y = native.gemm_tf32x3(a, bh, bl)
    assert rel_err(y, ref64(a, b)) < 1e-4
    # K that is not a multiple of 32: the tail k-block is zero-filled by TMA
    a48, b48 = torch.randn(8, 48, device=DEV, generator=g), torch.randn(8, 48, device=DEV, generator=g)
    assert rel_err(native.gemm_tf32x3(a48, *native.split_tf32(b48)), ref64(a48, b48)) < 1e-4
    with pytest.raises(RuntimeError, match="multiples of 4"):
        native.gemm_tf32x3(torch.randn(8, 30, device=DEV), *native.split_tf32(torch.randn(8, 30, device=DEV)))


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K,batch", [(128, 128, 32, 1), (300, 200, 104, 1), (1000, 256, 2048, 1),
                                         (256, 100, 120, 3), (4096, 256, 120, 2)])
def test_general_gemm_operand_majors_and_inkernel_split(a_mn, b_mn, M, N, K, batch):
    """MN-major operands (stored transposed) and in-kernel splitting of B; K tails are zero-filled."""
    if (K % 4 or M % 4 or N % 4):
        pytest.skip("strides must be multiples of 4 elements")
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N * 3 + K + batch)
    a = torch.randn(batch, M, K, device=DEV, generator=g)
    b = torch.randn(batch, N, K, device=DEV, generator=g) / K ** 0.5
    a_in = a.transpose(1, 2).contiguous() if a_mn else a
    b_in = b.transpose(1, 2).contiguous() if b_mn else b
    y = native.gemm_general(a_in, b_in, a_mn=a_mn, b_mn=b_mn)
    r = ref64(a, b)
    assert y.shape == (batch, M, N)
    assert rel_err(y, r) < 1e-4, (a_mn, b_mn, M, N, K, batch, rel_err(y, r))


def test_matmul_tn_weight_gradient_shape():
    g = torch.Generator(device=DEV).manual_seed(11)
    T = 21504 * 2
    gy = torch.randn(T, 288, device=DEV, generator=g)
    x = torch.randn(T, 256, device=DEV, generator=g)
    gw = native.matmul_tn(gy, x)
    ref = gy.double().t() @ x.double()
    assert gw.shape == (288, 256)
    # entries are sums of 43008 O(1) products (|ref| ~ 200): compare against the scale of the reduction
    assert (gw.double() - ref).abs().max().item() / ref.abs().max().item() < 1e-4
    # a token count with no 32-divisible split still works (single split, zero-filled K tail)
    gw2 = native.matmul_tn(gy[:1004].contiguous(), x[:1004].contiguous())
    ref2 = gy[:1004].double().t() @ x[:1004].double()
    assert (gw2.double() - ref2).abs().max().item() / ref2.abs().max().item() < 1e-4


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K,batch,splits", [(120, 256, 4096, 2, 8), (256, 256, 1000, 1, 5), (288, 256, 43008, 1, 42)])
def test_split_k(a_mn, b_mn, M, N, K, batch, splits):
    """In-kernel split-K: partial products of K ranges from different CTAs, summed by the wrapper; the last
    range may be ragged (zero-filled)."""
    g = torch.Generator(device=DEV).manual_seed(M + N + K + splits)
    a = torch.randn(batch, M, K, device=DEV, generator=g)
    b = torch.randn(batch, N, K, device=DEV, generator=g)
    a_in = a.transpose(1, 2).contiguous() if a_mn else a
    b_in = b.transpose(1, 2).contiguous() if b_mn else b
    y = native.gemm_general(a_in, b_in, a_mn=a_mn, b_mn=b_mn, k_splits=splits)
    r = ref64(a, b)
    assert y.shape == (batch, M, N)
    assert (y.double() - r).abs().max().item() / r.abs().max().item() < 1e-4
