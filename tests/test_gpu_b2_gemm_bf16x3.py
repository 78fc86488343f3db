"""-m gpu: the tcgen05 bf16x3 GEMM (A split in-kernel, TMA-store epilogue) vs an fp64 reference.
Tolerance: |err| <= 2e-4 * max(1, |ref|) -- operands carry 16 significand bits (hi + lo bf16), i.e. a relative
error of ~2^-16 per product, 5x inside the path's 1e-3 fp32 contract (the 3xTF32 kernel is held to 1e-4)."""
import pytest
import torch

from mp_former_b200 import native

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 2e-4


def rel_err(y, r):
    return ((y.double() - r).abs() / r.abs().clamp(min=1.0)).max().item()


def ref64(a, b, bias=None, relu=False):
    y = a.double() @ b.double().transpose(-1, -2)
    if bias is not None:
        y = y + bias.double()
    return y.relu() if relu else y


def test_split_bf16_halves():
    x = torch.randn(1000, 64, device=DEV) * 3
    hi, lo = native.split_bf16(x)
    assert hi.dtype == torch.bfloat16 and lo.dtype == torch.bfloat16
    assert torch.equal(hi, x.to(torch.bfloat16))
    assert ((hi.float() + lo.float() - x).abs() <= x.abs() * 2.0 ** -16).all()


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 64, 256), (256, 128, 256), (1000, 100, 256),
                                   (4096, 288, 256), (300, 1024, 256), (777, 256, 1024), (128, 2048, 256),
                                   (65536, 100, 256), (21504, 768, 256), (40, 256, 48)])
def test_gemm_matches_fp64(M, N, K):
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    a = torch.randn(M, K, device=DEV, generator=g)
    b = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    bias = torch.randn(N, device=DEV, generator=g)
    bh, bl = native.split_bf16(b)
    for relu in (False, True):
        y = native.gemm(a, bh, bl, bias, relu=relu)
        assert y.shape == (M, N)
        err = rel_err(y, ref64(a, b, bias, relu))
        assert err < TOL, (M, N, K, relu, err)
    if M % 4 == 0:
        yt = native.gemm(a, bh, bl, None, transpose_c=True)
        assert yt.shape == (N, M)
        assert rel_err(yt.t(), ref64(a, b)) < TOL


def test_batched_transposed_mask_logit_shape():
    B, Q, C, H, W = 3, 120, 256, 64, 64
    g = torch.Generator(device=DEV).manual_seed(1)
    E = torch.randn(B, Q, C, device=DEV, generator=g) / C ** 0.5          # logits ~ N(0, 1)
    F_ = torch.randn(B, C, H, W, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    eh, el = native.split_bf16(E)
    a = F_.permute(0, 2, 3, 1).reshape(B, H * W, C)
    out = native.gemm(a, eh, el, None, transpose_c=True).view(B, Q, H, W)
    ref = torch.einsum("bqc,bchw->bqhw", E.double(), F_.double())
    assert rel_err(out, ref) < TOL


def test_epilogue_variants():
    g = torch.Generator(device=DEV).manual_seed(5)
    B, HW, E = 3, 1000, 256
    mem = torch.randn(B, HW, E, device=DEV, generator=g)
    w = torch.randn(E, E, device=DEV, generator=g) / 16
    bias = torch.randn(E, device=DEV, generator=g)
    pos_k = torch.randn(HW, E, device=DEV, generator=g)
    wh, wl = native.split_bf16(w)
    # row-periodic residual (K projection: memory Wk^T + (pos Wk^T + bk)), pre-split output
    hi, lo = native.gemm(mem.reshape(B * HW, E), wh, wl, None, resid=pos_k, resid_rows=HW, split_out=True)
    ref = (mem.double() @ w.double().t() + pos_k.double()).reshape(B * HW, E)
    assert rel_err(hi + lo, ref) < TOL
    assert torch.equal(hi.view(torch.int32) & 0x1FFF, torch.zeros_like(hi, dtype=torch.int32))   # TF32-exact halves
    # scale + bias (query projection), broadcast weight over a batch, transposed + split output (V^T)
    y = native.gemm(mem.reshape(B * HW, E), wh, wl, bias, alpha=0.25)
    assert rel_err(y, (mem.double().reshape(-1, E) @ w.double().t() + bias.double()) * 0.25) < TOL
    vt_hi, vt_lo = native.gemm(mem, wh[None].expand(B, -1, -1), wl[None].expand(B, -1, -1), bias, transpose_c=True,
                               split_out=True)
    refv = (mem.double() @ w.double().t() + bias.double()).transpose(1, 2)
    assert vt_hi.shape == (B, E, HW) and rel_err(vt_hi + vt_lo, refv) < TOL
    # gate (ReLU backward fused into the input-gradient GEMM of an FFN)
    T = 3000
    gy = torch.randn(T, 256, device=DEV, generator=g)
    w2 = torch.randn(256, 1024, device=DEV, generator=g) / 32
    hidden = torch.randn(T, 1024, device=DEV, generator=g).relu()
    w2t_hi, w2t_lo = native.split_bf16(w2.t().contiguous())
    gh = native.gemm_general(gy, w2t_hi, b_lo=w2t_lo, gate=hidden)
    refg = (gy.double() @ w2.double()) * (hidden > 0)
    assert rel_err(gh, refg) < TOL


def test_strided_a_and_k_tail():
    g = torch.Generator(device=DEV).manual_seed(2)
    big = torch.randn(512, 512, device=DEV, generator=g)
    a = big[:, :256]                                   # row stride 512, K = 256
    b = torch.randn(96, 256, device=DEV, generator=g)
    y = native.gemm(a, *native.split_bf16(b))
    assert rel_err(y, ref64(a, b)) < TOL * 16          # |b| ~ 1 here: outputs ~ N(0, 256)
    a48, b48 = torch.randn(8, 48, device=DEV, generator=g), torch.randn(8, 48, device=DEV, generator=g)
    assert rel_err(native.gemm(a48, *native.split_bf16(b48)), ref64(a48, b48)) < TOL * 4


def test_split_b_picks_kernel_by_shape():
    w = torch.randn(256, 256, device=DEV)
    assert native.split_b(w)[0].dtype == (torch.bfloat16 if native.GEMM_MODE == "bf16x3" else torch.float32)
    assert native.split_b(torch.randn(81, 256, device=DEV))[0].dtype == torch.float32      # N % 4 != 0 -> 3xTF32


@pytest.mark.parametrize("T,M,N,batch,splits", [(1024, 128, 256, 1, 1), (64, 128, 64, 1, 1), (43008, 288, 256, 1, 42),
                                                (1004, 256, 1024, 1, 3), (120, 4096, 256, 2, 1), (344, 100, 64, 3, 2),
                                                (5000, 1024, 256, 1, 5), (2048, 256, 192, 1, 4), (1920, 768, 256, 1, 2)])
def test_gemm_tn_matches_fp64(T, M, N, batch, splits):
    """C = A^T B over the token dimension (weight gradients, dF of the mask logits): both operands split in-kernel,
    MN-major SWIZZLE_128B operand tiles, ragged T / M / N tails, split-K."""
    g = torch.Generator(device=DEV).manual_seed(T + M + N + batch)
    a = torch.randn(batch, T, M, device=DEV, generator=g)
    b = torch.randn(batch, T, N, device=DEV, generator=g)
    y = native.gemm_tn(a, b, k_splits=splits)
    r = a.double().transpose(1, 2) @ b.double()
    assert y.shape == (batch, M, N)
    # entries are sums of T O(1) products: compare against the scale of the reduction (sqrt(T))
    assert (y.double() - r).abs().max().item() / T ** 0.5 < TOL, (y.double() - r).abs().max().item() / T ** 0.5
    if batch == 1:
        y2 = native.matmul_tn(a[0], b[0])
        assert (y2.double() - r[0]).abs().max().item() / T ** 0.5 < TOL


@pytest.mark.parametrize("T,M,N,batch,splits", [(1024, 128, 256, 1, 1), (43008, 288, 256, 1, 42), (1004, 256, 1024, 1, 3),
                                                (120, 4096, 256, 2, 1), (344, 100, 64, 3, 2), (440, 2048, 256, 1, 1),
                                                (3520, 256, 2048, 1, 4), (33, 84, 256, 1, 1)])
def test_gemm_tn_with_column_sums_matches_fp64(T, M, N, batch, splits):
    """Weight gradient dW = dY^T X and bias gradient db = dY^T 1 from ONE pass over dY (the converter warps of the TN
    kernel add up the A values they already hold): product as in the plain kernel, column sums exact to fp32 rounding
    of a T-term sum; ragged T / M tails, several N tiles (only the first one writes the sums), split-K, batches."""
    g = torch.Generator(device=DEV).manual_seed(T + M + N + batch)
    a = torch.randn(batch, T, M, device=DEV, generator=g)
    b = torch.randn(batch, T, N, device=DEV, generator=g)
    y, cs = native.gemm_tn(a, b, k_splits=splits, colsum=True)
    r = a.double().transpose(1, 2) @ b.double()
    assert y.shape == (batch, M, N) and cs.shape == (batch, M)
    assert (y.double() - r).abs().max().item() / T ** 0.5 < TOL
    assert torch.equal(y, native.gemm_tn(a, b, k_splits=splits))          # the product itself is unchanged
    assert (cs.double() - a.double().sum(1)).abs().max().item() / T ** 0.5 < 2e-6
    if batch == 1:
        y2, c2 = native.matmul_tn(a[0], b[0], with_colsum=True)
        assert (y2.double() - r[0]).abs().max().item() / T ** 0.5 < TOL
        assert (c2.double() - a[0].double().sum(0)).abs().max().item() / T ** 0.5 < 2e-6


@pytest.mark.parametrize("T,M,N,batch", [(120, 4096, 256, 2), (100, 1000, 64, 3), (37, 300, 128, 1)])
def test_gemm_tn_accumulate_into(T, M, N, batch):
    """TMA reduce-add epilogue: C += A^T B (sum over the prediction heads of the mask-feature gradient, ragged M tail
    included -- out-of-range rows of the last tile must not be added anywhere)."""
    g = torch.Generator(device=DEV).manual_seed(T + M + N)
    acc = torch.randn(batch, M, N, device=DEV, generator=g)
    ref = acc.double().clone()
    for i in range(3):
        a = torch.randn(batch, T, M, device=DEV, generator=g)
        b = torch.randn(batch, T, N, device=DEV, generator=g)
        out = native.gemm_tn(a, b, accumulate_into=acc)
        assert out.data_ptr() == acc.data_ptr()
        ref += a.double().transpose(1, 2) @ b.double()
    assert (acc.double() - ref).abs().max().item() / (3 * T) ** 0.5 < TOL
    with pytest.raises(RuntimeError):
        native.gemm_tn(a, b, accumulate_into=acc[:, :, : N // 2])


@pytest.mark.parametrize("batch,R,Cc", [(2, 256, 64), (1, 1000, 100), (3, 68, 256), (1, 4096, 256)])
def test_transpose_split_bf16(batch, R, Cc):
    """hi + lo of the transposed tensor: hi = bf16_rn(x^T) exactly, hi + lo reproduces x to 2^-16 relative."""
    g = torch.Generator(device=DEV).manual_seed(R + Cc)
    x = torch.randn(batch, R, Cc, device=DEV, generator=g) * 3
    hi, lo = native.transpose_split_bf16(x)
    assert hi.shape == (batch, Cc, R) and hi.dtype == torch.bfloat16
    xt = x.transpose(1, 2)
    assert torch.equal(hi, xt.to(torch.bfloat16))
    assert torch.equal(lo, (xt - hi.float()).to(torch.bfloat16))
    assert ((hi.float() + lo.float() - xt).abs() <= xt.abs() * 2.0 ** -16 + 1e-30).all()


@pytest.mark.parametrize("batch,M,N,K,splits", [(2, 300, 256, 4096, 5), (1, 1200, 64, 2048, 3), (3, 100, 128, 1000, 1)])
def test_gemm_bf16x3_split_k_with_transposed_operand(batch, M, N, K, splits):
    """The batched dE of the prediction heads: A [M x K] (fp32, split in-kernel) times a K-major B produced by the
    transposing split, reduction K cut across CTAs."""
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    a = torch.randn(batch, M, K, device=DEV, generator=g)
    f = torch.randn(batch, K, N, device=DEV, generator=g)                  # stored reduction-major, like the tokens
    hi, lo = native.transpose_split_bf16(f)
    y = native.gemm_bf16x3_splitk(a, hi, lo, splits)
    r = a.double() @ f.double()
    assert y.shape == (batch, M, N)
    assert (y.double() - r).abs().max().item() / K ** 0.5 < TOL


@pytest.mark.parametrize("M,N,K", [(1000, 1024, 256), (128, 32, 64), (4099, 256, 1024)])
def test_relu_bit_mask_epilogues(M, N, K):
    """relu output + one bit per element; the backward GEMM gated by those bits equals gating by the activation."""
    g = torch.Generator(device=DEV).manual_seed(M + N)
    a = torch.randn(M, K, device=DEV, generator=g)
    w = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    b = torch.randn(N, device=DEV, generator=g)
    w_hi, w_lo = native.split_bf16(w)
    y, bits = native.gemm_relu_bits(a, w_hi, w_lo, b, relu_bits_out=True)
    ref = torch.relu(a.double() @ w.double().t() + b.double())
    assert (y.double() - ref).abs().max().item() < TOL * max(1.0, ref.abs().max().item())
    assert bits.shape == (N // 32, M)
    unpacked = ((bits.view(N // 32, M, 1) >> torch.arange(32, device=DEV).view(1, 1, 32)) & 1).bool()   # [N/32, M, 32]
    assert torch.equal(unpacked.permute(1, 0, 2).reshape(M, N), y > 0)
    # backward: dH = (dY @ W2) gated
    w2 = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5          # acts as W2^T: [d_ffn, d_model]
    dy = torch.randn(M, K, device=DEV, generator=g)
    w2_hi, w2_lo = native.split_bf16(w2)
    gh = native.gemm_relu_bits(dy, w2_hi, w2_lo, gate_bits=bits)
    gh_ref = (dy.double() @ w2.double().t()) * (y > 0)
    assert (gh.double() - gh_ref).abs().max().item() < TOL * max(1.0, gh_ref.abs().max().item())
    assert torch.equal(gh == 0, ~(y > 0) | (gh == 0))


@pytest.fixture
def pair_mode():
    """Switches the K-major bf16x3 GEMM between single CTAs (0) and forced CTA pairs (2); restores the setting."""
    from mp_former_b200 import _lib
    lib = _lib.load()
    prev = lib.mpf_gemm_bf16x3_set_pair_mode(-1)
    yield lib.mpf_gemm_bf16x3_set_pair_mode
    lib.mpf_gemm_bf16x3_set_pair_mode(prev)


@pytest.mark.parametrize("M,N,K", [(256, 128, 256), (1000, 100, 256), (4096, 288, 256), (300, 1024, 256),
                                   (777, 256, 1024), (65536, 100, 256), (21504, 768, 256), (43008, 256, 256),
                                   (384, 256, 32)])
def test_gemm_cta_pairs_equal_single_cta(pair_mode, M, N, K):
    """tcgen05 cta_group::2 variant (clusters of two CTAs, M = 256 UMMA, half of the B tile per CTA) against the
    single-CTA kernel: the same products accumulate in the same order per element, so the results are bit-identical --
    with every epilogue variant (bias + ReLU, ReLU bits, gate bits, residual, transposed store, pre-split output), odd
    numbers of M tiles (the phantom tile of the last pair) and ragged N."""
    g = torch.Generator(device=DEV).manual_seed(M * 3 + N + K)
    a = torch.randn(M, K, device=DEV, generator=g)
    b = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    bias = torch.randn(N, device=DEV, generator=g)
    resid = torch.randn(M, N, device=DEV, generator=g)
    bh, bl = native.split_bf16(b)

    def run():
        out = [native.gemm(a, bh, bl, bias, relu=True), native.gemm(a, bh, bl, None, resid=resid, alpha=0.5)]
        if M % 4 == 0:
            out.append(native.gemm(a, bh, bl, None, transpose_c=True))
        out += list(native.gemm(a, bh, bl, bias, split_out=True))
        if N % 32 == 0:
            y, bits = native.gemm_relu_bits(a, bh, bl, bias, relu_bits_out=True)
            out += [y, bits, native.gemm_relu_bits(a, bh, bl, None, gate_bits=bits)]
        return out

    pair_mode(0)
    single = run()
    pair_mode(2)
    paired = run()
    torch.cuda.synchronize()
    assert rel_err(single[0], ref64(a, b, bias, True)) < TOL
    for s_, p_ in zip(single, paired):
        assert torch.equal(s_, p_)


def test_conv3x3_cta_pairs_equal_single_cta(pair_mode):
    g = torch.Generator(device=DEV).manual_seed(11)
    x = torch.randn(2, 40, 200, 64, device=DEV, generator=g)                      # channels-last [B, H, W, Cin]
    w = torch.randn(128, 9 * 64, device=DEV, generator=g) / 24
    bias = torch.randn(128, device=DEV, generator=g)
    wh, wl = native.split_bf16(w)
    pair_mode(0)
    a = native.conv3x3_cl(x, wh, wl, bias)
    pair_mode(2)
    b = native.conv3x3_cl(x, wh, wl, bias)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
