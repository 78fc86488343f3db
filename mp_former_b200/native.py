"""Python launchers for the C-ABI kernels other than MSDeformAttn (tensor in, tensor out; the caller
owns autograd).  Every function requires CUDA fp32 tensors and raises RuntimeError otherwise."""
import os

import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32c(t, name):
    _lib.require_cuda(t, name)
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name}: expected float32, got {t.dtype}")
    return t


def split_tf32(x):
    """x = hi + lo with hi = rn_tf32(x), lo = rn_tf32(x - hi); returns (hi, lo), same shape."""
    x = _f32c(x, "x").contiguous()
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = _lib.load().mpf_split_tf32(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), _stream())
    _lib.check(rc, "split_tf32")
    return hi, lo


GEMM_MODE = os.environ.get("MPF_GEMM", "bf16x3")       # "bf16x3" (default) | "tf32x3"
# split-K of the TN GEMM: partial slabs + a sum kernel (default, bit-reproducible); MPF_TN_REDUCE_ADD=1 lets the
# splits add into one zeroed output through the TMA reduce-add epilogue instead -- measured no faster on the B200
# (139.8 vs 140.4 ms per step on one box: the memsets and the L2 read-modify-write cost what the tiny sums cost)
SPLITK_REDUCE_ADD = bool(os.environ.get("MPF_TN_REDUCE_ADD"))

# Optional per-launch device timing of the GEMM kernels (bench.py's roofline leg): CUDA events on the launching
# stream right around the C-ABI call, with the launch's algorithmic flops / bytes.  Off by default.
_PROFILE = None


def profile_begin():
    global _PROFILE
    _PROFILE = []


def profile_end():
    """-> {kernel: {"ms", "launches", "flops", "bytes"}} summed over the launches since profile_begin()."""
    global _PROFILE
    p, _PROFILE = _PROFILE, None
    torch.cuda.synchronize()
    out = {}
    for kind, flops, nbytes, a, b in p or []:
        r = out.setdefault(kind, {"ms": 0.0, "launches": 0, "flops": 0.0, "bytes": 0.0})
        r["ms"] += a.elapsed_time(b)
        r["launches"] += 1
        r["flops"] += flops
        r["bytes"] += nbytes
    return out


class _Timed:
    def __init__(self, kind, flops, nbytes):
        self.rec = (kind, flops, nbytes)

    def __enter__(self):
        if _PROFILE is not None:
            self.a, self.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if _PROFILE is not None:
            self.b.record()
            _PROFILE.append(self.rec + (self.a, self.b))
        return False


def split_bf16(x):
    """x = hi + lo with hi = bf16_rn(x), lo = bf16_rn(x - hi); returns (hi, lo) as bfloat16 tensors."""
    x = _f32c(x, "x").contiguous()
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty_like(hi)
    with torch.cuda.device(x.device):
        rc = _lib.load().mpf_split_bf16(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), _stream())
    _lib.check(rc, "split_bf16")
    return hi, lo


class WeightOperandCache:
    """The split weight operands of a training step, refreshed by ONE launch (csrc/weights.cu) instead of a split
    (and, for the input-gradient operand, a transposing copy) per use: ~300 launches per step of this head.

        cache = native.WeightOperandCache(model.parameters()); native.set_weight_cache(cache)
        for every step:  cache.begin_step();  loss = ...;  loss.backward();  optimizer.step()

    ``split_b(w)`` / ``split_bt(w)`` (the operand of ``w`` / of ``w.t()``) look the request up by address, shape and
    strides -- ``w`` may be a row slice of a registered parameter, e.g. the q / k / v blocks of a packed in-projection.
    A request seen for the first time is served the slow way and recorded; the next ``begin_step`` outside a graph
    capture builds the arena and the device table for everything recorded.  An entry is only used while the
    parameter's version counter equals the one it was filled at (an optimizer step without a following ``begin_step``
    falls back to the slow way instead of serving stale halves), and inside a CUDA-graph capture only if ``begin_step``
    was captured too (the replayed graph then refreshes the arena itself)."""

    def __init__(self, params):
        self.params = {}                 # storage address -> parameter (version counter shared by its views)
        for p in params:
            if p.dtype == torch.float32 and p.is_cuda:
                key = p.untyped_storage().data_ptr()
                if key in self.params and self.params[key] is not p:
                    # (the staleness check reads ONE version counter per storage: parameters that are separate
                    # tensors over one flat buffer would be checked against each other's counters)
                    raise ValueError("WeightOperandCache: parameters must own their storage (two of them share one)")
                self.params[key] = p
        self.entries = {}                # key -> [hi, lo, param, filled version]
        self.pending = {}                # key -> (geometry, param) recorded since the last build
        self.table = None
        self.total_tiles = 0
        self._keep = []                  # every arena / table ever built: a captured graph may still use an older one
        self._fresh_in_capture = False
        self.hits = self.misses = 0

    @staticmethod
    def _key(w, transposed):
        return (w.data_ptr(), tuple(w.shape), tuple(w.stride()), bool(transposed))

    def lookup(self, w, transposed):
        """-> (hi, lo) or None; records unknown requests on registered parameters."""
        if w.dim() != 2 or w.dtype != torch.float32 or w.stride(1) != 1 or not w.is_cuda:
            return None
        p = self.params.get(w.untyped_storage().data_ptr())
        if p is None:
            return None
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing:
            self._fresh_in_capture = False
        key = self._key(w, transposed)
        ent = self.entries.get(key)
        if ent is not None and ent[3] == p._version and (not capturing or self._fresh_in_capture):
            self.hits += 1
            return ent[0], ent[1]
        self.misses += 1
        if ent is None and key not in self.pending:
            self.pending[key] = (w.data_ptr(), w.shape[0], w.shape[1], w.stride(0), bool(transposed), p)
        return None

    def _build(self):
        geoms = [(k, (e[4], e[5], e[6], e[7], k[3], e[2])) for k, e in self.entries.items()]
        geoms += [(k, (g[0], g[1], g[2], g[3], g[4], g[5])) for k, g in self.pending.items()]
        self.pending = {}
        dev = next(iter(self.params.values())).device
        total = sum(g[1] * g[2] for _, g in geoms)
        arena = torch.empty(2 * total, dtype=torch.bfloat16, device=dev)
        words, tile0, off = [], 0, 0
        self.entries = {}
        for key, (ptr, rows, cols, ld, transposed, p) in geoms:
            n = rows * cols
            shape = (cols, rows) if transposed else (rows, cols)
            hi, lo = arena[off:off + n].view(shape), arena[off + n:off + 2 * n].view(shape)
            off += 2 * n
            words += [ptr, hi.data_ptr(), lo.data_ptr(), rows | (cols << 32), ld, int(transposed) | (tile0 << 32)]
            tile0 += ((rows + 31) // 32) * ((cols + 31) // 32)
            self.entries[key] = [hi, lo, p, -1, ptr, rows, cols, ld]
        self.table = torch.tensor(words, dtype=torch.int64, device=dev)
        self.total_tiles = tile0
        self._keep += [arena, self.table]

    def begin_step(self):
        """Refreshes every recorded operand from the current weights (one launch); call before the forward."""
        capturing = torch.cuda.is_current_stream_capturing()
        if self.pending and not capturing:          # (the table upload is a host -> device copy: never inside a capture)
            self._build()
        if not self.entries:
            return
        with torch.cuda.device(self.table.device):
            rc = _lib.load().mpf_split_weights_f32(self.table.data_ptr(), len(self.entries), self.total_tiles, _stream())
        _lib.check(rc, "split_weights")
        for ent in self.entries.values():
            ent[3] = ent[2]._version
        self._fresh_in_capture = capturing


_WEIGHT_CACHE = None


def set_weight_cache(cache):
    """Installs (or, with None, removes) the process-wide ``WeightOperandCache`` consulted by split_b / split_bt."""
    global _WEIGHT_CACHE
    _WEIGHT_CACHE = cache
    return cache


def split_b(w):
    """Pre-split the B operand ([..., N, K], K contiguous) of ``gemm`` / ``gemm_general``: bf16 halves for the
    bf16x3 kernel when its TMA constraints hold (K % 8 == 0, N % 4 == 0), TF32 halves for the 3xTF32 kernel
    otherwise (or when MPF_GEMM=tf32x3)."""
    if GEMM_MODE == "bf16x3" and w.shape[-1] % 8 == 0 and w.shape[-2] % 4 == 0:
        if _WEIGHT_CACHE is not None:
            hit = _WEIGHT_CACHE.lookup(w, False)
            if hit is not None:
                return hit
        return split_bf16(w)
    return split_tf32(w)


def split_bt(w):
    """``split_b(w.t().contiguous())`` -- the operand of the input-gradient product dx = dy W for a weight W [N, K]."""
    if GEMM_MODE == "bf16x3" and w.dim() == 2 and w.shape[0] % 8 == 0 and w.shape[1] % 4 == 0 and _WEIGHT_CACHE is not None:
        hit = _WEIGHT_CACHE.lookup(w, True)
        if hit is not None:
            return hit
    return split_b(w.t().contiguous())


def transpose_split_bf16(x):
    """x [batch, R, Cc] fp32 contiguous -> (hi, lo) bf16 [batch, Cc, R]: bf16x3 halves of the transpose."""
    x = _f32c(x, "x")
    if x.dim() != 3 or not x.is_contiguous():
        raise RuntimeError("transpose_split_bf16: x must be a contiguous [batch, R, Cc] tensor")
    batch, R, Cc = x.shape
    hi = torch.empty((batch, Cc, R), dtype=torch.bfloat16, device=x.device)
    lo = torch.empty_like(hi)
    with torch.cuda.device(x.device):
        rc = _lib.load().mpf_transpose_split_bf16(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), batch, R, Cc, _stream())
    _lib.check(rc, "transpose_split_bf16")
    return hi, lo


def gemm_bf16x3_splitk(a, b_hi, b_lo, k_splits):
    """C[i] = a[i] @ b[i]^T for a [batch, M, K] fp32 and pre-split bf16 b_hi / b_lo [batch, N, K], with the long
    reduction K cut into ``k_splits`` ranges handled by different CTAs (partial slabs summed here)."""
    a = _f32c(a, "a")
    batch, M, K = a.shape
    N = b_hi.shape[1]
    k_splits = max(1, min(int(k_splits), (K + 31) // 32))
    if b_hi.shape != (batch, N, K) or b_lo.shape != b_hi.shape or N % 4 or a.stride(2) != 1:
        raise RuntimeError(f"gemm_bf16x3_splitk: shapes a={tuple(a.shape)} b={tuple(b_hi.shape)}")
    out = torch.empty((batch * k_splits, M, N), dtype=torch.float32, device=a.device)
    nbytes = 4.0 * batch * (M * K + k_splits * M * N) + 4.0 * batch * N * K
    with torch.cuda.device(a.device), _Timed("gemm_bf16x3_kernel", 2.0 * batch * M * N * K, nbytes):
        rc = _lib.load().mpf_gemm_bf16x3(
            a.data_ptr(), a.stride(1), a.stride(0), b_hi.data_ptr(), b_lo.data_ptr(), K, N * K, None, out.data_ptr(),
            None, N, M * N, None, 0, 0, 0, None, 0, 1.0, batch, M, N, K, k_splits, 0, 0, _stream())
    _lib.check(rc, "gemm_bf16x3 (split-K)")
    return out.view(batch, k_splits, M, N).sum(1) if k_splits > 1 else out


def gemm_relu_bits(a, b_hi, b_lo, bias=None, resid=None, relu_bits_out=False, gate_bits=None):
    """y = a @ b^T (+ bias + resid) for a [M, K] fp32 and pre-split bf16 b [N, K] (N % 32 == 0), with the ReLU pattern
    as one bit per element.  ``relu_bits_out=True``: y = relu(...), returns (y, bits) with bits int32 [N/32, M];
    ``gate_bits``: the elements of y whose bit is clear are zeroed (backward through the ReLU), returns y."""
    a = _f32c(a, "a")
    M, K = a.shape
    N = b_hi.shape[0]
    if a.stride(1) != 1 or a.stride(0) % 4 or b_hi.shape != (N, K) or b_hi.dtype != torch.bfloat16 or N % 32 or K % 8:
        raise RuntimeError(f"gemm_relu_bits: unsupported operands a={tuple(a.shape)} b={tuple(b_hi.shape)}")
    y = torch.empty((M, N), dtype=torch.float32, device=a.device)
    bits = torch.empty((N // 32, M), dtype=torch.int32, device=a.device) if relu_bits_out else None
    if gate_bits is not None and (gate_bits.shape != (N // 32, M) or not gate_bits.is_contiguous()):
        raise RuntimeError("gemm_relu_bits: gate_bits must be a contiguous int32 [N/32, M] tensor")
    if resid is not None:
        resid = _f32c(resid, "resid")
        if resid.shape != (M, N) or resid.stride(1) != 1:
            raise RuntimeError("gemm_relu_bits: resid must be [M, N] with unit column stride")
    nbytes = 4.0 * (M * K + M * N + (M * N if resid is not None else 0)) + 4.0 * N * K + M * N / 8.0
    with torch.cuda.device(a.device), _Timed("gemm_bf16x3_kernel", 2.0 * M * N * K, nbytes):
        rc = _lib.load().mpf_gemm_bf16x3_relubits(
            a.data_ptr(), a.stride(0), b_hi.contiguous().data_ptr(), b_lo.contiguous().data_ptr(), K,
            None if bias is None else _f32c(bias, "bias").contiguous().data_ptr(), y.data_ptr(), N,
            None if resid is None else resid.data_ptr(), 0 if resid is None else resid.stride(0), M, N, K,
            int(relu_bits_out), None if bits is None else bits.data_ptr(),
            None if gate_bits is None else gate_bits.data_ptr(), _stream())
    _lib.check(rc, "gemm_bf16x3_relubits")
    return (y, bits) if relu_bits_out else y


def _gemm_bf16x3(a, b_hi, b_lo, bias, relu, transpose_c, split_out, resid, resid_rows, resid_cols, alpha, gate=None):
    """a [batch, M, K] fp32 (K contiguous); b_hi / b_lo bf16 [batch or 1, N, K] contiguous."""
    batch, M, K = a.shape
    nb, N = b_hi.shape[0], b_hi.shape[1]
    if b_hi.shape[2] != K or nb not in (1, batch) or b_lo.shape != b_hi.shape:
        raise RuntimeError(f"gemm: shape mismatch a={tuple(a.shape)} b={tuple(b_hi.shape)}")
    if (M if transpose_c else N) % 4 != 0:
        raise RuntimeError("gemm (bf16x3): the contiguous output dimension must be a multiple of 4")
    shape = (batch, N, M) if transpose_c else (batch, M, N)
    out = torch.empty(shape, dtype=torch.float32, device=a.device)
    out_lo = torch.empty_like(out) if split_out else None
    nbytes = 4.0 * batch * (M * K + M * N * (2 if split_out else 1) + (M * N if gate is not None else 0)) + 4.0 * nb * N * K
    with torch.cuda.device(a.device), _Timed("gemm_bf16x3_kernel", 2.0 * batch * M * N * K, nbytes):
        rc = _lib.load().mpf_gemm_bf16x3(
            a.data_ptr(), a.stride(1), a.stride(0) if batch > 1 else M * a.stride(1),
            b_hi.data_ptr(), b_lo.data_ptr(), K, N * K if (nb > 1) else 0,
            None if bias is None else bias.data_ptr(), out.data_ptr(),
            None if out_lo is None else out_lo.data_ptr(), M if transpose_c else N, out.stride(0),
            None if resid is None else resid.data_ptr(), 0 if resid is None else resid.stride(0),
            int(resid_rows), int(resid_cols), None if gate is None else gate.data_ptr(),
            0 if gate is None else gate.stride(0), float(alpha), batch, M, N, K, 1, int(relu), int(transpose_c),
            _stream())
    _lib.check(rc, "gemm_bf16x3")
    return out, out_lo


def gemm_tf32x3(a, b_hi, b_lo, bias=None, relu=False, transpose_c=False):
    """C[b] = A[b] @ B[b]^T (+bias)(ReLU).  a: [M,K] or [batch,M,K]; b_hi/b_lo: [N,K] or [batch,N,K]
    (K contiguous, row strides multiples of 4).  Returns [.., M, N] or, with ``transpose_c``,
    [.., N, M]."""
    a = _f32c(a, "a")
    squeeze = a.dim() == 2
    if squeeze:
        a, b_hi, b_lo = a[None], b_hi[None], b_lo[None]
    if a.stride(2) != 1 or a.stride(1) % 4 or (a.shape[0] > 1 and a.stride(0) % 4):
        a = a.contiguous()
    if not (b_hi.is_contiguous() and b_lo.is_contiguous()):
        b_hi, b_lo = b_hi.contiguous(), b_lo.contiguous()
    batch, M, K = a.shape
    N = b_hi.shape[1]
    if b_hi.shape != (batch, N, K) or b_lo.shape != b_hi.shape:
        raise RuntimeError(f"gemm_tf32x3: shape mismatch a={tuple(a.shape)} b={tuple(b_hi.shape)}")
    if bias is not None:
        bias = _f32c(bias, "bias").contiguous()
    out = torch.empty((batch, N, M) if transpose_c else (batch, M, N), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        rc = _lib.load().mpf_gemm_tf32x3(
            a.data_ptr(), a.stride(1), a.stride(0) if batch > 1 else M * a.stride(1),
            b_hi.data_ptr(), b_lo.data_ptr(), K, N * K,
            None if bias is None else bias.data_ptr(), out.data_ptr(),
            M if transpose_c else N, out.stride(0), batch, M, N, K, int(relu), int(transpose_c), _stream())
    _lib.check(rc, "gemm_tf32x3")
    return out[0] if squeeze else out


def gemm(a, b_hi, b_lo, bias=None, relu=False, transpose_c=False, split_out=False, resid=None,
         resid_rows=0, resid_cols=0, alpha=1.0):
    """Extended 3xTF32 GEMM (mpf_gemm_tf32x3_ex): x = (A @ B^T + bias + resid) * alpha (+ReLU).
    ``resid`` [R, >=resid_cols] is added to output row r from resid row r % resid_rows (0: row r), to
    columns < resid_cols (0: all).  ``split_out`` returns (hi, lo) TF32-exact halves."""
    a = _f32c(a, "a")
    squeeze = a.dim() == 2
    if squeeze:
        a, b_hi, b_lo = a[None], b_hi[None], b_lo[None]
    if a.stride(2) != 1 or a.stride(1) % 4 or (a.shape[0] > 1 and a.stride(0) % 4):
        a = a.contiguous()
    if bias is not None:
        bias = _f32c(bias, "bias").contiguous()
    if resid is not None:
        resid = _f32c(resid, "resid")
        if resid.dim() != 2 or resid.stride(1) != 1:
            resid = resid.reshape(-1, resid.shape[-1]).contiguous()
    if b_hi.dtype == torch.bfloat16:
        if b_hi.dim() == 3 and b_hi.shape[0] > 1 and b_hi.stride(0) == 0:     # expanded shared weight
            b_hi, b_lo = b_hi[:1], b_lo[:1]
        batch, M, K = a.shape
        N = b_hi.shape[1]
        tiles = batch * ((M + 127) // 128) * ((N + 255) // 256)
        if (K >= 1024 and tiles * 4 <= 148 and not (relu or transpose_c or split_out) and resid is None
                and alpha == 1.0 and b_hi.shape[0] == batch and N % 4 == 0):
            # a long reduction over a handful of output tiles (the decoder's second FFN product: 440 - 3520 query rows,
            # K = 2048): every CTA would walk 64 k-blocks back to back on 4 - 28 SMs (51 us measured).  The reduction
            # is cut across CTAs instead; the partial slabs are small.
            splits = max(2, min(K // 256, 148 // tiles))
            out = gemm_bf16x3_splitk(a, b_hi.contiguous(), b_lo.contiguous(), splits)
            if bias is not None:
                out = out + bias
            out = out.reshape(batch, M, N)
            return out[0] if squeeze else out
        out, out_lo = _gemm_bf16x3(a, b_hi.contiguous(), b_lo.contiguous(), bias, relu, transpose_c, split_out,
                                   resid, resid_rows, resid_cols, alpha)
        if squeeze:
            out = out[0]
            out_lo = None if out_lo is None else out_lo[0]
        return (out, out_lo) if split_out else out
    if not (b_hi.is_contiguous() and b_lo.is_contiguous()):
        b_hi, b_lo = b_hi.contiguous(), b_lo.contiguous()
    batch, M, K = a.shape
    N = b_hi.shape[1]
    if b_hi.shape != (batch, N, K) or b_lo.shape != b_hi.shape:
        raise RuntimeError(f"gemm: shape mismatch a={tuple(a.shape)} b={tuple(b_hi.shape)}")
    shape = (batch, N, M) if transpose_c else (batch, M, N)
    out = torch.empty(shape, dtype=torch.float32, device=a.device)
    out_lo = torch.empty_like(out) if split_out else None
    with torch.cuda.device(a.device):
        rc = _lib.load().mpf_gemm_tf32x3_ex(
            a.data_ptr(), a.stride(1), a.stride(0) if batch > 1 else M * a.stride(1),
            b_hi.data_ptr(), b_lo.data_ptr(), K, N * K,
            None if bias is None else bias.data_ptr(), out.data_ptr(),
            None if out_lo is None else out_lo.data_ptr(), M if transpose_c else N, out.stride(0),
            None if resid is None else resid.data_ptr(), 0 if resid is None else resid.stride(0),
            int(resid_rows), int(resid_cols), float(alpha), batch, M, N, K, int(relu), int(transpose_c),
            _stream())
    _lib.check(rc, "gemm_tf32x3_ex")
    if squeeze:
        out = out[0]
        out_lo = None if out_lo is None else out_lo[0]
    return (out, out_lo) if split_out else out


def add_layernorm_fwd(x, r, gamma, beta, eps):
    """y = LayerNorm(x + r) (r may be None) over the last dimension; returns (y, mean, rstd)."""
    x = _f32c(x, "x")
    C = x.shape[-1]
    x2 = x.reshape(-1, C)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    r2 = None
    if r is not None:
        r2 = _f32c(r, "r").reshape(-1, C)
        if not r2.is_contiguous():
            r2 = r2.contiguous()
    rows = x2.shape[0]
    y = torch.empty_like(x2)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    with torch.cuda.device(x.device):
        rc = _lib.load().mpf_add_layernorm_fwd_f32(x2.data_ptr(), None if r2 is None else r2.data_ptr(),
                                                   gamma.contiguous().data_ptr(), beta.contiguous().data_ptr(),
                                                   float(eps), rows, C, y.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                   _stream())
    _lib.check(rc, "add_layernorm_fwd")
    return y.view(x.shape), mean, rstd


def add_layernorm_bwd(dy, x, r, gamma, mean, rstd, with_colsum=False):
    """-> (dx, dgamma, dbeta[, colsum(dx)]); dx is the gradient of both x and r."""
    C = x.shape[-1]
    dy2 = _f32c(dy, "dy").reshape(-1, C)
    if not dy2.is_contiguous():
        dy2 = dy2.contiguous()
    x2 = x.reshape(-1, C)
    r2 = None if r is None else r.reshape(-1, C)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    if r2 is not None and not r2.is_contiguous():
        r2 = r2.contiguous()
    rows = x2.shape[0]
    lib = _lib.load()
    dx = torch.empty_like(x2)
    partial = torch.empty((lib.mpf_add_layernorm_partials(rows), 3, C), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.mpf_add_layernorm_bwd_f32(dy2.data_ptr(), x2.data_ptr(), None if r2 is None else r2.data_ptr(),
                                           gamma.contiguous().data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, C,
                                           dx.data_ptr(), partial.data_ptr(), _stream())
    _lib.check(rc, "add_layernorm_bwd")
    dgb = partial.sum(0)
    if with_colsum:
        return dx.view(x.shape), dgb[0], dgb[1], dgb[2]
    return dx.view(x.shape), dgb[0], dgb[1]


def groupnorm_cl_ok(C, groups):
    cpg = C // groups if groups > 0 and C % groups == 0 else 0
    return cpg in (4, 8, 16, 32) and C % 4 == 0 and C // 4 <= 256


def groupnorm_cl_fwd(x, gamma, beta, eps, groups, relu):
    """x [B, HW, C] contiguous (channels-last tokens) -> (y, mean [B,G], rstd [B,G])."""
    x = _f32c(x, "x")
    B, HW, C = x.shape
    y = torch.empty_like(x)
    mean = torch.empty((B, groups), dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    ws = torch.empty((B, groups, 2), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().mpf_groupnorm_cl_fwd_f32(x.data_ptr(), gamma.contiguous().data_ptr(),
                                                  beta.contiguous().data_ptr(), float(eps), B, HW, C, groups, int(relu),
                                                  y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), ws.data_ptr(), _stream())
    _lib.check(rc, "groupnorm_cl_fwd")
    return y, mean, rstd


def groupnorm_cl_bwd(dy, x, gamma, beta, mean, rstd, groups, relu):
    """-> (dx [B,HW,C], dgamma [C], dbeta [C])."""
    dy = _f32c(dy, "dy")
    B, HW, C = x.shape
    dx = torch.empty_like(x)
    dgb = torch.empty((2, C), dtype=torch.float32, device=x.device)
    ws = torch.empty((B, groups, 2), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().mpf_groupnorm_cl_bwd_f32(dy.data_ptr(), x.data_ptr(), gamma.contiguous().data_ptr(),
                                                  beta.contiguous().data_ptr(), mean.data_ptr(), rstd.data_ptr(), B, HW, C,
                                                  groups, int(relu), dx.data_ptr(), dgb.data_ptr(), ws.data_ptr(), _stream())
    _lib.check(rc, "groupnorm_cl_bwd")
    return dx, dgb[0], dgb[1]


def groupnorm_nchw2cl_ok(C, groups, HW):
    cpg = C // groups if groups > 0 and C % groups == 0 else 0
    return C % 64 == 0 and cpg > 0 and cpg % 4 == 0 and 64 % cpg == 0 and HW % 4 == 0


def groupnorm_nchw2cl_fwd(x, gamma, beta, eps, groups, relu):
    """x [B, C, HW] contiguous (NCHW) -> (y [B, HW, C] channels-last tokens, mean [B,G], rstd [B,G])."""
    x = _f32c(x, "x")
    B, C, HW = x.shape
    y = torch.empty((B, HW, C), dtype=torch.float32, device=x.device)
    mean = torch.empty((B, groups), dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    ws = torch.empty((B, groups, 2), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().mpf_groupnorm_nchw2cl_fwd_f32(x.data_ptr(), gamma.contiguous().data_ptr(),
                                                       beta.contiguous().data_ptr(), float(eps), B, HW, C, groups,
                                                       int(relu), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                       ws.data_ptr(), _stream())
    _lib.check(rc, "groupnorm_nchw2cl_fwd")
    return y, mean, rstd


def groupnorm_nchw2cl_bwd(dy, x, gamma, beta, mean, rstd, groups, relu):
    """dy [B, HW, C] (channels-last), x [B, C, HW] (NCHW) -> (dx [B, C, HW], dgamma [C], dbeta [C])."""
    dy = _f32c(dy, "dy")
    B, C, HW = x.shape
    dx = torch.empty_like(x)
    dgb = torch.empty((2, C), dtype=torch.float32, device=x.device)
    ws = torch.empty((B, groups, 2), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().mpf_groupnorm_nchw2cl_bwd_f32(dy.data_ptr(), x.data_ptr(), gamma.contiguous().data_ptr(),
                                                       beta.contiguous().data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                       B, HW, C, groups, int(relu), dx.data_ptr(), dgb.data_ptr(),
                                                       ws.data_ptr(), _stream())
    _lib.check(rc, "groupnorm_nchw2cl_bwd")
    return dx, dgb[0], dgb[1]


def upsample2x_add_ok(H, W, C):
    return H % 2 == 0 and W % 4 == 0 and C % 64 == 0 and H >= 2 and W >= 4


def upsample2x_add_nchw_fwd(cur, prev):
    """cur [B, H, W, C], prev [B, H/2, W/2, C] (both channels-last, contiguous) -> cur + bilinear_x2(prev) as
    NCHW-contiguous [B, C, H, W]."""
    cur, prev = _f32c(cur, "cur"), _f32c(prev, "prev")
    B, H, W, C = cur.shape
    if prev.shape != (B, H // 2, W // 2, C) or not cur.is_contiguous() or not prev.is_contiguous():
        raise RuntimeError(f"upsample2x_add: shapes {tuple(cur.shape)} / {tuple(prev.shape)} (contiguous channels-last)")
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=cur.device)
    with torch.cuda.device(cur.device):
        rc = _lib.load().mpf_upsample2x_add_nchw_fwd_f32(cur.data_ptr(), prev.data_ptr(), B, H, W, C, out.data_ptr(),
                                                         _stream())
    _lib.check(rc, "upsample2x_add_nchw_fwd")
    return out


def upsample2x_add_nchw_bwd(g):
    """g [B, C, H, W] NCHW-contiguous -> (g_cur [B, H, W, C], g_prev [B, H/2, W/2, C])."""
    g = _f32c(g, "g")
    if not g.is_contiguous():
        g = g.contiguous()
    B, C, H, W = g.shape
    g_cur = torch.empty((B, H, W, C), dtype=torch.float32, device=g.device)
    g_prev = torch.empty((B, H // 2, W // 2, C), dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        rc = _lib.load().mpf_upsample2x_add_nchw_bwd_f32(g.data_ptr(), B, H, W, C, g_cur.data_ptr(), g_prev.data_ptr(),
                                                         _stream())
    _lib.check(rc, "upsample2x_add_nchw_bwd")
    return g_cur, g_prev


def conv3x3_cl_ok(H, W, Cin, Cout):
    """Geometry covered by the tensor-core 3x3 convolution (forward / input gradient / weight gradient)."""
    return GEMM_MODE == "bf16x3" and H > 0 and W > 0 and Cin % 64 == 0 and Cout % 64 == 0


def conv3x3_cl(x, w_hi, w_lo, bias=None, relu=False):
    """x [B, H, W, Cin] fp32 contiguous (channels-last); w_hi / w_lo bf16 [Cout, 9*Cin] with column (ky*3+kx)*Cin + ci
    -> y [B, H, W, Cout]: 3x3 convolution, stride 1, zero padding 1, on the bf16x3 tensor-core GEMM."""
    x = _f32c(x, "x")
    if x.dim() != 4 or not x.is_contiguous():
        raise RuntimeError("conv3x3_cl: x must be a contiguous [B, H, W, C] tensor")
    B, H, W, Cin = x.shape
    Cout = w_hi.shape[0]
    if w_hi.shape != (Cout, 9 * Cin) or w_lo.shape != w_hi.shape or w_hi.dtype != torch.bfloat16:
        raise RuntimeError(f"conv3x3_cl: weight halves must be bf16 [{Cout}, {9 * Cin}]")
    y = torch.empty((B, H, W, Cout), dtype=torch.float32, device=x.device)
    flops = 2.0 * B * H * W * Cout * 9 * Cin
    with torch.cuda.device(x.device), _Timed("gemm_bf16x3_kernel", flops, 4.0 * B * H * W * (Cin + Cout) + 4.0 * Cout * 9 * Cin):
        rc = _lib.load().mpf_conv3x3_cl_bf16x3(x.data_ptr(), w_hi.contiguous().data_ptr(), w_lo.contiguous().data_ptr(),
                                               None if bias is None else _f32c(bias, "bias").contiguous().data_ptr(),
                                               y.data_ptr(), B, H, W, Cin, Cout, int(relu), _stream())
    _lib.check(rc, "conv3x3_cl")
    return y


def conv3x3_cl_wgrad(dy, x, target_tiles=296):
    """dy [B, H, W, Cout], x [B, H, W, Cin] (channels-last, contiguous) -> dW as [Cout, 9*Cin] (column (ky*3+kx)*Cin+ci):
    reduction over all pixels on the TN GEMM, the input map consumed tap by tap through shifted TMA boxes."""
    dy, x = _f32c(dy, "dy"), _f32c(x, "x")
    if not (dy.is_contiguous() and x.is_contiguous()) or dy.shape[:3] != x.shape[:3]:
        raise RuntimeError("conv3x3_cl_wgrad: dy / x must be contiguous [B, H, W, C] tensors of the same map size")
    B, H, W, Cin = x.shape
    Cout = dy.shape[3]
    N = 9 * Cin
    tiles = ((Cout + 127) // 128) * ((N + 255) // 256)
    T = B * H * W
    splits = max(1, min(148, target_tiles // max(1, tiles), (T + 1023) // 1024))
    out = torch.empty((splits, Cout, N), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device), _Timed("gemm_bf16x3_tn_kernel", 2.0 * T * Cout * N, 4.0 * T * (Cout * ((N + 255) // 256) + N)):
        rc = _lib.load().mpf_conv3x3_cl_wgrad_bf16x3(dy.data_ptr(), x.data_ptr(), out.data_ptr(), B, H, W, Cin, Cout, splits,
                                                     _stream())
    _lib.check(rc, "conv3x3_cl_wgrad")
    return out.sum(0) if splits > 1 else out[0]


def upsample2x_add_cl_fwd(cur, prev):
    """cur [B, H, W, C], prev [B, H/2, W/2, C] (channels-last, contiguous) -> cur + bilinear_x2(prev), channels-last."""
    cur, prev = _f32c(cur, "cur"), _f32c(prev, "prev")
    B, H, W, C = cur.shape
    if prev.shape != (B, H // 2, W // 2, C) or not cur.is_contiguous() or not prev.is_contiguous():
        raise RuntimeError(f"upsample2x_add_cl: shapes {tuple(cur.shape)} / {tuple(prev.shape)} (contiguous channels-last)")
    out = torch.empty_like(cur)
    with torch.cuda.device(cur.device):
        rc = _lib.load().mpf_upsample2x_add_cl_fwd_f32(cur.data_ptr(), prev.data_ptr(), B, H, W, C, out.data_ptr(), _stream())
    _lib.check(rc, "upsample2x_add_cl_fwd")
    return out


def upsample2x_cl_bwd(g):
    """g [B, H, W, C] channels-last contiguous -> gradient of the x2-upsampled map [B, H/2, W/2, C]."""
    g = _f32c(g, "g")
    if not g.is_contiguous():
        g = g.contiguous()
    B, H, W, C = g.shape
    g_prev = torch.empty((B, H // 2, W // 2, C), dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        rc = _lib.load().mpf_upsample2x_cl_bwd_f32(g.data_ptr(), B, H, W, C, g_prev.data_ptr(), _stream())
    _lib.check(rc, "upsample2x_cl_bwd")
    return g_prev


def colsum(x2):
    """Column sums of a [rows, C] fp32 matrix (row stride a multiple of 4): the bias gradient of a Linear layer."""
    x2 = _f32c(x2, "x")
    rows, C = x2.shape
    if x2.stride(1) != 1 or x2.stride(0) % 4 or C % 4 or C > 1024 or rows < 2048:
        return x2.sum(0)
    out = torch.empty(C, dtype=torch.float32, device=x2.device)
    with torch.cuda.device(x2.device):
        rc = _lib.load().mpf_colsum_f32(x2.data_ptr(), rows, C, x2.stride(0), out.data_ptr(), _stream())
    _lib.check(rc, "colsum")
    return out


def mask_words(n_keys):
    """uint32 words per mask row: whole 64-key tiles (the attention kernel reads two words per tile)."""
    return 2 * ((n_keys + 63) // 64)


def attn_mask_bits(logits, target_size):
    """logits [B, Q, H, W] fp32 -> packed bits int32 [B, Q, mask_words(h*w)], 1 = masked."""
    logits = _f32c(logits, "logits")
    if not logits.is_contiguous():
        logits = logits.contiguous()
    B, Q, H, W = logits.shape
    h, w = int(target_size[0]), int(target_size[1])
    words = mask_words(h * w)
    bits = torch.empty((B, Q, words), dtype=torch.int32, device=logits.device)
    with torch.cuda.device(logits.device):
        rc = _lib.load().mpf_attn_mask_bits_f32(logits.data_ptr(), H * W, B * Q, H, W, h, w, bits.data_ptr(),
                                                words, _stream())
    _lib.check(rc, "attn_mask_bits")
    return bits


def pack_bool_bits(mask):
    """bool [..., n] -> int32 [..., mask_words(n)], 1 = True; padding bits are 1."""
    _lib.require_cuda(mask, "mask")
    m8 = mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8).contiguous()
    n = mask.shape[-1]
    rows = mask.numel() // n
    words = mask_words(n)
    bits = torch.empty(tuple(mask.shape[:-1]) + (words,), dtype=torch.int32, device=mask.device)
    with torch.cuda.device(mask.device):
        rc = _lib.load().mpf_pack_bool_bits(m8.data_ptr(), rows, n, bits.data_ptr(), words, _stream())
    _lib.check(rc, "pack_bool_bits")
    return bits


def gt_mask_area_bits(masks, target_size):
    """masks [n, H, W] bool / uint8 -> packed bits int32 [n, mask_words(h*w)]; bit = 1 (masked) where the
    area-downsampled mask is <= 1e-8, i.e. the cell holds no instance pixel (ref decoder :986-987)."""
    _lib.require_cuda(masks, "masks")
    if masks.dtype not in (torch.bool, torch.uint8) or masks.dim() != 3:
        raise RuntimeError("gt_mask_area_bits: expected bool/uint8 masks [n, H, W]")
    m8 = masks.contiguous().view(torch.uint8)
    n, H, W = m8.shape
    h, w = int(target_size[0]), int(target_size[1])
    words = mask_words(h * w)
    bits = torch.empty((n, words), dtype=torch.int32, device=masks.device)
    if n == 0:
        return bits
    with torch.cuda.device(masks.device):
        rc = _lib.load().mpf_gt_mask_area_bits(m8.data_ptr(), n, H, W, h, w, bits.data_ptr(), words, _stream())
    _lib.check(rc, "gt_mask_area_bits")
    return bits


def unpack_bits(bits, n):
    """int32 [..., W] -> bool [..., n] (host-side helper for tests / the library backward)."""
    shifts = torch.arange(32, device=bits.device, dtype=torch.int32)
    b = ((bits.unsqueeze(-1) >> shifts) & 1).to(torch.bool)
    return b.flatten(-2)[..., :n]


XATTN_KEY_SPLITS = int(os.environ.get("MPF_XATTN_KEY_SPLITS", "0"))     # 0 = automatic; A/B measurements


def xattn_key_splits(B, Qt, HW, heads):
    """How many CTAs share the key tiles of one (query tile, head, image): as many as it takes to give every one of the
    148 SMs a CTA (the fused attention kernels keep ONE 190 KB CTA per SM), never more than there are 64-key tiles.
    B = 16: 128 CTAs -> 1 (unsplit); B = 2 (BASELINE configs[2] per GPU): 16 CTAs -> 9."""
    tiles = (HW + 63) // 64
    if XATTN_KEY_SPLITS > 0:
        return max(1, min(XATTN_KEY_SPLITS, tiles))
    ctas = B * heads * ((Qt + 127) // 128)
    return max(1, min(tiles, 148 // max(1, ctas)))


def masked_xattn_fwd(q_hi, q_lo, k_hi, k_lo, vt_hi, vt_lo, bits, row_open, heads):
    """Fused masked cross-attention forward.  Returns (out [B,Qt,E], lse2 [B,heads,Qt])."""
    B, Qt, E = q_hi.shape
    HW = k_hi.shape[1]
    for t, n in ((q_hi, "q_hi"), (q_lo, "q_lo"), (k_hi, "k_hi"), (k_lo, "k_lo"), (vt_hi, "vt_hi"), (vt_lo, "vt_lo")):
        _f32c(t, n)
        if not t.is_contiguous():
            raise RuntimeError(f"masked_xattn_fwd: {n} must be contiguous")
    if vt_hi.shape != (B, E, HW) or bits.shape[:2] != (B, Qt) or bits.dtype != torch.int32:
        raise RuntimeError("masked_xattn_fwd: inconsistent shapes")
    out = torch.empty((B, Qt, E), dtype=torch.float32, device=q_hi.device)
    lse2 = torch.empty((B, heads, Qt), dtype=torch.float32, device=q_hi.device)
    bits = bits.contiguous()
    if row_open is not None:
        row_open = row_open.to(torch.uint8).contiguous()
    splits = xattn_key_splits(B, Qt, HW, heads)
    ws_o = ws_ml = None
    if splits > 1:
        ws_o = torch.empty((splits, B, Qt, E), dtype=torch.float32, device=q_hi.device)
        ws_ml = torch.empty((splits, B, heads, Qt, 2), dtype=torch.float32, device=q_hi.device)
    # algorithmic: QK^T and PV (2 * Qt * HW * hd flop each per head); bytes: K, V^T hi+lo read once + mask bits
    with torch.cuda.device(q_hi.device), _Timed("masked_xattn_fwd", 4.0 * B * Qt * HW * E,
                                                4.0 * B * HW * E * 4 + B * Qt * HW / 8.0):
        rc = _lib.load().mpf_masked_xattn_fwd_f32_ex(
            q_hi.data_ptr(), q_lo.data_ptr(), k_hi.data_ptr(), k_lo.data_ptr(), vt_hi.data_ptr(),
            vt_lo.data_ptr(), bits.data_ptr(), None if row_open is None else row_open.data_ptr(),
            out.data_ptr(), lse2.data_ptr(), B, Qt, HW, heads, E // heads, bits.shape[2], splits,
            None if ws_o is None else ws_o.data_ptr(), None if ws_ml is None else ws_ml.data_ptr(), _stream())
    _lib.check(rc, "masked_xattn_fwd")
    return out, lse2


def gemm_general(a, b, a_mn=False, b_mn=False, b_lo=None, bias=None, relu=False, transpose_c=False, alpha=1.0,
                 k_splits=1, gate=None):
    """C[i] = op(A[i]) @ op(B[i])^T with either operand optionally "MN-major" (stored [batch, K, M-or-N],
    i.e. already transposed) and B split into TF32 halves in-kernel when ``b_lo`` is None.
    Shapes (3-D, batch first; 2-D inputs are treated as batch 1):
      a: [batch, M, K] (a_mn False) or [batch, K, M] (a_mn True);  b: [batch, N, K] or [batch, K, N].
    The innermost dimension must be contiguous; row / batch strides must be multiples of 4 elements."""
    a = _f32c(a, "a")
    if b.dtype == torch.bfloat16:
        if a_mn or b_mn or b_lo is None or k_splits != 1:
            raise RuntimeError("gemm_general: bf16 halves are only accepted for K-major operands without split-K")
        squeeze = a.dim() == 2
        if squeeze:
            a, b, b_lo = a[None], b[None], b_lo[None]
        if a.stride(2) != 1 or a.stride(1) % 4 or (a.shape[0] > 1 and a.stride(0) % 4):
            a = a.contiguous()
        if bias is not None:
            bias = _f32c(bias, "bias").contiguous()
        out, _ = _gemm_bf16x3(a, b.contiguous(), b_lo.contiguous(), bias, relu, transpose_c, False, None, 0, 0,
                              alpha, gate=gate)
        return out[0] if squeeze else out
    b = _f32c(b, "b")
    squeeze = a.dim() == 2
    if squeeze:
        a, b = a[None], b[None]
        b_lo = None if b_lo is None else b_lo[None]

    def fix(t):
        if t.stride(2) != 1 or t.stride(1) % 4 or (t.shape[0] > 1 and t.stride(0) % 4):
            return t.contiguous()
        return t
    a, b = fix(a), fix(b)
    if b_lo is not None:
        b_lo = b_lo.contiguous()
        b = b.contiguous()
    batch = a.shape[0]
    M, K = (a.shape[2], a.shape[1]) if a_mn else (a.shape[1], a.shape[2])
    N, Kb = (b.shape[2], b.shape[1]) if b_mn else (b.shape[1], b.shape[2])
    if Kb != K or b.shape[0] != batch:
        raise RuntimeError(f"gemm_general: inner dimensions differ ({tuple(a.shape)} vs {tuple(b.shape)})")
    if bias is not None:
        bias = _f32c(bias, "bias").contiguous()
    k_splits = max(1, min(int(k_splits), (K + 31) // 32))
    shape = (batch * k_splits, N, M) if transpose_c else (batch * k_splits, M, N)
    out = torch.empty(shape, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        rc = _lib.load().mpf_gemm_tf32x3_general(
            a.data_ptr(), int(a_mn), a.stride(1), a.stride(0) if batch > 1 else a.shape[1] * a.stride(1),
            b.data_ptr(), None if b_lo is None else b_lo.data_ptr(), int(b_mn), b.stride(1),
            b.stride(0) if batch > 1 else b.shape[1] * b.stride(1),
            None if bias is None else bias.data_ptr(), out.data_ptr(), None, M if transpose_c else N,
            out.stride(0), None, 0, 0, 0, None if gate is None else gate.data_ptr(),
            0 if gate is None else gate.stride(0), float(alpha), batch, M, N, K, k_splits, int(relu),
            int(transpose_c), _stream())
    _lib.check(rc, "gemm_tf32x3_general")
    if k_splits > 1:
        out = out.view(batch, k_splits, *out.shape[1:]).sum(1)
    return out[0] if squeeze else out


def gemm_tn(a, b, k_splits=1, accumulate_into=None, colsum=False):
    """C[i] = a[i]^T @ b[i] for a [batch, T, M], b [batch, T, N] (fp32, row-major; 2-D inputs = batch 1) on the
    bf16x3 TN kernel: both operands are split in-kernel, reduction over T, optional split-K (partials summed here).
    ``accumulate_into`` [batch, M, N] (contiguous fp32): the product is ADDED to it by the kernel's TMA reduce-add
    epilogue and the same tensor is returned.  Split-K (``k_splits > 1``): partial slabs summed here (bit-reproducible); with
    MPF_TN_REDUCE_ADD=1 the splits add into one zeroed output through the same epilogue instead."""
    a, b = _f32c(a, "a"), _f32c(b, "b")
    squeeze = a.dim() == 2
    if squeeze:
        a, b = a[None], b[None]

    def fix(t):
        if t.stride(2) != 1 or t.stride(1) % 4 or (t.shape[0] > 1 and t.stride(0) % 4):
            return t.contiguous()
        return t
    a, b = fix(a), fix(b)
    batch, T, M = a.shape
    N = b.shape[2]
    if b.shape[0] != batch or b.shape[1] != T:
        raise RuntimeError(f"gemm_tn: shape mismatch a={tuple(a.shape)} b={tuple(b.shape)}")
    if N % 4:
        raise RuntimeError("gemm_tn: N must be a multiple of 4")
    k_splits = max(1, min(int(k_splits), (T + 31) // 32))
    own = False
    if accumulate_into is None and k_splits > 1 and SPLITK_REDUCE_ADD:
        # split-K without partial slabs: the splits add into one zeroed output through the TMA reduce-add epilogue
        accumulate_into = torch.zeros((batch, M, N), dtype=torch.float32, device=a.device)
        own = True
    elif accumulate_into is not None:
        k_splits = 1
    if accumulate_into is not None:
        acc = accumulate_into
        if (acc.dtype != torch.float32 or not acc.is_contiguous() or acc.numel() != batch * M * N
                or acc.device != a.device):
            raise RuntimeError("gemm_tn: accumulate_into must be a contiguous fp32 [batch, M, N] tensor on the same device")
        with torch.cuda.device(a.device), _Timed("gemm_bf16x3_tn_kernel", 2.0 * batch * M * N * T,
                                                 4.0 * batch * (T * M + T * N + 2 * M * N)):
            rc = _lib.load().mpf_gemm_bf16x3_tn_ex(
                a.data_ptr(), a.stride(1), a.stride(0) if batch > 1 else T * a.stride(1),
                b.data_ptr(), b.stride(1), b.stride(0) if batch > 1 else T * b.stride(1),
                acc.data_ptr(), N, M * N, batch, M, N, T, k_splits, 1, _stream())
        _lib.check(rc, "gemm_bf16x3_tn_ex")
        return acc[0] if (own and squeeze) else acc
    if colsum:
        # ``colsum=True``: also returns the column sums of ``a`` over T ([batch, M]: the bias gradient that goes with
        # the weight gradient), accumulated by the kernel's converter warps from the values they already hold
        # (each slab of C is followed by its column sums, so that ONE reduction over the splits sums both)
        slab = M * N + M
        buf = torch.empty((batch * k_splits, slab), dtype=torch.float32, device=a.device)
        with torch.cuda.device(a.device), _Timed("gemm_bf16x3_tn_kernel", 2.0 * batch * M * N * T,
                                                 4.0 * batch * (T * M + T * N + k_splits * M * N)):
            rc = _lib.load().mpf_gemm_bf16x3_tn_colsum(
                a.data_ptr(), a.stride(1), a.stride(0) if batch > 1 else T * a.stride(1),
                b.data_ptr(), b.stride(1), b.stride(0) if batch > 1 else T * b.stride(1),
                buf.data_ptr(), N, slab, batch, M, N, T, k_splits, buf.data_ptr() + 4 * M * N, slab, _stream())
        _lib.check(rc, "gemm_bf16x3_tn_colsum")
        if k_splits > 1:
            buf = buf.view(batch, k_splits, slab).sum(1)
        out, cs = buf[:, :M * N].view(batch, M, N), buf[:, M * N:]
        return (out[0], cs[0]) if squeeze else (out, cs)
    out = torch.empty((batch * k_splits, M, N), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device), _Timed("gemm_bf16x3_tn_kernel", 2.0 * batch * M * N * T,
                                             4.0 * batch * (T * M + T * N + k_splits * M * N)):
        rc = _lib.load().mpf_gemm_bf16x3_tn(
            a.data_ptr(), a.stride(1), a.stride(0) if batch > 1 else T * a.stride(1),
            b.data_ptr(), b.stride(1), b.stride(0) if batch > 1 else T * b.stride(1),
            out.data_ptr(), N, M * N, batch, M, N, T, k_splits, _stream())
    _lib.check(rc, "gemm_bf16x3_tn")
    if k_splits > 1:
        out = out.view(batch, k_splits, M, N).sum(1)
    return out[0] if squeeze else out


def matmul_tn(x, y, target_tiles=296, with_colsum=False):
    """x^T @ y for x [T, M], y [T, N] (reduction over the long token dimension T), e.g. the weight gradient
    dW = dY^T X of an nn.Linear: both operands are consumed MN-major straight from their row-major
    storage; T is cut into K-splits handled by different CTAs whose partial products are summed.
    ``with_colsum``: returns (x^T @ y, x.sum(0)) -- weight and bias gradient of the layer from one pass over dY."""
    T, M = x.shape
    N = y.shape[1]
    if GEMM_MODE == "bf16x3" and N % 4 == 0 and M % 4 == 0:
        tiles = ((M + 127) // 128) * ((N + 255) // 256)
        splits = max(1, min(148, target_tiles // max(1, tiles), (T + 1023) // 1024))
        return gemm_tn(x, y, k_splits=splits, colsum=with_colsum)
    if with_colsum:
        return matmul_tn(x, y, target_tiles), colsum(x)
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    splits = max(1, min(64, target_tiles // max(1, tiles), (T + 1023) // 1024))
    return gemm_general(x, y, a_mn=True, b_mn=True, k_splits=splits)


def masked_xattn_bwd(q_hi, q_lo, k_hi, k_lo, kt_hi, kt_lo, v_hi, v_lo, d_o, bits, row_open, lse2, delta, heads):
    """Backward of the fused masked cross-attention.  q_*: pre-scaled query halves [B,Qt,E]; k_*, v_*:
    [B,HW,E]; kt_*: K^T [B,E,HW]; d_o: gradient of the attention output [B,Qt,E]; lse2, delta: [B,heads,Qt].
    Returns (dq wrt the unscaled query projection, dk, dv)."""
    B, Qt, E = q_hi.shape
    HW = k_hi.shape[1]
    qt_ld = (Qt + 3) // 4 * 4
    do_hi, do_lo = split_tf32(d_o)

    def transposed(t):                     # [B,Qt,E] -> [B,E,qt_ld] (zero padded); exact, so halves stay halves
        out = torch.zeros((B, E, qt_ld), dtype=torch.float32, device=t.device)
        out[:, :, :Qt] = t.transpose(1, 2)
        return out
    qt_hi, qt_lo, dot_hi, dot_lo = transposed(q_hi), transposed(q_lo), transposed(do_hi), transposed(do_lo)
    dq = torch.empty((B, Qt, E), dtype=torch.float32, device=q_hi.device)
    dk = torch.empty((B, HW, E), dtype=torch.float32, device=q_hi.device)
    dv = torch.empty_like(dk)
    ro = None if row_open is None else row_open.to(torch.uint8).contiguous()
    ts = [q_hi, q_lo, qt_hi, qt_lo, k_hi, k_lo, kt_hi, kt_lo, v_hi, v_lo, do_hi, do_lo, dot_hi, dot_lo]
    for t in ts:
        if not t.is_contiguous():
            raise RuntimeError("masked_xattn_bwd: operands must be contiguous")
    bits = bits.contiguous()
    lse2, delta = lse2.contiguous(), delta.contiguous()
    splits = xattn_key_splits(B, Qt, HW, heads)
    ws_dq = torch.empty((splits, B, Qt, E), dtype=torch.float32, device=q_hi.device) if splits > 1 else None
    # algorithmic: S recomputed twice, dP, dQ, dK, dV (2 * Qt * HW * hd flop each per head)
    with torch.cuda.device(q_hi.device), _Timed("masked_xattn_bwd", 12.0 * B * Qt * HW * E,
                                                4.0 * B * HW * E * 8 + B * Qt * HW / 8.0):
        rc = _lib.load().mpf_masked_xattn_bwd_f32_ex(
            *[t.data_ptr() for t in ts], bits.data_ptr(), None if ro is None else ro.data_ptr(),
            lse2.data_ptr(), delta.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
            B, Qt, qt_ld, HW, heads, E // heads, bits.shape[2], splits,
            None if ws_dq is None else ws_dq.data_ptr(), _stream())
    _lib.check(rc, "masked_xattn_bwd")
    return dq, dk, dv


# ----------------------------------------------------------------------------------------------------------------
# Hungarian matching on the device (ref mask2former/modeling/matcher.py:96-157)
# ----------------------------------------------------------------------------------------------------------------
SHARED_POINT_BAND_BYTES = 96 << 10       # shared memory of one band of sample_shared_points (two CTAs per SM)


def shared_point_bands(H, W):
    """(rows per band, number of bands) of ``sample_shared_points`` for H x W maps."""
    rows = max(1, min(H, SHARED_POINT_BAND_BYTES // (4 * W) - 1))
    return rows, -(-H // rows)


def sample_shared_points(pred_masks, point_coords, band_lo, band_rows):
    """pred_masks [B, Q, H, W] f32 (each map contiguous, W % 4 == 0) sampled at the image's shared points
    point_coords [B, P, 2] -> [B, Q, P], by streaming every map once through shared memory (csrc/matcher.cu
    sample_shared_points_kernel).  ``band_lo`` int32 [B, n_bands + 1]: first point of every band of ``band_rows`` rows
    in the row-major ordered point list."""
    _f32c(pred_masks, "pred_masks")
    _f32c(point_coords, "point_coords")
    _lib.require_cuda(band_lo, "band_lo")
    B, Q, H, W = pred_masks.shape
    P = point_coords.shape[1]
    n_bands = band_lo.shape[1] - 1
    if pred_masks.stride(-1) != 1 or pred_masks.stride(-2) != W or W % 4 or band_lo.dtype != torch.int32 or \
            band_lo.shape[0] != B or not band_lo.is_contiguous() or point_coords.shape != (B, P, 2):
        raise RuntimeError("sample_shared_points: contiguous maps with W % 4 == 0, point_coords [B, P, 2] and int32 "
                           "band_lo [B, n_bands + 1] expected")
    point_coords = point_coords.contiguous()
    out = torch.empty((B, Q, P), dtype=torch.float32, device=pred_masks.device)
    with torch.cuda.device(pred_masks.device):
        rc = _lib.load().mpf_sample_shared_points_f32(
            pred_masks.data_ptr(), pred_masks.stride(0), pred_masks.stride(1), H, W, point_coords.data_ptr(),
            band_lo.data_ptr(), n_bands, int(band_rows), B, Q, P, out.data_ptr(), _stream())
    _lib.check(rc, "sample_shared_points")
    return out


def match_cost(pred_logits, pred_masks, tgt_mask_ptrs, tgt_is_f32, tgt_hw, tgt_labels, tgt_offsets, counts,
               point_coords, cost_class, cost_mask, cost_dice, sampled=None):
    """Cost matrices of a whole batch in one pass.  pred_logits [B, Q, K+1] f32, pred_masks [B, Q, H, W] f32 (may be
    a query slice of a larger tensor), tgt_mask_ptrs int64 [B] (device pointers to each image's [n_b, Hg, Wg] masks),
    tgt_labels int64 [ntot], tgt_offsets int32 [B+1], counts = the host list of n_b, point_coords [B, P, 2].
    ``sampled`` [B, Q, P]: the prediction logits already sampled at the points (``sample_shared_points``); pred_masks
    is then only consulted for its shape.
    Returns a flat float tensor: image b's row-major [Q, n_b] matrix at Q * offsets[b]."""
    for t, n in ((pred_logits, "pred_logits"), (pred_masks, "pred_masks"), (tgt_mask_ptrs, "tgt_mask_ptrs"),
                 (tgt_labels, "tgt_labels"), (tgt_offsets, "tgt_offsets"), (point_coords, "point_coords")):
        _lib.require_cuda(t, n)
    if pred_logits.dtype != torch.float32 or pred_masks.dtype != torch.float32 or point_coords.dtype != torch.float32:
        raise RuntimeError("match_cost: pred_logits, pred_masks and point_coords must be float32")
    if pred_logits.dim() != 3 or pred_masks.dim() != 4 or pred_logits.shape[:2] != pred_masks.shape[:2]:
        raise RuntimeError("match_cost: expected pred_logits [B, Q, K+1] and pred_masks [B, Q, H, W]")
    B, Q, K1 = pred_logits.shape
    H, W = pred_masks.shape[-2:]
    if pred_masks.stride(-1) != 1 or pred_masks.stride(-2) != W:
        pred_masks = pred_masks.contiguous()
    if pred_logits.stride(-1) != 1:
        pred_logits = pred_logits.contiguous()
    point_coords = point_coords.contiguous()
    if point_coords.shape[0] != B or point_coords.shape[-1] != 2 or point_coords.dim() != 3:
        raise RuntimeError("match_cost: point_coords must be [B, P, 2]")
    if tgt_mask_ptrs.dtype != torch.int64 or tgt_labels.dtype != torch.int64 or tgt_offsets.dtype != torch.int32:
        raise RuntimeError("match_cost: tgt_mask_ptrs / tgt_labels must be int64 and tgt_offsets int32")
    if len(counts) != B or tgt_offsets.numel() != B + 1 or tgt_mask_ptrs.numel() != B:
        raise RuntimeError("match_cost: one target entry per image expected")
    P = point_coords.shape[1]
    ntot, nmax = int(sum(counts)), int(max(counts))
    if tgt_labels.numel() != ntot:
        raise RuntimeError("match_cost: tgt_labels does not hold sum(counts) labels")
    cost = torch.empty(Q * ntot, dtype=torch.float32, device=pred_masks.device)
    if ntot == 0:
        return cost
    lib = _lib.load()
    ws_bytes = int(lib.mpf_match_cost_workspace_bytes(B, Q, ntot, nmax, P))
    if ws_bytes < 0:
        raise RuntimeError("match_cost: bad sizes")
    ws = torch.empty((ws_bytes + 3) // 4, dtype=torch.float32, device=pred_masks.device)
    with torch.cuda.device(pred_masks.device):
        if sampled is not None:
            if sampled.dtype != torch.float32 or tuple(sampled.shape) != (B, Q, P) or not sampled.is_contiguous():
                raise RuntimeError("match_cost: sampled must be a contiguous float32 [B, Q, P] tensor")
            rc = lib.mpf_match_cost_presampled_f32(
                pred_logits.data_ptr(), pred_logits.stride(0), pred_logits.stride(1), K1, sampled.data_ptr(), H, W,
                tgt_mask_ptrs.data_ptr(), int(bool(tgt_is_f32)), int(tgt_hw[0]), int(tgt_hw[1]),
                tgt_labels.data_ptr(), tgt_offsets.data_ptr(), ntot, nmax, point_coords.data_ptr(), B, Q, P,
                float(cost_class), float(cost_mask), float(cost_dice), ws.data_ptr(), ws.numel() * 4, cost.data_ptr(),
                _stream())
        else:
            rc = lib.mpf_match_cost_f32(
                pred_logits.data_ptr(), pred_logits.stride(0), pred_logits.stride(1), K1,
                pred_masks.data_ptr(), pred_masks.stride(0), pred_masks.stride(1), H, W,
                tgt_mask_ptrs.data_ptr(), int(bool(tgt_is_f32)), int(tgt_hw[0]), int(tgt_hw[1]),
                tgt_labels.data_ptr(), tgt_offsets.data_ptr(), ntot, nmax, point_coords.data_ptr(), B, Q, P,
                float(cost_class), float(cost_mask), float(cost_dice), ws.data_ptr(), ws.numel() * 4, cost.data_ptr(),
                _stream())
    _lib.check(rc, "match_cost")
    return cost


def lsap(cost, tgt_offsets, counts, num_queries):
    """Solves every image's [Q, n_b] assignment problem on the device (scipy's algorithm and tie rule).
    Returns (query_idx int64 [m], target_idx int64 [m], status int32 [1]) with m = sum_b min(Q, n_b); image b's pairs
    start at sum_{b'<b} min(Q, n_b').  No host synchronisation."""
    _lib.require_cuda(cost, "cost")
    _lib.require_cuda(tgt_offsets, "tgt_offsets")
    if cost.dtype != torch.float32 or tgt_offsets.dtype != torch.int32:
        raise RuntimeError("lsap: cost must be float32 and tgt_offsets int32")
    B, Q = len(counts), int(num_queries)
    if tgt_offsets.numel() != B + 1 or cost.numel() != Q * int(sum(counts)) or not cost.is_contiguous():
        raise RuntimeError("lsap: cost must hold one contiguous [Q, n_b] matrix per image")
    m = int(sum(min(Q, int(n)) for n in counts))
    out = torch.empty((2, m), dtype=torch.int64, device=cost.device)
    status = torch.zeros(1, dtype=torch.int32, device=cost.device)
    if m == 0:
        return out[0], out[1], status
    with torch.cuda.device(cost.device):
        rc = _lib.load().mpf_lsap_f32(cost.data_ptr(), tgt_offsets.data_ptr(), B, Q, int(max(counts)),
                                      out[0].data_ptr(), out[1].data_ptr(), status.data_ptr(), _stream())
    _lib.check(rc, "lsap")
    return out[0], out[1], status


# ----------------------------------------------------------------------------------------------------------------
# Point sampling for the criterion (ref mask2former/modeling/criterion.py:141-191)
# ----------------------------------------------------------------------------------------------------------------
def point_sample_rows(map_ptrs, maps_are_f32, hw, coords, neg_abs=False, out=None):
    """out[r, p] = bilinear(map_r, coords[r, p]) with map_r the H x W map (uint8 or float32) at device address
    map_ptrs[r] (int64 [R]); coords [R, P, 2] f32 in [0, 1].  No autograd (see PointSampleRows).  ``out``: optional
    contiguous float32 [R, P] destination."""
    _lib.require_cuda(map_ptrs, "map_ptrs")
    _lib.require_cuda(coords, "coords")
    if map_ptrs.dtype != torch.int64 or coords.dtype != torch.float32 or coords.dim() != 3 or coords.shape[-1] != 2:
        raise RuntimeError("point_sample_rows: map_ptrs int64 [R] and coords float32 [R, P, 2] expected")
    coords = coords.contiguous()
    R, P = coords.shape[:2]
    if map_ptrs.numel() != R:
        raise RuntimeError("point_sample_rows: one map pointer per row of coords expected")
    if out is None:
        out = torch.empty((R, P), dtype=torch.float32, device=coords.device)
    elif out.dtype != torch.float32 or tuple(out.shape) != (R, P) or not out.is_contiguous() or out.device != coords.device:
        raise RuntimeError("point_sample_rows: out must be a contiguous float32 [R, P] tensor on the coords' device")
    with torch.cuda.device(coords.device):
        rc = _lib.load().mpf_point_sample_rows(map_ptrs.data_ptr(), int(bool(maps_are_f32)), int(hw[0]), int(hw[1]),
                                               coords.data_ptr(), R, P, int(bool(neg_abs)), out.data_ptr(), _stream())
    _lib.check(rc, "point_sample_rows")
    return out


class PointSampleRows(torch.autograd.Function):
    """Differentiable sampling of rows of a float32 tensor of maps: ``maps`` [..., H, W] (any leading dims, each map
    contiguous), ``row_index`` int64 [R] = flat index of the map each row samples (rows may repeat a map).  Forward
    reads the maps in place; backward returns a dense zero gradient with the samples' contributions added."""

    @staticmethod
    def forward(ctx, maps, row_index, coords):
        _lib.require_cuda(maps, "maps")
        if maps.dtype != torch.float32 or maps.dim() < 3:
            raise RuntimeError("PointSampleRows: float32 maps [..., H, W] expected")
        H, W = maps.shape[-2:]
        if maps.stride(-1) != 1 or maps.stride(-2) != W:
            raise RuntimeError("PointSampleRows: every H x W map must be contiguous")
        if row_index.dtype != torch.int64 or row_index.dim() != 1 or row_index.device != maps.device:
            raise RuntimeError("PointSampleRows: row_index must be an int64 vector on the maps' device")
        lead = maps.shape[:-2]
        # element offset of every map: flat index -> multi-index over the leading dims -> strides
        offs = torch.zeros_like(row_index)
        rem = row_index
        for size, stride in zip(reversed(lead), reversed(maps.stride()[:-2])):
            offs = offs + (rem % size) * stride
            rem = rem // size
        coords = coords.contiguous()
        ctx.save_for_backward(row_index, coords)
        ctx.maps_shape = tuple(maps.shape)
        ptrs = maps.data_ptr() + 4 * offs
        return point_sample_rows(ptrs, True, (H, W), coords)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        row_index, coords = ctx.saved_tensors
        shape = ctx.maps_shape
        H, W = shape[-2:]
        grad = torch.zeros(shape, dtype=torch.float32, device=grad_out.device)       # contiguous: map i at i * H * W
        ptrs = grad.data_ptr() + (4 * H * W) * row_index
        point_sample_rows_bwd(ptrs, (H, W), coords, grad_out)
        return grad, None, None


def point_sample_rows_bwd(grad_map_ptrs, hw, coords, grad_out):
    """grad_map_r[corner] += w_corner * grad_out[r, p] (fp32 atomics) for the zero-initialised H x W float32 maps at
    device addresses grad_map_ptrs[r]."""
    _lib.require_cuda(grad_out, "grad_out")
    grad_out = grad_out.contiguous()
    if grad_out.dtype != torch.float32 or grad_out.shape != coords.shape[:2] or not coords.is_contiguous():
        raise RuntimeError("point_sample_rows_bwd: grad_out float32 [R, P] matching contiguous coords [R, P, 2] expected")
    R, P = grad_out.shape
    with torch.cuda.device(grad_out.device):
        rc = _lib.load().mpf_point_sample_rows_bwd_f32(grad_map_ptrs.data_ptr(), int(hw[0]), int(hw[1]),
                                                       coords.data_ptr(), R, P, grad_out.data_ptr(), _stream())
    _lib.check(rc, "point_sample_rows_bwd")


class PointSampleViews(torch.autograd.Function):
    """``PointSampleRows`` over SEVERAL map tensors at once (the mask logits of all prediction heads, each
    ``[B, n_i, H, W]`` float32 with contiguous H x W maps).  Row r samples the map at device address ``src_ptrs[r]``,
    which is map ``slot[r]`` of the virtual concatenation ``[B, sum n_i, H, W]`` of the tensors along dim 1 (flat
    index ``b * sum n_i + first_i + q``; the caller derives both vectors from the same (tensor, b, q) triples -- see
    ``view_tables``).  One forward launch for all rows; the backward allocates ONE zeroed buffer of that concatenated
    shape, adds every row's contributions into it in one launch and returns its slices as the tensors' gradients -- a
    consumer that needs them side by side (ops._HeadCollector) finds them already laid out."""

    @staticmethod
    def view_tables(maps):
        """Per-tensor tables (base address, stride(0), stride(1) in bytes, first column of the concatenation) as int64
        device vectors + the concatenated width; build once per step (host -> device uploads)."""
        dev = maps[0].device
        mk = (lambda v: torch.tensor(v, dtype=torch.int64, device=dev))
        first, total = [], 0
        for m in maps:
            first.append(total)
            total += m.shape[1]
        return (mk([m.data_ptr() for m in maps]), mk([4 * m.stride(0) for m in maps]),
                mk([4 * m.stride(1) for m in maps]), mk(first), total)

    @staticmethod
    def forward(ctx, coords, src_ptrs, slot, *maps):
        B, _, H, W = maps[0].shape
        for m in maps:
            _lib.require_cuda(m, "maps")
            if m.dtype != torch.float32 or m.dim() != 4 or m.shape[0] != B or tuple(m.shape[-2:]) != (H, W) or \
                    m.stride(-1) != 1 or m.stride(-2) != W:
                raise RuntimeError("PointSampleViews: float32 maps [B, n, H, W] of one size with contiguous H x W expected")
        if src_ptrs.dtype != torch.int64 or slot.dtype != torch.int64 or src_ptrs.shape != slot.shape:
            raise RuntimeError("PointSampleViews: src_ptrs / slot must be int64 vectors of one length")
        coords = coords.contiguous()
        ctx.save_for_backward(coords, slot)
        ctx.geom = (B, H, W, [m.shape[1] for m in maps])
        return point_sample_rows(src_ptrs, True, (H, W), coords)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        coords, slot = ctx.saved_tensors
        B, H, W, sizes = ctx.geom
        grad = torch.zeros((B, sum(sizes), H, W), dtype=torch.float32, device=grad_out.device)
        point_sample_rows_bwd(grad.data_ptr() + (4 * H * W) * slot, (H, W), coords, grad_out)
        return (None, None, None) + tuple(grad.split(sizes, dim=1))


class MaskLossRows(torch.autograd.Function):
    """(bce [R], dice [R]) of sampled logits x [R, P] against sampled targets y [R, P] (ref criterion.py:25-68; the
    caller sums over rows and divides by num_masks): one launch forward, one backward (csrc/mask_loss.cu)."""

    @staticmethod
    def forward(ctx, x, y):
        _f32c(x, "x")
        _f32c(y, "y")
        if x.dim() != 2 or x.shape != y.shape:
            raise RuntimeError("MaskLossRows: x and y must be [R, P] tensors of one shape")
        x, y = x.contiguous(), y.contiguous()
        R, P = x.shape
        bce = torch.empty(R, dtype=torch.float32, device=x.device)
        dice = torch.empty_like(bce)
        stats = torch.empty((R, 2), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.load().mpf_mask_loss_rows_fwd_f32(x.data_ptr(), y.data_ptr(), R, P, bce.data_ptr(), dice.data_ptr(),
                                                        stats.data_ptr(), _stream())
        _lib.check(rc, "mask_loss_rows_fwd")
        ctx.save_for_backward(x, y, stats)
        return bce, dice

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_bce, g_dice):
        x, y, stats = ctx.saved_tensors
        R, P = x.shape
        zero = (lambda g: torch.zeros(R, dtype=torch.float32, device=x.device) if g is None else g.contiguous())
        g_bce, g_dice = zero(g_bce), zero(g_dice)
        gx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            rc = _lib.load().mpf_mask_loss_rows_bwd_f32(x.data_ptr(), y.data_ptr(), stats.data_ptr(), g_bce.data_ptr(),
                                                        g_dice.data_ptr(), R, P, gx.data_ptr(), _stream())
        _lib.check(rc, "mask_loss_rows_bwd")
        return gx, None


SELF_ATTN_MAX_Q = 320


def self_attn_fwd(qkv, mask_u8, heads):
    """qkv [B, Qt, 3E] f32, mask_u8 uint8 [Qt, Qt] or None (1 = not allowed) -> (out [B, Qt, E], lse [B, heads, Qt])."""
    _f32c(qkv, "qkv")
    B, Qt, E3 = qkv.shape
    E = E3 // 3
    if not qkv.is_contiguous() or E3 != 3 * E or E % heads or E // heads != 32 or Qt > SELF_ATTN_MAX_Q:
        raise RuntimeError(f"self_attn: unsupported geometry {tuple(qkv.shape)}, {heads} heads")
    if mask_u8 is not None and (mask_u8.dtype != torch.uint8 or mask_u8.shape != (Qt, Qt) or not mask_u8.is_contiguous()):
        raise RuntimeError("self_attn: mask must be a contiguous uint8 [Qt, Qt] tensor")
    out = torch.empty((B, Qt, E), dtype=torch.float32, device=qkv.device)
    lse = torch.empty((B, heads, Qt), dtype=torch.float32, device=qkv.device)
    with torch.cuda.device(qkv.device):
        rc = _lib.load().mpf_self_attn_fwd_f32(qkv.data_ptr(), None if mask_u8 is None else mask_u8.data_ptr(), B, Qt,
                                               heads, 32, out.data_ptr(), lse.data_ptr(), _stream())
    _lib.check(rc, "self_attn_fwd")
    return out, lse


def self_attn_bwd(qkv, mask_u8, d_out, lse, heads):
    """-> d_qkv [B, Qt, 3E]"""
    _f32c(d_out, "d_out")
    B, Qt, E3 = qkv.shape
    d_out = d_out.contiguous()
    d_qkv = torch.empty_like(qkv)
    with torch.cuda.device(qkv.device):
        rc = _lib.load().mpf_self_attn_bwd_f32(qkv.data_ptr(), None if mask_u8 is None else mask_u8.data_ptr(),
                                               d_out.data_ptr(), lse.data_ptr(), B, Qt, heads, 32, d_qkv.data_ptr(),
                                               _stream())
    _lib.check(rc, "self_attn_bwd")
    return d_qkv


TOPK_GATHER_MAX_N = 49152


def topk_gather_rows(scores, payload, k):
    """scores [R, n] f32, payload [R, n, w] f32 (w = 1 or 2) -> payload rows of the k largest scores of every row,
    [R, k, w], in ascending index order (ref criterion.py:165-172: the most uncertain candidate points)."""
    _f32c(scores, "scores")
    _f32c(payload, "payload")
    R, n = scores.shape
    w = payload.shape[-1]
    if payload.shape != (R, n, w) or w not in (1, 2) or not 0 < k <= n <= TOPK_GATHER_MAX_N:
        raise RuntimeError(f"topk_gather_rows: unsupported shapes {tuple(scores.shape)} / {tuple(payload.shape)}, k={k}")
    scores, payload = scores.contiguous(), payload.contiguous()
    out = torch.empty((R, k, w), dtype=torch.float32, device=scores.device)
    with torch.cuda.device(scores.device):
        rc = _lib.load().mpf_topk_gather_rows_f32(scores.data_ptr(), R, n, k, payload.data_ptr(), w, out.data_ptr(),
                                                  _stream())
    _lib.check(rc, "topk_gather_rows")
    return out


# ----------------------------------------------------------------------------------------------------------------
# Instance-segmentation epilogue (ref mask2former/maskformer_model.py:239-260, 365-401)
# ----------------------------------------------------------------------------------------------------------------
def instance_masks(mask_logits, query_index, padded_size, image_size, out_size, mask_dtype=torch.uint8):
    """mask_logits [Q, h, w] f32 (one image), query_index int64 [R] -> (masks [R, oh, ow] uint8 | float32 of 0/1,
    sums [R, 2] = (sum of sigmoid over the foreground, foreground pixel count)).  Both bilinear resizes of the
    reference (to the padded input size; after the crop to ``image_size``, to ``out_size``) are evaluated on the fly."""
    _lib.require_cuda(mask_logits, "mask_logits")
    _lib.require_cuda(query_index, "query_index")
    if mask_logits.dtype != torch.float32 or mask_logits.dim() != 3 or query_index.dtype != torch.int64:
        raise RuntimeError("instance_masks: mask_logits float32 [Q, h, w] and query_index int64 [R] expected")
    if mask_dtype not in (torch.uint8, torch.float32):
        raise RuntimeError("instance_masks: mask_dtype must be torch.uint8 or torch.float32")
    Q, h, w = mask_logits.shape
    if mask_logits.stride(-1) != 1 or mask_logits.stride(-2) != w:
        mask_logits = mask_logits.contiguous()
    R = int(query_index.numel())
    oh, ow = int(out_size[0]), int(out_size[1])
    lib = _lib.load()
    blocks = int(lib.mpf_instance_masks_blocks(oh, ow))
    if blocks <= 0:
        raise RuntimeError("instance_masks: bad output size")
    masks = torch.empty((R, oh, ow), dtype=mask_dtype, device=mask_logits.device)
    partial = torch.empty((R, blocks, 2), dtype=torch.float32, device=mask_logits.device)
    with torch.cuda.device(mask_logits.device):
        rc = lib.mpf_instance_masks_f32(mask_logits.data_ptr(), mask_logits.stride(0), h, w, query_index.data_ptr(), R,
                                        int(padded_size[0]), int(padded_size[1]), int(image_size[0]),
                                        int(image_size[1]), oh, ow, masks.data_ptr(),
                                        int(mask_dtype == torch.float32), partial.data_ptr(), _stream())
    _lib.check(rc, "instance_masks")
    return masks, partial.sum(1)
