"""Autograd-aware operators over the C-ABI kernels (``include/vgslm.h``).

Each function here is the host-side mirror of one reference call site (cited per function); the
math runs in libvgslm's CUDA kernels.  There is no CPU path: CPU tensors raise in ``_lib.ptr``.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Tuple

import torch

from . import _lib as L
from ._lib import ACT_GELU, ACT_MULT, ACT_NONE, ACT_RELU, ACT_SILU, GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05

# GEMM backend used for bf16 operands: AUTO picks tcgen05 when the shape qualifies.
GEMM_BACKEND = GEMM_AUTO
# when a list, every bf16 GEMM launch appends (start_event, end_event, flops, dtype) — bench.py's roofline probe
PROFILE = None

ACT_IDS = {"none": ACT_NONE, None: ACT_NONE, "relu": ACT_RELU, "gelu": ACT_GELU, "silu": ACT_SILU,
           "ReLU": ACT_RELU, "GELU": ACT_GELU, "SiLU": ACT_SILU}


def _u8(mask: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """bool [B,T] (or [M]) validity mask → flat uint8 view (no copy)."""
    if mask is None:
        return None
    if mask.dtype == torch.bool:
        mask = mask.view(torch.uint8)
    return mask.reshape(-1)


def _rows2d(x: torch.Tensor) -> torch.Tensor:
    x2 = x.reshape(-1, x.shape[-1])
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    return x2


# ------------------------------------------------------------------------------------------- cast
_shadow_cache = {}


def lowp(weight: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Compute-dtype copy of an fp32 master weight, refreshed when the master changes.

    (The arena path in ``arena.py`` writes the bf16 shadow inside the fused AdamW kernel instead.)"""
    if weight.dtype == dtype:
        return weight.detach()
    shadow = getattr(weight, "_vg_shadow", None)
    if shadow is not None and shadow.dtype == dtype:
        return shadow
    # memoised ON the weight tensor (dies with it — an address-keyed cache could alias a freed tensor's entry)
    ent = getattr(weight, "_vg_lowp", None)
    ver = weight._version
    if ent is not None and ent[0] == ver and ent[1].dtype == dtype and ent[1].device == weight.device:
        return ent[1]
    src = weight.detach().contiguous()
    out = torch.empty(src.shape, dtype=dtype, device=src.device)
    L.call("vg_cast_f32_to_bf16", L.ptr(src), L.ptr(out), src.numel(), L.stream())
    try:
        weight._vg_lowp = (ver, out)
    except Exception:
        pass
    return out


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    src = x.detach().contiguous()
    out = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    L.call("vg_cast_f32_to_bf16", L.ptr(src), L.ptr(out), src.numel(), L.stream())
    return out


# ------------------------------------------------------------------------------------------- GEMM
def gemm(a: torch.Tensor, b: torch.Tensor, *, trans_a: bool = False, trans_b: bool = True,
         out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None,
         bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
         preact: Optional[torch.Tensor] = None, dact_src: Optional[torch.Tensor] = None, dact: int = ACT_NONE,
         residual: Optional[torch.Tensor] = None, row_mask: Optional[torch.Tensor] = None,
         mask_first: bool = False, beta: float = 0.0, backend: Optional[int] = None,
         preact_is_grad: bool = False) -> torch.Tensor:
    """C = epilogue(op(A)·op(B)); see vg_gemm in include/vgslm.h.  ``a``/``b`` are 2-D with unit inner stride."""
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    if trans_a:
        K, M = a.shape
    else:
        M, K = a.shape
    if trans_b:
        N, Kb = b.shape
    else:
        Kb, N = b.shape
    assert K == Kb, f"inner dimensions differ: {K} vs {Kb}"
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype or a.dtype, device=a.device)
    assert out.stride(1) == 1 and out.shape == (M, N)
    g = L.GemmArgs()
    g.M, g.N, g.K = M, N, K
    g.A, g.lda, g.trans_a = L.ptr(a), a.stride(0), int(trans_a)
    g.B, g.ldb, g.trans_b = L.ptr(b), b.stride(0), int(trans_b)
    g.C, g.ldc = L.ptr(out), out.stride(0)
    g.ab_dtype, g.c_dtype = L.dtype_id(a.dtype), L.dtype_id(out.dtype)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
        g.bias = L.ptr(bias)
    g.act, g.dact = act, dact
    for name, t, ld in (("preact", preact, "ld_preact"), ("dact_src", dact_src, "ld_dact"),
                        ("residual", residual, "ld_res")):
        if t is not None:
            assert t.dtype == out.dtype and t.shape == (M, N) and t.stride(1) == 1, name
            setattr(g, name, L.ptr(t))
            setattr(g, ld, t.stride(0))
    if row_mask is not None:
        assert row_mask.numel() == M
        g.row_mask = L.ptr(row_mask)
    g.mask_before_residual = int(mask_first)
    g.preact_is_grad = int(preact_is_grad)
    g.beta = beta
    be = GEMM_BACKEND if backend is None else backend
    prof = PROFILE if (PROFILE is not None and a.dtype == torch.bfloat16) else None
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if PROFILE_SPIN:
            # keep the device BEHIND the host: a short spin kernel in front of the start event gives the host time to
            # enqueue the GEMM, so that the event pair brackets the kernel and not the idle gap before its launch arrives
            torch.cuda._sleep(PROFILE_SPIN)
        e0.record()
    L.call("vg_gemm", C.byref(g), be, None, 0, L.stream())
    if prof is not None:
        e1.record()
        prof.append((e0, e1, 2.0 * M * N * K, a.dtype, (M, N, K, int(trans_a), int(trans_b), str(out.dtype)[6:], act, bias is not None)))
    return out


def colsum(x: torch.Tensor, out: Optional[torch.Tensor] = None, beta: float = 0.0) -> torch.Tensor:
    """out[n] = beta·out[n] + Σ_m x[m,n] (bias gradients)."""
    rows, cols = x.shape
    if out is None:
        assert beta == 0.0
        out = torch.empty(cols, dtype=torch.float32, device=x.device)
    assert out.dtype == torch.float32 and out.numel() == cols and out.is_contiguous()
    ws = L.workspace(L.load().vg_colsum_workspace(rows, cols), x.device)
    L.call("vg_colsum", L.ptr(x), x.stride(0), L.ptr(out), rows, cols, L.dtype_id(x.dtype), float(beta), L.ptr(ws),
           ws.numel(), L.stream())
    return out


# Parameter gradients (wgrad GEMMs, bias column sums) are off the critical path of backward: nothing downstream reads
# them before the optimizer.  With GRAD_STREAM set (TrainStep does) they are enqueued on that side stream — in a captured
# graph a parallel branch — so the memory-bound column sums run under the compute-bound dgrad GEMMs and the wgrad GEMMs
# fill the tail waves of the dgrad GEMMs.  The owner of the stream joins it before the optimizer / all-reduce.
GRAD_STREAM = None


class _on_grad_stream:
    def __init__(self, *tensors):
        self.tensors = tensors
        self.ctx = None

    def __enter__(self):
        gs = GRAD_STREAM
        if gs is None or not self.tensors[0].is_cuda:
            return self
        gs.wait_stream(torch.cuda.current_stream())
        for t in self.tensors:
            t.record_stream(gs)
        self.ctx = torch.cuda.stream(gs)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def _bgrad(dy2: torch.Tensor, bias: torch.Tensor) -> Optional[torch.Tensor]:
    """db = Σ_rows dy.  A bias that lives in a ParamArena gets the sum written (or accumulated, on later
    micro-batches) straight into its gradient slice and autograd receives None — no AccumulateGrad add kernel."""
    main = getattr(bias, "_vg_main_grad", None)
    if main is not None:
        arena = bias._vg_arena
        with _on_grad_stream(dy2):
            colsum(dy2, out=main, beta=arena.wgrad_beta(bias))
            arena.grad_ready(bias)
        return None
    return colsum(dy2)


def mask_rows_(x2: torch.Tensor, mask_u8: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = where(mask[row], x, 0) on a contiguous 2-D tensor (utils/tensormask.py:63-67)."""
    assert x2.is_contiguous()
    out = torch.empty_like(x2) if out is None else out
    L.call("vg_mask_rows", L.ptr(x2), L.ptr(mask_u8), L.ptr(out), x2.shape[0], x2.shape[1], L.dtype_id(x2.dtype),
           L.stream())
    return out


def act_bwd(dy: torch.Tensor, src: torch.Tensor, act: int) -> torch.Tensor:
    assert dy.is_contiguous() and src.is_contiguous() and dy.dtype == src.dtype
    out = torch.empty_like(dy)
    L.call("vg_act_bwd", L.ptr(dy), L.ptr(src), L.ptr(out), dy.numel(), act, L.dtype_id(dy.dtype), L.stream())
    return out


class _MaskRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mask_u8):
        ctx.save_for_backward(mask_u8)
        ctx.shape = x.shape
        return mask_rows_(_rows2d(x).contiguous(), mask_u8).view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        (mask_u8,) = ctx.saved_tensors
        return mask_rows_(_rows2d(dy).contiguous(), mask_u8).view(ctx.shape), None


def mask_rows(x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """TensorMask.apply_mask for [B,T,C] activations with C % 8 == 0."""
    return _MaskRows.apply(x, _u8(mask))


# ---------------------------------------------------------------------------------------- RMSNorm
class _RMSNorm(torch.autograd.Function):
    """modules/norm.py:22-32 (+ the mask of transformer/layers.py:53-54).

    With ``passthrough`` the function also returns its input (as an alias): a pre-LN block uses that alias as the
    residual branch, so the two gradients of x — through the norm and around it — meet HERE and the backward kernel
    adds them (``dres``) instead of autograd launching a separate add (transformer/layers.py:57,63)."""

    @staticmethod
    def forward(ctx, x, scale, eps, mask_u8, out_dtype, passthrough):
        ctx.set_materialize_grads(False)
        x2 = _rows2d(x).contiguous()
        rows, dim = x2.shape
        y = torch.empty((rows, dim), dtype=out_dtype or x.dtype, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        sc = scale.detach().float().contiguous()
        L.call("vg_rmsnorm_fwd", L.ptr(x2), L.ptr(sc), L.ptr(mask_u8), L.ptr(y), L.ptr(rstd), rows, dim, float(eps),
               L.dtype_id(x2.dtype), L.dtype_id(y.dtype), L.stream())
        ctx.save_for_backward(x2, sc, rstd, mask_u8)
        ctx.xshape = x.shape
        ctx.scale_param = scale
        ctx.passthrough = passthrough
        if passthrough:
            return x.view_as(x), y.view(x.shape)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, *grads):
        x2, sc, rstd, mask_u8 = ctx.saved_tensors
        dres, dy = (grads[0], grads[1]) if ctx.passthrough else (None, grads[0])
        if dy is None:                      # the normalised branch is unused: only the residual gradient flows
            return dres, None, None, None, None, None
        rows, dim = x2.shape
        dy2 = _rows2d(dy).contiguous()
        dres2 = None
        if dres is not None:
            dres2 = _rows2d(dres).contiguous()
            if dres2.dtype != x2.dtype:
                dres2 = dres2.to(x2.dtype)
        dx = torch.empty_like(x2)
        scale = ctx.scale_param
        main = getattr(scale, "_vg_main_grad", None)
        if main is not None:
            dscale, beta = main, scale._vg_arena.wgrad_beta(scale)
        else:
            dscale, beta = torch.empty(dim, dtype=torch.float32, device=x2.device), 0.0
        ws = L.workspace(L.load().vg_rmsnorm_bwd_workspace(rows, dim), x2.device)
        L.call("vg_rmsnorm_bwd", L.ptr(dy2), L.ptr(x2), L.ptr(sc), L.ptr(rstd), L.ptr(mask_u8), L.ptr(dres2),
               L.ptr(dx), L.ptr(dscale), float(beta), L.ptr(ws), ws.numel(), rows, dim, L.dtype_id(x2.dtype),
               L.dtype_id(dy2.dtype), L.stream())
        if main is not None:
            scale._vg_arena.grad_ready(scale)
            dscale = None
        return dx.view(ctx.xshape), dscale, None, None, None, None


def rmsnorm(x: torch.Tensor, scale: torch.Tensor, eps: float, mask: Optional[torch.Tensor] = None,
            out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    return _RMSNorm.apply(x, scale, eps, _u8(mask), out_dtype, False)


def rmsnorm_residual(x: torch.Tensor, scale: torch.Tensor, eps: float, mask: Optional[torch.Tensor] = None,
                     out_dtype: Optional[torch.dtype] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(x, RMSNorm(x)): use the returned x as the residual input of the block that consumes the norm."""
    return _RMSNorm.apply(x, scale, eps, _u8(mask), out_dtype, True)


# ------------------------------------------------------------- skinny linear (cached generation step, <= 256 rows)
PROFILE_SPIN = 0            # device clock cycles of spin in front of every profiled GEMM (bench.py's instrumented pass)
SKINNY_LINEAR = os.environ.get("VG_SKINNY_LINEAR", "1") != "0"
SKINNY_MAX_ROWS = int(os.environ.get("VG_SKINNY_MAX_ROWS", "64"))      # above: gemm_tc's skinny-M plan (profiles/r02_decode.md)
# up to this many rows the layers with <= 1024 output features (out-projection, W2 of the FFN: long k-range, 8..16 output
# tiles in the general GEMM) still go to vg_skinny_linear, whose cluster splits the k-range
SKINNY_MIXED_ROWS = int(os.environ.get("VG_SKINNY_MIXED_ROWS", "256"))


def _skinny_ok(x2: torch.Tensor, w: torch.Tensor) -> bool:
    """bf16 linear layers on few rows (one new frame per sequence) whose inputs need no gradient go to vg_skinny_linear
    (autograd.Function.forward runs with grad mode off even in training: the callers test ctx.needs_input_grad)"""
    return (SKINNY_LINEAR and x2.is_cuda and x2.dtype == torch.bfloat16
            and w.dtype == torch.bfloat16 and x2.shape[1] % 64 == 0 and w.shape[0] >= 8
            and (x2.shape[0] <= SKINNY_MAX_ROWS or (x2.shape[0] <= SKINNY_MIXED_ROWS and w.shape[0] <= 1024))
            and x2.stride(0) % 8 == 0 and x2.stride(1) == 1 and w.is_contiguous())


@torch.no_grad()
def skinny_linear(x2: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
                  residual: Optional[torch.Tensor] = None, row_mask: Optional[torch.Tensor] = None,
                  mask_first: bool = False, out_dtype: Optional[torch.dtype] = None, *,
                  out: Optional[torch.Tensor] = None, x_ss: Optional[torch.Tensor] = None, norm_eps: float = 0.0,
                  y_ss: Optional[torch.Tensor] = None, zero_ss: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = epilogue(x2 · wᵀ): x2 [B <= 256, K] bf16, w [N, K] bf16 (csrc/skinny_linear.cu: swap-AB tcgen05, cluster
    split-K through distributed shared memory).  With ``x_ss`` the RMSNorm in front is folded in (w pre-multiplied by the
    norm's scale, 1/rms applied per row); ``y_ss`` accumulates the row sums of squares of the result, ``zero_ss`` is cleared."""
    B, K = x2.shape
    N = w.shape[0]
    odt = out_dtype or (out.dtype if out is not None else x2.dtype)
    y = torch.empty((B, N), dtype=odt, device=x2.device) if out is None else out
    assert y.shape == (B, N) and y.dtype == odt and y.stride(1) == 1
    a = L.SkinnyLinearArgs()
    if x_ss is not None:        # RMSNorm folded in: ``w`` already carries the norm's scale vector, x_ss = sum_k x[b,k]^2
        assert x_ss.dtype == torch.float32 and x_ss.numel() >= B
        a.row_ss_in, a.ss_inv_k, a.ss_eps = L.ptr(x_ss), 1.0 / K, float(norm_eps)
    if y_ss is not None:
        assert y_ss.dtype == torch.float32 and y_ss.numel() >= B
        a.row_ss_out = L.ptr(y_ss)
    if zero_ss is not None:
        assert zero_ss.dtype == torch.float32 and zero_ss.numel() >= B
        a.zero_ss = L.ptr(zero_ss)
    a.x, a.ldx, a.w, a.ldw, a.y, a.ldy = L.ptr(x2), x2.stride(0), L.ptr(w), w.stride(0), L.ptr(y), y.stride(0)
    a.B, a.N, a.K = B, N, K
    a.y_dtype, a.act, a.mask_before_residual = L.dtype_id(odt), act, int(mask_first)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
        a.bias = L.ptr(bias)
    if residual is not None:
        assert residual.dtype == odt and residual.shape == (B, N) and residual.stride(1) == 1
        a.residual, a.ld_res = L.ptr(residual), residual.stride(0)
    if row_mask is not None:
        assert row_mask.numel() == B
        a.row_mask = L.ptr(row_mask)
    L.call("vg_skinny_linear", C.byref(a), L.stream())
    return y


# ----------------------------------------------------------------------------------------- Linear
def _wgrad(dy2: torch.Tensor, x2: torch.Tensor, weight: torch.Tensor) -> Optional[torch.Tensor]:
    """dW = dyᵀ·x in fp32.  Parameters that live in a ParamArena (arena.py) get the GEMM written straight into
    their gradient slice (beta = 1 on later micro-batches) and autograd receives None."""
    main = getattr(weight, "_vg_main_grad", None)
    if main is not None:
        arena = weight._vg_arena
        with _on_grad_stream(dy2, x2):
            gemm(dy2, x2, trans_a=True, trans_b=False, out=main.view(main.shape[0], -1), beta=arena.wgrad_beta(weight))
            arena.grad_ready(weight)
        return None
    dw = gemm(dy2, x2, trans_a=True, trans_b=False, out_dtype=torch.float32).view(weight.shape)   # 1x1 conv: [N,K,1]
    return dw if weight.dtype == torch.float32 else dw.to(weight.dtype)


class _Linear(torch.autograd.Function):
    """y = mask/residual epilogue(act(x·Wᵀ + b)) — one GEMM launch; backward = dgrad + wgrad (+ colsum)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, mask_u8, act, mask_first, out_dtype, w_compute):
        x2 = _rows2d(x)
        w = w_compute if w_compute is not None else lowp(weight, x2.dtype)
        if w.dim() == 3:                       # kernel-size-1 Conv1d weight [N,K,1] used as a linear layer
            w = w.view(w.shape[0], -1)
        N = w.shape[0]
        odt = out_dtype or x2.dtype
        res2 = _rows2d(residual) if residual is not None else None
        b = bias.detach().float() if bias is not None else None
        if not any(ctx.needs_input_grad) and _skinny_ok(x2, w) and (res2 is None or res2.dtype == odt):
            y = skinny_linear(x2, w, b, act, res2, mask_u8, mask_first, odt)
            return y.view(*x.shape[:-1], N)
        # ReLU's derivative can be read off the output; otherwise the epilogue stores act'(pre) next to act(pre)
        need_pre = act != ACT_NONE and (act != ACT_RELU or residual is not None or mask_u8 is not None)
        pre = torch.empty((x2.shape[0], N), dtype=odt, device=x.device) if need_pre else None
        y = gemm(x2, w, trans_b=True, out_dtype=odt, bias=b, act=act, preact=pre, preact_is_grad=True, residual=res2,
                 row_mask=mask_u8, mask_first=mask_first)
        ctx.save_for_backward(x2, w, pre if need_pre else (y if act != ACT_NONE else None), mask_u8)
        ctx.act_bwd = ACT_MULT if need_pre else act
        ctx.meta = (act, mask_first, x.shape, weight, bias, residual is not None,
                    residual.shape if residual is not None else None)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w, act_src, mask_u8 = ctx.saved_tensors
        act, mask_first, xshape, weight, bias, has_res, res_shape = ctx.meta
        dy2 = _rows2d(dy).contiguous()
        g = mask_rows_(dy2, mask_u8) if mask_u8 is not None else dy2
        dres = None
        if has_res and ctx.needs_input_grad[3]:
            dres = (dy2 if mask_first else g).view(res_shape)
        dpre = act_bwd(g, act_src, ctx.act_bwd) if act != ACT_NONE else g
        if dpre.dtype != x2.dtype:           # fp32 head outputs of bf16 GEMMs
            dpre = dpre.to(x2.dtype)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = gemm(dpre, w, trans_b=False).view(xshape)
        if ctx.needs_input_grad[1]:
            dw = _wgrad(dpre, x2, weight)
        if bias is not None and ctx.needs_input_grad[2]:
            db = _bgrad(dpre, bias)
        return dx, dw, db, dres, None, None, None, None, None


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
           act: int = ACT_NONE, residual: Optional[torch.Tensor] = None, row_mask: Optional[torch.Tensor] = None,
           mask_before_residual: bool = False, out_dtype: Optional[torch.dtype] = None,
           w_compute: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.Linear (+ fused activation / residual / TensorMask.apply_mask) on the last dimension of ``x``."""
    return _Linear.apply(x, weight, bias, residual, _u8(row_mask), act, mask_before_residual, out_dtype, w_compute)


class _FFN(torch.autograd.Function):
    """transformer/layers.py:82-86: out = mask(res + W2·act(W1·x + b1) + b2).

    Forward: GEMM1 stores the pre-activation and the activation from one epilogue; GEMM2 fuses bias,
    residual and row mask.  Backward: the GELU derivative rides in the epilogue of GEMM2's dgrad."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, residual, mask_u8, act):
        x2 = _rows2d(x)
        w1c, w2c = lowp(w1, x2.dtype), lowp(w2, x2.dtype)
        M, F = x2.shape[0], w1c.shape[0]
        if not any(ctx.needs_input_grad) and _skinny_ok(x2, w1c) and _skinny_ok(x2, w2c):
            h = skinny_linear(x2, w1c, b1.detach().float() if b1 is not None else None, act)
            y = skinny_linear(h, w2c, b2.detach().float() if b2 is not None else None, ACT_NONE,
                              _rows2d(residual) if residual is not None else None, mask_u8)
            return y.view(x.shape[:-1] + (w2c.shape[0],))
        if not any(ctx.needs_input_grad) and x2.dtype == torch.bfloat16 and x2.is_cuda:
            h = gemm(x2, w1c, trans_b=True, bias=b1.detach().float() if b1 is not None else None, act=act)
            if _skinny_ok(h, w2c):            # SKINNY_MIXED_ROWS: W2's k-range split inside a cluster
                y = skinny_linear(h, w2c, b2.detach().float() if b2 is not None else None, ACT_NONE,
                                  _rows2d(residual) if residual is not None else None, mask_u8)
                return y.view(x.shape[:-1] + (w2c.shape[0],))
            y = gemm(h, w2c, trans_b=True, bias=b2.detach().float() if b2 is not None else None,
                     residual=_rows2d(residual) if residual is not None else None, row_mask=mask_u8)
            return y.view(x.shape[:-1] + (w2c.shape[0],))
        pre = torch.empty((M, F), dtype=x2.dtype, device=x.device)
        h = gemm(x2, w1c, trans_b=True, bias=b1.detach().float() if b1 is not None else None, act=act, preact=pre,
                 preact_is_grad=True)          # `pre` holds act'(W1·x + b1): backward is a plain multiply
        res2 = _rows2d(residual) if residual is not None else None
        y = gemm(h, w2c, trans_b=True, bias=b2.detach().float() if b2 is not None else None, residual=res2,
                 row_mask=mask_u8)
        ctx.save_for_backward(x2, w1c, w2c, pre, h, mask_u8)
        ctx.meta = (act, x.shape, b1, b2, residual is not None, w1, w2)
        return y.view(x.shape[:-1] + (w2c.shape[0],))

    @staticmethod
    def backward(ctx, dy):
        x2, w1c, w2c, pre, h, mask_u8 = ctx.saved_tensors
        act, xshape, b1, b2, has_res, w1, w2 = ctx.meta
        dy2 = _rows2d(dy).contiguous()
        g = mask_rows_(dy2, mask_u8) if mask_u8 is not None else dy2
        dres = g.view(dy.shape) if has_res and ctx.needs_input_grad[5] else None
        dpre = gemm(g, w2c, trans_b=False, dact_src=pre, dact=ACT_MULT)       # (g·W2) ⊙ act'(pre), saved by forward
        dw2 = _wgrad(g, h, w2) if ctx.needs_input_grad[3] else None
        db2 = _bgrad(g, b2) if b2 is not None and ctx.needs_input_grad[4] else None
        dx = gemm(dpre, w1c, trans_b=False).view(xshape) if ctx.needs_input_grad[0] else None
        dw1 = _wgrad(dpre, x2, w1) if ctx.needs_input_grad[1] else None
        db1 = _bgrad(dpre, b1) if b1 is not None and ctx.needs_input_grad[2] else None
        return dx, dw1, db1, dw2, db2, dres, None, None


def ffn(x, w1, b1, w2, b2, residual=None, row_mask=None, act: int = ACT_GELU) -> torch.Tensor:
    return _FFN.apply(x, w1, b1, w2, b2, residual, _u8(row_mask), act)


# ------------------------------------------------------------------------- depthwise conv + channel LN
class _DwConvLN(torch.autograd.Function):
    """norm(conv1(x) [+ time_emb]) of a ResidualBlock on [B,T,C] rows (conv/layers.py:117-135,238-253; norm.py:43-47)."""

    @staticmethod
    def forward(ctx, x, conv_w, conv_b, t_add, ln_w, ln_b, pad_left, eps, out_cols):
        B, T, Cc = x.shape
        xc = x.contiguous()
        if conv_w is not None:
            taps = conv_w.shape[-1]
            w_t = conv_w.detach().reshape(Cc, taps).t().contiguous().float()         # [taps][C]
        else:
            taps, w_t = 1, None
        cb = _f32c(conv_b)
        ta = _f32c(t_add)
        lw, lb = _f32c(ln_w), _f32c(ln_b)
        ld_y = out_cols or Cc                 # the caller may ask for a wider row (room for concatenated conditions)
        y = torch.empty((B, T, ld_y), dtype=x.dtype, device=x.device)
        mean = torch.empty(B * T, dtype=torch.float32, device=x.device)
        rstd = torch.empty(B * T, dtype=torch.float32, device=x.device)
        L.call("vg_dwconv_ln_fwd", L.ptr(xc), L.ptr(w_t), L.ptr(cb), L.ptr(ta), L.ptr(lw), L.ptr(lb), L.ptr(y), ld_y,
               L.ptr(mean), L.ptr(rstd), B, T, Cc, taps, int(pad_left), float(eps), L.dtype_id(x.dtype), L.stream())
        ctx.save_for_backward(xc, w_t, cb, ta, lw, mean, rstd)
        ctx.meta = (taps, int(pad_left), conv_w.shape if conv_w is not None else None, conv_b is not None,
                    t_add is not None, ld_y)
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, w_t, cb, ta, lw, mean, rstd = ctx.saved_tensors
        taps, pad_left, w_shape, has_b, has_t, ld_y = ctx.meta
        B, T, Cc = xc.shape
        dyc = dy.contiguous()                                  # [B,T,ld_y]; only the first C columns are ours
        dev = xc.device
        dh = torch.empty_like(xc)
        dx = torch.empty_like(xc)
        f32 = dict(dtype=torch.float32, device=dev)
        dw_t = torch.empty((taps, Cc), **f32) if w_t is not None else None
        dlw, dlb = torch.empty(Cc, **f32), torch.empty(Cc, **f32)
        db = torch.empty(Cc, **f32) if has_b else None
        dt = torch.empty((B, Cc), **f32) if has_t else None          # Σ_t dh, from the strip kernel's partial sums
        ws = L.workspace(L.load().vg_dwconv_ln_bwd_workspace(B, T, Cc, taps), dev)
        L.call("vg_dwconv_ln_bwd", L.ptr(dyc), ld_y, L.ptr(xc), L.ptr(w_t), L.ptr(cb), L.ptr(ta), L.ptr(lw),
               L.ptr(mean), L.ptr(rstd), L.ptr(dh), L.ptr(dx), L.ptr(dw_t), L.ptr(dlw), L.ptr(dlb), L.ptr(db),
               L.ptr(dt), L.ptr(ws), ws.numel(), B, T, Cc, taps, pad_left, L.dtype_id(xc.dtype), L.stream())
        dconv_w = dw_t.t().reshape(w_shape) if w_t is not None else None
        return dx, dconv_w, db, dt, dlw, dlb, None, None, None


def dwconv_ln(x: torch.Tensor, conv_w: Optional[torch.Tensor], conv_b: Optional[torch.Tensor],
              t_add: Optional[torch.Tensor], ln_w: torch.Tensor, ln_b: torch.Tensor, pad_left: int, eps: float,
              out_cols: Optional[int] = None) -> torch.Tensor:
    """x [B,T,C] → LayerNorm_C(depthwise_conv(x) + bias + t_add[:,None,:]); conv_w [C,1,k] or None (identity).
    With ``out_cols`` > C the result has that many columns and only [:C] is written."""
    return _DwConvLN.apply(x, conv_w, conv_b, t_add, ln_w, ln_b, pad_left, eps, out_cols)


# ------------------------------------------------------------------ strided Conv1d as a GEMM (utterance encoder)
class _Im2col(torch.autograd.Function):
    """a[b,t,c*K+j] = f(x[b, S*t - pad + j, c]) on [B,T,C] rows, f = ReLU (of the previous ConvNormAct) or identity
    (conv/layers.py:549-593).  The column order is the memory order of the Conv1d weight [Cout,Cin,K]."""

    @staticmethod
    def forward(ctx, x, K, S, pad, relu):
        B, T, Cc = x.shape
        xc = x.contiguous()
        Tout = (T + 2 * pad - K) // S + 1
        a = torch.empty((B, Tout, Cc * K), dtype=x.dtype, device=x.device)
        L.call("vg_im2col_fwd", L.ptr(xc), L.ptr(a), B, T, Cc, K, S, pad, Tout, int(relu), L.dtype_id(x.dtype), L.stream())
        ctx.save_for_backward(xc if relu else None)
        ctx.meta = (B, T, Cc, K, S, pad, Tout)
        return a

    @staticmethod
    def backward(ctx, da):
        (xr,) = ctx.saved_tensors
        B, T, Cc, K, S, pad, Tout = ctx.meta
        dac = da.contiguous()
        dx = torch.empty((B, T, Cc), dtype=da.dtype, device=da.device)
        L.call("vg_im2col_bwd", L.ptr(dac), L.ptr(xr), L.ptr(dx), B, T, Cc, K, S, pad, Tout, L.dtype_id(da.dtype), L.stream())
        return dx, None, None, None, None


def im2col(x: torch.Tensor, kernel_size: int, stride: int, pad: int, relu: bool = False) -> torch.Tensor:
    """[B,T,C] → [B,T_out,C*K] window gather (optionally of relu(x)); ``im2col(x, 1, 1, 0, relu=True)`` is relu(x)."""
    return _Im2col.apply(x, int(kernel_size), int(stride), int(pad), bool(relu))


# -------------------------------------------------------------------------------------- attention
def alibi_slopes(nheads: int) -> list:
    """Head slopes of position/alibi.py:19-30 (geometric sequence; interleaved for non powers of two)."""
    def pow2(n):
        start = 2.0 ** (-(2.0 ** -(math.log2(n) - 3)))
        return [start * (start ** i) for i in range(n)]
    if math.log2(nheads).is_integer():
        return pow2(nheads)
    c = 2 ** math.floor(math.log2(nheads))
    return pow2(c) + alibi_slopes(2 * c)[0::2][: nheads - c]


class _Attention(torch.autograd.Function):
    """attention.py:52-78 — causal softmax(QKᵀ/√D + ALiBi) V with key-padding, packed qkv in/out."""

    @staticmethod
    def forward(ctx, qkv, kv_len, slopes, nheads, scale):
        B, T, C3 = qkv.shape
        HD = C3 // 3
        D = HD // nheads
        qkv = qkv.contiguous()
        q, k, v = qkv[..., :HD], qkv[..., HD:2 * HD], qkv[..., 2 * HD:]
        out = torch.empty((B, T, HD), dtype=qkv.dtype, device=qkv.device)
        lse = torch.empty((B, nheads, T), dtype=torch.float32, device=qkv.device)
        L.call("vg_attn_fwd", L.ptr(q), L.ptr(k), L.ptr(v), C3, C3, L.ptr(out), HD, L.ptr(lse), L.ptr(kv_len),
               L.ptr(slopes), B, nheads, T, T, D, 0, 0, 0, float(scale), L.dtype_id(qkv.dtype), L.stream())
        ctx.save_for_backward(qkv, out, lse, kv_len, slopes)
        ctx.meta = (nheads, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse, kv_len, slopes = ctx.saved_tensors
        nheads, scale = ctx.meta
        B, T, C3 = qkv.shape
        HD = C3 // 3
        D = HD // nheads
        dout = dout.contiguous()
        dqkv = torch.empty_like(qkv)
        q, k, v = qkv[..., :HD], qkv[..., HD:2 * HD], qkv[..., 2 * HD:]
        dq, dk, dv = dqkv[..., :HD], dqkv[..., HD:2 * HD], dqkv[..., 2 * HD:]
        ws = L.workspace(L.load().vg_attn_bwd_workspace(B, nheads, T, T, D), qkv.device)
        L.call("vg_attn_bwd", L.ptr(dout), HD, L.ptr(q), L.ptr(k), L.ptr(v), C3, C3, L.ptr(out), HD, L.ptr(lse),
               L.ptr(dq), L.ptr(dk), L.ptr(dv), C3, C3, L.ptr(kv_len), L.ptr(slopes), B, nheads, T, T, D, 0,
               float(scale), L.dtype_id(qkv.dtype), L.ptr(ws), ws.numel(), L.stream())
        return dqkv, None, None, None, None


def attention(qkv: torch.Tensor, nheads: int, kv_len: Optional[torch.Tensor] = None,
              slopes: Optional[torch.Tensor] = None, scale: Optional[float] = None) -> torch.Tensor:
    """Self-attention over a packed [B,T,3·H·D] projection; kv_len int32 [B] = valid (right-padded) lengths."""
    HD = qkv.shape[-1] // 3
    scale = scale if scale is not None else 1.0 / math.sqrt(HD // nheads)
    return _Attention.apply(qkv, kv_len, slopes, nheads, scale)


def attention_cached(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, nheads: int, q_offset: int,
                     slopes: Optional[torch.Tensor], scale: Optional[float] = None,
                     head_major: bool = False, tk: Optional[int] = None) -> torch.Tensor:
    """Inference attention of Tq new queries against Tk keys (q_offset = Tk − Tq), no grad.

    ``k``/``v`` are [B,Tk,H·D] views (packed or separate) or, with ``head_major``, a [B,H,Tmax,D] cache
    of which the first ``tk`` rows are attended."""
    B, Tq, HD = q.shape
    D = HD // nheads
    scale = scale if scale is not None else 1.0 / math.sqrt(D)
    assert q.stride(2) == 1 and q.stride(0) == Tq * q.stride(1)
    if head_major:
        assert k.is_contiguous() and v.is_contiguous() and k.shape == v.shape and k.shape[1] == nheads
        Tk, ld_kv, bs, hs = int(tk), D, k.stride(0), k.stride(1)
    else:
        Tk, ld_kv = k.shape[1], k.stride(1)
        assert k.stride(2) == 1 and v.stride(2) == 1 and v.stride(1) == ld_kv and k.stride(0) == v.stride(0)
        bs, hs = k.stride(0), D
    out = torch.empty((B, Tq, HD), dtype=q.dtype, device=q.device)
    lse = torch.empty((B, nheads, Tq), dtype=torch.float32, device=q.device)
    L.call("vg_attn_fwd", L.ptr(q), L.ptr(k), L.ptr(v), q.stride(1), ld_kv, L.ptr(out), HD, L.ptr(lse), None,
           L.ptr(slopes), B, nheads, Tq, Tk, D, q_offset, bs, hs, float(scale), L.dtype_id(q.dtype), L.stream())
    return out


@torch.no_grad()
def kv_append(k: torch.Tensor, v: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, pos: int) -> None:
    """copy new K/V rows [B,T,H·D] (views of a packed projection) into the head-major cache at ``pos``."""
    B, T, HD = k.shape
    _, H, Tmax, D = k_cache.shape
    assert k.stride(2) == 1 and v.stride(1) == k.stride(1) and k.stride(0) == T * k.stride(1)
    L.call("vg_kv_append", L.ptr(k), L.ptr(v), k.stride(1), L.ptr(k_cache), L.ptr(v_cache), B, H, D, T, Tmax, pos,
           L.dtype_id(k.dtype), L.stream())


class pdl_mode:
    """``with ops.pdl_mode(3):`` — vg_gemm / vg_rmsnorm_fwd launched inside run as a programmatic-dependent-launch chain
    (include/vgslm.h ``vg_set_pdl_mode``): bit 0 overlaps each kernel's prologue with the tail of the one in front, bit 1
    additionally declares every GEMM B operand a static weight (prefetched before the dependency wait).  For inference
    chains of short kernels only (the layer-by-layer cached generation step); process-wide, restored on exit."""
    _current = 0

    def __init__(self, mode: int) -> None:
        self.mode = mode

    def __enter__(self):
        self.prev = pdl_mode._current
        L.call("vg_set_pdl_mode", self.mode)
        pdl_mode._current = self.mode
        return self

    def __exit__(self, *exc):
        L.call("vg_set_pdl_mode", self.prev)
        pdl_mode._current = self.prev
        return False


def decode_splits(n_bh: int, horizon: int, slots: int = 3 * 148) -> int:
    """kv-splits per (sequence, head) for ``vg_attn_decode``.  Its persistent CTAs (3 per SM) each take a contiguous run of
    (b, h, split) items; a split costs a partial write, a ticket and a merge by the last arriver (~3 us per item on the
    critical path of a CTA, profiles/r02_decode.md), which is more than the imbalance it removes at every cache length of
    the generation recipe (<= 650 keys) — so splits are only used when a handful of long items would leave the GPU idle."""
    if n_bh >= slots // 2 or horizon <= 1024:
        return 1
    return max(1, min(16, slots // n_bh, horizon // 512))


@torch.no_grad()
def attention_decode(qkv: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, pos: int,
                     slopes: Optional[torch.Tensor], pos_dev: Optional[torch.Tensor] = None,
                     scale: Optional[float] = None, splits: Optional[int] = None,
                     out: Optional[torch.Tensor] = None, tickets: Optional[torch.Tensor] = None) -> torch.Tensor:
    """single-token step: append this token's k/v at ``pos`` and attend over cache rows [0, pos]."""
    B, C3 = qkv.shape
    _, H, Tmax, D = k_cache.shape
    scale = scale if scale is not None else 1.0 / math.sqrt(D)
    if splits is None:      # with a device-resident position (CUDA-graph replay) the split count is frozen at capture
        # time, so size it for the whole cache
        splits = decode_splits(B * H, Tmax if pos_dev is not None else pos + 1)
    if out is None:
        out = torch.empty((B, C3 // 3), dtype=qkv.dtype, device=qkv.device)
    assert out.shape == (B, C3 // 3) and out.is_contiguous() and out.dtype == qkv.dtype
    ws = L.workspace(L.load().vg_attn_decode_workspace(B, H, D, splits), qkv.device)
    if tickets is not None:      # zero-initialised int32 [B*H] counters: split partials merged inside the same launch
        assert tickets.dtype == torch.int32 and tickets.numel() >= B * H
    L.call("vg_attn_decode", L.ptr(qkv), L.ptr(k_cache), L.ptr(v_cache), L.ptr(out), L.ptr(slopes), B, H, D, Tmax,
           pos, L.ptr(pos_dev), splits, float(scale), L.dtype_id(qkv.dtype), L.ptr(ws), ws.numel(), L.ptr(tickets),
           L.stream())
    return out


# ----------------------------------------------------------------------------------------- latent
def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.detach().float().contiguous()


class _LatentFront(torch.autograd.Function):
    """lvtr.py:151-169 fused: posterior heads + reparameterisation + log_q + embedding + fuse + BOS shift."""

    @staticmethod
    def forward(ctx, h_enc, eps, ids, mask_u8, init_state, w_mean, b_mean, w_logstd, b_logstd, tok_emb, w_fuse,
                b_fuse, temperature, act_dtype):
        ctx.set_materialize_grads(False)
        B, T, Ld = h_enc.shape
        E = tok_emb.shape[1]
        dev = h_enc.device
        a = L.LatentFrontArgs()
        a.B, a.T, a.latent_dim, a.emb_dim, a.vocab = B, T, Ld, E, tok_emb.shape[0]
        keep = dict(h_enc=_f32c(h_enc), eps=_f32c(eps), ids=ids.contiguous(), mask=mask_u8,
                    init_state=_f32c(init_state), w_mean=_f32c(w_mean), b_mean=_f32c(b_mean),
                    w_logstd=_f32c(w_logstd), b_logstd=_f32c(b_logstd), tok_emb=_f32c(tok_emb),
                    w_fuse=_f32c(w_fuse), b_fuse=_f32c(b_fuse))
        assert keep["ids"].dtype == torch.int64
        for kname, t in keep.items():
            setattr(a, kname, L.ptr(t))
        a.temperature = float(temperature)
        f32 = dict(dtype=torch.float32, device=dev)
        mean, logstd, z, log_q = (torch.empty((B, T, Ld), **f32) for _ in range(4))
        u = torch.empty((B, T, E), dtype=act_dtype, device=dev)
        u_shift = torch.empty((B, T, E), dtype=act_dtype, device=dev)
        a.mean, a.logstd, a.z, a.log_q = L.ptr(mean), L.ptr(logstd), L.ptr(z), L.ptr(log_q)
        a.u, a.u_shift, a.act_dtype = L.ptr(u), L.ptr(u_shift), L.dtype_id(act_dtype)
        L.call("vg_latent_front_fwd", C.byref(a), L.stream())
        ctx.keep = keep
        ctx.outs = (mean, logstd, z, log_q)
        ctx.meta = (B, T, Ld, E, float(temperature), act_dtype)
        return mean, logstd, z, log_q, u, u_shift

    @staticmethod
    def backward(ctx, d_mean, d_logstd, d_z, d_log_q, d_u, d_u_shift):
        B, T, Ld, E, temperature, act_dtype = ctx.meta
        keep = ctx.keep
        mean, logstd, z, log_q = ctx.outs
        dev = mean.device
        g = L.LatentFrontBwdArgs()
        a = g.f
        a.B, a.T, a.latent_dim, a.emb_dim, a.vocab = B, T, Ld, E, keep["tok_emb"].shape[0]
        for kname, t in keep.items():
            setattr(a, kname, L.ptr(t))
        a.temperature = temperature
        a.mean, a.logstd, a.z, a.log_q = L.ptr(mean), L.ptr(logstd), L.ptr(z), L.ptr(log_q)
        a.u, a.u_shift, a.act_dtype = L.ptr(mean), L.ptr(mean), L.dtype_id(act_dtype)   # not read by backward
        holders = []

        def f32(t):
            if t is None:
                return None
            t = t.float().contiguous()
            holders.append(t)
            return L.ptr(t)

        def act(t):
            if t is None:
                return None
            t = t.to(act_dtype).contiguous()
            holders.append(t)
            return L.ptr(t)

        g.d_z, g.d_log_q, g.d_mean_out, g.d_logstd_out = f32(d_z), f32(d_log_q), f32(d_mean), f32(d_logstd)
        g.d_u, g.d_u_shift = act(d_u), act(d_u_shift)
        o = dict(dtype=torch.float32, device=dev)
        d_h = torch.empty((B, T, Ld), **o)
        dwm, dws = torch.empty((Ld, Ld), **o), torch.empty((Ld, Ld), **o)
        dbm, dbs = torch.empty(Ld, **o), torch.empty(Ld, **o)
        demb = torch.empty_like(keep["tok_emb"])
        dwf, dbf = torch.empty((E, Ld), **o), torch.empty(E, **o)
        g.d_h_enc = L.ptr(d_h)
        g.d_w_mean, g.d_b_mean, g.d_w_logstd, g.d_b_logstd = L.ptr(dwm), L.ptr(dbm), L.ptr(dws), L.ptr(dbs)
        g.d_tok_emb, g.d_w_fuse, g.d_b_fuse = L.ptr(demb), L.ptr(dwf), L.ptr(dbf)
        ws = L.workspace(L.load().vg_latent_front_bwd_workspace(B, T, Ld, E, keep["tok_emb"].shape[0]), dev)
        L.call("vg_latent_front_bwd", C.byref(g), L.ptr(ws), ws.numel(), L.stream())
        return d_h, None, None, None, None, dwm, dbm, dws, dbs, demb, dwf, dbf, None, None


def latent_front(h_enc, eps, ids, mask, init_state, w_mean, b_mean, w_logstd, b_logstd, tok_emb, w_fuse, b_fuse,
                 temperature: float = 1.0, act_dtype: torch.dtype = torch.float32):
    """→ (mean, logstd, z, log_q) f32 [B,T,L] and (u, u_shift) act-dtype [B,T,E]."""
    return _LatentFront.apply(h_enc, eps, ids, _u8(mask), init_state, w_mean, b_mean, w_logstd, b_logstd, tok_emb,
                              w_fuse, b_fuse, temperature, act_dtype)


def _fill_back_args(a, head, z, log_q, mask_u8, flow, ln_eps, lo, hi):
    w1, b1, lnw, lnb, w2, b2 = flow
    M = head.shape[0]
    a.M, a.latent_dim, a.hidden, a.n_layers = M, w2.shape[1], w1.shape[1], w1.shape[0]
    a.head, a.head_ld = L.ptr(head), head.stride(0)
    a.z, a.log_q, a.mask = L.ptr(z), L.ptr(log_q), L.ptr(mask_u8)
    a.w1, a.b1, a.ln_w, a.ln_b, a.w2, a.b2 = (L.ptr(t) for t in (w1, b1, lnw, lnb, w2, b2))
    a.ln_eps, a.scale_lo, a.scale_hi = float(ln_eps), float(lo), float(hi)


class _LatentBack(torch.autograd.Function):
    """lvtr.py:172-191 fused: prior head columns + 4 conditional coupling layers + log_p + KL."""

    @staticmethod
    def forward(ctx, head, z, log_q, mask_u8, w1, b1, lnw, lnb, w2, b2, ln_eps, lo, hi):
        ctx.set_materialize_grads(False)
        head2 = head.reshape(-1, head.shape[-1])
        assert head2.dtype == torch.float32 and head2.stride(1) == 1
        M = head2.shape[0]
        Ld = w2.shape[1]
        z2, lq2 = _f32c(z).reshape(M, Ld), _f32c(log_q).reshape(M, Ld)
        flow = tuple(_f32c(t) for t in (w1, b1, lnw, lnb, w2, b2))
        dev = head.device
        log_p = torch.empty((M, Ld), dtype=torch.float32, device=dev)
        y = torch.empty((M, Ld), dtype=torch.float32, device=dev)
        kl_frame = torch.empty(M, dtype=torch.float32, device=dev)
        kl_sum = torch.empty(1, dtype=torch.float32, device=dev)
        a = L.LatentBackArgs()
        _fill_back_args(a, head2, z2, lq2, mask_u8, flow, ln_eps, lo, hi)
        a.log_p, a.y, a.kl_frame, a.kl_sum = L.ptr(log_p), L.ptr(y), L.ptr(kl_frame), L.ptr(kl_sum)
        ws = L.workspace(L.load().vg_latent_back_workspace(M, Ld, w1.shape[1], w1.shape[0]), dev)
        L.call("vg_latent_back_fwd", C.byref(a), L.ptr(ws), ws.numel(), L.stream())
        ctx.keep = (head2, z2, lq2, mask_u8, flow, float(ln_eps), float(lo), float(hi))
        ctx.shapes = (head.shape, z.shape)
        lead = z.shape[:-1]
        return log_p.view(*lead, Ld), y.view(*lead, Ld), kl_sum.view(())

    @staticmethod
    def backward(ctx, d_log_p, d_y, d_kl):
        if d_y is not None:
            raise NotImplementedError("gradient through the flow output `y` alone is not part of the hot path")
        head2, z2, lq2, mask_u8, flow, ln_eps, lo, hi = ctx.keep
        head_shape, z_shape = ctx.shapes
        M, Ld = z2.shape
        dev = head2.device
        maskf = mask_u8.view(M, 1).float()
        dlp = torch.zeros((M, Ld), dtype=torch.float32, device=dev) if d_log_p is None \
            else d_log_p.reshape(M, Ld).float()
        d_log_q = None
        if d_kl is not None:      # kl = Σ_valid mean_c(log_q − log_p)
            dlp = dlp - (d_kl.float() / Ld) * maskf
            d_log_q = ((d_kl.float() / Ld) * maskf).expand(M, Ld).reshape(z_shape).contiguous()
        dlp = dlp.contiguous()
        g = L.LatentBackBwdArgs()
        _fill_back_args(g.f, head2, z2, lq2, mask_u8, flow, ln_eps, lo, hi)
        d_head = torch.empty_like(head2)
        if head2.shape[1] > 2 * Ld + flow[0].shape[0] * 2 * flow[0].shape[1]:
            d_head.zero_()           # padded columns beyond the used head layout
        d_z = torch.empty((M, Ld), dtype=torch.float32, device=dev)
        grads = tuple(torch.empty_like(t) for t in flow)
        g.d_log_p, g.d_head, g.d_z = L.ptr(dlp), L.ptr(d_head), L.ptr(d_z)
        g.d_w1, g.d_b1, g.d_ln_w, g.d_ln_b, g.d_w2, g.d_b2 = (L.ptr(t) for t in grads)
        ws = L.workspace(L.load().vg_latent_back_workspace(M, Ld, flow[0].shape[1], flow[0].shape[0]), dev)
        L.call("vg_latent_back_bwd", C.byref(g), L.ptr(ws), ws.numel(), L.stream())
        return (d_head.view(head_shape), d_z.view(z_shape), d_log_q, None, *grads, None, None, None)


def latent_back(head, z, log_q, mask, w1, b1, lnw, lnb, w2, b2, ln_eps: float, scale_lo: float, scale_hi: float):
    """→ (log_p [..,L] masked, y [..,L] flow output, kl_sum scalar)."""
    return _LatentBack.apply(head, z, log_q, _u8(mask), w1, b1, lnw, lnb, w2, b2, ln_eps, scale_lo, scale_hi)


@torch.no_grad()
def latent_prior_sample(head, eps, temperature, w1, b1, lnw, lnb, w2, b2, ln_eps, scale_lo, scale_hi):
    """lvtr.py:267-275: z = Flow⁻¹(mean_p + exp(logstd_p)·eps·τ; FiLM columns of ``head``)."""
    head2 = head.reshape(-1, head.shape[-1])
    M, Ld = head2.shape[0], w2.shape[1]
    flow = tuple(_f32c(t) for t in (w1, b1, lnw, lnb, w2, b2))
    a = L.LatentBackArgs()
    dummy = head2
    _fill_back_args(a, head2, dummy, dummy, None, flow, ln_eps, scale_lo, scale_hi)
    e = _f32c(eps).reshape(M, Ld) if eps is not None else None
    z = torch.empty((M, Ld), dtype=torch.float32, device=head.device)
    L.call("vg_latent_prior_sample", C.byref(a), L.ptr(e), float(temperature), L.ptr(z), L.stream())
    return z.view(*head.shape[:-1], Ld)


# ----------------------------------------------------------------------------------------- losses
class _SoftmaxCE(torch.autograd.Function):
    """losses.py:30-41 — Σ over valid rows of −log softmax(logits)[target]."""

    @staticmethod
    def forward(ctx, logits, targets, mask_u8):
        lg = _rows2d(logits)
        rows, vocab = lg.shape
        tg = targets.reshape(-1).contiguous()
        dev = lg.device
        lse = torch.empty(rows, dtype=torch.float32, device=dev)
        loss_rows = torch.empty(rows, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        ws = L.workspace(L.load().vg_softmax_ce_workspace(rows), dev)
        L.call("vg_softmax_ce_fwd", L.ptr(lg), lg.stride(0), L.ptr(tg), L.ptr(mask_u8), L.ptr(lse), L.ptr(loss_rows),
               L.ptr(loss), rows, vocab, L.dtype_id(lg.dtype), L.ptr(ws), ws.numel(), L.stream())
        ctx.save_for_backward(lg, tg, mask_u8, lse)
        ctx.shape = logits.shape
        return loss.view(())

    @staticmethod
    def backward(ctx, d_loss):
        lg, tg, mask_u8, lse = ctx.saved_tensors
        rows, vocab = lg.shape
        dl = d_loss.float().reshape(1).contiguous()
        d_logits = torch.empty((rows, vocab), dtype=lg.dtype, device=lg.device)
        L.call("vg_softmax_ce_bwd", L.ptr(lg), lg.stride(0), L.ptr(tg), L.ptr(mask_u8), L.ptr(lse), L.ptr(dl),
               L.ptr(d_logits), vocab, rows, vocab, L.dtype_id(lg.dtype), L.stream())
        return d_logits.view(ctx.shape), None, None


def softmax_ce(logits: torch.Tensor, targets: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    return _SoftmaxCE.apply(logits, targets, _u8(mask))


@torch.no_grad()
def qsample(x0, noise, t, sqrt_ac, sqrt_1mac, mask, x0_scale: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """ddpm.py:328-334,352-359: x_t = (√ᾱ_t·x0·scale + √(1−ᾱ_t)·noise)·mask ; target = noise·mask."""
    B, T, Cc = x0.shape
    x0c, nc = _f32c(x0), _f32c(noise)
    x_t, target = torch.empty_like(x0c), torch.empty_like(x0c)
    L.call("vg_qsample", L.ptr(x0c), L.ptr(nc), L.ptr(t.contiguous()), L.ptr(_f32c(sqrt_ac)), L.ptr(_f32c(sqrt_1mac)),
           L.ptr(_u8(mask)), float(x0_scale), L.ptr(x_t), L.ptr(target), B, T, Cc, L.stream())
    return x_t, target


class _MaskedL1(torch.autograd.Function):
    """losses.py:9-27,44-57 — Σ_{b,t valid} mean_c |pred − target|."""

    @staticmethod
    def forward(ctx, pred, target, mask_u8):
        B, T, Cc = pred.shape
        p, tg = pred.contiguous(), target.contiguous()
        assert p.dtype == torch.float32 and tg.dtype == torch.float32
        loss = torch.empty(1, dtype=torch.float32, device=p.device)
        ws = L.workspace(L.load().vg_masked_l1_workspace(B, T, Cc), p.device)
        L.call("vg_masked_l1_fwd", L.ptr(p), L.ptr(tg), L.ptr(mask_u8), L.ptr(loss), B, T, Cc, L.ptr(ws), ws.numel(),
               L.stream())
        ctx.save_for_backward(p, tg, mask_u8)
        return loss.view(())

    @staticmethod
    def backward(ctx, d_loss):
        p, tg, mask_u8 = ctx.saved_tensors
        B, T, Cc = p.shape
        dl = d_loss.float().reshape(1).contiguous()
        d_pred = torch.empty_like(p)
        L.call("vg_masked_l1_bwd", L.ptr(p), L.ptr(tg), L.ptr(mask_u8), L.ptr(dl), L.ptr(d_pred), B, T, Cc, L.stream())
        return d_pred, None, None


def masked_l1(pred: torch.Tensor, target: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    return _MaskedL1.apply(pred.float(), target.float(), _u8(mask))


@torch.no_grad()
def sample_token(logits: torch.Tensor, u: Optional[torch.Tensor], temperature: float = 1.0) -> torch.Tensor:
    """lvtr.py:277-285.  ``u`` = U[0,1) per row → inverse-CDF multinomial; ``u=None`` → greedy argmax."""
    lg = _rows2d(logits)
    rows, vocab = lg.shape
    out = torch.empty(rows, dtype=torch.int64, device=lg.device)
    uu = _f32c(u).reshape(-1) if u is not None else None
    L.call("vg_sample_token", L.ptr(lg), lg.stride(0), L.ptr(uu), float(temperature), L.ptr(out), rows, vocab,
           L.dtype_id(lg.dtype), L.stream())
    return out.view(logits.shape[:-1])


# ------------------------------------------------------------------------------- decode (skinny) linear
def decode_linear(x: torch.Tensor, w: torch.Tensor, ws: torch.Tensor, *, norm_scale: Optional[torch.Tensor] = None,
                  x_ss: Optional[torch.Tensor] = None, norm_eps: float = 0.0, bias: Optional[torch.Tensor] = None,
                  act: int = ACT_NONE, residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                  out_f32: Optional[torch.Tensor] = None, y_ss: Optional[torch.Tensor] = None,
                  zero_ss: Optional[torch.Tensor] = None, overlap: bool = True) -> None:
    """Weight-streaming linear for the cached generation step (vg_decode_linear): x [B,K] bf16 (row stride free),
    w [N,K] bf16, ``ws`` a zero-filled workspace of ``decode_linear_workspace`` bytes (left zeroed)."""
    B, K = x.shape
    N = w.shape[0]
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.stride(1) == 1 and w.stride(1) == 1
    a = L.DecodeLinearArgs()
    a.B, a.N, a.K = B, N, K
    a.x, a.ldx, a.W, a.ldw = L.ptr(x), x.stride(0), L.ptr(w), w.stride(0)
    if norm_scale is not None:
        assert norm_scale.dtype == torch.float32 and x_ss is not None and x_ss.dtype == torch.float32
        a.norm_scale, a.x_ss, a.norm_eps = L.ptr(norm_scale), L.ptr(x_ss), float(norm_eps)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
        a.bias = L.ptr(bias)
    a.act = act
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.shape == (B, N) and residual.stride(1) == 1
        a.residual, a.ld_res = L.ptr(residual), residual.stride(0)
    if out is not None:
        assert out.dtype == torch.bfloat16 and out.shape == (B, N) and out.stride(1) == 1
        a.y, a.ldy = L.ptr(out), out.stride(0)
    if out_f32 is not None:
        assert out_f32.dtype == torch.float32 and out_f32.shape == (B, N) and out_f32.stride(1) == 1
        a.y_f32, a.ldy_f32 = L.ptr(out_f32), out_f32.stride(0)
    a.y_ss, a.zero_ss = L.ptr(y_ss), L.ptr(zero_ss)
    a.allow_overlap = int(overlap)
    L.call("vg_decode_linear", C.byref(a), L.ptr(ws), ws.numel(), L.stream())


def decode_linear_workspace(max_batch: int, max_n: int, device) -> torch.Tensor:
    return torch.zeros(int(L.load().vg_decode_linear_workspace(max_batch, max_n)), dtype=torch.uint8, device=device)
