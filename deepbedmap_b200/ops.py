"""Thin Python wrappers over the C ABI: they only compute pointers/strides from torch tensors
(torch is the allocator and stream host) and enqueue the CUDA kernels."""
from __future__ import annotations

import torch

from . import _lib

call = _lib.call


_AUX = None
_AUX2 = None


def _aux_stream2():
    """Second side stream (the discriminator's weight gradients; the generator's use _aux_stream)."""
    global _AUX2
    if _AUX2 is None:
        _AUX2 = torch.cuda.Stream()
    return _AUX2


def _aux_stream():
    """Side stream for gradient work nothing downstream in backward waits for (always joined by its caller)."""
    global _AUX
    if _AUX is None:
        _AUX = torch.cuda.Stream()
    return _AUX


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t, elem_offset: int = 0) -> int:
    return t.data_ptr() + elem_offset * t.element_size()


def _chk(t, dtype=torch.float32):
    if not (t.is_cuda and t.is_contiguous() and t.dtype == dtype):
        raise ValueError(f"expected contiguous CUDA {dtype} tensor, got {t.dtype} {t.device} contiguous={t.is_contiguous()}")


def empty(*shape, dtype=torch.float32):
    return torch.empty(shape, dtype=dtype, device="cuda")


def zeros(*shape, dtype=torch.float32):
    t = torch.empty(shape, dtype=dtype, device="cuda")
    if dtype == torch.float32:
        fill(t, 0.0)
    else:
        # raw memset through the same kernel on the fp32 view (sizes are multiples of 4 bytes)
        n32 = t.numel() * t.element_size() // 4
        call("dbm_fill_f32", t.data_ptr(), 0.0, n32, stream())
    return t


def fill(t, v: float):
    _chk(t)
    call("dbm_fill_f32", t.data_ptr(), float(v), t.numel(), stream())


# ---- fp32 conv family (NCHW, channel-slice views via batch strides) ----------------------------
def conv_out_hw(h, w, k, s, p):
    return (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1


def conv2d_fwd(x, x_c0, cin, w, b, y, y_c0, k, s, p, act=False):
    """y[:, y_c0:y_c0+O] = conv(x[:, x_c0:x_c0+cin], w) + b, optional LeakyReLU(0.2)."""
    n, cx, h, wd = x.shape
    o = w.shape[0]
    ho, wo = conv_out_hw(h, wd, k, s, p)
    assert y.shape[0] == n and y.shape[2] == ho and y.shape[3] == wo, (x.shape, y.shape)
    assert w.shape[1] == cin and x_c0 + cin <= cx and y_c0 + o <= y.shape[1]
    call("dbm_conv2d_fwd_f32", _ptr(x, x_c0 * h * wd), cx * h * wd, w.data_ptr(), b.data_ptr() if b is not None else None,
         _ptr(y, y_c0 * ho * wo), y.shape[1] * ho * wo, n, cin, h, wd, o, k, s, p, int(act), stream())


def conv2d_bwd_data(dy, dy_c0, w, dx, dx_c0, cin, k, s, p, accumulate=False):
    n, cdx, h, wd = dx.shape
    o = w.shape[0]
    ho, wo = dy.shape[2], dy.shape[3]
    call("dbm_conv2d_bwd_data_f32", _ptr(dy, dy_c0 * ho * wo), dy.shape[1] * ho * wo, w.data_ptr(),
         _ptr(dx, dx_c0 * h * wd), cdx * h * wd, n, cin, h, wd, o, k, s, p, int(accumulate), stream())


def conv2d_bwd_weight(x, x_c0, cin, dy, dy_c0, dw, k, s, p, db=None):
    """dw += x (*) dy ; db += sum dy."""
    n, cx, h, wd = x.shape
    o = dw.shape[0]
    ho, wo = dy.shape[2], dy.shape[3]
    call("dbm_conv2d_bwd_weight_f32", _ptr(x, x_c0 * h * wd), cx * h * wd, _ptr(dy, dy_c0 * ho * wo),
         dy.shape[1] * ho * wo, dw.data_ptr(), n, cin, h, wd, o, k, s, p, stream())
    if db is not None:
        call("dbm_bias_grad_f32", _ptr(dy, dy_c0 * ho * wo), dy.shape[1] * ho * wo, db.data_ptr(), n, o, ho * wo, stream())


def gemm(a, lda_m, lda_k, a_bs, b, ldb_k, ldb_n, b_bs, c, ldc_m, ldc_n, c_bs, bias, m, n, k, batch=1, act=False,
         accumulate=0, tc=False):
    """Batched strided GEMM (+ bias per n, LeakyReLU). tc: operands rounded to bf16, tensor cores (dbm_gemm_bf16)."""
    call("dbm_gemm_bf16" if tc else "dbm_gemm_f32", a.data_ptr(), lda_m, lda_k, a_bs, b.data_ptr(), ldb_k, ldb_n, b_bs, c.data_ptr(), ldc_m,
         ldc_n, c_bs, bias.data_ptr() if bias is not None else None, m, n, k, batch, int(act), accumulate, stream())


def axpby(x, x_c0, y, y_c0, out, out_c0, channels, a, b):
    """out[:, out_c0:+channels] = a * x[:, x_c0:+channels] + b * y[:, y_c0:+channels] (y may be None)."""
    n = x.shape[0]
    hw = x.shape[2] * x.shape[3]
    call("dbm_axpby_f32", _ptr(x, x_c0 * hw), x.shape[1] * hw, _ptr(y, y_c0 * hw) if y is not None else None,
         (y.shape[1] * hw) if y is not None else 0, _ptr(out, out_c0 * hw), out.shape[1] * hw, float(a), float(b), n,
         channels * hw, stream())


def lrelu_bwd(dy, dy_c0, y, y_c0, dx, dx_c0, channels, accumulate=False):
    n = dy.shape[0]
    hw = dy.shape[2] * dy.shape[3] if dy.dim() == 4 else 1
    call("dbm_lrelu_bwd_f32", _ptr(dy, dy_c0 * hw), dy.shape[1] * hw, _ptr(y, y_c0 * hw), y.shape[1] * hw,
         _ptr(dx, dx_c0 * hw), dx.shape[1] * hw, n, channels * hw, int(accumulate), stream())


def upsample2_fwd(x):
    n, c, h, w = x.shape
    y = empty(n, c, 2 * h, 2 * w)
    call("dbm_upsample2_fwd_f32", x.data_ptr(), y.data_ptr(), n * c, h, w, stream())
    return y


def upsample2_bwd(dy):
    n, c, h2, w2 = dy.shape
    dx = empty(n, c, h2 // 2, w2 // 2)
    call("dbm_upsample2_bwd_f32", dy.data_ptr(), dx.data_ptr(), n * c, h2 // 2, w2 // 2, stream())
    return dx


# ---- layouts for the tensor-core path ---------------------------------------------------------
def nchw_to_slab8(src, dst, dst_cs0=0):
    n, c, h, w = src.shape
    call("dbm_nchw_to_slab8", src.data_ptr(), 0, dst.data_ptr(), n, c, h, w, dst.shape[1], dst_cs0, stream())


def slab8_to_nchw(src, c, src_cs0=0):
    n, cs, h, w, _ = src.shape
    dst = empty(n, c, h, w)
    call("dbm_slab8_to_nchw", src.data_ptr(), cs, src_cs0, dst.data_ptr(), 0, n, c, h, w, stream())
    return dst


def nchw_to_slab4(src):
    n, c, h, w = src.shape
    dst = empty(n, c // 4, h, w, 4)
    call("dbm_nchw_to_slab4", src.data_ptr(), 0, dst.data_ptr(), n, c, h, w, stream())
    return dst


def slab4_to_nchw(src, c_keep):
    n, cs, h, w, _ = src.shape
    dst = empty(n, c_keep, h, w)
    call("dbm_slab4_to_nchw", src.data_ptr(), dst.data_ptr(), 0, n, cs * 4, c_keep, h, w, stream())
    return dst


def pack_conv3x3(w, cout_padded, ck=32):
    o, cin = w.shape[0], w.shape[1]
    packed = empty(9 * cin * cout_padded, dtype=torch.bfloat16)
    call("dbm_pack_conv3x3_weights", w.data_ptr(), packed.data_ptr(), o, cin, cout_padded, ck, stream())
    return packed


def deform_conv_umma(x_s8, off_s4, wpacked, bias, out_s8, act=False, out_cs0=0, next_out1_w=None):
    """Deformable 3x3 conv 64->64 on the tensor cores: x (N,8,H,W,8) bf16, offsets (N,>=5,H,W,4) fp32.
    ``next_out1_w``: (1,64,3,3) filter of a following single-output deformable layer -> also returns its nine
    projected planes (N,9,H,W), computed in the epilogue (finish with deform_out1_sample)."""
    n, cs, h, w, _ = x_s8.shape
    assert cs == 8, "deform_conv_umma needs 64 input channels"
    proj = None
    if next_out1_w is not None:
        assert tuple(next_out1_w.shape) == (1, 64, 3, 3)
        proj = empty(n, 9, h, w)
    call("dbm_deform_conv_umma", x_s8.data_ptr(), off_s4.data_ptr(), off_s4.shape[1], wpacked.data_ptr(),
         bias.data_ptr(), n, h, w, int(act), out_s8.data_ptr(), out_s8.shape[1], out_cs0,
         next_out1_w.data_ptr() if proj is not None else None, proj.data_ptr() if proj is not None else None, stream())
    return proj


def deform_out1_sample(proj, off_s4, bias):
    """Sampling half of the single-output deformable layer on planes projected by deform_conv_umma."""
    n, _, h, wd = proj.shape
    y = empty(n, 1, h, wd)
    call("dbm_deform_out1_sample", proj.data_ptr(), off_s4.data_ptr(), off_s4.shape[1], bias.data_ptr(), y.data_ptr(),
         n, h, wd, stream())
    return y


def deform_conv_out1(x_s8, off_s4, w, bias):
    """Final deformable conv 64->1: fp32 NCHW output (N,1,H,W)."""
    n, cs, h, wd, _ = x_s8.shape
    assert cs == 8 and tuple(w.shape) == (1, 64, 3, 3)
    y = empty(n, 1, h, wd)
    proj = empty(n, 9, h, wd)
    call("dbm_deform_conv_out1", x_s8.data_ptr(), off_s4.data_ptr(), off_s4.shape[1], w.data_ptr(), bias.data_ptr(),
         y.data_ptr(), proj.data_ptr(), n, h, wd, stream())
    return y


def conv3x3_umma(inp, cin, wpacked, bias, cout_padded, *, beta=0.0, act=False, up2=False, out=None, out_cs0=0,
                 out_f32=None, out_f32_cs0=0, res1=None, res2=None):
    """tcgen05 implicit-GEMM 3x3 conv on slab8 bf16 input (N, CS, H, W, 8)."""
    n, cs, h, w, _ = inp.shape
    call("dbm_conv3x3_umma", inp.data_ptr(), cs, cin, wpacked.data_ptr(), bias.data_ptr(), cout_padded, n, h, w,
         float(beta), int(act), int(up2),
         out.data_ptr() if out is not None else None, out.shape[1] if out is not None else 0, out_cs0,
         out_f32.data_ptr() if out_f32 is not None else None, out_f32.shape[1] if out_f32 is not None else 0,
         out_f32_cs0, res1.data_ptr() if res1 is not None else None, res2.data_ptr() if res2 is not None else None,
         stream())


# ---- deformable conv (fp32 path) -----------------------------------------------------------------
def deform_sample(x, offset):
    """Bilinear samples of every (channel, tap): cols (N, C*9, H*W) fp32 (the operand of the weight gradient)."""
    n, c, h, wd = x.shape
    cols = empty(n, c * 9, h * wd)
    call("dbm_deform_sample_f32", x.data_ptr(), offset.data_ptr(), cols.data_ptr(), n, c, h, wd, stream())
    return cols


def deform_sample_slab8(x8, offset):
    """The same cols from the bf16 slab8 copy of the input the fused forward gathered from (16-byte corner loads)."""
    n, _, h, wd, _ = x8.shape
    cols = empty(n, 576, h * wd)
    call("dbm_deform_sample_slab8_f32", x8.data_ptr(), offset.data_ptr(), cols.data_ptr(), n, h, wd, stream())
    return cols


def deform_conv_fwd_fused(x, offset, wpacked_ck64, b, act=False, keep_slab8=False):
    """Forward of a 64 -> 64 deformable conv in one tcgen05 kernel (gather -> UMMA -> bias / LeakyReLU): x (N,64,H,W)
    fp32, offset (N,18,H,W) fp32 -> y (N,64,H,W) fp32. Bilinear samples and filter are rounded to bf16 exactly as
    deform_conv_fwd(tc=True) does; no cols buffer is produced (backward re-samples, deform_sample)."""
    n, c, h, wd = x.shape
    assert c == 64 and tuple(offset.shape) == (n, 18, h, wd)
    x8 = empty(n, 8, h, wd, 8, dtype=torch.bfloat16)
    nchw_to_slab8(x, x8)
    y = empty(n, 64, h, wd)
    call("dbm_deform_conv_umma_nchw", x8.data_ptr(), offset.data_ptr(), wpacked_ck64.data_ptr(), b.data_ptr(), n, h, wd,
         int(act), y.data_ptr(), stream())
    return (y, x8) if keep_slab8 else y


def deform_conv_fwd(x, offset, w, b, act=False, tc=False):
    """x (N,C,H,W), offset (N,18,H,W), w (O,C,3,3) -> y (N,O,H,W), cols (N, C*9, H*W) kept for backward.
    tc: the contraction on the tensor cores with bf16-rounded operands (bf16 training path)."""
    n, c, h, wd = x.shape
    o = w.shape[0]
    hw = h * wd
    cols = empty(n, c * 9, hw)
    call("dbm_deform_sample_f32", x.data_ptr(), offset.data_ptr(), cols.data_ptr(), n, c, h, wd, stream())
    y = empty(n, o, h, wd)
    k = c * 9
    # per image: y[o, p] = sum_k cols[k, p] * w[o, k]   (M = pixels, N = O)
    gemm(cols, 1, hw, k * hw, w, 1, k, 0, y, 1, hw, o * hw, b, hw, o, k, batch=n, act=act, tc=tc)
    return y, cols


def deform_conv_bwd(x, offset, w, cols, dy, dw, db, dx, tc=False):
    """Accumulates dw, db; dx += d/dx; returns doffset (N,18,H,W)."""
    n, c, h, wd = x.shape
    o = w.shape[0]
    hw = h * wd
    k = c * 9
    # dw[o, kk] += sum_{n,p} dy[n,o,p] * cols[n,kk,p]   (atomic across the batch)
    # The weight / bias gradients only feed the optimizer: they run on a side stream beside the cols gradient and
    # its scatter (an atomics-bound kernel that leaves most of the GPU idle), joined before returning.
    cur, aux = torch.cuda.current_stream(), _aux_stream()
    aux.wait_stream(cur)
    with torch.cuda.stream(aux):
        gemm(dy, hw, 1, o * hw, cols, 1, hw, k * hw, dw, k, 1, 0, None, o, k, hw, batch=n, accumulate=2, tc=tc)
        call("dbm_bias_grad_f32", dy.data_ptr(), o * hw, db.data_ptr(), n, o, hw, stream())
    # dcols[n, kk, p] = sum_o w[o, kk] * dy[n, o, p]
    dcols = empty(n, k, hw)
    gemm(dy, 1, hw, o * hw, w, k, 1, 0, dcols, 1, hw, k * hw, None, hw, k, o, batch=n, tc=tc)
    doff = empty(n, 18, h, wd)
    call("dbm_deform_bwd_f32", x.data_ptr(), offset.data_ptr(), dcols.data_ptr(),
         dx.data_ptr() if dx is not None else None, doff.data_ptr(), n, c, h, wd, stream())
    cur.wait_stream(aux)
    return doff


def deform1_conv_fwd(x, offset, w, b):
    """Single-output deformable conv by tap projection: x (N,C,H,W), w (1,C,3,3) -> y (N,1,H,W) and the projected
    planes (N,9,H,W) kept for backward."""
    n, c, h, wd = x.shape
    assert tuple(w.shape) == (1, c, 3, 3)
    y = empty(n, 1, h, wd)
    proj = empty(n, 9, h, wd)
    call("dbm_deform1_fwd_f32", x.data_ptr(), offset.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(),
         proj.data_ptr(), n, c, h, wd, stream())
    return y, proj


def deform1_conv_bwd(x, offset, w, proj, dy, dw, db, dx, accumulate_dx=False):
    """Accumulates dw, db; dx = (or +=) d/dx; returns doffset (N,18,H,W)."""
    n, c, h, wd = x.shape
    hw = h * wd
    call("dbm_bias_grad_f32", dy.data_ptr(), hw, db.data_ptr(), n, 1, hw, stream())
    doff = empty(n, 18, h, wd)
    scratch = empty(n, 9, h, wd)
    call("dbm_deform1_bwd_f32", x.data_ptr(), offset.data_ptr(), w.data_ptr(), proj.data_ptr(), dy.data_ptr(),
         dw.data_ptr(), dx.data_ptr() if dx is not None else None, int(accumulate_dx), doff.data_ptr(),
         scratch.data_ptr(), n, c, h, wd, stream())
    return doff
