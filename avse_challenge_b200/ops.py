"""Thin torch-tensor wrappers over the C ABI (one function per ``mtn_*`` entry point).

These exist for the parity tests and for ``engine.py``; they allocate outputs with torch's caching
allocator, pass raw device pointers + the current CUDA stream, and raise on any error.  PyTorch is
plumbing only (memory + streams); every computation below runs in ``libmtn_b200.so``.
"""
from __future__ import annotations

import math

import torch

from . import _lib
from ._lib import GemmArgs, GnApplyArgs, ScanArgs, StreamPushArgs, check, ptr

LOG2E = 1.4426950408889634


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.MtnError("mtn ops need CUDA tensors (there is no CPU path)")


def rp_for(dt_rank: int) -> int:
    """K extent of the tensor-core dt_proj operand: the dt columns padded to a multiple of 16."""
    return 16 if dt_rank <= 16 else 32


def n_dbl_for(dt_rank: int) -> int:
    """Padded width of one direction's [dt | B | C] row: 48 for R <= 16, 64 for R = 32."""
    w = dt_rank + 32
    return 48 if w <= 48 else 64


def split_planes(x: torch.Tensor, planes: int) -> torch.Tensor:
    """fp32 [rows, cols] -> bf16 [planes, rows, cols] (hi / lo split when planes == 2)."""
    _req_cuda(x)
    assert x.dim() == 2 and x.dtype == torch.float32 and x.stride(1) == 1
    out = torch.empty((planes, x.shape[0], x.shape[1]), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().mtn_split_planes(ptr(x), x.stride(0), ptr(out), x.shape[0], x.shape[1], planes, _stream()),
          "mtn_split_planes")
    return out


def planes_to_float(p: torch.Tensor) -> torch.Tensor:
    """Test helper: reconstruct fp32 from planes (sum over the plane axis)."""
    return p.float().sum(dim=0)


def encoder_cln(mix, w_enc, gamma, beta, planes, eps=1e-8, mix_w=None, yn=None, T=None):
    """mix [B, ld] fp32 (row stride a multiple of 4 floats; ``T`` = valid samples per row, default ld)."""
    _req_cuda(mix, w_enc)
    B = mix.shape[0]
    if T is None:
        T = mix.shape[1]
    if mix.stride(0) % 4 != 0 or mix.stride(1) != 1 or mix.data_ptr() % 16 != 0:  # re-pitch for 128-bit frame loads
        padded = torch.zeros((B, (T + 3) // 4 * 4), dtype=torch.float32, device=mix.device)
        padded[:, :T] = mix[:, :T]
        mix = padded
    N = w_enc.shape[0]
    L = (T - 16) // 8 + 1
    if mix_w is None:
        mix_w = torch.empty((B * L, N), dtype=torch.float32, device=mix.device)
    if yn is None:
        yn = torch.empty((planes, B * L, N), dtype=torch.bfloat16, device=mix.device)
    check(_lib.load().mtn_encoder_cln_fwd(ptr(mix), mix.stride(0), ptr(w_enc), ptr(gamma), ptr(beta), ptr(mix_w), ptr(yn),
                                          B, T, L, N, planes, eps, _stream()), "mtn_encoder_cln_fwd")
    return mix_w, yn


def cln(x, gamma, beta, planes, eps=1e-8, yn=None):
    _req_cuda(x)
    M, N = x.shape
    if yn is None:
        yn = torch.empty((planes, M, N), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().mtn_cln_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(yn), M, N, planes, eps, _stream()), "mtn_cln_fwd")
    return yn


def gemm(a_planes, w_planes, M, N, K, *, out=None, groups=1, out_group_stride=0, epilogue=_lib.EPI_STORE,
         epi_param=0, aux=None, out_bf16=False, ldo=None, max_ctas=0, out2=None, rowsum=None, rowsq=None,
         rowsq_scale=0.0, rowsq_eps=0.0):
    """out[M, groups x N] = epilogue(A[M, K_g] @ W_g[N, K]^T); A planes [P, rows, lda], W planes [P, groups*N, K].
    ``rowsq`` (fp32 [parts, M], written by an ``EPI_RESADD`` call): accumulator row r is scaled by
    rsqrt(sum_k rowsq[k, r] * rowsq_scale + rowsq_eps) first (RMSNorm of the A rows folded in).  ``EPI_RESADD``: out (fp32,
    in/out) += result when ``epi_param`` else := result; the updated rows also go to ``out2`` (bf16 planes [P, rows, N])
    and their sums of squares to ``rowsum`` (fp32 [rowsum_parts(N), M], partial sums, plain stores)."""
    _req_cuda(a_planes, w_planes)
    P, a_rows, lda = a_planes.shape
    assert w_planes.shape == (P, groups * N, K), (tuple(w_planes.shape), (P, groups * N, K))
    if out is None:
        width = N if groups == 1 else out_group_stride * groups
        out = torch.empty((M, width), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=a_planes.device)
    args = GemmArgs(a=ptr(a_planes), w=ptr(w_planes), out=ptr(out), aux=ptr(aux), M=M, N=N, K=K, a_rows=a_rows,
                    lda=lda, ldo=ldo if ldo is not None else out.stride(0),
                    ld_aux=(aux.stride(0) if aux is not None and aux.dim() == 2 else 0), planes=P, groups=groups,
                    out_group_stride=out_group_stride, epilogue=epilogue, epi_param=epi_param,
                    out_bf16=int(out_bf16), max_ctas=max_ctas, out2=ptr(out2), rowsum=ptr(rowsum),
                    ldo2=(out2.stride(1) if out2 is not None else 0), a2_rows=(out2.shape[1] if out2 is not None else 0),
                    rowsq=ptr(rowsq), rowsq_scale=rowsq_scale, rowsq_eps=rowsq_eps,
                    rowsq_parts=(rowsq.shape[0] if rowsq is not None else 0))
    check(_lib.load().mtn_gemm_fwd(args, _stream()), "mtn_gemm_fwd")
    return out


def rowsum_parts(N: int) -> int:
    """Partial-sum planes an ``EPI_RESADD`` GEMM with N output columns writes (see ``mtn_gemm_args.rowsum``)."""
    return int(_lib.load().mtn_gemm_rowsum_parts(N))


def add_rmsnorm(h, res, res_valid, g, planes, eps=1e-5, xn=None, out_f32=None, beta=None):
    """``out_f32`` (fp32 [M, D]): also (or, with ``xn=False``, only) write the normalised rows in fp32.
    ``beta`` ([D]): LayerNorm (mean removed, weight ``g``, bias ``beta``) instead of RMSNorm."""
    _req_cuda(res, g, beta)
    M, D = res.shape
    if xn is None:
        xn = torch.empty((planes, M, D), dtype=torch.bfloat16, device=res.device)
    elif xn is False:
        assert out_f32 is not None
        xn = None
    check(_lib.load().mtn_add_norm_fwd(ptr(h), ptr(res), int(res_valid), ptr(g), ptr(beta), ptr(xn), ptr(out_f32), M, D, planes,
                                       eps, _stream()), "mtn_add_norm_fwd")
    return xn if xn is not None else out_f32


def conv_silu(xz, conv_w, conv_b, batch, L, di, planes, u=None, halo_lo=None, halo_hi=None, dir_mask=3):
    """xz [M, >= di] (fp32 or bf16; xs = first di columns) -> u planes [P, M, 2*di].
    ``halo_lo`` / ``halo_hi`` (fp32 [batch, 3, di]): xs rows just before / after this time chunk (None = zero pad).
    ``dir_mask``: 1 = forward half only (unidirectional stacks; the backward columns of ``u`` are left untouched)."""
    _req_cuda(xz, conv_w, conv_b, halo_lo, halo_hi)
    M = batch * L
    if u is None:
        u = torch.empty((planes, M, 2 * di), dtype=torch.bfloat16, device=xz.device)
    for h in (halo_lo, halo_hi):
        if h is not None:
            assert h.dtype == torch.float32 and h.is_contiguous() and tuple(h.shape) == (batch, 3, di)
    check(_lib.load().mtn_conv_silu_dir_fwd(ptr(xz), xz.stride(0), int(xz.dtype == torch.bfloat16), ptr(conv_w),
                                            ptr(conv_b), ptr(u), u.shape[1], ptr(halo_lo), ptr(halo_hi), batch, L, di,
                                            planes, dir_mask, _stream()), "mtn_conv_silu_dir_fwd")
    return u


def conv_xproj(xz, conv_w, conv_b, w_x, batch, L, di, planes, n_dbl, u=None, dbl=None):
    """``conv_silu`` of both directions fused with the two-group x_proj GEMM (``mtn_conv_xproj_fwd``): xz [M, >= di] ->
    (u planes [P, M, 2*di], dbl fp32 [M, 2*n_dbl]); ``w_x`` bf16 planes [P, 2*n_dbl, di].  Whole sequences only (no halos)."""
    _req_cuda(xz, conv_w, conv_b, w_x, u, dbl)
    M = batch * L
    assert tuple(w_x.shape) == (planes, 2 * n_dbl, di) and w_x.is_contiguous(), tuple(w_x.shape)
    if u is None:
        u = torch.empty((planes, M, 2 * di), dtype=torch.bfloat16, device=xz.device)
    if dbl is None:
        dbl = torch.empty((M, 2 * n_dbl), dtype=torch.float32, device=xz.device)
    check(_lib.load().mtn_conv_xproj_fwd(ptr(xz), xz.stride(0), int(xz.dtype == torch.bfloat16), ptr(conv_w), ptr(conv_b), ptr(u),
                                         u.shape[1], ptr(w_x), ptr(dbl), dbl.stride(0), n_dbl, batch, L, di, planes, _stream()),
          "mtn_conv_xproj_fwd")
    return u, dbl


def scan(u, dbl, z, z_col0, w_dt, dt_bias, A2, Dskip, batch, L, di, R, *, y=None, h_in=None, h_out=None, dir_mask=3,
         sum_delta=None, L_last=0, summary_only=False, dtp=None):
    """Both-direction selective scan; see ``mtn_scan_args`` in include/mtn_b200.h.
    ``summary_only``: no output is written (chunk-summary pass: ``h_out`` and ``sum_delta`` only).
    ``dtp`` (bf16 [M, 2, 2, RP], from ``gemm(..., epilogue=EPI_XPROJ)``): run dt_proj on the tensor cores."""
    _req_cuda(u, dbl, z)
    P = u.shape[0]
    nd = n_dbl_for(R)
    if y is None and not summary_only:
        y = torch.empty_like(u)
    args = ScanArgs(u=ptr(u), dbl=ptr(dbl), z=ptr(z), w_dt=ptr(w_dt), dt_bias=ptr(dt_bias), A2=ptr(A2),
                    Dskip=ptr(Dskip), y=None if summary_only else ptr(y), h_in=ptr(h_in), h_out=ptr(h_out),
                    batch=batch, L=L, di=di, R=R, n_dbl=nd, ld_dbl=dbl.stride(0), ldz=z.stride(0), z_col0=z_col0,
                    planes=P, z_bf16=int(z.dtype == torch.bfloat16), dir_mask=dir_mask, sum_delta=ptr(sum_delta),
                    L_last=L_last, dtp=ptr(dtp))
    check(_lib.load().mtn_scan_fwd(args, _stream()), "mtn_scan_fwd")
    return y


def softmax_mask(score, mix_w, rows, N, n_spk):
    """In place: softmax over each speaker's N channels of score [rows, n_spk*N] (x mix_w [rows, N] when given)."""
    _req_cuda(score, mix_w)
    assert score.dtype == torch.float32 and score.is_contiguous()
    check(_lib.load().mtn_softmax_mask_fwd(ptr(score), ptr(mix_w), rows, N, n_spk, _stream()), "mtn_softmax_mask_fwd")
    return score


def fold_states(h_end, sum_delta, A2, g0, n_out, *, h0=None, want_final=False, dir_mask=3):
    """Compose chunk operators (see ``mtn_fold_states_fwd``).  h_end [2, G, di, 16], sum_delta [2, G, di] ->
    (h_in [2, n_out, di, 16], h_final [2, di, 16] or None)."""
    _req_cuda(h_end, sum_delta, A2, h0)
    _, G, di, ns = h_end.shape
    assert ns == 16 and tuple(sum_delta.shape) == (2, G, di) and h_end.is_contiguous() and sum_delta.is_contiguous()
    h_in = torch.zeros((2, n_out, di, 16), dtype=torch.float32, device=h_end.device)
    h_final = torch.zeros((2, di, 16), dtype=torch.float32, device=h_end.device) if want_final else None
    check(_lib.load().mtn_fold_states_fwd(ptr(h_end), ptr(sum_delta), ptr(A2), ptr(h0), ptr(h_in), ptr(h_final), G, di,
                                          g0, n_out, dir_mask, _stream()), "mtn_fold_states_fwd")
    return h_in, h_final


def fold_states_packed(pack, A2, W, cmax, di, g0, n_out, *, h_in=None, h0=None, h_final=None, dir_mask=3):
    """``fold_states`` over the all-gathered packed summaries (see ``mtn_fold_states_packed_fwd``): ``pack`` fp32
    ``[W, 2*cmax*di*17]`` -> ``h_in [2, n_out, di, 16]`` (written into the given buffer when passed)."""
    _req_cuda(pack, A2, h0, h_in, h_final)
    assert pack.dtype == torch.float32 and pack.is_contiguous() and pack.numel() == W * 2 * cmax * di * 17
    if h_in is None:
        h_in = torch.empty((2, n_out, di, 16), dtype=torch.float32, device=pack.device)
    check(_lib.load().mtn_fold_states_packed_fwd(ptr(pack), ptr(A2), ptr(h0), ptr(h_in), ptr(h_final), W, cmax, di, g0,
                                                 n_out, dir_mask, _stream()), "mtn_fold_states_packed_fwd")
    return h_in


def decoder(sep, w_dec, batch, T, L, N, n_spk=2, est=None, frames=None, tail=None):
    """``tail`` (fp32 [batch, n_spk, 8], in/out): streaming overlap-add carry (then ``T`` must be ``8 * L``)."""
    _req_cuda(sep, w_dec, tail)
    if frames is None:
        frames = torch.empty((batch * L, n_spk, 16), dtype=torch.float32, device=sep.device)
    if est is None:
        est = torch.empty((batch, T, n_spk), dtype=torch.float32, device=sep.device)
    if tail is not None:
        assert tail.dtype == torch.float32 and tail.is_contiguous() and tuple(tail.shape) == (batch, n_spk, 8)
    check(_lib.load().mtn_decoder_stream_fwd(ptr(sep), ptr(w_dec), ptr(frames), ptr(est), ptr(tail), batch, T, L, N, n_spk,
                                             _stream()), "mtn_decoder_stream_fwd")
    return est


def stream_push(mix, in_tail, est, halo, h, ola_tail, head, bot_frag, mask_frag, layer_vec, layer_frag, *, B, F, N, D,
                di, R, n_spk, n_layers, first, eps_cln=1e-8, eps_rms=1e-5, timeline=None, halo_strides=None, halo_rows=3,
                stack_x=None, stack_out=None, dsl=64):
    """One streaming chunk of ``F <= 32`` frames per stream through the whole causal separator in one launch
    (``mtn_stream_push_fwd``; weight layouts in include/mtn_b200.h, packer in stream_fused.py).  ``stack_x`` / ``stack_out``
    (fp32 [B, F, D]): stack-only mode = ``MambaBlocksSequential.forward(x, inference_params)``; mix / est / tails are unused.
    ``halo_strides`` = (stream, layer) strides of ``halo`` in floats when it is not a dense [n_layers, B, 3, di] tensor."""
    _req_cuda(mix, in_tail, est, halo, h, ola_tail, head, bot_frag, mask_frag, layer_vec, layer_frag, timeline, stack_x, stack_out)
    if stack_x is None:
        assert mix.dtype == torch.float32 and mix.stride(1) == 1 and est.is_contiguous() and tuple(est.shape) == (B, 8 * F, n_spk)
        assert mix.shape[1] == 8 * F + (8 if first else 0) and in_tail.is_contiguous() and tuple(in_tail.shape) == (B, 8)
        assert ola_tail.is_contiguous() and tuple(ola_tail.shape) == (B, n_spk, 8)
    else:
        assert stack_x.dtype == torch.float32 and stack_x.is_contiguous() and tuple(stack_x.shape) == (B, F, D)
        assert stack_out.dtype == torch.float32 and stack_out.is_contiguous() and tuple(stack_out.shape) == (B, F, D)
    if halo_strides is None:
        assert halo.is_contiguous() and tuple(halo.shape) == (n_layers, B, 3, di)
        halo_strides = (3 * di, B * 3 * di)
    assert h.is_contiguous() and h.shape[0] == n_layers and tuple(h.shape[-3:]) == (B, di, 16)
    assert timeline is None or (timeline.dtype == torch.int64 and timeline.numel() >= (n_layers + 2) * 16)
    assert layer_vec.is_contiguous() and layer_frag.is_contiguous()
    args = StreamPushArgs(mix=ptr(mix), in_tail=ptr(in_tail), first=int(bool(first)), timeline=ptr(timeline), est=ptr(est),
                          halo=ptr(halo), h=ptr(h), ola_tail=ptr(ola_tail), head=ptr(head),
                          bot_frag=ptr(bot_frag), mask_frag=ptr(mask_frag), layer_vec=ptr(layer_vec),
                          layer_frag=ptr(layer_frag), h_layer_stride=h.stride(0),
                          layer_vec_stride=layer_vec.stride(0), layer_frag_stride=layer_frag.stride(0) * layer_frag.element_size(),
                          B=B, F=F, N=N, D=D, di=di, R=R, n_spk=n_spk, n_layers=n_layers, ld_mix=(mix.stride(0) if mix is not None else 0),
                          eps_cln=eps_cln, eps_rms=eps_rms, halo_stream_stride=halo_strides[0], halo_layer_stride=halo_strides[1], halo_rows=halo_rows,
                          stack_x=ptr(stack_x), stack_out=ptr(stack_out), dsl=dsl)
    check(_lib.load().mtn_stream_push_fwd(args, _stream()), "mtn_stream_push_fwd")
    return est if stack_x is None else stack_out


# ---------------------------------------------------------------------------------------------------- DPMamba glue
def dp_num_chunks(L: int, K: int) -> int:
    """Chunks S of ``Dual_Path_Model._Segmentation`` for L frames, chunk size K (needs no GPU)."""
    S = int(_lib.load().mtn_dp_num_chunks(L, K))
    if S < 0:
        raise _lib.MtnError(f"dp_num_chunks: bad L={L} K={K}")
    return S


def gn_partials(batch, rows, C, device):
    n = int(_lib.load().mtn_gn_partials_bytes(batch, rows, C))
    return torch.empty(n // 8, dtype=torch.float64, device=device)


def gn_stats(x, batch, rows, C, partials=None):
    """Pass 1 of GroupNorm(1, C): x fp32 [batch, rows, C] (any row order) -> per-utterance partial sums."""
    _req_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.numel() == batch * rows * C
    if partials is None:
        partials = gn_partials(batch, rows, C, x.device)
    check(_lib.load().mtn_gn_stats_fwd(ptr(x), ptr(partials), batch, rows, C, _stream()), "mtn_gn_stats_fwd")
    return partials


def gn_apply(x, partials, w, bias, batch, S, K, C, *, eps=1e-8, skip=None, out_a=None, out_a2=None, out_t=None, planes=None,
             x_transposed=False, blend=None):
    _req_cuda(x, partials, w, bias, skip, out_a, out_a2, out_t, planes, blend)
    args = GnApplyArgs(x=ptr(x), partials=ptr(partials), w=ptr(w), bias=ptr(bias), skip=ptr(skip), out_a=ptr(out_a),
                       out_a2=ptr(out_a2), out_t=ptr(out_t), planes=ptr(planes), batch=batch, S=S, K=K, C=C,
                       x_transposed=int(x_transposed), n_planes=(planes.shape[0] if planes is not None else 0),
                       plane_rows=(planes.shape[1] if planes is not None else 0), eps=eps, blend=ptr(blend))
    check(_lib.load().mtn_gn_apply_fwd(args, _stream()), "mtn_gn_apply_fwd")


def gn_apply_norm(x, partials, w, bias, batch, S, K, C, *, res_next, xn_next, g_next, next_transposed, eps=1e-8, rms_eps=1e-5,
                  skip=None, out_a=None, x_transposed=False, blend=None):
    """``gn_apply`` fused with the next stack's opening Add -> RMSNorm (see ``mtn_gn_apply_norm_fwd``)."""
    _req_cuda(x, partials, w, bias, skip, out_a, blend, res_next, xn_next, g_next)
    args = GnApplyArgs(x=ptr(x), partials=ptr(partials), w=ptr(w), bias=ptr(bias), skip=ptr(skip), out_a=ptr(out_a),
                       out_a2=None, out_t=None, planes=None, batch=batch, S=S, K=K, C=C, x_transposed=int(x_transposed),
                       n_planes=xn_next.shape[0], plane_rows=xn_next.shape[1], eps=eps, blend=ptr(blend))
    check(_lib.load().mtn_gn_apply_norm_fwd(args, ptr(res_next), ptr(xn_next), ptr(g_next), int(next_transposed), rms_eps,
                                            _stream()), "mtn_gn_apply_norm_fwd")


def dp_segment(x, batch, L, C, K, S, out_a, out_a2=None):
    _req_cuda(x, out_a, out_a2)
    check(_lib.load().mtn_dp_segment_fwd(ptr(x), ptr(out_a), ptr(out_a2), batch, L, C, K, S, _stream()), "mtn_dp_segment_fwd")
    return out_a


def dp_overadd_prelu(X, prelu_w, planes, batch, L, C, K, S):
    _req_cuda(X, prelu_w, planes)
    check(_lib.load().mtn_dp_overadd_prelu_fwd(ptr(X), ptr(prelu_w), ptr(planes), planes.shape[1], batch, L, C, K, S,
                                               planes.shape[0], _stream()), "mtn_dp_overadd_prelu_fwd")
    return planes


def bias_planes(x, bias, bias_scale, planes, rows, C):
    _req_cuda(x, bias, planes)
    check(_lib.load().mtn_bias_planes_fwd(ptr(x), x.stride(0), ptr(bias), bias_scale, ptr(planes), planes.shape[1], rows, C,
                                          planes.shape[0], _stream()), "mtn_bias_planes_fwd")
    return planes


def gate_planes(og, bo, bg, planes, rows, groups, D):
    _req_cuda(og, bo, bg, planes)
    check(_lib.load().mtn_gate_planes_fwd(ptr(og), ptr(bo), ptr(bg), ptr(planes), planes.shape[1], rows, groups, D,
                                          planes.shape[0], _stream()), "mtn_gate_planes_fwd")
    return planes
