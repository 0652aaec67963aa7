// HBM-bound streaming kernels of the Mamba-TasNet path (sm_100a): waveform encoder + channel LayerNorm,
// residual add + RMSNorm, bidirectional depthwise causal conv + SiLU, overlap-add decoder, plane packing.
// All activations channel-last; all global accesses are warp-contiguous (>= 64 B per warp instruction,
// 128-bit per lane where the row length allows).
#include "mtn_ptx.cuh"
#include "mtn_host.h"

namespace mtn {

// Operand-plane store of four consecutive values with PACKED conversions (F2FP.BF16.PACK_AB on the ALU pipe): the scalar
// __float2bfloat16_rn of store_planes4 compiles to F2F, which shares the XU pipe with the SiLU's MUFUs -- in the fp32-mode conv
// 8 of 24 XU operations per (4 channels, step) were conversions.  Same roundings, bit-identical planes.
template <int P>
__device__ __forceinline__ void store_planes4_packed(__nv_bfloat16* base, size_t plane_stride, size_t off, float4 v) {
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&h01);
    pk.y = *reinterpret_cast<const uint32_t*>(&h23);
    *reinterpret_cast<uint2*>(base + off) = pk;
    if (P == 2) {
        const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - __uint_as_float(pk.x << 16), v.y - __uint_as_float(pk.x & 0xffff0000u));
        const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - __uint_as_float(pk.y << 16), v.w - __uint_as_float(pk.y & 0xffff0000u));
        uint2 pl;
        pl.x = *reinterpret_cast<const uint32_t*>(&l01);
        pl.y = *reinterpret_cast<const uint32_t*>(&l23);
        *reinterpret_cast<uint2*>(base + plane_stride + off) = pl;
    }
}

// ------------------------------------------------------------------------------------------------
// Encoder (Conv1d 1->N, k=16, s=8, no bias, ReLU) + ChannelwiseLayerNorm.  One warp per frame.
// Reference: speechbrain dual_path.Encoder == baseline/avse2/model.py:14-24; cLN call
// Mamba-TasNet/modules/mamba_masknet.py:118 (biased variance, eps 1e-8).
// ------------------------------------------------------------------------------------------------
template <int P, int NJ>  // NJ = N / 32 channels per lane
__global__ void __launch_bounds__(256)
encoder_cln_kernel(const float* __restrict__ mix, const float* __restrict__ w_enc, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ mix_w, __nv_bfloat16* __restrict__ yn, int batch,
                   int ld_mix, int L, float eps) {
    constexpr int N = NJ * 32;
    extern __shared__ float s_w[];  // [16][N] transposed filter bank
    for (int i = threadIdx.x; i < 16 * N; i += blockDim.x) {
        const int n = i / 16, k = i % 16;
        s_w[k * N + n] = w_enc[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const size_t tokens = size_t(batch) * L;
    const size_t plane_stride = tokens * N;
    float g[NJ], bt[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        g[j] = gamma[lane + 32 * j];
        bt[j] = beta[lane + 32 * j];
    }
    // TPI tokens per warp iteration: every filter tap fetched from shared memory is used TPI times (one token at a time the
    // kernel was paced by its 16 x NJ scalar LDS per token: 0.109 ms at BASELINE config 2 against 0.04 ms of HBM time).
    constexpr int TPI = NJ >= 16 ? 2 : 4;
    const size_t groups = (tokens + TPI - 1) / TPI;
    for (size_t grp = size_t(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); grp < groups;
         grp += size_t(gridDim.x) * warps_per_block) {
        const size_t tok0 = grp * TPI;
        float x[TPI][16];
#pragma unroll
        for (int i = 0; i < TPI; ++i) {
            const size_t tok = tok0 + i < tokens ? tok0 + i : tokens - 1;   // tail: recompute the last token, store nothing
            const int b = int(tok / L), l = int(tok % L);
            const float4* xp = reinterpret_cast<const float4*>(mix + size_t(b) * ld_mix + size_t(l) * 8);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 f = __ldg(xp + q);
                x[i][4 * q] = f.x;
                x[i][4 * q + 1] = f.y;
                x[i][4 * q + 2] = f.z;
                x[i][4 * q + 3] = f.w;
            }
        }
        float v[TPI][NJ];
        float s[TPI];
#pragma unroll
        for (int i = 0; i < TPI; ++i) s[i] = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            float acc[TPI];
#pragma unroll
            for (int i = 0; i < TPI; ++i) acc[i] = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const float w = s_w[k * N + lane + 32 * j];
#pragma unroll
                for (int i = 0; i < TPI; ++i) acc[i] = fmaf(w, x[i][k], acc[i]);
            }
#pragma unroll
            for (int i = 0; i < TPI; ++i) {
                v[i][j] = fmaxf(acc[i], 0.f);
                s[i] += v[i][j];
            }
        }
#pragma unroll
        for (int i = 0; i < TPI; ++i) {
            const float mean = warp_sum(s[i]) * (1.0f / N);
            float sq = 0.f;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const float d = v[i][j] - mean;
                sq = fmaf(d, d, sq);
            }
            const float rstd = rsqrtf(warp_sum(sq) * (1.0f / N) + eps);
            if (tok0 + i < tokens) {
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const size_t off = (tok0 + i) * N + lane + 32 * j;
                    mix_w[off] = v[i][j];
                    store_planes1<P>(yn, plane_stride, off, fmaf(g[j] * (v[i][j] - mean), rstd, bt[j]));
                }
            }
        }
    }
}

// ChannelwiseLayerNorm alone (stand-alone MaskNet module on an external mix_w).  One warp per token.
template <int P, int NJ>
__global__ void __launch_bounds__(256)
cln_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
           __nv_bfloat16* __restrict__ yn, int M, float eps) {
    constexpr int N = NJ * 32;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const size_t plane_stride = size_t(M) * N;
    for (int tok = blockIdx.x * warps_per_block + (threadIdx.x >> 5); tok < M; tok += gridDim.x * warps_per_block) {
        float v[NJ];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            v[j] = x[size_t(tok) * N + lane + 32 * j];
            s += v[j];
        }
        const float mean = warp_sum(s) * (1.0f / N);
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const float d = v[j] - mean;
            sq = fmaf(d, d, sq);
        }
        const float rstd = rsqrtf(warp_sum(sq) * (1.0f / N) + eps);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int n = lane + 32 * j;
            store_planes1<P>(yn, plane_stride, size_t(tok) * N + n, fmaf(gamma[n] * (v[j] - mean), rstd, beta[n]));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Residual add + RMSNorm -> bf16 planes.  One warp per token, 128-bit accesses.
// Reference: Block.forward Mamba-TasNet/modules/mamba/bimamba.py:446-447; final add + norm_f
// modules/mamba_blocks.py:196-197; RMSNorm math = mamba-ssm rms_norm_ref (fp32, eps 1e-5).
// ------------------------------------------------------------------------------------------------
// LN: nn.LayerNorm instead of RMSNorm (`rms_norm=False`, modules/mamba_blocks.py:36-41: mean removed, biased variance,
// weight and bias) -- the row is already in registers, so the centred second moment costs one more warp reduction.
template <int P, int NV, bool VEC, bool LN>  // VEC: NV float4 per lane (D = 128*NV); else NV floats per lane (D = 32*NV)
__global__ void __launch_bounds__(256)
add_rmsnorm_kernel(const float* __restrict__ h, float* __restrict__ res, int res_valid, const float* __restrict__ g,
                   const float* __restrict__ beta, __nv_bfloat16* __restrict__ xn, float* __restrict__ out_f32, int M,
                   float eps) {
    constexpr int D = VEC ? 128 * NV : 32 * NV;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const size_t plane_stride = size_t(M) * D;
    for (int tok = blockIdx.x * warps_per_block + (threadIdx.x >> 5); tok < M; tok += gridDim.x * warps_per_block) {
        const size_t base = size_t(tok) * D;
        if (VEC) {
            float4 r[NV];
            float sq = 0.f;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const size_t off = base + 128 * j + 4 * lane;
                float4 a = h ? *reinterpret_cast<const float4*>(h + off) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (res_valid) {
                    const float4 b = *reinterpret_cast<const float4*>(res + off);
                    a.x += b.x;
                    a.y += b.y;
                    a.z += b.z;
                    a.w += b.w;
                }
                r[j] = a;
                sq += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
            }
            float mean = 0.f;
            if (LN) {
                float sm = 0.f;
#pragma unroll
                for (int j = 0; j < NV; ++j) sm += (r[j].x + r[j].y) + (r[j].z + r[j].w);
                mean = warp_sum(sm) * (1.0f / D);
                sq = 0.f;
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    const float a0 = r[j].x - mean, a1 = r[j].y - mean, a2 = r[j].z - mean, a3 = r[j].w - mean;
                    sq += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
                }
            }
            const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const size_t off = base + 128 * j + 4 * lane;
                if (h) *reinterpret_cast<float4*>(res + off) = r[j];
                const float4 gg = *reinterpret_cast<const float4*>(g + 128 * j + 4 * lane);
                float4 o;
                if (LN) {
                    const float4 bb = *reinterpret_cast<const float4*>(beta + 128 * j + 4 * lane);
                    o.x = fmaf((r[j].x - mean) * rstd, gg.x, bb.x);
                    o.y = fmaf((r[j].y - mean) * rstd, gg.y, bb.y);
                    o.z = fmaf((r[j].z - mean) * rstd, gg.z, bb.z);
                    o.w = fmaf((r[j].w - mean) * rstd, gg.w, bb.w);
                } else {
                o.x = r[j].x * rstd * gg.x;
                o.y = r[j].y * rstd * gg.y;
                o.z = r[j].z * rstd * gg.z;
                o.w = r[j].w * rstd * gg.w;
                }
                if (xn) store_planes4_packed<P>(xn, plane_stride, off, o);
                if (out_f32) *reinterpret_cast<float4*>(out_f32 + off) = o;
            }
        } else {
            float r[NV];
            float sq = 0.f;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const size_t off = base + 32 * j + lane;
                float a = h ? h[off] : 0.f;
                if (res_valid) a += res[off];
                r[j] = a;
                sq = fmaf(a, a, sq);
            }
            float mean = 0.f;
            if (LN) {
                float sm = 0.f;
#pragma unroll
                for (int j = 0; j < NV; ++j) sm += r[j];
                mean = warp_sum(sm) * (1.0f / D);
                sq = 0.f;
#pragma unroll
                for (int j = 0; j < NV; ++j) sq = fmaf(r[j] - mean, r[j] - mean, sq);
            }
            const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const size_t off = base + 32 * j + lane;
                if (h) res[off] = r[j];
                const float o = LN ? fmaf((r[j] - mean) * rstd, g[32 * j + lane], beta[32 * j + lane])
                                   : r[j] * rstd * g[32 * j + lane];
                if (xn) store_planes1<P>(xn, plane_stride, off, o);
                if (out_f32) out_f32[off] = o;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Depthwise causal conv (width 4) + bias + SiLU, forward AND time-reversed direction in one pass over xs.
// Reference: causal_conv1d_cuda.causal_conv1d_fwd call, Mamba-TasNet/modules/mamba/selective_scan_interface.py:182;
// the backward direction runs the same op on xz.flip(-1) (modules/mamba/bimamba.py:237), i.e. anti-causal taps.
// Thread = 4 channels, walks TT consecutive frames with a 7-row register window.
// ------------------------------------------------------------------------------------------------
template <typename XT>
__device__ __forceinline__ float4 load_x4(const XT* p);
template <>
__device__ __forceinline__ float4 load_x4<float>(const float* p) {
    return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 load_x4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 raw = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}

constexpr int CONV_TT = 32;

// NX rows are requested together per iteration: 8 for both directions from fp32 rows (128 B in flight per thread).  Measured on
// B200 at BASELINE config 2 (tools/conv_bench.py, profiles/r02/conv_variants_S_fp32.jsonl, all bit-identical): (time tile, NX) =
// (32, 4) 0.158 ms = 76 % of the HBM peak, **(32, 8) 0.142 ms = 85 %**, (64, 8) 0.143, (128, 8) 0.148, (24, 12) 0.150, (32, 16)
// 0.178 (182 registers).
template <int P, typename XT, int DIRS, int NX = (DIRS == 3 && sizeof(XT) == 4) ? 8 : 4>   // DIRS: 3 = both directions, 1 = forward only (compile time: keeps the hot both-direction code branch-free)
__global__ void __launch_bounds__(256)
conv_silu_kernel(const XT* __restrict__ xz, int ldxz, const float* __restrict__ conv_w, const float* __restrict__ conv_b,
                 __nv_bfloat16* __restrict__ u, size_t u_rows, const float* __restrict__ halo_lo,
                 const float* __restrict__ halo_hi, int batch, int L, int di) {
    const int c = (blockIdx.z * blockDim.x + threadIdx.x) * 4;
    if (c >= di) return;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * CONV_TT;
    const int t1 = min(t0 + CONV_TT, L);
    const size_t plane_stride = u_rows * 2 * di;
    constexpr bool do_f = DIRS & 1, do_b = DIRS & 2;  // unidirectional stacks run the forward half only
    float4 wf[4], wb[4];  // wf[k] = tap k for channels c..c+3
    {
        float tf[4][4], tb[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(conv_w + size_t(c + i) * 4);
            const float4 bq = do_b ? *reinterpret_cast<const float4*>(conv_w + size_t(di + c + i) * 4) : a;
            tf[i][0] = a.x; tf[i][1] = a.y; tf[i][2] = a.z; tf[i][3] = a.w;
            tb[i][0] = bq.x; tb[i][1] = bq.y; tb[i][2] = bq.z; tb[i][3] = bq.w;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            wf[k] = make_float4(tf[0][k], tf[1][k], tf[2][k], tf[3][k]);
            wb[k] = make_float4(tb[0][k], tb[1][k], tb[2][k], tb[3][k]);
        }
    }
    const float4 bf = *reinterpret_cast<const float4*>(conv_b + c);
    const float4 bb = do_b ? *reinterpret_cast<const float4*>(conv_b + di + c) : bf;
    const XT* xbase = xz + size_t(b) * L * ldxz + c;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    // rows outside [0, L): the neighbouring chunk's rows when a halo is given, else the zero padding of the conv
    const float* hlo = halo_lo ? halo_lo + size_t(b) * 3 * di + c : nullptr;
    const float* hhi = halo_hi ? halo_hi + size_t(b) * 3 * di + c : nullptr;
    auto ld = [&](int t) -> float4 {
        if (t >= 0 && t < L) return load_x4<XT>(xbase + size_t(t) * ldxz);
        if (t < 0) return (hlo && t >= -3) ? *reinterpret_cast<const float4*>(hlo + size_t(t + 3) * di) : zero;
        return (hhi && t < L + 3) ? *reinterpret_cast<const float4*>(hhi + size_t(t - L) * di) : zero;
    };
    // window w[i] = x[t - 3 + i], i = 0..6
    float4 w0 = ld(t0 - 3), w1 = ld(t0 - 2), w2 = ld(t0 - 1), w3 = ld(t0), w4 = ld(t0 + 1), w5 = ld(t0 + 2), w6;
    for (int t = t0; t < t1; t += NX) {
        float4 nx[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) nx[i] = ld(t + 3 + i);
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            w6 = nx[i];
            if (t + i < t1) {
                float4 f, r;
                f.x = fmaf(wf[0].x, w0.x, fmaf(wf[1].x, w1.x, fmaf(wf[2].x, w2.x, fmaf(wf[3].x, w3.x, bf.x))));
                f.y = fmaf(wf[0].y, w0.y, fmaf(wf[1].y, w1.y, fmaf(wf[2].y, w2.y, fmaf(wf[3].y, w3.y, bf.y))));
                f.z = fmaf(wf[0].z, w0.z, fmaf(wf[1].z, w1.z, fmaf(wf[2].z, w2.z, fmaf(wf[3].z, w3.z, bf.z))));
                f.w = fmaf(wf[0].w, w0.w, fmaf(wf[1].w, w1.w, fmaf(wf[2].w, w2.w, fmaf(wf[3].w, w3.w, bf.w))));
                // time-reversed direction: tap k multiplies x[t + 3 - k]
                r.x = fmaf(wb[0].x, w6.x, fmaf(wb[1].x, w5.x, fmaf(wb[2].x, w4.x, fmaf(wb[3].x, w3.x, bb.x))));
                r.y = fmaf(wb[0].y, w6.y, fmaf(wb[1].y, w5.y, fmaf(wb[2].y, w4.y, fmaf(wb[3].y, w3.y, bb.y))));
                r.z = fmaf(wb[0].z, w6.z, fmaf(wb[1].z, w5.z, fmaf(wb[2].z, w4.z, fmaf(wb[3].z, w3.z, bb.z))));
                r.w = fmaf(wb[0].w, w6.w, fmaf(wb[1].w, w5.w, fmaf(wb[2].w, w4.w, fmaf(wb[3].w, w3.w, bb.w))));
                constexpr bool FS = (P == 1);   // one bf16 plane: the result is rounded to 8 mantissa bits anyway
                if (do_f) f = make_float4(silu_sel<FS>(f.x), silu_sel<FS>(f.y), silu_sel<FS>(f.z), silu_sel<FS>(f.w));
                if (do_b) r = make_float4(silu_sel<FS>(r.x), silu_sel<FS>(r.y), silu_sel<FS>(r.z), silu_sel<FS>(r.w));
                const size_t off = (size_t(b) * L + (t + i)) * (2 * di) + c;
                if (do_f) store_planes4_packed<P>(u, plane_stride, off, f);
                if (do_b) store_planes4_packed<P>(u, plane_stride, off + di, r);
            }
            w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; w5 = w6;
        }
    }
}

// bf16 mode (bf16 xz, one output plane): 8 channels per thread so that every global access is 16 bytes (the 4-channel
// kernel above moves 8 bytes per access there and reached only 2.1 TB/s at BASELINE config 3).  One direction per CTA
// (blockIdx.z: low bit = direction when both are requested) keeps the register footprint at 32 weights + a 4-row fp32
// window; the second read of xs hits L1/L2.  Same fmaf nesting as conv_silu_kernel -> identical results.
// Time tile 64: the per-CTA start-up (32 weight rows, bias, three halo rows per thread) is paid half as often as with 32 rows.
// Measured on B200 (L hparams, 64 x 3999, profiles/r02/conv_variants_L_bf16.jsonl, bit-identical): 32 rows 0.411 ms, **64 rows
// 0.381 ms** (63 % of the HBM peak; ncu: issue 59 %, ALU pipe 46 %, 16 resident warps per SM at 128 registers); 16 rows in flight
// instead of 8 (184 registers) 0.73 ms, double-buffered row blocks 0.39 ms -- neither the bytes in flight nor load latency bound it.
constexpr int CONV8_TT = 64;

__device__ __forceinline__ void unpack_bf16x8(const uint4 raw, float (&v)[8]) {
    const uint32_t r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(r[i] << 16);
        v[2 * i + 1] = __uint_as_float(r[i] & 0xffff0000u);
    }
}

// `dir` is a template parameter: indexing the register window with a run-time direction sent it to local memory (stack
// frame, 19 % of HBM peak); the kernel below branches once per CTA into the two instantiations.
template <int dir, bool BOTH_>
__device__ __forceinline__ void conv_silu_bf16x8_body(const __nv_bfloat16* __restrict__ xz, int ldxz,
                                                      const float* __restrict__ conv_w, const float* __restrict__ conv_b,
                                                      __nv_bfloat16* __restrict__ u, const float* __restrict__ halo_lo,
                                                      const float* __restrict__ halo_hi, int L, int di, int cblk) {
    const int c = (cblk * blockDim.x + threadIdx.x) * 8;
    if (c >= di) return;
    const int b = blockIdx.y;
    const int t0 = int(blockIdx.x >> (BOTH_ ? 1 : 0)) * CONV8_TT;
    const int t1 = min(t0 + CONV8_TT, L);
    float w[4][8], bias[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(conv_w + (size_t(dir) * di + c + i) * 4);
        w[0][i] = a.x; w[1][i] = a.y; w[2][i] = a.z; w[3][i] = a.w;
        bias[i] = conv_b[size_t(dir) * di + c + i];
    }
    const __nv_bfloat16* xbase = xz + size_t(b) * L * ldxz + c;
    const float* hlo = halo_lo ? halo_lo + size_t(b) * 3 * di + c : nullptr;
    const float* hhi = halo_hi ? halo_hi + size_t(b) * 3 * di + c : nullptr;
    auto ld_halo = [&](int t, float (&v)[8]) {      // rows outside [0, L): neighbouring chunk's rows or zero padding
        const float* h = (t < 0) ? ((hlo && t >= -3) ? hlo + size_t(t + 3) * di : nullptr)
                                 : ((hhi && t < L + 3) ? hhi + size_t(t - L) * di : nullptr);
        if (h) {
            const float4 a = *reinterpret_cast<const float4*>(h), q = *reinterpret_cast<const float4*>(h + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
    };
    auto ld = [&](int t, float (&v)[8]) {
        if (t >= 0 && t < L) unpack_bf16x8(*reinterpret_cast<const uint4*>(xbase + size_t(t) * ldxz), v);
        else ld_halo(t, v);
    };
    // forward: out[t] = sum_k w[k] x[t-3+k] -> window rows r[k] = x[t-3+k]; backward: out[t] = sum_k w[k] x[t+3-k] ->
    // r[k] = x[t+3-k].  Either way the window slides by one row per step in the direction of increasing t: forward drops
    // r[0] and appends x[t] as r[3]; backward drops r[3] and prepends x[t+3] as r[0].
    // The four live rows sit in a ring of four register rows addressed by compile-time indices (the row loop is unrolled PF = 8
    // steps, a multiple of the ring size, so the assignment repeats every outer iteration): forward, x[t0 - 3 + m] lives in
    // slot m & 3, step j appends x[t] to slot (j + 3) & 3 and tap k reads slot (j + k) & 3; backward, x[t0 + m] lives in slot
    // m & 3, step j appends x[t + 3] to slot (j + 3) & 3 and tap k reads slot (j + 3 - k) & 3.  (Shifting a 4 x 8 window by
    // register moves cost 24 MOV per row next to 40 FFMA: 373 of the kernel's ~2 300 instructions.)
    float r[4][8];
#pragma unroll
    for (int k = 0; k < 3; ++k) ld(dir ? (t0 + k) : (t0 - 3 + k), r[k]);
    __nv_bfloat16* ubase = u + (size_t(b) * L) * (2 * size_t(di)) + size_t(dir) * di + c;
    // PF rows are requested together (PF x 16 B in flight per thread: with 16 resident warps per SM a single outstanding
    // load per thread left the kernel latency-bound at 2.2 TB/s), then consumed one step at a time.
    constexpr int PF = 8;
    static_assert(PF % 4 == 0 && CONV8_TT % PF == 0, "the register ring needs whole periods per outer iteration");
    constexpr int shift = dir ? 3 : 0;                 // the new row of step t is x[t + shift]
#pragma unroll 1
    for (int tb = t0; tb < t1; tb += PF) {
        uint4 raw[PF];
#pragma unroll
        for (int j = 0; j < PF; ++j) {
            const int tr = tb + j + shift;
            raw[j] = (tb + j < t1 && tr < L) ? *reinterpret_cast<const uint4*>(xbase + size_t(tr) * ldxz)
                                             : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int j = 0; j < PF; ++j) {
            const int t = tb + j;
            if (t < t1) {
                const int tr = t + shift;
                float(&rn)[8] = r[(j + 3) & 3];
                if (tr < L) unpack_bf16x8(raw[j], rn);
                else ld_halo(tr, rn);
                const float(&r0)[8] = r[dir ? (j + 3) & 3 : j & 3];            // tap 0 .. tap 3 rows
                const float(&r1)[8] = r[dir ? (j + 2) & 3 : (j + 1) & 3];
                const float(&r2)[8] = r[dir ? (j + 1) & 3 : (j + 2) & 3];
                const float(&r3)[8] = r[dir ? j & 3 : (j + 3) & 3];
                uint32_t pk[4];
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    float f0 = fmaf(w[0][i], r0[i], fmaf(w[1][i], r1[i], fmaf(w[2][i], r2[i], fmaf(w[3][i], r3[i], bias[i]))));
                    float f1 = fmaf(w[0][i + 1], r0[i + 1],
                                    fmaf(w[1][i + 1], r1[i + 1], fmaf(w[2][i + 1], r2[i + 1], fmaf(w[3][i + 1], r3[i + 1], bias[i + 1]))));
                    f0 = silu_bf16_f(f0);
                    f1 = silu_bf16_f(f1);
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(f0, f1);
                    pk[i / 2] = *reinterpret_cast<const uint32_t*>(&h2);
                }
                *reinterpret_cast<uint4*>(ubase + size_t(t) * (2 * size_t(di))) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
    }
}

template <bool BOTH>
__global__ void __launch_bounds__(128, 4)
conv_silu_bf16x8_kernel(const __nv_bfloat16* __restrict__ xz, int ldxz, const float* __restrict__ conv_w,
                        const float* __restrict__ conv_b, __nv_bfloat16* __restrict__ u, const float* __restrict__ halo_lo,
                        const float* __restrict__ halo_hi, int L, int di) {
    // both directions of a time tile are neighbours in launch order (blockIdx.x = 2 * tile + dir), so the second read
    // of the tile's xs rows hits L2 instead of DRAM (direction-major order read xs twice from DRAM: ncu 1.06 vs 0.52 GB)
    if (BOTH && (blockIdx.x & 1))
        conv_silu_bf16x8_body<1, BOTH>(xz, ldxz, conv_w, conv_b, u, halo_lo, halo_hi, L, di, blockIdx.z);
    else
        conv_silu_bf16x8_body<0, BOTH>(xz, ldxz, conv_w, conv_b, u, halo_lo, halo_hi, L, di, blockIdx.z);
}

// ------------------------------------------------------------------------------------------------
// Decoder: ConvTranspose1d(N -> 1, k=16, s=8, no bias) per speaker = frame GEMV + overlap-add.
// Reference: speechbrain dual_path.Decoder == baseline/avse2/model.py:27-37; speaker loop, cat and
// pad/trim Mamba-TasNet/train_wsj0mix.py:95-109.
// ------------------------------------------------------------------------------------------------
constexpr int DEC_WLD = 20;  // padded filter row (16 taps + 4): conflict-free float4 reads

template <int NJ>
__global__ void __launch_bounds__(256)
decoder_frames_kernel(const float* __restrict__ sep, const float* __restrict__ w_dec, float* __restrict__ frames,
                      size_t items) {
    constexpr int N = NJ * 32;
    extern __shared__ float s_w[];  // [N][DEC_WLD]
    for (int i = threadIdx.x; i < N * 16; i += blockDim.x) s_w[(i / 16) * DEC_WLD + (i % 16)] = w_dec[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    // One warp per group of RPW = 4 (token, speaker) rows: every filter tap read from shared memory serves four rows (the kernel
    // is paced by those 64-byte-per-lane reads: 16 shared-memory wavefronts per 32 channels), and the 16 per-lane partial sums of
    // each PAIR of rows are reduced by a transposing butterfly (offsets 16, 8, 4, 2, 1; every step halves the values a lane
    // keeps) -- 31 shuffles per pair instead of 80 per row, the same summation tree as a plain warp sum (bit-identical), and
    // lane l ends up with tap l & 15 of row l >> 4 of the pair: one coalesced 128-byte store per pair.
    constexpr int RPW = 4;
    const size_t groups = (items + RPW - 1) / RPW;
    for (size_t grp = size_t(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); grp < groups;
         grp += size_t(gridDim.x) * warps_per_block) {
        const size_t item0 = RPW * grp;
        const float* row[RPW];   // sep[token][s*N + n] -> contiguous in (token, s); rows past the end repeat the first one
#pragma unroll
        for (int i = 0; i < RPW; ++i) row[i] = sep + (item0 + i < items ? item0 + i : item0) * N;
        float a[RPW][16];
#pragma unroll
        for (int i = 0; i < RPW; ++i)
#pragma unroll
            for (int k = 0; k < 16; ++k) a[i][k] = 0.f;
#pragma unroll 2
        for (int j = 0; j < NJ; ++j) {
            const int n = lane + 32 * j;
            float v[RPW];
#pragma unroll
            for (int i = 0; i < RPW; ++i) v[i] = row[i][n];
            const float4* wp = reinterpret_cast<const float4*>(&s_w[n * DEC_WLD]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 w = wp[q];
#pragma unroll
                for (int i = 0; i < RPW; ++i) {
                    a[i][4 * q] = fmaf(v[i], w.x, a[i][4 * q]);
                    a[i][4 * q + 1] = fmaf(v[i], w.y, a[i][4 * q + 1]);
                    a[i][4 * q + 2] = fmaf(v[i], w.z, a[i][4 * q + 2]);
                    a[i][4 * q + 3] = fmaf(v[i], w.w, a[i][4 * q + 3]);
                }
            }
        }
        const bool up16 = lane & 16, up8 = lane & 8, up4 = lane & 4, up2 = lane & 2, up1 = lane & 1;
#pragma unroll
        for (int pq = 0; pq < RPW / 2; ++pq) {
            const float(&a0)[16] = a[2 * pq];
            const float(&a1)[16] = a[2 * pq + 1];
            float x[16], y[8], z[4], u2[2];
#pragma unroll
            for (int k = 0; k < 16; ++k)   // lanes 0..15 keep the pair's first row, lanes 16..31 its second
                x[k] = (up16 ? a1[k] : a0[k]) + __shfl_xor_sync(0xffffffffu, up16 ? a0[k] : a1[k], 16);
#pragma unroll
            for (int k = 0; k < 8; ++k) y[k] = (up8 ? x[8 + k] : x[k]) + __shfl_xor_sync(0xffffffffu, up8 ? x[k] : x[8 + k], 8);
#pragma unroll
            for (int k = 0; k < 4; ++k) z[k] = (up4 ? y[4 + k] : y[k]) + __shfl_xor_sync(0xffffffffu, up4 ? y[k] : y[4 + k], 4);
#pragma unroll
            for (int k = 0; k < 2; ++k) u2[k] = (up2 ? z[2 + k] : z[k]) + __shfl_xor_sync(0xffffffffu, up2 ? z[k] : z[2 + k], 2);
            const float mine = (up1 ? u2[1] : u2[0]) + __shfl_xor_sync(0xffffffffu, up1 ? u2[0] : u2[1], 1);
            const size_t it0 = item0 + 2 * pq;
            if (it0 + (lane >> 4) < items) frames[it0 * 16 + lane] = mine;
        }
    }
}

__global__ void __launch_bounds__(256)
decoder_ola_kernel(const float* __restrict__ frames, float* __restrict__ est, int batch, int T, int L, int S) {
    const size_t total = size_t(batch) * T * S;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const int s = int(i % S);
        const size_t bt = i / S;
        const int t = int(bt % T);
        const int b = int(bt / T);
        const int l1 = t >> 3, k = t & 7;
        float v = 0.f;
        if (l1 < L) v += frames[((size_t(b) * L + l1) * S + s) * 16 + k];
        if (l1 >= 1 && l1 - 1 < L) v += frames[((size_t(b) * L + l1 - 1) * S + s) * 16 + 8 + k];
        est[i] = v;
    }
}

// Streaming overlap-add: this chunk's F frames finalise samples [0, 8F); the first 8 of them also receive the second
// half of the previous chunk's last frame (tail, in), and the second half of frame F-1 becomes the new tail (out).
// The thread that reads tail[b][s][k] is the one that overwrites it.
__global__ void __launch_bounds__(256)
decoder_ola_stream_kernel(const float* __restrict__ frames, float* __restrict__ est, float* __restrict__ tail, int batch,
                          int F, int S) {
    const int T = 8 * F;
    const size_t total = size_t(batch) * T * S;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const int s = int(i % S);
        const size_t bt = i / S;
        const int t = int(bt % T);
        const int b = int(bt / T);
        const int l1 = t >> 3, k = t & 7;
        float v = frames[((size_t(b) * F + l1) * S + s) * 16 + k];
        if (l1 >= 1) {
            v += frames[((size_t(b) * F + l1 - 1) * S + s) * 16 + 8 + k];
        } else {
            float* tp = tail + (size_t(b) * S + s) * 8 + k;
            v += *tp;
            *tp = frames[((size_t(b) * F + F - 1) * S + s) * 16 + 8 + k];
        }
        est[i] = v;
    }
}

// mask_nonlinear = "softmax" (modules/mamba_masknet.py:133-134): `F.softmax(score, dim=2)` on score [n_spk, B, N, L], i.e.
// over the N encoder channels of one (speaker, utterance, frame).  One warp per (frame, speaker); in place; optionally
// times mix_w (the mask application of train_wsj0mix.py:91-92).
template <int NJ>
__global__ void __launch_bounds__(256)
softmax_mask_kernel(float* __restrict__ score, const float* __restrict__ mix_w, size_t items, int S) {
    constexpr int N = 32 * NJ;
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (size_t item = size_t(blockIdx.x) * wpb + (threadIdx.x >> 5); item < items; item += size_t(gridDim.x) * wpb) {
        float* row = score + item * N;
        float v[NJ];
        float m = -3.0e38f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            v[j] = row[lane + 32 * j];
            m = fmaxf(m, v[j]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            v[j] = ex2_approx(1.4426950408889634f * (v[j] - m));
            sum += v[j];
        }
        const float inv = 1.0f / warp_sum(sum);
        const float* mw = mix_w ? mix_w + (item / S) * N : nullptr;
#pragma unroll
        for (int j = 0; j < NJ; ++j) row[lane + 32 * j] = v[j] * inv * (mw ? mw[lane + 32 * j] : 1.0f);
    }
}

template <int P>
__global__ void __launch_bounds__(256)
split_planes_kernel(const float* __restrict__ src, int ld, __nv_bfloat16* __restrict__ dst, int rows, int cols) {
    const size_t total = size_t(rows) * cols;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const size_t r = i / cols, c = i % cols;
        store_planes1<P>(dst, total, i, src[r * ld + c]);
    }
}

static int grid_for(size_t items, int per_block, int waves = 8) {
    size_t need = (items + per_block - 1) / per_block;
    size_t cap = size_t(num_sms()) * waves;
    size_t g = need < cap ? need : cap;
    return g < 1 ? 1 : int(g);
}

}  // namespace mtn

using namespace mtn;

extern "C" int mtn_encoder_cln_fwd(const float* mix, int ld_mix, const float* w_enc, const float* gamma,
                                   const float* beta, float* mix_w, void* yn_planes, int batch, int T, int L, int N,
                                   int planes, float eps, mtn_stream_t stream) {
    MTN_REQUIRE(mix && w_enc && gamma && beta && mix_w && yn_planes, "encoder: null pointer");
    MTN_REQUIRE(batch > 0 && T >= 16 && L == (T - 16) / 8 + 1, "encoder: bad shape batch=%d T=%d L=%d", batch, T, L);
    MTN_REQUIRE(ld_mix >= T && ld_mix % 4 == 0 && (reinterpret_cast<uintptr_t>(mix) & 15) == 0,
                "encoder: ld_mix must be >= T and a multiple of 4, mix 16-byte aligned (128-bit frame loads)");
    MTN_REQUIRE(planes == 1 || planes == 2, "encoder: planes=%d", planes);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const size_t tokens = size_t(batch) * L;
    const int grid = grid_for(tokens, 8, 4);
    const size_t smem = size_t(16) * N * sizeof(float);
#define MTN_ENC(PP, NJ)                                                                                          \
    do {                                                                                                         \
        auto k = encoder_cln_kernel<PP, NJ>;                                                                     \
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
        k<<<grid, 256, smem, s>>>(mix, w_enc, gamma, beta, mix_w, reinterpret_cast<__nv_bfloat16*>(yn_planes),  \
                                  batch, ld_mix, L, eps);                                                           \
    } while (0)
#define MTN_ENC_N(PP)                                                         \
    switch (N) {                                                              \
        case 64: MTN_ENC(PP, 2); break;                                       \
        case 128: MTN_ENC(PP, 4); break;                                      \
        case 256: MTN_ENC(PP, 8); break;                                      \
        case 512: MTN_ENC(PP, 16); break;                                     \
        default: set_error("encoder: unsupported N=%d", N); return MTN_EINVAL; \
    }
    if (planes == 2) { MTN_ENC_N(2) } else { MTN_ENC_N(1) }
#undef MTN_ENC_N
#undef MTN_ENC
    MTN_CUDA_LAUNCH_CHECK("encoder_cln");
    return MTN_OK;
}

extern "C" int mtn_cln_fwd(const float* x, const float* gamma, const float* beta, void* yn_planes, int M, int N,
                           int planes, float eps, mtn_stream_t stream) {
    MTN_REQUIRE(x && gamma && beta && yn_planes && M > 0, "cln: bad arguments");
    MTN_REQUIRE(planes == 1 || planes == 2, "cln: planes=%d", planes);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int grid = grid_for(size_t(M), 8, 8);
    __nv_bfloat16* yn = reinterpret_cast<__nv_bfloat16*>(yn_planes);
#define MTN_CLN(PP, NJ) cln_kernel<PP, NJ><<<grid, 256, 0, s>>>(x, gamma, beta, yn, M, eps)
#define MTN_CLN_N(PP)                                                     \
    switch (N) {                                                          \
        case 64: MTN_CLN(PP, 2); break;                                   \
        case 128: MTN_CLN(PP, 4); break;                                  \
        case 256: MTN_CLN(PP, 8); break;                                  \
        case 512: MTN_CLN(PP, 16); break;                                 \
        default: set_error("cln: unsupported N=%d", N); return MTN_EINVAL; \
    }
    if (planes == 2) { MTN_CLN_N(2) } else { MTN_CLN_N(1) }
#undef MTN_CLN_N
#undef MTN_CLN
    MTN_CUDA_LAUNCH_CHECK("cln");
    return MTN_OK;
}

extern "C" int mtn_add_rmsnorm_fwd(const float* h, float* res, int res_valid, const float* g, void* xn_planes, int M,
                                   int D, int planes, float eps, mtn_stream_t stream) {
    MTN_REQUIRE(xn_planes, "add_rmsnorm: null pointer");
    return mtn_add_rmsnorm_out_fwd(h, res, res_valid, g, xn_planes, nullptr, M, D, planes, eps, stream);
}

extern "C" int mtn_add_rmsnorm_out_fwd(const float* h, float* res, int res_valid, const float* g, void* xn_planes,
                                       float* out_f32, int M, int D, int planes, float eps, mtn_stream_t stream) {
    return mtn_add_norm_fwd(h, res, res_valid, g, nullptr, xn_planes, out_f32, M, D, planes, eps, stream);
}

extern "C" int mtn_add_norm_fwd(const float* h, float* res, int res_valid, const float* g, const float* beta, void* xn_planes,
                                float* out_f32, int M, int D, int planes, float eps, mtn_stream_t stream) {
    MTN_REQUIRE(res && g && (xn_planes || out_f32), "add_rmsnorm: null pointer");
    MTN_REQUIRE(h || res_valid, "add_rmsnorm: need h or a valid residual");
    MTN_REQUIRE(M > 0 && planes >= 1 && planes <= 2, "add_rmsnorm: bad M=%d planes=%d", M, planes);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int grid = grid_for(size_t(M), 8, 8);
    __nv_bfloat16* xn = reinterpret_cast<__nv_bfloat16*>(xn_planes);
#define MTN_RMS(PP, NV, VEC)                                                                                          \
    do {                                                                                                             \
        if (beta) add_rmsnorm_kernel<PP, NV, VEC, true><<<grid, 256, 0, s>>>(h, res, res_valid, g, beta, xn, out_f32, M, eps);  \
        else add_rmsnorm_kernel<PP, NV, VEC, false><<<grid, 256, 0, s>>>(h, res, res_valid, g, beta, xn, out_f32, M, eps);      \
    } while (0)
#define MTN_RMS_D(PP)                                                             \
    switch (D) {                                                                  \
        case 64: MTN_RMS(PP, 2, false); break;                                    \
        case 128: MTN_RMS(PP, 1, true); break;                                    \
        case 256: MTN_RMS(PP, 2, true); break;                                    \
        case 512: MTN_RMS(PP, 4, true); break;                                    \
        default: set_error("add_rmsnorm: unsupported D=%d", D); return MTN_EINVAL; \
    }
    if (planes == 2) { MTN_RMS_D(2) } else { MTN_RMS_D(1) }
#undef MTN_RMS_D
#undef MTN_RMS
    MTN_CUDA_LAUNCH_CHECK("add_rmsnorm");
    return MTN_OK;
}

extern "C" int mtn_conv_silu_fwd(const void* xz, int ldxz, int xz_bf16, const float* conv_w, const float* conv_b,
                                 void* u_planes, int batch, int L, int di, int planes, mtn_stream_t stream) {
    return mtn_conv_silu_halo_fwd(xz, ldxz, xz_bf16, conv_w, conv_b, u_planes, batch * L, nullptr, nullptr, batch, L, di,
                                  planes, stream);
}

extern "C" int mtn_conv_silu_halo_fwd(const void* xz, int ldxz, int xz_bf16, const float* conv_w, const float* conv_b,
                                      void* u_planes, int u_rows, const float* halo_lo, const float* halo_hi,
                                      int batch, int L, int di, int planes, mtn_stream_t stream) {
    return mtn_conv_silu_dir_fwd(xz, ldxz, xz_bf16, conv_w, conv_b, u_planes, u_rows, halo_lo, halo_hi, batch, L, di,
                                 planes, 3, stream);
}

extern "C" int mtn_conv_silu_dir_fwd(const void* xz, int ldxz, int xz_bf16, const float* conv_w, const float* conv_b,
                                     void* u_planes, int u_rows, const float* halo_lo, const float* halo_hi,
                                     int batch, int L, int di, int planes, int dir_mask, mtn_stream_t stream) {
    MTN_REQUIRE(xz && conv_w && conv_b && u_planes, "conv_silu: null pointer");
    MTN_REQUIRE(dir_mask >= 1 && dir_mask <= 3, "conv_silu: dir_mask=%d", dir_mask);
    MTN_REQUIRE(batch > 0 && L > 0 && di > 0 && di % 4 == 0 && ldxz % 4 == 0, "conv_silu: bad shape");
    MTN_REQUIRE(planes == 1 || planes == 2, "conv_silu: planes=%d", planes);
    MTN_REQUIRE(batch <= 65535, "conv_silu: batch too large for grid.y");
    MTN_REQUIRE(u_rows >= batch * L, "conv_silu: u_rows=%d < batch*L=%d", u_rows, batch * L);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int threads_needed = di / 4;
    const int block = threads_needed >= 256 ? 256 : ((threads_needed + 31) / 32) * 32;
    dim3 grid((L + CONV_TT - 1) / CONV_TT, batch, (threads_needed + block - 1) / block);
    __nv_bfloat16* u = reinterpret_cast<__nv_bfloat16*>(u_planes);
    MTN_REQUIRE(dir_mask == 3 || dir_mask == 1, "conv_silu: dir_mask=%d (3 = both, 1 = forward only)", dir_mask);
    if (xz_bf16 && planes == 1 && di % 8 == 0 && ldxz % 8 == 0 && u_rows == batch * L &&
        (reinterpret_cast<uintptr_t>(xz) & 15) == 0 && (reinterpret_cast<uintptr_t>(u_planes) & 15) == 0) {
        const int tn = di / 8;
        const int blk = tn >= 128 ? 128 : ((tn + 31) / 32) * 32;
        const int cblks = (tn + blk - 1) / blk;
        dim3 g8(((L + CONV8_TT - 1) / CONV8_TT) * (dir_mask == 3 ? 2 : 1), batch, cblks);
        const __nv_bfloat16* x8 = reinterpret_cast<const __nv_bfloat16*>(xz);
        if (dir_mask == 3) conv_silu_bf16x8_kernel<true><<<g8, blk, 0, s>>>(x8, ldxz, conv_w, conv_b, u, halo_lo, halo_hi, L, di);
        else conv_silu_bf16x8_kernel<false><<<g8, blk, 0, s>>>(x8, ldxz, conv_w, conv_b, u, halo_lo, halo_hi, L, di);
        MTN_CUDA_LAUNCH_CHECK("conv_silu(bf16x8)");
        return MTN_OK;
    }
#define MTN_CONV(PP, XT_, DD) \
    conv_silu_kernel<PP, XT_, DD><<<grid, block, 0, s>>>(reinterpret_cast<const XT_*>(xz), ldxz, conv_w, conv_b, u, size_t(u_rows), halo_lo, halo_hi, batch, L, di)
#define MTN_CONV_D(PP, XT_) do { if (dir_mask == 3) MTN_CONV(PP, XT_, 3); else MTN_CONV(PP, XT_, 1); } while (0)
    if (xz_bf16) {
        if (planes == 2) MTN_CONV_D(2, __nv_bfloat16); else MTN_CONV_D(1, __nv_bfloat16);
    } else {
        if (planes == 2) MTN_CONV_D(2, float); else MTN_CONV_D(1, float);
    }
#undef MTN_CONV_D
#undef MTN_CONV
    MTN_CUDA_LAUNCH_CHECK("conv_silu");
    return MTN_OK;
}

extern "C" int mtn_decoder_fwd(const float* sep, const float* w_dec, float* frames, float* est, int batch, int T, int L,
                               int N, int n_spk, mtn_stream_t stream) {
    return mtn_decoder_stream_fwd(sep, w_dec, frames, est, nullptr, batch, T, L, N, n_spk, stream);
}

extern "C" int mtn_decoder_stream_fwd(const float* sep, const float* w_dec, float* frames, float* est, float* tail,
                                      int batch, int T, int L, int N, int n_spk, mtn_stream_t stream) {
    MTN_REQUIRE(sep && w_dec && frames && est, "decoder: null pointer");
    MTN_REQUIRE(batch > 0 && T > 0 && L > 0 && n_spk >= 1, "decoder: bad shape");
    MTN_REQUIRE(!tail || T == 8 * L, "decoder(stream): a chunk of L frames finalises exactly 8*L samples (T=%d, L=%d)", T, L);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const size_t tokens = size_t(batch) * L;
    const int grid = grid_for(tokens * n_spk, 8, 4);
    const size_t smem = size_t(N) * DEC_WLD * sizeof(float);
#define MTN_DEC(NJ)                                                                                            \
    do {                                                                                                       \
        auto k = decoder_frames_kernel<NJ>;                                                                    \
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        k<<<grid, 256, smem, s>>>(sep, w_dec, frames, tokens * n_spk);                                                \
    } while (0)
    switch (N) {
        case 64: MTN_DEC(2); break;
        case 128: MTN_DEC(4); break;
        case 256: MTN_DEC(8); break;
        case 512: MTN_DEC(16); break;
        default: set_error("decoder: unsupported N=%d", N); return MTN_EINVAL;
    }
#undef MTN_DEC
    MTN_CUDA_LAUNCH_CHECK("decoder_frames");
    if (tail) decoder_ola_stream_kernel<<<grid_for(size_t(batch) * T * n_spk, 256, 8), 256, 0, s>>>(frames, est, tail, batch, L, n_spk);
    else decoder_ola_kernel<<<grid_for(size_t(batch) * T * n_spk, 256, 8), 256, 0, s>>>(frames, est, batch, T, L, n_spk);
    MTN_CUDA_LAUNCH_CHECK("decoder_ola");
    return MTN_OK;
}

extern "C" int mtn_softmax_mask_fwd(float* score, const float* mix_w, int rows, int N, int n_spk, mtn_stream_t stream) {
    MTN_REQUIRE(score && rows > 0 && n_spk >= 1, "softmax_mask: bad arguments");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const size_t items = size_t(rows) * n_spk;
    const int grid = grid_for(items, 8, 8);
    switch (N) {
        case 64: softmax_mask_kernel<2><<<grid, 256, 0, s>>>(score, mix_w, items, n_spk); break;
        case 128: softmax_mask_kernel<4><<<grid, 256, 0, s>>>(score, mix_w, items, n_spk); break;
        case 256: softmax_mask_kernel<8><<<grid, 256, 0, s>>>(score, mix_w, items, n_spk); break;
        case 512: softmax_mask_kernel<16><<<grid, 256, 0, s>>>(score, mix_w, items, n_spk); break;
        default: set_error("softmax_mask: unsupported N=%d", N); return MTN_EINVAL;
    }
    MTN_CUDA_LAUNCH_CHECK("softmax_mask");
    return MTN_OK;
}

extern "C" int mtn_split_planes(const float* src, int ld, void* dst_planes, int rows, int cols, int planes,
                                mtn_stream_t stream) {
    MTN_REQUIRE(src && dst_planes && rows > 0 && cols > 0 && ld >= cols, "split_planes: bad arguments");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int grid = grid_for(size_t(rows) * cols, 256, 8);
    __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst_planes);
    if (planes == 2) split_planes_kernel<2><<<grid, 256, 0, s>>>(src, ld, d, rows, cols);
    else if (planes == 1) split_planes_kernel<1><<<grid, 256, 0, s>>>(src, ld, d, rows, cols);
    else { set_error("split_planes: planes=%d", planes); return MTN_EINVAL; }
    MTN_CUDA_LAUNCH_CHECK("split_planes");
    return MTN_OK;
}
