// Depthwise conv + SiLU (both directions) fused with the x_proj contraction (sm_100a).
//
// The unfused plan writes u = silu(conv1d(x)) for both directions (conv_silu_kernel) and reads it back twice: once by the
// x_proj GEMM (TMA) and once by the scan.  Here the conv warps of a CTA produce each 128-token x 64-channel piece of u
// ONCE: into global memory for the scan and -- as bf16 hi / lo planes in the 128-byte-swizzled K-major layout a TMA box would
// have produced -- into shared memory as the A operand of `tcgen05.mma`; W_x chunks arrive by TMA, [dt | B | C] accumulates in
// TMEM over the channel chunks and leaves through `tcgen05.ld`.  u is never re-read by a GEMM: 6.5 KB instead of 10.6 KB of
// HBM traffic per token at S hparams (fp32 mode).
//
// Replaces `causal_conv1d_cuda.causal_conv1d_fwd` x 2 + `xz.flip` + `F.linear(x_proj)` x 2
// (Mamba-TasNet/modules/mamba/selective_scan_interface.py:182-186, modules/mamba/bimamba.py:237).
// Same arithmetic as the unfused pair, in the same order: the conv uses conv_silu_kernel's fmaf nesting and the MMAs run in
// mtn_gemm's order (k-block, k16 step, hi*hi + lo*hi + hi*lo), so u and dbl are bit-identical to the two-kernel plan.
//
// Roles (320 threads): warps 0..7 conv producers + epilogue, warp 8 MMA issuer (one thread), warp 9 TMA producer of the W_x
// chunks (one thread).  Two A stages (2 directions x P planes x 16 KB each), two W stages.
#include "mtn_ptx.cuh"
#include "mtn_host.h"

namespace mtn {

constexpr int CX_BM = 128;        // tokens per tile (one utterance, consecutive frames)
constexpr int CX_BK = 64;         // channels per chunk = one 128-byte swizzled row of bf16
constexpr int CX_CONV_THREADS = 256;
constexpr int CX_THREADS = CX_CONV_THREADS + 64;
constexpr int CX_ASTAGES = 2;
constexpr int CX_WSTAGES = 2;
constexpr int CX_ACC_STRIDE = 64;  // TMEM columns between the two directions' accumulators

struct ConvXParams {
    const void* xz;
    const float* conv_w;   // [2][di][4]
    const float* conv_b;   // [2][di]
    __nv_bfloat16* u;      // [P][u_rows][2*di]
    float* dbl;            // [M][ld_dbl]: direction d at columns d*nd
    size_t u_plane;        // elements between the planes of u
    int ldxz, ld_dbl, nd, batch, L, di, tiles_per_seq, total_tiles;
};

template <int P, int ND>
struct ConvXCfg {
    static constexpr int A_TILE = CX_BM * CX_BK * 2;              // 16 KB: one (direction, plane) operand tile
    static constexpr int A_STAGE = 2 * P * A_TILE;
    static constexpr int W_TILE = ND * CX_BK * 2;                 // one (direction, plane) chunk of W_x
    static constexpr int W_STAGE = 2 * P * W_TILE;
    static constexpr int BAR_BYTES = 128;
    static constexpr int SMEM_BYTES = 1024 + CX_ASTAGES * A_STAGE + CX_WSTAGES * W_STAGE + BAR_BYTES;
    static_assert(W_TILE % 1024 == 0, "W chunk must keep 1024-byte alignment for SWIZZLE_128B");
};

template <typename XT>
__device__ __forceinline__ float4 cx_load4(const XT* p);
template <>
__device__ __forceinline__ float4 cx_load4<float>(const float* p) {
    return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 cx_load4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 raw = *reinterpret_cast<const uint2*>(p);
    const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
    const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}

// four consecutive k of row r of a [128][64] bf16 tile in the SWIZZLE_128B K-major layout (what a TMA box of 64 x 128 writes):
// 8-row groups 1024 B apart, rows 128 B apart, the 16-byte chunk index XORed with the row's index inside its group
__device__ __forceinline__ uint32_t cx_swz(int r, int k) {
    return uint32_t((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + (k & 7) * 2);
}

template <int P>
__device__ __forceinline__ void cx_put(uint8_t* tile_hi, uint8_t* tile_lo, uint32_t off, float4 v) {
    __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
    if (P == 2) {
        split_bf16(v.x, h0, l0);
        split_bf16(v.y, h1, l1);
        split_bf16(v.z, h2, l2);
        split_bf16(v.w, h3, l3);
    } else {
        h0 = __float2bfloat16_rn(v.x);
        h1 = __float2bfloat16_rn(v.y);
        h2 = __float2bfloat16_rn(v.z);
        h3 = __float2bfloat16_rn(v.w);
    }
    __nv_bfloat162 a = __halves2bfloat162(h0, h1), b = __halves2bfloat162(h2, h3);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&a);
    pk.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(tile_hi + off) = pk;
    if (P == 2) {
        __nv_bfloat162 c = __halves2bfloat162(l0, l1), d = __halves2bfloat162(l2, l3);
        pk.x = *reinterpret_cast<uint32_t*>(&c);
        pk.y = *reinterpret_cast<uint32_t*>(&d);
        *reinterpret_cast<uint2*>(tile_lo + off) = pk;
    }
}

template <int P, int ND, typename XT>
__global__ void __launch_bounds__(CX_THREADS, 1)
conv_xproj_kernel(const __grid_constant__ CUtensorMap mapW, const ConvXParams p) {
    using Cfg = ConvXCfg<P, ND>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_base = smem;
    uint8_t* w_base = smem + CX_ASTAGES * Cfg::A_STAGE;
    uint64_t* afull = reinterpret_cast<uint64_t*>(w_base + CX_WSTAGES * Cfg::W_STAGE);
    uint64_t* aempty = afull + CX_ASTAGES;
    uint64_t* wfull = aempty + CX_ASTAGES;
    uint64_t* wempty = wfull + CX_WSTAGES;
    uint64_t* tfull = wempty + CX_WSTAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapW);
        for (int s = 0; s < CX_ASTAGES; ++s) {
            mbar_init(&afull[s], CX_CONV_THREADS);
            mbar_init(&aempty[s], 1);
        }
        for (int s = 0; s < CX_WSTAGES; ++s) {
            mbar_init(&wfull[s], 1);
            mbar_init(&wempty[s], 1);
        }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 2 * CX_ACC_STRIDE);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int kchunks = p.di / CX_BK;

    if (warp == 9) {
        // ------------------------------------------------------------ TMA producer of the W_x chunks (one thread)
        if (lane == 0) {
            int ws = 0;
            uint32_t wph = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&wempty[ws], wph ^ 1);
                    mbar_arrive_expect_tx(&wfull[ws], Cfg::W_STAGE);
                    uint8_t* wb = w_base + ws * Cfg::W_STAGE;
#pragma unroll
                    for (int d = 0; d < 2; ++d)
#pragma unroll
                        for (int pl = 0; pl < P; ++pl)
                            tma_load_3d(wb + (d * P + pl) * Cfg::W_TILE, &mapW, &wfull[ws], kc * CX_BK, d * ND, pl);
                    if (++ws == CX_WSTAGES) {
                        ws = 0;
                        wph ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 8) {
        // ------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(CX_BM, ND);
            int as = 0, ws = 0;
            uint32_t aph = 0, wph = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&wfull[ws], wph);
                    mbar_wait(&afull[as], aph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(a_base + as * Cfg::A_STAGE);
                    const uint32_t sw = smem_u32(w_base + ws * Cfg::W_STAGE);
#pragma unroll
                    for (int d = 0; d < 2; ++d) {
                        const uint32_t d_tmem = tmem_base + d * CX_ACC_STRIDE;
#pragma unroll
                        for (int kk = 0; kk < CX_BK / 16; ++kk) {
                            const uint64_t a_hi = make_smem_desc_sw128(sa + (d * P) * Cfg::A_TILE + kk * 32);
                            const uint64_t b_hi = make_smem_desc_sw128(sw + (d * P) * Cfg::W_TILE + kk * 32);
                            tc_mma_bf16(d_tmem, a_hi, b_hi, idesc, (kc | kk) != 0 ? 1u : 0u);
                            if (P == 2) {
                                const uint64_t a_lo = make_smem_desc_sw128(sa + (d * P + 1) * Cfg::A_TILE + kk * 32);
                                const uint64_t b_lo = make_smem_desc_sw128(sw + (d * P + 1) * Cfg::W_TILE + kk * 32);
                                tc_mma_bf16(d_tmem, a_lo, b_hi, idesc, 1u);
                                tc_mma_bf16(d_tmem, a_hi, b_lo, idesc, 1u);
                            }
                        }
                    }
                    tc_commit(&aempty[as]);   // operand slots reusable once these MMAs retire
                    tc_commit(&wempty[ws]);
                    if (++as == CX_ASTAGES) {
                        as = 0;
                        aph ^= 1;
                    }
                    if (++ws == CX_WSTAGES) {
                        ws = 0;
                        wph ^= 1;
                    }
                }
                tc_commit(tfull);   // both accumulators of this tile complete
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ conv producers (warps 0..7), then the tile's epilogue
        const int tid = threadIdx.x;
        const int cq = tid & 15;   // channel quad inside the 64-channel chunk
        const int rg = tid >> 4;   // 8-row group of the tile
        const size_t u_plane = p.u_plane;
        int as = 0;
        uint32_t aph = 0, tph = 0;
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int b = tile / p.tiles_per_seq;
            const int t0 = (tile - b * p.tiles_per_seq) * CX_BM;
            const XT* xrow = reinterpret_cast<const XT*>(p.xz) + size_t(b) * p.L * p.ldxz;
            for (int kc = 0; kc < kchunks; ++kc) {
                const int c = kc * CX_BK + 4 * cq;
                // tap k of the forward filter multiplies x[t - 3 + k] (causal), of the backward filter x[t + 3 - k] (the flipped sequence)
                float4 wf[4], wb[4];   // here indexed by CHANNEL: wf[i] = the four taps of channel c + i
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    wf[i] = __ldg(reinterpret_cast<const float4*>(p.conv_w + size_t(c + i) * 4));
                    wb[i] = __ldg(reinterpret_cast<const float4*>(p.conv_w + size_t(p.di + c + i) * 4));
                }
                const float4 bf = __ldg(reinterpret_cast<const float4*>(p.conv_b + c));
                const float4 bb = __ldg(reinterpret_cast<const float4*>(p.conv_b + p.di + c));
                const int tr = t0 + rg * 8;
                auto ld = [&](int t) -> float4 {
                    return (t >= 0 && t < p.L) ? cx_load4<XT>(xrow + size_t(t) * p.ldxz + c) : zero;
                };
                // two halves of four rows each: a 10-row window per half keeps the thread under the 168-register cap of a
                // 12-warp CTA (one 14-row window spilled ~170 B); the second half's window is re-read from L1
                float4 xw[10];   // x[tr + 4*half - 3 .. tr + 4*half + 6]
#pragma unroll
                for (int i = 0; i < 10; ++i) xw[i] = ld(tr - 3 + i);
                mbar_wait(&aempty[as], aph ^ 1);
                uint8_t* sa = a_base + as * Cfg::A_STAGE;
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    if (half == 1) {
#pragma unroll
                        for (int i = 0; i < 10; ++i) xw[i] = ld(tr + 1 + i);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = rg * 8 + half * 4 + i, t = tr + half * 4 + i;
                        const float4 w0 = xw[i], w1 = xw[i + 1], w2 = xw[i + 2], w3 = xw[i + 3], w4 = xw[i + 4], w5 = xw[i + 5],
                                     w6 = xw[i + 6];
                        float4 f, r4;
                        f.x = fmaf(wf[0].x, w0.x, fmaf(wf[0].y, w1.x, fmaf(wf[0].z, w2.x, fmaf(wf[0].w, w3.x, bf.x))));
                        f.y = fmaf(wf[1].x, w0.y, fmaf(wf[1].y, w1.y, fmaf(wf[1].z, w2.y, fmaf(wf[1].w, w3.y, bf.y))));
                        f.z = fmaf(wf[2].x, w0.z, fmaf(wf[2].y, w1.z, fmaf(wf[2].z, w2.z, fmaf(wf[2].w, w3.z, bf.z))));
                        f.w = fmaf(wf[3].x, w0.w, fmaf(wf[3].y, w1.w, fmaf(wf[3].z, w2.w, fmaf(wf[3].w, w3.w, bf.w))));
                        r4.x = fmaf(wb[0].x, w6.x, fmaf(wb[0].y, w5.x, fmaf(wb[0].z, w4.x, fmaf(wb[0].w, w3.x, bb.x))));
                        r4.y = fmaf(wb[1].x, w6.y, fmaf(wb[1].y, w5.y, fmaf(wb[1].z, w4.y, fmaf(wb[1].w, w3.y, bb.y))));
                        r4.z = fmaf(wb[2].x, w6.z, fmaf(wb[2].y, w5.z, fmaf(wb[2].z, w4.z, fmaf(wb[2].w, w3.z, bb.z))));
                        r4.w = fmaf(wb[3].x, w6.w, fmaf(wb[3].y, w5.w, fmaf(wb[3].z, w4.w, fmaf(wb[3].w, w3.w, bb.w))));
                        constexpr bool FS = (P == 1);   // one bf16 plane: the result is rounded to 8 mantissa bits anyway
                        f = make_float4(silu_sel<FS>(f.x), silu_sel<FS>(f.y), silu_sel<FS>(f.z), silu_sel<FS>(f.w));
                        r4 = make_float4(silu_sel<FS>(r4.x), silu_sel<FS>(r4.y), silu_sel<FS>(r4.z), silu_sel<FS>(r4.w));
                        if (t >= p.L) {   // rows past the end of the utterance: finite operands, nothing stored
                            f = zero;
                            r4 = zero;
                        } else {
                            const size_t off = (size_t(b) * p.L + t) * (2 * size_t(p.di)) + c;
                            store_planes4<P>(p.u, u_plane, off, f);
                            store_planes4<P>(p.u, u_plane, off + p.di, r4);
                        }
                        const uint32_t so = cx_swz(r, 4 * cq);
                        cx_put<P>(sa, sa + Cfg::A_TILE, so, f);
                        cx_put<P>(sa + P * Cfg::A_TILE, sa + (P + 1) * Cfg::A_TILE, so, r4);
                    }
                }
                fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async proxy
                mbar_arrive(&afull[as]);
                if (++as == CX_ASTAGES) {
                    as = 0;
                    aph ^= 1;
                }
            }
            // ---- epilogue: [dt | B | C] of both directions, thread = token row; warps 0..3 direction 0, 4..7 direction 1
            mbar_wait(tfull, tph);
            tph ^= 1;
            tc_fence_after();
            {
                const int q = warp & 3, d = warp >> 2;
                const int t = t0 + q * 32 + lane;
                float* drow = p.dbl + (size_t(b) * p.L + t) * p.ld_dbl + d * ND;
#pragma unroll
                for (int c0 = 0; c0 < ND; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld_x16(tmem_base + d * CX_ACC_STRIDE + c0 + (uint32_t(q * 32) << 16), v);
                    tmem_ld_wait();
                    if (t < p.L) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<float4*>(drow + c0 + 4 * j) =
                                make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                            __uint_as_float(v[4 * j + 3]));
                    }
                }
            }
            tc_fence_before();   // the next tile's first MMA overwrites the accumulators: ordered behind these reads by afull
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * CX_ACC_STRIDE);
    }
}

template <int P, int ND, typename XT>
static int launch_conv_xproj(const ConvXParams& p, const void* wx, cudaStream_t s) {
    using Cfg = ConvXCfg<P, ND>;
    CUtensorMap mapW;
    {
        uint64_t dims[3] = {(uint64_t)p.di, (uint64_t)2 * ND, (uint64_t)P};
        uint64_t str[2] = {(uint64_t)p.di * 2, (uint64_t)2 * ND * p.di * 2};
        uint32_t box[3] = {CX_BK, ND, 1};
        if (!encode_tmap(&mapW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, wx, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return MTN_ECUDA;
    }
    auto kern = conv_xproj_kernel<P, ND, XT>;
    static std::atomic<unsigned long long> attr_done{0};
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES, attr_done, "conv_xproj")) return rc;
    int grid = num_sms();
    if (grid > p.total_tiles) grid = p.total_tiles;
    kern<<<grid, CX_THREADS, Cfg::SMEM_BYTES, s>>>(mapW, p);
    MTN_CUDA_LAUNCH_CHECK("conv_xproj");
    return MTN_OK;
}

}  // namespace mtn

using namespace mtn;

extern "C" int mtn_conv_xproj_fwd(const void* xz, int ldxz, int xz_bf16, const float* conv_w, const float* conv_b, void* u_planes,
                                  int u_rows, const void* wx_planes, float* dbl, int ld_dbl, int n_dbl, int batch, int L, int di,
                                  int planes, mtn_stream_t stream) {
    MTN_REQUIRE(xz && conv_w && conv_b && u_planes && wx_planes && dbl, "conv_xproj: null pointer");
    MTN_REQUIRE(batch > 0 && L > 0 && di > 0 && di % CX_BK == 0, "conv_xproj: bad shape batch=%d L=%d di=%d (di %% 64 == 0)", batch, L, di);
    MTN_REQUIRE(planes == 1 || planes == 2, "conv_xproj: planes=%d", planes);
    MTN_REQUIRE(n_dbl == 48 || n_dbl == 64, "conv_xproj: n_dbl=%d (48 or 64)", n_dbl);
    MTN_REQUIRE(ld_dbl >= 2 * n_dbl && ld_dbl % 4 == 0 && (reinterpret_cast<uintptr_t>(dbl) & 15) == 0, "conv_xproj: dbl layout");
    MTN_REQUIRE(ldxz % (xz_bf16 ? 4 : 4) == 0 && (reinterpret_cast<uintptr_t>(xz) & 15) == 0, "conv_xproj: xz alignment");
    MTN_REQUIRE(u_rows >= batch * L, "conv_xproj: u_rows=%d < batch*L", u_rows);
    MTN_REQUIRE((long long)batch * ((L + CX_BM - 1) / CX_BM) < (1ll << 31), "conv_xproj: too many tiles");
    ConvXParams p;
    p.xz = xz;
    p.conv_w = conv_w;
    p.conv_b = conv_b;
    p.u = reinterpret_cast<__nv_bfloat16*>(u_planes);
    p.dbl = dbl;
    p.u_plane = size_t(u_rows) * 2 * di;
    p.ldxz = ldxz;
    p.ld_dbl = ld_dbl;
    p.nd = n_dbl;
    p.batch = batch;
    p.L = L;
    p.di = di;
    p.tiles_per_seq = (L + CX_BM - 1) / CX_BM;
    p.total_tiles = batch * p.tiles_per_seq;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
#define MTN_CX(PP, NDD)                                                                                   \
    do {                                                                                                  \
        if (xz_bf16) return launch_conv_xproj<PP, NDD, __nv_bfloat16>(p, wx_planes, s);                   \
        return launch_conv_xproj<PP, NDD, float>(p, wx_planes, s);                                        \
    } while (0)
    if (planes == 2) {
        if (n_dbl == 48) MTN_CX(2, 48);
        MTN_CX(2, 64);
    }
    if (n_dbl == 48) MTN_CX(1, 48);
    MTN_CX(1, 64);
#undef MTN_CX
}
