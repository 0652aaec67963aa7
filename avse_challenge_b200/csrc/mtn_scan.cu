// Selective-scan recurrence of the bidirectional Mamba mixer (sm_100a), both time directions in one launch.
//
// Replaces, per direction: the dt_proj GEMM (Mamba-TasNet/modules/mamba/selective_scan_interface.py:187) and
// selective_scan_cuda.fwd (ssi.py:218-220; semantics = selective_scan_ref, ssi.py:91-157):
//     delta = softplus(dbl[:, :R] . W_dt[d, :] + dt_bias[d])          (ssi.py:187, :110-112)
//     h[d, n] = exp(delta * A[d, n]) * h[d, n] + delta * B_t[n] * u_t[d]   (ssi.py:126-139)
//     y_t[d] = sum_n h[d, n] * C_t[n] + D[d] * u_t[d]                 (ssi.py:144, :153)
//     out    = 0.5 * y * silu(z)                                      (ssi.py:155; 0.5 = bimamba.py:253)
// The backward direction walks t = L-1..0 on the same buffers (the reference flips xz instead, bimamba.py:237).
//
// Mapping: one thread owns one channel d of one utterance and keeps its 16 SSM states in registers (packed
// f32x2 FMAs); a CTA = 128 channels + 1 producer warp.  The producer streams (u, silu(z), [dt|B|C]) time tiles
// through a 4-stage TMA/mbarrier ring, so the consumers never touch global memory for inputs; B_t/C_t/dt_t are
// warp-broadcast shared-memory reads.  delta is never materialised in HBM (dt_proj is R FMAs from registers).
// The kernel is MUFU-bound before it is HBM-bound (16 ex2 per (t, d)); see DESIGN.md for both rooflines.
#include "mtn_ptx.cuh"
#include "mtn_host.h"

namespace mtn {

constexpr int SC_CH = 128;
constexpr int SC_TT = 16;
constexpr int SC_STAGES = 4;
constexpr int SC_NS = 16;

struct ScanParams {
    const float* w_dt;
    const float* dt_bias;
    const float* A2;
    const float* Dskip;
    __nv_bfloat16* y;
    const float* h_in;
    float* h_out;
    int batch, L, di, n_dbl, z_col0;
    int dir0;  // first direction handled by blockIdx.z == 0
};

template <int P, int NDBL, typename ZT>
struct ScanSmem {
    static constexpr int U_BYTES = P * SC_TT * SC_CH * 2;
    static constexpr int Z_BYTES = SC_TT * SC_CH * int(sizeof(ZT));
    static constexpr int D_BYTES = SC_TT * NDBL * 4;
    static constexpr int STAGE_BYTES = U_BYTES + Z_BYTES + D_BYTES;
    static constexpr int Y_BYTES = (SC_CH / 32) * P * SC_TT * 32 * 2;  // per-warp output staging
    static constexpr int TOTAL = 128 + SC_STAGES * STAGE_BYTES + 2 * SC_STAGES * 8 + Y_BYTES;
};

__device__ __forceinline__ float ldz(const float* p) { return *p; }
__device__ __forceinline__ float ldz(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// softplus with ONE MUFU op: max(x,0) + log1p(exp(-|x|)); log1p(e) = e * Q(e) on [0,1], Q = degree-8 Chebyshev fit of
// log1p(e)/e (max rel. err 9e-8, so tiny deltas keep full relative accuracy).  Equals torch's softplus incl. its
// "linear above 20" branch to fp32 rounding (exp(-20) ~ 2e-9 vanishes against x).
__device__ __forceinline__ float softplus_1mufu(float x) {
    const float e = ex2_approx(-1.4426950408889634f * fabsf(x));
    float q = 0.0051261021414032125f;
    q = fmaf(q, e, -0.02907406467853027f);
    q = fmaf(q, e, 0.07751608674076167f);
    q = fmaf(q, e, -0.13602247622393474f);
    q = fmaf(q, e, 0.19076880735651539f);
    q = fmaf(q, e, -0.24835398988480129f);
    q = fmaf(q, e, 0.3331812170752912f);
    q = fmaf(q, e, -0.49999444976340335f);
    q = fmaf(q, e, 0.9999999659255092f);
    return fmaf(q, e, fmaxf(x, 0.f));
}

// One direction of one CTA's channels.  REV is a template parameter so that, with the 16-step tile fully unrolled,
// every per-step register array index is a compile-time constant.
template <int P, int R, int NDBL, typename ZT, bool REV>
__device__ __forceinline__ void scan_consumer(uint8_t* smem, uint64_t* full_bar, uint64_t* empty_bar,
                                              __nv_bfloat16* sy, const ScanParams& p, int ch0, int b, int dir) {
    using SM = ScanSmem<P, NDBL, ZT>;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int L = p.L;
    const int ntiles = (L + SC_TT - 1) / SC_TT;
    const int d = ch0 + tid;
    const size_t pd = size_t(dir) * p.di + d;
    float2 h2[SC_NS / 2], A2[SC_NS / 2];
    float wdt[R];
    {
        const float4* ap = reinterpret_cast<const float4*>(p.A2 + pd * SC_NS);
#pragma unroll
        for (int q = 0; q < SC_NS / 4; ++q) {
            const float4 a = ap[q];
            A2[2 * q] = make_float2(a.x, a.y);
            A2[2 * q + 1] = make_float2(a.z, a.w);
        }
        const float4* wp = reinterpret_cast<const float4*>(p.w_dt + pd * R);
#pragma unroll
        for (int q = 0; q < R / 4; ++q) {
            const float4 w = wp[q];
            wdt[4 * q] = w.x;
            wdt[4 * q + 1] = w.y;
            wdt[4 * q + 2] = w.z;
            wdt[4 * q + 3] = w.w;
        }
        if (p.h_in) {
            const float4* hp = reinterpret_cast<const float4*>(p.h_in + ((size_t(dir) * p.batch + b) * p.di + d) * SC_NS);
#pragma unroll
            for (int q = 0; q < SC_NS / 4; ++q) {
                const float4 a = hp[q];
                h2[2 * q] = make_float2(a.x, a.y);
                h2[2 * q + 1] = make_float2(a.z, a.w);
            }
        } else {
#pragma unroll
            for (int q = 0; q < SC_NS / 2; ++q) h2[q] = make_float2(0.f, 0.f);
        }
    }
    const float bias = p.dt_bias[pd];
    const float Dv = p.Dskip[pd];
    const size_t M = size_t(p.batch) * L;
    const size_t y_plane = M * 2 * p.di;
    // this warp's 32-channel column block of y; rows are 2*di bf16 apart
    __nv_bfloat16* ywarp = p.y + size_t(dir) * p.di + ch0 + warp * 32;
    __nv_bfloat16* sy_w = sy + warp * (P * SC_TT * 32);  // warp-private staging [P][TT][32]

    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < ntiles; ++i) {
        const int tile = REV ? (ntiles - 1 - i) : i;
        const int t0 = tile * SC_TT;
        const int nvalid = min(SC_TT, L - t0);
        mbar_wait(&full_bar[stage], phase);
        const uint8_t* st = smem + stage * SM::STAGE_BYTES;
        const __nv_bfloat16* su = reinterpret_cast<const __nv_bfloat16*>(st);
        const ZT* sz = reinterpret_cast<const ZT*>(st + SM::U_BYTES);
        const float* sd = reinterpret_cast<const float*>(st + SM::U_BYTES + SM::Z_BYTES);

        // ---- phase A: per-step scalars for the whole tile (16 independent chains -> ILP)
        float delta[SC_TT], du[SC_TT];
#pragma unroll
        for (int j = 0; j < SC_TT; ++j) {
            const float* drow = sd + j * NDBL;
            float acc0 = bias, acc1 = 0.f;
#pragma unroll
            for (int q = 0; q < R / 4; ++q) {
                const float4 x = *reinterpret_cast<const float4*>(drow + 4 * q);
                acc0 = fmaf(x.x, wdt[4 * q], acc0);
                acc1 = fmaf(x.y, wdt[4 * q + 1], acc1);
                acc0 = fmaf(x.z, wdt[4 * q + 2], acc0);
                acc1 = fmaf(x.w, wdt[4 * q + 3], acc1);
            }
            float dl = softplus_1mufu(acc0 + acc1);
            dl = (j < nvalid) ? dl : 0.f;  // rows past the utterance end: exp2(0)=1, dBu=0 -> state unchanged
            float uval = __bfloat162float(su[j * SC_CH + tid]);
            if (P == 2) uval += __bfloat162float(su[SC_TT * SC_CH + j * SC_CH + tid]);
            delta[j] = dl;
            du[j] = dl * uval;
        }
        // ---- phase B: the recurrence, steps in processing order; exps of later steps are independent of h
#pragma unroll
        for (int jj = 0; jj < SC_TT; ++jj) {
            const int j = REV ? (SC_TT - 1 - jj) : jj;
            const float* drow = sd + j * NDBL;
            const float2 delta2 = make_float2(delta[j], delta[j]);
            const float2 du2 = make_float2(du[j], du[j]);
            float2 ya = make_float2(0.f, 0.f), yb = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < SC_NS / 4; ++q) {
                const float4 Bq = *reinterpret_cast<const float4*>(drow + R + 4 * q);
                const float4 Cq = *reinterpret_cast<const float4*>(drow + R + SC_NS + 4 * q);
                {
                    const float2 a = __fmul2_rn(delta2, A2[2 * q]);
                    const float2 e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
                    const float2 bu = __fmul2_rn(du2, make_float2(Bq.x, Bq.y));
                    h2[2 * q] = __ffma2_rn(e, h2[2 * q], bu);
                    ya = __ffma2_rn(h2[2 * q], make_float2(Cq.x, Cq.y), ya);
                }
                {
                    const float2 a = __fmul2_rn(delta2, A2[2 * q + 1]);
                    const float2 e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
                    const float2 bu = __fmul2_rn(du2, make_float2(Bq.z, Bq.w));
                    h2[2 * q + 1] = __ffma2_rn(e, h2[2 * q + 1], bu);
                    yb = __ffma2_rn(h2[2 * q + 1], make_float2(Cq.z, Cq.w), yb);
                }
            }
            float uval = __bfloat162float(su[j * SC_CH + tid]);
            if (P == 2) uval += __bfloat162float(su[SC_TT * SC_CH + j * SC_CH + tid]);
            const float zval = ldz(sz + j * SC_CH + tid);
            const float y = ((ya.x + ya.y) + (yb.x + yb.y) + Dv * uval) * (0.5f * zval);
            if (P == 2) {
                __nv_bfloat16 hi, lo;
                split_bf16(y, hi, lo);
                sy_w[j * 32 + lane] = hi;
                sy_w[SC_TT * 32 + j * 32 + lane] = lo;
            } else {
                sy_w[j * 32 + lane] = __float2bfloat16_rn(y);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);  // inputs of this stage are consumed
        // ---- flush this warp's staged y rows: 4 lanes x 16 B cover one 64-byte row segment
#pragma unroll
        for (int pl = 0; pl < P; ++pl) {
#pragma unroll
            for (int it = 0; it < SC_TT / 8; ++it) {
                const int row = it * 8 + (lane >> 2);
                const int seg = lane & 3;
                if (row < nvalid) {
                    const uint4 v = *reinterpret_cast<const uint4*>(sy_w + pl * SC_TT * 32 + row * 32 + seg * 8);
                    const size_t off = (size_t(b) * L + t0 + row) * (2 * size_t(p.di)) + seg * 8;
                    *reinterpret_cast<uint4*>(ywarp + pl * y_plane + off) = v;
                }
            }
        }
        __syncwarp();
        if (++stage == SC_STAGES) {
            stage = 0;
            phase ^= 1;
        }
    }
    if (p.h_out) {
        float4* hp = reinterpret_cast<float4*>(p.h_out + ((size_t(dir) * p.batch + b) * p.di + d) * SC_NS);
#pragma unroll
        for (int q = 0; q < SC_NS / 4; ++q)
            hp[q] = make_float4(h2[2 * q].x, h2[2 * q].y, h2[2 * q + 1].x, h2[2 * q + 1].y);
    }
}

template <int P, int R, int NDBL, typename ZT>
__global__ void __launch_bounds__(SC_CH + 32)
scan_kernel(const __grid_constant__ CUtensorMap mapU, const __grid_constant__ CUtensorMap mapZ,
            const __grid_constant__ CUtensorMap mapD, const ScanParams p) {
    using SM = ScanSmem<P, NDBL, ZT>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // align inside the shared window without leaving the shared address space (keeps LDS/STS, not generic LD/ST)
    uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SC_STAGES * SM::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + SC_STAGES;
    __nv_bfloat16* sy = reinterpret_cast<__nv_bfloat16*>(smem + SC_STAGES * SM::STAGE_BYTES + 2 * SC_STAGES * 8);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int ch0 = blockIdx.x * SC_CH;
    const int b = blockIdx.y;
    const int dir = p.dir0 + blockIdx.z;
    const bool rev = dir == 1;
    const int L = p.L;
    const int ntiles = (L + SC_TT - 1) / SC_TT;

    if (tid == 0) {
        tma_prefetch_desc(&mapU);
        tma_prefetch_desc(&mapZ);
        tma_prefetch_desc(&mapD);
        for (int s = 0; s < SC_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], SC_CH / 32);
        }
        fence_barrier_init();
    }
    __syncthreads();

    if (warp == SC_CH / 32) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < ntiles; ++i) {
                const int tile = rev ? (ntiles - 1 - i) : i;
                const int row0 = b * L + tile * SC_TT;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full_bar[stage], SM::STAGE_BYTES);
                uint8_t* st = smem + stage * SM::STAGE_BYTES;
                tma_load_3d(st, &mapU, &full_bar[stage], dir * p.di + ch0, row0, 0);
                tma_load_2d(st + SM::U_BYTES, &mapZ, &full_bar[stage], p.z_col0 + ch0, row0);
                tma_load_2d(st + SM::U_BYTES + SM::Z_BYTES, &mapD, &full_bar[stage], dir * p.n_dbl, row0);
                if (++stage == SC_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        return;
    }
    if (rev)
        scan_consumer<P, R, NDBL, ZT, true>(smem, full_bar, empty_bar, sy, p, ch0, b, dir);
    else
        scan_consumer<P, R, NDBL, ZT, false>(smem, full_bar, empty_bar, sy, p, ch0, b, dir);
}

template <int P, int R, int NDBL, typename ZT>
static int launch_scan(const mtn_scan_args* a, cudaStream_t stream) {
    using SM = ScanSmem<P, NDBL, ZT>;
    const uint64_t M = uint64_t(a->batch) * a->L;
    CUtensorMap mapU, mapZ, mapD;
    {
        uint64_t dims[3] = {uint64_t(2) * a->di, M, uint64_t(P)};
        uint64_t str[2] = {uint64_t(2) * a->di * 2, M * 2 * a->di * 2};
        uint32_t box[3] = {SC_CH, SC_TT, uint32_t(P)};
        if (!encode_tmap(&mapU, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a->u, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE))
            return MTN_ECUDA;
    }
    {
        uint64_t dims[2] = {uint64_t(a->ldz), M};
        uint64_t str[1] = {uint64_t(a->ldz) * sizeof(ZT)};
        uint32_t box[2] = {SC_CH, SC_TT};
        const CUtensorMapDataType dt =
            sizeof(ZT) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
        if (!encode_tmap(&mapZ, dt, 2, a->z, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return MTN_ECUDA;
    }
    {
        uint64_t dims[2] = {uint64_t(a->ld_dbl), M};
        uint64_t str[1] = {uint64_t(a->ld_dbl) * 4};
        uint32_t box[2] = {uint32_t(NDBL), SC_TT};
        if (!encode_tmap(&mapD, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a->dbl, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE))
            return MTN_ECUDA;
    }
    ScanParams p;
    p.w_dt = a->w_dt;
    p.dt_bias = a->dt_bias;
    p.A2 = a->A2;
    p.Dskip = a->Dskip;
    p.y = reinterpret_cast<__nv_bfloat16*>(a->y);
    p.h_in = a->h_in;
    p.h_out = a->h_out;
    p.batch = a->batch;
    p.L = a->L;
    p.di = a->di;
    p.n_dbl = a->n_dbl;
    p.z_col0 = a->z_col0;
    p.dir0 = (a->dir_mask & 1) ? 0 : 1;
    const int ndirs = (a->dir_mask == 3) ? 2 : 1;
    auto kern = scan_kernel<P, R, NDBL, ZT>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
        if (e != cudaSuccess) {
            set_error("scan: cudaFuncSetAttribute(%d B smem) failed: %s", SM::TOTAL, cudaGetErrorString(e));
            return MTN_ECUDA;
        }
        attr_set = true;
    }
    dim3 grid(a->di / SC_CH, a->batch, ndirs);
    kern<<<grid, SC_CH + 32, SM::TOTAL, stream>>>(mapU, mapZ, mapD, p);
    MTN_CUDA_LAUNCH_CHECK("scan");
    return MTN_OK;
}

template <int P, typename ZT>
static int dispatch_scan_r(const mtn_scan_args* a, cudaStream_t s) {
    if (a->R == 4 && a->n_dbl == 48) return launch_scan<P, 4, 48, ZT>(a, s);
    if (a->R == 8 && a->n_dbl == 48) return launch_scan<P, 8, 48, ZT>(a, s);
    if (a->R == 16 && a->n_dbl == 48) return launch_scan<P, 16, 48, ZT>(a, s);
    if (a->R == 32 && a->n_dbl == 64) return launch_scan<P, 32, 64, ZT>(a, s);
    set_error("scan: unsupported dt_rank R=%d / n_dbl=%d (supported: 4|8|16 with 48, 32 with 64)", a->R, a->n_dbl);
    return MTN_EINVAL;
}

}  // namespace mtn

extern "C" int mtn_scan_fwd(const mtn_scan_args* a, mtn_stream_t stream) {
    using namespace mtn;
    MTN_REQUIRE(a && a->u && a->dbl && a->z && a->w_dt && a->dt_bias && a->A2 && a->Dskip && a->y, "scan: null pointer");
    MTN_REQUIRE(a->batch > 0 && a->batch <= 65535 && a->L > 0, "scan: bad batch=%d L=%d", a->batch, a->L);
    MTN_REQUIRE(a->di > 0 && a->di % SC_CH == 0, "scan: di=%d must be a multiple of %d", a->di, SC_CH);
    MTN_REQUIRE(a->dir_mask >= 1 && a->dir_mask <= 3, "scan: dir_mask=%d", a->dir_mask);
    MTN_REQUIRE(a->ld_dbl % 4 == 0 && a->ld_dbl >= 2 * a->n_dbl, "scan: ld_dbl=%d too small for 2 x n_dbl=%d", a->ld_dbl,
                a->n_dbl);
    MTN_REQUIRE(a->ldz % 8 == 0 && a->z_col0 + a->di <= a->ldz, "scan: z view out of range");
    MTN_REQUIRE(uint64_t(a->batch) * a->L < (1ull << 31), "scan: too many tokens for 32-bit TMA coordinates");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (a->planes == 2 && !a->z_bf16) return dispatch_scan_r<2, float>(a, s);
    if (a->planes == 1 && a->z_bf16) return dispatch_scan_r<1, __nv_bfloat16>(a, s);
    if (a->planes == 1 && !a->z_bf16) return dispatch_scan_r<1, float>(a, s);
    set_error("scan: unsupported planes=%d z_bf16=%d", a->planes, a->z_bf16);
    return MTN_EINVAL;
}
