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
// Mapping (default, "split"): TWO lanes (l, l+16 of one warp) own one channel d of one utterance, 8 of its 16 SSM
// states each, in registers (packed f32x2 FMAs); a CTA = 128 channels = 8 consumer warps + 1 producer warp.  The
// time recurrence offers only batch*d_inner*2 independent sequences (32 768 at BASELINE config 2, i.e. 2 warps per
// SM sub-partition if one thread owned a whole channel); splitting the state vector doubles the resident warps so
// the MUFU / FMA / LSU latencies overlap across warps, at the price of three shuffles per step (delta, delta*u and
// the partial y).  Per-step scalar work (dt_proj, softplus, gating, stores) is shared between the two lanes by
// time parity: lane half h prepares and finalises the rows t = 2k + h of each 16-row tile.
// The producer streams (u, silu(z), [dt|B|C]) time tiles through a TMA/mbarrier ring, so consumers never touch
// global memory for inputs; B_t/C_t/dt_t are warp-broadcast shared-memory reads.  delta is never materialised in
// HBM (dt_proj is R FMAs from registers).  The kernel is MUFU-bound before it is HBM-bound (16 ex2 per (t, d));
// KP > 0 moves KP of every 4 state pairs of a lane to an FMA-pipe polynomial exp2 (Cody-Waite + degree-5 minimax,
// 2e-7 relative) to balance the MUFU and FMA pipes.  See DESIGN.md for both rooflines.
// MTN_SCAN_VARIANT (debug knob, read per call): 0/unset = tensor-core dt_proj when args.dtp is given, FMA dt_proj
// otherwise; 2 = FMA dt_proj even with dtp.  (Builds with MTN_SCAN_ABLATIONS=1 add timing-only experiment variants.)
#include "mtn_ptx.cuh"
#include "mtn_host.h"
#include <stdlib.h>

namespace mtn {

constexpr int SC_CH = 128;
constexpr int SC_TT = 16;
constexpr int SC_NS = 16;

struct ScanParams {
    const float* w_dt;
    const float* dt_bias;
    const float* A2;
    const float* Dskip;
    __nv_bfloat16* y;
    const float* h_in;
    float* h_out;
    float* sum_delta;
    int batch, L, di, n_dbl, z_col0;
    int L_last;  // valid length of the last sequence of the batch (== L when not ragged)
    int dir0;   // first direction of this launch
    int ndirs;  // 1 or 2
    const void* z;  // gate column block base (pair mapping reads it straight from global)
    int ldz;
};

static ScanParams make_scan_params(const mtn_scan_args* a) {
    ScanParams p;
    p.w_dt = a->w_dt;
    p.dt_bias = a->dt_bias;
    p.A2 = a->A2;
    p.Dskip = a->Dskip;
    p.y = reinterpret_cast<__nv_bfloat16*>(a->y);
    p.h_in = a->h_in;
    p.h_out = a->h_out;
    p.sum_delta = a->sum_delta;
    p.L_last = a->L_last > 0 ? a->L_last : a->L;
    p.batch = a->batch;
    p.L = a->L;
    p.di = a->di;
    p.n_dbl = a->n_dbl;
    p.z_col0 = a->z_col0;
    p.dir0 = (a->dir_mask & 1) ? 0 : 1;
    p.ndirs = (a->dir_mask == 3) ? 2 : 1;
    p.z = a->z;
    p.ldz = a->ldz;
    return p;
}

template <int P, int NDBL, typename ZT, int CH>
struct ScanSmem {
    // CH = channels per CTA: 128 (8 warps) for large grids, 32 (2 warps) when the problem offers too few 128-channel
    // CTAs to load every SM evenly (BASELINE config 2: 256 CTAs on 148 SMs -> 1024 CTAs, 6.9 per SM)
    static constexpr int STAGES = CH == 128 ? 5 : 4;
    static constexpr int U_BYTES = P * SC_TT * CH * 2;
    static constexpr int Z_BYTES = SC_TT * CH * int(sizeof(ZT));
    static constexpr int D_BYTES = SC_TT * NDBL * 4;
    static constexpr int STAGE_BYTES = U_BYTES + Z_BYTES + D_BYTES;
    static constexpr int Y_BYTES = P * SC_TT * CH * 2;  // output staging, one column block per consumer warp
    static constexpr int TOTAL = 128 + STAGES * STAGE_BYTES + 2 * STAGES * 8 + Y_BYTES;
};

__device__ __forceinline__ float ldz(const float* p) { return *p; }
__device__ __forceinline__ float ldz(const __nv_bfloat16* p) {
    return __uint_as_float(uint32_t(*reinterpret_cast<const uint16_t*>(p)) << 16);
}
__device__ __forceinline__ float bf16_bits_to_float(const __nv_bfloat16* p) {
    return __uint_as_float(uint32_t(*reinterpret_cast<const uint16_t*>(p)) << 16);
}
// fp32 -> bf16 through the packed ALU-pipe convert (F2FP); the scalar F2F form runs on the XU pipe, which is the
// bottleneck of this kernel.
__device__ __forceinline__ uint16_t f2bf_bits(float x) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(x, 0.f);
    return uint16_t(*reinterpret_cast<const uint32_t*>(&v) & 0xFFFFu);
}

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

// 2^(dl * a) for a pair of states on the FMA pipe: n = round(dl*a) via the 1.5*2^23 trick, f = dl*a - n in
// [-0.5, 0.5] (single-rounded through the FMA), degree-5 polynomial with p(0) = 1 exactly (max rel. err 2e-7 in
// fp32 Horner), exponent added with integer arithmetic.  dl*a must be >= -126 (caller clamps dl per channel).
__device__ __forceinline__ float2 ex2_poly2(float2 dl2, float2 a2) {
    const float2 t = __ffma2_rn(dl2, a2, make_float2(12582912.f, 12582912.f));
    const float2 r = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
    const float2 f = __ffma2_rn(dl2, a2, make_float2(-r.x, -r.y));
    float2 q = make_float2(MTN_EX2_C5, MTN_EX2_C5);
    q = __ffma2_rn(q, f, make_float2(MTN_EX2_C4, MTN_EX2_C4));
    q = __ffma2_rn(q, f, make_float2(MTN_EX2_C3, MTN_EX2_C3));
    q = __ffma2_rn(q, f, make_float2(MTN_EX2_C2, MTN_EX2_C2));
    q = __ffma2_rn(q, f, make_float2(MTN_EX2_C1, MTN_EX2_C1));
    q = __ffma2_rn(q, f, make_float2(1.f, 1.f));
    return make_float2(__uint_as_float(__float_as_uint(q.x) + (__float_as_uint(t.x) << 23)),
                       __uint_as_float(__float_as_uint(q.y) + (__float_as_uint(t.y) << 23)));
}

// ---------------------------------------------------------------------------------------------------------------
// Split mapping: 256 consumer threads per CTA, lane = (half, cl): channel = warp*16 + cl, states [8*half, 8*half+8).
//
// Instruction mix is kept uniform over time: the per-row scalar work ("prep": dt_proj, softplus, delta*u) for tile
// i+1 is interleaved into the recurrence steps of tile i (one row every two steps, written over the register slot
// whose row was just finalised), so every warp presents the same MUFU : FMA : LSU ratio at all times and the four
// warps of an SM sub-partition never queue up on the XU pipe together.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sts_b16(void* p, uint32_t v) {
    // no "memory" clobber on purpose: it would pin every shared-memory load of the recurrence behind this store;
    // the staging buffer is only read after the __syncwarp() that follows the tile.
    asm volatile("{\n\t.reg .b16 t;\n\tcvt.u16.u32 t, %1;\n\tst.shared.b16 [%0], t;\n\t}" ::"r"(smem_u32(p)), "r"(v));
}
// bf16(x) in the low 16 bits (F2FP on the ALU pipe; the scalar F2F form would use the XU pipe)
__device__ __forceinline__ uint32_t f2bf_lo(float x) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(0.f), "f"(x));
    return r;
}

template <int P, int R, int NDBL, typename ZT, int CH, bool REV, int KP, bool WY>
__device__ __forceinline__ void scan_consumer_split(uint8_t* smem, uint64_t* full_bar, uint64_t* empty_bar,
                                                    __nv_bfloat16* sy, const ScanParams& p, int ch0, int b, int dir,
                                                    const CUtensorMap* mapU, const CUtensorMap* mapZ,
                                                    const CUtensorMap* mapD) {
    using SM = ScanSmem<P, NDBL, ZT, CH>;
    constexpr int SC_STAGES = SM::STAGES;
    constexpr int NW = CH / 16;  // consumer warps
    constexpr uint32_t FULL = 0xffffffffu;
    // KP >= 100: timing-only ablations (WRONG results; MTN_SCAN_VARIANT=101..107): bit0 drops the dt_proj FMAs, bit1
    // the B/C shared-memory loads, bit2 the MUFU ex2.  They measure the marginal cost of each instruction group.
    constexpr int ABL = KP >= 100 ? KP - 100 : 0;
    constexpr int KPE = KP >= 100 ? 0 : KP;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int half = lane >> 4, cl = lane & 15;
    const int chl = warp * 16 + cl;
    const int L = p.L;                                       // row stride between sequences
    const int Lb = (b == p.batch - 1) ? p.L_last : p.L;      // valid steps of this sequence
    const int ntiles = (Lb + SC_TT - 1) / SC_TT;
    const int d = ch0 + chl;
    const size_t pd = size_t(dir) * p.di + d;
    float sdl = 0.f;  // sum of this lane's deltas (rows 2k + half)
    float2 h2[4], A2[4];
    float2 wdt2[R / 2];
    {
        const float4* ap = reinterpret_cast<const float4*>(p.A2 + pd * SC_NS + half * 8);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float4 a = ap[q];
            A2[2 * q] = make_float2(a.x, a.y);
            A2[2 * q + 1] = make_float2(a.z, a.w);
        }
        const float4* wp = reinterpret_cast<const float4*>(p.w_dt + pd * R);
#pragma unroll
        for (int q = 0; q < R / 4; ++q) {
            const float4 w = wp[q];
            wdt2[2 * q] = make_float2(w.x, w.y);
            wdt2[2 * q + 1] = make_float2(w.z, w.w);
        }
        if (p.h_in) {
            const float4* hp =
                reinterpret_cast<const float4*>(p.h_in + ((size_t(dir) * p.batch + b) * p.di + d) * SC_NS + half * 8);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float4 a = hp[q];
                h2[2 * q] = make_float2(a.x, a.y);
                h2[2 * q + 1] = make_float2(a.z, a.w);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) h2[q] = make_float2(0.f, 0.f);
        }
    }
    const float bias = p.dt_bias[pd];
    const float Dv = p.Dskip[pd];
    // polynomial exp2 needs dl * A2 >= -126 for its KP pairs: clamp delta for those pairs only
    float dl_lim = 3.0e38f;
    if (KPE > 0) {
        float amin = -1e-30f;
#pragma unroll
        for (int q = 0; q < KPE; ++q) amin = fminf(amin, fminf(A2[q].x, A2[q].y));
        dl_lim = -125.f / amin;
    }
    const size_t M = size_t(p.batch) * L;
    const size_t y_plane = M * 2 * p.di;
    // this warp's 16-channel column block of y; rows are 2*di bf16 apart
    __nv_bfloat16* ywarp = p.y + size_t(dir) * p.di + ch0 + warp * 16;
    uint16_t* sy_w = reinterpret_cast<uint16_t*>(sy) + warp * (P * SC_TT * 16);  // warp-private staging [P][TT][16]
    const int src0 = cl, src1 = cl | 16;

    // per-row scalars of the tile in flight: slot k <-> row 2k + half
    float dmine[SC_TT / 2], dumine[SC_TT / 2], umine[SC_TT / 2];
    auto prep_row = [&](int k, const uint8_t* st, int nvalid_t) {
        const __nv_bfloat16* su_h = reinterpret_cast<const __nv_bfloat16*>(st) + half * CH + chl;
        const float* drow = reinterpret_cast<const float*>(st + SM::U_BYTES + SM::Z_BYTES) + (2 * k + half) * NDBL;
        float acc0 = bias, acc1 = 0.f;
        if (ABL & 1) {
            acc1 = drow[0] * wdt2[0].x;
        } else {
            // dt_proj: R MACs as R/2 packed FFMA2 on two independent accumulator pairs
            float2 pa = make_float2(bias, 0.f), pb = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < R / 4; ++q) {
                const float4 x = *reinterpret_cast<const float4*>(drow + 4 * q);
                pa = __ffma2_rn(make_float2(x.x, x.y), wdt2[2 * q], pa);
                pb = __ffma2_rn(make_float2(x.z, x.w), wdt2[2 * q + 1], pb);
            }
            const float2 pab = __fadd2_rn(pa, pb);
            acc0 = pab.x;
            acc1 = pab.y;
        }
        float dl = softplus_1mufu(acc0 + acc1);
        dl = (2 * k + half < nvalid_t) ? dl : 0.f;  // rows past the utterance end: exp2(0)=1, dBu=0 -> state unchanged
        float uval = bf16_bits_to_float(su_h + k * 2 * CH);
        if (P == 2) uval += bf16_bits_to_float(su_h + SC_TT * CH + k * 2 * CH);
        sdl += dl;
        dmine[k] = dl;
        dumine[k] = dl * uval;
        umine[k] = uval;
    };
    auto make_exps = [&](float2(&ex)[4], float dl) {
        const float2 dl2 = make_float2(dl, dl);
        const float dlc = fminf(dl, dl_lim);
        const float2 dlc2 = make_float2(dlc, dlc);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (q < KPE) {
                ex[q] = ex2_poly2(dlc2, A2[q]);
            } else {
                const float2 a = __fmul2_rn(dl2, A2[q]);
                if (ABL & 4) ex[q] = __ffma2_rn(a, make_float2(0.5f, 0.5f), make_float2(1.f, 1.f));
                else ex[q] = make_float2(ex2_approx(a.x), ex2_approx(a.y));
            }
        }
    };

    // No producer warp: lane 0 of warp 0 issues the TMA loads.  Tile i+3 (processing order) is requested at the start
    // of tile i into the stage tile i-2 used (5-stage ring), whose "empty" barrier has normally completed long before (no stall);
    // an even warp count per CTA also lifts the register cap from 96 to 128 at two CTAs per SM.
    auto issue_tile = [&](int i2, int stg) {
        const int tile2 = REV ? (ntiles - 1 - i2) : i2;
        const int row0 = b * L + tile2 * SC_TT;
        mbar_arrive_expect_tx(&full_bar[stg], SM::STAGE_BYTES);
        uint8_t* dst = smem + stg * SM::STAGE_BYTES;
        tma_load_3d(dst, mapU, &full_bar[stg], dir * p.di + ch0, row0, 0);
        tma_load_2d(dst + SM::U_BYTES, mapZ, &full_bar[stg], p.z_col0 + ch0, row0);
        tma_load_2d(dst + SM::U_BYTES + SM::Z_BYTES, mapD, &full_bar[stg], dir * p.n_dbl, row0);
    };
    if (tid == 0) {
#pragma unroll
        for (int t = 0; t < SC_STAGES - 2; ++t)
            if (t < ntiles) issue_tile(t, t);
    }
    int stage = 0;
    uint32_t phase = 0;
    // prologue: rows of the first tile
    {
        const int tile0 = REV ? (ntiles - 1) : 0;
        mbar_wait(&full_bar[0], 0);
        const int nv0 = min(SC_TT, Lb - tile0 * SC_TT);
#pragma unroll
        for (int k = 0; k < SC_TT / 2; ++k) prep_row(k, smem, nv0);
    }
    for (int i = 0; i < ntiles; ++i) {
        const int tile = REV ? (ntiles - 1 - i) : i;
        const int t0 = tile * SC_TT;
        const int nvalid = min(SC_TT, Lb - t0);
        const uint8_t* st = smem + stage * SM::STAGE_BYTES;
        const ZT* sz = reinterpret_cast<const ZT*>(st + SM::U_BYTES);
        const float* sd = reinterpret_cast<const float*>(st + SM::U_BYTES + SM::Z_BYTES);
        // the next tile's inputs are prepared while this tile is scanned
        const bool has_next = i + 1 < ntiles;
        const int stage_n = (stage + 1 == SC_STAGES) ? 0 : stage + 1;
        const uint8_t* st_n = smem + stage_n * SM::STAGE_BYTES;
        int nvalid_n = 0;  // no next tile: the (unconditional) row prep then yields delta = 0 everywhere
        if (warp == (i % NW) && i + (SC_STAGES - 2) < ntiles) {   // producer duty rotates over the warps (see scan_consumer_tc)
            // that stage was last used by tile i-2, whose "empty" phase (parity ((i-2)/STAGES)&1) is normally long
            // complete.  The WHOLE warp waits (converged): a lone lane sleeping in try_wait on this barrier while its
            // 31 siblings sleep on a different one (the "full" barrier below) produced millisecond stragglers.
            const int stg3 = (stage + SC_STAGES - 2) % SC_STAGES;
            if (i >= 2) mbar_wait(&empty_bar[stg3], uint32_t((i - 2) / SC_STAGES) & 1u);
            if (lane == 0) issue_tile(i + SC_STAGES - 2, stg3);
            __syncwarp();
        }
        if (has_next) {
            const int tile_n = REV ? (tile - 1) : (tile + 1);
            nvalid_n = min(SC_TT, Lb - tile_n * SC_TT);
            mbar_wait(&full_bar[stage_n], (stage + 1 == SC_STAGES) ? (phase ^ 1) : phase);
        }

        // Software pipeline over the 16 steps of the tile (processing order jj, row j):
        //   step jj issues   - the B/C shared-memory loads of step jj+1,
        //                    - the (delta, delta*u) shuffles of step jj+2,
        //                    - the 8 exps of step jj+1,
        //                    - the finalisation of the row pair completed at step jj-1 (its y shuffle is in flight),
        //   then updates the state with values that were all requested one or two steps earlier.
        const float* sdB = sd + R + half * 8;  // this lane's 8 B values; its C values are 16 floats further
        auto jrow = [](int jj) { return REV ? (SC_TT - 1 - jj) : jj; };
        auto shfl_row = [&](const float(&arr)[SC_TT / 2], int jj) {
            const int j = jrow(jj);
            return __shfl_sync(FULL, arr[j >> 1], (j & 1) ? src1 : src0);
        };
        float2 e[4];
        float4 Bc[2], Cc[2];
        float du_c, dl_n, du_n;
        {
            const float dl0 = shfl_row(dmine, 0);
            du_c = shfl_row(dumine, 0);
            dl_n = shfl_row(dmine, 1);
            du_n = shfl_row(dumine, 1);
            const float* brow = sdB + jrow(0) * NDBL;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (ABL & 2) {
                    Bc[q] = make_float4(du_c, dl_n, du_n, 0.5f);
                    Cc[q] = make_float4(dl_n, du_c, 0.25f, du_n);
                } else {
                    Bc[q] = *reinterpret_cast<const float4*>(brow + 4 * q);
                    Cc[q] = *reinterpret_cast<const float4*>(brow + SC_NS + 4 * q);
                }
            }
            make_exps(e, dl0);
        }
        auto finalize = [&](int k, float mine, float recv) {
            if (WY) {
                const int row = 2 * k + half;
                const float zval = ldz(sz + row * CH + chl);
                const float y = fmaf(Dv, umine[k], mine + recv) * (0.5f * zval);
                const uint32_t hi = f2bf_lo(y);
                sts_b16(sy_w + row * 16 + cl, hi);
                if (P == 2) sts_b16(sy_w + SC_TT * 16 + row * 16 + cl, f2bf_lo(y - __uint_as_float(hi << 16)));
            }
            // slot k is free now: fill it with the next tile's row.  Unconditional on purpose (keeps the 16 steps one
            // basic block): after the last tile it reads a stale stage and the values are never used.
            prep_row(k, st_n, nvalid_n);
        };
        float yprev = 0.f, mine_p = 0.f, recv_p = 0.f;
#pragma unroll
        for (int jj = 0; jj < SC_TT; ++jj) {
            const int j = jrow(jj);
            float4 Bn[2], Cn[2];
            float2 en[4];
            float dl_nn = 0.f, du_nn = 0.f;
            if (jj + 1 < SC_TT) {
                const float* brow = sdB + jrow(jj + 1) * NDBL;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (ABL & 2) {
                        Bn[q] = make_float4(du_c, dl_n, du_n, 0.5f);
                        Cn[q] = make_float4(dl_n, du_c, 0.25f, du_n);
                    } else {
                        Bn[q] = *reinterpret_cast<const float4*>(brow + 4 * q);
                        Cn[q] = *reinterpret_cast<const float4*>(brow + SC_NS + 4 * q);
                    }
                }
            }
            if (jj + 2 < SC_TT) {
                dl_nn = shfl_row(dmine, jj + 2);
                du_nn = shfl_row(dumine, jj + 2);
            }
            if (jj + 1 < SC_TT) make_exps(en, dl_n);
            if (jj >= 2 && (jj & 1) == 0) finalize(jrow(jj - 1) >> 1, mine_p, recv_p);
            const float2 du2 = make_float2(du_c, du_c);
            float2 ya = make_float2(0.f, 0.f), yb = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float2 bu0 = __fmul2_rn(du2, make_float2(Bc[q].x, Bc[q].y));
                const float2 bu1 = __fmul2_rn(du2, make_float2(Bc[q].z, Bc[q].w));
                h2[2 * q] = __ffma2_rn(e[2 * q], h2[2 * q], bu0);
                h2[2 * q + 1] = __ffma2_rn(e[2 * q + 1], h2[2 * q + 1], bu1);
                if (WY) {
                    ya = __ffma2_rn(h2[2 * q], make_float2(Cc[q].x, Cc[q].y), ya);
                    yb = __ffma2_rn(h2[2 * q + 1], make_float2(Cc[q].z, Cc[q].w), yb);
                }
            }
            const float2 yab = __fadd2_rn(ya, yb);
            const float ypart = yab.x + yab.y;
            if ((jj & 1) == 0) {
                yprev = ypart;
            } else if (WY) {
                // rows {2k, 2k+1} are complete in both lane halves: half h will finalise row 2k + h (the row it
                // prepared); it needs the partner's partial sum for that row.
                const float y_r0 = REV ? ypart : yprev;  // partial sums of rows 2k / 2k+1 over this lane's 8 states
                const float y_r1 = REV ? yprev : ypart;
                mine_p = half ? y_r1 : y_r0;
                recv_p = __shfl_xor_sync(FULL, half ? y_r0 : y_r1, 16);
            }
            if (jj + 1 < SC_TT) {
#pragma unroll
                for (int q = 0; q < 4; ++q) e[q] = en[q];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    Bc[q] = Bn[q];
                    Cc[q] = Cn[q];
                }
                du_c = du_n;
                dl_n = dl_nn;
                du_n = du_nn;
            }
        }
        finalize(jrow(SC_TT - 1) >> 1, mine_p, recv_p);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);  // inputs of this stage are consumed
        // ---- flush this warp's staged y rows: 2 lanes x 16 B cover one 32-byte row segment
#pragma unroll
        for (int pl = 0; pl < P; ++pl) {
            const int row = lane >> 1;
            const int seg = lane & 1;
            if (WY && row < nvalid) {
                const uint4 v = *reinterpret_cast<const uint4*>(sy_w + pl * SC_TT * 16 + row * 16 + seg * 8);
                const size_t off = (size_t(b) * L + t0 + row) * (2 * size_t(p.di)) + seg * 8;
                *reinterpret_cast<uint4*>(ywarp + pl * y_plane + off) = v;
            }
        }
        __syncwarp();
        if (++stage == SC_STAGES) {
            stage = 0;
            phase ^= 1;
        }
    }
    if (p.h_out) {
        float4* hp = reinterpret_cast<float4*>(p.h_out + ((size_t(dir) * p.batch + b) * p.di + d) * SC_NS + half * 8);
#pragma unroll
        for (int q = 0; q < 2; ++q) hp[q] = make_float4(h2[2 * q].x, h2[2 * q].y, h2[2 * q + 1].x, h2[2 * q + 1].y);
    }
    if (p.sum_delta) {
        const float tot = sdl + __shfl_xor_sync(FULL, sdl, 16);
        if (half == 0) p.sum_delta[(size_t(dir) * p.batch + b) * p.di + d] = tot;
    }
}

// CH channels per CTA = CH/16 warps, 2 lanes per channel; no producer warp (the duty rotates over the warps).
#ifdef MTN_SCAN_ABLATIONS
__device__ unsigned long long g_scan_dbg[3 * 8192];  // per CTA: smid, start ns, end ns (timing experiments only)
#endif

template <int P, int R, int NDBL, typename ZT, int CH, int KP, bool WY>
__global__ void __launch_bounds__(CH * 2, CH == 128 ? 2 : 7)
scan_kernel(const __grid_constant__ CUtensorMap mapU, const __grid_constant__ CUtensorMap mapZ,
            const __grid_constant__ CUtensorMap mapD, const ScanParams p) {
    using SM = ScanSmem<P, NDBL, ZT, CH>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // align inside the shared window without leaving the shared address space (keeps LDS/STS, not generic LD/ST)
    uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::STAGES * SM::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + SM::STAGES;
    __nv_bfloat16* sy = reinterpret_cast<__nv_bfloat16*>(smem + SM::STAGES * SM::STAGE_BYTES + 2 * SM::STAGES * 8);

    const int tid = threadIdx.x;
    const int nchb = p.di / CH;
    // direction-major grid: blockIdx.x = dir * nchb + channel block (CTAs sharing an SM then mix both directions,
    // which measured ~15 % faster than pairing CTAs of the same direction)
    const int ch0 = (blockIdx.x % nchb) * CH;
    const int dir = p.dir0 + blockIdx.x / nchb;
    const int b = blockIdx.y;

    if (tid == 0) {
        tma_prefetch_desc(&mapU);
        tma_prefetch_desc(&mapZ);
        tma_prefetch_desc(&mapD);
        for (int s = 0; s < SM::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CH / 16);
        }
        fence_barrier_init();
    }
    __syncthreads();
#ifdef MTN_SCAN_ABLATIONS
    const int cta = blockIdx.y * gridDim.x + blockIdx.x;
    if (tid == 32 && cta < 8192) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_scan_dbg[3 * cta] = smid;
        g_scan_dbg[3 * cta + 1] = global_timer_ns();
    }
#endif
    if (dir == 1)
        scan_consumer_split<P, R, NDBL, ZT, CH, true, KP, WY>(smem, full_bar, empty_bar, sy, p, ch0, b, dir, &mapU, &mapZ,
                                                             &mapD);
    else
        scan_consumer_split<P, R, NDBL, ZT, CH, false, KP, WY>(smem, full_bar, empty_bar, sy, p, ch0, b, dir, &mapU, &mapZ,
                                                              &mapD);
#ifdef MTN_SCAN_ABLATIONS
    if (tid == 32 && cta < 8192) g_scan_dbg[3 * cta + 2] = global_timer_ns();
#endif
}

template <int P, int R, int NDBL, typename ZT, int CH, int KP, bool WY>
static int launch_scan(const mtn_scan_args* a, cudaStream_t stream) {
    using SM = ScanSmem<P, NDBL, ZT, CH>;
    const uint64_t M = uint64_t(a->batch) * a->L;
    CUtensorMap mapU, mapZ, mapD;
    {
        uint64_t dims[3] = {uint64_t(2) * a->di, M, uint64_t(P)};
        uint64_t str[2] = {uint64_t(2) * a->di * 2, M * 2 * a->di * 2};
        uint32_t box[3] = {uint32_t(CH), SC_TT, uint32_t(P)};
        if (!encode_tmap(&mapU, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a->u, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE))
            return MTN_ECUDA;
    }
    {
        uint64_t dims[2] = {uint64_t(a->ldz), M};
        uint64_t str[1] = {uint64_t(a->ldz) * sizeof(ZT)};
        uint32_t box[2] = {uint32_t(CH), SC_TT};
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
    ScanParams p = make_scan_params(a);
    const int ndirs = p.ndirs;
    auto kern = scan_kernel<P, R, NDBL, ZT, CH, KP, WY>;
    static std::atomic<unsigned long long> attr_done{0};   // per template instantiation, one bit per device
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), SM::TOTAL, attr_done, "scan")) return rc;
    dim3 grid(ndirs * (a->di / CH), a->batch, 1);
    kern<<<grid, CH * 2, SM::TOTAL, stream>>>(mapU, mapZ, mapD, p);
    MTN_CUDA_LAUNCH_CHECK("scan");
    return MTN_OK;
}


// ===============================================================================================================
// Tensor-core dt_proj variant (default when mtn_scan_args.dtp is given).
//
// Same split mapping and software pipeline as scan_consumer_split, but delta_pre = dt . W_dt^T -- R MACs per
// (step, channel), a quarter of the kernel's FMA-pipe work and its longest dependent chain -- moves to tcgen05:
//     D[128 channels (TMEM lanes), 16 steps (columns)] = W_dt[128, RP] (smem, K-major) x dt_tile[16, RP]^T
// once per 16-step tile, in split-bf16 form (all four hi/lo products, fp32 accumulate in TMEM).  The B operand (dt
// columns of the tile as hi | lo bf16) is written by the x_proj GEMM epilogue (MTN_EPI_XPROJ) and arrives through
// the same TMA ring as u / z / B / C; the A operand (W_dt rows of the CTA's 128 channels) is built once per CTA.
// TMEM row m <-> channel: warp w may only read TMEM lanes [32 (w%4), +32), so channel 16 w + c sits at row
// 32 (w%4) + 16 (w/4) + c: after tcgen05.ld the lanes of half (w/4) hold their own channel's 16 steps and hand the
// partner half its 8 rows with one shuffle per row.
// ===============================================================================================================
template <int P, int NDBL, typename ZT, int RP>
struct ScanSmemTC {
    static constexpr int U_BYTES = P * SC_TT * SC_CH * 2;
    static constexpr int Z_BYTES = SC_TT * SC_CH * int(sizeof(ZT));
    static constexpr int D_BYTES = SC_TT * 2 * SC_NS * 4;          // only [B | C] of each row (dt comes from TMEM)
    static constexpr int DT_BYTES = 2 * RP * SC_TT * 2;            // [plane][k8][16 rows][8] bf16, UMMA canonical
    static constexpr int STAGE_BYTES = U_BYTES + Z_BYTES + D_BYTES + DT_BYTES;
    static constexpr int Y_BYTES = P * SC_TT * SC_CH * 2;
    static constexpr int WA_BYTES = 2 * RP * SC_CH * 2;            // [plane][k8][128 rows][8] bf16, UMMA canonical
    static constexpr int BAR_BYTES = 128;
    static constexpr int FIXED = 128 + BAR_BYTES + Y_BYTES + WA_BYTES;
    static constexpr int PER_CTA_MAX = (233472 - 2 * 1024) / 2;    // two CTAs per SM
    static constexpr int STAGES = (FIXED + 5 * STAGE_BYTES <= PER_CTA_MAX) ? 5 : 4;
    static constexpr int TOTAL = FIXED + STAGES * STAGE_BYTES;
    static_assert(FIXED + 4 * STAGE_BYTES <= PER_CTA_MAX, "scan smem: two CTAs per SM must fit");
};

// K-major, no swizzle: 8-row x 16-byte core matrices, 128 B apart along M/N (SBO), `lbo` bytes apart along K
__device__ __forceinline__ uint64_t make_smem_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return uint64_t((saddr & 0x3FFFF) >> 4) | (uint64_t(lbo >> 4) << 16) | (uint64_t(sbo >> 4) << 32) |
           (uint64_t(1) << 46);
}

template <int P, int R, int NDBL, typename ZT, bool REV, bool WY>
__device__ __forceinline__ void scan_consumer_tc(uint8_t* smem, uint64_t* full_bar, uint64_t* empty_bar,
                                                 uint64_t* dtfull_bar, uint8_t* wa, uint32_t tmem_base,
                                                 __nv_bfloat16* sy, const ScanParams& p, int ch0, int b, int dir,
                                                 const CUtensorMap* mapU, const CUtensorMap* mapZ,
                                                 const CUtensorMap* mapD, const CUtensorMap* mapT) {
    constexpr int RP = R <= 16 ? 16 : 32;
    using SM = ScanSmemTC<P, NDBL, ZT, RP>;
    constexpr int S = SM::STAGES;
    constexpr int AHEAD = S - 2;  // tile i + AHEAD is requested at the start of tile i
    // The dt_proj MMA of tile t is issued LEAD tiles before the tile whose steps interleave tile t's row preparation
    // (i.e. at the top of tile t-1-... see the loop), so its latency never shows; 4 TMEM accumulators of 16 columns
    // (tile t -> buffer t & 3) are enough because warp 0 only proceeds once every warp has finished tile i-2.
    constexpr int LEAD = S >= 5 ? 2 : 1;
    constexpr int NB = 2 * SC_NS; // floats per staged [B | C] row
    constexpr uint32_t FULL = 0xffffffffu;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int half = lane >> 4, cl = lane & 15;
    const int chl = warp * 16 + cl;
    const int L = p.L;
    const int Lb = (b == p.batch - 1) ? p.L_last : p.L;
    const int ntiles = (Lb + SC_TT - 1) / SC_TT;
    const int d = ch0 + chl;
    const size_t pd = size_t(dir) * p.di + d;
    const int hs = warp >> 2;                      // the lane half whose TMEM rows hold this warp's own channels
    const int srcl = (hs << 4) | cl;               // lane that holds channel cl's 16 steps after tcgen05.ld
    const uint32_t taddr = tmem_base + (uint32_t(32 * (warp & 3)) << 16);
    float sdl = 0.f;
    float2 h2[4], A2[4];
    {
        const float4* ap = reinterpret_cast<const float4*>(p.A2 + pd * SC_NS + half * 8);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float4 a = ap[q];
            A2[2 * q] = make_float2(a.x, a.y);
            A2[2 * q + 1] = make_float2(a.z, a.w);
        }
        if (p.h_in) {
            const float4* hp =
                reinterpret_cast<const float4*>(p.h_in + ((size_t(dir) * p.batch + b) * p.di + d) * SC_NS + half * 8);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float4 a = hp[q];
                h2[2 * q] = make_float2(a.x, a.y);
                h2[2 * q + 1] = make_float2(a.z, a.w);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) h2[q] = make_float2(0.f, 0.f);
        }
    }
    const float bias = p.dt_bias[pd];
    const float Dv = p.Dskip[pd];
    const size_t M = size_t(p.batch) * L;
    const size_t y_plane = M * 2 * p.di;
    __nv_bfloat16* ywarp = p.y + size_t(dir) * p.di + ch0 + warp * 16;
    uint16_t* sy_w = reinterpret_cast<uint16_t*>(sy) + warp * (P * SC_TT * 16);
    const int src0 = cl, src1 = cl | 16;

    float dmine[SC_TT / 2], dumine[SC_TT / 2], umine[SC_TT / 2];
    uint32_t raw[SC_TT];  // delta_pre of the tile being prepared, one TMEM row (channel) x 16 steps
#pragma unroll
    for (int j = 0; j < SC_TT; ++j) raw[j] = 0u;

    auto issue_tile = [&](int i2, int stg) {
        const int tile2 = REV ? (ntiles - 1 - i2) : i2;
        const int row0 = b * L + tile2 * SC_TT;
        mbar_arrive_expect_tx(&full_bar[stg], SM::STAGE_BYTES);
        uint8_t* dst = smem + stg * SM::STAGE_BYTES;
        tma_load_3d(dst, mapU, &full_bar[stg], dir * p.di + ch0, row0, 0);
        tma_load_2d(dst + SM::U_BYTES, mapZ, &full_bar[stg], p.z_col0 + ch0, row0);
        tma_load_2d(dst + SM::U_BYTES + SM::Z_BYTES, mapD, &full_bar[stg], dir * p.n_dbl + R, row0);
        uint8_t* dt = dst + SM::U_BYTES + SM::Z_BYTES + SM::D_BYTES;
#pragma unroll
        for (int c = 0; c < 2 * RP / 8; ++c)   // chunk c = plane * (RP/8) + k8 -> one [16 rows][16 B] slab each
            tma_load_2d(dt + c * (SC_TT * 16), mapT, &full_bar[stg], dir * 2 * RP + c * 8, row0);
    };
    // one elected thread: D[buf] = W_dt x dt_tile(stage)^T, completion signalled on dtfull_bar[buf]
    auto issue_dt_mma = [&](int stg, int buf) {
        constexpr uint32_t idesc = make_idesc_bf16(SC_CH, SC_TT);
        tc_fence_after();
        const uint32_t a0 = smem_u32(wa);
        const uint32_t b0 = smem_u32(smem + stg * SM::STAGE_BYTES + SM::U_BYTES + SM::Z_BYTES + SM::D_BYTES);
        const uint32_t dcol = tmem_base + buf * SC_TT;
        constexpr uint32_t A_SLAB = SC_CH * 16, B_SLAB = SC_TT * 16, A_PLANE = (RP / 8) * A_SLAB,
                           B_PLANE = (RP / 8) * B_SLAB;
#pragma unroll
        for (int kk = 0; kk < RP / 16; ++kk) {
            const uint64_t a_hi = make_smem_desc_nosw(a0 + 2 * kk * A_SLAB, A_SLAB, 128);
            const uint64_t a_lo = make_smem_desc_nosw(a0 + A_PLANE + 2 * kk * A_SLAB, A_SLAB, 128);
            const uint64_t b_hi = make_smem_desc_nosw(b0 + 2 * kk * B_SLAB, B_SLAB, 128);
            const uint64_t b_lo = make_smem_desc_nosw(b0 + B_PLANE + 2 * kk * B_SLAB, B_SLAB, 128);
            tc_mma_bf16(dcol, a_hi, b_hi, idesc, kk > 0 ? 1u : 0u);
            tc_mma_bf16(dcol, a_lo, b_hi, idesc, 1u);
            tc_mma_bf16(dcol, a_hi, b_lo, idesc, 1u);
            tc_mma_bf16(dcol, a_lo, b_lo, idesc, 1u);  // free here (tiny MMA), keeps delta_pre at full fp32 accuracy
        }
        tc_commit(&dtfull_bar[buf]);
    };
    auto load_raw = [&](int t) {   // accumulator of tile t (processing order) -> registers, whole warp
        mbar_wait(&dtfull_bar[t & 3], uint32_t(t >> 2) & 1u);
        tc_fence_after();
        tmem_ld_x16(taddr + (t & 3) * SC_TT, raw);
        tmem_ld_wait();
        tc_fence_before();
    };
    auto prep_row = [&](int k, const uint8_t* st, int nvalid_t) {
        // delta_pre of row 2k + half: own register for the lanes of half hs, one shuffle from lane srcl for the others
        const uint32_t for_partner = hs ? raw[2 * k] : raw[2 * k + 1];
        const uint32_t own = hs ? raw[2 * k + 1] : raw[2 * k];
        const uint32_t got = __shfl_sync(FULL, for_partner, srcl);
        const float dtv = __uint_as_float(half == hs ? own : got);
        const __nv_bfloat16* su_h = reinterpret_cast<const __nv_bfloat16*>(st) + half * SC_CH + chl;
        float dl = softplus_1mufu(bias + dtv);
        dl = (2 * k + half < nvalid_t) ? dl : 0.f;
        float uval = bf16_bits_to_float(su_h + k * 2 * SC_CH);
        if (P == 2) uval += bf16_bits_to_float(su_h + SC_TT * SC_CH + k * 2 * SC_CH);
        sdl += dl;
        dmine[k] = dl;
        dumine[k] = dl * uval;
        umine[k] = uval;
    };
    auto make_exps = [&](float2(&ex)[4], float dl) {
        const float2 dl2 = make_float2(dl, dl);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 a = __fmul2_rn(dl2, A2[q]);
            ex[q] = make_float2(ex2_approx(a.x), ex2_approx(a.y));
        }
    };

    if (tid == 0) {
#pragma unroll
        for (int t = 0; t < AHEAD; ++t)
            if (t < ntiles) issue_tile(t, t);
    }
    int stage = 0;
    uint32_t phase = 0;
    {
        const int tile0 = REV ? (ntiles - 1) : 0;
        mbar_wait(&full_bar[0], 0);
        if (warp == 0) {
            if (lane == 0) issue_dt_mma(0, 0);
#pragma unroll
            for (int t = 1; t < LEAD; ++t) {
                if (t < ntiles) {
                    mbar_wait(&full_bar[t], 0);
                    if (lane == 0) issue_dt_mma(t, t);
                }
            }
            __syncwarp();
        }
        load_raw(0);
        const int nv0 = min(SC_TT, Lb - tile0 * SC_TT);
#pragma unroll
        for (int k = 0; k < SC_TT / 2; ++k) prep_row(k, smem, nv0);
    }
    for (int i = 0; i < ntiles; ++i) {
        const int tile = REV ? (ntiles - 1 - i) : i;
        const int t0 = tile * SC_TT;
        const int nvalid = min(SC_TT, Lb - t0);
        const uint8_t* st = smem + stage * SM::STAGE_BYTES;
        const ZT* sz = reinterpret_cast<const ZT*>(st + SM::U_BYTES);
        const float* sd = reinterpret_cast<const float*>(st + SM::U_BYTES + SM::Z_BYTES);
        const bool has_next = i + 1 < ntiles;
        const int stage_n = (stage + 1 == S) ? 0 : stage + 1;
        const uint8_t* st_n = smem + stage_n * SM::STAGE_BYTES;
        int nvalid_n = 0;
        if (warp == (i & 7)) {
            // The producer duties (TMA requests, dt_proj MMA) rotate over the 8 warps, one tile each: the warps advance
            // in lock step at tile granularity, so a fixed producer warp would slow every tile by its extra work.
            // Whole warp waits (converged): all 8 warps finished tile i-2, i.e. its smem stage is free for tile
            // i+AHEAD and the TMEM accumulator about to be overwritten (read during tile i-2 at the latest) is free
            if (i >= 2) mbar_wait(&empty_bar[(stage + S - 2) % S], uint32_t((i - 2) / S) & 1u);
            if (lane == 0 && i + AHEAD < ntiles) issue_tile(i + AHEAD, (stage + AHEAD) % S);
            if (i + LEAD < ntiles) {
                const int tl = i + LEAD;                       // its inputs were requested AHEAD - LEAD tiles ago
                mbar_wait(&full_bar[tl % S], uint32_t(tl / S) & 1u);
                if (lane == 0) issue_dt_mma(tl % S, tl & 3);
            }
            __syncwarp();
        }
        if (has_next) {
            const int tile_n = REV ? (tile - 1) : (tile + 1);
            nvalid_n = min(SC_TT, Lb - tile_n * SC_TT);
            mbar_wait(&full_bar[stage_n], (stage + 1 == S) ? (phase ^ 1) : phase);
        }

        const float* sdB = sd + half * 8;
        auto jrow = [](int jj) { return REV ? (SC_TT - 1 - jj) : jj; };
        auto shfl_row = [&](const float(&arr)[SC_TT / 2], int jj) {
            const int j = jrow(jj);
            return __shfl_sync(FULL, arr[j >> 1], (j & 1) ? src1 : src0);
        };
        float2 e[4];
        float4 Bc[2], Cc[2];
        float du_c, dl_n, du_n;
        {
            const float dl0 = shfl_row(dmine, 0);
            du_c = shfl_row(dumine, 0);
            dl_n = shfl_row(dmine, 1);
            du_n = shfl_row(dumine, 1);
            const float* brow = sdB + jrow(0) * NB;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                Bc[q] = *reinterpret_cast<const float4*>(brow + 4 * q);
                Cc[q] = *reinterpret_cast<const float4*>(brow + SC_NS + 4 * q);
            }
            make_exps(e, dl0);
        }
        auto finalize = [&](int k, float mine, float recv) {
            if (WY) {
                const int row = 2 * k + half;
                const float zval = ldz(sz + row * SC_CH + chl);
                const float y = fmaf(Dv, umine[k], mine + recv) * (0.5f * zval);
                const uint32_t hi = f2bf_lo(y);
                sts_b16(sy_w + row * 16 + cl, hi);
                if (P == 2) sts_b16(sy_w + SC_TT * 16 + row * 16 + cl, f2bf_lo(y - __uint_as_float(hi << 16)));
            }
            prep_row(k, st_n, nvalid_n);
        };
        float yprev = 0.f, mine_p = 0.f, recv_p = 0.f;
#pragma unroll
        for (int jj = 0; jj < SC_TT; ++jj) {
            float4 Bn[2], Cn[2];
            float2 en[4];
            float dl_nn = 0.f, du_nn = 0.f;
            if (jj + 1 < SC_TT) {
                const float* brow = sdB + jrow(jj + 1) * NB;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    Bn[q] = *reinterpret_cast<const float4*>(brow + 4 * q);
                    Cn[q] = *reinterpret_cast<const float4*>(brow + SC_NS + 4 * q);
                }
            }
            if (jj + 2 < SC_TT) {
                dl_nn = shfl_row(dmine, jj + 2);
                du_nn = shfl_row(dumine, jj + 2);
            }
            if (jj + 1 < SC_TT) make_exps(en, dl_n);
            if (jj == 2 && has_next) load_raw(i + 1);  // that MMA was issued LEAD - 1 tiles + 2 steps ago
            if (jj >= 2 && (jj & 1) == 0) finalize(jrow(jj - 1) >> 1, mine_p, recv_p);
            const float2 du2 = make_float2(du_c, du_c);
            float2 ya = make_float2(0.f, 0.f), yb = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float2 bu0 = __fmul2_rn(du2, make_float2(Bc[q].x, Bc[q].y));
                const float2 bu1 = __fmul2_rn(du2, make_float2(Bc[q].z, Bc[q].w));
                h2[2 * q] = __ffma2_rn(e[2 * q], h2[2 * q], bu0);
                h2[2 * q + 1] = __ffma2_rn(e[2 * q + 1], h2[2 * q + 1], bu1);
                if (WY) {
                    ya = __ffma2_rn(h2[2 * q], make_float2(Cc[q].x, Cc[q].y), ya);
                    yb = __ffma2_rn(h2[2 * q + 1], make_float2(Cc[q].z, Cc[q].w), yb);
                }
            }
            const float2 yab = __fadd2_rn(ya, yb);
            const float ypart = yab.x + yab.y;
            if ((jj & 1) == 0) {
                yprev = ypart;
            } else if (WY) {
                const float y_r0 = REV ? ypart : yprev;
                const float y_r1 = REV ? yprev : ypart;
                mine_p = half ? y_r1 : y_r0;
                recv_p = __shfl_xor_sync(FULL, half ? y_r0 : y_r1, 16);
            }
            if (jj + 1 < SC_TT) {
#pragma unroll
                for (int q = 0; q < 4; ++q) e[q] = en[q];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    Bc[q] = Bn[q];
                    Cc[q] = Cn[q];
                }
                du_c = du_n;
                dl_n = dl_nn;
                du_n = du_nn;
            }
        }
        finalize(jrow(SC_TT - 1) >> 1, mine_p, recv_p);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
#pragma unroll
        for (int pl = 0; pl < P; ++pl) {
            const int row = lane >> 1;
            const int seg = lane & 1;
            if (WY && row < nvalid) {
                const uint4 v = *reinterpret_cast<const uint4*>(sy_w + pl * SC_TT * 16 + row * 16 + seg * 8);
                const size_t off = (size_t(b) * L + t0 + row) * (2 * size_t(p.di)) + seg * 8;
                *reinterpret_cast<uint4*>(ywarp + pl * y_plane + off) = v;
            }
        }
        __syncwarp();
        if (++stage == S) {
            stage = 0;
            phase ^= 1;
        }
    }
    if (p.h_out) {
        float4* hp = reinterpret_cast<float4*>(p.h_out + ((size_t(dir) * p.batch + b) * p.di + d) * SC_NS + half * 8);
#pragma unroll
        for (int q = 0; q < 2; ++q) hp[q] = make_float4(h2[2 * q].x, h2[2 * q].y, h2[2 * q + 1].x, h2[2 * q + 1].y);
    }
    if (p.sum_delta) {
        const float tot = sdl + __shfl_xor_sync(FULL, sdl, 16);
        if (half == 0) p.sum_delta[(size_t(dir) * p.batch + b) * p.di + d] = tot;
    }
}

template <int P, int R, int NDBL, typename ZT, bool WY>
__global__ void __launch_bounds__(256, 2)
scan_kernel_tc(const __grid_constant__ CUtensorMap mapU, const __grid_constant__ CUtensorMap mapZ,
               const __grid_constant__ CUtensorMap mapD, const __grid_constant__ CUtensorMap mapT, const ScanParams p) {
    constexpr int RP = R <= 16 ? 16 : 32;
    using SM = ScanSmemTC<P, NDBL, ZT, RP>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint8_t* fixed = smem + SM::STAGES * SM::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(fixed);
    uint64_t* empty_bar = full_bar + SM::STAGES;
    uint64_t* dtfull_bar = empty_bar + SM::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dtfull_bar + 4);
    __nv_bfloat16* sy = reinterpret_cast<__nv_bfloat16*>(fixed + SM::BAR_BYTES);
    uint8_t* wa = fixed + SM::BAR_BYTES + SM::Y_BYTES;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int nchb = p.di / SC_CH;
    const int ch0 = (blockIdx.x % nchb) * SC_CH;
    const int dir = p.dir0 + blockIdx.x / nchb;
    const int b = blockIdx.y;

    if (tid == 0) {
        tma_prefetch_desc(&mapU);
        tma_prefetch_desc(&mapZ);
        tma_prefetch_desc(&mapD);
        tma_prefetch_desc(&mapT);
        for (int s = 0; s < SM::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 8);
        }
        for (int t = 0; t < 4; ++t) mbar_init(&dtfull_bar[t], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 64);
    if (tid < SC_CH) {
        // A operand: row m = tid holds W_dt of channel 16 w + c with m = 32 (w%4) + 16 (w/4) + c; split hi | lo bf16
        const int m = tid;
        const int q = m >> 5, r = m & 31;
        const int w = q + 4 * (r >> 4);
        const size_t pdm = size_t(dir) * p.di + ch0 + 16 * w + (r & 15);
#pragma unroll
        for (int k8 = 0; k8 < RP / 8; ++k8) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
                const int k = k8 * 8 + 2 * e2;
                const float v0 = k < R ? p.w_dt[pdm * R + k] : 0.f;
                const float v1 = k + 1 < R ? p.w_dt[pdm * R + k + 1] : 0.f;
                const uint32_t h0 = f2bf_lo(v0), h1 = f2bf_lo(v1);
                const uint32_t l0 = f2bf_lo(v0 - __uint_as_float(h0 << 16)), l1 = f2bf_lo(v1 - __uint_as_float(h1 << 16));
                hi[e2] = h0 | (h1 << 16);
                lo[e2] = l0 | (l1 << 16);
            }
            *reinterpret_cast<uint4*>(wa + k8 * (SC_CH * 16) + m * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(wa + (RP / 8 + k8) * (SC_CH * 16) + m * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (dir == 1)
        scan_consumer_tc<P, R, NDBL, ZT, true, WY>(smem, full_bar, empty_bar, dtfull_bar, wa, tmem_base, sy, p, ch0, b, dir,
                                                  &mapU, &mapZ, &mapD, &mapT);
    else
        scan_consumer_tc<P, R, NDBL, ZT, false, WY>(smem, full_bar, empty_bar, dtfull_bar, wa, tmem_base, sy, p, ch0, b,
                                                   dir, &mapU, &mapZ, &mapD, &mapT);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 64);
    }
}

template <int P, int R, int NDBL, typename ZT, bool WY>
static int launch_scan_tc(const mtn_scan_args* a, cudaStream_t stream) {
    constexpr int RP = R <= 16 ? 16 : 32;
    using SM = ScanSmemTC<P, NDBL, ZT, RP>;
    const uint64_t M = uint64_t(a->batch) * a->L;
    CUtensorMap mapU, mapZ, mapD, mapT;
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
        uint32_t box[2] = {uint32_t(2 * SC_NS), SC_TT};   // [B | C] only
        if (!encode_tmap(&mapD, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a->dbl, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE))
            return MTN_ECUDA;
    }
    {
        uint64_t dims[2] = {uint64_t(4 * RP), M};           // [dir][plane][RP] per row
        uint64_t str[1] = {uint64_t(4 * RP) * 2};
        uint32_t box[2] = {8, SC_TT};                       // one 16-byte K chunk x 16 rows = one canonical slab
        if (!encode_tmap(&mapT, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->dtp, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE))
            return MTN_ECUDA;
    }
    ScanParams p = make_scan_params(a);
    const int ndirs = p.ndirs;
    auto kern = scan_kernel_tc<P, R, NDBL, ZT, WY>;
    static std::atomic<unsigned long long> attr_done{0};   // per template instantiation, one bit per device
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), SM::TOTAL, attr_done, "scan(tc)")) return rc;
    dim3 grid(ndirs * (a->di / SC_CH), a->batch, 1);
    kern<<<grid, 256, SM::TOTAL, stream>>>(mapU, mapZ, mapD, mapT, p);
    MTN_CUDA_LAUNCH_CHECK("scan(tc)");
    return MTN_OK;
}


}  // namespace mtn
#include "mtn_scan_pair.cuh"
namespace mtn {

template <int P, int R, int NDBL, typename ZT>
static int dispatch_scan_variant(const mtn_scan_args* a, cudaStream_t s) {
    int variant = 0;
    if (const char* v = getenv("MTN_SCAN_VARIANT")) variant = atoi(v);
#ifdef MTN_SCAN_DEV   // experiments on the pair kernel (dev builds of tools/devbuild.sh only)
    if (a->y) {
        switch (variant) {
            case 61: return launch_scan_pair<P, R, NDBL, ZT, true, 1>(a, s);               // 1 state pair on the FMA-pipe exp2
            case 62: return launch_scan_pair<P, R, NDBL, ZT, true, 2>(a, s);               // 2 state pairs
            case 68: return launch_scan_pair<P, R, NDBL, ZT, true, 0, false, 32>(a, s);    // 64 ns helper poll (r02 start)
            case 69: return launch_scan_pair<P, R, NDBL, ZT, true, 0, false, 64>(a, s);    // the other softplus form
            case 70: return launch_scan_duo<P, R, NDBL, ZT>(a, s);                          // two 8-state recurrence warps per group
            case 71: return launch_scan_duo<P, R, NDBL, ZT, 4>(a, s);                       // ... timing only: no MUFU
            case 74: return launch_scan_pair<P, R, NDBL, ZT, true, 0, false, 128>(a, s);   // recurrence warps = the higher warp ids
            case 75: return launch_scan_duo<P, R, NDBL, ZT, 128>(a, s);                     // duo, helper = the lowest warp ids
            case 72: return launch_scan_duo<P, R, NDBL, ZT, 3>(a, s);                       // ... timing only: idle helper
            case 73: return launch_scan_duo<P, R, NDBL, ZT, 7>(a, s);                       // ... timing only: idle helper, no MUFU
            // timing-only ablations (WRONG results)
            case 63: return launch_scan_pair<P, R, NDBL, ZT, true, 0, false, 3>(a, s);     // helper: no dt_proj/softplus/gate
            case 64: return launch_scan_pair<P, R, NDBL, ZT, true, 0, false, 4>(a, s);     // recurrence: no MUFU
            case 65: return launch_scan_pair<P, R, NDBL, ZT, true, 0, false, 2>(a, s);     // helper: no gate
            case 66: return launch_scan_pair<P, R, NDBL, ZT, true, 0, false, 1>(a, s);     // helper: no dt_proj/softplus
            case 67: return launch_scan_pair<P, R, NDBL, ZT, true, 0, false, 7>(a, s);     // 63 + 64
            default: break;
        }
    }
#endif
    // Recurrence / helper warp pairs (mtn_scan_pair.cuh) for every pass that writes y.  Measured on B200
    // (tools/scan_bench.py, DESIGN.md 4.1, profiles/r02/scan_*.jsonl): faster than the split mapping at every shipped
    // shape (S fp32 0.90 -> 0.82 ms, L fp32 2.00 -> 1.73, L bf16 3.29 -> 2.82).  The summary pass (no y: nothing for
    // the helper to take over) measures the same in both mappings (0.72 ms) and stays on the split mapping.
    // variant 5 / 6 force the pair kernel (also for summary passes), 2 / 3 the split mapping.
    if (variant == 5 || variant == 6 || (variant == 0 && a->y != nullptr)) {
        if (!a->y) return launch_scan_pair<P, R, NDBL, ZT, false>(a, s);
        return launch_scan_pair<P, R, NDBL, ZT, true>(a, s);
    }
    if (a->dtp && (variant == 0 || variant == 3)) {  // split mapping, dt_proj on the tensor cores
        if (!a->y) return launch_scan_tc<P, R, NDBL, ZT, false>(a, s);
        return launch_scan_tc<P, R, NDBL, ZT, true>(a, s);
    }
    // 128-channel CTAs.  (Measured alternatives that did NOT pay, see DESIGN.md 4.1: 32-channel CTAs to even out the
    // SM load at small grids; four lanes per channel; direction-paired CTAs.)
    if (variant == 0 || variant == 2 || variant == 3) {
        if (!a->y) return launch_scan<P, R, NDBL, ZT, 128, 0, false>(a, s);
        return launch_scan<P, R, NDBL, ZT, 128, 0, true>(a, s);
    }
#ifdef MTN_SCAN_ABLATIONS
    if (a->y) {
        switch (variant) {
            case 4: return launch_scan<P, R, NDBL, ZT, 32, 0, true>(a, s);   // 32-channel CTAs
            case 11: return launch_scan<P, R, NDBL, ZT, 128, 1, true>(a, s);
            case 101: return launch_scan<P, R, NDBL, ZT, 128, 101, true>(a, s);
            case 102: return launch_scan<P, R, NDBL, ZT, 128, 102, true>(a, s);
            case 104: return launch_scan<P, R, NDBL, ZT, 128, 104, true>(a, s);
            case 103: return launch_scan<P, R, NDBL, ZT, 128, 103, true>(a, s);
            case 107: return launch_scan<P, R, NDBL, ZT, 128, 107, true>(a, s);
            default: break;
        }
    }
#endif
    set_error("scan: unknown MTN_SCAN_VARIANT=%d", variant);
    return MTN_EINVAL;
}

template <int P, typename ZT>
static int dispatch_scan_r(const mtn_scan_args* a, cudaStream_t s) {
#ifdef MTN_SCAN_DEV  // fast-iteration build: BASELINE config 2 / 3 instantiations only
    if (a->R == 16 && a->n_dbl == 48) return dispatch_scan_variant<P, 16, 48, ZT>(a, s);
    if (a->R == 32 && a->n_dbl == 64) return dispatch_scan_variant<P, 32, 64, ZT>(a, s);
    set_error("scan: MTN_SCAN_DEV build");
    return MTN_EINVAL;
#else
    if (a->R == 4 && a->n_dbl == 48) return dispatch_scan_variant<P, 4, 48, ZT>(a, s);
    if (a->R == 8 && a->n_dbl == 48) return dispatch_scan_variant<P, 8, 48, ZT>(a, s);
    if (a->R == 16 && a->n_dbl == 48) return dispatch_scan_variant<P, 16, 48, ZT>(a, s);
    if (a->R == 32 && a->n_dbl == 64) return dispatch_scan_variant<P, 32, 64, ZT>(a, s);
    set_error("scan: unsupported dt_rank R=%d / n_dbl=%d (supported: 4|8|16 with 48, 32 with 64)", a->R, a->n_dbl);
    return MTN_EINVAL;
#endif
}

}  // namespace mtn

#ifdef MTN_SCAN_ABLATIONS
extern "C" int mtn_debug_scan_times(unsigned long long* host_dst, int n_ctas) {
    return cudaMemcpyFromSymbol(host_dst, mtn::g_scan_dbg, sizeof(unsigned long long) * 3 * n_ctas) == cudaSuccess ? 0 : -2;
}
#endif

extern "C" int mtn_scan_fwd(const mtn_scan_args* a, mtn_stream_t stream) {
    using namespace mtn;
    MTN_REQUIRE(a && a->u && a->dbl && a->z && a->w_dt && a->dt_bias && a->A2 && a->Dskip, "scan: null pointer");
    MTN_REQUIRE(a->y || a->h_out, "scan: y == NULL (summary pass) needs h_out");
    MTN_REQUIRE(a->L_last >= 0 && a->L_last <= a->L, "scan: L_last=%d out of range (L=%d)", a->L_last, a->L);
    MTN_REQUIRE(a->batch > 0 && a->batch <= 65535 && a->L > 0, "scan: bad batch=%d L=%d", a->batch, a->L);
    MTN_REQUIRE(a->di > 0 && a->di % SC_CH == 0, "scan: di=%d must be a multiple of %d", a->di, SC_CH);
    MTN_REQUIRE(a->dir_mask >= 1 && a->dir_mask <= 3, "scan: dir_mask=%d", a->dir_mask);
    MTN_REQUIRE(a->ld_dbl % 4 == 0 && a->ld_dbl >= 2 * a->n_dbl, "scan: ld_dbl=%d too small for 2 x n_dbl=%d", a->ld_dbl,
                a->n_dbl);
    MTN_REQUIRE(a->ldz % 8 == 0 && a->z_col0 + a->di <= a->ldz, "scan: z view out of range");
    MTN_REQUIRE(uint64_t(a->batch) * a->L < (1ull << 31), "scan: too many tokens for 32-bit TMA coordinates");
    MTN_REQUIRE(!a->dtp || (reinterpret_cast<uintptr_t>(a->dtp) & 15) == 0, "scan: dtp must be 16-byte aligned");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (a->planes == 2 && !a->z_bf16) return dispatch_scan_r<2, float>(a, s);
    if (a->planes == 1 && a->z_bf16) return dispatch_scan_r<1, __nv_bfloat16>(a, s);
#ifndef MTN_SCAN_DEV
    if (a->planes == 1 && !a->z_bf16) return dispatch_scan_r<1, float>(a, s);
#endif
    set_error("scan: unsupported planes=%d z_bf16=%d", a->planes, a->z_bf16);
    return MTN_EINVAL;
}
