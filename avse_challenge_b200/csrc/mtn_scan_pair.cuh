// Selective scan, "pair" mapping (sm_100a): one RECURRENCE warp + one HELPER warp per 32 channels.
//
// Same arithmetic and citations as mtn_scan.cu (ssi.py:91-157, :187, bimamba.py:253).  Why a second mapping: the
// recurrence is bounded by the XU pipe (16 ex2 per (step, channel): 128 XU cycles per warp-step of 32 channels), but
// the split mapping of mtn_scan.cu spends 150 issue slots per 32 channel-steps on a serial path with 3 shuffles per
// step, so it runs at ~58 % of the XU floor.  Here the serial path is stripped to what depends on the state:
//   * recurrence warp (lane = channel, all 16 states in registers): per step 2 private LDS (delta, u), 8 broadcast
//     LDS.128 (B_t, C_t), 16 packed multiplies, 16 MUFU, 16 packed FMAs, a 4-op reduction and one STS of y: ~75 issue
//     slots per 128 XU cycles, no shuffles, no global memory access, no conversions.
//   * helper warp (the same 32 channels, but a whole 16-step tile at once, i.e. time-parallel work): dt_proj +
//     softplus -> delta, bf16 planes -> fp32 u, written to a pair-private shared-memory slot two tiles ahead of the
//     recurrence; afterwards gate (0.5 * silu(z)), hi/lo bf16 split and the global stores of y.  It also owns the
//     TMA ring (duty rotates over the 4 helpers) and sum_delta.
// The two warps of a pair share an SM sub-partition (warp w and w + 4), so the helper fills the issue slots the
// recurrence warp leaves idle while it waits on the XU pipe.  Hand-offs are pair-private mbarriers (prepped / ydone);
// only the TMA ring (u planes + [dt|B|C] rows) is shared by the CTA.  y overwrites delta in place in the slot.
//
// History (profiles/r01, profiles/r02): the first helper kept lane = channel for everything and needed
// 72 issue slots per warp-step next to the recurrence warp's 75; the helper below needs ~30, and the kernel went from
// issue-bound (XU 61 % busy) to XU-bound (81 % on the sub-partitions that hold two pairs).
#pragma once

namespace mtn {

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void sts_f32_nofence(void* p, float v) {
    // no "memory" clobber on purpose (see sts_b16): the slot is read by the helper only after the ydone barrier
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(smem_u32(p)), "f"(v));
}

// ------------------------------------------------------------------------------------------------ recurrence warp
// One instruction stream serves both directions (with one per direction and role, the four unrolled tile bodies, 80 KB,
// thrashed the instruction cache: 25 % of the warp samples were stall_no_inst): the slot row pointer and the [B | C] row
// pointer step by signed strides, everything else has immediate offsets.
// DS = slot row stride in floats (36: the helper's 16-byte row-segment accesses need an odd multiple of 16 bytes to stay
// bank-conflict free, and so do the rows this warp reads lane = channel).  KP = state pairs per step whose decay factor is computed by the
// FMA-pipe polynomial ex2_poly2 instead of MUFU.EX2 (balances the XU against the FMA pipe once the helper is light).
template <int S, int NB, int BOFF, bool WY, int ABL, bool RG = false, int DS = 32, int KP = 0>
__device__ __forceinline__ void scan_pair_recur(const uint8_t* ring, float* slots, int stage_bytes, int u_bytes,
                                                uint64_t* full_bar, uint64_t* empty_bar, uint64_t* prepped,
                                                uint64_t* ydone, const ScanParams& p, int w, int lane, int d, int b,
                                                int dir, int ntiles, int Lb) {
    constexpr int NSLOT = 3;
    const size_t pd = size_t(dir) * p.di + d;
    float2 h2[8], A2[8];
    {
        const float4* ap = reinterpret_cast<const float4*>(p.A2 + pd * SC_NS);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 a = ap[q];
            A2[2 * q] = make_float2(a.x, a.y);
            A2[2 * q + 1] = make_float2(a.z, a.w);
        }
        if (p.h_in) {
            const float4* hp = reinterpret_cast<const float4*>(p.h_in + ((size_t(dir) * p.batch + b) * p.di + d) * SC_NS);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 a = hp[q];
                h2[2 * q] = make_float2(a.x, a.y);
                h2[2 * q + 1] = make_float2(a.z, a.w);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) h2[q] = make_float2(0.f, 0.f);
        }
    }
    const float Dv = p.Dskip[pd];
    const int bc_first = dir ? (SC_TT - 1) * NB : 0;   // floats
    const int bc_step = dir ? -NB : NB;
    const int sl_first = dir ? (SC_TT - 1) * DS : 0;
    const int sl_step = dir ? -DS : DS;
    constexpr int UOFF = SC_TT * DS;            // u plane of the pair's slot, floats after the delta -> y plane
    constexpr int PAIRF = 2 * SC_TT * DS;       // floats per pair slot
    // polynomial exp2 needs delta * A2 >= -126 for its KP pairs: clamp delta for those pairs only (2^-126 ~ 0 either way)
    float dl_cap = 3.0e38f;
    if (KP > 0) {
        float amin = 0.f;
#pragma unroll
        for (int q = 0; q < KP; ++q) amin = fminf(amin, fminf(A2[q].x, A2[q].y));
        dl_cap = amin < 0.f ? -126.f / amin : 3.0e38f;
    }
    auto decay = [&](float dl, int q) -> float2 {
        if (q < KP) {
            const float dc = fminf(dl, dl_cap);
            return ex2_poly2(make_float2(dc, dc), A2[q]);
        }
        const float2 a = __fmul2_rn(make_float2(dl, dl), A2[q]);
        return (ABL & 4) ? __ffma2_rn(a, make_float2(0.5f, 0.5f), make_float2(1.f, 1.f))
                         : make_float2(ex2_approx(a.x), ex2_approx(a.y));
    };

    for (int i = 0; i < ntiles; ++i) {
        const int stg = i % S, sl = i % NSLOT;
        mbar_wait(&prepped[w * NSLOT + sl], uint32_t(i / NSLOT) & 1u);
        mbar_wait(&full_bar[stg], uint32_t(i / S) & 1u);  // complete long ago; orders this warp after the TMA writes
        // row pointer into the pair's slot: delta -> y at [row][32], u fp32 SC_TT*32 floats further
        float* psl = slots + sl * (4 * PAIRF) + w * PAIRF + lane + sl_first;
        const float* pbc = reinterpret_cast<const float*>(ring + stg * stage_bytes + u_bytes) + BOFF + bc_first;

        // Ragged tile (only the sequence's last 16-row tile can be one; it is processed first in the backward direction):
        // walk just its valid rows with a rolled, unpipelined step instead of 16 masked ones.  Many short sequences
        // (DPMamba's inter model: 34 steps = 2 tiles + 2 rows) otherwise pay 48 steps for 34: 2.51 -> 2.20 ms at 8000 x 34.
        // RG is a template parameter because merely compiling this branch into the long-sequence kernel cost it 2 %
        // (BASELINE config 2: 0.901 -> 0.919 ms); the launcher enables it for short sequences only.
        const int nvalid = RG ? min(SC_TT, Lb - (dir ? (ntiles - 1 - i) : i) * SC_TT) : SC_TT;
        if (RG && nvalid < SC_TT) {
            float* rowbase = psl - sl_first;
            const float* bcbase = pbc - bc_first;
#pragma unroll 1
            for (int s2 = 0; s2 < nvalid; ++s2) {
                const int row = dir ? (nvalid - 1 - s2) : s2;
                float* py = rowbase + row * DS;
                const float dl = py[0], uu = py[UOFF];
                const float* pb = bcbase + row * NB;
                const float du = dl * uu;
                const float2 du2 = make_float2(du, du);
                float2 ya[4] = {make_float2(Dv * uu, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 Bv = *reinterpret_cast<const float4*>(pb + 4 * q);
                    const float2 e0 = decay(dl, 2 * q), e1 = decay(dl, 2 * q + 1);
                    h2[2 * q] = __ffma2_rn(e0, h2[2 * q], __fmul2_rn(du2, make_float2(Bv.x, Bv.y)));
                    h2[2 * q + 1] = __ffma2_rn(e1, h2[2 * q + 1], __fmul2_rn(du2, make_float2(Bv.z, Bv.w)));
                    if (WY) {
                        const float4 Cv = *reinterpret_cast<const float4*>(pb + SC_NS + 4 * q);
                        // same accumulator assignment and summation tree as the unrolled path: bit-identical y
                        ya[2 * (q & 1)] = __ffma2_rn(h2[2 * q], make_float2(Cv.x, Cv.y), ya[2 * (q & 1)]);
                        ya[2 * (q & 1) + 1] = __ffma2_rn(h2[2 * q + 1], make_float2(Cv.z, Cv.w), ya[2 * (q & 1) + 1]);
                    }
                }
                if (WY) {
                    const float2 sy = __fadd2_rn(__fadd2_rn(ya[0], ya[1]), __fadd2_rn(ya[2], ya[3]));
                    sts_f32_nofence(py, sy.x + sy.y);
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&ydone[w * NSLOT + sl]);
                mbar_arrive(&empty_bar[stg]);
            }
            continue;
        }

        // Software pipeline: the decay factors run TWO steps ahead in their own registers (e_c: this step, e_n: next
        // step, both already requested), so a MUFU never waits for the FMA that consumes its predecessor and the XU
        // queue is fed independently of the state update chain.
        float2 e_c[8], e_n[8];
        float4 Bq[4];
        float dl_c = psl[0], u_c = psl[UOFF];
        float dl_n = psl[sl_step], u_n = psl[sl_step + UOFF];
        {
#pragma unroll
            for (int q = 0; q < 8; ++q) e_c[q] = decay(dl_c, q);
#pragma unroll
            for (int q = 0; q < 8; ++q) e_n[q] = decay(dl_n, q);
#pragma unroll
            for (int q = 0; q < 4; ++q) Bq[q] = *reinterpret_cast<const float4*>(pbc + 4 * q);
        }
        // Fully unrolled: a rolled 2 x 8 form (smaller code) measured 8 % slower (register rotation, lost overlap
        // across the loop edge).
#pragma unroll
        for (int jj = 0; jj < SC_TT; ++jj) {
            float4 Cq[4];
            if (WY) {
#pragma unroll
                for (int q = 0; q < 4; ++q) Cq[q] = *reinterpret_cast<const float4*>(pbc + SC_NS + 4 * q);
            }
            pbc += bc_step;
            float* py = psl;
            psl += sl_step;
            float dl_nn = 0.f, u_nn = 0.f;
            if (jj + 2 < SC_TT) {
                dl_nn = psl[sl_step];
                u_nn = psl[sl_step + UOFF];
            }
            const float du = dl_c * u_c;
            const float2 du2 = make_float2(du, du);
            float2 bu[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                bu[2 * q] = __fmul2_rn(du2, make_float2(Bq[q].x, Bq[q].y));
                bu[2 * q + 1] = __fmul2_rn(du2, make_float2(Bq[q].z, Bq[q].w));
            }
            if (jj + 1 < SC_TT) {  // B of the next step: its registers are free now
#pragma unroll
                for (int q = 0; q < 4; ++q) Bq[q] = *reinterpret_cast<const float4*>(pbc + 4 * q);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                h2[q] = __ffma2_rn(e_c[q], h2[q], bu[q]);
                e_c[q] = e_n[q];
                if (jj + 2 < SC_TT) e_n[q] = decay(dl_nn, q);
            }
            if (WY) {
                float2 y0 = make_float2(Dv * u_c, 0.f), y1 = make_float2(0.f, 0.f), y2 = y1, y3 = y1;
#pragma unroll
                for (int q = 0; q < 4; q += 2) {
                    y0 = __ffma2_rn(h2[2 * q], make_float2(Cq[q].x, Cq[q].y), y0);
                    y1 = __ffma2_rn(h2[2 * q + 1], make_float2(Cq[q].z, Cq[q].w), y1);
                    y2 = __ffma2_rn(h2[2 * q + 2], make_float2(Cq[q + 1].x, Cq[q + 1].y), y2);
                    y3 = __ffma2_rn(h2[2 * q + 3], make_float2(Cq[q + 1].z, Cq[q + 1].w), y3);
                }
                const float2 s = __fadd2_rn(__fadd2_rn(y0, y1), __fadd2_rn(y2, y3));
                sts_f32_nofence(py, s.x + s.y);
            }
            dl_c = dl_n;
            u_c = u_n;
            dl_n = dl_nn;
            u_n = u_nn;
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&ydone[w * NSLOT + sl]);
            mbar_arrive(&empty_bar[stg]);
        }
    }
    if (p.h_out) {
        float4* hp = reinterpret_cast<float4*>(p.h_out + ((size_t(dir) * p.batch + b) * p.di + d) * SC_NS);
#pragma unroll
        for (int q = 0; q < 4; ++q) hp[q] = make_float4(h2[2 * q].x, h2[2 * q].y, h2[2 * q + 1].x, h2[2 * q + 1].y);
    }
}

// --------------------------------------------------------------------------------------------------- helper warp
// None of the helper's work cares which lane owns which element, so every phase uses the layout that makes it cheap (a
// lane = channel helper pays scalar instructions per (step, channel): R FMAs of dt_proj, 13 for softplus, 16-bit loads /
// stores for u and y, one 4-byte cp.async per gate value):
//   * dt_proj: delta_pre[16 steps, 32 channels] = dt[16, R] x W_dt[32, R]^T as warp-level mma.sync.m16n8k16 (bf16 hi/lo
//     split of both operands, 4 passes, fp32 accumulate, dt_bias as the C operand): 16 HMMA per tile instead of
//     224 LDS/FFMA2/FADD.  The accumulator fragment leaves every lane 2 rows x 8 channels; the column -> channel
//     assignment of the four n-blocks is permuted (channel = col/2 + 4*(2*nb + col%2)) so that the fragment's STS.32
//     into the slot are bank-conflict free.
//   * softplus on that fragment in packed form (FFMA2 over the two accumulator columns): 15 slots per 2 elements.
//   * u planes -> fp32, gate, hi/lo split and the global stores in a (row, 8-channel segment) layout: 128-bit shared
//     loads / stores, packed converts, y goes from registers straight to global (no staging pass), silu(z) arrives by
//     16-byte cp.async that each lane issues for exactly the elements it gates itself.
// Slot rows are DS = 36 floats apart (144 B, an odd multiple of 16 B): the recurrence warp's row reads (lane =
// channel), the fragment stores and the 16-byte segment accesses are all conflict free.
constexpr int P2_DS = 36;

// DUO (mtn "duo" mapping below): TWO recurrence warps per 32 channels, 8 states each.  The pair slot then holds three planes
// [delta copy of half 0 -> y0 | delta copy of half 1 -> y1 | u] (each half overwrites ITS OWN delta copy with its partial y, so
// the halves never have to wait for each other inside a tile) and there are two slots instead of three.
template <int P, int NDBL, bool DUO = false>
struct ScanSmemPair {
    static constexpr int SLOTS = DUO ? 2 : 3;
    static constexpr int PLANES = DUO ? 3 : 2;
    static constexpr int U_BYTES = P * SC_TT * SC_CH * 2;
    static constexpr int NB = NDBL;
    static constexpr int D_BYTES = SC_TT * NB * 4;
    static constexpr int STAGE_BYTES = U_BYTES + D_BYTES;
    static constexpr int PAIR_SLOT_BYTES = PLANES * SC_TT * P2_DS * 4;   // [delta -> y | u] x 16 rows x 36 floats
    static constexpr int SLOT_BYTES = 4 * PAIR_SLOT_BYTES;
    static constexpr int ZBUF_BYTES = 4 * SC_TT * 32 * 4;                 // per pair 2 KB: [chunk j][lane][16 B]
    static constexpr int BAR_BYTES = (2 * 4 + 2 * 4 * SLOTS) * 8 + 16;
    static constexpr int FIXED = 128 + SLOTS * SLOT_BYTES + ZBUF_BYTES + BAR_BYTES;
    static constexpr int PER_CTA_MAX = 233472 / 2 - 1024;                 // two CTAs per SM
    static constexpr int STAGES = (FIXED + 4 * STAGE_BYTES <= PER_CTA_MAX) ? 4 : 3;
    static constexpr int TOTAL = FIXED + STAGES * STAGE_BYTES;
    static_assert(STAGE_BYTES % 128 == 0 && U_BYTES % 128 == 0, "TMA destinations must stay 128-byte aligned");
    static_assert(TOTAL <= PER_CTA_MAX, "two CTAs per SM must fit");
};

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// two fp32 -> one bf16x2 word, element 0 in the low half (F2FP on the ALU pipe)
__device__ __forceinline__ uint32_t pack_bf16x2(float e0, float e1) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));
    return r;
}
__device__ __forceinline__ float bf16lo_f(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi_f(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
// hi / lo bf16 planes of a pair of fp32 values (same rounding as split_bf16)
__device__ __forceinline__ void split2_bf16(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(x0, x1);
    lo = pack_bf16x2(x0 - bf16lo_f(hi), x1 - bf16hi_f(hi));
}
// D[16x8] = A[16x16] * B[16x8] + C, bf16 operands, fp32 accumulate (legacy warp-level tensor path: HMMA.16816).
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1,
                                               const float (&c)[4]) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}
// softplus_1mufu (mtn_scan.cu) on two values at once: the same operations per element, Horner steps as FFMA2
__device__ __forceinline__ float2 softplus2_1mufu(float x0, float x1) {
    const float2 e = make_float2(ex2_approx(-1.4426950408889634f * fabsf(x0)), ex2_approx(-1.4426950408889634f * fabsf(x1)));
    float2 q = make_float2(0.0051261021414032125f, 0.0051261021414032125f);
    q = __ffma2_rn(q, e, make_float2(-0.02907406467853027f, -0.02907406467853027f));
    q = __ffma2_rn(q, e, make_float2(0.07751608674076167f, 0.07751608674076167f));
    q = __ffma2_rn(q, e, make_float2(-0.13602247622393474f, -0.13602247622393474f));
    q = __ffma2_rn(q, e, make_float2(0.19076880735651539f, 0.19076880735651539f));
    q = __ffma2_rn(q, e, make_float2(-0.24835398988480129f, -0.24835398988480129f));
    q = __ffma2_rn(q, e, make_float2(0.3331812170752912f, 0.3331812170752912f));
    q = __ffma2_rn(q, e, make_float2(-0.49999444976340335f, -0.49999444976340335f));
    q = __ffma2_rn(q, e, make_float2(0.9999999659255092f, 0.9999999659255092f));
    return __ffma2_rn(q, e, make_float2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
}

// n-way bf16 split of a pair of fp32 values: x = p[0] + p[1] (+ p[2]) with 8 mantissa bits per plane (3 planes
// represent every fp32 value exactly up to its last bit or two)
template <int NS>
__device__ __forceinline__ void splitn_bf16(float x0, float x1, uint32_t (&pl)[NS]) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        pl[s] = pack_bf16x2(x0, x1);
        if (s + 1 < NS) {
            x0 -= bf16lo_f(pl[s]);
            x1 -= bf16hi_f(pl[s]);
        }
    }
}

// NS = planes of the dt_proj operands: 3 in fp32 mode (products with plane indices i + j <= 2: six MMA passes,
// error ~2^-24 like the FMA form), 2 in bf16 mode (three passes, ~2^-16; the reference's autocast path rounds this GEMM's
// operands AND its result to bf16, selective_scan_interface.py:174-176,187).
// ABL (dev builds, timing only, WRONG results): bit 0 skips the MMA + softplus, bit 1 the gate / store.
// softplus with TWO MUFU ops and no polynomial: max(x,0) + ln2 * lg2(1 + exp(-|x|)).  1 + e rounds e to 2^-24 absolute,
// so tiny deltas lose RELATIVE accuracy (6e-8 absolute on delta) -- irrelevant once the activations around the scan are
// bf16 (bf16 mode), where it saves 64 of the helper's FMA-pipe instructions per tile; fp32 mode keeps softplus2_1mufu.
__device__ __forceinline__ float2 softplus2_2mufu(float x0, float x1) {
    const float e0 = ex2_approx(-1.4426950408889634f * fabsf(x0)), e1 = ex2_approx(-1.4426950408889634f * fabsf(x1));
    const float2 l = make_float2(lg2_approx(1.0f + e0), lg2_approx(1.0f + e1));
    return __ffma2_rn(l, make_float2(0.6931471805599453f, 0.6931471805599453f), make_float2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
}

template <int P, int R, int NDBL, typename ZT, bool WY, int ABL = 0, bool DUO = false>
__device__ __forceinline__ void scan_pair_helper(uint8_t* ring, float* slots, uint8_t* zbuf_all, uint64_t* full_bar,
                                                  uint64_t* empty_bar, uint64_t* prepped, uint64_t* ydone,
                                                  const ScanParams& p, int w, int lane, int ch0, int b, int dir,
                                                  int ntiles, int Lb, const CUtensorMap* mapU, const CUtensorMap* mapD) {
    using SM = ScanSmemPair<P, NDBL, DUO>;
    constexpr int S = SM::STAGES, NSLOT = SM::SLOTS, NB = SM::NB, DS = P2_DS;
    constexpr int PLANE = SC_TT * DS, PAIRF = SM::PLANES * PLANE, UPL = (SM::PLANES - 1) * PLANE;
    constexpr int KS = (R + 15) / 16;                            // k-steps of the dt_proj MMA (K = 16 each)
    constexpr int NS = P == 2 ? 3 : 2;
    constexpr bool ZF = sizeof(ZT) == 4;
    constexpr int ZCH = ZF ? 2 : 1;                              // 16-byte chunks of silu(z) per (row, 8 channels)
    const int L = p.L;
    const int g = lane >> 2, tig = lane & 3;                     // mma fragment coordinates; also (row, segment) of the
    const int wch0 = ch0 + w * 32;                               // row-segment layout: rows g / g + 8, channels [8 tig, +8)
    const size_t pdw = size_t(dir) * p.di + wch0;

    // ---- dt_proj operands that never change: W_dt fragments (hi | lo) and the bias accumulator fragment
    uint32_t bw[KS][4][2][NS];
    float cb[4][2];
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
        const int chn = (g >> 1) + 4 * (2 * nb + (g & 1));       // channel of this lane's B-fragment column (n = g)
        const float* wrow = p.w_dt + (pdw + chn) * R;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int k = ks * 16 + 2 * tig + 8 * j;
                const float v0 = k < R ? wrow[k] : 0.f, v1 = k + 1 < R ? wrow[k + 1] : 0.f;
                splitn_bf16<NS>(v0, v1, bw[ks][nb][j]);
            }
        }
        cb[nb][0] = p.dt_bias[pdw + tig + 8 * nb];               // accumulator columns 2 tig, 2 tig + 1 of n-block nb
        cb[nb][1] = p.dt_bias[pdw + tig + 8 * nb + 4];
    }
    float2 sdl_sum[4];
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) sdl_sum[nb] = make_float2(0.f, 0.f);

    const size_t M = size_t(p.batch) * L;
    const size_t y_plane = M * 2 * p.di;
    const int ldy = 2 * p.di;                                     // y row stride, bf16 elements
    __nv_bfloat16* ybase = p.y + size_t(dir) * p.di + (wch0 + tig * 8);
    uint4* zbuf = reinterpret_cast<uint4*>(zbuf_all + w * (SC_TT * 32 * 4)) + lane;   // chunk j at zbuf[32 * j]
    const ZT* zsrc = reinterpret_cast<const ZT*>(p.z) + p.z_col0 + wch0 + tig * 8;
    auto tile_of = [&](int i) { return dir ? (ntiles - 1 - i) : i; };
    auto issue_tile = [&](int i2) {
        const int stg = i2 % S;
        const int row0 = b * L + tile_of(i2) * SC_TT;
        mbar_arrive_expect_tx(&full_bar[stg], SM::STAGE_BYTES);
        uint8_t* dst = ring + stg * SM::STAGE_BYTES;
        tma_load_3d(dst, mapU, &full_bar[stg], dir * p.di + ch0, row0, 0);
        tma_load_2d(dst + SM::U_BYTES, mapD, &full_bar[stg], dir * p.n_dbl, row0);
    };
    if (w == 0) {
        if (lane == 0) {
#pragma unroll 1
            for (int t = 0; t < S; ++t)
                if (t < ntiles) issue_tile(t);
        }
        __syncwarp();
    }

#pragma unroll 1
    for (int k = 0; k < ntiles + 2; ++k) {
        // ---- gate + store the tile the recurrence warp finished two hand-offs ago
        const int pt = k - 2;
        if (pt >= 0) {
            const int sl = pt % NSLOT;
            mbar_wait_long_sleep(&ydone[w * NSLOT + sl], uint32_t(pt / NSLOT) & 1u, (ABL & 32) ? 64u : 500u);
            if (WY && !(ABL & 2)) {
                const int t0 = tile_of(pt) * SC_TT;
                const int nvalid = min(SC_TT, Lb - t0);
                const float* sy = slots + sl * (4 * PAIRF) + w * PAIRF + g * DS + tig * 8;
                __nv_bfloat16* ytile = ybase + (size_t(b) * L + t0 + g) * size_t(ldy);
                cp_async_wait_all();   // this lane's own silu(z) chunks, requested one iteration ago
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    if (g + 8 * hh < nvalid) {
                        float4 ya = *reinterpret_cast<const float4*>(sy + hh * 8 * DS);
                        float4 yb = *reinterpret_cast<const float4*>(sy + hh * 8 * DS + 4);
                        if (DUO) {   // y = (states 0-7 + D * u) + (states 8-15), the two recurrence warps' partial sums
                            const float4 yc = *reinterpret_cast<const float4*>(sy + PLANE + hh * 8 * DS);
                            const float4 yd = *reinterpret_cast<const float4*>(sy + PLANE + hh * 8 * DS + 4);
                            const float2 s0 = __fadd2_rn(make_float2(ya.x, ya.y), make_float2(yc.x, yc.y));
                            const float2 s1 = __fadd2_rn(make_float2(ya.z, ya.w), make_float2(yc.z, yc.w));
                            const float2 s2 = __fadd2_rn(make_float2(yb.x, yb.y), make_float2(yd.x, yd.y));
                            const float2 s3 = __fadd2_rn(make_float2(yb.z, yb.w), make_float2(yd.z, yd.w));
                            ya = make_float4(s0.x, s0.y, s1.x, s1.y);
                            yb = make_float4(s2.x, s2.y, s3.x, s3.y);
                        }
                        float zf[8];
                        if (ZF) {
                            const uint4 za = zbuf[32 * (2 * hh)], zb = zbuf[32 * (2 * hh + 1)];
                            zf[0] = __uint_as_float(za.x); zf[1] = __uint_as_float(za.y);
                            zf[2] = __uint_as_float(za.z); zf[3] = __uint_as_float(za.w);
                            zf[4] = __uint_as_float(zb.x); zf[5] = __uint_as_float(zb.y);
                            zf[6] = __uint_as_float(zb.z); zf[7] = __uint_as_float(zb.w);
                        } else {
                            const uint4 za = zbuf[32 * hh];
                            zf[0] = bf16lo_f(za.x); zf[1] = bf16hi_f(za.x); zf[2] = bf16lo_f(za.y); zf[3] = bf16hi_f(za.y);
                            zf[4] = bf16lo_f(za.z); zf[5] = bf16hi_f(za.z); zf[6] = bf16lo_f(za.w); zf[7] = bf16hi_f(za.w);
                        }
                        const float yv[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            // y * (0.5 * silu(z)): two roundings (0.5 * z is exact)
                            const float2 gz = __fmul2_rn(make_float2(zf[2 * q], zf[2 * q + 1]), make_float2(0.5f, 0.5f));
                            const float2 o = __fmul2_rn(make_float2(yv[2 * q], yv[2 * q + 1]), gz);
                            if (P == 2) split2_bf16(o.x, o.y, hi[q], lo[q]);
                            else hi[q] = pack_bf16x2(o.x, o.y);
                        }
                        __nv_bfloat16* yrow = ytile + size_t(hh * 8) * size_t(ldy);
                        *reinterpret_cast<uint4*>(yrow) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        if (P == 2) *reinterpret_cast<uint4*>(yrow + y_plane) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
            }
        }
        // ---- producer duty of this iteration (rotates over the four helpers): refill the ring stage of tile k-2
        if ((k & 3) == w) {
            const bool refill = pt >= 0 && pt + S < ntiles;
            if (refill) {
                mbar_wait_sleep(&empty_bar[pt % S], uint32_t(pt / S) & 1u);   // every recurrence warp is done with it
                if (lane == 0) issue_tile(pt + S);
            }
            __syncwarp();
        }
        // ---- request silu(z) of the tile gated in the NEXT iteration: 16-byte chunks, each lane its own elements
        if (WY && !(ABL & 2) && k >= 1 && k - 1 < ntiles) {
            const int t0 = tile_of(k - 1) * SC_TT;
            const int nvalid = min(SC_TT, Lb - t0);
            const ZT* zs = zsrc + (size_t(b) * L + t0 + g) * size_t(p.ldz);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                if (g + 8 * hh < nvalid) {
#pragma unroll
                    for (int c = 0; c < ZCH; ++c)
                        cp_async_16(zbuf + 32 * (ZCH * hh + c), zs + size_t(hh * 8) * size_t(p.ldz) + c * 4);
                }
            }
        }
        cp_async_commit();
        // ---- delta and u of tile k
        if (k < ntiles) {
            const int stg = k % S, sl = k % NSLOT;
            const int t0 = tile_of(k) * SC_TT;
            const int nvalid = min(SC_TT, Lb - t0);
            mbar_wait_sleep(&full_bar[stg], uint32_t(k / S) & 1u);
            const uint8_t* st = ring + stg * SM::STAGE_BYTES;
            float* slot = slots + sl * (4 * PAIRF) + w * PAIRF;
            // dt operand fragments: rows g, g + 8 of the staged [dt | B | C] rows, split into bf16 hi | lo
            const float* sd = reinterpret_cast<const float*>(st + SM::U_BYTES) + g * NB + 2 * tig;
            float acc[4][4];
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                acc[nb][0] = acc[nb][2] = cb[nb][0];
                acc[nb][1] = acc[nb][3] = cb[nb][1];
            }
            if (!(ABL & 1)) {
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    uint32_t aw[NS][4];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const float2 v0 = *reinterpret_cast<const float2*>(sd + ks * 16 + 8 * j);
                        const float2 v1 = *reinterpret_cast<const float2*>(sd + 8 * NB + ks * 16 + 8 * j);
                        uint32_t p0[NS], p1[NS];
                        splitn_bf16<NS>(v0.x, v0.y, p0);
                        splitn_bf16<NS>(v1.x, v1.y, p1);
#pragma unroll
                        for (int s2 = 0; s2 < NS; ++s2) {
                            aw[s2][2 * j] = p0[s2];
                            aw[s2][2 * j + 1] = p1[s2];
                        }
                    }
                    // smallest products first (plane index sum NS-1 down to 0); the four n-blocks are independent chains
#pragma unroll
                    for (int tsum = NS - 1; tsum >= 0; --tsum) {
#pragma unroll
                        for (int i = 0; i <= tsum; ++i) {
#pragma unroll
                            for (int nb = 0; nb < 4; ++nb)
                                mma_bf16_16816(acc[nb], aw[i], bw[ks][nb][0][tsum - i], bw[ks][nb][1][tsum - i], acc[nb]);
                        }
                    }
                }
            }
            float* sdl = slot + g * DS + tig;
            const bool v0 = g < nvalid, v1 = g + 8 < nvalid;   // rows past the utterance end: delta = 0 -> state unchanged
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                constexpr bool SP2 = (P == 1) != bool(ABL & 64);   // ABL bit 6 (dev builds): the other form, for A/B timing
                float2 d0 = (ABL & 1) ? make_float2(acc[nb][0], acc[nb][1])
                            : SP2     ? softplus2_2mufu(acc[nb][0], acc[nb][1]) : softplus2_1mufu(acc[nb][0], acc[nb][1]);
                float2 d1 = (ABL & 1) ? make_float2(acc[nb][2], acc[nb][3])
                            : SP2     ? softplus2_2mufu(acc[nb][2], acc[nb][3]) : softplus2_1mufu(acc[nb][2], acc[nb][3]);
                if (nvalid < SC_TT) {
                    if (!v0) d0 = make_float2(0.f, 0.f);
                    if (!v1) d1 = make_float2(0.f, 0.f);
                }
                sdl_sum[nb] = __fadd2_rn(sdl_sum[nb], __fadd2_rn(d0, d1));
                sdl[8 * nb] = d0.x;
                sdl[8 * nb + 4] = d0.y;
                sdl[8 * DS + 8 * nb] = d1.x;
                sdl[8 * DS + 8 * nb + 4] = d1.y;
                if (DUO) {   // the second recurrence warp's own copy
                    sdl[PLANE + 8 * nb] = d0.x;
                    sdl[PLANE + 8 * nb + 4] = d0.y;
                    sdl[PLANE + 8 * DS + 8 * nb] = d1.x;
                    sdl[PLANE + 8 * DS + 8 * nb + 4] = d1.y;
                }
            }
            // u planes -> fp32, rows g / g + 8, channels [8 tig, 8 tig + 8) of this warp's 32
            const __nv_bfloat16* su = reinterpret_cast<const __nv_bfloat16*>(st) + g * SC_CH + w * 32 + tig * 8;
            float* suo = slot + UPL + g * DS + tig * 8;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const uint4 uh = *reinterpret_cast<const uint4*>(su + hh * 8 * SC_CH);
                const uint32_t hw[4] = {uh.x, uh.y, uh.z, uh.w};
                float2 f[4];
                if (P == 2) {
                    const uint4 ul = *reinterpret_cast<const uint4*>(su + (SC_TT + hh * 8) * SC_CH);
                    const uint32_t lw[4] = {ul.x, ul.y, ul.z, ul.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        f[q] = __fadd2_rn(make_float2(bf16lo_f(hw[q]), bf16hi_f(hw[q])),
                                          make_float2(bf16lo_f(lw[q]), bf16hi_f(lw[q])));
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) f[q] = make_float2(bf16lo_f(hw[q]), bf16hi_f(hw[q]));
                }
                *reinterpret_cast<float4*>(suo + hh * 8 * DS) = make_float4(f[0].x, f[0].y, f[1].x, f[1].y);
                *reinterpret_cast<float4*>(suo + hh * 8 * DS + 4) = make_float4(f[2].x, f[2].y, f[3].x, f[3].y);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&prepped[w * NSLOT + sl]);
        }
    }
    if (p.sum_delta) {   // per channel: sum over the rows held by the 8 lanes that share tig
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                sdl_sum[nb].x += __shfl_xor_sync(0xffffffffu, sdl_sum[nb].x, o);
                sdl_sum[nb].y += __shfl_xor_sync(0xffffffffu, sdl_sum[nb].y, o);
            }
        }
        if (g == 0) {
            float* dst = p.sum_delta + (size_t(dir) * p.batch + b) * p.di + wch0 + tig;
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                dst[8 * nb] = sdl_sum[nb].x;
                dst[8 * nb + 4] = sdl_sum[nb].y;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ "duo" mapping
// Two recurrence warps per 32 channels, 8 states each (HS = 0: states 0-7 and the D * u skip term, HS = 1: states 8-15), and
// the same helper warp: 12 warps per CTA, warps w / w + 4 / w + 8 of a channel group on one SM sub-partition.
// Why: with one recurrence warp per group a sub-partition holds at most two of them (registers), and two in-order warps
// cannot keep the XU fed -- config 2 ran 392 cycles per step of a sub-partition's two groups against 256 cycles of MUFU work
// (profiles/r02/scan_v7_pair2_ncu_summary.txt: XU 70 %, issue 55 %, 27 % of the recurrence warps' samples in fixed-latency
// `wait`; without any MUFU the recurrence still took 0.59 ms).  Halving the state per warp doubles the independent
// instruction streams at the same total MUFU / FMA work; it costs the duplicated per-step scalars (delta, u, delta * u,
// pointer steps: ~8 issue slots per step) and one extra add per element in the helper's gate.  No shuffles, no atomics:
// each half owns a delta copy in the slot and overwrites it in place with its partial y.
template <int HS, int S, int NB, int BOFF, int ABL, int DS>
__device__ __forceinline__ void scan_duo_recur(const uint8_t* ring, float* slots, int stage_bytes, int u_bytes,
                                               uint64_t* full_bar, uint64_t* empty_bar, uint64_t* prepped,
                                               uint64_t* ydone, const ScanParams& p, int w, int lane, int d, int b,
                                               int dir, int ntiles) {
    constexpr int NSLOT = 2;
    constexpr int PLANE = SC_TT * DS, PAIRF = 3 * PLANE;
    constexpr int UOFF = (2 - HS) * PLANE;      // u plane relative to this half's delta -> y plane
    constexpr int SOFF = 8 * HS;                // first state of this half
    const size_t pd = size_t(dir) * p.di + d;
    float2 h2[4], A2[4];
    {
        const float4* ap = reinterpret_cast<const float4*>(p.A2 + pd * SC_NS + SOFF);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float4 a = ap[q];
            A2[2 * q] = make_float2(a.x, a.y);
            A2[2 * q + 1] = make_float2(a.z, a.w);
        }
        if (p.h_in) {
            const float4* hp = reinterpret_cast<const float4*>(p.h_in + ((size_t(dir) * p.batch + b) * p.di + d) * SC_NS + SOFF);
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
    const float Dv = HS == 0 ? p.Dskip[pd] : 0.f;
    const int bc_first = dir ? (SC_TT - 1) * NB : 0;   // floats
    const int bc_step = dir ? -NB : NB;
    const int sl_first = dir ? (SC_TT - 1) * DS : 0;
    const int sl_step = dir ? -DS : DS;
    auto decay = [&](float dl, int q) -> float2 {
        const float2 a = __fmul2_rn(make_float2(dl, dl), A2[q]);
        return (ABL & 4) ? __ffma2_rn(a, make_float2(0.5f, 0.5f), make_float2(1.f, 1.f))
                         : make_float2(ex2_approx(a.x), ex2_approx(a.y));
    };
    for (int i = 0; i < ntiles; ++i) {
        const int stg = i % S, sl = i % NSLOT;
        mbar_wait(&prepped[w * NSLOT + sl], uint32_t(i / NSLOT) & 1u);
        mbar_wait(&full_bar[stg], uint32_t(i / S) & 1u);  // complete long ago; orders this warp after the TMA writes
        float* psl = slots + sl * (4 * PAIRF) + w * PAIRF + HS * PLANE + lane + sl_first;
        const float* pbc = reinterpret_cast<const float*>(ring + stg * stage_bytes + u_bytes) + BOFF + SOFF + bc_first;
        // decay factors two steps ahead in their own registers (see scan_pair_recur)
        float2 e_c[4], e_n[4];
        float4 Bq[2];
        float dl_c = psl[0], u_c = psl[UOFF];
        float dl_n = psl[sl_step], u_n = psl[sl_step + UOFF];
#pragma unroll
        for (int q = 0; q < 4; ++q) e_c[q] = decay(dl_c, q);
#pragma unroll
        for (int q = 0; q < 4; ++q) e_n[q] = decay(dl_n, q);
#pragma unroll
        for (int q = 0; q < 2; ++q) Bq[q] = *reinterpret_cast<const float4*>(pbc + 4 * q);
#pragma unroll
        for (int jj = 0; jj < SC_TT; ++jj) {
            float4 Cq[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) Cq[q] = *reinterpret_cast<const float4*>(pbc + SC_NS + 4 * q);
            pbc += bc_step;
            float* py = psl;
            psl += sl_step;
            float dl_nn = 0.f, u_nn = 0.f;
            if (jj + 2 < SC_TT) {
                dl_nn = psl[sl_step];
                u_nn = psl[sl_step + UOFF];
            }
            const float du = dl_c * u_c;
            const float2 du2 = make_float2(du, du);
            float2 bu[4];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                bu[2 * q] = __fmul2_rn(du2, make_float2(Bq[q].x, Bq[q].y));
                bu[2 * q + 1] = __fmul2_rn(du2, make_float2(Bq[q].z, Bq[q].w));
            }
            if (jj + 1 < SC_TT) {
#pragma unroll
                for (int q = 0; q < 2; ++q) Bq[q] = *reinterpret_cast<const float4*>(pbc + 4 * q);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                h2[q] = __ffma2_rn(e_c[q], h2[q], bu[q]);
                e_c[q] = e_n[q];
                if (jj + 2 < SC_TT) e_n[q] = decay(dl_nn, q);
            }
            float2 y0 = make_float2(HS == 0 ? Dv * u_c : 0.f, 0.f), y1 = make_float2(0.f, 0.f);
            y0 = __ffma2_rn(h2[0], make_float2(Cq[0].x, Cq[0].y), y0);
            y1 = __ffma2_rn(h2[1], make_float2(Cq[0].z, Cq[0].w), y1);
            y0 = __ffma2_rn(h2[2], make_float2(Cq[1].x, Cq[1].y), y0);
            y1 = __ffma2_rn(h2[3], make_float2(Cq[1].z, Cq[1].w), y1);
            const float2 sy = __fadd2_rn(y0, y1);
            sts_f32_nofence(py, sy.x + sy.y);
            dl_c = dl_n;
            u_c = u_n;
            dl_n = dl_nn;
            u_n = u_nn;
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&ydone[w * NSLOT + sl]);
            mbar_arrive(&empty_bar[stg]);
        }
    }
    if (p.h_out) {
        float4* hp = reinterpret_cast<float4*>(p.h_out + ((size_t(dir) * p.batch + b) * p.di + d) * SC_NS + SOFF);
#pragma unroll
        for (int q = 0; q < 2; ++q) hp[q] = make_float4(h2[2 * q].x, h2[2 * q].y, h2[2 * q + 1].x, h2[2 * q + 1].y);
    }
}

template <int P, int R, int NDBL, typename ZT, int ABL>
__global__ void __launch_bounds__(384, 2)
scan_kernel_duo(const __grid_constant__ CUtensorMap mapU, const __grid_constant__ CUtensorMap mapD, const ScanParams p) {
    using SM = ScanSmemPair<P, NDBL, true>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint8_t* ring = smem;
    float* slots = reinterpret_cast<float*>(smem + SM::STAGES * SM::STAGE_BYTES);
    uint8_t* zbuf = smem + SM::STAGES * SM::STAGE_BYTES + SM::SLOTS * SM::SLOT_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(zbuf + SM::ZBUF_BYTES);
    uint64_t* empty_bar = full_bar + 4;
    uint64_t* prepped = empty_bar + 4;
    uint64_t* ydone = prepped + 4 * SM::SLOTS;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int nchb = p.di / SC_CH;
    const int ch0 = (blockIdx.x % nchb) * SC_CH;
    const int dir = p.dir0 + blockIdx.x / nchb;
    const int b = blockIdx.y;
    if (tid == 0) {
        tma_prefetch_desc(&mapU);
        tma_prefetch_desc(&mapD);
        for (int s = 0; s < SM::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 8);     // both halves of the four channel groups
        }
        for (int s = 0; s < 4 * SM::SLOTS; ++s) {
            mbar_init(&prepped[s], 1);
            mbar_init(&ydone[s], 2);         // both halves
        }
        fence_barrier_init();
    }
    __syncthreads();
    const int Lb = (b == p.batch - 1) ? p.L_last : p.L;
    const int ntiles = (Lb + SC_TT - 1) / SC_TT;
    const int w = warp & 3;
    const int role = (ABL & 128) ? (warp < 4 ? 2 : (warp < 8 ? 0 : 1)) : (warp >> 2);   // bit 7: helper = the LOWEST warp ids
    if (role == 0)
        scan_duo_recur<0, SM::STAGES, SM::NB, R, ABL, P2_DS>(ring, slots, SM::STAGE_BYTES, SM::U_BYTES, full_bar, empty_bar,
                                                            prepped, ydone, p, w, lane, ch0 + w * 32 + lane, b, dir, ntiles);
    else if (role == 1)
        scan_duo_recur<1, SM::STAGES, SM::NB, R, ABL, P2_DS>(ring, slots, SM::STAGE_BYTES, SM::U_BYTES, full_bar, empty_bar,
                                                            prepped, ydone, p, w, lane, ch0 + w * 32 + lane, b, dir, ntiles);
    else
        scan_pair_helper<P, R, NDBL, ZT, true, ABL, true>(ring, slots, zbuf, full_bar, empty_bar, prepped, ydone, p, w, lane,
                                                          ch0, b, dir, ntiles, Lb, &mapU, &mapD);
}

template <int P, int R, int NDBL, typename ZT, int ABL = 0>
static int launch_scan_duo(const mtn_scan_args* a, cudaStream_t stream) {
    using SM = ScanSmemPair<P, NDBL, true>;
    MTN_REQUIRE((reinterpret_cast<uintptr_t>(a->z) & 15) == 0 && (a->ldz * sizeof(ZT)) % 16 == 0 &&
                    (a->z_col0 * sizeof(ZT)) % 16 == 0,
                "scan(duo): the gate block must be 16-byte aligned (z=%p ldz=%d z_col0=%d)", a->z, a->ldz, a->z_col0);
    MTN_REQUIRE(a->y && (reinterpret_cast<uintptr_t>(a->y) & 15) == 0, "scan(duo): y must be 16-byte aligned");
    const uint64_t M = uint64_t(a->batch) * a->L;
    CUtensorMap mapU, mapD;
    {
        uint64_t dims[3] = {uint64_t(2) * a->di, M, uint64_t(P)};
        uint64_t str[2] = {uint64_t(2) * a->di * 2, M * 2 * a->di * 2};
        uint32_t box[3] = {uint32_t(SC_CH), SC_TT, uint32_t(P)};
        if (!encode_tmap(&mapU, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a->u, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE))
            return MTN_ECUDA;
    }
    {
        uint64_t dims[2] = {uint64_t(a->ld_dbl), M};
        uint64_t str[1] = {uint64_t(a->ld_dbl) * 4};
        uint32_t box[2] = {uint32_t(SM::NB), SC_TT};
        if (!encode_tmap(&mapD, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a->dbl, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE))
            return MTN_ECUDA;
    }
    ScanParams p = make_scan_params(a);
    auto kern = scan_kernel_duo<P, R, NDBL, ZT, ABL>;
    static std::atomic<unsigned long long> attr_done{0};   // per template instantiation, one bit per device
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), SM::TOTAL, attr_done, "scan(duo)")) return rc;
    dim3 grid(p.ndirs * (a->di / SC_CH), a->batch, 1);
    kern<<<grid, 384, SM::TOTAL, stream>>>(mapU, mapD, p);
    MTN_CUDA_LAUNCH_CHECK("scan(duo)");
    return MTN_OK;
}

template <int P, int R, int NDBL, typename ZT, bool WY, bool RG, int KP, int ABL>
__global__ void __launch_bounds__(256, 2)
scan_kernel_pair(const __grid_constant__ CUtensorMap mapU, const __grid_constant__ CUtensorMap mapD, const ScanParams p) {
    using SM = ScanSmemPair<P, NDBL>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint8_t* ring = smem;
    float* slots = reinterpret_cast<float*>(smem + SM::STAGES * SM::STAGE_BYTES);
    uint8_t* zbuf = smem + SM::STAGES * SM::STAGE_BYTES + SM::SLOTS * SM::SLOT_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(zbuf + SM::ZBUF_BYTES);
    uint64_t* empty_bar = full_bar + 4;
    uint64_t* prepped = empty_bar + 4;
    uint64_t* ydone = prepped + 4 * SM::SLOTS;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int nchb = p.di / SC_CH;
    const int ch0 = (blockIdx.x % nchb) * SC_CH;
    const int dir = p.dir0 + blockIdx.x / nchb;
    const int b = blockIdx.y;
    if (tid == 0) {
        tma_prefetch_desc(&mapU);
        tma_prefetch_desc(&mapD);
        for (int s = 0; s < SM::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 4);
        }
        for (int s = 0; s < 4 * SM::SLOTS; ++s) {
            mbar_init(&prepped[s], 1);
            mbar_init(&ydone[s], 1);
        }
        fence_barrier_init();
    }
    __syncthreads();
    const int Lb = (b == p.batch - 1) ? p.L_last : p.L;
    const int ntiles = (Lb + SC_TT - 1) / SC_TT;
    const int w = warp & 3;
    if ((ABL & 128) ? warp >= 4 : warp < 4)   // ABL bit 7 (dev builds): roles swapped, recurrence = the HIGHER warp ids
        scan_pair_recur<SM::STAGES, SM::NB, R, WY, ABL, RG, P2_DS, KP>(ring, slots, SM::STAGE_BYTES, SM::U_BYTES, full_bar,
                                                                    empty_bar, prepped, ydone, p, w, lane,
                                                                    ch0 + w * 32 + lane, b, dir, ntiles, Lb);
    else
        scan_pair_helper<P, R, NDBL, ZT, WY, ABL>(ring, slots, zbuf, full_bar, empty_bar, prepped, ydone, p, w, lane, ch0, b,
                                              dir, ntiles, Lb, &mapU, &mapD);
}

template <int P, int R, int NDBL, typename ZT, bool WY, int KP = 0, bool RG = false, int ABL = 0>
static int launch_scan_pair(const mtn_scan_args* a, cudaStream_t stream) {
    // short ragged sequences (many of them: DPMamba's inter model) take the instantiation with the ragged-tile path
    if (!RG && KP == 0 && ABL == 0 && WY && a->L <= 128 && (a->L % SC_TT) != 0 && a->L_last == 0)
        return launch_scan_pair<P, R, NDBL, ZT, WY, KP, true>(a, stream);
    using SM = ScanSmemPair<P, NDBL>;
    MTN_REQUIRE((reinterpret_cast<uintptr_t>(a->z) & 15) == 0 && (a->ldz * sizeof(ZT)) % 16 == 0 &&
                    (a->z_col0 * sizeof(ZT)) % 16 == 0,
                "scan(pair): the gate block must be 16-byte aligned (z=%p ldz=%d z_col0=%d)", a->z, a->ldz, a->z_col0);
    MTN_REQUIRE(!a->y || (reinterpret_cast<uintptr_t>(a->y) & 15) == 0, "scan(pair): y must be 16-byte aligned");
    const uint64_t M = uint64_t(a->batch) * a->L;
    CUtensorMap mapU, mapD;
    {
        uint64_t dims[3] = {uint64_t(2) * a->di, M, uint64_t(P)};
        uint64_t str[2] = {uint64_t(2) * a->di * 2, M * 2 * a->di * 2};
        uint32_t box[3] = {uint32_t(SC_CH), SC_TT, uint32_t(P)};
        if (!encode_tmap(&mapU, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a->u, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE))
            return MTN_ECUDA;
    }
    {
        uint64_t dims[2] = {uint64_t(a->ld_dbl), M};
        uint64_t str[1] = {uint64_t(a->ld_dbl) * 4};
        uint32_t box[2] = {uint32_t(SM::NB), SC_TT};
        if (!encode_tmap(&mapD, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a->dbl, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE))
            return MTN_ECUDA;
    }
    ScanParams p = make_scan_params(a);
    auto kern = scan_kernel_pair<P, R, NDBL, ZT, WY, RG, KP, ABL>;
    static std::atomic<unsigned long long> attr_done{0};   // per template instantiation, one bit per device
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), SM::TOTAL, attr_done, "scan(pair)")) return rc;
    dim3 grid(p.ndirs * (a->di / SC_CH), a->batch, 1);
    kern<<<grid, 256, SM::TOTAL, stream>>>(mapU, mapD, p);
    MTN_CUDA_LAUNCH_CHECK("scan(pair)");
    return MTN_OK;
}

}  // namespace mtn
