// One-launch streaming push of the causal Mamba-TasNet separator (sm_100a).
//
// A streaming chunk is a handful of frames (a 20 ms push at 8 kHz = 20 frames): every GEMM of the path has M <= 32 rows and
// the ~100-kernel chain of the batch plan is pure launch latency there (0.56 ms per push, profiles/r01/stream_push_*).
// This kernel runs the WHOLE push -- encoder + cLN + bottleneck, every Mamba block (Add -> RMSNorm -> in_proj -> causal
// conv + SiLU -> x_proj -> dt_proj + softplus -> selective-scan steps -> gate -> out_proj), norm_f, mask, decoder
// overlap-add -- as ONE launch: one thread-block CLUSTER per stream, the channels of d_inner cut over the CTAs of the
// cluster (DSL = 32, 64 or 128 per CTA, a template parameter: clusters of 16 / 8 / 4 CTAs for the S recipe, picked by the
// caller from the number of streams), weights streamed from L2 exactly once per CTA.
//
// Replaces, for chunks of <= 32 frames, the reference's token-at-a-time `Mamba.step` + `inference_params` caches
// (Mamba-TasNet/modules/mamba/bimamba.py:320-372, caches :374-404) driven through `MambaBlocksSequential.forward(x,
// inference_params)` (modules/mamba_blocks.py:186-197), plus the Encoder / MaskNet / Decoder calls around it
// (train_wsj0mix.py:86-111).
//
// Arithmetic: the contractions run on warp-level mma.sync.m16n8k16 (bf16 hi/lo split of both operands, three passes,
// fp32 accumulate = the same fp32-class operand model as the batch plan's tcgen05 GEMMs).  The WEIGHT is the M-side
// operand (16 output channels per tile, fragments pre-packed per lane at load time so a warp reads 512 contiguous bytes per
// k-step straight from L2 into registers) and the FRAMES are the N side (8 per tile: F = 20 costs 24 rows, not 32).
// Everything else (norms, conv, dt_proj, the recurrence, overlap-add) is fp32 SIMT on shared memory.
//
// Cluster dataflow per block (rank r owns d_inner channels [DSL r, DSL r + DSL) and d_model columns [DSL/2 r, DSL/2 r + DSL/2)):
//   [A] every CTA gathers the residual slices of all ranks -> RMSNorm of all rows -> in_proj (its x / z columns) -> conv +
//   SiLU -> x_proj partial over its channels [B] sum of the partials of all ranks -> dt_proj + softplus -> scan (its
//   channels, 16 states, state in / out of the cache) -> gate -> out_proj partial over its channels [C] sum of its
//   columns over all ranks, residual slice += that -> next block.
// The three exchanges per block are reads of the peers' shared memory (DSMEM, ld.shared::cluster) behind a cluster barrier:
// a first version that exchanged through global memory spent 2.2 us per dependent read of a line another SM had just
// written (tools/stream_push_timeline.py), ten times the DSMEM latency.  Every exchanged buffer is rewritten only after a
// later cluster barrier, so nothing is double-buffered.
#include "mtn_ptx.cuh"
#include "mtn_host.h"

namespace mtn {

constexpr int SP_THREADS = 512;
constexpr int SP_WARPS = SP_THREADS / 32;
// Channels of d_inner per CTA (DSL, a template parameter: 32, 64 or 128) fix the partition: cluster size CL = 2 * d_model / DSL,
// d_model columns per CTA CSL = DSL / 2, mask columns per CTA MSL = DSL (n_spk * N / CL with N == d_model, 2 speakers).  Small
// DSL = more CTAs per stream = lower latency; large DSL = more streams resident at once.
constexpr int SP_PAD = 8;    // bf16 row padding of the operand planes: row stride = 4 (mod 32) words, conflict-free fragments

// Shared-memory carve-up (bytes), compile-time per (frame tiles, d_model): every buffer is `smem + constant`, which keeps ~30
// pointers out of the register file (a run-time layout spilled 1.4 KB per thread).  Identical in every CTA of a cluster:
// peers address each other's buffers by the same offsets.
template <int CT, int KSTEPS>
struct GemmShape {
    static constexpr int ks_() {
        int KS = 1;
        while (KS * 2 * CT <= SP_WARPS && KS * 2 <= KSTEPS && KSTEPS % (KS * 2) == 0) KS *= 2;
        return KS;
    }
    static constexpr int KS = ks_(), KPER = KSTEPS / KS, UNITS = CT * KS, NPRE = KPER < 8 ? KPER : 8, LDR = 16 * CT + 4;
};

constexpr int sp_al16(int bytes) { return (bytes + 15) / 16 * 16; }
constexpr int sp_max(int a, int b) { return a > b ? a : b; }
template <int NTF, int D, int DSL>
struct SpL {
    static constexpr int R = D / 16, NXp = (R + 32 + 15) / 16 * 16, CL = 2 * D / DSL, CSL = DSL / 2, MSL = DSL;
    static constexpr int Fp = 8 * NTF, lda = D + SP_PAD, ldu = DSL + SP_PAD;
    static constexpr int act_hi = 0;
    static constexpr int act_lo = act_hi + sp_al16(Fp * lda * 2);
    static constexpr int su_hi = act_lo + sp_al16(Fp * lda * 2);
    static constexpr int su_lo = su_hi + sp_al16(Fp * ldu * 2);
    static constexpr int planes_end = su_lo + sp_al16(Fp * ldu * 2);
    static constexpr int xs = planes_end;
    static constexpr int zs = xs + sp_al16((Fp + 3) * DSL * 4);
    static constexpr int us = zs + sp_al16(Fp * DSL * 4);
    static constexpr int dl = us + sp_al16(Fp * DSL * 4);
    static constexpr int dbl = dl + sp_al16(Fp * DSL * 4);
    static constexpr int xd = dbl + sp_al16(Fp * NXp * 4);        // x_proj partial of this CTA (read by every peer)
    static constexpr int res = xd + sp_al16(Fp * NXp * 4);        // residual slice (read by every peer)
    static constexpr int mixw = res + sp_al16(Fp * CSL * 4);
    static constexpr int frs = mixw + sp_al16(Fp * MSL * 4);      // decoder frames: this CTA's partial (read by rank 0)
    static constexpr int lvec = frs + sp_al16(Fp * 16 * 4);       // this CTA's slice of the layer's small vectors
    static constexpr int norm = lvec + sp_al16(DSL * (23 + R) * 4);   // RMSNorm weights, double-buffered one layer ahead
    // K-split partial sums of one GEMM (plane 0 of out_proj's is read by every peer): the widest of the five contractions
    template <int CT, int KSTEPS>
    static constexpr int width() { return GemmShape<CT, KSTEPS>::KS * GemmShape<CT, KSTEPS>::LDR; }
    static constexpr int widest = sp_max(sp_max(sp_max(width<2 * DSL / 16, D / 16>(), width<NXp / 16, DSL / 16>()),
                                                sp_max(width<D / 16, DSL / 16>(), width<CSL / 16, D / 16>())),
                                         sp_max(width<MSL / 16, D / 16>(), 2 * 16 + 4));
    static constexpr int red = norm + sp_al16(2 * D * 4);
    static constexpr int total = red + sp_al16(Fp * widest * 4);
};

__device__ __forceinline__ void mma16816(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void put_planes(__nv_bfloat16* hi, __nv_bfloat16* lo, int idx, float v) {
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[idx] = h;
    lo[idx] = l;
}

// ---- distributed shared memory: the same offset in cluster rank `rank`'s shared memory
__device__ __forceinline__ uint32_t dsmem_addr(const void* local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(rank));
    return r;
}
__device__ __forceinline__ float ld_dsmem(uint32_t addr) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_dsmem4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
// sum over the cluster ranks of the float4 at the same shared-memory address: eight loads in flight before the first add.
// Rank `rank` starts with its right-hand neighbour, so the eight CTAs never queue on the same peer's shared-memory port; the
// order of the additions therefore differs from rank to rank (and is fixed for a rank: results are reproducible).
template <int CL>
__device__ __forceinline__ float4 rank_sum4(const float* local, int rank) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r0 = 0; r0 < CL; r0 += 8) {
        constexpr int NB = CL < 8 ? CL : 8;
        float4 v[NB];
#pragma unroll
        for (int r = 0; r < NB; ++r) v[r] = ld_dsmem4(dsmem_addr(local, (rank + 1 + r0 + r) & (CL - 1)));
#pragma unroll
        for (int r = 0; r < NB; ++r) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
    }
    return s;
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

#define SP_MARK(k)                                                                                                          \
    do {                                                                                                                    \
        if (a.timeline && blockIdx.x == 0 && tid == 0) a.timeline[size_t(tl_row) * 16 + (k)] = global_timer_ns();          \
    } while (0)

// ---- CTA-wide GEMM on pre-packed weight fragments.
// CT tiles of 16 output channels x KSTEPS k-steps of 16.  With fewer tiles than warps the k range is cut KS ways (partial sums
// in red[ks][frame][ldr], ldr = 16*CT + 4); a warp works on units (tile, k part) warp, warp + 16, ...  The weights of a
// warp's FIRST unit (up to 8 k-steps = 64 registers) can be requested long before the operand planes exist (`issue`), so the
// L2 round trip and the 512 B / k-step / warp stream hide behind the phase in between.

__device__ __forceinline__ uint4 ld_weights(const uint4* p) {   // volatile: stays where it is issued
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int NB>
struct WFrag {
    uint4 h[NB], l[NB];
};

template <int NB>
__device__ __forceinline__ void wfrag_load(WFrag<NB>& w, const uint4* __restrict__ wf, int lane) {
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        w.h[i] = ld_weights(wf + (i * 2) * 32 + lane);
        w.l[i] = ld_weights(wf + (i * 2 + 1) * 32 + lane);
    }
}

template <int CT, int KSTEPS>
__device__ __forceinline__ void gemm_issue(WFrag<GemmShape<CT, KSTEPS>::NPRE>& w, const uint4* __restrict__ wfrag, int warp, int lane) {
    using G = GemmShape<CT, KSTEPS>;
    if (warp < G::UNITS) wfrag_load<G::NPRE>(w, wfrag + (size_t(warp / G::KS) * KSTEPS + size_t(warp % G::KS) * G::KPER) * 64, lane);
}

// acc += W[NB k-steps] x act rows (bh / bl point at this lane's element of k-step 0)
template <int NTF, int NB>
__device__ __forceinline__ void mma_batch(const WFrag<NB>& w, const __nv_bfloat16* bh, const __nv_bfloat16* bl, int lda,
                                          float (&acc)[NTF][4]) {
#pragma unroll
    for (int i = 0; i < NB; ++i) {
#pragma unroll
        for (int j = 0; j < NTF; ++j) {
            const uint32_t h0 = *reinterpret_cast<const uint32_t*>(bh + j * 8 * lda + i * 16);
            const uint32_t h1 = *reinterpret_cast<const uint32_t*>(bh + j * 8 * lda + i * 16 + 8);
            const uint32_t l0 = *reinterpret_cast<const uint32_t*>(bl + j * 8 * lda + i * 16);
            const uint32_t l1 = *reinterpret_cast<const uint32_t*>(bl + j * 8 * lda + i * 16 + 8);
            mma16816(acc[j], w.h[i], h0, h1);
            mma16816(acc[j], w.l[i], h0, h1);
            mma16816(acc[j], w.h[i], l0, l1);
        }
    }
}

template <int NTF, int CT, int KSTEPS>
__device__ __forceinline__ void gemm_run(const WFrag<GemmShape<CT, KSTEPS>::NPRE>& pre, const uint4* __restrict__ wfrag,
                                         const __nv_bfloat16* ahi, const __nv_bfloat16* alo, int lda, float* red, int warp, int lane) {
    using G = GemmShape<CT, KSTEPS>;
    constexpr int Fp = 8 * NTF;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
    for (int unit = warp; unit < G::UNITS; unit += SP_WARPS) {
        const int ct = unit / G::KS, ks = unit % G::KS;
        const uint4* wf = wfrag + (size_t(ct) * KSTEPS + size_t(ks) * G::KPER) * 64;
        const __nv_bfloat16* bh = ahi + g * lda + ks * G::KPER * 16 + 2 * t;
        const __nv_bfloat16* bl = alo + g * lda + ks * G::KPER * 16 + 2 * t;
        float acc[NTF][4];
#pragma unroll
        for (int j = 0; j < NTF; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        if (unit == warp) {
            mma_batch<NTF, G::NPRE>(pre, bh, bl, lda, acc);
            if constexpr (G::KPER > G::NPRE) {
                static_assert(G::KPER <= G::NPRE || (G::KPER - G::NPRE) % 4 == 0, "k-step remainder in batches of four");
#pragma unroll 1
                for (int k = G::NPRE; k < G::KPER; k += 4) {
                    WFrag<4> w;
                    wfrag_load<4>(w, wf + size_t(k) * 64, lane);
                    mma_batch<NTF, 4>(w, bh + k * 16, bl + k * 16, lda, acc);
                }
            }
        } else {
            constexpr int NB = G::KPER % 4 == 0 ? 4 : (G::KPER % 2 == 0 ? 2 : 1);
#pragma unroll 1
            for (int k = 0; k < G::KPER; k += NB) {
                WFrag<NB> w;
                wfrag_load<NB>(w, wf + size_t(k) * 64, lane);
                mma_batch<NTF, NB>(w, bh + k * 16, bl + k * 16, lda, acc);
            }
        }
        float* r = red + size_t(ks) * Fp * G::LDR + ct * 16 + g;
#pragma unroll
        for (int j = 0; j < NTF; ++j) {
            const int f0 = j * 8 + 2 * t;
            r[f0 * G::LDR] = acc[j][0];
            r[(f0 + 1) * G::LDR] = acc[j][1];
            r[f0 * G::LDR + 8] = acc[j][2];
            r[(f0 + 1) * G::LDR + 8] = acc[j][3];
        }
    }
}

__device__ __forceinline__ float red_sum(const float* red, int KS, int Fp, int ldr, int f, int c) {
    float v = red[f * ldr + c];
    for (int ks = 1; ks < KS; ++ks) v += red[(size_t(ks) * Fp + f) * ldr + c];
    return v;
}

// RMSNorm of all F rows of the residual, whose 32-column slices live in the `res` buffers of the cluster's CTAs, under the
// weight g (shared memory) -> operand planes.  Lane = column inside a slice, so a warp reads 128 contiguous bytes per peer.
// Split in two so that the caller can put its weight requests BETWEEN the peer reads and their use: issued first, 128 KB of
// weight loads per CTA queue in front of the DSMEM reads in the load pipe and delay the norm by a microsecond.
template <int D, int CSL, int ROWS>   // ROWS = rows per warp (1 for F <= 16, 2 for F <= 32); CSL = residual columns per rank
struct NormRows {
    static constexpr int NL = D / 32;   // loads per row: lane = column inside a 32-column block
    float v[ROWS][NL];
    // block j of this rank's rotation (every rank starts at its own columns, so the peers' ports are hit evenly)
    __device__ __forceinline__ static int block_of(int j, int rank) { return (j + rank * CSL / 32) & (NL - 1); }
    __device__ __forceinline__ void gather(const float* res_local, int F, int warp, int lane, int rank) {
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
            const int f = warp + i * SP_WARPS;
            if (f < F) {
#pragma unroll
                for (int j = 0; j < NL; ++j) {
                    const int col = 32 * block_of(j, rank) + lane;
                    v[i][j] = ld_dsmem(dsmem_addr(res_local + f * CSL + col % CSL, col / CSL));
                }
            }
        }
    }
    // stack-only mode: the normalised rows ARE the result; every CTA has all of them, rank r stores its own columns
    __device__ __forceinline__ void finish_out(const float* g, int F, float eps, float* out, int warp, int lane, int rank) {
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
            const int f = warp + i * SP_WARPS;
            if (f < F) {
                float sq = 0.f;
#pragma unroll
                for (int j = 0; j < NL; ++j) sq = fmaf(v[i][j], v[i][j], sq);
                const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
#pragma unroll
                for (int j = 0; j < NL; ++j) {
                    const int col = 32 * block_of(j, rank) + lane;
                    if (col / CSL == rank) out[size_t(f) * D + col] = v[i][j] * rstd * g[col];
                }
            }
        }
    }
    __device__ __forceinline__ void finish(const float* g, int F, float eps, __nv_bfloat16* ahi, __nv_bfloat16* alo, int lda, int warp,
                                           int lane, int rank) {
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
            const int f = warp + i * SP_WARPS;
            if (f < F) {
                float sq = 0.f;
#pragma unroll
                for (int j = 0; j < NL; ++j) sq = fmaf(v[i][j], v[i][j], sq);
                const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
#pragma unroll
                for (int j = 0; j < NL; ++j) {
                    const int col = 32 * block_of(j, rank) + lane;
                    put_planes(ahi, alo, f * lda + col, v[i][j] * rstd * g[col]);
                }
            }
        }
    }
};

// sum of v[k] over the TPC lanes that share a channel, scattered: lane sp returns the total of element sp.
// log2(TPC) rounds of TPC/2, TPC/4, ... independent shuffles instead of TPC x log2(TPC) dependent ones.
template <int TPC>
__device__ __forceinline__ float reduce_scatter(float (&v)[TPC], int sp) {
#pragma unroll
    for (int half = TPC / 2; half >= 1; half >>= 1) {
        const bool up = sp & half;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float keep = up ? v[i + half] : v[i];
            const float send = up ? v[i] : v[i + half];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

template <int NTF, int D, int DSL>
__global__ void __launch_bounds__(SP_THREADS, 1) stream_push_kernel(const mtn_stream_push_args a) {
    using L = SpL<NTF, D, DSL>;
    constexpr int Fp = 8 * NTF, N = D, di = 2 * D, R = L::R, NXp = L::NXp, CL = L::CL, lda = L::lda, ldu = L::ldu;
    constexpr int SP_DSL = DSL, SP_CSL = L::CSL, SP_MSL = L::MSL;   // d_inner channels / d_model columns / mask columns per CTA
    constexpr int TPC = SP_THREADS / DSL, SPT = 16 / TPC;           // scan: threads per channel, states per thread
    using Norm = NormRows<D, SP_CSL, (NTF + 1) / 2>;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = int(cluster_ctarank());
    const int b = blockIdx.x / CL;
    const int F = a.F;
    __nv_bfloat16* const act_hi = reinterpret_cast<__nv_bfloat16*>(smem + L::act_hi);
    __nv_bfloat16* const act_lo = reinterpret_cast<__nv_bfloat16*>(smem + L::act_lo);
    __nv_bfloat16* const su_hi = reinterpret_cast<__nv_bfloat16*>(smem + L::su_hi);
    __nv_bfloat16* const su_lo = reinterpret_cast<__nv_bfloat16*>(smem + L::su_lo);
    float* const xs = reinterpret_cast<float*>(smem + L::xs);
    float* const zs = reinterpret_cast<float*>(smem + L::zs);
    float* const us = reinterpret_cast<float*>(smem + L::us);
    float* const dl = reinterpret_cast<float*>(smem + L::dl);
    float* const dbl = reinterpret_cast<float*>(smem + L::dbl);
    float* const xd = reinterpret_cast<float*>(smem + L::xd);
    float* const res = reinterpret_cast<float*>(smem + L::res);
    float* const mixw = reinterpret_cast<float*>(smem + L::mixw);
    float* const frs = reinterpret_cast<float*>(smem + L::frs);
    float* const lvec = reinterpret_cast<float*>(smem + L::lvec);
    float* const s_norm = reinterpret_cast<float*>(smem + L::norm);
    float* const red = reinterpret_cast<float*>(smem + L::red);
    int tl_row = 0;

    // head blob: w_enc_t [16][N] | gamma [N] | beta [N] | norm_f [D] | w_dec [N][16]
    const float* w_enc_t = a.head;
    const float* gamma = w_enc_t + 16 * N;
    const float* beta = gamma + N;
    const float* norm_f = beta + N;
    const float* w_dec = norm_f + D;

    // rows >= F of the operand planes are multiplied too (their results are never read): keep them finite
    for (int i = tid; i < L::planes_end / 4; i += SP_THREADS) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    // RMSNorm weight of the first block
    for (int i = tid; i < D / 4; i += SP_THREADS) cp_async16(s_norm + 4 * i, a.layer_vec + 4 * i);
    cp_async_commit();
    __syncthreads();

    const bool stack_only = a.stack_x != nullptr;   // MambaBlocksSequential.forward(x, inference_params): no encoder / mask / decoder
    if (stack_only) {
        // the stack's input rows are the first residual (bimamba.py:446 with residual None): this CTA keeps its 32 columns
        const float* xb = a.stack_x + size_t(b) * F * D + rank * SP_CSL;
        for (int i = tid; i < F * SP_CSL; i += SP_THREADS) res[i] = __ldg(xb + size_t(i / SP_CSL) * D + i % SP_CSL);
    } else {
    // bottleneck weights: requested now, used after the encoder
    WFrag<GemmShape<SP_CSL / 16, N / 16>::NPRE> w_bot;
    gemm_issue<SP_CSL / 16, N / 16>(w_bot, reinterpret_cast<const uint4*>(a.bot_frag) + size_t(rank) * (SP_CSL / 16) * (N / 16) * 64, warp, lane);

    // ------------------------------------------------------------------ encoder + cLN (every CTA, all rows)
    {
        constexpr int nj = N / 32;
        const int m0 = (rank * SP_MSL) % N;   // first encoder channel of this CTA's mask columns
        // the encoder sees [8 carried samples | chunk]; the very first push has no carried samples
        const float* mixb = a.mix + size_t(b) * a.ld_mix;
        const float* tl = a.in_tail + size_t(b) * 8;
        const int shift = a.first ? 0 : 8;
        for (int f = warp; f < F; f += SP_WARPS) {
            float x[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int j = 8 * f + k - shift;
                x[k] = j >= 0 ? __ldg(mixb + j) : tl[j + 8];
            }
            float v[16];
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < nj) {
                    float acc = 0.f;
#pragma unroll
                    for (int k = 0; k < 16; ++k) acc = fmaf(__ldg(w_enc_t + k * N + lane + 32 * j), x[k], acc);
                    v[j] = fmaxf(acc, 0.f);
                    s += v[j];
                }
            const float mean = warp_sum(s) * (1.0f / N);
            float sq = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < nj) {
                    const float d = v[j] - mean;
                    sq = fmaf(d, d, sq);
                }
            const float rstd = rsqrtf(warp_sum(sq) * (1.0f / N) + a.eps_cln);
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < nj) {
                    const int n = lane + 32 * j;
                    put_planes(act_hi, act_lo, f * lda + n, fmaf(__ldg(gamma + n) * (v[j] - mean), rstd, __ldg(beta + n)));
                    if (n >= m0 && n < m0 + SP_MSL) mixw[f * SP_MSL + n - m0] = v[j];
                }
        }
    }
    __syncthreads();

    // ------------------------------------------------------------------ bottleneck: this CTA's 32 columns of h; residual := h
    {
        using G = GemmShape<SP_CSL / 16, N / 16>;
        const uint4* wf = reinterpret_cast<const uint4*>(a.bot_frag) + size_t(rank) * (SP_CSL / 16) * (N / 16) * 64;
        gemm_run<NTF, SP_CSL / 16, N / 16>(w_bot, wf, act_hi, act_lo, lda, red, warp, lane);
        __syncthreads();
        for (int i = tid; i < F * SP_CSL; i += SP_THREADS) res[i] = red_sum(red, G::KS, Fp, G::LDR, i / SP_CSL, i % SP_CSL);
    }
    }   // !stack_only
    cp_async_wait_all();
    cluster_sync_all();   // [A]

    // ------------------------------------------------------------------ Mamba blocks
    constexpr int ct_in = 2 * SP_DSL / 16, ks_d = D / 16, ct_x = NXp / 16, ks_c = SP_DSL / 16, ct_o = D / 16;
    constexpr size_t in_units = size_t(ct_in) * ks_d, x_units = size_t(ct_x) * ks_c, o_units = size_t(ct_o) * ks_c;
    const int ch0 = rank * SP_DSL;
    // this CTA's slice of a layer's small vectors in shared memory:
    // conv_w [64][4] | conv_b [64] | w_dt^T [R][64] | dt_bias [64] | A2 [64][16] | D [64]
    float* s_conv_w = lvec;
    float* s_conv_b = s_conv_w + SP_DSL * 4;
    float* s_w_dt = s_conv_b + SP_DSL;
    float* s_dt_bias = s_w_dt + SP_DSL * R;
    float* s_A2 = s_dt_bias + SP_DSL;
    float* s_D = s_A2 + SP_DSL * 16;
    using GI = GemmShape<ct_in, ks_d>;
    using GX = GemmShape<ct_x, ks_c>;
    using GO = GemmShape<ct_o, ks_c>;
    // asynchronous copy of a layer's small vectors (this CTA's slices) + the norm weight of the layer after it
    auto prefetch_vectors = [&](int layer) {
        const float* lv = a.layer_vec + size_t(layer) * a.layer_vec_stride;
        const float* conv_w = lv + D;
        const float* conv_b = conv_w + size_t(di) * 4;
        const float* w_dt = conv_b + di;
        const float* dt_bias = w_dt + size_t(di) * R;
        const float* A2 = dt_bias + di;
        const float* Dskip = A2 + size_t(di) * 16;
        const float* next_norm = layer + 1 < a.n_layers ? lv + a.layer_vec_stride : norm_f;
        constexpr int q_cw = SP_DSL, q_v = SP_DSL / 4, q_wd = R * (SP_DSL / 4), q_a2 = SP_DSL * 4, q_n = D / 4;   // 16-byte chunks
        for (int i = tid; i < q_cw + 3 * q_v + q_wd + q_a2 + q_n; i += SP_THREADS) {
            int j = i;
            if (j < q_cw) { cp_async16(s_conv_w + 4 * j, conv_w + size_t(ch0) * 4 + 4 * j); continue; }
            j -= q_cw;
            if (j < q_v) { cp_async16(s_conv_b + 4 * j, conv_b + ch0 + 4 * j); continue; }
            j -= q_v;
            if (j < q_wd) { cp_async16(s_w_dt + 4 * j, w_dt + size_t(j / (SP_DSL / 4)) * di + ch0 + 4 * (j % (SP_DSL / 4))); continue; }
            j -= q_wd;
            if (j < q_v) { cp_async16(s_dt_bias + 4 * j, dt_bias + ch0 + 4 * j); continue; }
            j -= q_v;
            if (j < q_a2) { cp_async16(s_A2 + 4 * j, A2 + size_t(ch0) * 16 + 4 * j); continue; }
            j -= q_a2;
            if (j < q_v) { cp_async16(s_D + 4 * j, Dskip + ch0 + 4 * j); continue; }
            j -= q_v;
            cp_async16(s_norm + ((layer + 1) & 1) * D + 4 * j, next_norm + 4 * j);
        }
        cp_async_commit();
    };
    prefetch_vectors(0);
    float h_prev[SPT] = {};
    float* hst_prev = nullptr;
    SP_MARK(0);
    for (int layer = 0; layer < a.n_layers; ++layer) {
        tl_row = 1 + layer;
        // layer fragment blob: in_proj [CL][ct_in][ks_d] | x_proj [CL][ct_x][ks_c] | out_proj [CL][ct_o][ks_c], 1 KiB each
        const uint4* lf = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(a.layer_frag) +
                                                         size_t(layer) * a.layer_frag_stride);
        const uint4* f_in = lf + size_t(rank) * in_units * 64;
        const uint4* f_x = lf + (size_t(CL) * in_units + size_t(rank) * x_units) * 64;
        const uint4* f_o = lf + (size_t(CL) * (in_units + x_units) + size_t(rank) * o_units) * 64;
        float* halo = a.halo + size_t(layer) * a.halo_layer_stride + size_t(b) * a.halo_stream_stride + ch0;
        float* hst = a.h + size_t(layer) * a.h_layer_stride + (size_t(b) * di + ch0) * 16;

        // Add -> RMSNorm (bimamba.py:446-447), all rows in every CTA.  Order of the requests: the peers' residual slices, then
        // everything that does not depend on this layer's activations (in_proj weights of this warp's unit, the SSM state of
        // this thread's (channel, state pair), the conv history), then the norm itself.
        Norm nr;
        nr.gather(res, F, warp, lane, rank);
        float h[SPT];
        {
            const float* hp = hst + (tid / TPC) * 16 + SPT * (tid % TPC);
#pragma unroll
            for (int q = 0; q < SPT; ++q) asm volatile("ld.global.f32 %0, [%1];" : "=f"(h[q]) : "l"(hp + q));
        }
        float halo_v = 0.f;
        // a.halo_rows = 4: the buffer is the reference's 4-wide conv_state (row 0 = the oldest input, never read: bimamba.py:274-277)
        static_assert(4 * SP_DSL <= SP_THREADS, "one thread per element of the conv history");
        if (tid < 3 * SP_DSL)
            asm volatile("ld.global.f32 %0, [%1];" : "=f"(halo_v) : "l"(halo + (a.halo_rows - 3 + tid / SP_DSL) * di + tid % SP_DSL));
        // the previous block's final SSM state goes out here, a whole phase away from the next cluster barrier: a global store
        // still in flight at `barrier.cluster.arrive.release` makes every thread sit in its memory barrier
        if (layer > 0) {
#pragma unroll
            for (int q = 0; q < SPT; ++q) hst_prev[(tid / TPC) * 16 + SPT * (tid % TPC) + q] = h_prev[q];
        }
        WFrag<GI::NPRE> w_in;
        gemm_issue<ct_in, ks_d>(w_in, f_in, warp, lane);
        nr.finish(s_norm + (layer & 1) * D, F, a.eps_rms, act_hi, act_lo, lda, warp, lane, rank);
        __syncthreads();
        SP_MARK(1);

        // in_proj: x columns [ch0, ch0+64) and z columns di + [ch0, ch0+64)
        gemm_run<NTF, ct_in, ks_d>(w_in, f_in, act_hi, act_lo, lda, red, warp, lane);
        if (tid < 3 * SP_DSL) xs[tid] = halo_v;   // conv history = rows -3..-1
        WFrag<GX::NPRE> w_x;
        gemm_issue<ct_x, ks_c>(w_x, f_x, warp, lane);
        __syncthreads();
        SP_MARK(2);
#pragma unroll
        for (int it = 0; it < (Fp * 2 * SP_DSL + SP_THREADS - 1) / SP_THREADS; ++it) {   // a thread's iterations are independent: unrolled for ILP
            const int i = tid + it * SP_THREADS;
            if (i < F * 2 * SP_DSL) {
                const int f = i / (2 * SP_DSL), c = i % (2 * SP_DSL);
                const float v = red_sum(red, GI::KS, Fp, GI::LDR, f, c);
                if (c < SP_DSL) xs[(3 + f) * SP_DSL + c] = v;
                else zs[f * SP_DSL + c - SP_DSL] = silu_f(v);
            }
        }
        cp_async_wait_all();
        __syncthreads();
        SP_MARK(3);

        // causal depthwise conv (width 4) + bias + SiLU (ssi.py:182); new history = last three conv inputs
#pragma unroll
        for (int it = 0; it < (Fp * SP_DSL + SP_THREADS - 1) / SP_THREADS; ++it) {
            const int i = tid + it * SP_THREADS;
            if (i < F * SP_DSL) {
                const int f = i / SP_DSL, c = i % SP_DSL;
                const float4 w = *reinterpret_cast<const float4*>(s_conv_w + 4 * c);
                const float* x = xs + f * SP_DSL + c;
                const float pre = fmaf(w.x, x[0], fmaf(w.y, x[SP_DSL], fmaf(w.z, x[2 * SP_DSL], fmaf(w.w, x[3 * SP_DSL], s_conv_b[c]))));
                const float u = silu_f(pre);
                us[i] = u;
                put_planes(su_hi, su_lo, f * ldu + c, u);
            }
        }
        __syncthreads();
        SP_MARK(4);

        // x_proj over this CTA's 64 channels: partial [dt | B | C] rows, left in xd for the peers
        gemm_run<NTF, ct_x, ks_c>(w_x, f_x, su_hi, su_lo, ldu, red, warp, lane);
        WFrag<GO::NPRE> w_o;
        gemm_issue<ct_o, ks_c>(w_o, f_o, warp, lane);   // out_proj weights travel while the partials are summed and the scan runs
        __syncthreads();
        for (int i = tid; i < F * NXp; i += SP_THREADS) xd[i] = red_sum(red, GX::KS, Fp, GX::LDR, i / NXp, i % NXp);
        SP_MARK(5);
        cluster_sync_all();   // [B]
        SP_MARK(6);
        // new conv history = last three conv inputs (xs is not touched again before the next block's in_proj epilogue); stored
        // here, right behind a barrier, for the reason given at the state store above
        if (tid < a.halo_rows * SP_DSL)
            halo[(tid / SP_DSL) * di + tid % SP_DSL] = xs[(F + 3 - a.halo_rows + tid / SP_DSL) * SP_DSL + tid % SP_DSL];
        for (int i = tid; i < F * NXp / 4; i += SP_THREADS)
            *reinterpret_cast<float4*>(dbl + 4 * i) = rank_sum4<CL>(xd + 4 * i, rank);
        __syncthreads();
        SP_MARK(7);

        // dt_proj + softplus (ssi.py:187, delta_softplus of :218)
#pragma unroll
        for (int it = 0; it < (Fp * SP_DSL + SP_THREADS - 1) / SP_THREADS; ++it) {
            const int i = tid + it * SP_THREADS;
            if (i < F * SP_DSL) {
                const int f = i / SP_DSL, c = i % SP_DSL;
                float acc0 = s_dt_bias[c], acc1 = 0.f;   // two chains halve the dependent-FMA depth
#pragma unroll
                for (int r = 0; r < R; r += 2) {
                    acc0 = fmaf(dbl[f * NXp + r], s_w_dt[r * SP_DSL + c], acc0);
                    acc1 = fmaf(dbl[f * NXp + r + 1], s_w_dt[(r + 1) * SP_DSL + c], acc1);
                }
                dl[i] = softplus_f(acc0 + acc1);
            }
        }
        __syncthreads();
        SP_MARK(8);

        // selective-scan steps: TPC threads per channel, SPT of the 16 states each (selective_scan_ref, ssi.py:91-157);
        // y = (sum_n C h + D u) * silu(z) -> operand planes of out_proj (over the u planes, which x_proj is done with).
        // TPC steps of operands are fetched before their recurrence runs and the TPC results are stored after it (the
        // compiler cannot move shared loads across the plane stores of an earlier step); the state sums of those steps are
        // reduce-scattered over the channel's lanes, so lane sp finishes step f0 + sp.
        {
            const int c = tid / TPC, sp = tid % TPC;
            float A[SPT];
#pragma unroll
            for (int q = 0; q < SPT; ++q) A[q] = s_A2[c * 16 + SPT * sp + q];
            const float Dp = s_D[c];
            const float* bc = dbl + R + SPT * sp;
#pragma unroll 1
            for (int f0 = 0; f0 < F; f0 += TPC) {
                float d[TPC], u[TPC], yv[TPC], Bv[TPC][SPT], Cv[TPC][SPT];
#pragma unroll
                for (int k = 0; k < TPC; ++k) {
                    const int f = f0 + k < F ? f0 + k : F - 1;
                    d[k] = dl[f * SP_DSL + c];
                    u[k] = us[f * SP_DSL + c];
#pragma unroll
                    for (int q = 0; q < SPT; ++q) {
                        Bv[k][q] = bc[f * NXp + q];
                        Cv[k][q] = bc[f * NXp + 16 + q];
                    }
                }
#pragma unroll
                for (int k = 0; k < TPC; ++k) {
                    const bool live = f0 + k < F;
                    const float du = d[k] * u[k];
                    float y = 0.f;
#pragma unroll
                    for (int q = 0; q < SPT; ++q) {
                        if (live) h[q] = fmaf(ex2_approx(d[k] * A[q]), h[q], du * Bv[k][q]);
                        y = fmaf(Cv[k][q], h[q], y);
                    }
                    yv[k] = y;
                }
                const float ysum = reduce_scatter<TPC>(yv, sp);
                float um = u[0];
#pragma unroll
                for (int k = 1; k < TPC; ++k) um = sp == k ? u[k] : um;
                if (f0 + sp < F) put_planes(su_hi, su_lo, (f0 + sp) * ldu + c, fmaf(Dp, um, ysum) * zs[(f0 + sp) * SP_DSL + c]);
            }
#pragma unroll
            for (int q = 0; q < SPT; ++q) h_prev[q] = h[q];
            hst_prev = hst;
        }
        __syncthreads();
        SP_MARK(9);
        if (layer + 1 < a.n_layers) prefetch_vectors(layer + 1);   // this layer's readers of the vector slices are done

        // out_proj over this CTA's 64 channels: partial h [F][D], left in plane 0 of red for the peers
        gemm_run<NTF, ct_o, ks_c>(w_o, f_o, su_hi, su_lo, ldu, red, warp, lane);
        if constexpr (GO::KS > 1) {
            __syncthreads();
            for (int i = tid; i < F * D; i += SP_THREADS) {
                const int f = i / D, c = i % D;
                red[f * (D + 4) + c] = red_sum(red, GO::KS, Fp, GO::LDR, f, c);   // each element has one owner: in place
            }
        }
        SP_MARK(10);
        cluster_sync_all();   // [C]
        SP_MARK(11);
        for (int i = tid; i < F * SP_CSL / 4; i += SP_THREADS) {
            const int f = i / (SP_CSL / 4), c = 4 * (i % (SP_CSL / 4));
            const float4 v = rank_sum4<CL>(red + f * (D + 4) + rank * SP_CSL + c, rank);
            float4 r = *reinterpret_cast<float4*>(res + f * SP_CSL + c);
            r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w;
            *reinterpret_cast<float4*>(res + f * SP_CSL + c) = r;
        }
        SP_MARK(12);
        cluster_sync_all();   // [A] of the next block / of norm_f
        SP_MARK(13);
    }
#pragma unroll
    for (int q = 0; q < SPT; ++q) hst_prev[(tid / TPC) * 16 + SPT * (tid % TPC) + q] = h_prev[q];   // last block's final state
    tl_row = 1 + a.n_layers;
    SP_MARK(0);

    if (stack_only) {   // norm_f (modules/mamba_blocks.py:196-197) is the output
        Norm nr;
        nr.gather(res, F, warp, lane, rank);
        nr.finish_out(s_norm + (a.n_layers & 1) * D, F, a.eps_rms, a.stack_out + size_t(b) * F * D, warp, lane, rank);
        cluster_sync_all();   // peers keep their residual slices alive until every CTA has gathered them
        return;
    }
    // ------------------------------------------------------------------ norm_f -> mask conv + ReLU -> mask * mix_w
    const uint4* f_mask = reinterpret_cast<const uint4*>(a.mask_frag) + size_t(rank) * (SP_MSL / 16) * (D / 16) * 64;
    WFrag<GemmShape<SP_MSL / 16, D / 16>::NPRE> w_mask;
    {
        Norm nr;
        nr.gather(res, F, warp, lane, rank);
        gemm_issue<SP_MSL / 16, D / 16>(w_mask, f_mask, warp, lane);
        nr.finish(s_norm + (a.n_layers & 1) * D, F, a.eps_rms, act_hi, act_lo, lda, warp, lane, rank);
    }
    __syncthreads();
    {
        using G = GemmShape<SP_MSL / 16, D / 16>;
        gemm_run<NTF, SP_MSL / 16, D / 16>(w_mask, f_mask, act_hi, act_lo, lda, red, warp, lane);
        __syncthreads();
        for (int i = tid; i < F * SP_MSL; i += SP_THREADS) {
            const int f = i / SP_MSL, c = i % SP_MSL;
            us[i] = fmaxf(red_sum(red, G::KS, Fp, G::LDR, f, c), 0.f) * mixw[i];   // sep slice (train_wsj0mix.py:91-92)
        }
    }
    __syncthreads();
    // decoder frames (ConvTranspose1d k=16 s=8): partial over this CTA's 64 encoder channels of ONE speaker
    {
        const int m0 = (rank * SP_MSL) % N;
        for (int i = tid; i < F * 16; i += SP_THREADS) {
            const int f = i >> 4, j = i & 15;
            float acc = 0.f;
#pragma unroll 8
            for (int c = 0; c < SP_MSL; ++c) acc = fmaf(us[f * SP_MSL + c], __ldg(w_dec + (m0 + c) * 16 + j), acc);
            frs[i] = acc;
        }
    }
    SP_MARK(1);
    cluster_sync_all();   // [D]
    // overlap-add on rank 0: frame f finalises samples [8f, 8f+8) together with the second half of frame f-1 (or the
    // carried tail); the second half of the last frame becomes the new tail
    if (rank == 0) {
        constexpr int Sn = 2, rps = N / SP_MSL;   // ranks per speaker
        float* tail = a.ola_tail + size_t(b) * Sn * 8;
        float* fsum = red;   // [Sn][F][16]
        for (int i = tid; i < Sn * F * 16; i += SP_THREADS) {
            const int s = i / (F * 16), fj = i % (F * 16);
            float v = 0.f;
            for (int r = s * rps; r < (s + 1) * rps; ++r) v += ld_dsmem(dsmem_addr(frs + fj, r));
            fsum[i] = v;
        }
        __syncthreads();
        float* est = a.est + size_t(b) * (8 * F) * Sn;
        for (int i = tid; i < 8 * F * Sn; i += SP_THREADS) {
            const int s = i % Sn, t = i / Sn, f = t >> 3, k = t & 7;
            float v = fsum[(s * F + f) * 16 + k];
            v += f >= 1 ? fsum[(s * F + f - 1) * 16 + 8 + k] : tail[s * 8 + k];
            est[i] = v;
        }
        __syncthreads();
        if (tid < Sn * 8) tail[tid] = fsum[((tid >> 3) * F + F - 1) * 16 + 8 + (tid & 7)];
        // the last 8 samples of this chunk open the next push's first frame (every CTA of the cluster is past its encoder)
        if (tid >= 32 && tid < 40)
            a.in_tail[size_t(b) * 8 + tid - 32] = a.mix[size_t(b) * a.ld_mix + (a.first ? 8 * F + 8 : 8 * F) - 8 + tid - 32];
        SP_MARK(2);
    }
    cluster_sync_all();   // [E] peers keep their shared memory alive until rank 0 has read the frames
}

}  // namespace mtn

using namespace mtn;

extern "C" size_t mtn_sizeof_stream_push_args(void) { return sizeof(mtn_stream_push_args); }

namespace {
template <int NTF>
size_t sp_smem_ntf(int D, int dsl) {
#define SP_SZ(D_)                                                     \
    if (D == D_) {                                                    \
        if (dsl == 64) return SpL<NTF, D_, 64>::total;                \
        if (dsl == 128) return SpL<NTF, D_, 128>::total;              \
        if constexpr (D_ <= 256) { if (dsl == 32) return SpL<NTF, D_, 32>::total; } \
        return 0;                                                     \
    }
    SP_SZ(64) SP_SZ(128) SP_SZ(256) SP_SZ(512)
#undef SP_SZ
    return 0;
}
}  // namespace

// Dynamic shared memory one CTA of the push kernel needs for F frames at d_model D and dsl channels per CTA; 0 = that
// combination does not exist (cluster size 2 * D / dsl outside 2..16).  A launch needs <= 227 KiB.
extern "C" size_t mtn_stream_push_smem_bytes(int F, int D, int dsl) {
    if (F < 1 || F > 32 || dsl <= 0 || 2 * D / dsl > 16 || D < dsl) return 0;
    switch ((F + 7) / 8) {
        case 1: return sp_smem_ntf<1>(D, dsl);
        case 2: return sp_smem_ntf<2>(D, dsl);
        case 3: return sp_smem_ntf<3>(D, dsl);
        default: return sp_smem_ntf<4>(D, dsl);
    }
}

extern "C" int mtn_stream_push_fwd(const mtn_stream_push_args* args, mtn_stream_t stream) {
    MTN_REQUIRE(args, "stream_push: null args");
    const mtn_stream_push_args& a = *args;
    MTN_REQUIRE(a.halo && a.h && a.head && a.layer_vec && a.layer_frag, "stream_push: null pointer");
    if (a.stack_x) MTN_REQUIRE(a.stack_out, "stream_push: stack_x without stack_out");
    else MTN_REQUIRE(a.mix && a.in_tail && a.est && a.ola_tail && a.bot_frag && a.mask_frag, "stream_push: null pointer");
    MTN_REQUIRE(a.halo_rows == 3 || a.halo_rows == 4, "stream_push: halo_rows=%d (3, or 4 = the reference's conv_state)", a.halo_rows);
    MTN_REQUIRE(a.halo_stream_stride >= size_t(a.halo_rows) * size_t(a.di) && a.halo_layer_stride >= a.halo_stream_stride * a.B,
                "stream_push: halo strides (floats): stream >= 3*di, layer >= B*stream");
    MTN_REQUIRE(a.B >= 1 && a.F >= 1 && a.F <= 32, "stream_push: B=%d F=%d (1 <= F <= 32 frames per push)", a.B, a.F);
    MTN_REQUIRE(a.N == a.D && a.di == 2 * a.D && a.n_spk == 2, "stream_push: needs enc_dim == d_model, expand 2, 2 speakers");
    MTN_REQUIRE(a.D == 64 || a.D == 128 || a.D == 256 || a.D == 512, "stream_push: d_model=%d (64, 128, 256 or 512)", a.D);
    MTN_REQUIRE(a.n_layers >= 1, "stream_push: layers=%d", a.n_layers);
    MTN_REQUIRE(a.stack_x || a.ld_mix >= 8 * a.F + (a.first ? 8 : 0), "stream_push: ld_mix=%d shorter than the chunk", a.ld_mix);
    MTN_REQUIRE(a.layer_vec_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(a.layer_vec) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a.layer_frag) & 15) == 0 && a.layer_frag_stride % 16 == 0 &&
                    (reinterpret_cast<uintptr_t>(a.head) & 15) == 0,
                "stream_push: weight blobs must be 16-byte aligned");
    const int DSL = a.dsl ? a.dsl : 64;
    MTN_REQUIRE(DSL == 32 || DSL == 64 || DSL == 128, "stream_push: dsl=%d channels per CTA (32, 64 or 128; 0 = 64)", a.dsl);
    const int CL = 2 * a.D / DSL;
    // CL >= 2: a CTA's mask columns (dsl of them) must lie inside one speaker's enc_dim = d_model columns
    MTN_REQUIRE(CL >= 2 && CL <= 16, "stream_push: d_model %d at %d channels per CTA needs a cluster of %d CTAs (2..16)", a.D, DSL, CL);
    const int NTF = (a.F + 7) / 8;
    MTN_REQUIRE(a.R == a.D / 16, "stream_push: dt_rank=%d, the recipes' ceil(d_model / 16) = %d is compiled in", a.R, a.D / 16);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const void* kern = nullptr;
    int smem_bytes = 0;
    static std::atomic<unsigned long long> done[4 * 4 * 3];
#define SP_PICK(NTF_, D_, DSL_)                                                        \
    do {                                                                               \
        kern = reinterpret_cast<const void*>(&stream_push_kernel<NTF_, D_, DSL_>);     \
        smem_bytes = SpL<NTF_, D_, DSL_>::total;                                       \
    } while (0)
#define SP_PICK_DSL(NTF_, D_)                                                          \
    do {                                                                               \
        if (DSL == 64) SP_PICK(NTF_, D_, 64);                                          \
        else if (DSL == 128) SP_PICK(NTF_, D_, 128);                                   \
        else if constexpr (D_ <= 256) SP_PICK(NTF_, D_, 32);                           \
    } while (0)
#define SP_PICK_D(NTF_)                                   \
    switch (a.D) {                                        \
        case 64: SP_PICK_DSL(NTF_, 64); break;            \
        case 128: SP_PICK_DSL(NTF_, 128); break;          \
        case 256: SP_PICK_DSL(NTF_, 256); break;          \
        default: SP_PICK_DSL(NTF_, 512); break;           \
    }
    switch (NTF) {
        case 1: SP_PICK_D(1) break;
        case 2: SP_PICK_D(2) break;
        case 3: SP_PICK_D(3) break;
        default: SP_PICK_D(4) break;
    }
#undef SP_PICK_D
#undef SP_PICK_DSL
#undef SP_PICK
    MTN_REQUIRE(kern, "stream_push: no kernel for d_model %d at %d channels per CTA", a.D, DSL);
    MTN_REQUIRE(smem_bytes <= 227 * 1024, "stream_push: %d B of shared memory needed", smem_bytes);
    const int slot = ((NTF - 1) * 4 + (a.D == 64 ? 0 : a.D == 128 ? 1 : a.D == 256 ? 2 : 3)) * 3 + (DSL == 32 ? 0 : DSL == 64 ? 1 : 2);
    {
        int dev = -1;
        cudaGetDevice(&dev);
        const bool tracked = dev >= 0 && dev < 64;
        if (!tracked || !((done[slot].load(std::memory_order_acquire) >> dev) & 1ull)) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (e != cudaSuccess) {
                set_error("stream_push: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
                return MTN_ECUDA;
            }
            if (tracked) done[slot].fetch_or(1ull << dev, std::memory_order_release);
        }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(a.B) * CL, 1, 1);
    cfg.blockDim = dim3(SP_THREADS, 1, 1);
    cfg.dynamicSmemBytes = size_t(smem_bytes);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = unsigned(CL);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    void* kargs[1] = {const_cast<mtn_stream_push_args*>(args)};
    cudaError_t e = cudaLaunchKernelExC(&cfg, kern, kargs);
    if (e != cudaSuccess) {
        set_error("stream_push: launch failed (cluster of %d CTAs, %d B smem): %s", CL, smem_bytes, cudaGetErrorString(e));
        return MTN_ECUDA;
    }
    return MTN_OK;
}
