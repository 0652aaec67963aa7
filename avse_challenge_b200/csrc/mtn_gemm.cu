// Dense contractions of the Mamba-TasNet path on tcgen05 tensor cores (sm_100a).
//
//   out[M, N] = epilogue( A[M, K] * W[N, K]^T )        (x groups)
//
// Replaces cuBLAS/cuDNN calls of the reference: in_proj (modules/mamba/bimamba.py:192-196), x_proj
// (modules/mamba/selective_scan_interface.py:186), out_proj (bimamba.py:253), bottleneck / mask 1x1 convs
// (modules/mamba_masknet.py:121,123).
//
// Design (B200-first, not a cuBLAS wrapper):
//  * A and W live in HBM as bf16 "planes".  P = 2 is the fp32-accurate mode: value = hi + lo and the kernel issues
//    three tcgen05.mma per K-step (hi*hi + lo*hi + hi*lo) into one fp32 TMEM accumulator -- ~2^-17 operand
//    precision at the HBM traffic of plain fp32, with no in-kernel conversion pass.  P = 1 is bf16 mode.
//  * Persistent, warp-specialised CTA (one per SM): warp 0 = TMA producer (128B-swizzled boxes straight into the
//    UMMA canonical K-major layout), warp 1 = single-thread tcgen05.mma issuer, warps 2-9 = epilogue.  Two TMEM
//    accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
//  * These GEMMs are HBM-bound (K <= 1024, output-write dominated), so the epilogue is the part that matters:
//    TMEM -> registers -> per-warp padded smem transpose -> fully coalesced 128-bit global stores, with the
//    SiLU(z) gate / relu(mask)*mix_w fused in.
#include "mtn_ptx.cuh"
#include "mtn_host.h"
#include <stdlib.h>
#include <string.h>

namespace mtn {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int STG_LD = 20;  // floats per staging row (16 + 4 pad: conflict-free 128-bit accesses)
constexpr int EPI_WARPS = 8;  // two per TMEM lane quarter: they take alternate 16-column chunks of the accumulator

struct GemmParams {
    void* out;
    const float* aux;
    int M, N, K;
    int ldo, ld_aux;
    int groups, out_group_stride;
    int epi_param;
    int tiles_m, tiles_n;
    int desc;        // 1: row tiles are walked from the last to the first (see launch_gemm)
    __nv_bfloat16* out2;
    float* rowsum;
    const float* rowsq;
    int ldo2;
    size_t plane2;   // elements between the planes of out2
    float rowsq_scale, rowsq_eps;
    int rowsq_parts;
};

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (2-CTA cluster, cta_group::2) per 256 x BN tile: each CTA stages its
// own 128 rows of A and HALF of the B tile per k-block, so a stage is 2/3 of the one-CTA stage (64 KB instead of 96 KB at
// P = 2, BN = 256: three stages instead of two) and each SM pulls a third less operand data through the crossbar -- the
// wide fp32-mode GEMMs sat at 55 % / 72 % tensor-pipe utilisation with a two-deep ring (profiles/r01/gemm_ncu_metrics.txt).
// BKT = K extent of one pipeline stage: 64 (rows of 128 B, SWIZZLE_128B) or 32 (rows of 64 B, SWIZZLE_64B).  Half-depth stages
// make the ring twice as deep in the same shared memory: at P = 2, BN = 256 a 64-deep stage is 96 KB, so only two fit and the
// TMA request for k-block k + 2 goes out when k-block k retires -- one k-block (1 536 tensor-pipe cycles) before its data is
// needed, less than a 96 KB load takes under load (in_proj: tensor pipe 55 % active).  With 32-deep stages (48 KB, four of them)
// the request leads by three stages.
template <int P, int BN, int CG = 1, int BKT = BK, int EW = EPI_WARPS, bool TS = false>
struct GemmCfg {
    static_assert(EW == 8 || EW == 16, "epilogue warps: two or four per TMEM lane quarter");
    static_assert(BKT == 64 || BKT == 32, "stage depth: 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B) bf16 elements");
    static constexpr int A_BYTES = BM * BKT * 2;
    static constexpr int B_BYTES = (BN / CG) * BKT * 2;      // bytes of B this CTA stages per plane
    static constexpr int STAGE_BYTES = P * (A_BYTES + B_BYTES);
    // TS (bulk tensor stores): two dense 32 x 16 boxes (2 KB each, 512-byte aligned for the swizzle) per epilogue warp
    static constexpr int STAGING_BYTES = TS ? EW * 4096 : EW * 32 * STG_LD * 4;
    static constexpr int BAR_BYTES = 256;
    static constexpr int BUDGET = 227 * 1024 - 1024 - STAGING_BYTES - BAR_BYTES;
    static constexpr int STAGES_RAW = BUDGET / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
    static constexpr int ACC_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
    static constexpr int TMEM_COLS = 2 * ACC_COLS;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + STAGING_BYTES + BAR_BYTES;
    static_assert(STAGES >= 2, "need at least a double-buffered operand pipeline");
    static_assert(B_BYTES % 1024 == 0, "B tile must keep 1024B alignment for SWIZZLE_128B");
};

template <int P, int BN, int EPI, bool OUT_BF16, int CG = 1, int BKT = BK, int EW = EPI_WARPS, bool TS = false>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const __grid_constant__ CUtensorMap mapO, const GemmParams p) {
    using Cfg = GemmCfg<P, BN, CG, BKT, EW, TS>;
    static_assert(!TS || EPI == MTN_EPI_STORE || EPI == MTN_EPI_INPROJ, "bulk-store epilogue: plain store and in_proj only");
    constexpr int ESPLIT = EW / 4;   // epilogue warps per TMEM lane quarter: each takes every ESPLIT-th 16-column chunk
    static_assert(EPI != MTN_EPI_RESADD || EW == 8, "the row-sum planes of the resadd epilogue are laid out for two warps per quarter");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-B align inside the shared window without leaving the shared address space (keeps LDS/STS)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* staging = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES);
    uint64_t* empty_bar = full_bar + Cfg::STAGES;
    uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0);   // warp-uniform for the compiler too (role / chunk tests)
    const int lane = threadIdx.x & 31;
    const int crank = CG == 2 ? int(cluster_ctarank()) : 0;   // 0 = the pair's leader (issues the MMAs)

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);            // CG 2: only the leader's is used (both CTAs' loads signal it)
            mbar_init(&empty_bar[s], 1);           // CG 2: the leader's commit arrives on both CTAs' barriers
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], CG * EW);  // one arrive per epilogue warp (CG 2: of both CTAs, on the leader's)
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if (CG == 2) tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
        else tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all();   // the peer's barriers must be initialised before anything is signalled on them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int kblocks = p.K / BKT;
    // CG 2: the tile loop runs over PAIR tiles of 256 rows; CTA `crank` owns rows [mt * 256 + crank * 128, +128)
    const int tiles_m = CG == 2 ? (p.tiles_m + 1) / 2 : p.tiles_m;
    const int tiles_per_group = tiles_m * p.tiles_n;
    const int total_tiles = tiles_per_group * p.groups;
    const int tile0 = CG == 2 ? int(blockIdx.x) / 2 : int(blockIdx.x);
    const int tile_step = CG == 2 ? int(gridDim.x) / 2 : int(gridDim.x);

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (one thread)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = tile0; tile < total_tiles; tile += tile_step) {
                const int g = tile / tiles_per_group;
                const int r = tile - g * tiles_per_group;
                const int mq = r / p.tiles_n;
                const int nt = r - mq * p.tiles_n;
                const int mt = p.desc ? tiles_m - 1 - mq : mq;
                const int arow = (mt * CG + crank) * BM;                  // rows past M are zero-filled by TMA
                const int brow = g * p.N + nt * BN + crank * (BN / CG);   // this CTA's share of the B tile
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (CG == 1 || crank == 0) mbar_arrive_expect_tx(&full_bar[stage], CG * Cfg::STAGE_BYTES);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + P * Cfg::A_BYTES;
#pragma unroll
                    for (int pl = 0; pl < P; ++pl) {
                        if (CG == 2) {
                            tma_load_3d_2sm(sa + pl * Cfg::A_BYTES, &mapA, &full_bar[stage], g * p.K + kb * BKT, arow, pl);
                            tma_load_3d_2sm(sb + pl * Cfg::B_BYTES, &mapB, &full_bar[stage], kb * BKT, brow, pl);
                        } else {
                            tma_load_3d(sa + pl * Cfg::A_BYTES, &mapA, &full_bar[stage], g * p.K + kb * BKT, arow, pl);
                            tma_load_3d(sb + pl * Cfg::B_BYTES, &mapB, &full_bar[stage], kb * BKT, brow, pl);
                        }
                    }
                    if (++stage == Cfg::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0 && crank == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(CG * BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = tile0; tile < total_tiles; tile += tile_step) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + P * Cfg::A_BYTES;
#pragma unroll
                    for (int kk = 0; kk < BKT / 16; ++kk) {
                        auto make_desc = [](uint32_t addr) { return BKT == 64 ? make_smem_desc_sw128(addr) : make_smem_desc_sw64(addr); };
                        const uint64_t a_hi = make_desc(sa + kk * 32);
                        const uint64_t b_hi = make_desc(sb + kk * 32);
                        auto mma = [&](uint64_t da, uint64_t db, uint32_t accum) {
                            if (CG == 2) tc_mma_bf16_2sm(d_tmem, da, db, idesc, accum);
                            else tc_mma_bf16(d_tmem, da, db, idesc, accum);
                        };
                        mma(a_hi, b_hi, (kb | kk) != 0 ? 1u : 0u);
                        if (P == 2) {
                            const uint64_t a_lo = make_desc(sa + Cfg::A_BYTES + kk * 32);
                            const uint64_t b_lo = make_desc(sb + Cfg::B_BYTES + kk * 32);
                            mma(a_lo, b_hi, 1u);
                            mma(a_hi, b_lo, 1u);
                        }
                    }
                    // smem slot reusable once these MMAs retire (CG 2: in both CTAs)
                    if (CG == 2) tc_commit_2sm(&empty_bar[stage]);
                    else tc_commit(&empty_bar[stage]);
                    if (++stage == Cfg::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (CG == 2) tc_commit_2sm(&tfull_bar[acc]);   // accumulator complete -> both CTAs' epilogues
                else tc_commit(&tfull_bar[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue warps (2..9)
        // One warp alone on an SM sub-partition is latency-bound (~5 cycles per instruction), and the epilogue, not the
        // MMA, paced these output-heavy GEMMs: eight warps, two per TMEM lane quarter, split the accumulator columns.
        const int e = warp - 2;
        const int q = warp & 3;    // TMEM lane quarter this warp may read (warps 2..9 cover every quarter twice)
        const int chalf = e >> 2;  // this warp takes the 16-column chunks c0 = 16*chalf, +32, ...
        float* stg = staging + e * 32 * STG_LD;
        int acc = 0;
        uint32_t acc_phase = 0;
        float ss_next[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t nstore = 0;   // TS: bulk stores issued by this warp so far (selects the staging box)
        auto load_ss = [&](int tile2) {  // sum of the partial-sum planes (fixed order) for this lane's 4 rows of tile2
            if (tile2 >= total_tiles) return;
            const int r2 = tile2 % tiles_per_group;
            const int mq2 = r2 / p.tiles_n;
            const int rowb = ((p.desc ? tiles_m - 1 - mq2 : mq2) * CG + crank) * BM + q * 32 + (lane >> 2);
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int grow = rowb + it * 8;
                float acc2 = 0.f;
                if (grow < p.M)
                    for (int k = 0; k < p.rowsq_parts; ++k) acc2 += p.rowsq[size_t(k) * p.M + grow];
                ss_next[it] = acc2;
            }
        };
        auto load_res = [&](float4(&dst)[4], int row0_, int gcol0) {  // old residual, 4 row groups x this lane's float4
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int grow = row0_ + it * 8 + (lane >> 2);
                dst[it] = (p.epi_param && grow < p.M)
                              ? *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.out) +
                                                                 size_t(grow) * p.ldo + gcol0 + (lane & 3) * 4)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto release_acc = [&](int acc_) {   // hand the accumulator back to the (leader's) MMA thread
            if (CG == 2 && crank != 0) mbar_arrive_remote(&tempty_bar[acc_], 0);
            else mbar_arrive(&tempty_bar[acc_]);
        };
        if (p.rowsq) load_ss(tile0);
        for (int tile = tile0; tile < total_tiles; tile += tile_step) {
            const int g = tile / tiles_per_group;
            const int r = tile - g * tiles_per_group;
            const int mq = r / p.tiles_n;
            const int mt = (p.desc ? tiles_m - 1 - mq : mq) * CG + crank;     // this CTA's 128-row tile
            const int nt = r - mq * p.tiles_n;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_base = tmem_base + acc * Cfg::ACC_COLS + (uint32_t(q * 32) << 16);
            const int row0 = mt * BM + q * 32;
            bool released = false;
            float sqacc[4] = {0.f, 0.f, 0.f, 0.f};  // RESADD: sum of squares of this lane's 4 rows over the warp's chunks
            // folded RMSNorm: one scale per row.  The partial sums of the NEXT tile's rows are requested now and turned
            // into rstd when that tile starts, so their global latency never sits on the epilogue's critical path.
            float rstd[4] = {1.f, 1.f, 1.f, 1.f};
            if (p.rowsq) {
#pragma unroll
                for (int it = 0; it < 4; ++it) rstd[it] = rsqrtf(fmaf(ss_next[it], p.rowsq_scale, p.rowsq_eps));
                load_ss(tile + tile_step);
            }
            float4 auxn[4];  // RESADD: residual rows of the warp's next chunk, requested one chunk ahead
            if (EPI == MTN_EPI_RESADD) load_res(auxn, row0, nt * BN + 16 * chalf);
            // Control flow below is warp-uniform on purpose (chunk index, activation / dt-column tests per CHUNK, row validity
            // only as a predicate on the memory accesses): per-lane `if (row < M) { ... }` / `if (col >= split)` blocks had cost
            // ~15 branch / convergence-barrier / predicate instructions per 16-byte store (ncu source page, in_proj: 3.5 M BRA,
            // 2.4 M BSSY + 2.4 M BSYNC, 2.6 M ISETP, 2.3 M R2UR for 1.0 M STG), in warps that run latency-bound.  Rows past M hold
            // zeros (TMA zero-fills the A rows, guarded aux loads return zeros), so computing them is harmless.
            constexpr int NCH = BN / 16;
            if constexpr (TS) {
                // Bulk-store epilogue (plain store / in_proj, no folded norm): the thread that read row `lane` of the chunk from
                // TMEM applies the activation in registers, writes its 16 values (64 B fp32 / 32 B bf16) into a dense, swizzled
                // 32 x 16 box in shared memory, and one lane hands the box to the TMA unit -- no transposing read-back, no
                // per-row address arithmetic, no global store instructions in the warp.  Rows past M are clipped by the tensor map.
                uint8_t* sbox = reinterpret_cast<uint8_t*>(staging) + e * 4096;   // two boxes, used alternately (nstore)
#pragma unroll
                for (int j = 0; j < (NCH + ESPLIT - 1) / ESPLIT; ++j) {
                    const int ch = j * ESPLIT + chalf;
                    if (NCH % ESPLIT != 0 && ch >= NCH) break;
                    const int c0 = ch * 16;
                    uint32_t v[16];
                    tmem_ld_x16(t_base + c0, v);
                    tmem_ld_wait();
                    if (ch + ESPLIT >= NCH) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) release_acc(acc);
                        released = true;
                    }
                    const int ccol = nt * BN + c0;
                    if (EPI == MTN_EPI_INPROJ && ccol >= p.epi_param) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(silu_sel<OUT_BF16>(__uint_as_float(v[i])));
                    }
                    uint8_t* box = sbox + (nstore++ & 1) * 2048;
                    if (lane == 0) bulk_wait_group_read<1>();   // the store issued from this buffer two chunks ago has read it
                    __syncwarp();
                    if (OUT_BF16) {     // rows of 32 B, SWIZZLE_32B: 16-byte chunk c of row r sits at c ^ ((r >> 2) & 1)
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                            pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
                        }
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                            *reinterpret_cast<uint4*>(box + lane * 32 + ((c ^ ((lane >> 2) & 1)) << 4)) =
                                make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                    } else {            // rows of 64 B, SWIZZLE_64B: chunk c of row r sits at c ^ ((r >> 1) & 3)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            *reinterpret_cast<uint4*>(box + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) =
                                make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&mapO, box, g * p.out_group_stride + ccol, row0);
                        bulk_commit_group();
                    }
                }
            } else
#pragma unroll
            for (int j = 0; j < (NCH + ESPLIT - 1) / ESPLIT; ++j) {
                const int ch = j * ESPLIT + chalf;      // this warp's j-th 16-column chunk
                if (NCH % ESPLIT != 0 && ch >= NCH) break;
                const int c0 = ch * 16;
                uint32_t v[16];
                tmem_ld_x16(t_base + c0, v);
                tmem_ld_wait();
                if (ch + ESPLIT >= NCH) {
                    // last TMEM read of this accumulator by this warp: hand it back to the MMA warp early
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) release_acc(acc);
                    released = true;
                }
                // registers (thread = row) -> padded smem
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    *reinterpret_cast<float4*>(&stg[lane * STG_LD + 4 * jj]) =
                        make_float4(__uint_as_float(v[4 * jj]), __uint_as_float(v[4 * jj + 1]), __uint_as_float(v[4 * jj + 2]),
                                    __uint_as_float(v[4 * jj + 3]));
                __syncwarp();
                // smem -> global, row-contiguous: 4 lanes cover one 64-byte row segment, 8 rows per pass
                const int rsub = lane >> 2;
                const int c4 = lane & 3;
                const int ccol = nt * BN + c0;           // first column of the chunk within the group's N (warp-uniform)
                const int gcol = ccol + c4 * 4;
                const bool chunk_flag = EPI == MTN_EPI_INPROJ ? ccol >= p.epi_param      // SiLU half (split is a multiple of 16)
                                      : EPI == MTN_EPI_XPROJ ? ccol < p.epi_param        // dt columns (RP = 16 or 32)
                                                             : false;
                float4 aux[4];
                if (EPI == MTN_EPI_RESADD) {
#pragma unroll
                    for (int it = 0; it < 4; ++it) aux[it] = auxn[it];
                    if (ch + ESPLIT < NCH) load_res(auxn, row0, ccol + 16 * ESPLIT);
                }
                if (EPI == MTN_EPI_MASK) {  // issue the four mix_w loads together (they were the epilogue's critical path)
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int grow = row0 + it * 8 + rsub;
                        aux[it] = grow < p.M ? *reinterpret_cast<const float4*>(p.aux + size_t(grow) * p.ld_aux +
                                                                               (gcol % p.epi_param))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int rr = it * 8 + rsub;
                    const int grow = row0 + rr;
                    const bool ok = grow < p.M;
                    float4 f = *reinterpret_cast<const float4*>(&stg[rr * STG_LD + c4 * 4]);
                    if (p.rowsq) {  // RMSNorm of the A operand's rows, applied after the contraction
                        f.x *= rstd[it];
                        f.y *= rstd[it];
                        f.z *= rstd[it];
                        f.w *= rstd[it];
                    }
                    if (EPI == MTN_EPI_RESADD) {
                        f.x += aux[it].x;
                        f.y += aux[it].y;
                        f.z += aux[it].z;
                        f.w += aux[it].w;
                        // the planes of the updated residual: A operand of the next in_proj / mask GEMM (optional: without out2
                        // and rowsum the epilogue is just `res += result`, the residual add of bimamba.py:446 moved into out_proj)
                        if (p.out2 != nullptr) {
                        __nv_bfloat16* pr = p.out2 + size_t(grow) * p.ldo2 + gcol;
                        __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
                        if (P == 2) {
                            split_bf16(f.x, h0, l0);
                            split_bf16(f.y, h1, l1);
                            split_bf16(f.z, h2, l2);
                            split_bf16(f.w, h3, l3);
                        } else {
                            h0 = __float2bfloat16_rn(f.x);
                            h1 = __float2bfloat16_rn(f.y);
                            h2 = __float2bfloat16_rn(f.z);
                            h3 = __float2bfloat16_rn(f.w);
                        }
                        __nv_bfloat162 a2 = __halves2bfloat162(h0, h1), b2 = __halves2bfloat162(h2, h3);
                        uint2 ph;
                        ph.x = *reinterpret_cast<uint32_t*>(&a2);
                        ph.y = *reinterpret_cast<uint32_t*>(&b2);
                        if (ok) *reinterpret_cast<uint2*>(pr) = ph;
                        if (P == 2) {
                            __nv_bfloat162 c2 = __halves2bfloat162(l0, l1), d2 = __halves2bfloat162(l2, l3);
                            uint2 pl;
                            pl.x = *reinterpret_cast<uint32_t*>(&c2);
                            pl.y = *reinterpret_cast<uint32_t*>(&d2);
                            if (ok) *reinterpret_cast<uint2*>(pr + p.plane2) = pl;
                        }
                        }
                        sqacc[it] = fmaf(f.x, f.x, fmaf(f.y, f.y, fmaf(f.z, f.z, fmaf(f.w, f.w, sqacc[it]))));
                    }
                    if (EPI == MTN_EPI_INPROJ) {
                        if (chunk_flag) {
                            f.x = silu_sel<OUT_BF16>(f.x);   // bf16 output: 1-MUFU form (rounded to bf16 right below)
                            f.y = silu_sel<OUT_BF16>(f.y);
                            f.z = silu_sel<OUT_BF16>(f.z);
                            f.w = silu_sel<OUT_BF16>(f.w);
                        }
                    } else if (EPI == MTN_EPI_RELU) {
                        f.x = fmaxf(f.x, 0.f);
                        f.y = fmaxf(f.y, 0.f);
                        f.z = fmaxf(f.z, 0.f);
                        f.w = fmaxf(f.w, 0.f);
                    } else if (EPI == MTN_EPI_MASK) {
                        f.x = fmaxf(f.x, 0.f) * aux[it].x;
                        f.y = fmaxf(f.y, 0.f) * aux[it].y;
                        f.z = fmaxf(f.z, 0.f) * aux[it].z;
                        f.w = fmaxf(f.w, 0.f) * aux[it].w;
                    }
                    if (EPI == MTN_EPI_XPROJ) {
                        // dt columns also go out as hi | lo bf16 planes: the B operand of the scan's dt_proj MMA
                        if (chunk_flag) {
                            __nv_bfloat16* drow = reinterpret_cast<__nv_bfloat16*>(const_cast<float*>(p.aux)) +
                                                  (size_t(grow) * p.groups + g) * 2 * p.epi_param + gcol;
                            __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
                            split_bf16(f.x, h0, l0);
                            split_bf16(f.y, h1, l1);
                            split_bf16(f.z, h2, l2);
                            split_bf16(f.w, h3, l3);
                            __nv_bfloat162 a2 = __halves2bfloat162(h0, h1), b2 = __halves2bfloat162(h2, h3);
                            __nv_bfloat162 c2 = __halves2bfloat162(l0, l1), d2 = __halves2bfloat162(l2, l3);
                            uint2 ph, pl;
                            ph.x = *reinterpret_cast<uint32_t*>(&a2);
                            ph.y = *reinterpret_cast<uint32_t*>(&b2);
                            pl.x = *reinterpret_cast<uint32_t*>(&c2);
                            pl.y = *reinterpret_cast<uint32_t*>(&d2);
                            if (ok) {
                                *reinterpret_cast<uint2*>(drow) = ph;
                                *reinterpret_cast<uint2*>(drow + p.epi_param) = pl;
                            }
                        }
                    }
                    const size_t off = size_t(grow) * p.ldo + size_t(g) * p.out_group_stride + gcol;
                    if (OUT_BF16) {
                        __nv_bfloat162 lo2 = __floats2bfloat162_rn(f.x, f.y);
                        __nv_bfloat162 hi2 = __floats2bfloat162_rn(f.z, f.w);
                        uint2 pk;
                        pk.x = *reinterpret_cast<uint32_t*>(&lo2);
                        pk.y = *reinterpret_cast<uint32_t*>(&hi2);
                        if (ok) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off) = pk;
                    } else {
                        if (ok) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off) = f;
                    }
                }
                __syncwarp();
            }
            if (EPI == MTN_EPI_RESADD && p.rowsum != nullptr) {
                // row sums of squares: 4 lanes share a row; every (N tile, warp half) owns one partial-sum plane, so the
                // sums are plain stores (nothing to zero, bit-reproducible); the consumer adds the planes in order
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    float sq = sqacc[it];
                    sq += __shfl_xor_sync(0xffffffffu, sq, 1);
                    sq += __shfl_xor_sync(0xffffffffu, sq, 2);
                    const int grow = row0 + it * 8 + (lane >> 2);
                    if ((lane & 3) == 0 && grow < p.M) p.rowsum[size_t(nt * 2 + chalf) * p.M + grow] = sq;
                }
            }
            if (!released) {  // a warp with no chunk of its own (cannot happen for BN >= 32) still has to release
                tc_fence_before();
                __syncwarp();
                if (lane == 0) release_acc(acc);
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }

    if (TS && warp >= 2 && lane == 0) bulk_wait_group_all();   // this lane's bulk stores have left shared memory and are performed
    tc_fence_before();
    if (CG == 2) cluster_sync_all();   // neither CTA may leave (or free TMEM) while the pair's MMAs / remote arrives are in flight
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
        else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int P, int BN, int EPI, bool OUT_BF16, int CG = 1, int BKT = BK, int EW = EPI_WARPS, bool TS = false>
static int launch_gemm(const mtn_gemm_args* a, cudaStream_t stream) {
    using Cfg = GemmCfg<P, BN, CG, BKT, EW, TS>;
    constexpr CUtensorMapSwizzle swz = BKT == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUtensorMap mapA, mapB;
    {
        uint64_t dims[3] = {(uint64_t)a->lda, (uint64_t)a->a_rows, (uint64_t)P};
        uint64_t str[2] = {(uint64_t)a->lda * 2, (uint64_t)a->a_rows * a->lda * 2};
        uint32_t box[3] = {BKT, BM, 1};
        // Narrow-N fp32-mode GEMMs (x_proj) stream A once from HBM, 128 B of every row per k-block: promoting those requests to
        // 256 B brings the next k-block's piece of the row along.  B200, S 32 x 3 999: x_proj 0.1056 -> 0.0982 ms (83 -> 89 % of
        // the HBM peak), bit-identical; neutral to -4 % for the wide tiles and the bf16-mode x_proj (97 % already), so only here
        // (profiles/r02/gemm_l2_promotion_S_fp32.jsonl).
        const CUtensorMapL2promotion promoA =
            (P == 2 && BN <= 64) ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
        if (!encode_tmap(&mapA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a->a, dims, str, box, swz, promoA))
            return MTN_ECUDA;
    }
    {
        uint64_t rows = (uint64_t)a->groups * a->N;
        uint64_t dims[3] = {(uint64_t)a->K, rows, (uint64_t)P};
        uint64_t str[2] = {(uint64_t)a->K * 2, rows * a->K * 2};
        uint32_t box[3] = {BKT, BN / CG, 1};
        if (!encode_tmap(&mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a->w, dims, str, box, swz))
            return MTN_ECUDA;
    }
    CUtensorMap mapO;
    memset(&mapO, 0, sizeof(mapO));
    if (TS) {   // output as a 2-D tensor [M][ldo]; one box = 32 rows x 16 columns of a chunk
        uint64_t dims[2] = {(uint64_t)a->ldo, (uint64_t)a->M};
        uint64_t str[1] = {(uint64_t)a->ldo * (OUT_BF16 ? 2 : 4)};
        uint32_t box[2] = {16, 32};
        if (!encode_tmap(&mapO, OUT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a->out, dims, str,
                         box, OUT_BF16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B))
            return MTN_ECUDA;
    }
    GemmParams p;
    p.out = a->out;
    p.aux = reinterpret_cast<const float*>(a->aux);
    p.M = a->M;
    p.N = a->N;
    p.K = a->K;
    p.ldo = a->ldo;
    p.ld_aux = a->ld_aux;
    p.groups = a->groups;
    p.out_group_stride = a->out_group_stride;
    p.epi_param = a->epi_param;
    p.out2 = reinterpret_cast<__nv_bfloat16*>(a->out2);
    p.rowsum = a->rowsum;
    p.rowsq = a->rowsq;
    p.ldo2 = a->ldo2;
    p.plane2 = size_t(a->a2_rows) * a->ldo2;
    p.rowsq_scale = a->rowsq_scale;
    p.rowsq_eps = a->rowsq_eps;
    p.rowsq_parts = a->rowsq_parts;
    p.tiles_m = (a->M + BM - 1) / BM;
    p.tiles_n = a->N / BN;
    // Row tiles are walked from the LAST to the first.  Every other kernel of the layer (norm, conv, scan, decoder) walks the
    // tokens upwards, and each kernel's input is what the previous one has just written or read, 0.1 - 0.5 GB against 126 MB of
    // L2: walking the same way, a consumer starts at the rows that were evicted first and pushes out the newest ones before it
    // reaches them; walking the opposite way it starts on whatever part of the stream is still in L2.  MTN_GEMM_ORDER = 0 in
    // the environment restores ascending order (A/B runs).
    p.desc = 1;
    if (const char* v = getenv("MTN_GEMM_ORDER")) p.desc = atoi(v) != 0;
    auto kern = gemm_tcgen05_kernel<P, BN, EPI, OUT_BF16, CG, BKT, EW, TS>;
    static std::atomic<unsigned long long> attr_done{0};   // per template instantiation, one bit per device
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES, attr_done, "gemm")) return rc;
    int cap = a->max_ctas > 0 ? a->max_ctas : num_sms();
    if (CG == 2) {
        // one CTA pair (a 2-CTA cluster: both SMs of a TPC) per 256-row tile; persistent over at most cap / 2 pairs
        const int pairs_total = ((p.tiles_m + 1) / 2) * p.tiles_n * p.groups;
        int pairs = cap / 2 < pairs_total ? cap / 2 : pairs_total;
        if (pairs < 1) pairs = 1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs, 1, 1);
        cfg.blockDim = dim3(64 + 32 * EW, 1, 1);
        cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mapA, mapB, mapO, p);
        if (e != cudaSuccess) {
            set_error("gemm(2cta): cluster launch failed: %s", cudaGetErrorString(e));
            return MTN_ECUDA;
        }
        return MTN_OK;
    }
    int total = p.tiles_m * p.tiles_n * p.groups;
    int grid = total < cap ? total : cap;
    kern<<<grid, 64 + 32 * EW, Cfg::SMEM_BYTES, stream>>>(mapA, mapB, mapO, p);
    MTN_CUDA_LAUNCH_CHECK("gemm");
    return MTN_OK;
}

// CTA pairs (cta_group::2) for the mainloop-bound GEMMs: 256-wide N tiles, K >= 1024 and at least one 256-row pair tile per
// SM pair.  Measured on B200 (tools/gemm_bench.py --ab, profiles/r02/gemm_2cta_ab_*.jsonl; outputs bit-identical):
// out_proj (K = 1024) 0.177 -> 0.156 ms = 94 % of the sustained bf16 tensor peak, L bf16 out_proj (K = 2048) 0.489 -> 0.445;
// but in_proj (K = 256: four k-blocks per tile) 0.206 -> 0.227 and mask 0.195 -> 0.216 -- with so short a main loop the
// accumulator hand-off between MMA and epilogue is paid per tile, and in a pair it crosses two SMs.  Hence the K rule; it
// was K >= 512 at first, which also paired the L recipe's in_proj / bottleneck (K = 512, eight k-blocks): 0.636 -> 0.707 ms and
// 0.175 -> 0.203 ms in bf16 mode (gemm_2cta_ab_L_bf16.jsonl), so the bar is sixteen k-blocks.
// MTN_GEMM_2CTA = 0 / 1 / 2 in the environment: never / by this rule (default) / whenever the shape allows (A/B runs).
static bool use_cta_pairs(const mtn_gemm_args* a, int bn) {
    if (bn != 256 || a->groups != 1) return false;
    int mode = 1;
    if (const char* v = getenv("MTN_GEMM_2CTA")) mode = atoi(v);
    if (mode == 0) return false;
    if (mode == 1 && a->K < 1024) return false;
    const long pair_tiles = long((a->M + 2 * BM - 1) / (2 * BM)) * (a->N / bn);
    const long cap = a->max_ctas > 0 ? a->max_ctas : num_sms();
    return pair_tiles >= cap / 2;
}

// MTN_GEMM_HALF_STAGES in the environment (A/B runs of tools/gemm_bench.py): see dispatch_epi.  Unset = -1 = the shipped rule.
static int half_depth_stages() {
    if (const char* v = getenv("MTN_GEMM_HALF_STAGES")) return atoi(v);
    return -1;
}

// Bulk-store epilogue (GemmCfg TS): plain-store and in_proj GEMMs without the folded norm, output rows 16-byte aligned.
// MTN_GEMM_TMA_STORE = 0 in the environment keeps the register -> shared -> STG epilogue (A/B runs).
static bool bulk_store_ok(const mtn_gemm_args* a) {
    if (a->rowsq || (a->epilogue != MTN_EPI_STORE && a->epilogue != MTN_EPI_INPROJ)) return false;
    if (const char* v = getenv("MTN_GEMM_TMA_STORE")) if (atoi(v) == 0) return false;
    return (size_t(a->ldo) * (a->out_bf16 ? 2 : 4)) % 16 == 0;
}

template <int P, int BN>
static int dispatch_epi(const mtn_gemm_args* a, cudaStream_t s) {
    if (BN == 256 && use_cta_pairs(a, BN)) {
        if (a->epilogue == MTN_EPI_STORE && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_STORE, false, 2>(a, s);
        if (a->epilogue == MTN_EPI_INPROJ && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_INPROJ, false, 2>(a, s);
        if (a->epilogue == MTN_EPI_INPROJ && a->out_bf16) return launch_gemm<P, BN, MTN_EPI_INPROJ, true, 2>(a, s);
        if (a->epilogue == MTN_EPI_MASK && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_MASK, false, 2>(a, s);
        if (a->epilogue == MTN_EPI_RESADD && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_RESADD, false, 2>(a, s);
    }
    if constexpr (P == 2 && BN == 256) {
        // fp32-mode in_proj / bottleneck: the bulk-store epilogue (B200, S 32 x 3 999, profiles/r02/gemm_tma_store_*.jsonl, bit-identical):
        // in_proj 0.202 -> 0.192 ms, bottleneck 0.058 -> 0.056; no gain for x_proj's narrow tiles or the bf16-mode GEMMs (not wired).
        if (bulk_store_ok(a)) {
            if (a->epilogue == MTN_EPI_STORE && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_STORE, false, 1, BK, 8, true>(a, s);
            if (a->epilogue == MTN_EPI_INPROJ && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_INPROJ, false, 1, BK, 8, true>(a, s);
        }
    }
    if constexpr (P == 2 && BN == 256) {
        // Mask GEMM (fp32 mode): its epilogue waits on four global mix_w loads per chunk, so it runs with sixteen epilogue warps
        // (four per TMEM lane quarter) and, to make room for their staging buffers, half-depth stages: 0.195 -> 0.139 ms at
        // BASELINE config 2 (profiles/r02/gemm_epilogue_variants_S_fp32.jsonl).  in_proj / bottleneck do not gain from either
        // (0.201 -> 0.209 / 0.211 ms): MTN_GEMM_HALF_STAGES = 1 / 2 selects them there for A/B runs, 0 switches the mask rule off.
        const int hs = half_depth_stages();
        if (a->epilogue == MTN_EPI_MASK && !a->out_bf16 && hs != 0) return launch_gemm<P, BN, MTN_EPI_MASK, false, 1, 32, 16>(a, s);
        if (hs == 2) {
            if (a->epilogue == MTN_EPI_STORE && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_STORE, false, 1, 32, 16>(a, s);
            if (a->epilogue == MTN_EPI_INPROJ && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_INPROJ, false, 1, 32, 16>(a, s);
        } else if (hs == 1) {
            if (a->epilogue == MTN_EPI_STORE && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_STORE, false, 1, 32>(a, s);
            if (a->epilogue == MTN_EPI_INPROJ && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_INPROJ, false, 1, 32>(a, s);
        }
    }
    if constexpr (P == 1 && BN == 256) {
        // bf16 mode: 48 KB stages, three of them fit beside sixteen warps' staging buffers, and the output-heavy GEMMs are
        // epilogue-bound there (one tensor pass per k-step): L hparams, 64 x 3999 (profiles/r02/gemm_ew16_L_bf16.jsonl), in_proj
        // 0.563 -> 0.518 ms, mask 0.682 -> 0.446 ms, bit-identical.  MTN_GEMM_HALF_STAGES = 0 keeps eight warps (A/B runs).
        if (half_depth_stages() != 0) {
            if (a->epilogue == MTN_EPI_MASK && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_MASK, false, 1, 64, 16>(a, s);
            if (a->epilogue == MTN_EPI_INPROJ && a->out_bf16) return launch_gemm<P, BN, MTN_EPI_INPROJ, true, 1, 64, 16>(a, s);
        }
    }
    if (a->epilogue == MTN_EPI_STORE && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_STORE, false>(a, s);
    if (BN <= 64) {
        if (a->epilogue == MTN_EPI_XPROJ && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_XPROJ, false>(a, s);
    }
    if (BN == 64) {
        // in_proj of a single-row-tile problem (streaming chunk): see dispatch_bn
        if (a->epilogue == MTN_EPI_INPROJ && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_INPROJ, false>(a, s);
        if (a->epilogue == MTN_EPI_INPROJ && a->out_bf16) return launch_gemm<P, BN, MTN_EPI_INPROJ, true>(a, s);
        if (a->epilogue == MTN_EPI_RESADD && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_RESADD, false>(a, s);
        // per-speaker grouped end_conv1x1 of DPMamba at enc_dim = 64 (unit-test sizes)
        if (a->epilogue == MTN_EPI_MASK && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_MASK, false>(a, s);
        if (a->epilogue == MTN_EPI_RELU && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_RELU, false>(a, s);
    }
    if (BN >= 128) {
        if (a->epilogue == MTN_EPI_RESADD && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_RESADD, false>(a, s);
        if (a->epilogue == MTN_EPI_INPROJ && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_INPROJ, false>(a, s);
        if (a->epilogue == MTN_EPI_INPROJ && a->out_bf16) return launch_gemm<P, BN, MTN_EPI_INPROJ, true>(a, s);
        if (a->epilogue == MTN_EPI_MASK && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_MASK, false>(a, s);
        if (a->epilogue == MTN_EPI_RELU && !a->out_bf16) return launch_gemm<P, BN, MTN_EPI_RELU, false>(a, s);
    }
    set_error("gemm: unsupported epilogue %d / out_bf16 %d for N tile %d", a->epilogue, a->out_bf16, BN);
    return MTN_EINVAL;
}

static int tile_n_for(int N) { return N % 256 == 0 ? 256 : N % 128 == 0 ? 128 : N; }

template <int P>
static int dispatch_bn(const mtn_gemm_args* a, cudaStream_t s) {
    const int N = a->N;
    // Under-filled problems (fewer 256-wide tiles than SMs: streaming chunks, short utterances) are latency-bound, not
    // throughput-bound: narrower N tiles put up to 4x as many CTAs on them and cut each CTA's weight load and epilogue
    // by as much (B200, S hparams, one stream, 20 ms pushes = one row tile: 0.733 ms per push with 256-wide tiles, 0.590
    // with 128, 0.574 with 64).  Same K order per output element, so the result is bit-identical.  (RESADD keeps its
    // tile width: the row-sum plane count depends on it; XPROJ only exists for narrow N.)
    const bool narrowable = a->epilogue == MTN_EPI_STORE || a->epilogue == MTN_EPI_INPROJ ||
                            a->epilogue == MTN_EPI_MASK || a->epilogue == MTN_EPI_RELU;
    const long row_tiles = long((a->M + BM - 1) / BM) * a->groups;
    const long cap = a->max_ctas > 0 ? a->max_ctas : num_sms();
    if (narrowable && N > 64) {
        if (N % 256 == 0 && row_tiles * (N / 256) >= cap) return dispatch_epi<P, 256>(a, s);
        if (N % 128 == 0 && row_tiles * (N / 128) >= cap) return dispatch_epi<P, 128>(a, s);
        if (N % 64 == 0) return dispatch_epi<P, 64>(a, s);
    }
    if (N % 256 == 0) return dispatch_epi<P, 256>(a, s);
    if (N % 128 == 0) return dispatch_epi<P, 128>(a, s);
    if (N == 64) return dispatch_epi<P, 64>(a, s);
    if (N == 48) return dispatch_epi<P, 48>(a, s);
    set_error("gemm: unsupported N=%d (need 48, 64 or a multiple of 128)", N);
    return MTN_EINVAL;
}

}  // namespace mtn

extern "C" int mtn_gemm_rowsum_parts(int N) { return 2 * (N / mtn::tile_n_for(N)); }

extern "C" int mtn_gemm_fwd(const mtn_gemm_args* a, mtn_stream_t stream) {
    using namespace mtn;
    MTN_REQUIRE(a && a->a && a->w && a->out, "gemm: null pointer");
    MTN_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0 && a->groups >= 1, "gemm: bad shape M=%d N=%d K=%d groups=%d", a->M,
                a->N, a->K, a->groups);
    MTN_REQUIRE(a->K % BK == 0, "gemm: K=%d must be a multiple of %d", a->K, BK);
    MTN_REQUIRE(a->lda % 8 == 0 && a->lda >= a->groups * a->K, "gemm: lda=%d invalid", a->lda);
    MTN_REQUIRE(a->a_rows >= a->M, "gemm: a_rows=%d < M=%d", a->a_rows, a->M);
    MTN_REQUIRE(a->ldo % 4 == 0 && a->out_group_stride % 4 == 0, "gemm: ldo/out_group_stride must be multiples of 4");
    MTN_REQUIRE((reinterpret_cast<uintptr_t>(a->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->w) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
                "gemm: pointers must be 16-byte aligned");
    if (a->epilogue == MTN_EPI_MASK)
        MTN_REQUIRE(a->aux && a->epi_param > 0 && a->epi_param % 4 == 0 && a->ld_aux % 4 == 0,
                    "gemm: mask epilogue needs aux / enc_dim");
    if (a->epilogue == MTN_EPI_INPROJ) MTN_REQUIRE(a->epi_param % 16 == 0, "gemm: inproj split must be a multiple of 16");
    if (a->rowsq) MTN_REQUIRE(a->rowsq_parts >= 1 && a->rowsq_parts <= 64, "gemm: rowsq_parts=%d", a->rowsq_parts);
    if (a->epilogue == MTN_EPI_RESADD)
        MTN_REQUIRE(!a->out_bf16 && a->groups == 1 &&
                        ((!a->out2 && !a->rowsum) || (a->out2 && a->rowsum && a->ldo2 % 4 == 0 && a->a2_rows >= a->M &&
                                                      (reinterpret_cast<uintptr_t>(a->out2) & 7) == 0)),
                    "gemm: resadd epilogue needs fp32 out and either no out2 / rowsum (plain `out += result`) or both (out2 planes "
                    "8-byte aligned, ldo2 %% 4 == 0, a2_rows >= M)");
    if (a->epilogue == MTN_EPI_XPROJ)
        MTN_REQUIRE(a->aux && (a->epi_param == 16 || a->epi_param == 32) && a->epi_param <= a->N &&
                        (reinterpret_cast<uintptr_t>(a->aux) & 15) == 0,
                    "gemm: xproj epilogue needs a 16-byte aligned dtp buffer and RP = 16 or 32");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (a->planes == 2) return dispatch_bn<2>(a, s);
    if (a->planes == 1) return dispatch_bn<1>(a, s);
    set_error("gemm: planes=%d (must be 1 or 2)", a->planes);
    return MTN_EINVAL;
}
