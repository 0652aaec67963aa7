// Selective scan, "trio" mapping (sm_100a): per 32 channels one RECURRENCE warp + one PREP warp + one POST warp.
//
// Same arithmetic, shared-memory layout and recurrence code as the pair mapping (mtn_scan_pair.cuh); the pair's helper
// warp is cut in two along its two independent jobs, because the measurements showed the helper's serial
// instruction stream -- not the XU pipe, not HBM -- to be what bounds the kernel with four in-order warps per scheduler:
//   * PREP warp  (warp 4 + w): TMA ring duty (rotating), dt_proj (FFMA2) + 1-MUFU softplus -> delta, bf16 planes -> u,
//     written into the pair-private slot of tile k; runs up to two tiles ahead of the recurrence;
//   * POST warp  (warp 8 + w): silu(z) fetch (cp.async, double buffered, one tile ahead), gate, hi/lo split, staging and
//     the 16-byte row-segment stores of y; runs one tile behind the recurrence and releases the slot (slotfree) to PREP.
// Six warps per scheduler instead of four give the schedulers two more independent instruction streams to cover the
// LDS / MUFU / FMA latencies.  Registers: the kernel is launched with 80 per thread (384 threads, 2 CTAs per SM) and
// re-partitioned per warp group with setmaxnreg: recurrence 128, prep 64, post 48 (sum = 3 x 80).
#pragma once

namespace mtn {

template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

template <int P, int NDBL>
struct ScanSmemTrio {
    using Base = ScanSmemPair<P, NDBL, false, 16>;
    static constexpr int ZBUF2_BYTES = Base::ZBUF_BYTES;     // second silu(z) buffer per pair (double buffering)
    static constexpr int EXTRA_BAR = 128;                     // slotfree[4 * SLOTS]
    static constexpr int TOTAL = Base::TOTAL + ZBUF2_BYTES + EXTRA_BAR;
    static_assert(2 * (TOTAL + 1024) <= 233472, "two CTAs per SM must fit");
};

// ------------------------------------------------------------------------------------------------------ PREP warp
template <int P, int R, int NDBL, int ABL>
__device__ __forceinline__ void scan_trio_prep(uint8_t* ring, float* slots, uint64_t* full_bar, uint64_t* empty_bar,
                                               uint64_t* prepped, uint64_t* slotfree, const ScanParams& p, int w, int lane,
                                               int ch0, int d, int b, int dir, int ntiles, int Lb,
                                               const CUtensorMap* mapU, const CUtensorMap* mapD) {
    using SM = ScanSmemPair<P, NDBL, false, 16>;
    constexpr int S = SM::STAGES, NSLOT = SM::SLOTS, NB = SM::NB;
    constexpr int G = 4;                                          // rows per pass (register budget of this warp: 64)
    const int L = p.L;
    const size_t pd = size_t(dir) * p.di + d;
    const int chl = w * 32 + lane;
    float2 wdt2[R / 2];
    {
        const float4* wp = reinterpret_cast<const float4*>(p.w_dt + pd * R);
#pragma unroll
        for (int q = 0; q < R / 4; ++q) {
            const float4 x = wp[q];
            wdt2[2 * q] = make_float2(x.x, x.y);
            wdt2[2 * q + 1] = make_float2(x.z, x.w);
        }
    }
    const float bias = p.dt_bias[pd];
    float sdl_sum = 0.f;
    auto tile_of = [&](int i) { return dir ? (ntiles - 1 - i) : i; };
    auto issue_tile = [&](int i2) {
        const int stg = i2 % S;
        const int row0 = b * L + tile_of(i2) * SC_TT;
        mbar_arrive_expect_tx(&full_bar[stg], SM::STAGE_BYTES);
        uint8_t* dst = ring + stg * SM::STAGE_BYTES;
        tma_load_3d(dst, mapU, &full_bar[stg], dir * p.di + ch0, row0, 0);
        tma_load_2d(dst + SM::U_BYTES, mapD, &full_bar[stg], dir * p.n_dbl, row0);
    };
    if (w == 0 && lane == 0) {
#pragma unroll 1
        for (int t = 0; t < S; ++t)
            if (t < ntiles) issue_tile(t);
    }
    __syncwarp();
#pragma unroll 1
    for (int k = 0; k < ntiles; ++k) {
        // producer duty of this iteration (rotates over the four prep warps): refill the ring stage of tile k-2 once
        // every recurrence warp released it (the prep warps finished with that stage before the recurrence started it)
        const int pt = k - 2;
        if ((k & 3) == w) {
            const bool refill = pt >= 0 && pt + S < ntiles;
            if (refill) {
                mbar_wait_sleep(&empty_bar[pt % S], uint32_t(pt / S) & 1u);
                if (lane == 0) issue_tile(pt + S);
            }
            __syncwarp();
        }
        const int stg = k % S, sl = k % NSLOT;
        // the slot is free once the post warp has stored tile k - NSLOT
        if (k >= NSLOT) mbar_wait_sleep(&slotfree[w * NSLOT + sl], uint32_t(k / NSLOT - 1) & 1u);
        const int t0 = tile_of(k) * SC_TT;
        const int nvalid = min(SC_TT, Lb - t0);
        mbar_wait_sleep(&full_bar[stg], uint32_t(k / S) & 1u);
        const uint8_t* st = ring + stg * SM::STAGE_BYTES;
        const __nv_bfloat16* su = reinterpret_cast<const __nv_bfloat16*>(st) + chl;
        const float* sd = reinterpret_cast<const float*>(st + SM::U_BYTES);
        float* sdl = slots + sl * (4 * 2 * SC_TT * 32) + w * (2 * SC_TT * 32) + lane;
#pragma unroll 1
        for (int g = 0; g < SC_TT; g += G) {
            float pre[G];
            {
                float2 pa[G], pb[G];
#pragma unroll
                for (int r = 0; r < G; ++r) {
                    pa[r] = make_float2(bias, 0.f);
                    pb[r] = make_float2(0.f, 0.f);
                }
#pragma unroll
                for (int q = 0; q < ((ABL & 8) ? 1 : R / 4); ++q) {
#pragma unroll
                    for (int r = 0; r < G; ++r) {
                        const float4 x = *reinterpret_cast<const float4*>(sd + (g + r) * NB + 4 * q);
                        pa[r] = __ffma2_rn(make_float2(x.x, x.y), wdt2[2 * q], pa[r]);
                        pb[r] = __ffma2_rn(make_float2(x.z, x.w), wdt2[2 * q + 1], pb[r]);
                    }
                }
#pragma unroll
                for (int r = 0; r < G; ++r) {
                    const float2 pab = __fadd2_rn(pa[r], pb[r]);
                    pre[r] = pab.x + pab.y;
                }
            }
            float ev[G], qv[G];
#pragma unroll
            for (int r = 0; r < G; ++r) {
                ev[r] = (ABL & 1) ? 0.5f : ex2_approx(-1.4426950408889634f * fabsf(pre[r]));
                qv[r] = 0.0051261021414032125f;
            }
            constexpr float SPC[8] = {-0.02907406467853027f, 0.07751608674076167f, -0.13602247622393474f,
                                      0.19076880735651539f,  -0.24835398988480129f, 0.3331812170752912f,
                                      -0.49999444976340335f, 0.9999999659255092f};
#pragma unroll
            for (int c = 0; c < ((ABL & 1) ? 1 : 8); ++c) {
#pragma unroll
                for (int r = 0; r < G; ++r) qv[r] = fmaf(qv[r], ev[r], SPC[c]);
            }
#pragma unroll
            for (int r = 0; r < G; ++r) {
                float dl = fmaf(qv[r], ev[r], fmaxf(pre[r], 0.f));   // softplus_1mufu, see mtn_scan.cu
                dl = (g + r < nvalid) ? dl : 0.f;  // rows past the utterance end: exp2(0) = 1, delta*u = 0 -> state unchanged
                float uval = bf16_bits_to_float(su + (g + r) * SC_CH);
                if (P == 2) uval += bf16_bits_to_float(su + (SC_TT + g + r) * SC_CH);
                sdl_sum += dl;
                sdl[(g + r) * 32] = dl;
                sdl[(SC_TT + g + r) * 32] = uval;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&prepped[w * NSLOT + sl]);
    }
    if (p.sum_delta) p.sum_delta[(size_t(dir) * p.batch + b) * p.di + d] = sdl_sum;
}

// ------------------------------------------------------------------------------------------------------ POST warp
template <int P, int NDBL, typename ZT, int ABL>
__device__ __forceinline__ void scan_trio_post(float* slots, uint8_t* zbuf_all, uint64_t* ydone, uint64_t* slotfree,
                                               const ScanParams& p, int w, int lane, int ch0, int b, int dir, int ntiles,
                                               int Lb) {
    using SM = ScanSmemPair<P, NDBL, false, 16>;
    constexpr int NSLOT = SM::SLOTS;
    constexpr bool ZF = sizeof(ZT) == 4;
    constexpr int ZPAIR = SC_TT * 32 * 4;                        // bytes of one pair's z tile buffer
    const int L = p.L;
    const size_t M = size_t(p.batch) * L;
    const size_t y_plane = M * 2 * p.di;
    const int ldy = 2 * p.di;                                     // y row stride, bf16 elements
    const int frow = lane >> 2, fseg = lane & 3;                  // flush mapping: (row, 16-byte segment of 8 channels)
    __nv_bfloat16* yflush = p.y + size_t(dir) * p.di + (ch0 + w * 32 + fseg * 8);
    // gate copies (4 bytes each): fp32 z: lane = channel, one row per copy; bf16 z: lane = (row parity, channel pair)
    const int zc_row0 = ZF ? 0 : (lane >> 4);
    const int zc_col = ZF ? lane : 2 * (lane & 15);
    const ZT* zsrc = reinterpret_cast<const ZT*>(p.z) + p.z_col0 + ch0 + w * 32 + zc_col;
    constexpr int ZROWS = ZF ? 1 : 2;
    auto tile_of = [&](int i) { return dir ? (ntiles - 1 - i) : i; };
    auto zbuf_of = [&](int i) { return reinterpret_cast<ZT*>(zbuf_all + (i & 1) * (4 * ZPAIR) + w * ZPAIR); };
    auto request_z = [&](int i) {
        const int t0 = tile_of(i) * SC_TT;
        const int nvalid = min(SC_TT, Lb - t0);
        const ZT* zs = zsrc + (size_t(b) * L + t0 + zc_row0) * size_t(p.ldz);
        const size_t zstep = size_t(ZROWS) * p.ldz;
        ZT* zdst = zbuf_of(i) + zc_row0 * 32 + zc_col;
#pragma unroll
        for (int r = 0; r < SC_TT; r += ZROWS) {
            if (r + zc_row0 < nvalid) cp_async_4(zdst + r * 32, zs);
            zs += zstep;
        }
        cp_async_commit();
    };
    if (ntiles > 0) request_z(0);
#pragma unroll 1
    for (int pt = 0; pt < ntiles; ++pt) {
        const int sl = pt % NSLOT;
        // tile pt + 1's gate goes into the other z buffer (tile pt - 1's was consumed in the previous iteration)
        if (pt + 1 < ntiles) {
            request_z(pt + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");   // tile pt's copies have landed
        } else {
            cp_async_wait_all();
        }
        mbar_wait_sleep(&ydone[w * NSLOT + sl], uint32_t(pt / NSLOT) & 1u);
        __syncwarp();
        if (!(ABL & 2)) {
            const int t0 = tile_of(pt) * SC_TT;
            const int nvalid = min(SC_TT, Lb - t0);
            float* slot = slots + sl * (4 * 2 * SC_TT * 32) + w * (2 * SC_TT * 32);
            const float* sy = slot + lane;
            const ZT* zb = zbuf_of(pt);
            // the u half of the slot is dead now: stage the gated outputs there as bf16 [P][16 rows][32 ch]
            uint16_t* stage = reinterpret_cast<uint16_t*>(slot + SC_TT * 32);
#pragma unroll 1
            for (int g = 0; g < SC_TT; g += 8) {
                float yv[8], zv[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    yv[r] = sy[(g + r) * 32];
                    zv[r] = ldz(zb + (g + r) * 32 + lane);
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float y = yv[r] * (0.5f * zv[r]);
                    const uint32_t hi = f2bf_lo(y);
                    sts_b16(stage + (g + r) * 32 + lane, hi);
                    if (P == 2) sts_b16(stage + (SC_TT + g + r) * 32 + lane, f2bf_lo(y - __uint_as_float(hi << 16)));
                }
            }
            __syncwarp();
            __nv_bfloat16* ytile = yflush + (size_t(b) * L + t0) * size_t(ldy);
#pragma unroll
            for (int pl = 0; pl < P; ++pl) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int row = frow + 8 * hh;
                    if (row < nvalid) {
                        const uint4 v = *reinterpret_cast<const uint4*>(stage + (pl * SC_TT + row) * 32 + fseg * 8);
                        *reinterpret_cast<uint4*>(ytile + pl * y_plane + size_t(row * ldy)) = v;
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&slotfree[w * NSLOT + sl]);
    }
}

template <int P, int R, int NDBL, typename ZT, int ABL>
__global__ void __launch_bounds__(384, 2)
scan_kernel_trio(const __grid_constant__ CUtensorMap mapU, const __grid_constant__ CUtensorMap mapD, const ScanParams p) {
    using SM = ScanSmemPair<P, NDBL, false, 16>;
    using ST = ScanSmemTrio<P, NDBL>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint8_t* ring = smem;
    float* slots = reinterpret_cast<float*>(smem + SM::STAGES * SM::STAGE_BYTES);
    uint8_t* zbuf = smem + SM::STAGES * SM::STAGE_BYTES + SM::SLOTS * SM::SLOT_BYTES;       // two buffers of ZBUF_BYTES
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(zbuf + SM::ZBUF_BYTES + ST::ZBUF2_BYTES);
    uint64_t* empty_bar = full_bar + 4;
    uint64_t* prepped = empty_bar + 4;
    uint64_t* ydone = prepped + 4 * SM::SLOTS;
    uint64_t* slotfree = ydone + 4 * SM::SLOTS;

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
            mbar_init(&slotfree[s], 1);
        }
        fence_barrier_init();
    }
    __syncthreads();
    const int Lb = (b == p.batch - 1) ? p.L_last : p.L;
    const int ntiles = (Lb + SC_TT - 1) / SC_TT;
    const int w = warp & 3;
    const int d = ch0 + w * 32 + lane;
    if (warp < 4) {
        reg_inc<128>();
        scan_pair_recur<SM::STAGES, SM::NB, R, true, ABL>(ring, slots, SM::STAGE_BYTES, SM::U_BYTES, full_bar, empty_bar,
                                                         prepped, ydone, p, w, lane, d, b, dir, ntiles, Lb);
    } else if (warp < 8) {
        reg_dec<64>();
        scan_trio_prep<P, R, NDBL, ABL>(ring, slots, full_bar, empty_bar, prepped, slotfree, p, w, lane, ch0, d, b, dir,
                                        ntiles, Lb, &mapU, &mapD);
    } else {
        reg_dec<48>();
        scan_trio_post<P, NDBL, ZT, ABL>(slots, zbuf, ydone, slotfree, p, w, lane, ch0, b, dir, ntiles, Lb);
    }
}

template <int P, int R, int NDBL, typename ZT, int ABL = 0>
static int launch_scan_trio(const mtn_scan_args* a, cudaStream_t stream) {
    using SM = ScanSmemPair<P, NDBL, false, 16>;
    using ST = ScanSmemTrio<P, NDBL>;
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
    auto kern = scan_kernel_trio<P, R, NDBL, ZT, ABL>;
    static std::atomic<unsigned long long> attr_done{0};   // per template instantiation, one bit per device
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), ST::TOTAL, attr_done, "scan(trio)")) return rc;
    dim3 grid(p.ndirs * (a->di / SC_CH), a->batch, 1);
    kern<<<grid, 384, ST::TOTAL, stream>>>(mapU, mapD, p);
    MTN_CUDA_LAUNCH_CHECK("scan(trio)");
    return MTN_OK;
}

}  // namespace mtn
