// On-device SI-SNR with n-speaker PIT (n <= 4) and SI-SNR improvement over the mixture (sm_100a).
//
// Replaces, for the evaluation front end of the reference (Mamba-TasNet/train_wsj0mix.py:503-604 `save_results`):
//     sisnr          = compute_objectives(predictions, targets)          (train_wsj0mix.py:548; loss =
//                      speechbrain get_si_snr_with_pitwrapper [3P], hparams/WSJ0Mix/mambatasnet_S.yaml:163)
//     sisnr_baseline = compute_objectives(stack([mix] * n_spk), targets) (train_wsj0mix.py:551-557)
//     sisnr_i        = sisnr - sisnr_baseline                            (:558)
// SI-SNR itself is cal_si_snr of baseline/avse2/utils/dnn.py:15-57 (EPS 1e-8, zero-mean over time, the in-repo
// statement of the speechbrain loss): with a = est - mean(est), s = src - mean(src),
//     proj = <a, s> s / (|s|^2 + EPS),  e = a - proj,  si_snr = 10 log10(|proj|^2 / (|e|^2 + EPS) + EPS).
// Everything is a function of first and second moments, so one streaming pass over est / src / mix suffices:
// stage 1 accumulates 4n + n^2 + n + 2 sums per (utterance, time chunk) in fp64 (HBM-bound: 4 (2n + 1) bytes per sample),
// stage 2 adds the chunks in order (bit-reproducible), forms the n x n pair matrix, picks the assignment with the largest
// mean SI-SNR out of the n! (PitWrapper minimises the mean loss = -SI-SNR) and the mixture baseline.
#include "mtn_ptx.cuh"
#include "mtn_host.h"

namespace mtn {

constexpr int SCORE_CHUNK = 32768;  // samples per stage-1 CTA
constexpr int SCORE_MAX_SPK = 4;    // n! <= 24 permutations; the reference's recipes use 2 or 3 (mambatasnet_S.yaml:39)

// Sums per (utterance, chunk), NS = number of speakers, row of SCORE_ROW(NS) doubles:
//   [0, NS) e_i | [NS, 2NS) s_j | [2NS, 3NS) e_i e_i | [3NS, 4NS) s_j s_j | [4NS, 4NS + NS^2) e_i s_j |
//   then m, m m, m s_j (NS)
__host__ __device__ constexpr int score_row(int ns) { return 4 * ns + ns * ns + 2 + ns; }

template <int NS>
__global__ void __launch_bounds__(256)
score_partial_kernel(const float* __restrict__ est, const float* __restrict__ src, const float* __restrict__ mix,
                     int ld_mix, int T, int nchunks, double* __restrict__ partial) {
    constexpr int ROW = score_row(NS);
    const int b = blockIdx.y, c = blockIdx.x;
    const int t0 = c * SCORE_CHUNK, t1 = min(T, t0 + SCORE_CHUNK);
    const float* e = est + size_t(b) * T * NS;
    const float* s = src + size_t(b) * T * NS;
    const float* m = mix + size_t(b) * ld_mix;
    double acc[ROW];
#pragma unroll
    for (int i = 0; i < ROW; ++i) acc[i] = 0.0;
    for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
        double ev[NS], sv[NS];
        if (NS == 2) {   // interleaved pairs: one 8-byte load each
            const float2 e2 = reinterpret_cast<const float2*>(e)[t], s2 = reinterpret_cast<const float2*>(s)[t];
            ev[0] = e2.x; ev[NS - 1] = e2.y; sv[0] = s2.x; sv[NS - 1] = s2.y;
        } else {
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                ev[i] = e[size_t(t) * NS + i];
                sv[i] = s[size_t(t) * NS + i];
            }
        }
        const double mm = m[t];   // products of fp32 values are exact only as fp64: promote before multiplying
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            acc[i] += ev[i];
            acc[NS + i] += sv[i];
            acc[2 * NS + i] = fma(ev[i], ev[i], acc[2 * NS + i]);
            acc[3 * NS + i] = fma(sv[i], sv[i], acc[3 * NS + i]);
#pragma unroll
            for (int j = 0; j < NS; ++j) acc[4 * NS + i * NS + j] = fma(ev[i], sv[j], acc[4 * NS + i * NS + j]);
            acc[4 * NS + NS * NS + 2 + i] = fma(mm, sv[i], acc[4 * NS + NS * NS + 2 + i]);
        }
        acc[4 * NS + NS * NS] += mm;
        acc[4 * NS + NS * NS + 1] = fma(mm, mm, acc[4 * NS + NS * NS + 1]);
    }
    __shared__ double red[8][ROW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < ROW; ++i) {
        double v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < ROW) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        partial[(size_t(b) * nchunks + c) * ROW + threadIdx.x] = v;
    }
}

__device__ __forceinline__ double si_snr_from_moments(double sa, double saa, double ss, double sss, double sas, double n) {
    const double EPS = 1e-8;
    const double dot = sas - sa * ss / n;          // <a, s>, zero-mean
    const double es = sss - ss * ss / n;           // |s|^2
    const double ea = saa - sa * sa / n;           // |a|^2
    const double k = dot / (es + EPS);
    const double proj2 = k * k * es;               // |proj|^2
    double noise2 = ea - 2.0 * k * dot + proj2;    // |a - proj|^2
    noise2 = noise2 > 0.0 ? noise2 : 0.0;
    return 10.0 * log10(proj2 / (noise2 + EPS) + EPS);
}

// out row (out_stride floats): si_snr (PIT, mean over speakers), si_snr_i, permutation index (lexicographic rank of the
// assignment; for NS = 2: 0 direct, 1 swapped), baseline, pair matrix [est i][src j] (NS*NS), and -- when the row has room
// -- the assignment itself: NS entries, entry i = the source that estimate i is matched to.
template <int NS>
__global__ void score_final_kernel(const double* __restrict__ partial, int nchunks, int T, int batch,
                                   float* __restrict__ out, int out_stride) {
    constexpr int ROW = score_row(NS);
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    double s[ROW];
    for (int i = 0; i < ROW; ++i) s[i] = 0.0;
    for (int c = 0; c < nchunks; ++c)       // chunks in order: bit-reproducible
        for (int i = 0; i < ROW; ++i) s[i] += partial[(size_t(b) * nchunks + c) * ROW + i];
    const double n = T;
    double pair[NS][NS];
    double base = 0.0;
    for (int j = 0; j < NS; ++j) {
        for (int i = 0; i < NS; ++i)
            pair[i][j] = si_snr_from_moments(s[i], s[2 * NS + i], s[NS + j], s[3 * NS + j], s[4 * NS + i * NS + j], n);
        base += si_snr_from_moments(s[4 * NS + NS * NS], s[4 * NS + NS * NS + 1], s[NS + j], s[3 * NS + j],
                                    s[4 * NS + NS * NS + 2 + j], n);
    }
    base /= NS;
    // all NS! assignments in lexicographic order (PitWrapper minimises the mean loss = maximises the mean SI-SNR; ties keep
    // the first, like argmin)
    int perm[NS], bestp[NS];
    for (int i = 0; i < NS; ++i) perm[i] = bestp[i] = i;
    double best = -1e300;
    int besti = 0, idx = 0;
    for (;;) {
        double tot = 0.0;
        for (int i = 0; i < NS; ++i) tot += pair[i][perm[i]];
        tot /= NS;
        if (tot > best) {
            best = tot;
            besti = idx;
            for (int i = 0; i < NS; ++i) bestp[i] = perm[i];
        }
        ++idx;
        int k = NS - 2;                      // next lexicographic permutation
        while (k >= 0 && perm[k] > perm[k + 1]) --k;
        if (k < 0) break;
        int l = NS - 1;
        while (perm[l] < perm[k]) --l;
        int t = perm[k]; perm[k] = perm[l]; perm[l] = t;
        for (int lo = k + 1, hi = NS - 1; lo < hi; ++lo, --hi) { t = perm[lo]; perm[lo] = perm[hi]; perm[hi] = t; }
    }
    float* o = out + size_t(b) * out_stride;
    o[0] = float(best);
    o[1] = float(best - base);
    o[2] = float(besti);
    o[3] = float(base);
    for (int i = 0; i < NS; ++i)
        for (int j = 0; j < NS; ++j) o[4 + i * NS + j] = float(pair[i][j]);
    if (out_stride >= 4 + NS * NS + NS)
        for (int i = 0; i < NS; ++i) o[4 + NS * NS + i] = float(bestp[i]);
}

template <int NS>
static int launch_score(const float* est, const float* src, const float* mix, int ld_mix, int batch, int T, double* work,
                        float* out, int out_stride, cudaStream_t s) {
    const int nchunks = (T + SCORE_CHUNK - 1) / SCORE_CHUNK;
    score_partial_kernel<NS><<<dim3(nchunks, batch), 256, 0, s>>>(est, src, mix, ld_mix, T, nchunks, work);
    MTN_CUDA_LAUNCH_CHECK("si_snr_pit(partial)");
    score_final_kernel<NS><<<(batch + 127) / 128, 128, 0, s>>>(work, nchunks, T, batch, out, out_stride);
    MTN_CUDA_LAUNCH_CHECK("si_snr_pit(final)");
    return MTN_OK;
}

}  // namespace mtn

extern "C" size_t mtn_si_snr_workspace_bytes_n(int batch, int T, int n_spk) {
    const size_t nchunks = (size_t(T) + mtn::SCORE_CHUNK - 1) / mtn::SCORE_CHUNK;
    return size_t(batch) * nchunks * mtn::score_row(n_spk) * sizeof(double);
}
extern "C" size_t mtn_si_snr_workspace_bytes(int batch, int T) { return mtn_si_snr_workspace_bytes_n(batch, T, 2); }

extern "C" int mtn_si_snr_pit_n_fwd(const float* est, const float* src, const float* mix, int ld_mix, int batch, int T,
                                    int n_spk, void* workspace, size_t workspace_bytes, float* out, int out_stride,
                                    mtn_stream_t stream) {
    using namespace mtn;
    MTN_REQUIRE(est && src && mix && workspace && out, "si_snr_pit: null pointer");
    MTN_REQUIRE(n_spk >= 1 && n_spk <= SCORE_MAX_SPK, "si_snr_pit: n_spk=%d (1..%d)", n_spk, SCORE_MAX_SPK);
    MTN_REQUIRE(batch > 0 && batch <= 65535 && T > 0 && ld_mix >= T, "si_snr_pit: bad batch=%d T=%d ld_mix=%d", batch, T,
                ld_mix);
    MTN_REQUIRE(out_stride >= 4 + n_spk * n_spk, "si_snr_pit: out_stride=%d < %d", out_stride, 4 + n_spk * n_spk);
    MTN_REQUIRE((reinterpret_cast<uintptr_t>(est) & 7) == 0 && (reinterpret_cast<uintptr_t>(src) & 7) == 0 &&
                    (reinterpret_cast<uintptr_t>(workspace) & 7) == 0,
                "si_snr_pit: est / src / workspace must be 8-byte aligned");
    MTN_REQUIRE(workspace_bytes >= mtn_si_snr_workspace_bytes_n(batch, T, n_spk), "si_snr_pit: workspace too small");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    double* w = reinterpret_cast<double*>(workspace);
    switch (n_spk) {
        case 1: return launch_score<1>(est, src, mix, ld_mix, batch, T, w, out, out_stride, s);
        case 2: return launch_score<2>(est, src, mix, ld_mix, batch, T, w, out, out_stride, s);
        case 3: return launch_score<3>(est, src, mix, ld_mix, batch, T, w, out, out_stride, s);
        default: return launch_score<4>(est, src, mix, ld_mix, batch, T, w, out, out_stride, s);
    }
}

extern "C" int mtn_si_snr_pit_fwd(const float* est, const float* src, const float* mix, int ld_mix, int batch, int T,
                                  void* workspace, size_t workspace_bytes, float* out, mtn_stream_t stream) {
    return mtn_si_snr_pit_n_fwd(est, src, mix, ld_mix, batch, T, 2, workspace, workspace_bytes, out, 8, stream);
}
