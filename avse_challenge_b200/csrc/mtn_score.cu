// On-device SI-SNR with 2-speaker PIT and SI-SNR improvement over the mixture (sm_100a).
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
// stage 1 accumulates 15 sums per (utterance, time chunk) in fp64 (HBM-bound: 20 bytes per sample), stage 2 adds the
// chunks in order (bit-reproducible), forms the 2 x 2 pair matrix, picks the permutation with the larger mean SI-SNR
// (PitWrapper minimises the mean loss = -SI-SNR) and the mixture baseline.
#include "mtn_ptx.cuh"
#include "mtn_host.h"

namespace mtn {

constexpr int SCORE_CHUNK = 32768;  // samples per stage-1 CTA

// sums: 0 e0, 1 e1, 2 s0, 3 s1, 4 m, 5 e0e0, 6 e1e1, 7 s0s0, 8 s1s1, 9 mm, 10 e0s0, 11 e0s1, 12 e1s0, 13 e1s1,
//       14 ms0   (+ ms1 kept in slot 15 of the padded row)
__global__ void __launch_bounds__(256)
score_partial_kernel(const float* __restrict__ est, const float* __restrict__ src, const float* __restrict__ mix,
                     int ld_mix, int T, int nchunks, double* __restrict__ partial) {
    const int b = blockIdx.y, c = blockIdx.x;
    const int t0 = c * SCORE_CHUNK, t1 = min(T, t0 + SCORE_CHUNK);
    const float2* e2 = reinterpret_cast<const float2*>(est) + size_t(b) * T;
    const float2* s2 = reinterpret_cast<const float2*>(src) + size_t(b) * T;
    const float* m = mix + size_t(b) * ld_mix;
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.0;
    for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
        const float2 e = e2[t], s = s2[t];
        const float mv = m[t];
        // products in fp32 are exact enough only as fp64: promote before multiplying
        const double e0 = e.x, e1 = e.y, s0 = s.x, s1 = s.y, mm = mv;
        acc[0] += e0;
        acc[1] += e1;
        acc[2] += s0;
        acc[3] += s1;
        acc[4] += mm;
        acc[5] = fma(e0, e0, acc[5]);
        acc[6] = fma(e1, e1, acc[6]);
        acc[7] = fma(s0, s0, acc[7]);
        acc[8] = fma(s1, s1, acc[8]);
        acc[9] = fma(mm, mm, acc[9]);
        acc[10] = fma(e0, s0, acc[10]);
        acc[11] = fma(e0, s1, acc[11]);
        acc[12] = fma(e1, s0, acc[12]);
        acc[13] = fma(e1, s1, acc[13]);
        acc[14] = fma(mm, s0, acc[14]);
        acc[15] = fma(mm, s1, acc[15]);
    }
    __shared__ double red[8][16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        double v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        partial[(size_t(b) * nchunks + c) * 16 + threadIdx.x] = v;
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

// out [B][8]: si_snr (PIT), si_snr_i, permutation (0: est0->src0, 1: est0->src1), baseline, pair 00, 01, 10, 11
__global__ void score_final_kernel(const double* __restrict__ partial, int nchunks, int T, int batch,
                                   float* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    double s[16];
    for (int i = 0; i < 16; ++i) s[i] = 0.0;
    for (int c = 0; c < nchunks; ++c)
        for (int i = 0; i < 16; ++i) s[i] += partial[(size_t(b) * nchunks + c) * 16 + i];
    const double n = T;
    const double p00 = si_snr_from_moments(s[0], s[5], s[2], s[7], s[10], n);
    const double p01 = si_snr_from_moments(s[0], s[5], s[3], s[8], s[11], n);
    const double p10 = si_snr_from_moments(s[1], s[6], s[2], s[7], s[12], n);
    const double p11 = si_snr_from_moments(s[1], s[6], s[3], s[8], s[13], n);
    const double m0 = si_snr_from_moments(s[4], s[9], s[2], s[7], s[14], n);
    const double m1 = si_snr_from_moments(s[4], s[9], s[3], s[8], s[15], n);
    const double direct = 0.5 * (p00 + p11), swapped = 0.5 * (p01 + p10);
    const bool sw = swapped > direct;
    const double best = sw ? swapped : direct;
    const double base = 0.5 * (m0 + m1);
    float* o = out + size_t(b) * 8;
    o[0] = float(best);
    o[1] = float(best - base);
    o[2] = sw ? 1.f : 0.f;
    o[3] = float(base);
    o[4] = float(p00);
    o[5] = float(p01);
    o[6] = float(p10);
    o[7] = float(p11);
}

}  // namespace mtn

extern "C" size_t mtn_si_snr_workspace_bytes(int batch, int T) {
    const size_t nchunks = (size_t(T) + mtn::SCORE_CHUNK - 1) / mtn::SCORE_CHUNK;
    return size_t(batch) * nchunks * 16 * sizeof(double);
}

extern "C" int mtn_si_snr_pit_fwd(const float* est, const float* src, const float* mix, int ld_mix, int batch, int T,
                                  void* workspace, size_t workspace_bytes, float* out, mtn_stream_t stream) {
    using namespace mtn;
    MTN_REQUIRE(est && src && mix && workspace && out, "si_snr_pit: null pointer");
    MTN_REQUIRE(batch > 0 && batch <= 65535 && T > 0 && ld_mix >= T, "si_snr_pit: bad batch=%d T=%d ld_mix=%d", batch, T,
                ld_mix);
    MTN_REQUIRE((reinterpret_cast<uintptr_t>(est) & 7) == 0 && (reinterpret_cast<uintptr_t>(src) & 7) == 0 &&
                    (reinterpret_cast<uintptr_t>(workspace) & 7) == 0,
                "si_snr_pit: est / src / workspace must be 8-byte aligned");
    MTN_REQUIRE(workspace_bytes >= mtn_si_snr_workspace_bytes(batch, T), "si_snr_pit: workspace too small");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int nchunks = (T + SCORE_CHUNK - 1) / SCORE_CHUNK;
    score_partial_kernel<<<dim3(nchunks, batch), 256, 0, s>>>(est, src, mix, ld_mix, T, nchunks,
                                                              reinterpret_cast<double*>(workspace));
    MTN_CUDA_LAUNCH_CHECK("si_snr_pit(partial)");
    score_final_kernel<<<(batch + 127) / 128, 128, 0, s>>>(reinterpret_cast<const double*>(workspace), nchunks, T, batch,
                                                           out);
    MTN_CUDA_LAUNCH_CHECK("si_snr_pit(final)");
    return MTN_OK;
}
