// Chunk-state folding for the chunked / sequence-parallel selective scan (sm_100a).
//
// A long recording is cut into G time chunks (G = ranks x sub-chunks per rank).  A summary pass of the scan kernel
// (mtn_scan_fwd with y == NULL) gives, per chunk g, direction and channel, the chunk's affine transfer operator
//     h_out = exp2(A2 * sum_delta_g) * h_in + h_end_g            (h_end_g = final state for h_in = 0)
// which is exact because the recurrence h <- exp(delta_t A) h + delta_t B_t u_t is linear in h with a diagonal,
// time-varying decay (Mamba-TasNet/modules/mamba/selective_scan_interface.py:126-139).  This kernel composes the
// operators in time order (forward: g = 0..G-1, backward: g = G-1..0) and emits the state ENTERING each chunk of the
// caller's range [g0, g0 + n_out); a second scan pass seeded with those states then equals the unchunked scan.
// The reference has no chunked mode (its oracle has no initial-state argument, ssi.py:124); this is the
// B200-side answer to BASELINE config 5.  One thread per (direction, channel, state); G <= a few thousand.
#include "mtn_ptx.cuh"
#include "mtn_host.h"

namespace mtn {

__global__ void __launch_bounds__(256)
fold_states_kernel(const float* __restrict__ h_end, const float* __restrict__ sum_delta, const float* __restrict__ A2,
                   const float* __restrict__ h0, float* __restrict__ h_in, float* __restrict__ h_final, int G, int di,
                   int g0, int n_out, int dir0) {
    const int dir = dir0 + blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // d * 16 + n
    if (idx >= di * 16) return;
    const int d = idx >> 4;
    const float a2 = A2[size_t(dir) * di * 16 + idx];
    float h = h0 ? h0[size_t(dir) * di * 16 + idx] : 0.f;
    const float* he = h_end + size_t(dir) * G * di * 16 + idx;
    const float* sd = sum_delta + size_t(dir) * G * di + d;
    float* out = h_in + size_t(dir) * n_out * di * 16 + idx;
    for (int s = 0; s < G; ++s) {
        const int g = dir == 0 ? s : G - 1 - s;
        if (g >= g0 && g < g0 + n_out) out[size_t(g - g0) * di * 16] = h;
        h = fmaf(ex2_approx(a2 * sd[size_t(g) * di]), h, he[size_t(g) * di * 16]);
    }
    if (h_final) h_final[size_t(dir) * di * 16 + idx] = h;
}

// Same composition over summaries as they arrive from ONE all-gather: every rank contributes one packed record
// [h_end (2 x cmax x di x 16) | sum_delta (2 x cmax x di)] (cmax = sub-chunks per rank, padded with identity operators =
// zeros), so global chunk g = rank * cmax + c lives at pack[rank][dir][c].  Reading that layout in place removes the two
// transposing copies (and one of the two collectives) per layer of the sequence-parallel forward.
__global__ void __launch_bounds__(256)
fold_states_packed_kernel(const float* __restrict__ pack, const float* __restrict__ A2, const float* __restrict__ h0,
                          float* __restrict__ h_in, float* __restrict__ h_final, int W, int cmax, int di, int g0, int n_out,
                          int dir0) {
    const int dir = dir0 + blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // d * 16 + n
    if (idx >= di * 16) return;
    const int d = idx >> 4;
    const size_t rec = size_t(2) * cmax * di * 17;          // floats per rank record
    const size_t sd_off = size_t(2) * cmax * di * 16;
    const float a2 = A2[size_t(dir) * di * 16 + idx];
    float h = h0 ? h0[size_t(dir) * di * 16 + idx] : 0.f;
    float* out = h_in + size_t(dir) * n_out * di * 16 + idx;
    const int G = W * cmax;
    for (int s = 0; s < G; ++s) {
        const int g = dir == 0 ? s : G - 1 - s;
        const int w = g / cmax, c = g - w * cmax;
        const float* r = pack + size_t(w) * rec;
        if (g >= g0 && g < g0 + n_out) out[size_t(g - g0) * di * 16] = h;
        const float sd = r[sd_off + (size_t(dir) * cmax + c) * di + d];
        h = fmaf(ex2_approx(a2 * sd), h, r[(size_t(dir) * cmax + c) * di * 16 + idx]);
    }
    if (h_final) h_final[size_t(dir) * di * 16 + idx] = h;
}

}  // namespace mtn

extern "C" int mtn_fold_states_packed_fwd(const float* pack, const float* A2, const float* h0, float* h_in, float* h_final,
                                          int W, int cmax, int di, int g0, int n_out, int dir_mask, mtn_stream_t stream) {
    using namespace mtn;
    MTN_REQUIRE(pack && A2 && h_in, "fold_states_packed: null pointer");
    MTN_REQUIRE(W > 0 && cmax > 0 && di > 0 && g0 >= 0 && n_out > 0 && g0 + n_out <= W * cmax,
                "fold_states_packed: bad range W=%d cmax=%d g0=%d n_out=%d", W, cmax, g0, n_out);
    MTN_REQUIRE(dir_mask >= 1 && dir_mask <= 3, "fold_states_packed: dir_mask=%d", dir_mask);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int dir0 = (dir_mask & 1) ? 0 : 1;
    const int ndirs = dir_mask == 3 ? 2 : 1;
    dim3 grid((di * 16 + 255) / 256, ndirs);
    fold_states_packed_kernel<<<grid, 256, 0, s>>>(pack, A2, h0, h_in, h_final, W, cmax, di, g0, n_out, dir0);
    MTN_CUDA_LAUNCH_CHECK("fold_states_packed");
    return MTN_OK;
}

extern "C" int mtn_fold_states_fwd(const float* h_end, const float* sum_delta, const float* A2, const float* h0,
                                   float* h_in, float* h_final, int G, int di, int g0, int n_out, int dir_mask,
                                   mtn_stream_t stream) {
    using namespace mtn;
    MTN_REQUIRE(h_end && sum_delta && A2 && h_in, "fold_states: null pointer");
    MTN_REQUIRE(G > 0 && di > 0 && g0 >= 0 && n_out > 0 && g0 + n_out <= G, "fold_states: bad range G=%d g0=%d n_out=%d",
                G, g0, n_out);
    MTN_REQUIRE(dir_mask >= 1 && dir_mask <= 3, "fold_states: dir_mask=%d", dir_mask);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int dir0 = (dir_mask & 1) ? 0 : 1;
    const int ndirs = dir_mask == 3 ? 2 : 1;
    dim3 grid((di * 16 + 255) / 256, ndirs);
    fold_states_kernel<<<grid, 256, 0, s>>>(h_end, sum_delta, A2, h0, h_in, h_final, G, di, g0, n_out, dir0);
    MTN_CUDA_LAUNCH_CHECK("fold_states");
    return MTN_OK;
}
