// DPMamba (dual-path) glue kernels, sm_100a: GroupNorm(1, C) over a whole utterance, 50 %-overlap segmentation and
// overlap-add, the intra <-> inter row transposition, PReLU, bias + operand planes, tanh * sigmoid output gate.
// Everything is channel-last fp32 rows of C channels; every access is a 128-bit vector inside a contiguous row, so a warp
// always touches >= 512 contiguous bytes.  All of it is HBM-bound streaming work.
//
// Reference (paths relative to /root/reference): the model is speechbrain 1.0.0 `lobes.models.dual_path.Dual_Path_Model`
// [third party, not vendored] as instantiated by Mamba-TasNet/hparams/WSJ0Mix/dpmamba_*.yaml:164-174; its forward is
// restated line by line in the vendored subclass Mamba-TasNet/modules/dual_path.py:56-150 (identical for skip_n_block = 0).
// Layout used here: the 4-D tensor [B, N, K, S] of the reference is kept as rows (b, s, k) x N channels ("layout A",
// what the intra model consumes, dual_path.py / speechbrain Dual_Computation_Block `x.permute(0,3,2,1).view(B*S,K,N)`); the
// inter model consumes rows (b, k, s) ("layout T", `permute(0,2,3,1).view(B*K,S,N)`).
#include "mtn_ptx.cuh"
#include "mtn_host.h"

namespace mtn {

constexpr int GN_THREADS = 256;
constexpr int GN_MAX_BLK = 64;

static int gn_nblk(size_t elems) {
    size_t n = (elems + 32767) / 32768;
    return n < 1 ? 1 : (n > GN_MAX_BLK ? GN_MAX_BLK : int(n));
}

// partial (sum, sum of squares) of utterance b's elements, slice blockIdx.x of gridDim.x: fp32 per thread over <= a few
// hundred elements, fp64 across the block.  Deterministic: fixed slices, fixed reduction tree, plain stores.
__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(const float* __restrict__ x, double2* __restrict__ partials, size_t elems_per_b) {
    const int b = blockIdx.y;
    const size_t n4 = elems_per_b / 4;
    const size_t per = (n4 + gridDim.x - 1) / gridDim.x;
    const size_t lo = per * blockIdx.x, hi = lo + per < n4 ? lo + per : n4;
    const float4* p = reinterpret_cast<const float4*>(x + size_t(b) * elems_per_b);
    double s = 0.0, q = 0.0;
    for (size_t i0 = lo; i0 < hi; i0 += size_t(GN_THREADS) * 64) {   // fp32 runs of <= 64 float4, then promoted
        float fs = 0.f, fq = 0.f;
        const size_t i1 = i0 + size_t(GN_THREADS) * 64 < hi ? i0 + size_t(GN_THREADS) * 64 : hi;
        for (size_t i = i0 + threadIdx.x; i < i1; i += GN_THREADS) {
            const float4 v = p[i];
            fs += (v.x + v.y) + (v.z + v.w);
            fq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, fq))));
        }
        s += double(fs);
        q += double(fq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    __shared__ double ss[GN_THREADS / 32], sq[GN_THREADS / 32];
    if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = s; sq[threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int w = 0; w < GN_THREADS / 32; ++w) { a += ss[w]; c += sq[w]; }
        partials[size_t(b) * gridDim.x + blockIdx.x] = make_double2(a, c);
    }
}

struct GnApplyParams {
    const float* x;
    const double2* partials;
    const float* w;
    const float* bias;
    const float* skip;
    const float* blend;      // nullable: result = 0.5 * result + 0.5 * blend (Dual_Path_Model_Skip, dual_path.py:114-116)
    float* out_a;
    float* out_a2;
    float* out_t;
    __nv_bfloat16* planes;
    size_t plane_stride;
    int S, K, C, nblk, x_transposed;
    float eps;
};

// y[b,s,k,:] = (x[src] - mean_b) * rstd_b * w + bias (+ skip[b,s,k,:]); src = (b,k,s) when x_transposed.
// Written to any of: out_a / out_a2 (rows (b,s,k)), out_t (rows (b,k,s)), operand planes (rows (b,s,k)).
template <int P>
__global__ void __launch_bounds__(256)
gn_apply_kernel(GnApplyParams p) {
    const int b = blockIdx.y;
    __shared__ float s_mean, s_rstd;
    if (threadIdx.x < 32) {
        double s = 0.0, q = 0.0;
        for (int i = threadIdx.x; i < p.nblk; i += 32) {
            const double2 v = p.partials[size_t(b) * p.nblk + i];
            s += v.x;
            q += v.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (threadIdx.x == 0) {
            const double n = double(p.S) * p.K * p.C;
            const double mean = s / n;
            double var = q / n - mean * mean;       // biased variance, like torch GroupNorm
            if (var < 0.0) var = 0.0;
            s_mean = float(mean);
            s_rstd = float(1.0 / sqrt(var + double(p.eps)));
        }
    }
    __syncthreads();
    const float mean = s_mean, rstd = s_rstd;
    const int c4n = p.C / 4;
    const size_t rows = size_t(p.S) * p.K;
    const size_t items = rows * c4n;
    const size_t base_b = size_t(b) * rows;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < items; i += size_t(gridDim.x) * blockDim.x) {
        const size_t r = i / c4n;
        const int c = int(i - r * c4n) * 4;
        const int s = int(r / p.K), k = int(r - size_t(s) * p.K);
        const size_t row_a = base_b + r;
        const size_t row_t = base_b + size_t(k) * p.S + s;
        const float4 v = *reinterpret_cast<const float4*>(p.x + (p.x_transposed ? row_t : row_a) * p.C + c);
        const float4 g = *reinterpret_cast<const float4*>(p.w + c);
        const float4 bb = *reinterpret_cast<const float4*>(p.bias + c);
        float4 o;
        o.x = fmaf((v.x - mean) * rstd, g.x, bb.x);
        o.y = fmaf((v.y - mean) * rstd, g.y, bb.y);
        o.z = fmaf((v.z - mean) * rstd, g.z, bb.z);
        o.w = fmaf((v.w - mean) * rstd, g.w, bb.w);
        if (p.skip) {
            const float4 k4 = *reinterpret_cast<const float4*>(p.skip + row_a * p.C + c);
            o.x += k4.x; o.y += k4.y; o.z += k4.z; o.w += k4.w;
        }
        if (p.blend) {
            const float4 r4 = *reinterpret_cast<const float4*>(p.blend + row_a * p.C + c);
            o.x = 0.5f * o.x + 0.5f * r4.x; o.y = 0.5f * o.y + 0.5f * r4.y;
            o.z = 0.5f * o.z + 0.5f * r4.z; o.w = 0.5f * o.w + 0.5f * r4.w;
        }
        if (p.out_a) *reinterpret_cast<float4*>(p.out_a + row_a * p.C + c) = o;
        if (p.out_a2) *reinterpret_cast<float4*>(p.out_a2 + row_a * p.C + c) = o;
        if (p.out_t) *reinterpret_cast<float4*>(p.out_t + row_t * p.C + c) = o;
        if (p.planes) store_planes4<P>(p.planes, p.plane_stride, row_a * p.C + c, o);
    }
}

// Same values as gn_apply_kernel, one warp per row, fused with the Add -> RMSNorm that opens the next stack
// (Block.forward with residual None, bimamba.py:446-447): the row is written to out_a (the block's residual copy), to the
// next stack's residual stream `res_next` and, RMS-normalised with that stack's first norm weight, to its operand planes --
// both at the row position the next stack uses (transposed for the inter model).  Replaces gn_apply's second output and the
// next stack's first add_rmsnorm launch (5 row transfers -> 3).
template <int P, int NV>   // C = 128 * NV
__global__ void __launch_bounds__(256)
gn_apply_norm_kernel(GnApplyParams p, float* __restrict__ res_next, __nv_bfloat16* __restrict__ xn_next,
                     const float* __restrict__ g_next, int next_transposed, float rms_eps) {
    constexpr int C = 128 * NV;
    const int b = blockIdx.y;
    __shared__ float s_mean, s_rstd;
    if (threadIdx.x < 32) {
        double s = 0.0, q = 0.0;
        for (int i = threadIdx.x; i < p.nblk; i += 32) {
            const double2 v = p.partials[size_t(b) * p.nblk + i];
            s += v.x;
            q += v.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (threadIdx.x == 0) {
            const double n = double(p.S) * p.K * C;
            const double mean = s / n;
            double var = q / n - mean * mean;
            if (var < 0.0) var = 0.0;
            s_mean = float(mean);
            s_rstd = float(1.0 / sqrt(var + double(p.eps)));
        }
    }
    __syncthreads();
    const float mean = s_mean, rstd = s_rstd;
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const size_t rows = size_t(p.S) * p.K;
    const size_t base_b = size_t(b) * rows;
    for (size_t r = size_t(blockIdx.x) * wpb + (threadIdx.x >> 5); r < rows; r += size_t(gridDim.x) * wpb) {
        const int s = int(r / p.K), k = int(r - size_t(s) * p.K);
        const size_t row_a = base_b + r;
        const size_t row_t = base_b + size_t(k) * p.S + s;
        const size_t row_x = p.x_transposed ? row_t : row_a;
        const size_t row_n = next_transposed ? row_t : row_a;
        float4 o[NV];
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = 128 * j + 4 * lane;
            const float4 v = *reinterpret_cast<const float4*>(p.x + row_x * C + c);
            const float4 g = *reinterpret_cast<const float4*>(p.w + c);
            const float4 bb = *reinterpret_cast<const float4*>(p.bias + c);
            float4 t;
            t.x = fmaf((v.x - mean) * rstd, g.x, bb.x);
            t.y = fmaf((v.y - mean) * rstd, g.y, bb.y);
            t.z = fmaf((v.z - mean) * rstd, g.z, bb.z);
            t.w = fmaf((v.w - mean) * rstd, g.w, bb.w);
            if (p.skip) {
                const float4 k4 = *reinterpret_cast<const float4*>(p.skip + row_a * C + c);
                t.x += k4.x; t.y += k4.y; t.z += k4.z; t.w += k4.w;
            }
            if (p.blend) {
                const float4 r4 = *reinterpret_cast<const float4*>(p.blend + row_a * C + c);
                t.x = 0.5f * t.x + 0.5f * r4.x; t.y = 0.5f * t.y + 0.5f * r4.y;
                t.z = 0.5f * t.z + 0.5f * r4.z; t.w = 0.5f * t.w + 0.5f * r4.w;
            }
            o[j] = t;
            sq += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;     // same order as add_rmsnorm_kernel
        }
        const float rr = rsqrtf(warp_sum(sq) * (1.0f / C) + rms_eps);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = 128 * j + 4 * lane;
            if (p.out_a) *reinterpret_cast<float4*>(p.out_a + row_a * C + c) = o[j];
            *reinterpret_cast<float4*>(res_next + row_n * C + c) = o[j];
            const float4 gg = *reinterpret_cast<const float4*>(g_next + c);
            float4 n4;
            n4.x = o[j].x * rr * gg.x;
            n4.y = o[j].y * rr * gg.y;
            n4.z = o[j].z * rr * gg.z;
            n4.w = o[j].w * rr * gg.w;
            store_planes4<P>(xn_next, p.plane_stride, row_n * C + c, n4);
        }
    }
}

// speechbrain Dual_Path_Model._padding + _Segmentation: chunk s, offset k <- frame l = s*(K/2) + k - K/2 (zero outside [0, L))
__global__ void __launch_bounds__(256)
dp_segment_kernel(const float* __restrict__ x, float* __restrict__ out_a, float* __restrict__ out_a2, int L, int C, int K,
                  int S) {
    const int b = blockIdx.y;
    const int c4n = C / 4, Ph = K / 2;
    const size_t rows = size_t(S) * K;
    const size_t items = rows * c4n;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < items; i += size_t(gridDim.x) * blockDim.x) {
        const size_t r = i / c4n;
        const int c = int(i - r * c4n) * 4;
        const int s = int(r / K), k = int(r - size_t(s) * K);
        const int l = s * Ph + k - Ph;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (l >= 0 && l < L) v = *reinterpret_cast<const float4*>(x + (size_t(b) * L + l) * C + c);
        const size_t off = (size_t(b) * rows + r) * C + c;
        *reinterpret_cast<float4*>(out_a + off) = v;
        if (out_a2) *reinterpret_cast<float4*>(out_a2 + off) = v;
    }
}

__device__ __forceinline__ float prelu_f(float v, float a) { return v >= 0.f ? v : a * v; }

// PReLU (one shared slope, dual_path.py:112) then Dual_Path_Model._over_add: frame l = sum of the two chunks that cover
// padded position p = l + K/2: the even chunk 2*(p / K) at offset p % K and the odd chunk 2*((p - K/2) / K) + 1 at
// (p - K/2) % K.  Output as GEMM operand planes for the conv2d 1x1 that follows (moved behind the overlap-add: it is
// linear, so conv2d(over_add(x)) + 2*bias == over_add(conv2d(x) + bias), and half the rows).
template <int P>
__global__ void __launch_bounds__(256)
dp_overadd_prelu_kernel(const float* __restrict__ X, const float* __restrict__ prelu_w, __nv_bfloat16* __restrict__ planes,
                        size_t plane_stride, int L, int C, int K, int S) {
    const int b = blockIdx.y;
    const float a = prelu_w[0];
    const int c4n = C / 4, Ph = K / 2;
    const size_t items = size_t(L) * c4n;
    const size_t rows = size_t(S) * K;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < items; i += size_t(gridDim.x) * blockDim.x) {
        const int l = int(i / c4n);
        const int c = int(i - size_t(l) * c4n) * 4;
        const int p = l + Ph;
        const int s1 = 2 * (p / K), k1 = p % K;
        const int s2 = 2 * ((p - Ph) / K) + 1, k2 = (p - Ph) % K;
        const float4 v1 = *reinterpret_cast<const float4*>(X + (size_t(b) * rows + size_t(s1) * K + k1) * C + c);
        const float4 v2 = *reinterpret_cast<const float4*>(X + (size_t(b) * rows + size_t(s2) * K + k2) * C + c);
        float4 o;
        o.x = prelu_f(v1.x, a) + prelu_f(v2.x, a);
        o.y = prelu_f(v1.y, a) + prelu_f(v2.y, a);
        o.z = prelu_f(v1.z, a) + prelu_f(v2.z, a);
        o.w = prelu_f(v1.w, a) + prelu_f(v2.w, a);
        store_planes4<P>(planes, plane_stride, (size_t(b) * L + l) * C + c, o);
    }
}

// planes = x + bias_scale * bias  (conv2d bias, applied after the GEMM)
template <int P>
__global__ void __launch_bounds__(256)
bias_planes_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ bias, float bias_scale,
                   __nv_bfloat16* __restrict__ planes, size_t plane_stride, size_t rows, int C) {
    const int c4n = C / 4;
    const size_t items = rows * c4n;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < items; i += size_t(gridDim.x) * blockDim.x) {
        const size_t r = i / c4n;
        const int c = int(i - r * c4n) * 4;
        float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
        const float4 bb = *reinterpret_cast<const float4*>(bias + c);
        v.x = fmaf(bias_scale, bb.x, v.x);
        v.y = fmaf(bias_scale, bb.y, v.y);
        v.z = fmaf(bias_scale, bb.z, v.z);
        v.w = fmaf(bias_scale, bb.w, v.w);
        store_planes4<P>(planes, plane_stride, r * C + c, v);
    }
}

__device__ __forceinline__ float tanh_f(float x) {      // 1 - 2 / (exp(2x) + 1), saturates cleanly at +-1
    const float e = ex2_approx(2.885390081777927f * x);
    return 1.0f - 2.0f * rcp_approx(e + 1.0f);
}
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }

// og fp32 [rows][groups][2*D] = (output pre-activation | gate pre-activation) per speaker group ->
// planes [P][rows][groups*D] = tanh(o + bo) * sigmoid(g + bg)     (dual_path.py:133)
template <int P>
__global__ void __launch_bounds__(256)
gate_planes_kernel(const float* __restrict__ og, const float* __restrict__ bo, const float* __restrict__ bg,
                   __nv_bfloat16* __restrict__ planes, size_t plane_stride, size_t rows, int groups, int D) {
    const int d4n = D / 4;
    const size_t items = rows * groups * d4n;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < items; i += size_t(gridDim.x) * blockDim.x) {
        const size_t rg = i / d4n;                 // row * groups + g
        const int d = int(i - rg * d4n) * 4;
        const float4 o = *reinterpret_cast<const float4*>(og + rg * 2 * D + d);
        const float4 g = *reinterpret_cast<const float4*>(og + rg * 2 * D + D + d);
        const float4 b1 = *reinterpret_cast<const float4*>(bo + d);
        const float4 b2 = *reinterpret_cast<const float4*>(bg + d);
        float4 v;
        v.x = tanh_f(o.x + b1.x) * sigmoid_f(g.x + b2.x);
        v.y = tanh_f(o.y + b1.y) * sigmoid_f(g.y + b2.y);
        v.z = tanh_f(o.z + b1.z) * sigmoid_f(g.z + b2.z);
        v.w = tanh_f(o.w + b1.w) * sigmoid_f(g.w + b2.w);
        store_planes4<P>(planes, plane_stride, rg * D + d, v);
    }
}

static int dp_grid(size_t items, int waves = 8) {
    size_t need = (items + 255) / 256;
    size_t cap = size_t(num_sms()) * waves;
    size_t g = need < cap ? need : cap;
    return g < 1 ? 1 : int(g);
}

}  // namespace mtn

using namespace mtn;

extern "C" size_t mtn_gn_partials_bytes(int batch, int rows, int C) {
    if (batch <= 0 || rows <= 0 || C <= 0) return 0;
    return size_t(batch) * gn_nblk(size_t(rows) * C) * sizeof(double2);
}

extern "C" int mtn_gn_stats_fwd(const float* x, void* partials, int batch, int rows, int C, mtn_stream_t stream) {
    MTN_REQUIRE(x && partials, "gn_stats: null pointer");
    MTN_REQUIRE(batch > 0 && batch <= 65535 && rows > 0 && C > 0 && C % 4 == 0, "gn_stats: bad shape batch=%d rows=%d C=%d",
                batch, rows, C);
    MTN_REQUIRE((reinterpret_cast<uintptr_t>(partials) & 15) == 0, "gn_stats: partials must be 16-byte aligned");
    const size_t elems = size_t(rows) * C;
    dim3 grid(gn_nblk(elems), batch);
    gn_stats_kernel<<<grid, GN_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        x, reinterpret_cast<double2*>(partials), elems);
    MTN_CUDA_LAUNCH_CHECK("gn_stats");
    return MTN_OK;
}

extern "C" int mtn_gn_apply_fwd(const mtn_gn_apply_args* a, mtn_stream_t stream) {
    MTN_REQUIRE(a && a->x && a->partials && a->w && a->bias, "gn_apply: null pointer");
    MTN_REQUIRE(a->out_a || a->out_a2 || a->out_t || a->planes, "gn_apply: no output requested");
    MTN_REQUIRE(a->batch > 0 && a->batch <= 65535 && a->S > 0 && a->K > 0 && a->C > 0 && a->C % 4 == 0,
                "gn_apply: bad shape batch=%d S=%d K=%d C=%d", a->batch, a->S, a->K, a->C);
    MTN_REQUIRE(!a->planes || a->n_planes == 1 || a->n_planes == 2, "gn_apply: n_planes=%d", a->n_planes);
    GnApplyParams p;
    p.x = a->x; p.partials = reinterpret_cast<const double2*>(a->partials); p.w = a->w; p.bias = a->bias; p.skip = a->skip;
    p.blend = a->blend;
    p.out_a = a->out_a; p.out_a2 = a->out_a2; p.out_t = a->out_t;
    p.planes = reinterpret_cast<__nv_bfloat16*>(a->planes);
    const size_t rows = size_t(a->S) * a->K;
    p.plane_stride = (a->plane_rows > 0 ? size_t(a->plane_rows) : size_t(a->batch) * rows) * a->C;
    p.S = a->S; p.K = a->K; p.C = a->C; p.nblk = gn_nblk(rows * a->C); p.x_transposed = a->x_transposed; p.eps = a->eps;
    const size_t items = rows * (a->C / 4);
    size_t gx = (items + 255) / 256;
    const size_t cap = size_t(num_sms()) * 8 / a->batch + 1;
    if (gx > cap) gx = cap;
    dim3 grid((unsigned)gx, a->batch);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (a->planes && a->n_planes == 2) gn_apply_kernel<2><<<grid, 256, 0, s>>>(p);
    else gn_apply_kernel<1><<<grid, 256, 0, s>>>(p);
    MTN_CUDA_LAUNCH_CHECK("gn_apply");
    return MTN_OK;
}

extern "C" int mtn_gn_apply_norm_fwd(const mtn_gn_apply_args* a, float* res_next, void* xn_next_planes, const float* g_next,
                                     int next_transposed, float rms_eps, mtn_stream_t stream) {
    MTN_REQUIRE(a && a->x && a->partials && a->w && a->bias && res_next && xn_next_planes && g_next,
                "gn_apply_norm: null pointer");
    MTN_REQUIRE(a->batch > 0 && a->batch <= 65535 && a->S > 0 && a->K > 0, "gn_apply_norm: bad shape");
    MTN_REQUIRE(a->C == 128 || a->C == 256 || a->C == 512, "gn_apply_norm: C=%d (supported: 128, 256, 512)", a->C);
    MTN_REQUIRE(a->n_planes == 1 || a->n_planes == 2, "gn_apply_norm: n_planes=%d", a->n_planes);
    MTN_REQUIRE(!a->out_a2 && !a->out_t && !a->planes, "gn_apply_norm: only out_a is supported beside the fused outputs");
    GnApplyParams p;
    p.x = a->x; p.partials = reinterpret_cast<const double2*>(a->partials); p.w = a->w; p.bias = a->bias; p.skip = a->skip;
    p.blend = a->blend;
    p.out_a = a->out_a; p.out_a2 = nullptr; p.out_t = nullptr; p.planes = nullptr;
    const size_t rows = size_t(a->S) * a->K;
    p.plane_stride = (a->plane_rows > 0 ? size_t(a->plane_rows) : size_t(a->batch) * rows) * a->C;
    p.S = a->S; p.K = a->K; p.C = a->C; p.nblk = gn_nblk(rows * a->C); p.x_transposed = a->x_transposed; p.eps = a->eps;
    size_t gx = (rows + 7) / 8;
    const size_t cap = size_t(num_sms()) * 8 / a->batch + 1;
    if (gx > cap) gx = cap;
    dim3 grid((unsigned)gx, a->batch);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    __nv_bfloat16* xn = reinterpret_cast<__nv_bfloat16*>(xn_next_planes);
#define MTN_GNN(PP, NV) gn_apply_norm_kernel<PP, NV><<<grid, 256, 0, s>>>(p, res_next, xn, g_next, next_transposed, rms_eps)
    if (a->n_planes == 2) {
        if (a->C == 128) MTN_GNN(2, 1); else if (a->C == 256) MTN_GNN(2, 2); else MTN_GNN(2, 4);
    } else {
        if (a->C == 128) MTN_GNN(1, 1); else if (a->C == 256) MTN_GNN(1, 2); else MTN_GNN(1, 4);
    }
#undef MTN_GNN
    MTN_CUDA_LAUNCH_CHECK("gn_apply_norm");
    return MTN_OK;
}

extern "C" int mtn_dp_num_chunks(int L, int K) {
    if (L <= 0 || K <= 0 || K % 2) return -1;
    const int P = K / 2;
    const int gap = K - (P + L % K) % K;         // speechbrain Dual_Path_Model._padding
    return 2 * ((L + gap + P) / K);              // chunks of K at stride K/2 over the padded length L + gap + 2P
}

extern "C" int mtn_dp_segment_fwd(const float* x, float* out_a, float* out_a2, int batch, int L, int C, int K, int S,
                                  mtn_stream_t stream) {
    MTN_REQUIRE(x && out_a, "dp_segment: null pointer");
    MTN_REQUIRE(batch > 0 && batch <= 65535 && L > 0 && C > 0 && C % 4 == 0 && K > 0 && K % 2 == 0, "dp_segment: bad shape");
    MTN_REQUIRE(S == mtn_dp_num_chunks(L, K), "dp_segment: S=%d but L=%d K=%d gives %d chunks", S, L, K,
                mtn_dp_num_chunks(L, K));
    const size_t items = size_t(S) * K * (C / 4);
    dim3 grid(dp_grid(items, 8) / (batch < 8 ? 1 : 8) + 1, batch);
    dp_segment_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, out_a, out_a2, L, C, K, S);
    MTN_CUDA_LAUNCH_CHECK("dp_segment");
    return MTN_OK;
}

extern "C" int mtn_dp_overadd_prelu_fwd(const float* X, const float* prelu_w, void* planes, int plane_rows, int batch, int L,
                                        int C, int K, int S, int n_planes, mtn_stream_t stream) {
    MTN_REQUIRE(X && prelu_w && planes, "dp_overadd: null pointer");
    MTN_REQUIRE(batch > 0 && batch <= 65535 && L > 0 && C > 0 && C % 4 == 0 && K > 0 && K % 2 == 0, "dp_overadd: bad shape");
    MTN_REQUIRE(S == mtn_dp_num_chunks(L, K), "dp_overadd: S=%d but L=%d K=%d gives %d chunks", S, L, K,
                mtn_dp_num_chunks(L, K));
    MTN_REQUIRE(n_planes == 1 || n_planes == 2, "dp_overadd: n_planes=%d", n_planes);
    MTN_REQUIRE(plane_rows >= batch * L, "dp_overadd: plane_rows=%d < batch*L", plane_rows);
    const size_t items = size_t(L) * (C / 4);
    dim3 grid(dp_grid(items, 8) / (batch < 8 ? 1 : 8) + 1, batch);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(planes);
    const size_t stride = size_t(plane_rows) * C;
    if (n_planes == 2) dp_overadd_prelu_kernel<2><<<grid, 256, 0, s>>>(X, prelu_w, pl, stride, L, C, K, S);
    else dp_overadd_prelu_kernel<1><<<grid, 256, 0, s>>>(X, prelu_w, pl, stride, L, C, K, S);
    MTN_CUDA_LAUNCH_CHECK("dp_overadd_prelu");
    return MTN_OK;
}

extern "C" int mtn_bias_planes_fwd(const float* x, int ldx, const float* bias, float bias_scale, void* planes, int plane_rows,
                                   int rows, int C, int n_planes, mtn_stream_t stream) {
    MTN_REQUIRE(x && bias && planes, "bias_planes: null pointer");
    MTN_REQUIRE(rows > 0 && C > 0 && C % 4 == 0 && ldx >= C && ldx % 4 == 0 && plane_rows >= rows, "bias_planes: bad shape");
    MTN_REQUIRE(n_planes == 1 || n_planes == 2, "bias_planes: n_planes=%d", n_planes);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(planes);
    const int grid = dp_grid(size_t(rows) * (C / 4));
    const size_t stride = size_t(plane_rows) * C;
    if (n_planes == 2) bias_planes_kernel<2><<<grid, 256, 0, s>>>(x, ldx, bias, bias_scale, pl, stride, size_t(rows), C);
    else bias_planes_kernel<1><<<grid, 256, 0, s>>>(x, ldx, bias, bias_scale, pl, stride, size_t(rows), C);
    MTN_CUDA_LAUNCH_CHECK("bias_planes");
    return MTN_OK;
}

extern "C" int mtn_gate_planes_fwd(const float* og, const float* bo, const float* bg, void* planes, int plane_rows, int rows,
                                   int groups, int D, int n_planes, mtn_stream_t stream) {
    MTN_REQUIRE(og && bo && bg && planes, "gate_planes: null pointer");
    MTN_REQUIRE(rows > 0 && groups > 0 && D > 0 && D % 4 == 0 && plane_rows >= rows, "gate_planes: bad shape");
    MTN_REQUIRE(n_planes == 1 || n_planes == 2, "gate_planes: n_planes=%d", n_planes);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(planes);
    const int grid = dp_grid(size_t(rows) * groups * (D / 4));
    const size_t stride = size_t(plane_rows) * groups * D;
    if (n_planes == 2) gate_planes_kernel<2><<<grid, 256, 0, s>>>(og, bo, bg, pl, stride, size_t(rows), groups, D);
    else gate_planes_kernel<1><<<grid, 256, 0, s>>>(og, bo, bg, pl, stride, size_t(rows), groups, D);
    MTN_CUDA_LAUNCH_CHECK("gate_planes");
    return MTN_OK;
}
