// Micro-benchmark: per-sub-partition issue / pipe cost of the instructions the scan recurrence is made of (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
// One CTA on one SM, W warps per sub-partition (block = 128 * W threads); every test runs ITER iterations of an unrolled
// body of independent chains and reports cycles per warp-instruction per sub-partition (= 1 / throughput).
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
#define ITER 2000
__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d; asm volatile("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(*(uint64_t*)&d) : "l"(*(uint64_t*)&a), "l"(*(uint64_t*)&b), "l"(*(uint64_t*)&c)); return d; }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    float2 d; asm volatile("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(*(uint64_t*)&d) : "l"(*(uint64_t*)&a), "l"(*(uint64_t*)&b)); return d; }

template <int T>
__global__ void k(float* out, long long* cyc, float seed) {
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 0.001f * i;
    __syncthreads();
    float2 h[8], e[8], b[8];
    for (int q = 0; q < 8; ++q) { h[q] = make_float2(seed * q, seed); e[q] = make_float2(0.999f - 0.01f * q * seed, 0.998f - 0.013f * q * seed); b[q] = make_float2(1e-3f * threadIdx.x + q * seed, 1e-3f * q + seed); }
    float s = seed;
    float2 a2[8];
    for (int q = 0; q < 8; ++q) a2[q] = make_float2(-0.01f * (2 * q + 1) * seed, -0.01f * (2 * q + 2) * seed);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        if (T == 0) {          // 16 FFMA (3-reg, d = a*d + c)
#pragma unroll
            for (int q = 0; q < 8; ++q) { h[q].x = fmaf(e[q].x, h[q].x, b[q].x); h[q].y = fmaf(e[q].y, h[q].y, b[q].y); }
        } else if (T == 1) {   // 16 FFMA2: h = e*h + b  (two rounds of 8)
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < 8; ++q) h[q] = ffma2(e[q], h[q], b[q]);
        } else if (T == 2) {   // 16 FMUL2 pair * pair -> consumed by nothing but kept live via xor-free trick: h = h * e
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < 8; ++q) h[q] = fmul2(h[q], e[q]);
        } else if (T == 3) {   // 16 FMUL2 with a broadcast scalar operand: h = (s,s) * h
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < 8; ++q) h[q] = fmul2(make_float2(s, s), h[q]);
        } else if (T == 4) {   // 16 MUFU.EX2
#pragma unroll
            for (int q = 0; q < 8; ++q) { h[q].x = ex2a(h[q].x); h[q].y = ex2a(h[q].y); }
        } else if (T == 5 || T == 6 || T == 7 || T == 10 || T == 11 || T == 12 || T == 13) {   // one recurrence step's arithmetic: 8 FMUL2 (dl*A2) + 16 MUFU + 8 FMUL2 (du*B) + 8 FFMA2 (h) + 8 FFMA2 (y)
            float2 x[8], bu[8];
            float2 y0 = make_float2(0.f, 0.f), y1 = y0, y2 = y0, y3 = y0;
            float4 Bv[4], Cv[4];
            constexpr bool LD = (T == 6 || T == 7 || T >= 11);
            constexpr bool MU = (T < 10 || T >= 12);
            if (LD) {      // + 8 broadcast LDS.128 (B_t, C_t)
#pragma unroll
                for (int q = 0; q < 4; ++q) { Bv[q] = *reinterpret_cast<float4*>(&sm[(it & 15) * 32 + 4 * q]); Cv[q] = *reinterpret_cast<float4*>(&sm[(it & 15) * 32 + 16 + 4 * q]); }
            }
            if (T == 7 || T >= 11) s = sm[512 + (it & 15) * 32 + (threadIdx.x & 31)];   // + a lane-private LDS
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = fmul2(make_float2(s, s), e[q]);
#pragma unroll
            for (int q = 0; q < 8; ++q) { if (MU) { x[q].x = ex2a(x[q].x); x[q].y = ex2a(x[q].y); } }
#pragma unroll
            for (int q = 0; q < 8; ++q) bu[q] = fmul2(make_float2(s, s), LD ? ((q & 1) ? make_float2(Bv[q >> 1].z, Bv[q >> 1].w) : make_float2(Bv[q >> 1].x, Bv[q >> 1].y)) : b[q]);
#pragma unroll
            for (int q = 0; q < 8; ++q) h[q] = ffma2(x[q], h[q], bu[q]);
#pragma unroll
            for (int q = 0; q < 8; q += 4) {
                float2 c0 = LD ? make_float2(Cv[q >> 1].x, Cv[q >> 1].y) : b[q], c1 = LD ? make_float2(Cv[q >> 1].z, Cv[q >> 1].w) : b[q + 1];
                float2 c2 = LD ? make_float2(Cv[(q >> 1) + 1].x, Cv[(q >> 1) + 1].y) : b[q + 2], c3 = LD ? make_float2(Cv[(q >> 1) + 1].z, Cv[(q >> 1) + 1].w) : b[q + 3];
                y0 = ffma2(h[q], c0, y0); y1 = ffma2(h[q + 1], c1, y1); y2 = ffma2(h[q + 2], c2, y2); y3 = ffma2(h[q + 3], c3, y3);
            }
            if (T == 13) { const float2 sy = make_float2(y0.x + y1.x + y2.x + y3.x, y0.y + y1.y + y2.y + y3.y); sm[600 + (it & 15) * 32 + (threadIdx.x & 31)] = sy.x + sy.y; }
            s = s * 0.999f + (y0.x + y1.x + y2.x + y3.x + y0.y + y1.y + y2.y + y3.y) * 1e-30f;
        } else if (T == 15) {  // the same step, software-pipelined the way the kernel is: the decay factors of step it+1 (8 FMUL2 + 16 MUFU)
                               // do not depend on the state, so they are issued among the state updates of step it; two steps per
                               // iteration so that no register moves are needed.  Nothing feeds back from y into the next step.
            float4 Bv[4], Cv[4];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float2* cur = half ? b : e;     // e / b double as the (current, next) decay registers
                float2* nxt = half ? e : b;
                const int row = ((2 * it + half) & 15) * 32;
#pragma unroll
                for (int q = 0; q < 4; ++q) { Bv[q] = *reinterpret_cast<float4*>(&sm[row + 4 * q]); Cv[q] = *reinterpret_cast<float4*>(&sm[row + 16 + 4 * q]); }
                const float sn = sm[512 + row + (threadIdx.x & 31)];
                const float du = sm[768 + (row >> 1) + (threadIdx.x & 15)];
                float2 y0 = make_float2(0.f, 0.f), y1 = y0, y2 = y0, y3 = y0;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float2 x = fmul2(make_float2(sn, sn), a2[q]);
                    x.x = ex2a(x.x); x.y = ex2a(x.y);
                    nxt[q] = x;
                    const float2 bq = (q & 1) ? make_float2(Bv[q >> 1].z, Bv[q >> 1].w) : make_float2(Bv[q >> 1].x, Bv[q >> 1].y);
                    const float2 cq = (q & 1) ? make_float2(Cv[q >> 1].z, Cv[q >> 1].w) : make_float2(Cv[q >> 1].x, Cv[q >> 1].y);
                    h[q] = ffma2(cur[q], h[q], fmul2(make_float2(du, du), bq));
                    if ((q & 3) == 0) y0 = ffma2(h[q], cq, y0); else if ((q & 3) == 1) y1 = ffma2(h[q], cq, y1);
                    else if ((q & 3) == 2) y2 = ffma2(h[q], cq, y2); else y3 = ffma2(h[q], cq, y3);
                }
                const float2 sy = make_float2(y0.x + y1.x + y2.x + y3.x, y0.y + y1.y + y2.y + y3.y);
                sm[600 + ((row >> 1) & 127) + (threadIdx.x & 31)] = sy.x + sy.y;
            }
        } else if (T == 8) {   // FFMA2 with an immediate-like constant third operand: h = e*h + const
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < 8; ++q) h[q] = ffma2(e[q], h[q], make_float2(0.5f, 0.5f));
        } else if (T == 14) {
            float2 d[8];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int q = 0; q < 8; ++q) d[q] = ffma2(e[q], h[q], b[q]);
#pragma unroll
            for (int q = 0; q < 8; ++q) s += d[q].x;
            }
        } else if (T == 9) {   // 16 scalar FMUL
#pragma unroll
            for (int q = 0; q < 8; ++q) { h[q].x = h[q].x * e[q].x; h[q].y = h[q].y * e[q].y; }
        }
    }
    long long t1 = clock64();
    float acc = s;
    for (int q = 0; q < 8; ++q) acc += h[q].x + h[q].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int T>
void run(const char* name, int ninstr) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int W = 1; W <= 6; ++W) {
        if (W == 5) continue;
        k<T><<<1, 128 * W, 8192>>>(out, cyc, 0.5f);
        k<T><<<1, 128 * W, 8192>>>(out, cyc, 0.5f);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-44s W=%d  cycles/iter %8.1f  cycles per warp-instr per SMSP %6.2f\n", name, W, double(c) / ITER, double(c) / ITER / (ninstr * W));
    }
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("16 FFMA (3 reg)", 16);
    run<9>("16 FMUL", 16);
    run<1>("16 FFMA2 h=e*h+b", 16);
    run<8>("16 FFMA2 h=e*h+const", 16);
    run<2>("16 FMUL2 pair*pair", 16);
    run<3>("16 FMUL2 (s,s)*pair", 16);
    run<4>("16 MUFU.EX2", 16);
    run<5>("step: 16 FMUL2+16 MUFU+16 FFMA2 (48 instr)", 48);
    run<6>("step + 8 LDS.128 broadcast (56 instr)", 56);
    run<7>("step + 8 LDS.128 + 1 LDS (57 instr)", 57);
    run<10>("step without MUFU (32 instr)", 32);
    run<11>("step without MUFU + 8 LDS.128 + 1 LDS (41)", 41);
    run<12>("step + 8 LDS.128 + 1 LDS (57) again", 57);
    run<13>("step + 8 LDS.128 + 1 LDS + y reduce + STS (~62)", 62);
    run<14>("16 FFMA2 d=a*b+c, d distinct", 16);
    run<15>("2 pipelined steps + LDS + y reduce + STS (~124)", 124);
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
    return 0;
}
