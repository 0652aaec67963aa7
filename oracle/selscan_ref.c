/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's selective_scan_ref
 * (Mamba-TasNet/modules/mamba/selective_scan_interface.py:91-157), channel-last layout.
 *
 *   delta = softplus(delta_pre + bias)            (:110-112; torch softplus: linear above 20)
 *   h     = exp(delta * A) * h + delta * B * u    (:126,:131,:139)
 *   y     = sum_n h[n] * C[n] + D * u             (:144,:153)
 *   out   = y * silu(z)                           (:155)   (skipped when z == NULL)
 *
 * fp32 arithmetic throughout like the reference on fp32 inputs; `reverse` walks t = L-1..0
 * (the reference flips its inputs instead, bimamba.py:237,253).  h_in (nullable) / h_last are
 * the chunk-carry extension (reference: zeros at :124, optional last state at :147-148).
 * sum_delta (nullable) returns the per-channel sum of delta over the chunk.
 *
 * Never linked into the product; used by tests/, smoke() and bench.py's cpu_baseline leg only.
 */
#include <math.h>
#include <stddef.h>

static inline float softplus_f(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
static inline float silu_f(float x) { return x / (1.0f + expf(-x)); }

void selscan_ref_f32(const float* u, const float* delta_pre, const float* A, const float* Bm,
                     const float* Cm, const float* D, const float* z, const float* bias,
                     const float* h_in, float* out, int batch, int L, int di, int Ns, int reverse,
                     float* h_last, float* sum_delta)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < batch; ++b) {
        for (int d = 0; d < di; ++d) {
            float h[64];
            for (int n = 0; n < Ns; ++n) h[n] = h_in ? h_in[((size_t)b * di + d) * Ns + n] : 0.0f;
            float sd = 0.0f;
            for (int i = 0; i < L; ++i) {
                int t = reverse ? (L - 1 - i) : i;
                size_t idx = ((size_t)b * L + t) * di + d;
                const float* Bt = Bm + ((size_t)b * L + t) * Ns;
                const float* Ct = Cm + ((size_t)b * L + t) * Ns;
                float dl = softplus_f(delta_pre[idx] + bias[d]);
                float uu = u[idx];
                float y = 0.0f;
                sd += dl;
                for (int n = 0; n < Ns; ++n) {
                    float dA = expf(dl * A[(size_t)d * Ns + n]);
                    float dBu = dl * Bt[n] * uu;
                    h[n] = dA * h[n] + dBu;
                    y += h[n] * Ct[n];
                }
                y += uu * D[d];
                if (z) y *= silu_f(z[idx]);
                out[idx] = y;
            }
            if (h_last) for (int n = 0; n < Ns; ++n) h_last[((size_t)b * di + d) * Ns + n] = h[n];
            if (sum_delta) sum_delta[(size_t)b * di + d] = sd;
        }
    }
}
