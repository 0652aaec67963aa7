/* mtn_b200.h -- C ABI of the B200-native Mamba-TasNet separator forward.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.
 * Every entry point launches hand-written sm_100a kernels on the caller's stream and returns
 * immediately (stream-ordered, re-entrant, CUDA-graph capturable).  The caller owns every
 * buffer; the library allocates nothing and keeps no state besides the last error string.
 * Return value: 0 on success, negative MTN_E* on error (never throws, never exits, never
 * falls back to another implementation).
 *
 * What each entry replaces in the reference (paths relative to /root/reference):
 *   mtn_encoder_cln_fwd  speechbrain dual_path.Encoder.forward (== baseline/avse2/model.py:14-24,
 *                        called Mamba-TasNet/train_wsj0mix.py:89) + ChannelwiseLayerNorm
 *                        (Mamba-TasNet/modules/mamba_masknet.py:118)
 *   mtn_gemm_fwd         every dense contraction of the path: bottleneck / mask 1x1 convs
 *                        (mamba_masknet.py:121,123), in_proj (modules/mamba/bimamba.py:192-196),
 *                        x_proj (modules/mamba/selective_scan_interface.py:186), out_proj
 *                        (bimamba.py:253); epilogues fuse SiLU(z) (ssi.py:155) and
 *                        relu(mask)*mix_w (mamba_masknet.py:136 + train_wsj0mix.py:91-92)
 *   mtn_add_rmsnorm_fwd  Block.forward add + RMSNorm (bimamba.py:446-447) and the final
 *                        add + norm_f (modules/mamba_blocks.py:196-197)
 *   mtn_conv_silu_fwd    causal_conv1d_cuda.causal_conv1d_fwd (ssi.py:182), both directions
 *                        (the reference runs the 2nd on xz.flip(-1), bimamba.py:237)
 *   mtn_scan_fwd         dt_proj GEMM (ssi.py:187) + selective_scan_cuda.fwd (ssi.py:218-220,
 *                        oracle selective_scan_ref ssi.py:91-157), both directions, incl. the 0.5
 *                        averaging of bimamba.py:253; optional initial/final state for the
 *                        sequence-parallel mode
 *   mtn_conv_silu_halo_fwd / mtn_fold_states_fwd / mtn_scan_args.{sum_delta,L_last,h_in,h_out}
 *                        the sequence-parallel long-form mode (BASELINE config 5); new functionality, the
 *                        reference has no counterpart
 *   mtn_cln_fwd          ChannelwiseLayerNorm alone (mamba_masknet.py:118) for the stand-alone MaskNet module
 *   mtn_decoder_fwd      speechbrain dual_path.Decoder (== baseline/avse2/model.py:27-37), both
 *                        speakers, cat + pad/trim (train_wsj0mix.py:95-109)
 *
 * Layout conventions: every activation is channel-last [tokens = batch*frames, channels].
 * "planes" P: a GEMM A-operand tensor is stored as P bf16 planes [P][rows][ld]:
 *   P = 2 ("fp32 mode"): value = hi + lo (split-bf16; the GEMM computes hi*hi + lo*hi + hi*lo in
 *          fp32 TMEM accumulators: ~2^-17 relative operand error), P = 1 ("bf16 mode"): plain bf16.
 */
#ifndef MTN_B200_H
#define MTN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Bumped whenever a struct below or an entry point's signature changes; mtn_abi_version() returns the value the library
 * was built with (experiment builds of tools/devbuild.sh add 1000 so that a product binding refuses them). */
#define MTN_ABI_VERSION 7

#define MTN_OK 0
#define MTN_EINVAL (-1)  /* bad shape / alignment / unsupported option */
#define MTN_ECUDA (-2)   /* CUDA runtime or driver error; see mtn_last_error_string() */

typedef void* mtn_stream_t; /* cudaStream_t */

enum { MTN_EPI_STORE = 0, MTN_EPI_INPROJ = 1, MTN_EPI_MASK = 2, MTN_EPI_RELU = 3, MTN_EPI_XPROJ = 4,
       MTN_EPI_RESADD = 5 };

typedef struct {
    const void* a;       /* bf16 [planes][a_rows][lda]; group g reads columns [g*K, (g+1)*K) */
    const void* w;       /* bf16 [planes][groups*N][K]  (row-major, K contiguous) */
    void* out;           /* fp32 (or bf16 if out_bf16) [M][ldo]; group g writes columns g*out_group_stride + [0,N) */
    const void* aux;     /* MTN_EPI_MASK: mix_w fp32 [M][ld_aux].  MTN_EPI_XPROJ: OUTPUT dtp, bf16 [M][groups][2][RP]
                            (hi plane | lo plane of the first RP = epi_param columns of each group: the dt columns of
                            x_proj in the operand layout of the scan's tensor-core dt_proj).  Else NULL */
    int M, N, K;         /* per-group problem; N % 16 == 0, N <= 256 or N % 256 == 0 or N % 128 == 0; K % 64 == 0 */
    int a_rows;          /* rows allocated per A plane (>= M) */
    int lda, ldo, ld_aux;
    int planes;          /* 1 or 2 */
    int groups;          /* >= 1 */
    int out_group_stride;
    int epilogue;        /* MTN_EPI_* */
    int epi_param;       /* INPROJ: first column that gets SiLU; MASK: enc_dim (aux column = col % enc_dim);
                            XPROJ: RP (16 or 32); RESADD: 1 = out holds a valid residual to add to, 0 = first block
                            (residual := result, bimamba.py:446 `residual = hidden_states`) */
    int out_bf16;        /* 0: fp32 output, 1: bf16 output */
    int max_ctas;        /* 0 = one persistent CTA per SM */
    /* --- RMSNorm folded into the GEMMs (ABI >= 4); all optional ------------------------------------------------
     * Block.forward is Add -> RMSNorm -> Mixer (bimamba.py:446-447).  RMSNorm(res) * g = rstd[row] * res * g[col]: g is
     * folded into the consuming weight (W' = W * g, done once at load), rstd is a per-row scalar that commutes with
     * the contraction.  So the producer GEMM (out_proj / bottleneck, MTN_EPI_RESADD) adds its result to the residual,
     * stores it as fp32 AND as operand planes and accumulates sum(res^2) per row; the consumer GEMM (in_proj / mask)
     * reads those planes and scales its accumulator rows by rsqrt(rowsq * rowsq_scale + rowsq_eps) before its own
     * epilogue.  No separate add/norm kernel, no round trip of the mixer output through HBM.
     * MTN_EPI_RESADD with out2 == NULL and rowsum == NULL is the residual add alone: out (fp32, in/out) += result (epi_param 1)
     * or := result (epi_param 0) -- the default plan's out_proj: the norm kernel then reads the residual stream only. */
    void* out2;          /* MTN_EPI_RESADD: bf16 planes [planes][a2_rows][ldo2] of the updated residual (next GEMM's A); nullable
                            together with rowsum */
    float* rowsum;       /* MTN_EPI_RESADD: fp32 [parts][M], parts = mtn_gemm_rowsum_parts(N): plane k holds the sum of
                            res^2 over the columns one epilogue warp group of one N tile owns (plain stores: nothing to
                            zero, bit-reproducible) */
    int ldo2, a2_rows;
    const float* rowsq;  /* any epilogue: nullable fp32 [rowsq_parts][M] written by a RESADD call; accumulator row r is
                            scaled by rsqrt(sum_k rowsq[k][r] * rowsq_scale + rowsq_eps) */
    float rowsq_scale, rowsq_eps;
    int rowsq_parts;
} mtn_gemm_args;

typedef struct {
    const void* u;        /* bf16 [planes][M][2*di]: conv output, direction d at columns [d*di, (d+1)*di) */
    const float* dbl;     /* fp32 [M][ld_dbl]: direction d at columns d*n_dbl + [dt(R) | B(16) | C(16)] */
    const void* z;        /* silu(z), fp32 (or bf16 if z_bf16) [M][ldz], columns z_col0 + [0, di) */
    const float* w_dt;    /* [2][di][R] */
    const float* dt_bias; /* [2][di] */
    const float* A2;      /* [2][di][16] = -exp(A_log) * log2(e) */
    const float* Dskip;   /* [2][di] */
    void* y;              /* bf16 [planes][M][2*di]: 0.5 * (scan + D*u) * silu(z), same column convention as u */
    const float* h_in;    /* nullable: fp32 [2][batch][di][16] initial state per direction */
    float* h_out;         /* nullable: fp32 [2][batch][di][16] final state per direction */
    int batch, L, di, R, n_dbl, ld_dbl, ldz, z_col0;
    int planes;           /* 1 or 2 */
    int z_bf16;
    int dir_mask;         /* bit0 forward, bit1 backward (3 = both in one launch) */
    /* --- chunked / sequence-parallel scans (ABI >= 2); all optional ----------------------------------------- */
    float* sum_delta;     /* nullable: fp32 [2][batch][di] = sum_t delta_t of each sequence.  With h_out it is the
                             chunk summary of the reduce-then-scan scheme: h_end(h_in) = exp2(A2*sum_delta)*h_in +
                             h_out(h_in = 0).  y may be NULL in that case (summary pass, no output written). */
    int L_last;           /* 0, or the valid length (1..L) of the LAST sequence of the batch: rows beyond it are
                             ignored (a long recording cut into `batch` equal chunks of L frames, ragged tail) */
    /* --- dt_proj on the tensor cores (ABI >= 3); optional ---------------------------------------------------- */
    const void* dtp;      /* nullable: bf16 [M][2][2][RP] written by mtn_gemm_fwd(MTN_EPI_XPROJ), RP = 16 (R <= 16) or
                             32.  When given, delta_pre = dtp . w_dt is a tcgen05.mma per 16-step tile (W_dt rows as
                             the M = 128 operand, accumulator in TMEM) instead of R FMAs per (step, channel). */
} mtn_scan_args;

/* mix [batch][ld_mix >= T] fp32 -> mix_w [batch*L][N] fp32 = relu(conv1d(k=16,s=8)), L = (T-16)/8+1;
 * yn planes = cLN(mix_w).  ld_mix % 4 == 0 (128-bit frame loads); T itself is arbitrary (>= 16). */
int mtn_encoder_cln_fwd(const float* mix, int ld_mix, const float* w_enc /*[N][16]*/, const float* gamma,
                        const float* beta, float* mix_w, void* yn_planes, int batch, int T, int L, int N, int planes,
                        float eps, mtn_stream_t stream);

int mtn_gemm_fwd(const mtn_gemm_args* args, mtn_stream_t stream);
/* number of partial-sum planes a MTN_EPI_RESADD call with this N writes to `rowsum` */
int mtn_gemm_rowsum_parts(int N);

/* res = (h ? h : 0) + (res_valid ? res : 0); xn planes = res * rsqrt(mean(res^2)+eps) * g.
 * h fp32 [M][D] (nullable), res fp32 [M][D] in/out. */
int mtn_add_rmsnorm_fwd(const float* h, float* res, int res_valid, const float* g, void* xn_planes, int M, int D,
                        int planes, float eps, mtn_stream_t stream);

/* Same; additionally (or instead: xn_planes may be NULL) writes the normalised rows as fp32 [M][D].  This is the output
 * of a whole MambaBlocksSequential.forward (modules/mamba_blocks.py:196-197) when the stack is used as a stand-alone
 * sequence model, e.g. as the intra / inter model of DPMamba (hparams/WSJ0Mix/dpmamba_L.yaml:139-161). */
int mtn_add_rmsnorm_out_fwd(const float* h, float* res, int res_valid, const float* g, void* xn_planes, float* out_f32,
                            int M, int D, int planes, float eps, mtn_stream_t stream);

/* Same with an optional bias `beta` [D]: beta != NULL selects nn.LayerNorm (mean removed, biased variance, weight g and
 * bias beta) instead of RMSNorm -- `rms_norm=False` in create_block / norm_f (modules/mamba_blocks.py:36-41,167-169). */
int mtn_add_norm_fwd(const float* h, float* res, int res_valid, const float* g, const float* beta, void* xn_planes,
                     float* out_f32, int M, int D, int planes, float eps, mtn_stream_t stream);

/* xs = xz[:, 0:di] (fp32 or bf16, row stride ldxz) -> u planes [P][M][2*di]:
 * u_fwd[t] = silu(b + sum_k w[k]*xs[t-3+k]),  u_bwd[t] = silu(b' + sum_k w'[k]*xs[t+3-k]), zero padded per
 * utterance.  conv_w [2][di][4], conv_b [2][di]. */
int mtn_conv_silu_fwd(const void* xz, int ldxz, int xz_bf16, const float* conv_w, const float* conv_b, void* u_planes,
                      int batch, int L, int di, int planes, mtn_stream_t stream);

/* Same, for a time chunk of a longer sequence (sequence-parallel mode): halo_lo / halo_hi (nullable, fp32
 * [batch][3][di]) hold the xs rows t = -3..-1 and t = L..L+2 owned by the neighbouring chunks; NULL = zero padding
 * (true utterance edge).  The reference has no such mode (its conv always sees the whole utterance,
 * selective_scan_interface.py:182); the arithmetic per output element is unchanged. */
int mtn_conv_silu_halo_fwd(const void* xz, int ldxz, int xz_bf16, const float* conv_w, const float* conv_b,
                           void* u_planes, int u_rows /* rows allocated per plane, >= batch*L */, const float* halo_lo,
                           const float* halo_hi, int batch, int L, int di, int planes, mtn_stream_t stream);

/* Same with a direction mask (3 = both directions, 1 = forward only; backward-only is not built): a unidirectional stack (`bidirectional=False`,
 * modules/mamba_blocks.py:128 -> mamba_ssm.Mamba; the vendored non-fused branch modules/mamba/bimamba.py:271-285) runs
 * dir_mask = 1, conv_w/conv_b then only need their first [di] rows.  With halo_lo = the last 3 conv inputs of the previous
 * chunk this is also the streaming form of the reference's conv_state cache (bimamba.py:274-277, step :327-333). */
int mtn_conv_silu_dir_fwd(const void* xz, int ldxz, int xz_bf16, const float* conv_w, const float* conv_b,
                          void* u_planes, int u_rows, const float* halo_lo, const float* halo_hi, int batch, int L,
                          int di, int planes, int dir_mask, mtn_stream_t stream);

/* conv_silu (both directions) fused with the x_proj contraction (ABI >= 7): one pass over xs writes u (for the scan) and
 * dbl = [dt | B | C] of both directions; u is never re-read by a GEMM.  Bit-identical to mtn_conv_silu_fwd followed by
 * mtn_gemm_fwd(groups = 2) on the same inputs.  Replaces causal_conv1d_fwd x 2 + F.linear(x_proj) x 2
 * (modules/mamba/selective_scan_interface.py:182-186, modules/mamba/bimamba.py:237).  wx_planes: bf16 [planes][2*n_dbl][di]
 * (rows dt | B | C | zero pad per direction, as for mtn_gemm_fwd); dbl: fp32 [batch*L][ld_dbl], direction d at columns
 * d*n_dbl; di % 64 == 0; n_dbl 48 or 64; no halo rows (whole sequences only). */
int mtn_conv_xproj_fwd(const void* xz, int ldxz, int xz_bf16, const float* conv_w, const float* conv_b, void* u_planes,
                       int u_rows, const void* wx_planes, float* dbl, int ld_dbl, int n_dbl, int batch, int L, int di,
                       int planes, mtn_stream_t stream);

int mtn_scan_fwd(const mtn_scan_args* args, mtn_stream_t stream);

/* Chunked scan, step 2 of 3 (summary pass -> fold -> seeded pass): compose the per-chunk operators
 * h -> exp2(A2 * sum_delta[g]) * h + h_end[g] in time order and write the state entering each chunk of
 * [g0, g0 + n_out).  h_end fp32 [2][G][di][16], sum_delta fp32 [2][G][di] (chunk g of direction d at [d][g]),
 * A2 [2][di][16], h0 nullable [2][di][16] = state entering the first chunk in time order (forward: before g = 0,
 * backward: after g = G-1), h_in out [2][n_out][di][16], h_final nullable [2][di][16] = state after the last chunk.
 * dir_mask as in mtn_scan_args.  No counterpart in the reference (it never chunks the scan across devices). */
int mtn_fold_states_fwd(const float* h_end, const float* sum_delta, const float* A2, const float* h0, float* h_in,
                        float* h_final, int G, int di, int g0, int n_out, int dir_mask, mtn_stream_t stream);

/* The same composition over the summaries exactly as ONE all-gather delivers them (ABI >= 6): `pack` fp32 [W][rec], rank w's
 * record rec = [ h_end [2][cmax][di][16] | sum_delta [2][cmax][di] ] (cmax sub-chunks per rank, unused ones zero = identity
 * operators); global chunk g = w * cmax + c.  h0 / h_in / h_final / g0 / n_out / dir_mask as above. */
int mtn_fold_states_packed_fwd(const float* pack, const float* A2, const float* h0, float* h_in, float* h_final, int W,
                               int cmax, int di, int g0, int n_out, int dir_mask, mtn_stream_t stream);

/* sep fp32 [batch*L][n_spk*N] (speaker-major channels) -> est [batch][T][n_spk] fp32:
 * est[b, 8l+k, s] = sum over frames/taps of sum_n w_dec[n][k] * sep[b,l,s*N+n]; zero-padded / trimmed to T.
 * frames: scratch fp32 [batch*L][n_spk][16]. */
int mtn_decoder_fwd(const float* sep, const float* w_dec /*[N][16]*/, float* frames, float* est, int batch, int T, int L,
                    int N, int n_spk, mtn_stream_t stream);

/* Streaming decoder: the chunk's L frames finalise est [batch][T = 8*L][n_spk]; `tail` fp32 [batch][n_spk][8] holds the
 * second half of the previous chunk's last frame on entry (zeros before the first chunk) and of this chunk's last frame on
 * return.  tail == NULL is mtn_decoder_fwd.  Concatenating the chunks' outputs (plus the final tail) reproduces the
 * one-shot ConvTranspose1d overlap-add bit for bit. */
int mtn_decoder_stream_fwd(const float* sep, const float* w_dec, float* frames, float* est, float* tail, int batch, int T,
                           int L, int N, int n_spk, mtn_stream_t stream);

/* ChannelwiseLayerNorm alone (MaskNet called on an externally produced mix_w): x fp32 [M][N] -> yn planes */
int mtn_cln_fwd(const float* x, const float* gamma, const float* beta, void* yn_planes, int M, int N, int planes,
                float eps, mtn_stream_t stream);

/* mask_nonlinear = "softmax" of MaskNet (modules/mamba_masknet.py:133-134: F.softmax(score, dim=2) on [n_spk, B, N, L],
 * i.e. over the N encoder channels): score fp32 [rows][n_spk*N] in place; mix_w nullable fp32 [rows][N]: the result is
 * also multiplied by it (mask application, train_wsj0mix.py:91-92). */
int mtn_softmax_mask_fwd(float* score, const float* mix_w, int rows, int N, int n_spk, mtn_stream_t stream);

/* fp32 [rows][cols] (row stride ld) -> bf16 planes [planes][rows][cols] (weight packing helper) */
int mtn_split_planes(const float* src, int ld, void* dst_planes, int rows, int cols, int planes, mtn_stream_t stream);

/* Evaluation front end: SI-SNR with 2-speaker PIT and SI-SNR improvement over the mixture, as
 * `save_results` computes them per utterance (Mamba-TasNet/train_wsj0mix.py:548-558; SI-SNR = cal_si_snr,
 * baseline/avse2/utils/dnn.py:15-57).  est, src fp32 [batch][T][2], mix fp32 [batch][ld_mix >= T].
 * out fp32 [batch][8] = { si_snr (best permutation, mean over speakers), si_snr_i, permutation (0 direct, 1 swapped),
 * si_snr of the mixture, pair matrix est0/src0, est0/src1, est1/src0, est1/src1 }, all in dB (positive = better).
 * workspace: mtn_si_snr_workspace_bytes(batch, T) bytes, 8-byte aligned. */
size_t mtn_si_snr_workspace_bytes(int batch, int T);
int mtn_si_snr_pit_fwd(const float* est, const float* src, const float* mix, int ld_mix, int batch, int T,
                       void* workspace, size_t workspace_bytes, float* out, mtn_stream_t stream);
/* The same for n_spk = 1..4 speakers (`num_spks: 3` is the wsj0-3mix setting of every recipe,
 * hparams/WSJ0Mix/mambatasnet_S.yaml:39; save_results appends s3_sig then, train_wsj0mix.py:537-538) (ABI >= 6).
 * est, src fp32 [batch][T][n_spk].  out fp32 [batch][out_stride >= 4 + n^2]: { si_snr (best of the n! assignments, mean
 * over speakers), si_snr_i, lexicographic rank of the best assignment, si_snr of the mixture, pair matrix [est i][src j] }
 * and, when out_stride >= 4 + n^2 + n, the assignment itself (entry i = source matched to estimate i). */
size_t mtn_si_snr_workspace_bytes_n(int batch, int T, int n_spk);
int mtn_si_snr_pit_n_fwd(const float* est, const float* src, const float* mix, int ld_mix, int batch, int T, int n_spk,
                         void* workspace, size_t workspace_bytes, float* out, int out_stride, mtn_stream_t stream);

/* ---- DPMamba (dual-path) glue, SURVEY 8f rank 1 ------------------------------------------------------------------
 * The mask network of the dpmamba_* recipes is speechbrain 1.0.0 `Dual_Path_Model` [third party, not vendored]
 * (Mamba-TasNet/hparams/WSJ0Mix/dpmamba_L.yaml:164-174) with MambaBlocksSequential as intra and inter model; its forward
 * is restated in the vendored subclass Mamba-TasNet/modules/dual_path.py:56-150.  The reference's 4-D tensor [B, N, K, S]
 * is held as fp32 rows (b, s, k) x C ("layout A", consumed by the intra model) or (b, k, s) x C ("layout T", inter). */

/* GroupNorm(1, C, eps) = `select_norm("ln", ...)` of speechbrain: one mean / biased variance per utterance over all
 * rows x C elements.  Pass 1: deterministic partial sums, `partials` = mtn_gn_partials_bytes(batch, rows, C) bytes,
 * 16-byte aligned.  Replaces `self.norm` (dual_path.py:83) and intra_norm / inter_norm of every Dual_Computation_Block. */
size_t mtn_gn_partials_bytes(int batch, int rows, int C);
int mtn_gn_stats_fwd(const float* x /* [batch][rows][C] */, void* partials, int batch, int rows, int C, mtn_stream_t stream);

typedef struct {
    const float* x;        /* fp32 [batch][S*K][C]; rows ordered (s, k), or (k, s) when x_transposed */
    const void* partials;  /* from mtn_gn_stats_fwd on x */
    const float* w;        /* [C] GroupNorm weight */
    const float* bias;     /* [C] GroupNorm bias */
    const float* skip;     /* nullable fp32 [batch][S][K][C]: added after the affine (the block's residual connections) */
    float* out_a;          /* nullable fp32 [batch][S][K][C] */
    float* out_a2;         /* nullable second copy in the same layout (the next stack's input buffer, which it consumes) */
    float* out_t;          /* nullable fp32 [batch][K][S][C]: the same values with rows transposed (inter-model input) */
    void* planes;          /* nullable bf16 operand planes [n_planes][plane_rows][C], rows as out_a */
    int batch, S, K, C;
    int x_transposed;
    int n_planes, plane_rows;
    float eps;
    const float* blend;    /* nullable fp32 [batch][S][K][C]: result = 0.5 * result + 0.5 * blend afterwards -- the periodic
                              skip to the segmented input of Dual_Path_Model_Skip (modules/dual_path.py:114-116), folded
                              into the norm that ends the preceding block */
} mtn_gn_apply_args;
/* Pass 2: y = (x - mean) * rstd * w + bias (+ skip), written to every non-NULL output. */
int mtn_gn_apply_fwd(const mtn_gn_apply_args* args, mtn_stream_t stream);

/* Pass 2 fused with the Add -> RMSNorm that opens the NEXT Mamba stack (Block.forward with residual None,
 * modules/mamba/bimamba.py:446-447): besides args->out_a (optional) the row goes to that stack's residual stream res_next
 * (fp32 [batch][S*K][C]) and, times rsqrt(mean_C(row^2) + rms_eps) * g_next, to its operand planes xn_next
 * ([n_planes][plane_rows][C]); both at row (b,k,s) when next_transposed (inter model), else (b,s,k).  C in {128, 256, 512};
 * args->out_a2 / out_t / planes must be NULL. */
int mtn_gn_apply_norm_fwd(const mtn_gn_apply_args* args, float* res_next, void* xn_next_planes, const float* g_next,
                          int next_transposed, float rms_eps, mtn_stream_t stream);

/* Number of chunks S that Dual_Path_Model._Segmentation makes of L frames with chunk size K (50 % overlap, zero padded). */
int mtn_dp_num_chunks(int L, int K);
/* _padding + _Segmentation: x fp32 [batch][L][C] -> out_a (and out_a2, nullable) fp32 [batch][S][K][C]. */
int mtn_dp_segment_fwd(const float* x, float* out_a, float* out_a2, int batch, int L, int C, int K, int S,
                       mtn_stream_t stream);
/* PReLU (one slope, dual_path.py:112) + _over_add (dual_path.py:126): X fp32 [batch][S][K][C] -> operand planes
 * [n_planes][plane_rows >= batch*L][C] of the frame-domain sum, ready for the conv2d 1x1 (dual_path.py:117; applied after
 * the overlap-add here, which is the same linear map on half the rows -- its bias then counts twice). */
int mtn_dp_overadd_prelu_fwd(const float* X, const float* prelu_w, void* planes, int plane_rows, int batch, int L, int C,
                             int K, int S, int n_planes, mtn_stream_t stream);
/* planes = x + bias_scale * bias; x fp32 [rows][ldx], first C columns. */
int mtn_bias_planes_fwd(const float* x, int ldx, const float* bias, float bias_scale, void* planes, int plane_rows, int rows,
                        int C, int n_planes, mtn_stream_t stream);
/* Gated output layer `output(x) * output_gate(x)` (dual_path.py:127): og fp32 [rows][groups][2*D] = pre-activations
 * (tanh branch | sigmoid branch) per speaker group -> planes [n_planes][plane_rows][groups*D] =
 * tanh(o + bo) * sigmoid(g + bg). */
int mtn_gate_planes_fwd(const float* og, const float* bo, const float* bg, void* planes, int plane_rows, int rows, int groups,
                        int D, int n_planes, mtn_stream_t stream);

/* ---- one-launch streaming push (ABI >= 7) ------------------------------------------------------------------------------
 * The whole causal separator for one chunk of F <= 32 frames per stream, as ONE kernel launch: a thread-block cluster of
 * 2 * d_model / dsl CTAs per stream (dsl = 32, 64 or 128 d_inner channels each), contractions on warp-level mma.sync with the same split-bf16
 * operand model as mtn_gemm_fwd.  Replaces, for short chunks, the reference's per-token `Mamba.step` + `inference_params`
 * caches (Mamba-TasNet/modules/mamba/bimamba.py:320-372, :374-404) under `MambaBlocksSequential.forward(x,
 * inference_params)` (modules/mamba_blocks.py:186-197) together with the Encoder / MaskNet / Decoder calls around the
 * stack (train_wsj0mix.py:86-111); same carried state as the chunked batch plan (conv history, SSM state, overlap-add
 * tail).  Needs enc_dim == d_model in {64, 128, 256, 512}, expand 2, d_state 16, d_conv 4, RMSNorm blocks, ReLU mask,
 * 2 speakers, unidirectional mixers, fp32 mode.  Weight blobs are packed once by the caller (layouts below; the Python
 * packer is avse_challenge_b200/stream_fused.py):
 *   head       fp32: w_enc^T [16][N] | cLN gamma [N] | beta [N] | norm_f [D] | w_dec [N][16]
 *   layer_vec  fp32 per layer: norm [D] | conv_w [di][4] | conv_b [di] | w_dt^T [R][di] | dt_bias [di] | A2 [di][16] | D [di]
 *   *_frag     mma.sync A-operand fragments, bf16: [cluster rank][16-column tile][k-step of 16][plane hi|lo][lane][8]:
 *              with c = dsl: bot_frag (rank r: bottleneck rows (c/2)r..), mask_frag (rank r: mask rows c*r..), per layer
 *              in_proj (rank r: c rows from c*r of x, then c rows from di + c*r of z; K = D) | x_proj (rows dt|B|C padded to a
 *              multiple of 16; K = channels c*r..c*r+c-1) | out_proj (all D rows; K = the same channels) */
typedef struct {
    const float* mix;       /* [B][ld_mix]: this push's chunk, 8*F samples (first push: 8*F + 8) */
    float* in_tail;         /* in/out [B][8]: the last 8 samples of the previous chunk (the encoder window overlaps it);
                               ignored on input when first != 0, always updated */
    float* est;             /* [B][8*F][n_spk] */
    float* halo;            /* in/out [n_layers][B][3][di]: last three conv inputs (the reference's conv_state) */
    float* h;               /* in/out: layer l, stream b at h + l*h_layer_stride + b*di*16: [di][16] (ssm_state) */
    float* ola_tail;        /* in/out [B][n_spk][8]: second half of the last decoder frame */
    const float* head;
    const void* bot_frag;
    const void* mask_frag;
    const float* layer_vec;
    const void* layer_frag;
    size_t h_layer_stride;  /* floats */
    size_t layer_vec_stride;  /* floats */
    size_t layer_frag_stride; /* bytes */
    int B, F, N, D, di, R, n_spk, n_layers, ld_mix;
    int first;              /* != 0: first push of a stream (no carried samples; the chunk holds F + 1 hops) */
    float eps_cln, eps_rms;
    unsigned long long* timeline; /* nullable debug aid: [n_layers + 2][16] globaltimer stamps of stream 0, rank 0 */
    /* conv history addressing (floats): layer l, stream b at halo + l*halo_layer_stride + b*halo_stream_stride: [3][di].
     * Dense = {halo_rows*di, B*halo_rows*di}. */
    size_t halo_stream_stride, halo_layer_stride;
    int halo_rows;          /* 3, or 4: each [halo_rows][di] block is the reference's conv_state (time-major): rows 1..3 are read
                               as the history, all four are rewritten with the last four conv inputs */
    /* stack-only mode (both nullable together): `MambaBlocksSequential.forward(x, inference_params)` itself --
     * stack_x [B][F][D] fp32 is the stack's input, stack_out [B][F][D] = norm_f(blocks(x)); the encoder / mask / decoder
     * arguments (mix, in_tail, est, ola_tail, bot_frag, mask_frag) are then unused and may be NULL. */
    const float* stack_x;
    float* stack_out;
    int dsl;                /* d_inner channels per CTA: 32, 64 or 128 (0 = 64).  Cluster size = 2 * D / dsl (2..16); the weight
                               fragments must have been packed for the same value.  32: lowest latency; 128: most streams
                               resident at once */
} mtn_stream_push_args;

int mtn_stream_push_fwd(const mtn_stream_push_args* args, mtn_stream_t stream);
size_t mtn_sizeof_stream_push_args(void);
/* dynamic shared memory per CTA for F frames, d_model D, dsl channels per CTA; 0 = no such kernel; a launch needs <= 227 KiB */
size_t mtn_stream_push_smem_bytes(int F, int D, int dsl);

const char* mtn_last_error_string(void);
int mtn_abi_version(void);
/* sizeof() of the argument structs as the library was compiled: a binding generated from another revision of this header
 * (or a hand-copied ctypes / cgo / JNI struct) can check its layout before the first launch (ABI >= 6). */
size_t mtn_sizeof_gemm_args(void);
size_t mtn_sizeof_scan_args(void);
size_t mtn_sizeof_gn_apply_args(void);

#ifdef __cplusplus
}
#endif
#endif /* MTN_B200_H */
