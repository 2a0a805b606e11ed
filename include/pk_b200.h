/* pk_b200.h — C ABI of the B200-native acoustic-model hot path (libpk_b200.so).
 *
 * Boundary: the reference (mravanelli/pytorch-kaldi) has no native code; its hot path is the
 * stock-PyTorch forward/backward of the `neural_networks.py` module zoo called from
 * `utils.forward_model` (utils.py:2330/2339) inside `core.run_nn`'s minibatch loop
 * (core.py:577-699).  These entry points are what the `torch.autograd.Function` shims of the
 * drop-in `neural_networks.py` bind (through ctypes); each comment names the reference code
 * the call replaces.  See INTEGRATION.md for the reference-side binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; no torch types cross the ABI;
 *   - `stream` is a cudaStream_t passed as void* (the caller's current stream);
 *   - return value 0 = ok; non-zero = error, message via pk_last_error() (thread-local);
 *     nothing aborts or exits;
 *   - "row-major"      [rows][ld]  : rows = frames n = t*B + b (utils.py:2323 t-major order);
 *     "channel-major"  [chan][ld]  : one row per feature/unit, ld >= T*B, column = t*B + b;
 *   - fp16 operand buffers need ld % 8 == 0 and 16-byte aligned bases (TMA requirement),
 *     fp32 (tf32) operand buffers need ld % 4 == 0.
 */
#ifndef PK_B200_H_
#define PK_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PK_ABI_VERSION 2

/* operand types for pk_gemm_tn */
#define PK_F16 0
#define PK_TF32 2

/* activation ids == neural_networks.act_fun (neural_networks.py:36-57) */
#define PK_ACT_RELU 0
#define PK_ACT_TANH 1
#define PK_ACT_SIGMOID 2
#define PK_ACT_LEAKY_RELU 3
#define PK_ACT_ELU 4
#define PK_ACT_LINEAR 5

/* recurrent cell kinds (neural_networks.py classes liGRU :997, RNN :1319, GRU :486,
 * minimalGRU :1158, LSTM :300) */
#define PK_CELL_LIGRU 0
#define PK_CELL_RNN 1
#define PK_CELL_GRU 2
#define PK_CELL_MGRU 3
#define PK_CELL_LSTM 4
#define PK_CELL_MASK 0xff
/* optional flags OR-ed into `cell` */
/* kernel choice for the persistent liGRU / RNN recurrence.  Default (no flag): the faster of the two on this
 * part as measured (profiles/r2_selftest_tc_vs_ws.log) — the warp-specialised register-stationary mma.sync
 * kernels for H <= 560, the tcgen05 kernels (weights stationary in tensor memory) for 560 < H <= 1024. */
#define PK_REC_WS 0x8000   /* force the warp-specialised mma.sync kernels (H <= 560) */
#define PK_REC_TC 0x400000 /* force the tcgen05 kernels (H <= 1024) */
/* backward mma.sync kernel, formulation of the per-step exchange: all-gather of (da, dpz) blocks, or K-split with a
 * reduce-scatter of fp16 partial sums (half the DSMEM bytes per step).  Default: the measured-faster one. */
#define PK_REC_BWD_ALLGATHER 0x800000
#define PK_REC_BWD_KSPLIT 0x1000000
/* timing experiments only (results are incomplete): skip the global stores / loads */
#define PK_REC_DBG_NOSTORE 0x10000
#define PK_REC_DBG_NOLOAD 0x20000
#define PK_REC_GROUPS(n) ((n) << 19) /* tcgen05 kernels: n in 1..3 arrival-group barriers per step (default 1) */
#define PK_REC_DBG_NOPROXYFENCE 0x40000 /* tcgen05 kernels: no fence.proxy.async between chunk arrival and its MMAs */
#define PK_REC_DBG_BLOCKINGWAIT 0x200000 /* tcgen05 kernels: MMA issuer suspends in try_wait instead of spinning on test_wait */

const char* pk_last_error(void);
int pk_version(void);

/* C[M][ldc] (fp32) = alpha * (*alpha_dev) * A[M][a_k0 : a_k0+K] . B[N][b_k0 : b_k0+K]^T (+ bias);
 * both operands K-major ([rows][ld]).  tcgen05 / TMEM / TMA kernel.
 *   bias_mode 1: bias[n], 2: bias[m];  rowstats: optional [M][2] doubles accumulating per-row
 *   (sum, sum of squares) of the OUTPUT (BatchNorm batch statistics of a channel-major
 *   projection);  accumulate: C += ...;  split_k > 1 partitions K over gridDim.z with fp32
 *   atomics;  a_k0 / b_k0 (multiples of 16 bytes) select a sub-range of each operand's K axis;
 *   a_kext / b_kext: valid extent of that axis (0 -> k0 + K), reads past it return zeros;
 *   amax_bits: optional device uint32 receiving atomicMax of |C| (float bit pattern) so that the
 *   consumer's loss scale needs no extra pass (see pk_amax_finalize).
 * Replaces: nn.Linear forward/backward GEMMs (neural_networks.py:1114-1115, :432-435,
 * :609-611, :138-148 and their autograd transposes). */
int pk_gemm_tn(int dtype, int M, int N, int K, const void* A, int64_t lda, int64_t a_k0,
               int64_t a_kext, const void* B, int64_t ldb, int64_t b_k0, int64_t b_kext, float* C,
               int64_t ldc, const float* bias, int bias_mode, double* rowstats, float alpha,
               const float* alpha_dev, int accumulate, int split_k, void* amax_bits, void* stream);

/* outT[c][r] = in[r][c] (fp32, optional) plus optional fp16 copies scaled by *scale_dev:
 * outT16 (channel-major) and in16 (row-major).  Replaces flip/cat/view shuffles
 * (neural_networks.py:1095-1097, :1144-1150) and feeds the K-major GEMM operands. */
int pk_transpose_f32(const float* in, int64_t ldi, int R, int C, float* outT, int64_t ldo,
                     void* outT16, int64_t ldo16, void* in16, int64_t ldi16,
                     const float* scale_dev, void* amax_bits, void* stream);
int pk_convert_f16(const float* in, int64_t ldi, int R, int C, void* out, int64_t ldo,
                   const float* scale_dev, void* stream);

/* scale_out[0] = 2^k with amax(x) * 2^k in [2^(target_log2-1), 2^target_log2): the loss scale
 * that keeps fp16 gradient operands in range; scale_out[1] = 2^-k.  amax_scratch: 1 float. */
int pk_amax_scale(const float* x, int64_t ld, int R, int C, float target_log2,
                  float* amax_scratch, float* scale_out, void* stream);
/* same, from an amax accumulated by a producer kernel (pk_gemm_tn / pk_transpose_f32 `amax_bits`);
 * re-zeroes the accumulator for its next use. */
int pk_amax_finalize(void* amax_bits, float target_log2, float* scale_out, void* stream);

/* nn.BatchNorm1d(C, momentum) over the projection rows (neural_networks.py:1070-1071,
 * :1118-1124): stats = [C][2] doubles (sum, sumsq over the n_unique = T*B de-duplicated rows);
 * n_ref = number of rows the reference normalised (T*2B when bidirectional) for the unbiased
 * running_var.  Writes folded scale = gamma*rstd, shift = beta - mean*scale, and
 * mean/rstd for the backward.  training=0 uses running stats. */
int pk_bn_finalize(const double* stats, int C, int64_t n_unique, int64_t n_ref,
                   const float* gamma, const float* beta, float eps, float momentum,
                   int training, float* running_mean, float* running_var, int64_t* num_batches,
                   float* scale, float* shift, float* mean_out, float* rstd_out, void* stream);
/* no normalisation: scale = 1, shift = bias (or 0) */
int pk_fill_scale_shift(const float* bias, int C, float* scale, float* shift, void* stream);

/* Backward of the (optional) BatchNorm on the de-duplicated projection.  GT = [ndir][C][ldt]
 * gradients w.r.t. the normalised pre-activations in natural time (both directions are
 * summed; pass GT = NULL and GT16 = the fp16 copy scaled by *gscale to read that instead), PT = [C][ldp] raw projections.  Outputs dgamma/dbeta [C] (dbeta = bias grad when
 * use_bn=0) and dP as fp16 * (*gscale) in channel-major (dPT16) and row-major (dP16) form.
 * sums_scratch: 2*C doubles. */
int pk_bn_bwd(int C, int ndir, int64_t n, const float* GT, const void* GT16, int64_t ldt, const float* PT,
              int64_t ldp, int use_bn, int training, const float* mean, const float* rstd,
              const float* gamma, const float* gscale, float* dgamma, float* dbeta, void* dPT16,
              int64_t ld16t, void* dP16, int64_t ld16r, double* sums_scratch, void* stream);

/* One recurrent layer, all T steps in ONE persistent cluster kernel.
 * cell: PK_CELL_LIGRU, or PK_CELL_RNN (h = act(W x + U h) * mask, neural_networks.py:1438-1447) which runs
 * on the same kernel with the update gate pinned to 0: pass G = 2 gate blocks with the second one zero.
 *   PT    [G*H][ldp]  channel-major projections W x (G gates: liGRU h,z), shared by both
 *                     directions (direction 1 reads time T-1-k);
 *   scale/shift [G*H] folded BatchNorm (or 1 / bias);   U [G*H][H] recurrent weights (fp32);
 *   mask  [ndir*B][H] dropout mask of the candidate (neural_networks.py:1102-1111) or NULL
 *                     with mask_scalar = 1-p (eval);
 * outputs (any may be NULL): Y32 [T*B][ldy32] row-major [t,b,d*H+u] (= module output),
 * Y16 fp16 copy (next layer's GEMM operand), HT/HT16/ZT/HCT channel-major [ndir*H][ldt]
 * (state, update gate, masked candidate — saved for the backward), HP16 = fp16 copy of the
 * PREVIOUS state h_{k-1} stored at the column of step k (zeros at k=0): the K-major operand
 * of dU = sum_t G_t^T h_{t-1}, so that product needs no time shift.
 * Replaces: the `for k in range(x.shape[0])` loops (liGRU neural_networks.py:1130-1141). */
int pk_rnn_layer_fwd(int cell, int T, int B, int H, int ndir, int act, const float* PT,
                     int64_t ldp, const float* scale, const float* shift, const float* U,
                     const float* mask, float mask_scalar, float* Y32, int64_t ldy32, void* Y16,
                     int64_t ldy16, float* HT, void* HT16, void* HP16, float* ZT, float* HCT,
                     int64_t ldt, void* stream);

/* Reverse-time persistent kernel: dYT [ndir*H][ldt] channel-major gradient w.r.t. the layer
 * output -> GT16 [ndir][G*H][ldt] = fp16 gradients w.r.t. the normalised pre-activations scaled by
 * *gscale (operand of dU = sum_t G_t^T h_{t-1} and input of pk_bn_bwd).  GT (fp32, unscaled) is
 * optional: only the legacy kernels write it.
 * Replaces: autograd through the per-step graph (core.py:634 loss.backward()). */
int pk_rnn_layer_bwd(int cell, int T, int B, int H, int ndir, int act, const float* dYT,
                     const float* HT, const float* ZT, const float* HCT, int64_t ldt,
                     const float* U, const float* mask, float mask_scalar, const float* gscale,
                     float* GT, void* GT16, void* stream);

/* Step-wise recurrent path: one fused tensor-core kernel per time step, all T launches issued inside this
 * call.  Serves the cells / sizes the persistent kernels do not hold: PK_CELL_LSTM (neural_networks.py:
 * 300-483), PK_CELL_GRU (:486-654) and PK_CELL_MGRU (:1158-1316) — these two need two dependent products per
 * step and run as two launches per step — and PK_CELL_LIGRU with any H (e.g. 5x1024).  Same tensor conventions as pk_rnn_layer_fwd; gate
 * order of PT/scale/shift/U rows and of the saved tensors sv0..sv4 ([ndir*H][ldt] fp32, may be NULL when the
 * backward is not needed):
 *   liGRU: gates (h, z)       saved sv0 = z, sv1 = masked candidate
 *   LSTM : gates (f, i, o, c) saved sv0 = f, sv1 = act(c~)*mask, sv2 = i, sv3 = o, sv4 = cell state c
 *   GRU  : gates (h, z, r)    saved sv0 = z, sv1 = masked candidate, sv2 = r;   HX16 = fp16 (r * h_{k-1})
 *   mGRU : gates (h, z)       saved sv0 = z, sv1 = masked candidate;            HX16 = fp16 (z * h_{k-1})
 * HX16 [ndir*H][ldt] (GRU / minimalGRU only, else NULL) is the K-major operand of dUh like HP16 is for the others.
 * workspace: device scratch of pk_rnn_step_workspace_bytes(cell, T, B, H, ndir, backward) bytes (packed fp16
 * weights, fp32 state, double-buffered fp16 operands). */
int64_t pk_rnn_step_workspace_bytes(int cell, int T, int B, int H, int ndir, int backward);
/* number of __global__ launches one pk_rnn_step_fwd / _bwd call performs for this shape (weight packing + the
 * recurrent kernel(s)): 2 for the cluster-persistent kernels (csrc/pk_cell_cluster.cu LSTM, pk_cell_cluster2.cu GRU /
 * minimalGRU, H <= 560: 16 batch rows per thread-block cluster, the CTA's slice of the recurrent weights stationary in
 * shared memory, state / partial sums exchanged over distributed shared memory with st.async + mbarrier),
 * 2-3 for the cooperative step-wise kernels, T (2T for GRU / minimalGRU) + packs for per-step launches.
 * Environment: PK_LSTM_CLUSTER=0 / PK_GRU_CLUSTER=0 select the step-wise family for A/B runs. */
int pk_rnn_step_launches(int cell, int T, int B, int H, int ndir, int backward);
/* 1 if pk_rnn_step_fwd / _bwd run this (cell, H) on the cluster-persistent kernels, 0 for the step-wise family */
int pk_rnn_step_is_cluster(int cell, int H);
int pk_rnn_step_fwd(int cell, int T, int B, int H, int ndir, int act, const float* PT, int64_t ldp,
                    const float* scale, const float* shift, const float* U, const float* mask,
                    float mask_scalar, float* Y32, int64_t ldy32, void* Y16, int64_t ldy16, float* HT,
                    void* HT16, void* HP16, void* HX16, float* sv0, float* sv1, float* sv2, float* sv3,
                    float* sv4, int64_t ldt, void* workspace, int64_t workspace_bytes, void* stream);
/* reverse-time counterpart: GT16 [ndir][NG*H][ldt] fp16 scaled by *gscale (gate order as above). */
int pk_rnn_step_bwd(int cell, int T, int B, int H, int ndir, int act, const float* dYT, const float* HT,
                    const float* sv0, const float* sv1, const float* sv2, const float* sv3,
                    const float* sv4, int64_t ldt, const float* U, const float* mask, float mask_scalar,
                    const float* gscale, void* GT16, void* workspace, int64_t workspace_bytes,
                    void* stream);

/* ---- FusionLinearConv (neural_networks.py:2057-2099, used by liGRU_layer :795-995 / fusionRNN_jit :719-793) ----
 * The shared per-microphone affine map Conv1d(1, C, kernel = d, stride = d) is pk_gemm_tn over the zero-copy view
 * [N*M][d] of the input (row n*M + m = channel m of frame n) -> O [N*M][ldo] fp32 (+ bias).  These two calls do the
 * rest.  Columns are G gate blocks of Hh (e.g. wh | wz of a liGRU_layer), each with its own slope:
 *   fwd: P[n][c] = red * sum_m act(O[n*M+m][c]);  mode 0: act(x) = x > 0 ? x : slopes[c / Hh] * x (relu 0, leaky_relu
 *        0.01, prelu = the learnt parameter, read on the device), mode 1: tanh;  red = 1 ("sum") or 1/M ("mean")
 *   bwd: dO[n*M+m][c] = red * dP[n][c] * act'(O[n*M+m][c]);  dbias[c] += column sums of dO (conv bias gradient);
 *        dslope[c / Hh] += sum over O <= 0 of red * dP * O (PReLU gradient; may be NULL).  dbias / dslope must be
 *        zeroed by the caller (atomics). */
int pk_fusion_reduce_fwd(const float* O, int64_t ldo, int64_t N, int M, int C, int Hh, int mode, const float* slopes,
                         float red, float* P, int64_t ldp, void* stream);
int pk_fusion_reduce_bwd(const float* dP, int64_t lddp, const float* O, int64_t ldo, int64_t N, int M, int C, int Hh,
                         int mode, const float* slopes, float red, float* dO, int64_t lddo, float* dbias, float* dslope,
                         void* stream);

/* ---- conv front-ends: CNN (neural_networks.py:1464-1556), SincNet (:1559-1665), SincConv (:1668-1813) ----
 * Activations are position-major fp16 A16[n][l][c] (channel pitch Cp = pad8(C)); a stride-1 valid convolution
 * is pk_gemm_tn with A = the activation buffer viewed with row pitch Cp and K = k*Cp (overlapping rows), B =
 * W16 from pk_conv_pack_weights; rows l > L-k of every frame are garbage and ignored by pk_conv_post_fwd. */

/* ln0 = LayerNorm(input_dim) over the sample axis (:1541-1542, :1644-1645; unbiased std, eps added to std):
 * y [N][L] = gamma (x - mean)/(std + eps) + beta; stats [N][2] = mean, 1/(std+eps). */
int pk_rowln_fwd(const float* x, int64_t ldx, int N, int L, const float* gamma, const float* beta, float eps,
                 float* y, float* stats, void* stream);
/* gradients of ln0's gamma/beta [L] from G [N*L][ldg] = dO16 . Wflip16(viewed [k][Cop])^T (columns in the
 * flipped tap order of that operand), raw input x and the stats of pk_rowln_fwd. */
int pk_conv_ln0_bwd(const float* G, int64_t ldg, int N, int L, int Lout, int k, const float* x, int64_t ldx,
                    const float* stats, float* dgamma, float* dbeta, void* stream);
/* SincConv filter synthesis (:1777-1803): band-pass = 2 high sinc - 2 low sinc, max-normalised, Hamming
 * windowed; low_hz_/band_hz_ [C] (the module's parameters), k odd.  filt [C][k] fp32. */
int pk_sinc_filters_fwd(const float* low_hz_, const float* band_hz_, int C, int k, float sample_rate,
                        float min_low_hz, float min_band_hz, float* filt, void* stream);
int pk_sinc_filters_bwd(const float* low_hz_, const float* band_hz_, int C, int k, float sample_rate,
                        float min_low_hz, float min_band_hz, const float* dfilt, float* dlow, float* dband,
                        void* stream);
/* w [Co][Ci][k] fp32 (nn.Conv1d layout) -> W16 [Co][ldw], column kk*Cip + ci (forward operand) and/or
 * Wflip16 [Ci][ldf], column j*Cop + co holding w[co][ci][k-1-j] (operand of the input gradient).  Either may
 * be NULL; the caller zero-fills the padding. */
int pk_conv_pack_weights(const float* w, int Co, int Ci, int k, void* W16, int Cip, int64_t ldw, void* Wflip16,
                         int Cop, int64_t ldf, void* stream);
/* first layer (one input channel): explicit fp16 im2col Xcol [N*L][Kp] (rows l >= Lout zero) and its
 * transpose XcolT [k][ldp] (operand of dW); either may be NULL. */
int pk_conv_im2col0(const float* x, int64_t ldx, int N, int L, int k, int Lout, void* Xcol, int Kp, void* XcolT,
                    int64_t ldp, void* stream);
/* transposed im2col of a position-major activation for dW: XT [kk*Ci + ci][pos] = A16[pos + kk][ci]. */
int pk_conv_im2col_t(const void* A16, int64_t rows, int Cp, int Ci, int k, void* XT, int64_t ldp, void* stream);
/* fused layer epilogue drop(act(LN_L(max_pool1d(O)))) (:1547-1553, :1652-1661).  O [N*L][ldo] conv output,
 * Lout = L-k+1 valid positions per frame, pool p, Lp = Lout/p; gamma/beta [C][Lp] (LayerNorm over the length
 * axis) or NULL; keep16 [N][Lp][C] = dropout keep/(1-p) or NULL.  Saves P (pooled), arg, stats [N][C][2];
 * writes the next layer's operand A16n [N*Lp][Cpn] and/or the module output Y32 [N][C][Lp]. */
int pk_conv_post_fwd(const float* O, int64_t ldo, int N, int L, int Lout, int p, int Lp, int C, int act,
                     const float* gamma, const float* beta, float eps, const void* keep16, float* P, void* arg,
                     float* stats, void* A16n, int Cpn, float* Y32, void* stream);
/* its backward: dY[n*sn + l*sl + c*sc] -> dO [N*L][C] fp32 (zeros outside the arg-max positions), dgamma /
 * dbeta [C][Lp], dbias [C] (may be NULL), amax of |dO| into amax_bits (may be NULL). */
int pk_conv_post_bwd(const float* dY, int64_t sn, int64_t sl, int64_t sc, int N, int L, int Lout, int p, int Lp,
                     int C, int act, const float* gamma, const float* beta, float eps, const void* keep16,
                     const float* P, const void* arg, const float* stats, float* dgamma, float* dbeta,
                     float* dbias, float* dO, void* amax_bits, void* stream);

/* In place: logits [N][ld] -> log-posteriors (act_fun("softmax") = LogSoftmax(dim=1),
 * neural_networks.py:53-54).  With labels (int64, utils.py:2348-2352): acc[0] = sum_n
 * -logp[n,lab[n]] (nn.NLLLoss numerator, utils.py:2361), acc[1] = #(argmax != lab)
 * (cost_err, utils.py:2379-2380).  acc: 2 doubles, zeroed by the callee. */
int pk_logsoftmax_nll(int N, int S, float* logits, int64_t ld, const int64_t* labels,
                      double* acc, void* stream);

/* Gradient w.r.t. the logits as fp16 GEMM operands scaled by out_scale * (*scale_dev)
 * (row-major d16, channel-major dT16) and dbias [S] (unscaled).  Fused-NLL mode (dlogp NULL):
 * (exp(logp) - onehot(lab)) * gcoef.  General mode: dlogp - exp(logp) * rowsum(dlogp)
 * (rowsum_scratch: N floats). */
int pk_logsoftmax_bwd(int N, int S, const float* logp, int64_t ld, const int64_t* labels,
                      const float* dlogp, int64_t lddl, float gcoef, float out_scale,
                      const float* scale_dev, void* d16, int64_t ld16, void* dT16, int64_t ld16t,
                      float* dbias, float* rowsum_scratch, void* stream);

/* Dense (MLP hidden) layer epilogue: y = drop(act(scale * p + shift)) on the channel-major projection PT
 * (BatchNorm folded into scale/shift by pk_bn_finalize, nn.Dropout's inverted scaling folded into the fp16
 * keep mask keepT = 0 or 1/(1-p), channel-major, or NULL).  Writes YT16 (channel-major), Y16 (row-major) and
 * optionally Y32 (row-major module output).  Replaces drop(act(bn(W x + b))) of neural_networks.py:138-148. */
int pk_dense_act_fwd(int C, int64_t n, int act, const float* PT, int64_t ldp, const float* scale,
                     const float* shift, const void* keepT, int64_t ldk, void* YT16, int64_t ld16t,
                     void* Y16, int64_t ld16r, float* Y32, int64_t ld32, void* stream);
/* its backward up to the BatchNorm: GT16 [C][ldg] = fp16(*gscale * dYT * keep * act'(y)), the GT16 input of
 * pk_bn_bwd (ndir = 1); y is recovered from the saved YT16 / keepT. */
int pk_dense_act_bwd(int C, int64_t n, int act, const float* dYT, int64_t ldy, const void* YT16,
                     int64_t ld16t, const void* keepT, int64_t ldk, const float* gscale, void* GT16,
                     int64_t ldg, void* stream);

/* ---- input side of the path ----
 * Chunk preparation of data_io.load_chunk (data_io.py:255-272) on the device: context-window expansion
 * (out column (lag+left)*F + f = fea[i+left+lag][f], data_io.py:228-241), per-column mean / population-std
 * normalisation (:263, statistics in double), and the label column lab[i+left] - lab_min appended (:266-272; pass
 * lab = NULL for feature-only chunks).  fea [n_in][ldf]; out [n_in-left-right][ldo]; stats: 2*(left+right+1)*F doubles. */
int pk_chunk_prepare(const float* fea, int64_t ldf, const int64_t* lab, int64_t lab_min, int64_t n_in, int F, int left,
                     int right, double* stats, float* out, int64_t ldo, void* stream);
/* Minibatch assembly of core.run_nn (core.py:577-598) without the per-sentence host loop: inp [max_len][B][D];
 * desc = int64 [3][B] on the device: first frame, length and number of leading zero frames of every sentence
 * (the random left padding is drawn by the caller, core.py:592). */
int pk_batch_assemble(const float* data_set, int64_t ldd, int D, const int64_t* desc, int batch_size, int max_len,
                      float* inp, void* stream);

/* Kaldi CompressedMatrix ("CM ") payload -> fp32 on the device (data_io.py:1150-1196): col_headers
 * [cols][4] uint16 percentiles (0/25/75/100), data [cols][rows] uint8 column-major, global min_value / range from the
 * archive header; out [rows][ldo].  Bit-identical to the reference's numpy evaluation. */
int pk_cm_decode(const void* col_headers, const void* data, float min_value, float range, int rows, int cols,
                 float* out, int64_t ldo, void* stream);

/* ---- output side of the path: forward-phase posteriors -> scaled log-likelihoods, in place:
 * logp[n][s] -= log_prior[s] with log_prior = log(counts / sum(counts)) (core.py:664-667); the result is what
 * data_io.write_mat stores as a Kaldi "FM" matrix (data_io.py:1200-1239). */
int pk_sub_log_prior(float* logp, int64_t ld, int64_t n, int S, const float* log_prior, void* stream);

/* torch.optim.Adam (utils.py:2131-2145, amsgrad off): m/v = exponential averages (zero-initialised by the caller),
 * step counts from 1, g is multiplied by gscale (1/world after the allreduce), weight_decay is the L2 term. */
int pk_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                 float eps, float weight_decay, int64_t step, float gscale, void* stream);
/* torch.optim.RMSprop (momentum 0, not centered) / SGD steps over a flat buffer
 * (utils.py:2121-2162, core.py:640-642); gscale multiplies the gradient (1/world_size). */
int pk_rmsprop_step(float* p, const float* g, float* v, int64_t n, float lr, float alpha,
                    float eps, float gscale, void* stream);
int pk_sgd_step(float* p, const float* g, int64_t n, float lr, float gscale, void* stream);

/* ---- the reference's custom LayerNorm (neural_networks.py:23-33: unbiased std, eps added to the std, over the
 * feature axis) as used by MLP layers (`dnn_use_laynorm`, :129-145), on channel-major activations PT[C][ld]
 * (one column per frame).  Forward: in place, XH (optional) receives x_hat, stats[n][2] = (mean, std + eps).
 * Backward: dT16 (fp16, loss-scaled gradient w.r.t. the LayerNorm output) is replaced by the gradient w.r.t. its
 * input, dR16 (optional) receives the row-major copy, dgamma / dbeta are un-scaled with scale[1]. */
int pk_ln_cm_fwd(float* PT, int C, int64_t n, int64_t ld, const float* gamma, const float* beta, float eps, float* XH,
                 float* stats, void* stream);
int pk_ln_cm_bwd(void* dT16, int64_t ld16t, void* dR16, int64_t ld16r, const float* XH, int64_t ld, int C, int64_t n,
                 const float* gamma, const float* stats, float eps, const float* scale, float* dgamma, float* dbeta,
                 float* dbias /* optional: gradient of a bias added in front of the norm */, void* stream);
/* per-channel (sum, sum of squares) over n frames of PT[C][ld]: BatchNorm batch statistics when the GEMM epilogue
 * cannot provide them (BatchNorm applied to a LayerNorm output, :142-143) */
int pk_row_stats(const float* PT, int C, int64_t n, int64_t ld, double* stats, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PK_B200_H_ */
