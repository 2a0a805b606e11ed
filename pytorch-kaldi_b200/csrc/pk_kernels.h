// pk_kernels.h — internal (C++) launch interfaces shared by the .cu files and the C-ABI shim.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define PK_DT_F16 0
#define PK_DT_TF32 2

namespace pk {

struct GemmArgs {
  int dtype = PK_DT_F16;  // operand type of A and B
  int M = 0, N = 0, K = 0;
  const void* A = nullptr;  // [M][lda] K-major
  long long lda = 0;
  const void* B = nullptr;  // [N][ldb] K-major
  long long ldb = 0;
  float* C = nullptr;  // [M][ldc] fp32
  long long ldc = 0;
  const float* bias = nullptr;
  int bias_mode = 0;  // 1: along N, 2: along M
  double* rowstats = nullptr;  // [M][2] (sum, sumsq) accumulated with atomics, or null
  float alpha = 1.f;
  const float* alpha_dev = nullptr;
  int accumulate = 0;  // C += result
  int split_k = 1;
  // contraction range: A[:, a_k0 : a_k0+K] . B[:, b_k0 : b_k0+K]^T; *_kext = valid extent of the
  // operand's K axis (0 -> k0 + K); reads past it return zeros
  long long a_k0 = 0, b_k0 = 0, a_kext = 0, b_kext = 0;
  unsigned int* amax_bits = nullptr;  // optional atomicMax of |C| (float bits)
};
int gemm_tn(const GemmArgs& a, cudaStream_t stream);

struct RecFwdArgs {
  int T = 0, B = 0, H = 0, ndir = 1, act = 0;
  const float* PT = nullptr;  // [G*H][ldp] channel-major pre-activations, col = t*B + b
  long long ldp = 0;
  const float* scale = nullptr;  // [G*H]
  const float* shift = nullptr;  // [G*H]
  const float* U = nullptr;      // [G*H][H] fp32
  const float* mask = nullptr;   // [ndir*B][H] or null
  float mask_scalar = 1.f;
  float* Y32 = nullptr;  // [T*B][ldy32] row-major, col = d*H + u
  long long ldy32 = 0;
  __half* Y16 = nullptr;  // [T*B][ldy16]
  long long ldy16 = 0;
  float* HT = nullptr;     // [ndir*H][ldt] channel-major, natural time
  __half* HT16 = nullptr;  // same, fp16
  __half* HP16 = nullptr;  // fp16 PREVIOUS state h_{k-1} at the same column (operand of dU), zeros at k=0
  float* ZT = nullptr;
  float* HCT = nullptr;
  long long ldt = 0;
  int cluster = 0;  // 0 = auto; > 0 selects a legacy (non-specialised) kernel with that cluster size
  int legacy = 0;   // 0 = auto (faster kernel for this H), 3 = tcgen05 (pk_rnn_tc.cu), 2 = warp-specialised mma.sync
                    // (pk_rnn_ws.cu), 1 = pk_rnn.cu
  int force_z0 = 0; // RNN cell (reference :1438-1447): update gate pinned to 0 -> h = act(a) * mask
  long long* dbg_clk = nullptr;  // bring-up: per-phase cycle sums of CTA 0 / warp 0 (8 slots)
  int sync = -1;    // -1 = default (st.async + mbarrier), 0 = barrier.cluster, 1 = st.async
  int dbg = 0;      // timing experiments only: bit0 skip global stores, bit1 skip global loads
  int groups = 0;   // tcgen05 kernels: arrival-group barriers per buffer (1..3), 0 = default
};
int ligru_fwd(const RecFwdArgs& a, cudaStream_t stream);

struct RecBwdArgs {
  int T = 0, B = 0, H = 0, ndir = 1, act = 0;
  const float* dYT = nullptr;  // [ndir*H][ldt]
  const float* HT = nullptr;
  const float* ZT = nullptr;
  const float* HCT = nullptr;
  long long ldt = 0;
  const float* U = nullptr;
  const float* mask = nullptr;
  float mask_scalar = 1.f;
  const float* gscale = nullptr;  // device scalar: power-of-two loss scale for fp16 operands
  float* GT = nullptr;            // [ndir][2H][ldt] fp32
  __half* GT16 = nullptr;         // [ndir][2H][ldt] fp16, scaled by *gscale
  int cluster = 0;
  int legacy = 0;
  long long* dbg_clk = nullptr;
  int sync = -1;
  int dbg = 0;
  int groups = 0;
};
int ligru_bwd(const RecBwdArgs& a, cudaStream_t stream);
// warp-specialised variants (pk_rnn_ws.cu); the bwd one writes GT16 only
int ligru_fwd_ws(const RecFwdArgs& a, cudaStream_t stream);
int ligru_bwd_ws(const RecBwdArgs& a, cudaStream_t stream);
// tcgen05 variants (pk_rnn_tc.cu): weights stationary in tensor memory, H <= ligru_tc_max_hidden()
int ligru_fwd_tc(const RecFwdArgs& a, cudaStream_t stream);
int ligru_bwd_tc(const RecBwdArgs& a, cudaStream_t stream);
int ligru_tc_max_hidden();
void set_debug_clock_buffer_tc(long long* dev_ptr);
void set_debug_clock_buffer(long long* dev_ptr);

// ---- step-wise recurrent path (pk_cell_step.cu): one fused kernel per time step ----
enum CellKind { CELL_LIGRU = 0, CELL_RNN = 1, CELL_GRU = 2, CELL_MGRU = 3, CELL_LSTM = 4 };
// saved tensors SV[0..4], all [ndir*H][ldt] fp32 channel-major, natural time:
//   liGRU: z, hcand            LSTM: f, g (= act(c~)*mask), i, o, c
//   GRU:   z, hcand, r         minimalGRU: z, hcand
struct CellStepFwdArgs {
  int cell = 0, T = 0, B = 0, H = 0, ndir = 1, act = 0;
  const float* PT = nullptr;  // [NG*H][ldp]
  long long ldp = 0;
  const float* scale = nullptr;
  const float* shift = nullptr;
  const float* U = nullptr;  // [NG*H][H] fp32, gate order of the saved-tensor list above
  const float* mask = nullptr;
  float mask_scalar = 1.f;
  float* Y32 = nullptr;
  long long ldy32 = 0;
  __half* Y16 = nullptr;
  long long ldy16 = 0;
  float* HT = nullptr;
  __half* HT16 = nullptr;
  __half* HP16 = nullptr;
  __half* HX16 = nullptr;  // GRU / minimalGRU: fp16 (r*h_{k-1}) / (z*h_{k-1}), the operand of dUh
  float* SV[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  long long ldt = 0;
  void* workspace = nullptr;
  long long workspace_bytes = 0;
};
struct CellStepBwdArgs {
  int cell = 0, T = 0, B = 0, H = 0, ndir = 1, act = 0;
  const float* dYT = nullptr;
  const float* HT = nullptr;
  const float* SV[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  long long ldt = 0;
  const float* U = nullptr;
  const float* mask = nullptr;
  float mask_scalar = 1.f;
  const float* gscale = nullptr;
  __half* GT16 = nullptr;  // [ndir][NG*H][ldt] fp16 scaled
  void* workspace = nullptr;
  long long workspace_bytes = 0;
};
long long cell_step_workspace_bytes(int cell, int T, int B, int H, int ndir, int backward);
// kernel launches one cell_step_fwd / cell_step_bwd call performs for this shape (bench.py's gpu_launches claim)
int cell_step_launches(int cell, int T, int B, int H, int ndir, int backward);
// cluster-persistent LSTM kernels (pk_cell_cluster.cu): DSMEM state exchange instead of L2 + grid barrier, H <= 560
bool lstm_cluster_usable(int cell, int H);
long long lstm_cluster_pack_bytes(int H);
int lstm_cluster_fwd(const CellStepFwdArgs& a, __half* Wc, cudaStream_t stream);
int lstm_cluster_bwd(const CellStepBwdArgs& a, __half* Wc, cudaStream_t stream);
// cluster-persistent GRU / minimalGRU kernels (pk_cell_cluster2.cu): two exchanges per step
bool gru_cluster_usable(int cell, int H);
long long gru_cluster_pack_bytes(int cell, int H);
int gru_cluster_fwd(const CellStepFwdArgs& a, __half* Wc, cudaStream_t stream);
int gru_cluster_bwd(const CellStepBwdArgs& a, __half* Wc, cudaStream_t stream);
int cell_step_fwd(const CellStepFwdArgs& a, cudaStream_t stream);
int cell_step_bwd(const CellStepBwdArgs& a, cudaStream_t stream);

// ---- conv front-ends (pk_conv.cu): CNN / SincNet companions of the tcgen05 GEMM ----
int rowln_fwd(const float* x, long long ldx, int N, int L, const float* gamma, const float* beta, float eps, float* y,
              float* stats, cudaStream_t stream);
int conv_ln0_bwd(const float* G, long long ldg, int N, int L, int Lout, int k, const float* x, long long ldx,
                 const float* stats, float* dgamma, float* dbeta, cudaStream_t stream);
int sinc_filters_fwd(const float* low_hz_, const float* band_hz_, int C, int k, float sr, float min_low, float min_band,
                     float* filt, cudaStream_t stream);
int sinc_filters_bwd(const float* low_hz_, const float* band_hz_, int C, int k, float sr, float min_low, float min_band,
                     const float* dfilt, float* dlow, float* dband, cudaStream_t stream);
int conv_pack_weights(const float* w, int Co, int Ci, int k, __half* W16, int Cip, long long ldw, __half* Wf16, int Cop,
                      long long ldf, cudaStream_t stream);
int conv_im2col0(const float* x, long long ldx, int N, int L, int k, int Lout, __half* Xcol, int Kp, __half* XcolT,
                 long long ldp, cudaStream_t stream);
int conv_im2colT(const __half* A16, long long rows, int Cp, int Ci, int k, __half* XT, long long ldp, cudaStream_t stream);
struct ConvPostFwdArgs {
  const float* O = nullptr;  // [N*L][ldo] conv output (+bias), position-major
  long long ldo = 0;
  int N = 0, L = 0, Lout = 0, p = 1, Lp = 0, C = 0, act = 0;
  const float* gamma = nullptr;  // [C][Lp] LayerNorm affine, or null (no normalisation)
  const float* beta = nullptr;
  float eps = 1e-6f;
  const __half* keep = nullptr;  // [N][Lp][C] dropout keep * 1/(1-p), or null
  float* P = nullptr;            // [N][Lp][C] pooled values (saved)
  uint8_t* arg = nullptr;        // [N][Lp][C] arg-max inside the pooling window (saved)
  float* stats = nullptr;        // [N][C][2] mean, 1/(std+eps) (saved)
  __half* A16n = nullptr;        // [N*Lp (+tail)][Cpn] next layer's fp16 operand, or null
  int Cpn = 0;
  float* Y32 = nullptr;          // [N][C][Lp] module output (last layer), or null
};
int conv_post_fwd(const ConvPostFwdArgs& a, cudaStream_t stream);
struct ConvPostBwdArgs {
  const float* dY = nullptr;  // dY[n*sn + l*sl + c*sc]
  long long sn = 0, sl = 0, sc = 0;
  int N = 0, L = 0, Lout = 0, p = 1, Lp = 0, C = 0, act = 0;
  const float* gamma = nullptr;
  const float* beta = nullptr;
  float eps = 1e-6f;
  const __half* keep = nullptr;
  const float* P = nullptr;
  const uint8_t* arg = nullptr;
  const float* stats = nullptr;
  float* dgamma = nullptr;  // [C][Lp]
  float* dbeta = nullptr;
  float* dbias = nullptr;   // [C] or null
  float* dO = nullptr;      // [N*L][C] gradient w.r.t. the conv output (zeros outside the arg-max positions)
  unsigned int* amax_bits = nullptr;
};
int conv_post_bwd(const ConvPostBwdArgs& a, cudaStream_t stream);

// ---- FusionLinearConv companions (pk_fusion.cu): activation + reduction over the microphone channels ----
int fusion_reduce_fwd(const float* O, long long ldo, long long N, int M, int C, int Hh, int mode, const float* slopes,
                      float red, float* P, long long ldp, cudaStream_t stream);
int fusion_reduce_bwd(const float* dP, long long lddp, const float* O, long long ldo, long long N, int M, int C, int Hh,
                      int mode, const float* slopes, float red, float* dO, long long lddo, float* dbias, float* dslope,
                      cudaStream_t stream);

// ---- memory-bound helpers (pk_elementwise.cu) ----
// out[c][r] = in[r][c]; optional fp16 copies. in is [R][ldi] fp32.
int transpose_f32(const float* in, long long ldi, int R, int C, float* outT, long long ldo,
                  __half* outT16, long long ldo16, __half* in16, long long ldi16, const float* scale_dev,
                  unsigned int* amax_bits, cudaStream_t stream);
// scale_out[0] = 2^k, [1] = 2^-k from an amax accumulated by a producer; re-zeroes the accumulator
int amax_finalize(unsigned int* amax_bits, float target_log2, float* scale_out, cudaStream_t stream);
int convert_f16(const float* in, long long ldi, int R, int C, __half* out, long long ldo,
                const float* scale_dev, cudaStream_t stream);
// scale_out[0] = 2^k, scale_out[1] = 2^-k
int amax_scale(const float* x, long long ld, int R, int C, float target_log2, float* amax_scratch,
               float* scale_out, cudaStream_t stream);
int bn_finalize(const double* stats, int C, long long n_unique, long long n_ref, const float* gamma,
                const float* beta, float eps, float momentum, int training, float* running_mean,
                float* running_var, long long* num_batches, float* scale, float* shift, float* mean_out,
                float* rstd_out, cudaStream_t stream);
int fill_scale_shift(const float* bias, int C, float* scale, float* shift, cudaStream_t stream);
struct BnBwdArgs {
  int C = 0;          // channels (G*H)
  int ndir = 1;
  long long n = 0;    // unique rows T*B
  const float* GT = nullptr;   // [ndir][C][ldt] fp32 (unscaled), or null -> GT16
  const __half* GT16 = nullptr; // [ndir][C][ldt] fp16 scaled by *gscale
  const float* PT = nullptr;   // [C][ldp] pre-BN projections
  long long ldt = 0, ldp = 0;
  int use_bn = 1;
  int training = 1;
  const float* mean = nullptr;   // [C] batch (training) or running (eval)
  const float* rstd = nullptr;   // [C]
  const float* gamma = nullptr;  // [C]
  const float* gscale = nullptr; // device loss scale applied to the fp16 outputs
  float* dgamma = nullptr;       // [C]
  float* dbeta = nullptr;        // [C]  (= bias gradient when !use_bn)
  __half* dPT16 = nullptr;       // [C][ld16t]    channel-major, scaled
  long long ld16t = 0;
  __half* dP16 = nullptr;        // [n][ld16r]    row-major, scaled
  long long ld16r = 0;
  double* sums_scratch = nullptr; // [2][C]
};
int bn_bwd(const BnBwdArgs& a, cudaStream_t stream);

struct HeadFwdArgs {
  int N = 0, S = 0;
  float* logits = nullptr;  // [N][ld] in: logits, out: log-posteriors (in place)
  long long ld = 0;
  const long long* labels = nullptr;  // [N] or null
  double* acc = nullptr;              // [2] device accumulators: NLL sum, error count (zeroed by callee)
};
int logsoftmax_nll(const HeadFwdArgs& a, cudaStream_t stream);
struct HeadBwdArgs {
  int N = 0, S = 0;
  const float* logp = nullptr;  // [N][ld]
  long long ld = 0;
  const long long* labels = nullptr;  // fused-NLL mode (dlogp == null): d = (exp(logp) - onehot) * gcoef
  const float* dlogp = nullptr;       // general mode: d = dlogp - exp(logp) * rowsum(dlogp)
  long long lddl = 0;
  float gcoef = 1.f;                  // upstream grad / N for fused mode
  float out_scale = 1.f;              // fp16 loss scale (host value, power of two)
  const float* scale_dev = nullptr;   // optional extra device-side scale factor
  __half* d16 = nullptr;              // [N][ld16] row-major scaled
  long long ld16 = 0;
  __half* dT16 = nullptr;             // [S][ld16t] channel-major scaled
  long long ld16t = 0;
  float* dbias = nullptr;             // [S] column sums (unscaled)
  float* rowsum_scratch = nullptr;    // [N] (general mode only)
};
int logsoftmax_bwd(const HeadBwdArgs& a, cudaStream_t stream);
struct DenseFwdArgs {
  int C = 0, act = 0;
  long long n = 0;
  const float* PT = nullptr;  // [C][ldp]
  long long ldp = 0;
  const float* scale = nullptr;
  const float* shift = nullptr;
  const __half* keepT = nullptr;  // [C][ldk] 0 or 1/(1-p), or null
  long long ldk = 0;
  __half* YT16 = nullptr;  // [C][ld16t]
  long long ld16t = 0;
  __half* Y16 = nullptr;  // [n][ld16r]
  long long ld16r = 0;
  float* Y32 = nullptr;  // [n][ld32]
  long long ld32 = 0;
};
int dense_act_fwd(const DenseFwdArgs& a, cudaStream_t stream);
struct DenseBwdArgs {
  int C = 0, act = 0;
  long long n = 0;
  const float* dYT = nullptr;  // [C][ldy]
  long long ldy = 0;
  const __half* YT16 = nullptr;
  long long ld16t = 0;
  const __half* keepT = nullptr;
  long long ldk = 0;
  const float* gscale = nullptr;
  __half* GT16 = nullptr;  // [C][ldg]
  long long ldg = 0;
};
int dense_act_bwd(const DenseBwdArgs& a, cudaStream_t stream);
int rmsprop_step(float* p, const float* g, float* v, long long n, float lr, float alpha, float eps,
                 float gscale, cudaStream_t stream);
int sgd_step(float* p, const float* g, long long n, float lr, float gscale, cudaStream_t stream);
int chunk_prepare(const float* fea, long long ldf, const long long* lab, long long lab_min, long long n_in, int F, int left,
                  int right, double* stats, float* out, long long ldo, cudaStream_t stream);
int batch_assemble(const float* data, long long ldd, int D, const long long* desc, int Bsz, int max_len, float* inp,
                   cudaStream_t stream);
int cm_decode(const uint16_t* hdr, const uint8_t* data, float gmin, float grange, int rows, int cols, float* out, long long ldo,
              cudaStream_t stream);
int rows_sub_vec(float* x, long long ld, long long n, int C, const float* v, cudaStream_t stream);
int adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float wd,
              long long step, float gscale, cudaStream_t stream);

// reference LayerNorm (unbiased std, eps added to the std) over the feature axis of channel-major activations
int ln_cm_fwd(float* PT, int C, long long n, long long ld, const float* gamma, const float* beta, float eps, float* XH,
              float* stats, cudaStream_t stream);
int ln_cm_bwd(__half* dT16, long long ld16t, __half* dR16, long long ld16r, const float* XH, long long ld, int C, long long n,
              const float* gamma, const float* stats, float eps, const float* scale, float* dgamma, float* dbeta,
              float* dbias, cudaStream_t stream);
int row_stats(const float* PT, int C, long long n, long long ld, double* stats, cudaStream_t stream);

}  // namespace pk
