// pk_api.cu — extern "C" shim: validates arguments, forwards to the kernels, never throws.
#include "../../include/pk_b200.h"

#include "pk_common.cuh"
#include "pk_kernels.h"

#include <cstdarg>
#include <cstdio>

namespace pk {
namespace {
thread_local char g_err[1024] = "";
}
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace pk

using namespace pk;

namespace {
struct RecFlags {
  int cell, cluster, sync, dbg, legacy, groups;
};
RecFlags parse_cell(int cell) {
  RecFlags f;
  f.cluster = 0;
  f.sync = -1;
  f.dbg = ((cell & PK_REC_DBG_NOSTORE) ? 1 : 0) | ((cell & PK_REC_DBG_NOLOAD) ? 2 : 0) | ((cell & PK_REC_DBG_NOPROXYFENCE) ? 4 : 0) |
          ((cell & PK_REC_DBG_BLOCKINGWAIT) ? 8 : 0) | ((cell & PK_REC_BWD_ALLGATHER) ? 16 : 0) |
          ((cell & PK_REC_BWD_KSPLIT) ? 32 : 0);
  f.legacy = (cell & PK_REC_WS) ? 2 : ((cell & PK_REC_TC) ? 3 : 0);
  f.groups = (cell >> 19) & 3;
  f.cell = cell & PK_CELL_MASK;
  return f;
}
}  // namespace

extern "C" {

const char* pk_last_error(void) { return pk::g_err; }
int pk_version(void) { return PK_ABI_VERSION; }
// bring-up hook (not part of the documented ABI): per-phase cycle counters of the recurrent kernels
void pk_debug_set_clock_buffer(void* dev_ptr) { pk::set_debug_clock_buffer(static_cast<long long*>(dev_ptr)); }

int pk_gemm_tn(int dtype, int M, int N, int K, const void* A, int64_t lda, int64_t a_k0, int64_t a_kext,
               const void* B, int64_t ldb, int64_t b_k0, int64_t b_kext, float* C, int64_t ldc,
               const float* bias, int bias_mode, double* rowstats, float alpha, const float* alpha_dev,
               int accumulate, int split_k, void* amax_bits, void* stream) {
  PK_REQUIRE(A && B && C, "pk_gemm_tn: null operand");
  GemmArgs a;
  a.dtype = dtype; a.M = M; a.N = N; a.K = K;
  a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.C = C; a.ldc = ldc;
  a.a_k0 = a_k0; a.a_kext = a_kext; a.b_k0 = b_k0; a.b_kext = b_kext;
  a.bias = bias; a.bias_mode = bias_mode; a.rowstats = rowstats;
  a.alpha = alpha; a.alpha_dev = alpha_dev; a.accumulate = accumulate; a.split_k = split_k;
  a.amax_bits = static_cast<unsigned int*>(amax_bits);
  return gemm_tn(a, static_cast<cudaStream_t>(stream));
}

int pk_transpose_f32(const float* in, int64_t ldi, int R, int C, float* outT, int64_t ldo, void* outT16,
                     int64_t ldo16, void* in16, int64_t ldi16, const float* scale_dev, void* amax_bits,
                     void* stream) {
  PK_REQUIRE(in != nullptr, "pk_transpose_f32: null input");
  return transpose_f32(in, ldi, R, C, outT, ldo, static_cast<__half*>(outT16), ldo16,
                       static_cast<__half*>(in16), ldi16, scale_dev, static_cast<unsigned int*>(amax_bits),
                       static_cast<cudaStream_t>(stream));
}

int pk_convert_f16(const float* in, int64_t ldi, int R, int C, void* out, int64_t ldo, const float* scale_dev,
                   void* stream) {
  PK_REQUIRE(in && out, "pk_convert_f16: null pointer");
  return convert_f16(in, ldi, R, C, static_cast<__half*>(out), ldo, scale_dev,
                     static_cast<cudaStream_t>(stream));
}

int pk_amax_scale(const float* x, int64_t ld, int R, int C, float target_log2, float* amax_scratch,
                  float* scale_out, void* stream) {
  PK_REQUIRE(x && amax_scratch && scale_out, "pk_amax_scale: null pointer");
  return amax_scale(x, ld, R, C, target_log2, amax_scratch, scale_out, static_cast<cudaStream_t>(stream));
}

int pk_amax_finalize(void* amax_bits, float target_log2, float* scale_out, void* stream) {
  PK_REQUIRE(amax_bits && scale_out, "pk_amax_finalize: null pointer");
  return amax_finalize(static_cast<unsigned int*>(amax_bits), target_log2, scale_out, static_cast<cudaStream_t>(stream));
}

int pk_bn_finalize(const double* stats, int C, int64_t n_unique, int64_t n_ref, const float* gamma,
                   const float* beta, float eps, float momentum, int training, float* running_mean,
                   float* running_var, int64_t* num_batches, float* scale, float* shift, float* mean_out,
                   float* rstd_out, void* stream) {
  PK_REQUIRE(scale && shift, "pk_bn_finalize: null output");
  PK_REQUIRE(!training || stats, "pk_bn_finalize: training mode needs stats");
  return bn_finalize(stats, C, n_unique, n_ref, gamma, beta, eps, momentum, training, running_mean,
                     running_var, reinterpret_cast<long long*>(num_batches), scale, shift, mean_out, rstd_out,
                     static_cast<cudaStream_t>(stream));
}

int pk_fill_scale_shift(const float* bias, int C, float* scale, float* shift, void* stream) {
  PK_REQUIRE(scale && shift, "pk_fill_scale_shift: null output");
  return fill_scale_shift(bias, C, scale, shift, static_cast<cudaStream_t>(stream));
}

int pk_bn_bwd(int C, int ndir, int64_t n, const float* GT, const void* GT16, int64_t ldt, const float* PT, int64_t ldp, int use_bn,
              int training, const float* mean, const float* rstd, const float* gamma, const float* gscale,
              float* dgamma, float* dbeta, void* dPT16, int64_t ld16t, void* dP16, int64_t ld16r,
              double* sums_scratch, void* stream) {
  PK_REQUIRE(GT != nullptr || GT16 != nullptr, "pk_bn_bwd: need GT or GT16");
  PK_REQUIRE(!use_bn || (rstd && (!training || (PT && mean))), "pk_bn_bwd: BatchNorm mode needs PT/mean/rstd");
  BnBwdArgs a;
  a.C = C; a.ndir = ndir; a.n = n; a.GT = GT; a.GT16 = static_cast<const __half*>(GT16); a.ldt = ldt; a.PT = PT; a.ldp = ldp;
  a.use_bn = use_bn; a.training = training; a.mean = mean; a.rstd = rstd; a.gamma = gamma; a.gscale = gscale;
  a.dgamma = dgamma; a.dbeta = dbeta;
  a.dPT16 = static_cast<__half*>(dPT16); a.ld16t = ld16t;
  a.dP16 = static_cast<__half*>(dP16); a.ld16r = ld16r;
  a.sums_scratch = sums_scratch;
  return bn_bwd(a, static_cast<cudaStream_t>(stream));
}

int pk_rnn_layer_fwd(int cell, int T, int B, int H, int ndir, int act, const float* PT, int64_t ldp,
                     const float* scale, const float* shift, const float* U, const float* mask,
                     float mask_scalar, float* Y32, int64_t ldy32, void* Y16, int64_t ldy16, float* HT,
                     void* HT16, void* HP16, float* ZT, float* HCT, int64_t ldt, void* stream) {
  const RecFlags f = parse_cell(cell);
  PK_REQUIRE(f.cell == PK_CELL_LIGRU || f.cell == PK_CELL_RNN, "pk_rnn_layer_fwd: cell kind %d not implemented", f.cell);
  PK_REQUIRE(PT && scale && shift && U, "pk_rnn_layer_fwd: null input");
  PK_REQUIRE(act >= PK_ACT_RELU && act <= PK_ACT_LINEAR, "pk_rnn_layer_fwd: bad activation %d", act);
  RecFwdArgs a;
  a.T = T; a.B = B; a.H = H; a.ndir = ndir; a.act = act;
  a.PT = PT; a.ldp = ldp; a.scale = scale; a.shift = shift; a.U = U; a.mask = mask; a.mask_scalar = mask_scalar;
  a.Y32 = Y32; a.ldy32 = ldy32; a.Y16 = static_cast<__half*>(Y16); a.ldy16 = ldy16;
  a.HT = HT; a.HT16 = static_cast<__half*>(HT16); a.HP16 = static_cast<__half*>(HP16);
  a.ZT = ZT; a.HCT = HCT; a.ldt = ldt;
  a.cluster = f.cluster; a.sync = f.sync; a.dbg = f.dbg; a.legacy = f.legacy; a.groups = f.groups;
  a.force_z0 = (f.cell == PK_CELL_RNN) ? 1 : 0;
  return ligru_fwd(a, static_cast<cudaStream_t>(stream));
}

int pk_rnn_layer_bwd(int cell, int T, int B, int H, int ndir, int act, const float* dYT, const float* HT,
                     const float* ZT, const float* HCT, int64_t ldt, const float* U, const float* mask,
                     float mask_scalar, const float* gscale, float* GT, void* GT16, void* stream) {
  const RecFlags f = parse_cell(cell);
  PK_REQUIRE(f.cell == PK_CELL_LIGRU || f.cell == PK_CELL_RNN, "pk_rnn_layer_bwd: cell kind %d not implemented", f.cell);
  PK_REQUIRE(dYT && HT && ZT && HCT && U && (GT || GT16), "pk_rnn_layer_bwd: null input");
  RecBwdArgs a;
  a.T = T; a.B = B; a.H = H; a.ndir = ndir; a.act = act;
  a.dYT = dYT; a.HT = HT; a.ZT = ZT; a.HCT = HCT; a.ldt = ldt; a.U = U; a.mask = mask;
  a.mask_scalar = mask_scalar; a.gscale = gscale; a.GT = GT; a.GT16 = static_cast<__half*>(GT16);
  a.cluster = f.cluster; a.sync = f.sync; a.dbg = f.dbg; a.legacy = f.legacy; a.groups = f.groups;
  return ligru_bwd(a, static_cast<cudaStream_t>(stream));
}

int64_t pk_rnn_step_workspace_bytes(int cell, int T, int B, int H, int ndir, int backward) {
  return cell_step_workspace_bytes(cell & PK_CELL_MASK, T, B, H, ndir, backward);
}

int pk_rnn_step_launches(int cell, int T, int B, int H, int ndir, int backward) {
  return cell_step_launches(cell & PK_CELL_MASK, T, B, H, ndir, backward);
}

int pk_rnn_step_is_cluster(int cell, int H) {
  return (lstm_cluster_usable(cell & PK_CELL_MASK, H) || gru_cluster_usable(cell & PK_CELL_MASK, H)) ? 1 : 0;
}

int pk_rnn_step_fwd(int cell, int T, int B, int H, int ndir, int act, const float* PT, int64_t ldp,
                    const float* scale, const float* shift, const float* U, const float* mask, float mask_scalar,
                    float* Y32, int64_t ldy32, void* Y16, int64_t ldy16, float* HT, void* HT16, void* HP16,
                    void* HX16, float* sv0, float* sv1, float* sv2, float* sv3, float* sv4, int64_t ldt, void* workspace,
                    int64_t workspace_bytes, void* stream) {
  PK_REQUIRE(PT && scale && shift && U, "pk_rnn_step_fwd: null input");
  PK_REQUIRE(act >= PK_ACT_RELU && act <= PK_ACT_LINEAR, "pk_rnn_step_fwd: bad activation %d", act);
  PK_REQUIRE(T > 0 && B > 0 && H > 0 && (ndir == 1 || ndir == 2), "pk_rnn_step_fwd: bad shape");
  CellStepFwdArgs a;
  a.cell = cell & PK_CELL_MASK; a.T = T; a.B = B; a.H = H; a.ndir = ndir; a.act = act;
  a.PT = PT; a.ldp = ldp; a.scale = scale; a.shift = shift; a.U = U; a.mask = mask; a.mask_scalar = mask_scalar;
  a.Y32 = Y32; a.ldy32 = ldy32; a.Y16 = static_cast<__half*>(Y16); a.ldy16 = ldy16;
  a.HT = HT; a.HT16 = static_cast<__half*>(HT16); a.HP16 = static_cast<__half*>(HP16);
  a.HX16 = static_cast<__half*>(HX16);
  a.SV[0] = sv0; a.SV[1] = sv1; a.SV[2] = sv2; a.SV[3] = sv3; a.SV[4] = sv4; a.ldt = ldt;
  a.workspace = workspace; a.workspace_bytes = workspace_bytes;
  return cell_step_fwd(a, static_cast<cudaStream_t>(stream));
}

int pk_rnn_step_bwd(int cell, int T, int B, int H, int ndir, int act, const float* dYT, const float* HT,
                    const float* sv0, const float* sv1, const float* sv2, const float* sv3, const float* sv4,
                    int64_t ldt, const float* U, const float* mask, float mask_scalar, const float* gscale,
                    void* GT16, void* workspace, int64_t workspace_bytes, void* stream) {
  PK_REQUIRE(dYT && HT && sv0 && sv1 && U && GT16, "pk_rnn_step_bwd: null input");
  PK_REQUIRE((cell & PK_CELL_MASK) != PK_CELL_LSTM || (sv2 && sv3 && sv4), "pk_rnn_step_bwd: LSTM needs sv0..sv4");
  PK_REQUIRE((cell & PK_CELL_MASK) != PK_CELL_GRU || sv2, "pk_rnn_step_bwd: GRU needs sv2 (reset gate)");
  CellStepBwdArgs a;
  a.cell = cell & PK_CELL_MASK; a.T = T; a.B = B; a.H = H; a.ndir = ndir; a.act = act;
  a.dYT = dYT; a.HT = HT; a.SV[0] = sv0; a.SV[1] = sv1; a.SV[2] = sv2; a.SV[3] = sv3; a.SV[4] = sv4;
  a.ldt = ldt; a.U = U; a.mask = mask; a.mask_scalar = mask_scalar; a.gscale = gscale;
  a.GT16 = static_cast<__half*>(GT16); a.workspace = workspace; a.workspace_bytes = workspace_bytes;
  return cell_step_bwd(a, static_cast<cudaStream_t>(stream));
}

int pk_fusion_reduce_fwd(const float* O, int64_t ldo, int64_t N, int M, int C, int Hh, int mode, const float* slopes,
                         float red, float* P, int64_t ldp, void* stream) {
  PK_REQUIRE(O && P && N > 0 && M > 0 && C > 0 && Hh > 0 && C % Hh == 0, "pk_fusion_reduce_fwd: bad arguments");
  PK_REQUIRE(mode == 1 || (mode == 0 && slopes), "pk_fusion_reduce_fwd: mode 0 (piecewise linear) needs the slopes, mode 1 = tanh");
  return fusion_reduce_fwd(O, ldo, N, M, C, Hh, mode, slopes, red, P, ldp, static_cast<cudaStream_t>(stream));
}
int pk_fusion_reduce_bwd(const float* dP, int64_t lddp, const float* O, int64_t ldo, int64_t N, int M, int C, int Hh,
                         int mode, const float* slopes, float red, float* dO, int64_t lddo, float* dbias, float* dslope,
                         void* stream) {
  PK_REQUIRE(dP && O && dO && dbias && N > 0 && M > 0 && C > 0 && Hh > 0 && C % Hh == 0, "pk_fusion_reduce_bwd: bad arguments");
  PK_REQUIRE(mode == 1 || (mode == 0 && slopes), "pk_fusion_reduce_bwd: mode 0 (piecewise linear) needs the slopes, mode 1 = tanh");
  return fusion_reduce_bwd(dP, lddp, O, ldo, N, M, C, Hh, mode, slopes, red, dO, lddo, dbias, dslope,
                           static_cast<cudaStream_t>(stream));
}

int pk_rowln_fwd(const float* x, int64_t ldx, int N, int L, const float* gamma, const float* beta, float eps, float* y,
                 float* stats, void* stream) {
  PK_REQUIRE(x && gamma && beta && y && stats, "pk_rowln_fwd: null pointer");
  return rowln_fwd(x, ldx, N, L, gamma, beta, eps, y, stats, static_cast<cudaStream_t>(stream));
}
int pk_conv_ln0_bwd(const float* G, int64_t ldg, int N, int L, int Lout, int k, const float* x, int64_t ldx,
                    const float* stats, float* dgamma, float* dbeta, void* stream) {
  PK_REQUIRE(G && x && stats && dgamma && dbeta, "pk_conv_ln0_bwd: null pointer");
  return conv_ln0_bwd(G, ldg, N, L, Lout, k, x, ldx, stats, dgamma, dbeta, static_cast<cudaStream_t>(stream));
}
int pk_sinc_filters_fwd(const float* low_hz_, const float* band_hz_, int C, int k, float sample_rate, float min_low_hz,
                        float min_band_hz, float* filt, void* stream) {
  PK_REQUIRE(low_hz_ && band_hz_ && filt, "pk_sinc_filters_fwd: null pointer");
  return sinc_filters_fwd(low_hz_, band_hz_, C, k, sample_rate, min_low_hz, min_band_hz, filt,
                          static_cast<cudaStream_t>(stream));
}
int pk_sinc_filters_bwd(const float* low_hz_, const float* band_hz_, int C, int k, float sample_rate, float min_low_hz,
                        float min_band_hz, const float* dfilt, float* dlow, float* dband, void* stream) {
  PK_REQUIRE(low_hz_ && band_hz_ && dfilt && dlow && dband, "pk_sinc_filters_bwd: null pointer");
  return sinc_filters_bwd(low_hz_, band_hz_, C, k, sample_rate, min_low_hz, min_band_hz, dfilt, dlow, dband,
                          static_cast<cudaStream_t>(stream));
}
int pk_conv_pack_weights(const float* w, int Co, int Ci, int k, void* W16, int Cip, int64_t ldw, void* Wflip16, int Cop,
                         int64_t ldf, void* stream) {
  return conv_pack_weights(w, Co, Ci, k, static_cast<__half*>(W16), Cip, ldw, static_cast<__half*>(Wflip16), Cop, ldf,
                           static_cast<cudaStream_t>(stream));
}
int pk_conv_im2col0(const float* x, int64_t ldx, int N, int L, int k, int Lout, void* Xcol, int Kp, void* XcolT,
                    int64_t ldp, void* stream) {
  PK_REQUIRE(x && (Xcol || XcolT), "pk_conv_im2col0: null pointer");
  return conv_im2col0(x, ldx, N, L, k, Lout, static_cast<__half*>(Xcol), Kp, static_cast<__half*>(XcolT), ldp,
                      static_cast<cudaStream_t>(stream));
}
int pk_conv_im2col_t(const void* A16, int64_t rows, int Cp, int Ci, int k, void* XT, int64_t ldp, void* stream) {
  PK_REQUIRE(A16 && XT, "pk_conv_im2col_t: null pointer");
  return conv_im2colT(static_cast<const __half*>(A16), rows, Cp, Ci, k, static_cast<__half*>(XT), ldp,
                      static_cast<cudaStream_t>(stream));
}
int pk_conv_post_fwd(const float* O, int64_t ldo, int N, int L, int Lout, int p, int Lp, int C, int act,
                     const float* gamma, const float* beta, float eps, const void* keep16, float* P, void* arg,
                     float* stats, void* A16n, int Cpn, float* Y32, void* stream) {
  PK_REQUIRE(O && P && arg && (gamma == nullptr || (beta && stats)), "pk_conv_post_fwd: null pointer");
  PK_REQUIRE(act >= PK_ACT_RELU && act <= PK_ACT_LINEAR, "pk_conv_post_fwd: bad activation %d", act);
  ConvPostFwdArgs a;
  a.O = O; a.ldo = ldo; a.N = N; a.L = L; a.Lout = Lout; a.p = p; a.Lp = Lp; a.C = C; a.act = act;
  a.gamma = gamma; a.beta = beta; a.eps = eps; a.keep = static_cast<const __half*>(keep16);
  a.P = P; a.arg = static_cast<uint8_t*>(arg); a.stats = stats;
  a.A16n = static_cast<__half*>(A16n); a.Cpn = Cpn; a.Y32 = Y32;
  return conv_post_fwd(a, static_cast<cudaStream_t>(stream));
}
int pk_conv_post_bwd(const float* dY, int64_t sn, int64_t sl, int64_t sc, int N, int L, int Lout, int p, int Lp, int C,
                     int act, const float* gamma, const float* beta, float eps, const void* keep16, const float* P,
                     const void* arg, const float* stats, float* dgamma, float* dbeta, float* dbias, float* dO,
                     void* amax_bits, void* stream) {
  PK_REQUIRE(dY && P && arg && dO && (gamma == nullptr || (beta && stats && dgamma && dbeta)),
             "pk_conv_post_bwd: null pointer");
  ConvPostBwdArgs a;
  a.dY = dY; a.sn = sn; a.sl = sl; a.sc = sc;
  a.N = N; a.L = L; a.Lout = Lout; a.p = p; a.Lp = Lp; a.C = C; a.act = act;
  a.gamma = gamma; a.beta = beta; a.eps = eps; a.keep = static_cast<const __half*>(keep16);
  a.P = P; a.arg = static_cast<const uint8_t*>(arg); a.stats = stats;
  a.dgamma = dgamma; a.dbeta = dbeta; a.dbias = dbias; a.dO = dO; a.amax_bits = static_cast<unsigned int*>(amax_bits);
  return conv_post_bwd(a, static_cast<cudaStream_t>(stream));
}

int pk_logsoftmax_nll(int N, int S, float* logits, int64_t ld, const int64_t* labels, double* acc,
                      void* stream) {
  PK_REQUIRE(logits != nullptr, "pk_logsoftmax_nll: null logits");
  HeadFwdArgs a;
  a.N = N; a.S = S; a.logits = logits; a.ld = ld;
  a.labels = reinterpret_cast<const long long*>(labels); a.acc = acc;
  return logsoftmax_nll(a, static_cast<cudaStream_t>(stream));
}

int pk_logsoftmax_bwd(int N, int S, const float* logp, int64_t ld, const int64_t* labels, const float* dlogp,
                      int64_t lddl, float gcoef, float out_scale, const float* scale_dev, void* d16,
                      int64_t ld16, void* dT16, int64_t ld16t, float* dbias, float* rowsum_scratch,
                      void* stream) {
  PK_REQUIRE(logp != nullptr, "pk_logsoftmax_bwd: null logp");
  HeadBwdArgs a;
  a.N = N; a.S = S; a.logp = logp; a.ld = ld; a.labels = reinterpret_cast<const long long*>(labels);
  a.dlogp = dlogp; a.lddl = lddl; a.gcoef = gcoef; a.out_scale = out_scale; a.scale_dev = scale_dev;
  a.d16 = static_cast<__half*>(d16); a.ld16 = ld16; a.dT16 = static_cast<__half*>(dT16); a.ld16t = ld16t;
  a.dbias = dbias; a.rowsum_scratch = rowsum_scratch;
  return logsoftmax_bwd(a, static_cast<cudaStream_t>(stream));
}

int pk_dense_act_fwd(int C, int64_t n, int act, const float* PT, int64_t ldp, const float* scale, const float* shift,
                     const void* keepT, int64_t ldk, void* YT16, int64_t ld16t, void* Y16, int64_t ld16r, float* Y32,
                     int64_t ld32, void* stream) {
  PK_REQUIRE(PT && scale && shift, "pk_dense_act_fwd: null input");
  PK_REQUIRE(act >= PK_ACT_RELU && act <= PK_ACT_LINEAR, "pk_dense_act_fwd: bad activation %d", act);
  DenseFwdArgs a;
  a.C = C; a.n = n; a.act = act; a.PT = PT; a.ldp = ldp; a.scale = scale; a.shift = shift;
  a.keepT = static_cast<const __half*>(keepT); a.ldk = ldk;
  a.YT16 = static_cast<__half*>(YT16); a.ld16t = ld16t; a.Y16 = static_cast<__half*>(Y16); a.ld16r = ld16r;
  a.Y32 = Y32; a.ld32 = ld32;
  return dense_act_fwd(a, static_cast<cudaStream_t>(stream));
}
int pk_dense_act_bwd(int C, int64_t n, int act, const float* dYT, int64_t ldy, const void* YT16, int64_t ld16t,
                     const void* keepT, int64_t ldk, const float* gscale, void* GT16, int64_t ldg, void* stream) {
  PK_REQUIRE(dYT && YT16 && GT16, "pk_dense_act_bwd: null input");
  DenseBwdArgs a;
  a.C = C; a.n = n; a.act = act; a.dYT = dYT; a.ldy = ldy; a.YT16 = static_cast<const __half*>(YT16); a.ld16t = ld16t;
  a.keepT = static_cast<const __half*>(keepT); a.ldk = ldk; a.gscale = gscale;
  a.GT16 = static_cast<__half*>(GT16); a.ldg = ldg;
  return dense_act_bwd(a, static_cast<cudaStream_t>(stream));
}

int pk_rmsprop_step(float* p, const float* g, float* v, int64_t n, float lr, float alpha, float eps,
                    float gscale, void* stream) {
  PK_REQUIRE(n <= 0 || (p && g && v), "pk_rmsprop_step: null pointer");
  return rmsprop_step(p, g, v, n, lr, alpha, eps, gscale, static_cast<cudaStream_t>(stream));
}
int pk_chunk_prepare(const float* fea, int64_t ldf, const int64_t* lab, int64_t lab_min, int64_t n_in, int F, int left, int right,
                     double* stats, float* out, int64_t ldo, void* stream) {
  PK_REQUIRE(fea && stats && out, "pk_chunk_prepare: null pointer");
  return chunk_prepare(fea, ldf, reinterpret_cast<const long long*>(lab), lab_min, n_in, F, left, right, stats, out, ldo,
                       static_cast<cudaStream_t>(stream));
}
int pk_batch_assemble(const float* data_set, int64_t ldd, int D, const int64_t* desc, int batch_size, int max_len, float* inp,
                      void* stream) {
  PK_REQUIRE(data_set && desc && inp, "pk_batch_assemble: null pointer");
  return batch_assemble(data_set, ldd, D, reinterpret_cast<const long long*>(desc), batch_size, max_len, inp,
                        static_cast<cudaStream_t>(stream));
}
int pk_cm_decode(const void* col_headers, const void* data, float min_value, float range, int rows, int cols, float* out,
                 int64_t ldo, void* stream) {
  PK_REQUIRE(col_headers && data && out, "pk_cm_decode: null pointer");
  return cm_decode(static_cast<const uint16_t*>(col_headers), static_cast<const uint8_t*>(data), min_value, range, rows, cols, out,
                   ldo, static_cast<cudaStream_t>(stream));
}
int pk_sub_log_prior(float* logp, int64_t ld, int64_t n, int S, const float* log_prior, void* stream) {
  PK_REQUIRE(logp && log_prior, "pk_sub_log_prior: null pointer");
  return rows_sub_vec(logp, ld, n, S, log_prior, static_cast<cudaStream_t>(stream));
}
int pk_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, int64_t step, float gscale, void* stream) {
  PK_REQUIRE(p && g && m && v, "pk_adam_step: null pointer");
  return adam_step(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, gscale, static_cast<cudaStream_t>(stream));
}
int pk_sgd_step(float* p, const float* g, int64_t n, float lr, float gscale, void* stream) {
  PK_REQUIRE(n <= 0 || (p && g), "pk_sgd_step: null pointer");
  return sgd_step(p, g, n, lr, gscale, static_cast<cudaStream_t>(stream));
}

int pk_ln_cm_fwd(float* PT, int C, int64_t n, int64_t ld, const float* gamma, const float* beta, float eps, float* XH,
                 float* stats, void* stream) {
  PK_REQUIRE(PT && gamma && beta, "pk_ln_cm_fwd: null input");
  return ln_cm_fwd(PT, C, n, ld, gamma, beta, eps, XH, stats, static_cast<cudaStream_t>(stream));
}
int pk_ln_cm_bwd(void* dT16, int64_t ld16t, void* dR16, int64_t ld16r, const float* XH, int64_t ld, int C, int64_t n,
                 const float* gamma, const float* stats, float eps, const float* scale, float* dgamma, float* dbeta,
                 float* dbias, void* stream) {
  PK_REQUIRE(dT16 && XH && gamma && stats && scale && dgamma && dbeta, "pk_ln_cm_bwd: null input");
  return ln_cm_bwd(static_cast<__half*>(dT16), ld16t, static_cast<__half*>(dR16), ld16r, XH, ld, C, n, gamma, stats, eps, scale,
                   dgamma, dbeta, dbias, static_cast<cudaStream_t>(stream));
}
int pk_row_stats(const float* PT, int C, int64_t n, int64_t ld, double* stats, void* stream) {
  PK_REQUIRE(PT && stats, "pk_row_stats: null input");
  return row_stats(PT, C, n, ld, stats, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
