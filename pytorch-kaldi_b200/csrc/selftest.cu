// selftest.cu — standalone bring-up test for libpk_b200.so (GPU box only).
// Every kernel is called through the public C ABI and compared with a straightforward CPU
// loop written here (test infrastructure, not product code).  Also prints first timings.
//   usage: pk_selftest [quick|full]
#include "../../include/pk_b200.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>
#include <string>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)
#define PKC(x)                                                        \
  do {                                                                \
    int r_ = (x);                                                     \
    if (r_ != 0) {                                                    \
      printf("PK error %d: %s (%s:%d)\n", r_, pk_last_error(), __FILE__, __LINE__); \
      g_fail++;                                                       \
      return;                                                         \
    }                                                                 \
  } while (0)

extern "C" void pk_debug_set_clock_buffer(void*);  // bring-up hook, not in the public header
static int g_fail = 0;
static std::mt19937 rng(1234);

static float h2f(float x) { return __half2float(__float2half_rn(x)); }

template <typename T>
struct Dev {
  T* p = nullptr;
  size_t n = 0;
  explicit Dev(size_t n_) : n(n_) { CK(cudaMalloc(&p, n * sizeof(T))); CK(cudaMemset(p, 0, n * sizeof(T))); }
  ~Dev() { cudaFree(p); }
  void up(const std::vector<T>& h) { CK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)); }
  std::vector<T> down() const {
    std::vector<T> h(n);
    CK(cudaMemcpy(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost));
    return h;
  }
};

static std::vector<float> randn(size_t n, float sd = 1.f) {
  std::normal_distribution<float> d(0.f, sd);
  std::vector<float> v(n);
  for (auto& x : v) x = d(rng);
  return v;
}

static void report(const char* name, double maxerr, double tol, const std::string& extra = "") {
  const bool ok = (maxerr <= tol) && std::isfinite(maxerr);
  printf("[%s] %-44s max_err=%.3e tol=%.1e %s\n", ok ? " ok " : "FAIL", name, maxerr, tol, extra.c_str());
  if (!ok) g_fail++;
}

// ---------------------------------------------------------------------------------
static void test_gemm(int dtype, int M, int N, int K, int bias_mode, bool stats, int splitk, int accumulate) {
  const int esz = dtype == PK_F16 ? 2 : 4;
  const int align = 16 / esz;
  const long long lda = (K + align - 1) / align * align, ldb = lda, ldc = N + 3;
  std::vector<float> A = randn((size_t)M * lda), B = randn((size_t)N * ldb), bias = randn(std::max(M, N));
  // operands rounded so the tensor-core product is exact: fp16 values / tf32-representable values
  auto rnd = [&](float x) {
    if (dtype == PK_F16) return h2f(x);
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xFFFFE000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
  };
  for (auto& x : A) x = rnd(x);
  for (auto& x : B) x = rnd(x);
  std::vector<float> C0 = randn((size_t)M * ldc);
  Dev<float> dC((size_t)M * ldc), dbias(bias.size());
  Dev<double> dstats(2 * (size_t)M);
  dC.up(C0);
  dbias.up(bias);
  void *dA, *dB;
  CK(cudaMalloc(&dA, (size_t)M * lda * esz));
  CK(cudaMalloc(&dB, (size_t)N * ldb * esz));
  if (dtype == PK_F16) {
    std::vector<__half> hA(A.size()), hB(B.size());
    for (size_t i = 0; i < A.size(); ++i) hA[i] = __float2half_rn(A[i]);
    for (size_t i = 0; i < B.size(); ++i) hB[i] = __float2half_rn(B[i]);
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  } else {
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  }
  const float alpha = 0.5f;
  PKC(pk_gemm_tn(dtype, M, N, K, dA, lda, 0, 0, dB, ldb, 0, 0, dC.p, ldc, bias_mode ? dbias.p : nullptr, bias_mode,
                 stats ? dstats.p : nullptr, alpha, nullptr, accumulate, splitk, nullptr, nullptr));
  CK(cudaDeviceSynchronize());
  auto C = dC.down();
  auto st = dstats.down();
  double maxerr = 0, maxst = 0;
  std::vector<double> rs(M, 0.0), rs2(M, 0.0);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += (double)A[(size_t)m * lda + k] * B[(size_t)n * ldb + k];
      double ref = alpha * acc + (bias_mode == 1 ? bias[n] : bias_mode == 2 ? bias[m] : 0.0);
      rs[m] += ref;
      rs2[m] += ref * ref;
      if (accumulate) ref += C0[(size_t)m * ldc + n];
      maxerr = std::max(maxerr, std::fabs(ref - C[(size_t)m * ldc + n]) / (1.0 + std::fabs(ref)));
    }
  // padding columns of C must be untouched
  for (int m = 0; m < M; ++m)
    for (int n = N; n < ldc; ++n)
      if (C[(size_t)m * ldc + n] != C0[(size_t)m * ldc + n]) maxerr = 1e9;
  if (stats)
    for (int m = 0; m < M; ++m) {
      maxst = std::max(maxst, std::fabs(st[2 * m] - rs[m]) / (1.0 + std::fabs(rs[m])));
      maxst = std::max(maxst, std::fabs(st[2 * m + 1] - rs2[m]) / (1.0 + std::fabs(rs2[m])));
    }
  char name[128];
  snprintf(name, sizeof(name), "gemm %s M%d N%d K%d bias%d st%d sk%d acc%d", dtype == PK_F16 ? "f16" : "tf32", M, N,
           K, bias_mode, (int)stats, splitk, accumulate);
  report(name, std::max(maxerr, maxst), 2e-4);
  cudaFree(dA);
  cudaFree(dB);
}

// C = A[:, a0:a0+K] . B[:, b0:b0+K]^T with operand extents KX (the time-shifted dU product)
static void test_gemm_shift(int M, int N, int KX, int a0, int b0) {
  const int K = KX - std::max(a0, b0);
  const long long ld = (KX + 7) / 8 * 8 + 8;
  std::vector<float> A = randn((size_t)M * ld), B = randn((size_t)N * ld);
  for (auto& x : A) x = h2f(x);
  for (auto& x : B) x = h2f(x);
  std::vector<__half> hA(A.size()), hB(B.size());
  for (size_t i = 0; i < A.size(); ++i) hA[i] = __float2half_rn(A[i]);
  for (size_t i = 0; i < B.size(); ++i) hB[i] = __float2half_rn(B[i]);
  Dev<__half> dA(hA.size()), dB(hB.size());
  Dev<float> dC((size_t)M * N);
  dA.up(hA);
  dB.up(hB);
  PKC(pk_gemm_tn(PK_F16, M, N, K, dA.p, ld, a0, KX, dB.p, ld, b0, KX, dC.p, N, nullptr, 0, nullptr, 1.f, nullptr, 0, 3,
                 nullptr, nullptr));
  CK(cudaDeviceSynchronize());
  auto C = dC.down();
  double e = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += (double)A[(size_t)m * ld + a0 + k] * B[(size_t)n * ld + b0 + k];
      e = std::max(e, std::fabs(acc - C[(size_t)m * N + n]) / (1.0 + std::fabs(acc)));
    }
  char name[128];
  snprintf(name, sizeof(name), "gemm f16 shifted M%d N%d KX%d a0=%d b0=%d", M, N, KX, a0, b0);
  report(name, e, 2e-4);
}

// ---------------------------------------------------------------------------------
// liGRU layer CPU reference with the same operand quantisation as the kernels
struct LigruCase {
  int T, B, H, ndir, act;
  long long ld;  // T*B padded
  std::vector<float> PT, scale, shift, U, mask;
};
static float actf(int act, float x) {
  switch (act) {
    case PK_ACT_RELU: return x > 0 ? x : 0;
    case PK_ACT_TANH: return tanhf(x);
    case PK_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case PK_ACT_LEAKY_RELU: return x > 0 ? x : 0.2f * x;
    case PK_ACT_ELU: return x > 0 ? x : expm1f(x);
    default: return x;
  }
}
static float dactf(int act, float y) {
  switch (act) {
    case PK_ACT_RELU: return y > 0 ? 1.f : 0.f;
    case PK_ACT_TANH: return 1.f - y * y;
    case PK_ACT_SIGMOID: return y * (1.f - y);
    case PK_ACT_LEAKY_RELU: return y > 0 ? 1.f : 0.2f;
    case PK_ACT_ELU: return y > 0 ? 1.f : y + 1.f;
    default: return 1.f;
  }
}
static LigruCase make_case(int T, int B, int H, int ndir, int act) {
  LigruCase c{T, B, H, ndir, act};
  c.ld = ((long long)T * B + 7) / 8 * 8;
  c.PT = randn((size_t)2 * H * c.ld);
  c.scale = randn(2 * H, 0.3f);
  for (auto& x : c.scale) x += 1.f;
  c.shift = randn(2 * H, 0.2f);
  c.U = randn((size_t)2 * H * H, 1.0f / sqrtf((float)H));
  c.mask.resize((size_t)ndir * B * H);
  std::bernoulli_distribution bd(0.8);
  for (auto& x : c.mask) x = bd(rng) ? 1.f : 0.f;
  return c;
}
// outputs channel-major [ndir*H][ld]
static void ligru_fwd_cpu(const LigruCase& c, std::vector<float>& HT, std::vector<float>& ZT, std::vector<float>& HCT) {
  const int T = c.T, B = c.B, H = c.H;
  HT.assign((size_t)c.ndir * H * c.ld, 0.f);
  ZT = HT;
  HCT = HT;
  std::vector<float> Uq(c.U.size());
  for (size_t i = 0; i < Uq.size(); ++i) Uq[i] = h2f(c.U[i]);
  std::vector<float> h(H), hq(H), hn(H);
  for (int r = 0; r < c.ndir * B; ++r) {
    const int d = r >= B, b = r - d * B;
    std::fill(h.begin(), h.end(), 0.f);
    for (int k = 0; k < T; ++k) {
      const int t = d ? T - 1 - k : k;
      const long long col = (long long)t * B + b;
      for (int j = 0; j < H; ++j) hq[j] = h2f(h[j]);
      for (int u = 0; u < H; ++u) {
        double ah = 0, az = 0;
        for (int j = 0; j < H; ++j) {
          ah += (double)Uq[(size_t)u * H + j] * hq[j];
          az += (double)Uq[(size_t)(H + u) * H + j] * hq[j];
        }
        const float zt = 1.f / (1.f + expf(-(c.scale[H + u] * c.PT[(size_t)(H + u) * c.ld + col] + c.shift[H + u] + (float)az)));
        const float at = c.scale[u] * c.PT[(size_t)u * c.ld + col] + c.shift[u] + (float)ah;
        const float hc = actf(c.act, at) * c.mask[(size_t)r * H + u];
        hn[u] = zt * h[u] + (1.f - zt) * hc;
        const size_t idx = (size_t)(d * H + u) * c.ld + col;
        HT[idx] = hn[u];
        ZT[idx] = zt;
        HCT[idx] = hc;
      }
      h = hn;
    }
  }
}
static void ligru_bwd_cpu(const LigruCase& c, const std::vector<float>& dYT, const std::vector<float>& HT,
                          const std::vector<float>& ZT, const std::vector<float>& HCT, float s, std::vector<float>& GT) {
  const int T = c.T, B = c.B, H = c.H;
  GT.assign((size_t)c.ndir * 2 * H * c.ld, 0.f);
  std::vector<float> Uq(c.U.size());
  for (size_t i = 0; i < Uq.size(); ++i) Uq[i] = h2f(c.U[i]);
  std::vector<float> carry(H), da(H), dz(H), keep(H), gq(2 * H);
  for (int r = 0; r < c.ndir * B; ++r) {
    const int d = r >= B, b = r - d * B;
    std::fill(carry.begin(), carry.end(), 0.f);
    for (int k = T - 1; k >= 0; --k) {
      const int t = d ? T - 1 - k : k;
      const long long col = (long long)t * B + b;
      for (int u = 0; u < H; ++u) {
        const size_t idx = (size_t)(d * H + u) * c.ld + col;
        const float dh = dYT[idx] + carry[u];
        const float hp = k > 0 ? HT[idx + (d ? B : -B)] : 0.f;
        const float z = ZT[idx], hc = HCT[idx], m = c.mask[(size_t)r * H + u];
        const float y = m != 0.f ? hc / m : 0.f;
        da[u] = dh * (1.f - z) * m * dactf(c.act, y);
        dz[u] = dh * (hp - hc) * z * (1.f - z);
        keep[u] = dh * z;
        GT[(size_t)d * 2 * H * c.ld + (size_t)u * c.ld + col] = da[u];
        GT[(size_t)d * 2 * H * c.ld + (size_t)(H + u) * c.ld + col] = dz[u];
        gq[u] = h2f(da[u] * s);
        gq[H + u] = h2f(dz[u] * s);
      }
      for (int u = 0; u < H; ++u) {
        double acc = 0;
        for (int j = 0; j < H; ++j)
          acc += (double)Uq[(size_t)j * H + u] * gq[j] + (double)Uq[(size_t)(H + j) * H + u] * gq[H + j];
        carry[u] = keep[u] + (float)acc / s;
      }
    }
  }
}

static double maxrel(const std::vector<float>& a, const std::vector<float>& b, double floor_ = 1e-3) {
  double m = 0, scale = 0;
  for (size_t i = 0; i < a.size(); ++i) scale = std::max(scale, (double)std::fabs(b[i]));
  scale = std::max(scale, floor_);
  for (size_t i = 0; i < a.size(); ++i) {
    double e = std::fabs((double)a[i] - b[i]) / scale;
    if (!(e == e)) return 1e30;
    m = std::max(m, e);
  }
  return m;
}

static void test_ligru(int T, int B, int H, int ndir, int act) {
  LigruCase c = make_case(T, B, H, ndir, act);
  std::vector<float> HT, ZT, HCT;
  ligru_fwd_cpu(c, HT, ZT, HCT);
  const size_t nch = (size_t)ndir * H * c.ld;
  Dev<float> dPT(c.PT.size()), dsc(2 * H), dsh(2 * H), dU(c.U.size()), dmask(c.mask.size());
  dPT.up(c.PT); dsc.up(c.scale); dsh.up(c.shift); dU.up(c.U); dmask.up(c.mask);
  const long long ldy = (long long)ndir * H + 2;
  const long long ldy16 = ((long long)ndir * H + 7) / 8 * 8;
  const char* names[] = {"tc", "tc-3groups", "ws-allgather", "ws-ksplit"};
  const int vflags[] = {PK_REC_TC, PK_REC_TC | PK_REC_GROUPS(3), PK_REC_WS | PK_REC_BWD_ALLGATHER, PK_REC_WS | PK_REC_BWD_KSPLIT};
  const int npass = H > 560 ? 2 : 4;
  for (int pass = 0; pass < npass; ++pass) {
    const char* cl = names[pass];
    const int vflag = vflags[pass];
    Dev<float> dHT(nch), dZT(nch), dHCT(nch), dY((size_t)T * B * ldy);
    Dev<__half> dY16((size_t)T * B * ldy16), dHT16(nch), dHP16(nch);
    PKC(pk_rnn_layer_fwd(PK_CELL_LIGRU | vflag, T, B, H, ndir, act, dPT.p, c.ld, dsc.p, dsh.p, dU.p, dmask.p, 1.f, dY.p, ldy,
                         dY16.p, ldy16, dHT.p, dHT16.p, dHP16.p, dZT.p, dHCT.p, c.ld, nullptr));
    CK(cudaDeviceSynchronize());
    auto gHT = dHT.down(), gZT = dZT.down(), gHCT = dHCT.down(), gY = dY.down();
    auto gY16 = dY16.down();
    char name[128];
    snprintf(name, sizeof(name), "ligru_fwd cl%s T%d B%d H%d nd%d act%d", cl, T, B, H, ndir, act);
    double e = std::max(maxrel(gHT, HT), std::max(maxrel(gZT, ZT), maxrel(gHCT, HCT)));
    // row-major outputs must agree with the channel-major ones
    double ey = 0;
    for (int t = 0; t < T; ++t)
      for (int b = 0; b < B; ++b)
        for (int ch = 0; ch < ndir * H; ++ch) {
          const float ref = gHT[(size_t)ch * c.ld + (size_t)t * B + b];
          ey = std::max(ey, (double)std::fabs(gY[((size_t)t * B + b) * ldy + ch] - ref));
          ey = std::max(ey, (double)std::fabs(__half2float(gY16[((size_t)t * B + b) * ldy16 + ch]) - h2f(ref)));
        }
    {  // HP16 must be the fp16 state of the previous step at the same column
      auto gHP = dHP16.down();
      for (int r = 0; r < ndir * B; ++r) {
        const int d = r >= B, b = r - d * B;
        for (int k = 0; k < T; ++k) {
          const int t = d ? T - 1 - k : k;
          for (int uu = 0; uu < H; ++uu) {
            const size_t idx = (size_t)(d * H + uu) * c.ld + (size_t)t * B + b;
            const float ref = k > 0 ? h2f(gHT[idx + (d ? B : -B)]) : 0.f;
            ey = std::max(ey, (double)std::fabs(__half2float(gHP[idx]) - ref));
          }
        }
      }
    }
    report(name, std::max(e, ey), 2e-3);

    // backward
    std::vector<float> dYT = randn(nch, 1e-4f);
    const float s = 4096.f;
    std::vector<float> GT;
    ligru_bwd_cpu(c, dYT, gHT, gZT, gHCT, s, GT);
    Dev<float> ddYT(nch), dGT(GT.size()), dscale(1);
    Dev<__half> dGT16(GT.size());
    ddYT.up(dYT);
    dscale.up({s});
    PKC(pk_rnn_layer_bwd(PK_CELL_LIGRU | vflag, T, B, H, ndir, act, ddYT.p, dHT.p, dZT.p, dHCT.p, c.ld, dU.p, dmask.p, 1.f,
                         dscale.p, dGT.p, dGT16.p, nullptr));
    CK(cudaDeviceSynchronize());
    auto gGT = dGT.down();
    auto gGT16 = dGT16.down();
    double e16 = 0, gmax = 0;
    for (auto x : GT) gmax = std::max(gmax, (double)std::fabs(x));
    for (size_t i = 0; i < GT.size(); ++i)
      e16 = std::max(e16, std::fabs((double)__half2float(gGT16[i]) / s - GT[i]) / std::max(gmax, 1e-30));
    const bool ws = true;  // the kernels write GT16 only
    snprintf(name, sizeof(name), "ligru_bwd cl%s T%d B%d H%d nd%d act%d", cl, T, B, H, ndir, act);
    report(name, std::max(ws ? 0.0 : maxrel(gGT, GT, 1e-30), e16), 3e-3);
  }
}

// ---------------------------------------------------------------------------------
static void test_elementwise() {
  // transpose + conversions
  {
    const int R = 70, C = 45;
    const long long ldi = C + 1, ldo = R + 2, ldo16 = 72, ldi16 = 48;
    auto in = randn((size_t)R * ldi);
    Dev<float> din(in.size()), dout((size_t)C * ldo), dsc(1);
    Dev<__half> dT16((size_t)C * ldo16), d16((size_t)R * ldi16);
    din.up(in);
    dsc.up({2.f});
    Dev<float> dam(1), dsc2(2);
    PKC(pk_transpose_f32(din.p, ldi, R, C, dout.p, ldo, dT16.p, ldo16, d16.p, ldi16, dsc.p, dam.p, nullptr));
    PKC(pk_amax_finalize(dam.p, 8.f, dsc2.p, nullptr));
    CK(cudaDeviceSynchronize());
    auto o = dout.down();
    auto t16 = dT16.down();
    auto r16 = d16.down();
    double e = 0;
    for (int r = 0; r < R; ++r)
      for (int c = 0; c < C; ++c) {
        const float v = in[(size_t)r * ldi + c];
        e = std::max(e, (double)std::fabs(o[(size_t)c * ldo + r] - v));
        e = std::max(e, (double)std::fabs(__half2float(t16[(size_t)c * ldo16 + r]) - h2f(2.f * v)));
        e = std::max(e, (double)std::fabs(__half2float(r16[(size_t)r * ldi16 + c]) - h2f(2.f * v)));
      }
    {
      float am = 0;
      for (int r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c) am = std::max(am, std::fabs(in[(size_t)r * ldi + c]));
      const float s2 = dsc2.down()[0];
      if (!(am * s2 >= 128.f && am * s2 < 256.f) || dam.down()[0] != 0.f) e = 1;
    }
    report("transpose_f32 (+fp16 copies, fused amax)", e, 0);
  }
  // amax scale
  {
    auto x = randn(10000, 3e-5f);
    float amax = 0;
    for (auto v : x) amax = std::max(amax, std::fabs(v));
    Dev<float> dx(x.size()), dscr(1), dsc(2);
    dx.up(x);
    PKC(pk_amax_scale(dx.p, 100, 100, 100, 8.f, dscr.p, dsc.p, nullptr));
    CK(cudaDeviceSynchronize());
    const float s = dsc.down()[0];
    const bool ok = amax * s >= 128.f && amax * s < 256.f && std::exp2(std::round(std::log2(s))) == s &&
                    dsc.down()[1] == 1.f / s;
    report("amax_scale", ok ? 0 : 1, 0, "scale=" + std::to_string(s));
  }
  // log-softmax + NLL + err, and its backward (fused mode)
  {
    const int N = 77, S = 1936;
    const long long ld = S + 4, ld16 = (S + 7) / 8 * 8, ld16t = 80;
    auto x = randn((size_t)N * ld, 2.f);
    std::vector<long long> lab(N);
    for (auto& l : lab) l = rng() % S;
    Dev<float> dx(x.size()), dbias(S);
    Dev<long long> dlab(N);
    Dev<double> dacc(2);
    dx.up(x);
    dlab.up(lab);
    PKC(pk_logsoftmax_nll(N, S, dx.p, ld, (const int64_t*)dlab.p, dacc.p, nullptr));
    CK(cudaDeviceSynchronize());
    auto lp = dx.down();
    auto acc = dacc.down();
    double e = 0, loss = 0, err = 0;
    std::vector<double> ref((size_t)N * S);
    for (int n = 0; n < N; ++n) {
      double m = -1e30;
      int am = 0;
      for (int j = 0; j < S; ++j)
        if (x[(size_t)n * ld + j] > m) { m = x[(size_t)n * ld + j]; am = j; }
      double ssum = 0;
      for (int j = 0; j < S; ++j) ssum += std::exp(x[(size_t)n * ld + j] - m);
      const double lse = m + std::log(ssum);
      for (int j = 0; j < S; ++j) {
        ref[(size_t)n * S + j] = x[(size_t)n * ld + j] - lse;
        e = std::max(e, std::fabs(ref[(size_t)n * S + j] - lp[(size_t)n * ld + j]));
      }
      loss -= ref[(size_t)n * S + lab[n]];
      err += am != lab[n];
    }
    e = std::max(e, std::fabs(loss - acc[0]) / loss);
    e = std::max(e, std::fabs(err - acc[1]));
    report("logsoftmax_nll", e, 2e-5);
    Dev<__half> d16((size_t)N * ld16), dT16((size_t)S * ld16t);
    const float gcoef = 1.f / N, oscale = 1024.f;
    PKC(pk_logsoftmax_bwd(N, S, dx.p, ld, (const int64_t*)dlab.p, nullptr, 0, gcoef, oscale, nullptr, d16.p, ld16, dT16.p,
                          ld16t, dbias.p, nullptr, nullptr));
    CK(cudaDeviceSynchronize());
    auto g16 = d16.down();
    auto gT16 = dT16.down();
    auto gb = dbias.down();
    double eb = 0;
    std::vector<double> cs(S, 0.0);
    for (int n = 0; n < N; ++n)
      for (int j = 0; j < S; ++j) {
        const double d = (std::exp(ref[(size_t)n * S + j]) - (lab[n] == j)) * gcoef;
        cs[j] += d;
        eb = std::max(eb, std::fabs(__half2float(g16[(size_t)n * ld16 + j]) / oscale - d) / gcoef);
        eb = std::max(eb, std::fabs(__half2float(gT16[(size_t)j * ld16t + n]) / oscale - d) / gcoef);
      }
    for (int j = 0; j < S; ++j) eb = std::max(eb, std::fabs(gb[j] - cs[j]) / gcoef);
    report("logsoftmax_bwd (fused NLL)", eb, 1e-3);
  }
  // BatchNorm finalize + backward
  {
    const int C = 37, ndir = 2;
    const long long n = 203, ldt = 208, ldp = 208, ld16t = 208, ld16r = 40;
    auto PT = randn((size_t)C * ldp), GT = randn((size_t)ndir * C * ldt, 1e-3f), gamma = randn(C), beta = randn(C);
    std::vector<double> stats(2 * C, 0.0);
    for (int c = 0; c < C; ++c)
      for (long long i = 0; i < n; ++i) {
        stats[2 * c] += PT[(size_t)c * ldp + i];
        stats[2 * c + 1] += (double)PT[(size_t)c * ldp + i] * PT[(size_t)c * ldp + i];
      }
    Dev<double> dst(2 * C), dsums(2 * C);
    Dev<float> dg(C), db(C), drm(C), drv(C), dsc(C), dsh(C), dmean(C), drstd(C), dPT(PT.size()), dGT(GT.size()), ddg(C),
        ddb(C), dgs(1);
    Dev<long long> dnb(1);
    Dev<__half> dPT16((size_t)C * ld16t), dP16((size_t)n * ld16r);
    dst.up(stats); dg.up(gamma); db.up(beta); dPT.up(PT); dGT.up(GT);
    std::vector<float> ones(C, 1.f);
    drv.up(ones);
    const float gs = 512.f;
    dgs.up({gs});
    PKC(pk_bn_finalize(dst.p, C, n, 2 * n, dg.p, db.p, 1e-5f, 0.05f, 1, drm.p, drv.p, (int64_t*)dnb.p, dsc.p, dsh.p,
                       dmean.p, drstd.p, nullptr));
    PKC(pk_bn_bwd(C, ndir, n, dGT.p, nullptr, ldt, dPT.p, ldp, 1, 1, dmean.p, drstd.p, dg.p, dgs.p, ddg.p, ddb.p, dPT16.p, ld16t,
                  dP16.p, ld16r, dsums.p, nullptr));
    CK(cudaDeviceSynchronize());
    auto sc = dsc.down(), sh = dsh.down(), rm = drm.down(), rv = drv.down(), dgam = ddg.down(), dbet = ddb.down();
    auto pT16 = dPT16.down();
    auto p16 = dP16.down();
    double e = 0;
    for (int c = 0; c < C; ++c) {
      const double mean = stats[2 * c] / n, var = stats[2 * c + 1] / n - mean * mean, rstd = 1.0 / std::sqrt(var + 1e-5);
      e = std::max(e, std::fabs(sc[c] - gamma[c] * rstd) / (1 + std::fabs(gamma[c] * rstd)));
      e = std::max(e, std::fabs(sh[c] - (beta[c] - mean * gamma[c] * rstd)) / (1 + std::fabs(beta[c])));
      e = std::max(e, std::fabs(rm[c] - 0.05 * mean));
      e = std::max(e, std::fabs(rv[c] - (0.95 + 0.05 * var * (2.0 * n) / (2.0 * n - 1))));
      double s1 = 0, s2 = 0;
      for (long long i = 0; i < n; ++i) {
        const double g = (double)GT[(size_t)c * ldt + i] + GT[(size_t)(C + c) * ldt + i];
        s1 += g;
        s2 += g * (PT[(size_t)c * ldp + i] - mean) * rstd;
      }
      e = std::max(e, std::fabs(dbet[c] - s1) / 1e-3);
      e = std::max(e, std::fabs(dgam[c] - s2) / 1e-3);
      for (long long i = 0; i < n; ++i) {
        const double g = (double)GT[(size_t)c * ldt + i] + GT[(size_t)(C + c) * ldt + i];
        const double ph = (PT[(size_t)c * ldp + i] - mean) * rstd;
        const double dp = gamma[c] * rstd * (g - s1 / n - ph * s2 / n);
        e = std::max(e, std::fabs(__half2float(pT16[(size_t)c * ld16t + i]) / gs - dp) / 1e-2);
        e = std::max(e, std::fabs(__half2float(p16[(size_t)i * ld16r + c]) / gs - dp) / 1e-2);
      }
    }
    report("bn_finalize + bn_bwd", e, 2e-3, "nb=" + std::to_string(dnb.down()[0]));
  }
  // optimizers
  {
    const long long n = 1000;
    auto p = randn(n), g = randn(n), v = randn(n);
    for (auto& x : v) x = std::fabs(x);
    Dev<float> dp(n), dg(n), dv(n);
    dp.up(p); dg.up(g); dv.up(v);
    PKC(pk_rmsprop_step(dp.p, dg.p, dv.p, n, 4e-4f, 0.95f, 1e-8f, 0.5f, nullptr));
    CK(cudaDeviceSynchronize());
    auto gp = dp.down();
    double e = 0;
    for (long long i = 0; i < n; ++i) {
      const float gi = 0.5f * g[i];
      const float vi = 0.95f * v[i] + 0.05f * gi * gi;
      e = std::max(e, (double)std::fabs(gp[i] - (p[i] - 4e-4f * gi / (sqrtf(vi) + 1e-8f))));
    }
    report("rmsprop_step", e, 1e-6);
  }
}

// ---------------------------------------------------------------------------------
static float time_ms(int iters, const std::function<void()>& fn) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  fn();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) fn();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / iters;
}

static void bench_all() {
  printf("---- timings (config-2 shapes: T=500 B=32 H=550 bidir, S=1936) ----\n");
  {
    const int T = 500, B = 32, H = 550, ndir = 2;
    const long long ld = (long long)T * B;
    const size_t nch = (size_t)ndir * H * ld;
    Dev<float> dPT((size_t)2 * H * ld), dsc(2 * H), dsh(2 * H), dU((size_t)2 * H * H), dmask((size_t)ndir * B * H);
    dPT.up(randn((size_t)2 * H * ld));
    std::vector<float> ones(2 * H, 1.f);
    dsc.up(ones);
    dU.up(randn((size_t)2 * H * H, 1.f / sqrtf(550.f)));
    dmask.up(std::vector<float>((size_t)ndir * B * H, 1.f));
    Dev<float> dHT(nch), dZT(nch), dHCT(nch), dY((size_t)T * B * 1100), dGT(2 * nch), dgs(1), ddY(nch);
    Dev<__half> dY16((size_t)T * B * 1104), dHT16(nch), dHP16(nch), dGT16(2 * nch);
    dgs.up({1024.f});
    ddY.up(randn(nch, 1e-3f));
    struct V { const char* name; int flags; };
    const V vs[] = {{"default (auto)          ", 0},
                    {"tc tcgen05, 1 group     ", PK_REC_TC},
                    {"tc 2 groups             ", PK_REC_TC | PK_REC_GROUPS(2)},
                    {"tc 3 groups             ", PK_REC_TC | PK_REC_GROUPS(3)},
                    {"tc no proxy fence       ", PK_REC_TC | PK_REC_DBG_NOPROXYFENCE},
                    {"tc blocking wait        ", PK_REC_TC | PK_REC_DBG_BLOCKINGWAIT},
                    {"tc nostore              ", PK_REC_TC | PK_REC_DBG_NOSTORE},
                    {"tc noload/nostore       ", PK_REC_TC | PK_REC_DBG_NOSTORE | PK_REC_DBG_NOLOAD},
                    {"ws mma.sync all-gather  ", PK_REC_WS | PK_REC_BWD_ALLGATHER},
                    {"ws mma.sync bwd K-split ", PK_REC_WS | PK_REC_BWD_KSPLIT}};
    for (const V& v : vs) {
      const int flag = v.flags;
      int rc = 0;
      float ms = time_ms(3, [&] {
        rc |= pk_rnn_layer_fwd(PK_CELL_LIGRU | flag, T, B, H, ndir, PK_ACT_RELU, dPT.p, ld, dsc.p, dsh.p, dU.p, dmask.p, 1.f, dY.p, 1100,
                               dY16.p, 1104, dHT.p, dHT16.p, dHP16.p, dZT.p, dHCT.p, ld, nullptr);
      });
      printf("ligru_fwd %s: %.3f ms/layer  (%.3f us/step) rc=%d %s\n", v.name, ms, ms * 1000.f / T, rc,
             rc ? pk_last_error() : "");
      ms = time_ms(3, [&] {
        rc |= pk_rnn_layer_bwd(PK_CELL_LIGRU | flag, T, B, H, ndir, PK_ACT_RELU, ddY.p, dHT.p, dZT.p, dHCT.p, ld, dU.p, dmask.p, 1.f,
                               dgs.p, dGT.p, dGT16.p, nullptr);
      });
      printf("ligru_bwd %s: %.3f ms/layer  (%.3f us/step) rc=%d %s\n", v.name, ms, ms * 1000.f / T, rc,
             rc ? pk_last_error() : "");
    }
    {  // which of the forward kernel's output streams cost time? (null pointer = stream not written)
      struct A { const char* name; bool y, f32, f16; };
      const A as[] = {{"all outputs", true, true, true}, {"no Y16/Y32", false, true, true}, {"no fp32 HT/ZT/HCT", true, false, true},
                      {"no fp16 HT16/HP16", true, true, false}, {"Y16 only", true, false, false}, {"fp32 only", false, true, false}};
      for (const A& v : as) {
        int rc = 0;
        float ms = time_ms(3, [&] {
          rc |= pk_rnn_layer_fwd(PK_CELL_LIGRU, T, B, H, ndir, PK_ACT_RELU, dPT.p, ld, dsc.p, dsh.p, dU.p, dmask.p, 1.f, nullptr, 1100,
                                 v.y ? dY16.p : nullptr, 1104, v.f32 ? dHT.p : nullptr, v.f16 ? dHT16.p : nullptr,
                                 v.f16 ? dHP16.p : nullptr, v.f32 ? dZT.p : nullptr, v.f32 ? dHCT.p : nullptr, ld, nullptr);
        });
        printf("ligru_fwd ablation %-20s: %.3f us/step rc=%d\n", v.name, ms * 1000.f / T, rc);
      }
    }
    {  // per-phase cycle breakdown of the critical-path warp (CTA 0, warp 0)
      Dev<long long> dclk(8 + 8 * 16 + 16 * 64);
      pk_debug_set_clock_buffer(dclk.p);
      pk_rnn_layer_fwd(PK_CELL_LIGRU | PK_REC_TC, T, B, H, ndir, PK_ACT_RELU, dPT.p, ld, dsc.p, dsh.p, dU.p, dmask.p, 1.f, dY.p, 1100,
                       dY16.p, 1104, dHT.p, dHT16.p, dHP16.p, dZT.p, dHCT.p, ld, nullptr);
      CK(cudaDeviceSynchronize());
      auto c = dclk.down();
      printf("fwd phases (cycles/step): wait_acc %.0f | ld+xchg %.0f | gates+stage %.0f | push %.0f | rings(shadow) %.0f | - %.0f\n",
             c[0] / (double)T, c[1] / (double)T, c[2] / (double)T, c[3] / (double)T, c[4] / (double)T, c[5] / (double)T);
      {
        const long long base = c[8 + 4];
        printf("fwd trace (cycles rel. to acc wake of step 200): issuer[top first_ready last_ready commit] epi0[wake ld gates push rings] epi3[wake end]\n");
        for (int st = 0; st < 8; ++st) {
          const long long* t = &c[8 + st * 16];
          printf("  step %d: issuer %6lld %6lld %6lld %6lld | epi0 %6lld %6lld %6lld %6lld %6lld | epi3 %6lld %6lld\n", 200 + st, t[0] - base,
                 t[1] - base, t[2] - base, t[3] - base, t[4] - base, t[5] - base, t[6] - base, t[7] - base, t[8] - base, t[9] - base,
                 t[10] - base);
        }
      }
      {
        printf("fwd per-CTA global-time trace of cluster 0 (ns rel. to CTA 0 push start of step 202): [push_start push_end all_ready commit]\n");
        const long long base = c[136 + 2 * 8 + 0];
        for (int st = 2; st < 5; ++st)
          for (int cta = 0; cta < 9; ++cta) {
            const long long* t = &c[136 + cta * 64 + st * 8];
            printf("  step %d cta %d: %6lld %6lld %6lld %6lld\n", 200 + st, cta, t[0] - base, t[1] - base, t[2] - base, t[3] - base);
          }
      }
      pk_rnn_layer_bwd(PK_CELL_LIGRU | PK_REC_TC, T, B, H, ndir, PK_ACT_RELU, ddY.p, dHT.p, dZT.p, dHCT.p, ld, dU.p, dmask.p, 1.f, dgs.p,
                       dGT.p, dGT16.p, nullptr);
      CK(cudaDeviceSynchronize());
      c = dclk.down();
      printf("bwd phases (cycles/step): pointwise+stage %.0f | push %.0f | rings(shadow) %.0f | wait_acc %.0f | ld+xchg+carry %.0f | - %.0f\n",
             c[0] / (double)T, c[1] / (double)T, c[2] / (double)T, c[3] / (double)T, c[4] / (double)T, c[5] / (double)T);
      pk_rnn_layer_fwd(PK_CELL_LIGRU | PK_REC_WS, T, B, H, ndir, PK_ACT_RELU, dPT.p, ld, dsc.p, dsh.p, dU.p, dmask.p, 1.f, dY.p, 1100,
                       dY16.p, 1104, dHT.p, dHT16.p, dHP16.p, dZT.p, dHCT.p, ld, nullptr);
      CK(cudaDeviceSynchronize());
      c = dclk.down();
      printf("fwd ws phases (cycles/step): wait %.0f | ldmatrix+HMMA %.0f | gates %.0f | stage+push %.0f | rings(shadow) %.0f\n",
             c[0] / (double)T, c[1] / (double)T, c[2] / (double)T, c[3] / (double)T, c[4] / (double)T);
      for (int v = 0; v < 2; ++v) {
        pk_rnn_layer_bwd(PK_CELL_LIGRU | PK_REC_WS | (v ? PK_REC_BWD_KSPLIT : PK_REC_BWD_ALLGATHER), T, B, H, ndir, PK_ACT_RELU, ddY.p,
                         dHT.p, dZT.p, dHCT.p, ld, dU.p, dmask.p, 1.f, dgs.p, dGT.p, dGT16.p, nullptr);
        CK(cudaDeviceSynchronize());
        c = dclk.down();
        printf(v ? "bwd ws K-split phases (cycles/step): pointwise+local stage+bar %.0f | HMMA+stage+push %.0f | rings(shadow) %.0f | wait %.0f | reduce %.0f\n"
                 : "bwd ws all-gather phases (cycles/step): pointwise %.0f | stage+push %.0f | rings(shadow) %.0f | wait %.0f | HMMA+pair exchange %.0f\n",
               c[0] / (double)T, c[1] / (double)T, c[2] / (double)T, c[3] / (double)T, c[4] / (double)T);
      }
      pk_debug_set_clock_buffer(nullptr);
    }
  }
  {  // config-4 layer shape: H = 1024 (15 chunks of weights in tensor memory + 1 in shared memory)
    const int T = 500, B = 32, H = 1024, ndir = 2;
    const long long ld = (long long)T * B;
    const size_t nch = (size_t)ndir * H * ld;
    Dev<float> dPT((size_t)2 * H * ld), dsc(2 * H), dsh(2 * H), dU((size_t)2 * H * H), dmask((size_t)ndir * B * H);
    dPT.up(randn((size_t)2 * H * ld));
    dsc.up(std::vector<float>(2 * H, 1.f));
    dU.up(randn((size_t)2 * H * H, 1.f / 32.f));
    dmask.up(std::vector<float>((size_t)ndir * B * H, 1.f));
    Dev<float> dHT(nch), dZT(nch), dHCT(nch), dgs(1), ddY(nch);
    Dev<__half> dY16((size_t)T * B * 2048), dHT16(nch), dHP16(nch), dGT16(2 * nch);
    dgs.up({1024.f});
    ddY.up(randn(nch, 1e-3f));
    int rc = 0;
    float ms = time_ms(3, [&] {
      rc |= pk_rnn_layer_fwd(PK_CELL_LIGRU, T, B, H, ndir, PK_ACT_RELU, dPT.p, ld, dsc.p, dsh.p, dU.p, dmask.p, 1.f, nullptr, 2048,
                             dY16.p, 2048, dHT.p, dHT16.p, dHP16.p, dZT.p, dHCT.p, ld, nullptr);
    });
    printf("ligru_fwd H=1024 tcgen05: %.3f ms/layer  (%.3f us/step) rc=%d %s\n", ms, ms * 1000.f / T, rc, rc ? pk_last_error() : "");
    ms = time_ms(3, [&] {
      rc |= pk_rnn_layer_bwd(PK_CELL_LIGRU, T, B, H, ndir, PK_ACT_RELU, ddY.p, dHT.p, dZT.p, dHCT.p, ld, dU.p, dmask.p, 1.f, dgs.p,
                             nullptr, dGT16.p, nullptr);
    });
    printf("ligru_bwd H=1024 tcgen05: %.3f ms/layer  (%.3f us/step) rc=%d %s\n", ms, ms * 1000.f / T, rc, rc ? pk_last_error() : "");
  }
  struct G { const char* name; int M, N, K, sk; };
  const G gs[] = {{"proj  PT=W.X^T  ", 1100, 16000, 1100, 1}, {"head  logits    ", 16000, 1936, 1100, 1},
                  {"dW    dPT.XT^T  ", 1100, 1100, 16000, 8}, {"dU    GT.HT^T   ", 1100, 550, 15968, 16},
                  {"dXT   WT.dP^T   ", 1100, 16000, 1100, 1}, {"square 4096     ", 4096, 4096, 4096, 1}};
  for (const G& g : gs) {
    for (int dtype : {PK_F16, PK_TF32}) {
      const int esz = dtype == PK_F16 ? 2 : 4;
      const long long lda = (g.K + 7) / 8 * 8;
      void *dA, *dB;
      CK(cudaMalloc(&dA, (size_t)g.M * lda * esz));
      CK(cudaMalloc(&dB, (size_t)g.N * lda * esz));
      CK(cudaMemset(dA, 0, (size_t)g.M * lda * esz));
      CK(cudaMemset(dB, 0, (size_t)g.N * lda * esz));
      Dev<float> dC((size_t)g.M * g.N);
      int rc = 0;
      float ms = time_ms(5, [&] {
        rc |= pk_gemm_tn(dtype, g.M, g.N, g.K, dA, lda, 0, 0, dB, lda, 0, 0, dC.p, g.N, nullptr, 0, nullptr, 1.f, nullptr, 0, g.sk,
                         nullptr, nullptr);
      });
      printf("gemm %s %-4s M=%5d N=%5d K=%5d sk=%2d : %.3f ms  %.1f TFLOP/s rc=%d\n", g.name,
             dtype == PK_F16 ? "f16" : "tf32", g.M, g.N, g.K, g.sk, ms, 2.0 * g.M * g.N * g.K / ms * 1e-9, rc);
      cudaFree(dA);
      cudaFree(dB);
    }
  }
}

int main(int argc, char** argv) {
  const bool full = argc > 1 && std::string(argv[1]) == "full";
  if (argc > 1 && std::string(argv[1]) == "bench") {  // timings only (used under ncu)
    bench_all();
    return 0;
  }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs, pk_version=%d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount,
         pk_version());
  // GEMM: start tiny, then tails, epilogue variants, split-K
  test_gemm(PK_F16, 128, 128, 64, 0, false, 1, 0);
  test_gemm(PK_F16, 128, 128, 256, 0, false, 1, 0);
  test_gemm(PK_TF32, 128, 128, 32, 0, false, 1, 0);
  test_gemm(PK_TF32, 128, 128, 128, 0, false, 1, 0);
  test_gemm(PK_F16, 200, 136, 520, 1, false, 1, 0);
  test_gemm(PK_F16, 77, 300, 1100, 2, true, 1, 0);
  test_gemm(PK_TF32, 130, 257, 100, 1, true, 1, 0);
  test_gemm(PK_F16, 256, 130, 2000, 0, false, 4, 0);
  test_gemm(PK_F16, 256, 130, 2000, 2, false, 3, 1);
  test_gemm(PK_TF32, 100, 60, 900, 0, false, 1, 1);
  test_gemm_shift(150, 70, 1000, 8, 0);
  test_gemm_shift(150, 70, 1000, 0, 40);
  test_gemm_shift(64, 200, 333, 32, 0);
  test_elementwise();
  test_ligru(3, 2, 20, 1, PK_ACT_RELU);
  test_ligru(7, 5, 70, 2, PK_ACT_TANH);
  test_ligru(12, 8, 550, 2, PK_ACT_RELU);
  if (full) {
    test_ligru(40, 32, 550, 2, PK_ACT_RELU);
    test_ligru(9, 3, 512, 2, PK_ACT_LEAKY_RELU);
    test_ligru(9, 11, 300, 1, PK_ACT_SIGMOID);
    test_ligru(6, 8, 900, 2, PK_ACT_RELU);
    test_ligru(5, 8, 1024, 2, PK_ACT_RELU);
    test_ligru(4, 3, 1000, 1, PK_ACT_TANH);
    test_ligru(5, 4, 64, 1, PK_ACT_TANH);
  }
  bench_all();
  printf("SELFTEST %s (%d failures)\n", g_fail ? "FAILED" : "PASSED", g_fail);
  return g_fail ? 1 : 0;
}
