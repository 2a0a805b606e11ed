// pk_conv.cu — streaming companions of the conv front-ends (reference CNN neural_networks.py:1464-1556,
// SincNet :1559-1665, SincConv :1668-1813).
//
// Formulation.  Activations are POSITION-MAJOR fp16 `A16[n][l][c]` (channels contiguous, padded to a multiple
// of 8).  A stride-1 valid convolution is then ONE tcgen05 GEMM without im2col:
//     O[(n,l)][co] = sum_{kk,ci} A16[(n,l+kk)][ci] * W16[co][kk*Cp + ci]
// because the im2col row of position (n,l) is the contiguous range A16[(n*L+l)*Cp ... +k*Cp): the GEMM's A
// operand is the activation buffer itself viewed with row pitch Cp and K = k*Cp (overlapping rows; TMA only
// needs a 16-byte multiple pitch).  Rows with l > L-k run into the next frame and are ignored downstream.
// The first layer has one input channel (pitch 2 bytes), so it alone gets an explicit fp16 im2col.
// Backward: dX uses the same trick on the (front-padded) output gradient with flipped weights; dW contracts
// over positions and needs both operands channel-major, i.e. an explicit transposed im2col.
//
// Kernels here: input LayerNorm (ln0) fwd + its parameter gradients, sinc filter synthesis fwd/bwd, weight
// packing, im2col for layer 0 / transposed im2col, and the fused per-layer epilogue
//     drop(act(LN_L(max_pool1d(conv + b))))   fwd / bwd      (:1547-1553, :1652-1661)
// where LN_L is the reference's LayerNorm over the LAST (length) axis with a [C, Lp] affine (:23-33, :1505).
#include "pk_common.cuh"
#include "pk_kernels.h"

#include <algorithm>

namespace pk {

namespace {

__device__ __forceinline__ float block_sum(float v, float* red) {
  // blockDim.x threads, red has >= 32 floats; returns the sum to every thread
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < nw) ? red[l] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

// ---- input LayerNorm over the sample axis (ln0, :1541-1542 / :1644-1645) ----
__global__ void rowln_fwd_kernel(const float* __restrict__ x, long long ldx, int L, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, float* __restrict__ y, float* __restrict__ stats) {
  __shared__ float red[32];
  const int n = blockIdx.x;
  const float* xr = x + static_cast<long long>(n) * ldx;
  float s = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) s += xr[i];
  const float mean = block_sum(s, red) / L;
  float q = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) { const float d = xr[i] - mean; q += d * d; }
  const float stdv = sqrtf(block_sum(q, red) / (L - 1));   // unbiased, like torch.std (:31)
  const float inv = 1.f / (stdv + eps);
  for (int i = threadIdx.x; i < L; i += blockDim.x)
    y[static_cast<long long>(n) * L + i] = fmaf(gamma[i] * inv, xr[i] - mean, beta[i]);
  if (threadIdx.x == 0) { stats[2 * n] = mean; stats[2 * n + 1] = inv; }
}

// d(ln0 gamma/beta) from G[(n,l)][k-1-kk] = sum_co dO[(n,l)][co] F[co][kk] (columns in the flipped order of the
// packed Wflip operand):  dx[n][p] = sum_kk G[(n,p-kk)][k-1-kk]
__global__ void ln0_bwd_kernel(const float* __restrict__ G, long long ldg, int N, int L, int Lout, int k,
                               const float* __restrict__ x, long long ldx, const float* __restrict__ stats,
                               float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= L) return;
  float ag = 0.f, ab = 0.f;
  for (int n = blockIdx.y; n < N; n += gridDim.y) {
    float dx = 0.f;
    const int kk_lo = max(0, p - (Lout - 1)), kk_hi = min(k - 1, p);
    for (int kk = kk_lo; kk <= kk_hi; ++kk)
      dx += G[(static_cast<long long>(n) * L + (p - kk)) * ldg + (k - 1 - kk)];
    const float xh = (x[static_cast<long long>(n) * ldx + p] - stats[2 * n]) * stats[2 * n + 1];
    ag = fmaf(dx, xh, ag);
    ab += dx;
  }
  atomicAdd(dgamma + p, ag);
  atomicAdd(dbeta + p, ab);
}

// ---- SincConv filter synthesis (:1777-1803) ----
// one block per filter; thread j handles tap j
__global__ void sinc_fwd_kernel(const float* __restrict__ low_hz_, const float* __restrict__ band_hz_, int k, float sr,
                                float min_low, float min_band, float* __restrict__ filt) {
  extern __shared__ float sh[];  // [k] band-pass
  __shared__ float red[32];
  const int c = blockIdx.x;
  const float low = min_low / sr + fabsf(low_hz_[c]);
  const float high = low + min_band / sr + fabsf(band_hz_[c]);
  const int half = (k - 1) / 2;
  const float two_pi = 6.283185307179586f;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    float bp;
    if (j == half) {
      bp = 2.f * high - 2.f * low;
    } else {
      const int jl = j < half ? j : k - 1 - j;       // mirrored evaluation (:1762-1770)
      const float t = (static_cast<float>(jl) - static_cast<float>(half)) / sr;   // n_ (:1759-1760)
      const float a = two_pi * t * sr;
      const float xl = low * a, xh = high * a;
      bp = 2.f * high * (sinf(xh) / xh) - 2.f * low * (sinf(xl) / xl);
    }
    sh[j] = bp;
    mx = fmaxf(mx, bp);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  float m = red[0];
  for (int w = 1; w < (blockDim.x + 31) / 32; ++w) m = fmaxf(m, red[w]);
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    // Hamming window on linspace(0, k, k) (:1755-1756)
    const float nl = static_cast<float>(j) * static_cast<float>(k) / static_cast<float>(k - 1);
    const float win = 0.54f - 0.46f * cosf(two_pi * nl / static_cast<float>(k));
    filt[static_cast<long long>(c) * k + j] = sh[j] / m * win;
  }
}

__global__ void sinc_bwd_kernel(const float* __restrict__ low_hz_, const float* __restrict__ band_hz_, int k, float sr,
                                float min_low, float min_band, const float* __restrict__ dfilt, float* __restrict__ dlow_out,
                                float* __restrict__ dband_out) {
  // single thread per filter (k <= a few hundred taps, C <= a few hundred filters): tiny
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const float low = min_low / sr + fabsf(low_hz_[c]);
  const float high = low + min_band / sr + fabsf(band_hz_[c]);
  const int half = (k - 1) / 2;
  const float two_pi = 6.283185307179586f;
  // pass 1: band-pass, its max and arg-max, sum g*bp
  float m = -INFINITY;
  int jstar = 0;
  float gbp = 0.f;
  for (int j = 0; j < k; ++j) {
    float bp;
    if (j == half) {
      bp = 2.f * high - 2.f * low;
    } else {
      const int jl = j < half ? j : k - 1 - j;
      const float a = two_pi * (static_cast<float>(jl) - static_cast<float>(half));
      bp = 2.f * high * (sinf(high * a) / (high * a)) - 2.f * low * (sinf(low * a) / (low * a));
    }
    const float nl = static_cast<float>(j) * static_cast<float>(k) / static_cast<float>(k - 1);
    const float win = 0.54f - 0.46f * cosf(two_pi * nl / static_cast<float>(k));
    const float g = dfilt[static_cast<long long>(c) * k + j] * win;
    gbp = fmaf(g, bp, gbp);
    if (bp > m) { m = bp; jstar = j; }
  }
  const float corr = gbp / (m * m);
  float dhigh = 0.f, dlow = 0.f;
  for (int j = 0; j < k; ++j) {
    const float nl = static_cast<float>(j) * static_cast<float>(k) / static_cast<float>(k - 1);
    const float win = 0.54f - 0.46f * cosf(two_pi * nl / static_cast<float>(k));
    float dbp = dfilt[static_cast<long long>(c) * k + j] * win / m;
    if (j == jstar) dbp -= corr;
    float dh, dl;
    if (j == half) {
      dh = 2.f; dl = 2.f;
    } else {
      const int jl = j < half ? j : k - 1 - j;
      const float a = two_pi * (static_cast<float>(jl) - static_cast<float>(half));
      dh = 2.f * cosf(high * a); dl = 2.f * cosf(low * a);
    }
    dhigh = fmaf(dbp, dh, dhigh);
    dlow = fmaf(-dbp, dl, dlow);
  }
  dlow += dhigh;  // high = low + const + |band|
  const float sl = low_hz_[c] > 0.f ? 1.f : (low_hz_[c] < 0.f ? -1.f : 0.f);
  const float sb = band_hz_[c] > 0.f ? 1.f : (band_hz_[c] < 0.f ? -1.f : 0.f);
  dlow_out[c] = sl * dlow;
  dband_out[c] = sb * dhigh;
}

// ---- weight packing: W16[co][kk*Cip + ci] = w[co][ci][kk];  Wflip16[ci][j*Cop + co] = w[co][ci][k-1-j] ----
__global__ void conv_pack_kernel(const float* __restrict__ w, int Co, int Ci, int k, __half* __restrict__ W16, int Cip,
                                 long long ldw, __half* __restrict__ Wf16, int Cop, long long ldf) {
  const long long total = static_cast<long long>(Co) * Ci * k;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kk = static_cast<int>(e % k);
    const int ci = static_cast<int>((e / k) % Ci);
    const int co = static_cast<int>(e / (static_cast<long long>(k) * Ci));
    const __half v = f16_sat(w[e]);
    if (W16) W16[co * ldw + static_cast<long long>(kk) * Cip + ci] = v;
    if (Wf16) Wf16[ci * ldf + static_cast<long long>(k - 1 - kk) * Cop + co] = v;
  }
}

// ---- layer-0 im2col: Xcol[(n,l)][kk] = x[n][l+kk] (l < Lout, else 0);  XcolT[kk][(n,l)] likewise ----
__global__ void im2col0_kernel(const float* __restrict__ x, long long ldx, int N, int L, int k, int Lout,
                               __half* __restrict__ Xcol, int Kp, __half* __restrict__ XcolT, long long ldp) {
  // block: 128 positions of one frame; threads sweep taps
  const int n = blockIdx.y;
  const int l0 = blockIdx.x * 128;
  extern __shared__ float xs[];  // [128 + k]
  const float* xr = x + static_cast<long long>(n) * ldx;
  for (int i = threadIdx.x; i < 128 + k; i += blockDim.x) xs[i] = (l0 + i < L) ? xr[l0 + i] : 0.f;
  __syncthreads();
  if (Xcol) {
    for (int e = threadIdx.x; e < 128 * Kp; e += blockDim.x) {
      const int li = e / Kp, kk = e % Kp;
      const int l = l0 + li;
      if (l >= L) continue;
      const float v = (l < Lout && kk < k) ? xs[li + kk] : 0.f;
      Xcol[(static_cast<long long>(n) * L + l) * Kp + kk] = f16_sat(v);
    }
  }
  if (XcolT) {
    for (int e = threadIdx.x; e < 128 * k; e += blockDim.x) {
      const int kk = e / 128, li = e % 128;
      const int l = l0 + li;
      if (l >= L) continue;
      const float v = (l < Lout) ? xs[li + kk] : 0.f;
      XcolT[static_cast<long long>(kk) * ldp + static_cast<long long>(n) * L + l] = f16_sat(v);
    }
  }
}

// ---- transposed im2col for dW: XT[kk*Ci + ci][pos] = A16[pos + kk][ci] ----
__global__ void im2colT_kernel(const __half* __restrict__ A16, long long rows, int Cp, int Ci, int k,
                               __half* __restrict__ XT, long long ldp) {
  __shared__ __half tile[64][66];
  const long long p0 = static_cast<long long>(blockIdx.x) * 64;
  const int c0 = blockIdx.y * 64;
  const int kk = blockIdx.z;
  for (int e = threadIdx.x; e < 64 * 64; e += blockDim.x) {
    const int pi = e / 64, ci = e % 64;
    const long long pos = p0 + pi + kk;
    tile[pi][ci] = (p0 + pi < rows && c0 + ci < Ci) ? A16[pos * Cp + c0 + ci] : __float2half(0.f);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 64 * 64; e += blockDim.x) {
    const int ci = e / 64, pi = e % 64;
    if (c0 + ci < Ci && p0 + pi < rows)
      XT[(static_cast<long long>(kk) * Ci + c0 + ci) * ldp + p0 + pi] = tile[pi][ci];
  }
}

// ---- fused epilogue: drop(act(LN_L(max_pool1d(O)))) ----
// block (32 channels, NP length partitions); grid (ceil(C/32), N)
constexpr int kNP = 8;

struct PostFwd {
  const float* O; long long ldo;
  int N, L, Lout, p, Lp, C, act;
  const float* gamma; const float* beta; float eps;
  const __half* keep;   // [N][Lp][C] (0 or 1/(1-p)) or null
  float* P; uint8_t* arg; float* stats;
  __half* A16n; int Cpn;
  float* Y32;
};

__global__ void __launch_bounds__(32 * kNP) conv_post_fwd_kernel(const PostFwd a) {
  __shared__ float red[kNP][32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * 32 + tx;
  const int n = blockIdx.y;
  const bool cok = c < a.C;
  const int chunk = (a.Lp + kNP - 1) / kNP;
  const int l_lo = ty * chunk, l_hi = min(a.Lp, l_lo + chunk);
  const long long obase = static_cast<long long>(n) * a.L;
  const long long pbase = static_cast<long long>(n) * a.Lp;
  // pass 1: max-pool (first maximum wins, like F.max_pool1d), sum
  float s = 0.f;
  if (cok) {
    for (int lp = l_lo; lp < l_hi; ++lp) {
      float best = -INFINITY;
      int bi = 0;
      for (int j = 0; j < a.p; ++j) {
        const float v = a.O[(obase + lp * a.p + j) * a.ldo + c];
        if (v > best) { best = v; bi = j; }
      }
      a.P[(pbase + lp) * a.C + c] = best;
      a.arg[(pbase + lp) * a.C + c] = static_cast<uint8_t>(bi);
      s += best;
    }
  }
  float mean = 0.f, inv = 1.f;
  const bool ln = a.gamma != nullptr;
  if (ln) {
    red[ty][tx] = s;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kNP; ++i) t += red[i][tx];
    mean = t / a.Lp;
    __syncthreads();
    float q = 0.f;
    if (cok)
      for (int lp = l_lo; lp < l_hi; ++lp) { const float d = a.P[(pbase + lp) * a.C + c] - mean; q += d * d; }
    red[ty][tx] = q;
    __syncthreads();
    t = 0.f;
#pragma unroll
    for (int i = 0; i < kNP; ++i) t += red[i][tx];
    const float stdv = sqrtf(t / (a.Lp - 1));
    inv = 1.f / (stdv + a.eps);
    if (cok && ty == 0) {
      a.stats[(static_cast<long long>(n) * a.C + c) * 2] = mean;
      a.stats[(static_cast<long long>(n) * a.C + c) * 2 + 1] = inv;
    }
  }
  // pass 2: normalise, activate, drop, emit
  for (int lp = l_lo; lp < l_hi; ++lp) {
    float out = 0.f;
    if (cok) {
      float v = a.P[(pbase + lp) * a.C + c];
      if (ln) v = fmaf(a.gamma[static_cast<long long>(c) * a.Lp + lp] * inv, v - mean, a.beta[static_cast<long long>(c) * a.Lp + lp]);
      out = act_fwd(a.act, v);
      if (a.keep) out *= __half2float(a.keep[(pbase + lp) * a.C + c]);
      if (a.Y32) a.Y32[(static_cast<long long>(n) * a.C + c) * a.Lp + lp] = out;
    }
    if (a.A16n && c < a.Cpn) a.A16n[(pbase + lp) * a.Cpn + c] = f16_sat(out);
  }
}

struct PostBwd {
  const float* dY; long long sn, sl, sc;   // element strides of dY[n][l'][c]
  int N, L, Lout, p, Lp, C, act;
  const float* gamma; const float* beta; float eps;
  const __half* keep;
  const float* P; const uint8_t* arg; const float* stats;
  float* dgamma; float* dbeta; float* dbias;
  float* dO;            // [N*L][C] fp32
  unsigned int* amax_bits;
};

__global__ void __launch_bounds__(32 * kNP) conv_post_bwd_kernel(const PostBwd a) {
  __shared__ float red[2][kNP][32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * 32 + tx;
  const int n = blockIdx.y;
  const bool cok = c < a.C;
  const int chunk = (a.Lp + kNP - 1) / kNP;
  const int l_lo = ty * chunk, l_hi = min(a.Lp, l_lo + chunk);
  const long long pbase = static_cast<long long>(n) * a.Lp;
  const bool ln = a.gamma != nullptr;
  float mean = 0.f, inv = 1.f;
  if (ln && cok) {
    mean = a.stats[(static_cast<long long>(n) * a.C + c) * 2];
    inv = a.stats[(static_cast<long long>(n) * a.C + c) * 2 + 1];
  }
  auto grad_pre = [&](int lp, float& xc) -> float {   // gradient w.r.t. the (normalised) pre-activation
    const float pv = a.P[(pbase + lp) * a.C + c];
    xc = pv - mean;
    float y = pv;
    if (ln) y = fmaf(a.gamma[static_cast<long long>(c) * a.Lp + lp] * inv, xc, a.beta[static_cast<long long>(c) * a.Lp + lp]);
    const float out = act_fwd(a.act, y);
    float g = a.dY[n * a.sn + lp * a.sl + c * a.sc];
    if (a.keep) g *= __half2float(a.keep[(pbase + lp) * a.C + c]);
    return g * act_bwd_from_out(a.act, out);
  };
  float s1 = 0.f, s2 = 0.f;
  if (ln) {
    if (cok) {
      for (int lp = l_lo; lp < l_hi; ++lp) {
        float xc;
        const float g = grad_pre(lp, xc);
        atomicAdd(a.dgamma + static_cast<long long>(c) * a.Lp + lp, g * xc * inv);
        atomicAdd(a.dbeta + static_cast<long long>(c) * a.Lp + lp, g);
        const float gy = g * a.gamma[static_cast<long long>(c) * a.Lp + lp];
        s1 += gy;
        s2 = fmaf(gy, xc, s2);
      }
    }
    red[0][ty][tx] = s1; red[1][ty][tx] = s2;
    __syncthreads();
    s1 = 0.f; s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kNP; ++i) { s1 += red[0][i][tx]; s2 += red[1][i][tx]; }
  }
  // LN backward with std = 1/inv - eps (unbiased): dx = inv (gy - S1/Lp) - inv^2 S2 xc / ((Lp-1) std)
  const float stdv = ln ? fmaxf(1.f / inv - a.eps, 1e-30f) : 1.f;
  const float k1 = ln ? s1 / a.Lp : 0.f;
  const float k2 = ln ? inv * inv * s2 / ((a.Lp - 1) * stdv) : 0.f;
  float bsum = 0.f, amax = 0.f;
  if (cok) {
    for (int lp = l_lo; lp < l_hi; ++lp) {
      float xc;
      float d = grad_pre(lp, xc);
      if (ln) d = inv * (d * a.gamma[static_cast<long long>(c) * a.Lp + lp] - k1) - k2 * xc;
      const int bi = a.arg[(pbase + lp) * a.C + c];
      for (int j = 0; j < a.p; ++j)
        a.dO[(static_cast<long long>(n) * a.L + lp * a.p + j) * a.C + c] = (j == bi) ? d : 0.f;
      bsum += d;
      amax = fmaxf(amax, fabsf(d));
    }
    // positions past the pooled region (and the invalid conv tail) carry no gradient
    for (int l = a.Lp * a.p + ty; l < a.L; l += kNP) a.dO[(static_cast<long long>(n) * a.L + l) * a.C + c] = 0.f;
    if (a.dbias) atomicAdd(a.dbias + c, bsum);
  }
  if (a.amax_bits) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (tx == 0 && amax > 0.f) atomicMax(a.amax_bits, __float_as_uint(amax));
  }
}

}  // namespace

int rowln_fwd(const float* x, long long ldx, int N, int L, const float* gamma, const float* beta, float eps, float* y,
              float* stats, cudaStream_t stream) {
  PK_REQUIRE(N > 0 && L > 1, "rowln_fwd: bad shape");
  rowln_fwd_kernel<<<N, 256, 0, stream>>>(x, ldx, L, gamma, beta, eps, y, stats);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int conv_ln0_bwd(const float* G, long long ldg, int N, int L, int Lout, int k, const float* x, long long ldx,
                 const float* stats, float* dgamma, float* dbeta, cudaStream_t stream) {
  PK_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * L, stream));
  PK_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * L, stream));
  const dim3 grid((L + 127) / 128, std::min(N, 64));
  ln0_bwd_kernel<<<grid, 128, 0, stream>>>(G, ldg, N, L, Lout, k, x, ldx, stats, dgamma, dbeta);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int sinc_filters_fwd(const float* low_hz_, const float* band_hz_, int C, int k, float sr, float min_low, float min_band,
                     float* filt, cudaStream_t stream) {
  PK_REQUIRE(C > 0 && k > 2 && (k % 2) == 1, "sinc_filters_fwd: kernel size must be odd (got %d)", k);
  sinc_fwd_kernel<<<C, 128, k * sizeof(float), stream>>>(low_hz_, band_hz_, k, sr, min_low, min_band, filt);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int sinc_filters_bwd(const float* low_hz_, const float* band_hz_, int C, int k, float sr, float min_low, float min_band,
                     const float* dfilt, float* dlow, float* dband, cudaStream_t stream) {
  PK_REQUIRE(C > 0 && k > 2 && (k % 2) == 1, "sinc_filters_bwd: kernel size must be odd (got %d)", k);
  sinc_bwd_kernel<<<C, 1, 0, stream>>>(low_hz_, band_hz_, k, sr, min_low, min_band, dfilt, dlow, dband);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int conv_pack_weights(const float* w, int Co, int Ci, int k, __half* W16, int Cip, long long ldw, __half* Wf16, int Cop,
                      long long ldf, cudaStream_t stream) {
  PK_REQUIRE(w && (W16 || Wf16), "conv_pack_weights: null");
  conv_pack_kernel<<<148, 256, 0, stream>>>(w, Co, Ci, k, W16, Cip, ldw, Wf16, Cop, ldf);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int conv_im2col0(const float* x, long long ldx, int N, int L, int k, int Lout, __half* Xcol, int Kp, __half* XcolT,
                 long long ldp, cudaStream_t stream) {
  PK_REQUIRE(Kp >= k && Lout == L - k + 1, "conv_im2col0: bad sizes");
  const dim3 grid((L + 127) / 128, N);
  im2col0_kernel<<<grid, 256, (128 + k) * sizeof(float), stream>>>(x, ldx, N, L, k, Lout, Xcol, Kp, XcolT, ldp);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int conv_im2colT(const __half* A16, long long rows, int Cp, int Ci, int k, __half* XT, long long ldp, cudaStream_t stream) {
  const dim3 grid(static_cast<unsigned>((rows + 63) / 64), (Ci + 63) / 64, k);
  im2colT_kernel<<<grid, 256, 0, stream>>>(A16, rows, Cp, Ci, k, XT, ldp);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int conv_post_fwd(const ConvPostFwdArgs& f, cudaStream_t stream) {
  PK_REQUIRE(f.p >= 1 && f.p <= 255 && f.Lp == f.Lout / f.p && f.Lp >= 1, "conv_post_fwd: bad pooling sizes");
  PK_REQUIRE(f.gamma == nullptr || f.Lp > 1, "conv_post_fwd: LayerNorm over a single position");
  PostFwd a;
  a.O = f.O; a.ldo = f.ldo; a.N = f.N; a.L = f.L; a.Lout = f.Lout; a.p = f.p; a.Lp = f.Lp; a.C = f.C; a.act = f.act;
  a.gamma = f.gamma; a.beta = f.beta; a.eps = f.eps; a.keep = f.keep; a.P = f.P; a.arg = f.arg; a.stats = f.stats;
  a.A16n = f.A16n; a.Cpn = f.Cpn; a.Y32 = f.Y32;
  const int cmax = std::max(f.C, f.A16n ? f.Cpn : 0);
  const dim3 grid((cmax + 31) / 32, f.N), block(32, kNP);
  conv_post_fwd_kernel<<<grid, block, 0, stream>>>(a);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int conv_post_bwd(const ConvPostBwdArgs& f, cudaStream_t stream) {
  PostBwd a;
  a.dY = f.dY; a.sn = f.sn; a.sl = f.sl; a.sc = f.sc;
  a.N = f.N; a.L = f.L; a.Lout = f.Lout; a.p = f.p; a.Lp = f.Lp; a.C = f.C; a.act = f.act;
  a.gamma = f.gamma; a.beta = f.beta; a.eps = f.eps; a.keep = f.keep; a.P = f.P; a.arg = f.arg; a.stats = f.stats;
  a.dgamma = f.dgamma; a.dbeta = f.dbeta; a.dbias = f.dbias; a.dO = f.dO; a.amax_bits = f.amax_bits;
  if (f.gamma) {
    PK_CHECK_CUDA(cudaMemsetAsync(f.dgamma, 0, sizeof(float) * f.C * f.Lp, stream));
    PK_CHECK_CUDA(cudaMemsetAsync(f.dbeta, 0, sizeof(float) * f.C * f.Lp, stream));
  }
  if (f.dbias) PK_CHECK_CUDA(cudaMemsetAsync(f.dbias, 0, sizeof(float) * f.C, stream));
  const dim3 grid((f.C + 31) / 32, f.N), block(32, kNP);
  conv_post_bwd_kernel<<<grid, block, 0, stream>>>(a);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pk
