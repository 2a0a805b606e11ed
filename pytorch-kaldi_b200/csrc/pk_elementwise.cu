// pk_elementwise.cu — HBM-bound companions of the GEMM / recurrent kernels.
//
// All of these are streaming kernels (coalesced, vectorised where the layout allows, grids
// sized in multiples of the SM count); none of them is reshaped into a GEMM.
//   transpose / convert : fp32 -> fp16 operand staging (row-major and channel-major copies)
//   amax_scale          : power-of-two loss scale for fp16 gradient operands
//   bn_finalize         : BatchNorm1d statistics -> folded scale/shift + running stats
//                         (reference neural_networks.py:1118-1124, nn.BatchNorm1d(momentum=0.05))
//   bn_bwd              : BatchNorm backward on the de-duplicated projection
//   logsoftmax_nll      : LogSoftmax(dim=1) + NLLLoss + argmax error in one pass
//                         (neural_networks.py:53-54, utils.py:2344-2381)
//   logsoftmax_bwd      : gradient of the above, emitted directly as fp16 GEMM operands
//   rmsprop / sgd       : fused optimizer steps over a flat parameter buffer (utils.py:2121-2162)
#include "pk_common.cuh"
#include "pk_kernels.h"

#include <algorithm>
#include <cmath>

namespace pk {

namespace {

constexpr int kSMs = 148;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------
// transpose (+ optional fp16 copies).  32x32 tiles through padded shared memory.
// ------------------------------------------------------------------------------------
__global__ void transpose_kernel(const float* __restrict__ in, long long ldi, int R, int C,
                                 float* __restrict__ outT, long long ldo, __half* __restrict__ outT16,
                                 long long ldo16, __half* __restrict__ in16, long long ldi16,
                                 const float* __restrict__ scale_dev, unsigned int* amax_bits) {
  __shared__ float tile[32][33];
  float amax = 0.f;
  const float s = scale_dev ? __ldg(scale_dev) : 1.f;
  const int tiles_c = (C + 31) / 32;
  const long long ntiles = static_cast<long long>((R + 31) / 32) * tiles_c;
  for (long long tidx = blockIdx.x; tidx < ntiles; tidx += gridDim.x) {
    const int r0 = static_cast<int>(tidx / tiles_c) * 32;
    const int c0 = static_cast<int>(tidx % tiles_c) * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int r = r0 + i, c = c0 + threadIdx.x;
      float v = 0.f;
      if (r < R && c < C) {
        v = in[static_cast<long long>(r) * ldi + c];
        amax = fmaxf(amax, fabsf(v));
        if (in16) in16[static_cast<long long>(r) * ldi16 + c] = f16_sat(v * s);
      }
      tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int c = c0 + i, r = r0 + threadIdx.x;
      if (r < R && c < C) {
        const float v = tile[threadIdx.x][i];
        if (outT) outT[static_cast<long long>(c) * ldo + r] = v;
        if (outT16) outT16[static_cast<long long>(c) * ldo16 + r] = f16_sat(v * s);
      }
    }
    __syncthreads();
  }
  if (amax_bits) {
    amax = warp_max(amax);
    if (threadIdx.x == 0 && amax > 0.f) atomicMax(amax_bits, __float_as_uint(fminf(amax, 3.0e38f)));
  }
}

__global__ void convert_kernel(const float* __restrict__ in, long long ldi, int R, int C,
                               __half* __restrict__ out, long long ldo, const float* __restrict__ scale_dev) {
  const float s = scale_dev ? __ldg(scale_dev) : 1.f;
  const long long total = static_cast<long long>(R) * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / C;
    const int c = static_cast<int>(i - r * C);
    out[r * ldo + c] = f16_sat(in[r * ldi + c] * s);
  }
}

// ------------------------------------------------------------------------------------
// amax -> power-of-two scale
// ------------------------------------------------------------------------------------
__global__ void amax_kernel(const float* __restrict__ x, long long ld, int R, int C, unsigned int* amax_bits) {
  float m = 0.f;
  const long long total = static_cast<long long>(R) * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / C;
    const int c = static_cast<int>(i - r * C);
    const float v = fabsf(x[r * ld + c]);
    if (v == v) m = fmaxf(m, v);  // ignore NaN
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(amax_bits, __float_as_uint(fminf(m, 3.0e38f)));
}
__global__ void scale_from_amax_kernel(unsigned int* amax_bits, float target_log2, float* scale_out, int rezero) {
  const float amax = __uint_as_float(*amax_bits);
  float s = 1.f;
  if (amax > 0.f) {
    int e;
    frexpf(amax, &e);  // amax = f * 2^e, f in [0.5, 1)  => amax < 2^e
    float ex = target_log2 - static_cast<float>(e);
    ex = fminf(fmaxf(ex, -100.f), 100.f);
    s = exp2f(ex);  // amax * s in [2^(target-1), 2^target)
  }
  scale_out[0] = s;
  scale_out[1] = 1.f / s;
  if (rezero) *amax_bits = 0u;
}

// ------------------------------------------------------------------------------------
// BatchNorm finalize
// ------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const double* __restrict__ stats, int C, long long n_unique,
                                   long long n_ref, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, int training,
                                   float* running_mean, float* running_var, long long* num_batches,
                                   float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && training && num_batches) *num_batches += 1;
  if (c >= C) return;
  float mean, rstd;
  if (training) {
    const double m = stats[2 * c] / static_cast<double>(n_unique);
    double var = stats[2 * c + 1] / static_cast<double>(n_unique) - m * m;
    if (var < 0.0) var = 0.0;
    mean = static_cast<float>(m);
    rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    if (running_mean) {
      // torch updates running_var with the unbiased estimate over the rows the reference
      // actually normalised (T*2B duplicated rows for a bidirectional layer)
      const double unbiased = (n_ref > 1) ? var * static_cast<double>(n_ref) / static_cast<double>(n_ref - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
    }
  } else {
    mean = running_mean[c];
    rstd = 1.f / sqrtf(running_var[c] + eps);
  }
  const float gmm = gamma ? gamma[c] : 1.f;
  const float bt = beta ? beta[c] : 0.f;
  const float sc = gmm * rstd;
  scale[c] = sc;
  shift[c] = bt - mean * sc;
  if (mean_out) mean_out[c] = mean;
  if (rstd_out) rstd_out[c] = rstd;
}

__global__ void fill_scale_shift_kernel(const float* __restrict__ bias, int C, float* scale, float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    scale[c] = 1.f;
    shift[c] = bias ? bias[c] : 0.f;
  }
}

// ------------------------------------------------------------------------------------
// BatchNorm backward, pass 1: per-channel sums over the T*B unique rows (one CTA per channel)
// ------------------------------------------------------------------------------------
__global__ void bn_bwd_sums_kernel(const BnBwdArgs a, double* __restrict__ sums /* [2][C] */) {
  const int c = blockIdx.x;
  const long long o0 = static_cast<long long>(c) * a.ldt;
  const long long o1 = o0 + static_cast<long long>(a.C) * a.ldt;
  const float inv_s = (a.GT16 && a.gscale) ? 1.f / __ldg(a.gscale) : 1.f;
  const float* p = a.PT ? a.PT + static_cast<long long>(c) * a.ldp : nullptr;
  const float mean = (a.use_bn && a.mean) ? a.mean[c] : 0.f;
  const float rstd = (a.use_bn && a.rstd) ? a.rstd[c] : 1.f;
  double s1 = 0.0, s2 = 0.0;
  for (long long i = threadIdx.x; i < a.n; i += blockDim.x) {
    float g = a.GT ? a.GT[o0 + i] : __half2float(a.GT16[o0 + i]) * inv_s;
    if (a.ndir == 2) g += a.GT ? a.GT[o1 + i] : __half2float(a.GT16[o1 + i]) * inv_s;
    s1 += g;
    if (a.use_bn) s2 += static_cast<double>(g) * ((p[i] - mean) * rstd);
  }
  __shared__ double red[2][32];
  s1 = warp_sum_d(s1);
  s2 = warp_sum_d(s2);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = s1; red[1][w] = s2; }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    s1 = l < nw ? red[0][l] : 0.0;
    s2 = l < nw ? red[1][l] : 0.0;
    s1 = warp_sum_d(s1);
    s2 = warp_sum_d(s2);
    if (l == 0) {
      sums[c] = s1;
      sums[a.C + c] = s2;
      if (a.dbeta) a.dbeta[c] = static_cast<float>(s1);
      if (a.dgamma && a.use_bn) a.dgamma[c] = static_cast<float>(s2);
    }
  }
}

// pass 2: dP = gamma*rstd*(g - mean(g) - p_hat*mean(g*p_hat)), written as scaled fp16 in both
// layouts (channel-major for dW = dP^T X, row-major for dX = dP W).  64x64 tiles through shared memory,
// every global access is a 4-byte (half2) or 8-byte (float2) vector along the contiguous axis.
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BnBwdArgs a, const double* __restrict__ sums) {
  __shared__ float tile[64][65];
  const float s = a.gscale ? __ldg(a.gscale) : 1.f;
  const float inv_s = (a.GT16 && a.gscale) ? 1.f / s : 1.f;
  const int tx = threadIdx.x, ty = threadIdx.y;  // (32, 8)
  const int tiles_i = static_cast<int>((a.n + 63) / 64);
  const long long ntiles = static_cast<long long>((a.C + 63) / 64) * tiles_i;
  const double inv_n = 1.0 / static_cast<double>(a.n);
  const bool vec_in = ((a.ldt & 1) == 0) && ((a.ldp & 1) == 0);
  for (long long tidx = blockIdx.x; tidx < ntiles; tidx += gridDim.x) {
    const int c0 = static_cast<int>(tidx / tiles_i) * 64;
    const long long i0 = static_cast<long long>(tidx % tiles_i) * 64;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int cr = ty + 8 * jj;
      const int c = c0 + cr;
      const long long i = i0 + 2 * tx;
      float v0 = 0.f, v1 = 0.f;
      if (c < a.C && i < a.n) {
        const bool two = (i + 1 < a.n);
        const long long o0 = static_cast<long long>(c) * a.ldt + i;
        const long long o1 = static_cast<long long>(a.C + c) * a.ldt + i;
        float g0, g1 = 0.f;
        if (a.GT) {
          g0 = a.GT[o0];
          if (two) g1 = a.GT[o0 + 1];
          if (a.ndir == 2) { g0 += a.GT[o1]; if (two) g1 += a.GT[o1 + 1]; }
        } else if (vec_in && two) {
          float2 f = __half22float2(*reinterpret_cast<const __half2*>(a.GT16 + o0));
          if (a.ndir == 2) {
            const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(a.GT16 + o1));
            f.x += f2.x; f.y += f2.y;
          }
          g0 = f.x * inv_s; g1 = f.y * inv_s;
        } else {
          g0 = __half2float(a.GT16[o0]);
          if (two) g1 = __half2float(a.GT16[o0 + 1]);
          if (a.ndir == 2) { g0 += __half2float(a.GT16[o1]); if (two) g1 += __half2float(a.GT16[o1 + 1]); }
          g0 *= inv_s; g1 *= inv_s;
        }
        if (a.use_bn) {
          const float rstd = a.rstd[c];
          const float gam = a.gamma ? a.gamma[c] : 1.f;
          if (a.training) {
            const long long po = static_cast<long long>(c) * a.ldp + i;
            float p0, p1 = 0.f;
            if (vec_in && two) { const float2 pp = *reinterpret_cast<const float2*>(a.PT + po); p0 = pp.x; p1 = pp.y; }
            else { p0 = a.PT[po]; if (two) p1 = a.PT[po + 1]; }
            const float mean = a.mean[c];
            const float mg = static_cast<float>(sums[c] * inv_n);
            const float mgp = static_cast<float>(sums[a.C + c] * inv_n);
            const float k = gam * rstd;
            v0 = k * (g0 - mg - (p0 - mean) * rstd * mgp);
            v1 = k * (g1 - mg - (p1 - mean) * rstd * mgp);
          } else {
            v0 = gam * rstd * g0;
            v1 = gam * rstd * g1;
          }
        } else {
          v0 = g0; v1 = g1;
        }
        v0 *= s; v1 *= s;
        if (!two) v1 = 0.f;
        if (a.dPT16) {
          const long long oo = static_cast<long long>(c) * a.ld16t + i;
          if (two && ((a.ld16t & 1) == 0)) *reinterpret_cast<__half2*>(a.dPT16 + oo) = __halves2half2(f16_sat(v0), f16_sat(v1));
          else { a.dPT16[oo] = f16_sat(v0); if (two) a.dPT16[oo + 1] = f16_sat(v1); }
        }
      }
      tile[cr][2 * tx] = v0;
      tile[cr][2 * tx + 1] = v1;
    }
    __syncthreads();
    if (a.dP16) {
#pragma unroll
      for (int ii = 0; ii < 8; ++ii) {
        const int ir = ty + 8 * ii;
        const long long i = i0 + ir;
        const int c = c0 + 2 * tx;
        if (i < a.n && c < a.C) {
          const float v0 = tile[2 * tx][ir];
          const float v1 = tile[2 * tx + 1][ir];
          const long long oo = i * a.ld16r + c;
          if (c + 1 < a.C && ((a.ld16r & 1) == 0)) *reinterpret_cast<__half2*>(a.dP16 + oo) = __halves2half2(f16_sat(v0), f16_sat(v1));
          else { a.dP16[oo] = f16_sat(v0); if (c + 1 < a.C) a.dP16[oo + 1] = f16_sat(v1); }
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------
// Dense (MLP hidden) layer epilogues: y = drop(act(scale * p + shift))  (reference neural_networks.py:138-148,
// BatchNorm folded into scale/shift, nn.Dropout's inverted scaling folded into the keep mask)
// ------------------------------------------------------------------------------------
// PT [C][ldp] channel-major projections -> YT16 [C][ld16t] (operand of the next dW), Y16 [n][ld16r] (operand of
// the next projection), optional Y32 [n][ld32] (module output).  keepT: fp16 [C][ldk] with 0 or 1/(1-p), or null.
__global__ void __launch_bounds__(256) dense_act_fwd_kernel(const DenseFwdArgs a) {
  __shared__ float tile[64][65];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tiles_i = static_cast<int>((a.n + 63) / 64);
  const long long ntiles = static_cast<long long>((a.C + 63) / 64) * tiles_i;
  for (long long tidx = blockIdx.x; tidx < ntiles; tidx += gridDim.x) {
    const int c0 = static_cast<int>(tidx / tiles_i) * 64;
    const long long i0 = static_cast<long long>(tidx % tiles_i) * 64;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int cr = ty + 8 * jj;
      const int c = c0 + cr;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const long long i = i0 + 2 * tx + e;
        float y = 0.f;
        if (c < a.C && i < a.n) {
          const float p = a.PT[static_cast<long long>(c) * a.ldp + i];
          y = act_fwd(a.act, fmaf(a.scale[c], p, a.shift[c]));
          if (a.keepT) y *= __half2float(a.keepT[static_cast<long long>(c) * a.ldk + i]);
          if (a.YT16) a.YT16[static_cast<long long>(c) * a.ld16t + i] = f16_sat(y);
        }
        tile[cr][2 * tx + e] = y;
      }
    }
    __syncthreads();
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) {
      const int ir = ty + 8 * ii;
      const long long i = i0 + ir;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = c0 + 2 * tx + e;
        if (i < a.n && c < a.C) {
          const float y = tile[2 * tx + e][ir];
          if (a.Y16) a.Y16[i * a.ld16r + c] = f16_sat(y);
          if (a.Y32) a.Y32[i * a.ld32 + c] = y;
        }
      }
    }
    __syncthreads();
  }
}

// dYT [C][ldy] fp32 (gradient w.r.t. the layer output, channel-major) -> GT16 [C][ldg] = fp16(scale * dY * keep *
// act'(y)), the input of pk_bn_bwd (ndir = 1).  y is recovered from the saved post-dropout YT16.
__global__ void dense_act_bwd_kernel(const DenseBwdArgs a) {
  const float s = a.gscale ? __ldg(a.gscale) : 1.f;
  const long long total = static_cast<long long>(a.C) * a.n;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long c = e / a.n;
    const long long i = e - c * a.n;
    const float dy = a.dYT[c * a.ldy + i];
    const float keep = a.keepT ? __half2float(a.keepT[c * a.ldk + i]) : 1.f;
    float g = 0.f;
    if (keep != 0.f) {
      const float y = __half2float(a.YT16[c * a.ld16t + i]) / keep;
      g = dy * keep * act_bwd_from_out(a.act, y);
    }
    a.GT16[c * a.ldg + i] = f16_sat(g * s);
  }
}

// ------------------------------------------------------------------------------------
// LogSoftmax + NLL + argmax error, one warp per row (row stays L1-resident across passes)
// ------------------------------------------------------------------------------------
__global__ void logsoftmax_nll_kernel(const HeadFwdArgs a, double* __restrict__ acc /* [2] */) {
  const int warps_per_block = blockDim.x >> 5;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  double loss_local = 0.0, err_local = 0.0;
  for (long long row = static_cast<long long>(blockIdx.x) * warps_per_block + w; row < a.N;
       row += static_cast<long long>(gridDim.x) * warps_per_block) {
    float* x = a.logits + row * a.ld;
    float m = -INFINITY;
    int mi = 0x7fffffff;
    for (int j = l; j < a.S; j += 32) {
      const float v = x[j];
      if (v > m) { m = v; mi = j; }
    }
    // warp arg-max with first-index tie break (torch.max(dim=1) on CPU returns the first maximum)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, o);
      const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
      if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
    }
    float ssum = 0.f;
    for (int j = l; j < a.S; j += 32) ssum += expf(x[j] - m);
    ssum = warp_sum(ssum);
    const float lse = m + logf(ssum);
    float xlab = 0.f;
    long long lab = -1;
    if (a.labels) {
      lab = a.labels[row];
      if (lab >= 0 && lab < a.S) xlab = x[lab];  // read the raw logit before the in-place update
    }
    __syncwarp();
    for (int j = l; j < a.S; j += 32) x[j] = x[j] - lse;
    if (a.labels && l == 0) {
      if (lab >= 0 && lab < a.S) loss_local += static_cast<double>(lse - xlab);
      err_local += (static_cast<long long>(mi) != lab) ? 1.0 : 0.0;
    }
  }
  if (a.labels) {
    __shared__ double red[2][32];
    if (l == 0) { red[0][w] = loss_local; red[1][w] = err_local; }
    __syncthreads();
    if (w == 0) {
      double s1 = l < warps_per_block ? red[0][l] : 0.0;
      double s2 = l < warps_per_block ? red[1][l] : 0.0;
      s1 = warp_sum_d(s1);
      s2 = warp_sum_d(s2);
      if (l == 0) {
        atomicAdd(acc, s1);
        atomicAdd(acc + 1, s2);
      }
    }
  }
}

// rowsum of dlogp (general log-softmax backward)
__global__ void rowsum_kernel(const float* __restrict__ d, long long ld, int N, int S, float* __restrict__ out) {
  const int warps_per_block = blockDim.x >> 5;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (long long row = static_cast<long long>(blockIdx.x) * warps_per_block + w; row < N;
       row += static_cast<long long>(gridDim.x) * warps_per_block) {
    float s = 0.f;
    for (int j = l; j < S; j += 32) s += d[row * ld + j];
    s = warp_sum(s);
    if (l == 0) out[row] = s;
  }
}

// 64 x 64 tiles, block (32, 8): every global access is a 2-element vector (float2 loads of the posteriors, half2
// stores of both fp16 layouts); the channel-major copy goes through a shared-memory transpose.
__global__ void __launch_bounds__(256) logsoftmax_bwd_kernel(const HeadBwdArgs a, const float* __restrict__ rowsum) {
  __shared__ float tile[64][65];
  const float oscale = a.out_scale * (a.scale_dev ? __ldg(a.scale_dev) : 1.f);
  const int tiles_c = (a.S + 63) / 64;
  const long long ntiles = static_cast<long long>((a.N + 63) / 64) * tiles_c;
  const int tx = threadIdx.x, ty = threadIdx.y;
  // vector paths need even leading dimensions / 8-byte (4-byte) aligned bases
  const bool vin = ((a.ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.logp) & 7) == 0) &&
                   (!a.dlogp || (((a.lddl & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.dlogp) & 7) == 0)));
  const bool v16 = a.d16 && ((a.ld16 & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.d16) & 3) == 0);
  const bool vt16 = a.dT16 && ((a.ld16t & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.dT16) & 3) == 0);
  for (long long tidx = blockIdx.x; tidx < ntiles; tidx += gridDim.x) {
    const long long r0 = (tidx / tiles_c) * 64;
    const int c0 = static_cast<int>(tidx % tiles_c) * 64;
    const int c = c0 + 2 * tx;
#pragma unroll 2
    for (int i = ty; i < 64; i += 8) {
      const long long r = r0 + i;
      float d0 = 0.f, d1 = 0.f;
      if (r < a.N && c < a.S) {
        const bool two = c + 1 < a.S;
        float l0, l1 = 0.f;
        if (vin && two) {
          const float2 l = *reinterpret_cast<const float2*>(a.logp + r * a.ld + c);
          l0 = l.x; l1 = l.y;
        } else {
          l0 = a.logp[r * a.ld + c];
          if (two) l1 = a.logp[r * a.ld + c + 1];
        }
        const float p0 = __expf(l0), p1 = __expf(l1);
        if (a.dlogp) {
          const float rs = rowsum[r];
          d0 = a.dlogp[r * a.lddl + c] - p0 * rs;
          if (two) d1 = a.dlogp[r * a.lddl + c + 1] - p1 * rs;
        } else {
          const long long lab = a.labels[r];
          d0 = (p0 - ((lab == c) ? 1.f : 0.f)) * a.gcoef;
          if (two) d1 = (p1 - ((lab == c + 1) ? 1.f : 0.f)) * a.gcoef;
        }
        if (a.d16) {
          if (v16 && two) {
            *reinterpret_cast<__half2*>(a.d16 + r * a.ld16 + c) = __halves2half2(f16_sat(d0 * oscale), f16_sat(d1 * oscale));
          } else {
            a.d16[r * a.ld16 + c] = f16_sat(d0 * oscale);
            if (two) a.d16[r * a.ld16 + c + 1] = f16_sat(d1 * oscale);
          }
        }
      }
      tile[i][2 * tx] = d0;
      tile[i][2 * tx + 1] = d1;
    }
    __syncthreads();
#pragma unroll 2
    for (int i = ty; i < 64; i += 8) {
      const int cc = c0 + i;
      const long long r = r0 + 2 * tx;
      const float d0 = tile[2 * tx][i], d1 = tile[2 * tx + 1][i];   // rows past N hold zeros
      if (a.dT16 && cc < a.S && r < a.N) {
        if (vt16 && r + 1 < a.N) {
          *reinterpret_cast<__half2*>(a.dT16 + static_cast<long long>(cc) * a.ld16t + r) =
              __halves2half2(f16_sat(d0 * oscale), f16_sat(d1 * oscale));
        } else {
          a.dT16[static_cast<long long>(cc) * a.ld16t + r] = f16_sat(d0 * oscale);
          if (r + 1 < a.N) a.dT16[static_cast<long long>(cc) * a.ld16t + r + 1] = f16_sat(d1 * oscale);
        }
      }
      if (a.dbias) {
        const float cs = warp_sum(d0 + d1);  // the warp spans the 64 rows of column cc (zero padded)
        if (tx == 0 && cc < a.S) atomicAdd(a.dbias + cc, cs);
      }
    }
    __syncthreads();
  }
}

__global__ void rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ v,
                               long long n, float lr, float alpha, float eps, float gscale) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * gscale;
    const float vi = alpha * v[i] + (1.f - alpha) * gi * gi;
    v[i] = vi;
    p[i] -= lr * gi / (sqrtf(vi) + eps);
  }
}
// torch.optim.Adam (utils.py:2131-2145; amsgrad off): bias corrections bc1 = 1 - b1^t, bc2 = 1 - b2^t from the host
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float rsqrt_bc2,
                            float gscale) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float pi = p[i];
    const float gi = fmaf(wd, pi, g[i] * gscale);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - (lr / bc1) * mi / (sqrtf(vi) * rsqrt_bc2 + eps);
  }
}
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, long long n, float lr, float gscale) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    p[i] -= lr * g[i] * gscale;
}

// ---- input side of the path: chunk preparation (data_io.py:255-272) and minibatch assembly
// (core.py:577-598) on the device ----
// expanded column j = (lag + left) * F + f holds fea[i + left + lag][f];  stats[j] = (sum, sumsq) in double
__global__ void chunk_stats_kernel(const float* __restrict__ fea, long long ldf, long long n_out, int F, int left, int right,
                                   double* __restrict__ stats) {
  const int W = left + right + 1;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= W * F) return;
  const int lagi = j / F, f = j - lagi * F;
  double s = 0.0, q = 0.0;
  for (long long i = blockIdx.y; i < n_out; i += gridDim.y) {
    const double v = fea[(i + lagi) * ldf + f];
    s += v;
    q += v * v;
  }
  atomicAdd(stats + 2 * j, s);
  atomicAdd(stats + 2 * j + 1, q);
}
__global__ void chunk_write_kernel(const float* __restrict__ fea, long long ldf, const long long* __restrict__ lab, long long lab_min,
                                   long long n_out, int F, int left, int right, const double* __restrict__ stats,
                                   float* __restrict__ out, long long ldo) {
  const int W = left + right + 1;
  const int cols = W * F + (lab ? 1 : 0);
  const long long total = n_out * cols;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i = e / cols;
    const int j = static_cast<int>(e - i * cols);
    float v;
    if (j < W * F) {
      const int lagi = j / F, f = j - lagi * F;
      const double mean = stats[2 * j] / n_out;
      const double var = stats[2 * j + 1] / n_out - mean * mean;  // np.std: population variance (data_io.py:263)
      v = static_cast<float>((static_cast<double>(fea[(i + lagi) * ldf + f]) - mean) / sqrt(var > 0.0 ? var : 0.0));
    } else {
      v = static_cast<float>(lab[i + left] - lab_min);               // data_io.py:266-272
    }
    out[i * ldo + j] = v;
  }
}
// inp[t][k][:] = data_set[beg[k] + t - left[k]][:] for left[k] <= t < left[k] + len[k], else 0   (core.py:584-595)
__global__ void batch_assemble_kernel(const float* __restrict__ data, long long ldd, int D, const long long* __restrict__ desc,
                                      int Bsz, int max_len, float* __restrict__ inp) {
  const long long total = static_cast<long long>(max_len) * Bsz * D;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(e % D);
    const long long tk = e / D;
    const int k = static_cast<int>(tk % Bsz);
    const long long t = tk / Bsz;
    const long long beg = desc[k], len = desc[Bsz + k], lz = desc[2 * Bsz + k];
    float v = 0.f;
    if (t >= lz && t < lz + len) v = data[(beg + t - lz) * ldd + c];
    inp[e] = v;
  }
}

// input side: Kaldi CompressedMatrix "CM " payload -> fp32 (data_io.py:1150-1196).  hdr [cols][4] uint16
// percentiles (0, 25, 75, 100), data [cols][rows] uint8 (column-major); out [rows][ldo].  Every operation is a
// separately rounded fp32 op in the reference's order, so the result is bit-identical to its numpy evaluation.
__global__ void cm_decode_kernel(const uint16_t* __restrict__ hdr, const uint8_t* __restrict__ data, float gmin, float grange,
                                 int rows, int cols, float* __restrict__ out, long long ldo) {
  const long long total = static_cast<long long>(rows) * cols;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(e / rows);
    const int r = static_cast<int>(e - static_cast<long long>(c) * rows);
    float p[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      p[k] = __fadd_rn(__fmul_rn(__fmul_rn(static_cast<float>(hdr[4 * c + k]), grange), 1.52590218966964e-05f), gmin);
    const unsigned d = data[e];
    float v;
    if (d <= 64u)
      v = __fadd_rn(p[0], __fmul_rn(__fdiv_rn(__fsub_rn(p[1], p[0]), 64.0f), static_cast<float>(d)));
    else if (d > 192u)
      v = __fadd_rn(p[2], __fmul_rn(__fdiv_rn(__fsub_rn(p[3], p[2]), 63.0f), static_cast<float>(d - 192u)));
    else
      v = __fadd_rn(p[1], __fmul_rn(__fdiv_rn(__fsub_rn(p[2], p[1]), 128.0f), static_cast<float>(d - 64u)));
    out[static_cast<long long>(r) * ldo + c] = v;
  }
}

// output side: x[n][c] -= v[c]  (log-posteriors -> scaled log-likelihoods, core.py:664-667)
__global__ void rows_sub_vec_kernel(float* __restrict__ x, long long ld, long long n, int C, const float* __restrict__ v) {
  const long long total = n * C;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = e / C;
    const int c = static_cast<int>(e - r * C);
    x[r * ld + c] -= __ldg(v + c);
  }
}

inline int grid_for(long long work_items, int per_block) {
  long long b = (work_items + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(kSMs) * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}


// ---------------------------------------------------------------------------------------------
// Reference LayerNorm (neural_networks.py:23-33: y = gamma (x - mean) / (std_unbiased + eps) + beta over the
// FEATURE axis) on channel-major activations PT[C][ld]: one thread per frame (column), fully coalesced along n.
// Forward normalises in place, keeps x_hat for the backward and (mean, std + eps) per frame.
// ---------------------------------------------------------------------------------------------
__global__ void ln_cm_fwd_kernel(float* __restrict__ PT, int C, long long n, long long ld, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, float* __restrict__ XH, float* __restrict__ stats) {
  const long long col = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (col >= n) return;
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += PT[c * ld + col];
  const float mean = s / C;
  float ss = 0.f;
  for (int c = 0; c < C; ++c) {
    const float d = PT[c * ld + col] - mean;
    ss = fmaf(d, d, ss);
  }
  const float sd = sqrtf(ss / (C - 1)) + eps;  // torch.std: unbiased; eps is added to the std
  const float r = 1.f / sd;
  for (int c = 0; c < C; ++c) {
    const float xh = (PT[c * ld + col] - mean) * r;
    if (XH) XH[c * ld + col] = xh;
    PT[c * ld + col] = fmaf(__ldg(gamma + c), xh, __ldg(beta + c));
  }
  if (stats) {
    stats[2 * col] = mean;
    stats[2 * col + 1] = sd;
  }
}

// Backward of the same: dY16 (fp16, loss-scaled, gradient w.r.t. the LayerNorm output, channel-major) -> gradient
// w.r.t. its input in both operand layouts (same scale), dgamma / dbeta un-scaled.
//   dxh = dy * gamma;  dx = (dxh - mean_c(dxh)) / sd - xh * sum_c(dxh * xh) / ((C - 1) * (sd - eps))
__global__ void ln_cm_bwd_kernel(__half* __restrict__ dT16, long long ld16t, __half* __restrict__ dR16, long long ld16r,
                                 const float* __restrict__ XH, long long ld, int C, long long n,
                                 const float* __restrict__ gamma, const float* __restrict__ stats, float eps,
                                 const float* __restrict__ scale /* [2]: loss scale, 1 / loss scale */,
                                 float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias) {
  const long long col = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool ok = col < n;
  const float inv_scale = __ldg(scale + 1);
  float s1 = 0.f, q = 0.f;
  if (ok)
    for (int c = 0; c < C; ++c) {
      const float dxh = __half2float(dT16[c * ld16t + col]) * __ldg(gamma + c);
      s1 += dxh;
      q = fmaf(dxh, XH[c * ld + col], q);
    }
  const float sd = ok ? stats[2 * col + 1] : 1.f;
  const float r = 1.f / sd;
  const float m1 = s1 / C;
  const float k2 = q / ((C - 1) * fmaxf(sd - eps, 1e-30f));
  for (int c = 0; c < C; ++c) {
    float dy = 0.f, xh = 0.f, dx = 0.f;
    if (ok) {
      dy = __half2float(dT16[c * ld16t + col]);
      xh = XH[c * ld + col];
      dx = (dy * __ldg(gamma + c) - m1) * r - xh * k2;
      const __half h = f16_sat(dx);
      dT16[c * ld16t + col] = h;
      if (dR16) dR16[col * ld16r + c] = h;
    }
    // per-channel sums over this block's frames -> one atomic per warp (dbias: the Linear bias in front of the norm)
    float a = dy * xh, b = dy, e = dx;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
      e += __shfl_xor_sync(0xffffffffu, e, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(dgamma + c, a * inv_scale);
      atomicAdd(dbeta + c, b * inv_scale);
      if (dbias) atomicAdd(dbias + c, e * inv_scale);
    }
  }
}

// per-channel (sum, sum of squares) of PT[C][ld] over n frames: BatchNorm batch statistics of a LayerNorm output
__global__ void row_stats_kernel(const float* __restrict__ PT, long long ld, long long n, double* __restrict__ stats) {
  const float* row = PT + static_cast<long long>(blockIdx.x) * ld;
  double s = 0.0, ss = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = row[i];
    s += v;
    ss += v * v;
  }
  __shared__ double sh[2][32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) { a += sh[0][w]; b += sh[1][w]; }
    stats[2 * blockIdx.x] = a;
    stats[2 * blockIdx.x + 1] = b;
  }
}

}  // namespace

int transpose_f32(const float* in, long long ldi, int R, int C, float* outT, long long ldo, __half* outT16,
                  long long ldo16, __half* in16, long long ldi16, const float* scale_dev, unsigned int* amax_bits,
                  cudaStream_t stream) {
  PK_REQUIRE(R > 0 && C > 0, "transpose: empty");
  const long long ntiles = static_cast<long long>((R + 31) / 32) * ((C + 31) / 32);
  transpose_kernel<<<grid_for(ntiles, 1), dim3(32, 8), 0, stream>>>(in, ldi, R, C, outT, ldo, outT16, ldo16,
                                                                     in16, ldi16, scale_dev, amax_bits);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int convert_f16(const float* in, long long ldi, int R, int C, __half* out, long long ldo,
                const float* scale_dev, cudaStream_t stream) {
  PK_REQUIRE(R > 0 && C > 0, "convert: empty");
  convert_kernel<<<grid_for(static_cast<long long>(R) * C, 1024), 256, 0, stream>>>(in, ldi, R, C, out, ldo,
                                                                                     scale_dev);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int amax_scale(const float* x, long long ld, int R, int C, float target_log2, float* amax_scratch,
               float* scale_out, cudaStream_t stream) {
  PK_CHECK_CUDA(cudaMemsetAsync(amax_scratch, 0, sizeof(float), stream));
  amax_kernel<<<grid_for(static_cast<long long>(R) * C, 2048), 256, 0, stream>>>(
      x, ld, R, C, reinterpret_cast<unsigned int*>(amax_scratch));
  scale_from_amax_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<unsigned int*>(amax_scratch), target_log2,
                                              scale_out, 0);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int amax_finalize(unsigned int* amax_bits, float target_log2, float* scale_out, cudaStream_t stream) {
  scale_from_amax_kernel<<<1, 1, 0, stream>>>(amax_bits, target_log2, scale_out, 1);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int bn_finalize(const double* stats, int C, long long n_unique, long long n_ref, const float* gamma,
                const float* beta, float eps, float momentum, int training, float* running_mean,
                float* running_var, long long* num_batches, float* scale, float* shift, float* mean_out,
                float* rstd_out, cudaStream_t stream) {
  PK_REQUIRE(training || (running_mean && running_var), "bn_finalize: eval mode needs running stats");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(stats, C, n_unique, n_ref, gamma, beta, eps,
                                                          momentum, training, running_mean, running_var,
                                                          num_batches, scale, shift, mean_out, rstd_out);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int fill_scale_shift(const float* bias, int C, float* scale, float* shift, cudaStream_t stream) {
  fill_scale_shift_kernel<<<(C + 127) / 128, 128, 0, stream>>>(bias, C, scale, shift);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int bn_bwd(const BnBwdArgs& a, cudaStream_t stream) {
  PK_REQUIRE(a.C > 0 && a.n > 0, "bn_bwd: empty");
  PK_REQUIRE(a.sums_scratch != nullptr, "bn_bwd: sums_scratch (2*C doubles) required");
  double* sums = a.sums_scratch;
  bn_bwd_sums_kernel<<<a.C, 256, 0, stream>>>(a, sums);
  const long long ntiles = static_cast<long long>((a.C + 63) / 64) * ((a.n + 63) / 64);
  bn_bwd_apply_kernel<<<grid_for(ntiles, 1), dim3(32, 8), 0, stream>>>(a, sums);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int logsoftmax_nll(const HeadFwdArgs& a, cudaStream_t stream) {
  PK_REQUIRE(a.N > 0 && a.S > 0, "logsoftmax_nll: empty");
  PK_REQUIRE(!a.labels || a.acc, "logsoftmax_nll: acc (2 doubles: loss sum, error count) required with labels");
  if (a.acc) PK_CHECK_CUDA(cudaMemsetAsync(a.acc, 0, 2 * sizeof(double), stream));
  logsoftmax_nll_kernel<<<grid_for(a.N, 8), 256, 0, stream>>>(a, a.acc);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int logsoftmax_bwd(const HeadBwdArgs& a, cudaStream_t stream) {
  PK_REQUIRE(a.N > 0 && a.S > 0, "logsoftmax_bwd: empty");
  PK_REQUIRE(a.dlogp || a.labels, "logsoftmax_bwd: need labels (fused NLL) or dlogp (general)");
  float* rowsum = a.rowsum_scratch;
  if (a.dlogp) {
    PK_REQUIRE(rowsum != nullptr, "logsoftmax_bwd: rowsum_scratch (N floats) required in general mode");
    rowsum_kernel<<<grid_for(a.N, 8), 256, 0, stream>>>(a.dlogp, a.lddl, a.N, a.S, rowsum);
  }
  if (a.dbias) PK_CHECK_CUDA(cudaMemsetAsync(a.dbias, 0, sizeof(float) * a.S, stream));
  const long long ntiles = static_cast<long long>((a.N + 63) / 64) * ((a.S + 63) / 64);
  logsoftmax_bwd_kernel<<<grid_for(ntiles, 1), dim3(32, 8), 0, stream>>>(a, rowsum);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int dense_act_fwd(const DenseFwdArgs& a, cudaStream_t stream) {
  PK_REQUIRE(a.C > 0 && a.n > 0, "dense_act_fwd: empty");
  const long long ntiles = static_cast<long long>((a.C + 63) / 64) * ((a.n + 63) / 64);
  dense_act_fwd_kernel<<<grid_for(ntiles, 1), dim3(32, 8), 0, stream>>>(a);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int dense_act_bwd(const DenseBwdArgs& a, cudaStream_t stream) {
  PK_REQUIRE(a.C > 0 && a.n > 0, "dense_act_bwd: empty");
  dense_act_bwd_kernel<<<grid_for(static_cast<long long>(a.C) * a.n, 1024), 256, 0, stream>>>(a);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int rmsprop_step(float* p, const float* g, float* v, long long n, float lr, float alpha, float eps,
                 float gscale, cudaStream_t stream) {
  if (n <= 0) return 0;
  rmsprop_kernel<<<grid_for(n, 1024), 256, 0, stream>>>(p, g, v, n, lr, alpha, eps, gscale);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float wd,
              long long step, float gscale, cudaStream_t stream) {
  if (n <= 0) return 0;
  PK_REQUIRE(step >= 1, "adam_step: step counts from 1");
  const double bc1 = 1.0 - pow(static_cast<double>(b1), static_cast<double>(step));
  const double bc2 = 1.0 - pow(static_cast<double>(b2), static_cast<double>(step));
  adam_kernel<<<grid_for(n, 1024), 256, 0, stream>>>(p, g, m, v, n, lr, b1, b2, eps, wd, static_cast<float>(bc1),
                                                     static_cast<float>(1.0 / sqrt(bc2)), gscale);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int chunk_prepare(const float* fea, long long ldf, const long long* lab, long long lab_min, long long n_in, int F, int left,
                  int right, double* stats, float* out, long long ldo, cudaStream_t stream) {
  const long long n_out = n_in - left - right;
  PK_REQUIRE(n_out > 0 && F > 0 && left >= 0 && right >= 0, "chunk_prepare: bad sizes");
  const int W = left + right + 1;
  PK_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * W * F, stream));
  const dim3 g1((W * F + 127) / 128, static_cast<unsigned>(std::min<long long>(n_out, 512)));
  chunk_stats_kernel<<<g1, 128, 0, stream>>>(fea, ldf, n_out, F, left, right, stats);
  chunk_write_kernel<<<grid_for(n_out * (W * F + 1), 1024), 256, 0, stream>>>(fea, ldf, lab, lab_min, n_out, F, left, right, stats,
                                                                            out, ldo);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int batch_assemble(const float* data, long long ldd, int D, const long long* desc, int Bsz, int max_len, float* inp,
                   cudaStream_t stream) {
  PK_REQUIRE(Bsz > 0 && max_len > 0 && D > 0, "batch_assemble: bad sizes");
  batch_assemble_kernel<<<grid_for(static_cast<long long>(max_len) * Bsz * D, 1024), 256, 0, stream>>>(data, ldd, D, desc, Bsz,
                                                                                                     max_len, inp);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int cm_decode(const uint16_t* hdr, const uint8_t* data, float gmin, float grange, int rows, int cols, float* out, long long ldo,
              cudaStream_t stream) {
  PK_REQUIRE(rows > 0 && cols > 0, "cm_decode: empty matrix");
  cm_decode_kernel<<<grid_for(static_cast<long long>(rows) * cols, 1024), 256, 0, stream>>>(hdr, data, gmin, grange, rows, cols, out,
                                                                                          ldo);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int rows_sub_vec(float* x, long long ld, long long n, int C, const float* v, cudaStream_t stream) {
  if (n <= 0 || C <= 0) return 0;
  rows_sub_vec_kernel<<<grid_for(n * C, 1024), 256, 0, stream>>>(x, ld, n, C, v);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int sgd_step(float* p, const float* g, long long n, float lr, float gscale, cudaStream_t stream) {
  if (n <= 0) return 0;
  sgd_kernel<<<grid_for(n, 1024), 256, 0, stream>>>(p, g, n, lr, gscale);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}


int ln_cm_fwd(float* PT, int C, long long n, long long ld, const float* gamma, const float* beta, float eps, float* XH,
              float* stats, cudaStream_t stream) {
  PK_REQUIRE(C > 1 && n > 0, "ln_cm_fwd: need C > 1 features and n > 0 frames");
  ln_cm_fwd_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, stream>>>(PT, C, n, ld, gamma, beta, eps, XH, stats);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int ln_cm_bwd(__half* dT16, long long ld16t, __half* dR16, long long ld16r, const float* XH, long long ld, int C, long long n,
              const float* gamma, const float* stats, float eps, const float* scale, float* dgamma, float* dbeta,
              float* dbias, cudaStream_t stream) {
  PK_REQUIRE(C > 1 && n > 0, "ln_cm_bwd: need C > 1 features and n > 0 frames");
  PK_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * C, stream));
  PK_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * C, stream));
  if (dbias) PK_CHECK_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * C, stream));
  ln_cm_bwd_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, stream>>>(dT16, ld16t, dR16, ld16r, XH, ld, C, n, gamma, stats,
                                                                               eps, scale, dgamma, dbeta, dbias);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int row_stats(const float* PT, int C, long long n, long long ld, double* stats, cudaStream_t stream) {
  row_stats_kernel<<<C, 256, 0, stream>>>(PT, ld, n, stats);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pk
