// pk_fusion.cu — companions of the FusionLinearConv layer (reference neural_networks.py:2057-2099; Titouan P. et al.,
// "FusionRNN: Shared Neural Parameters for Multi-Channel Distant Speech Recognition").
//
// The layer applies ONE shared affine map to each of the M microphone channels of a frame (Conv1d(1, C, kernel = d,
// stride = d) over the [M*d] feature vector), an activation, and reduces over the channels.  Here the shared map is one
// pk_gemm_tn over the zero-copy view [N*M, d] of the input (rows n*M + m = channel m of frame n); these kernels do
// the rest:  P[n][c] = red * sum_m act(O[n*M + m][c])  and its backward (dO, column sums for the bias gradient, the
// PReLU slope gradient).  Columns are C = G*Hh: G gate blocks (FusionLinearConv instances, e.g. wh and wz of a
// liGRU_layer) side by side, each with its own activation slope.
#include "pk_common.cuh"
#include "pk_kernels.h"

#include <algorithm>

namespace pk {

namespace {

// mode 0: piecewise linear act(x) = x > 0 ? x : slope[c / Hh] * x   (relu: 0, leaky_relu: 0.01, prelu: learnt)
// mode 1: tanh
__global__ void fusion_reduce_fwd_kernel(const float* __restrict__ O, long long ldo, long long N, int M, int C, int Hh,
                                         int mode, const float* __restrict__ slopes, float red, float* __restrict__ P,
                                         long long ldp) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sl = (mode == 0) ? __ldg(slopes + c / Hh) : 0.f;
  for (long long n = blockIdx.y * static_cast<long long>(blockDim.y) + threadIdx.y; n < N;
       n += static_cast<long long>(gridDim.y) * blockDim.y) {
    const float* o = O + n * M * ldo + c;
    float acc = 0.f;
    for (int m = 0; m < M; ++m) {
      const float v = o[m * ldo];
      acc += (mode == 0) ? (v > 0.f ? v : sl * v) : tanhf(v);
    }
    P[n * ldp + c] = acc * red;
  }
}

// dO[n*M+m][c] = red * dP[n][c] * act'(O[n*M+m][c]);  dbias[c] += sum dO;  dslope[c / Hh] += sum red * dP * O over O < 0
__global__ void fusion_reduce_bwd_kernel(const float* __restrict__ dP, long long lddp, const float* __restrict__ O,
                                         long long ldo, long long N, int M, int C, int Hh, int mode,
                                         const float* __restrict__ slopes, float red, float* __restrict__ dO,
                                         long long lddo, float* __restrict__ dbias, float* __restrict__ dslope) {
  __shared__ float sb[8][33];
  __shared__ float ss[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < C;
  const float sl = (ok && mode == 0) ? __ldg(slopes + c / Hh) : 0.f;
  float db = 0.f, ds = 0.f;
  if (ok) {
    for (long long n = blockIdx.y * 8LL + threadIdx.y; n < N; n += static_cast<long long>(gridDim.y) * 8) {
      const float g = dP[n * lddp + c] * red;
      const float* o = O + n * M * ldo + c;
      float* d = dO + n * M * lddo + c;
      for (int m = 0; m < M; ++m) {
        const float v = o[m * ldo];
        float dv;
        if (mode == 0) {
          dv = v > 0.f ? g : sl * g;
          if (!(v > 0.f)) ds = fmaf(g, v, ds);
        } else {
          const float t = tanhf(v);
          dv = g * (1.f - t * t);
        }
        d[m * lddo] = dv;
        db += dv;
      }
    }
  }
  sb[threadIdx.y][threadIdx.x] = db;
  ss[threadIdx.y][threadIdx.x] = ds;
  __syncthreads();
  if (threadIdx.y == 0 && ok) {
    float b = 0.f, s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { b += sb[j][threadIdx.x]; s += ss[j][threadIdx.x]; }
    atomicAdd(dbias + c, b);
    if (mode == 0 && dslope) atomicAdd(dslope + c / Hh, s);
  }
}

}  // namespace

int fusion_reduce_fwd(const float* O, long long ldo, long long N, int M, int C, int Hh, int mode, const float* slopes,
                      float red, float* P, long long ldp, cudaStream_t stream) {
  const dim3 block(32, 8);
  const long long gy = std::min<long long>((N + 7) / 8, 1184);
  const dim3 grid((C + 31) / 32, static_cast<unsigned>(gy > 0 ? gy : 1));
  fusion_reduce_fwd_kernel<<<grid, block, 0, stream>>>(O, ldo, N, M, C, Hh, mode, slopes, red, P, ldp);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int fusion_reduce_bwd(const float* dP, long long lddp, const float* O, long long ldo, long long N, int M, int C, int Hh,
                      int mode, const float* slopes, float red, float* dO, long long lddo, float* dbias, float* dslope,
                      cudaStream_t stream) {
  const dim3 block(32, 8);
  const long long gy = std::min<long long>((N + 7) / 8, 592);
  const dim3 grid((C + 31) / 32, static_cast<unsigned>(gy > 0 ? gy : 1));
  fusion_reduce_bwd_kernel<<<grid, block, 0, stream>>>(dP, lddp, O, ldo, N, M, C, Hh, mode, slopes, red, dO, lddo, dbias,
                                                       dslope);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pk
