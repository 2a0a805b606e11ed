// pk_gemm.cu — C[M,N] = alpha * A[M,K] * B[N,K]^T (+bias) on tcgen05 tensor cores.
//
// Both operands are K-major ("TN"): every projection of the acoustic-model path is laid out
// so that the contraction index is contiguous ("Data layout in HBM").
//   * TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) stages 128x{64 f16 | 32 tf32} operand
//     tiles through a 6-deep shared-memory ring guarded by full/empty mbarriers;
//   * one elected thread issues tcgen05.mma.cta_group::1 (128x128xK16/K8) with the fp32
//     accumulator tile living in tensor memory (128 lanes x 128 columns);
//   * four epilogue warps pull the accumulator back with tcgen05.ld (32 lanes x 32 columns per
//     instruction) and fuse: alpha / device-side scale, bias along N or along M, per-row
//     sum / sum-of-squares (BatchNorm batch statistics of a channel-major projection), plain
//     or atomic (split-K / accumulate) stores.
//
// Replaces: every `nn.Linear` call of the reference module zoo (neural_networks.py:1114-1115
// liGRU input projections, :432-435 LSTM, :609-611 GRU, MLP :138-148) and the autograd GEMMs
// behind them.
#include "pk_common.cuh"
#include "pk_kernels.h"

#include <algorithm>
#include <cstdlib>
#include <mutex>

namespace pk {

namespace {

constexpr int kBM = 128;
constexpr int kBN = 128;
#ifndef PK_GEMM_STAGES
#define PK_GEMM_STAGES 3
#endif
constexpr int kStages = PK_GEMM_STAGES;  // 3 stages (98 KB) -> two CTAs per SM: one tile's epilogue overlaps the other's mainloop
constexpr int kTileBytes = kBM * 128;  // 128 rows x 128 bytes
constexpr int kGemmThreads = 192;
// BN = 128: kStages stages of (16 + 16) KB -> two CTAs per SM.  BN = 256 (wide-N problems): 4 stages of (16 + 32) KB,
// one CTA per SM, 1.36x the operand reuse per byte read from L2 (the 128x128 tile is L2-bandwidth bound on B200).
constexpr int stages_for(int bn) { return bn == 256 ? 4 : kStages; }
constexpr int smem_for(int bn) { return stages_for(bn) * (kTileBytes + bn * 128) + 256 + 1024; }  // + barriers + align slack

struct GemmDev {
  int M, N, K;
  float* C;
  long long ldc;
  const float* bias;
  int bias_mode;  // 0 none, 1 along N (C[m][n] += bias[n]), 2 along M (C[m][n] += bias[m])
  double* rowstats;  // [M][2] (sum, sumsq) or null
  float alpha;
  const float* alpha_dev;  // optional device multiplier (e.g. 1/loss_scale)
  int atomic;              // 1: atomicAdd into C
  int kb_per_split;
  int bk_elems;
  int a_k0, b_k0;  // element offsets of the contraction range inside each operand's rows
  unsigned int* amax_bits;  // optional: atomicMax of |C| (float bits) for the next loss scale
};

template <int DT, int BN>  // DT: 0 = f16 operands, 2 = tf32 (fp32 operands); BN: tile width 128 / 256
__global__ void __launch_bounds__(kGemmThreads, (BN == 128 && kStages <= 3) ? 2 : 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  constexpr int NST = stages_for(BN);
  constexpr int kBTile = BN * 128;
  uint8_t* sA = smem;
  uint8_t* sB = smem + NST * kTileBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NST * (kTileBytes + kBTile));
  uint64_t* empty_bar = full_bar + NST;
  uint64_t* acc_bar = empty_bar + NST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_blk = blockIdx.x;
  const int n_blk = blockIdx.y;
  const int kb_total = (p.K + p.bk_elems - 1) / p.bk_elems;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(kb_total, kb_begin + p.kb_per_split);
  const int nkb = kb_end - kb_begin;
  if (nkb <= 0) return;  // uniform across the CTA (only possible for trailing splits)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % NST;
        const uint32_t ph = (i / NST) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], kTileBytes + kBTile);
        const int kc = (kb_begin + i) * p.bk_elems;
        tma_load_2d(sA + s * kTileBytes, &tmA, &full_bar[s], kc + p.a_k0, m_blk * kBM);
        tma_load_2d(sB + s * kBTile, &tmB, &full_bar[s], kc + p.b_k0, n_blk * BN);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc(DT, kBM, BN);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % NST;
        const uint32_t ph = (i / NST) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + s * kTileBytes);
        const uint32_t b_base = smem_u32(sB + s * kBTile);
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 4 x 32 bytes = one 128-byte swizzle row
          const uint64_t ad = umma_desc_k_sw128(a_base + k * 32);
          const uint64_t bd = umma_desc_k_sw128(b_base + k * 32);
          if (DT == 0)
            umma_f16(tmem_base, ad, bd, idesc, (i | k) != 0);
          else
            umma_tf32(tmem_base, ad, bd, idesc, (i | k) != 0);
        }
        umma_commit(&empty_bar[s]);  // frees the stage once these MMAs retire
      }
      umma_commit(acc_bar);
    }
    __syncwarp();
  } else {
    // ---- epilogue: warps 2..5 own TMEM lane groups (warp % 4) ----
    const int lg = warp & 3;
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const int row = lg * 32 + lane;
    const long long gm = static_cast<long long>(m_blk) * kBM + row;
    float alpha = p.alpha;
    if (p.alpha_dev) alpha *= __ldg(p.alpha_dev);
    const bool add_bias = (p.bias != nullptr) && (blockIdx.z == 0);
    float bias_m = 0.f;
    if (add_bias && p.bias_mode == 2 && gm < p.M) bias_m = __ldg(p.bias + gm);
    double s1 = 0.0, s2 = 0.0;
    float amax = 0.f;
    float* crow = p.C + gm * p.ldc;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + c0, v);
      tmem_ld_wait();
      const int gn0 = n_blk * BN + c0;
      if (gm < p.M && gn0 < p.N) {
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = alpha * __uint_as_float(v[j]);
          if (add_bias) x += (p.bias_mode == 1) ? ((gn0 + j < p.N) ? __ldg(p.bias + gn0 + j) : 0.f)
                                                : bias_m;
          o[j] = x;
        }
        if (p.amax_bits) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (gn0 + j < p.N) amax = fmaxf(amax, fabsf(o[j]));
        }
        if (p.rowstats) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (gn0 + j < p.N) {
              s1 += o[j];
              s2 += static_cast<double>(o[j]) * o[j];
            }
        }
        if (p.atomic) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (gn0 + j < p.N) atomicAdd(crow + gn0 + j, o[j]);
        } else if (vec_ok && gn0 + 32 <= p.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(crow + gn0 + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (gn0 + j < p.N) crow[gn0 + j] = o[j];
        }
      }
    }
    if (p.amax_bits) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      if (lane == 0 && amax > 0.f) atomicMax(p.amax_bits, __float_as_uint(fminf(amax, 3.0e38f)));
    }
    if (p.rowstats && gm < p.M) {
      atomicAdd(p.rowstats + 2 * gm, s1);
      atomicAdd(p.rowstats + 2 * gm + 1, s2);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BN>(tmem_base);
}

// ---- persistent 128x256 variant: one CTA per SM walks over tiles; the fp32 accumulator is double-buffered in
// TMEM (2 x 256 columns), so the epilogue of tile j (tcgen05.ld -> scale/bias/statistics -> global) overlaps the
// TMA + MMA mainloop of tile j+1.  Used for single-pass f16 problems with N >= 1024 (projections, dX, logits).
constexpr int kPBN = 256;
constexpr int kPStages = 3;
constexpr int kPThreads = 320;  // TMA warp, MMA warp, 8 epilogue warps (4 TMEM lane groups x 2 column halves)
constexpr int kEpiRow = 36;     // floats per staged row (32 + 4 pad: conflict-free float4 both ways)
constexpr int kEpiBytes = 8 * 32 * kEpiRow * 4;  // per-warp 32 x 32 transposition buffers
constexpr int kPSmemBytes = kPStages * (kTileBytes + kPBN * 128) + kEpiBytes + 256 + 1024;

__global__ void __launch_bounds__(kPThreads, 1)
gemm_tn_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmDev p,
                       int m_tiles, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  constexpr int kBTile = kPBN * 128;
  uint8_t* sA = smem;
  uint8_t* sB = smem + kPStages * kTileBytes;
  float* epi = reinterpret_cast<float*>(smem + kPStages * (kTileBytes + kBTile));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kPStages * (kTileBytes + kBTile) + kEpiBytes);
  uint64_t* empty_bar = full_bar + kPStages;
  uint64_t* tfull_bar = empty_bar + kPStages;   // [2] accumulator ready for the epilogue
  uint64_t* tempty_bar = tfull_bar + 2;         // [2] accumulator drained by the 4 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkb = (p.K + p.bk_elems - 1) / p.bk_elems;
  const int total = m_tiles * n_tiles;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kPStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int it = 0;  // running k-block counter across tiles
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;  // m fastest: concurrent CTAs share the B tile in L2
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % kPStages;
          const uint32_t ph = (it / kPStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], kTileBytes + kBTile);
          const int kc = i * p.bk_elems;
          tma_load_2d(sA + s * kTileBytes, &tmA, &full_bar[s], kc + p.a_k0, m_blk * kBM);
          tma_load_2d(sB + s * kBTile, &tmB, &full_bar[s], kc + p.b_k0, n_blk * kPBN);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc(0, kBM, kPBN);
      int it = 0, j = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++j) {
        const int b = j & 1;
        const uint32_t use = static_cast<uint32_t>(j >> 1);
        mbar_wait(&tempty_bar[b], (use & 1) ^ 1);   // epilogue has drained this accumulator (first use passes)
        tc_fence_after();
        const uint32_t acc = tmem_base + static_cast<uint32_t>(b * kPBN);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % kPStages;
          const uint32_t ph = (it / kPStages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + s * kTileBytes);
          const uint32_t b_base = smem_u32(sB + s * kBTile);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = umma_desc_k_sw128(a_base + k * 32);
            const uint64_t bd = umma_desc_k_sw128(b_base + k * 32);
            umma_f16(acc, ad, bd, idesc, (i | k) != 0);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tfull_bar[b]);
      }
    }
    __syncwarp();
  } else {
    const int lg = warp & 3;               // TMEM lane group this warp may read (warp id % 4)
    const int chalf = (warp - 2) >> 2;     // which 128 columns of the 256-wide accumulator
    float alpha = p.alpha;
    if (p.alpha_dev) alpha *= __ldg(p.alpha_dev);
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    float amax = 0.f;
    int j = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++j) {
      const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
      const int b = j & 1;
      const uint32_t use = static_cast<uint32_t>(j >> 1);
      mbar_wait(&tfull_bar[b], use & 1);
      tc_fence_after();
      const int row = lg * 32 + lane;
      const long long gm = static_cast<long long>(m_blk) * kBM + row;
      const bool add_bias = p.bias != nullptr;
      float bias_m = 0.f;
      if (add_bias && p.bias_mode == 2 && gm < p.M) bias_m = __ldg(p.bias + gm);
      double s1 = 0.0, s2 = 0.0;
      float* crow = p.C + gm * p.ldc;
#pragma unroll 1
      for (int c0 = chalf * 128; c0 < chalf * 128 + 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + static_cast<uint32_t>(b * kPBN + c0), v);
        tmem_ld_wait();
        const int gn0 = n_blk * kPBN + c0;
        float o[32];
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) o[jj] = 0.f;
        if (gm < p.M && gn0 < p.N) {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            float x = alpha * __uint_as_float(v[jj]);
            if (add_bias) x += (p.bias_mode == 1) ? ((gn0 + jj < p.N) ? __ldg(p.bias + gn0 + jj) : 0.f) : bias_m;
            o[jj] = x;
          }
          if (p.amax_bits) {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj)
              if (gn0 + jj < p.N) amax = fmaxf(amax, fabsf(o[jj]));
          }
          if (p.rowstats) {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj)
              if (gn0 + jj < p.N) {
                s1 += o[jj];
                s2 += static_cast<double>(o[jj]) * o[jj];
              }
          }
        }
        // lane = row in registers -> stage the 32 x 32 block and write it row-contiguous: 4 full 128-byte lines per
        // store instruction instead of 32 scattered 16-byte pieces (the L1 store path was the epilogue's limiter)
        float* tw = epi + (warp - 2) * 32 * kEpiRow;
#pragma unroll
        for (int jj = 0; jj < 32; jj += 4)
          *reinterpret_cast<float4*>(tw + lane * kEpiRow + jj) = make_float4(o[jj], o[jj + 1], o[jj + 2], o[jj + 3]);
        __syncwarp();
        if (gn0 < p.N) {
          const int cc = (lane & 7) * 4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = (lane >> 3) + 4 * i;
            const long long gmr = static_cast<long long>(m_blk) * kBM + lg * 32 + r;
            if (gmr < p.M) {
              const float4 q = *reinterpret_cast<const float4*>(tw + r * kEpiRow + cc);
              float* dst = p.C + gmr * p.ldc + gn0 + cc;
              if (vec_ok && gn0 + cc + 4 <= p.N) {
                *reinterpret_cast<float4*>(dst) = q;
              } else {
                if (gn0 + cc < p.N) dst[0] = q.x;
                if (gn0 + cc + 1 < p.N) dst[1] = q.y;
                if (gn0 + cc + 2 < p.N) dst[2] = q.z;
                if (gn0 + cc + 3 < p.N) dst[3] = q.w;
              }
            }
          }
        }
        __syncwarp();
      }
      // this warp's TMEM reads of the buffer are complete -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[b]);
      if (p.rowstats && gm < p.M) {
        atomicAdd(p.rowstats + 2 * gm, s1);
        atomicAdd(p.rowstats + 2 * gm + 1, s2);
      }
    }
    if (p.amax_bits) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      if (lane == 0 && amax > 0.f) atomicMax(p.amax_bits, __float_as_uint(fminf(amax, 3.0e38f)));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---- host side: tensor-map encoding through the driver entry point (no -lcuda needed) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

int make_operand_map(CUtensorMap* map, const void* base, int dtype, long long rows, long long k,
                     long long ld_elems, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  PK_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  const int esz = (dtype == PK_DT_F16) ? 2 : 4;
  PK_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "GEMM operand base not 16-byte aligned");
  PK_REQUIRE((ld_elems * esz) % 16 == 0, "GEMM operand row pitch (%lld elems) not 16-byte multiple",
             ld_elems);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld_elems) * esz};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esz), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dtype == PK_DT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                   2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PK_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%lld k=%lld ld=%lld", (int)r,
             rows, k, ld_elems);
  return 0;
}

}  // namespace

int gemm_tn(const GemmArgs& a, cudaStream_t stream) {
  PK_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm_tn: empty problem M=%d N=%d K=%d", a.M, a.N, a.K);
  PK_REQUIRE(a.dtype == PK_DT_F16 || a.dtype == PK_DT_TF32, "gemm_tn: bad dtype %d", a.dtype);
  PK_REQUIRE(a.bias_mode >= 0 && a.bias_mode <= 2, "gemm_tn: bad bias mode");
  static PerDeviceOnce attr_once;
  const cudaError_t attr_err = attr_once.run([] {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel<0, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_for(128));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tn_kernel<2, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_for(128));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tn_kernel<0, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_for(256));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tn_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmemBytes);
    return e;
  });
  PK_CHECK_CUDA(attr_err);

  CUtensorMap tmA, tmB;
  // the K extent of each map bounds what TMA may read; everything beyond is zero-filled, so a
  // shifted contraction (a_k0 != b_k0) stays exact as long as one side runs out of bounds
  PK_REQUIRE(a.a_k0 >= 0 && a.b_k0 >= 0, "gemm_tn: negative k offset");
  {
    const int esz = (a.dtype == PK_DT_F16) ? 2 : 4;
    PK_REQUIRE((a.a_k0 * esz) % 16 == 0 && (a.b_k0 * esz) % 16 == 0,
               "gemm_tn: k offsets must be multiples of 16 bytes (TMA box start alignment)");
  }
  const long long a_ext = a.a_kext > 0 ? a.a_kext : a.a_k0 + a.K;
  const long long b_ext = a.b_kext > 0 ? a.b_kext : a.b_k0 + a.K;
  // wide tile for wide-N single-pass f16 problems (projections, dX, logits); env PK_GEMM_BN=128 pins the narrow one
  static const bool allow256 = [] { const char* e = getenv("PK_GEMM_BN"); return !(e && atoi(e) == 128); }();
  const bool wide = allow256 && a.dtype == PK_DT_F16 && a.N >= 1024 && a.split_k <= 1;
  const int bn = wide ? 256 : 128;
  if (int rc = make_operand_map(&tmA, a.A, a.dtype, a.M, a_ext, a.lda, kBM)) return rc;
  if (int rc = make_operand_map(&tmB, a.B, a.dtype, a.N, b_ext, a.ldb, bn)) return rc;

  GemmDev p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.C = a.C; p.ldc = a.ldc;
  p.bias = a.bias; p.bias_mode = a.bias ? a.bias_mode : 0;
  p.rowstats = a.rowstats;
  p.alpha = a.alpha;
  p.alpha_dev = a.alpha_dev;
  p.bk_elems = (a.dtype == PK_DT_F16) ? 64 : 32;
  p.amax_bits = a.amax_bits;
  PK_REQUIRE(!(a.amax_bits && (a.accumulate || a.split_k > 1)), "gemm_tn: amax needs a single-pass store epilogue");
  p.a_k0 = static_cast<int>(a.a_k0);
  p.b_k0 = static_cast<int>(a.b_k0);
  const int kb_total = (a.K + p.bk_elems - 1) / p.bk_elems;
  int splits = a.split_k > 0 ? a.split_k : 1;
  if (splits > kb_total) splits = kb_total;
  p.kb_per_split = (kb_total + splits - 1) / splits;
  splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;
  PK_REQUIRE(!(a.rowstats && splits > 1), "gemm_tn: rowstats with split-K unsupported");
  p.atomic = (splits > 1 || a.accumulate) ? 1 : 0;
  if (splits > 1 && !a.accumulate) {
    if (a.ldc == a.N) {
      PK_CHECK_CUDA(cudaMemsetAsync(a.C, 0, sizeof(float) * (size_t)a.M * a.N, stream));
    } else {
      PK_CHECK_CUDA(cudaMemset2DAsync(a.C, sizeof(float) * a.ldc, 0, sizeof(float) * a.N, a.M, stream));
    }
  }
  dim3 grid((a.M + kBM - 1) / kBM, (a.N + bn - 1) / bn, splits);
  static const int persist_mode = [] { const char* e = getenv("PK_GEMM_PERSIST"); return e ? atoi(e) : 1; }();
  if (wide && persist_mode && !p.atomic) {
    static int sms = 0;
    if (!sms) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int total = static_cast<int>(grid.x * grid.y);
    gemm_tn_persist_kernel<<<std::min(total, sms), kPThreads, kPSmemBytes, stream>>>(tmA, tmB, p, static_cast<int>(grid.x),
                                                                                        static_cast<int>(grid.y));
  } else if (wide)
    gemm_tn_kernel<0, 256><<<grid, kGemmThreads, smem_for(256), stream>>>(tmA, tmB, p);
  else if (a.dtype == PK_DT_F16)
    gemm_tn_kernel<0, 128><<<grid, kGemmThreads, smem_for(128), stream>>>(tmA, tmB, p);
  else
    gemm_tn_kernel<2, 128><<<grid, kGemmThreads, smem_for(128), stream>>>(tmA, tmB, p);
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pk
