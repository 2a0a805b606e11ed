// pk_cell_step.cu — step-wise recurrent path: one fused tensor-core kernel per time step.
//
// Covers the gate families / sizes the register-resident persistent kernels (pk_rnn_ws.cu) do not hold:
// LSTM (4 gates, reference neural_networks.py:300-483) and liGRU with H > 560 (e.g. the 5x1024 Librispeech
// stress shape).  All T launches of a layer are issued from inside ONE C-ABI call (no Python per step).
//
// Per step (forward):  pre[rows, NG*8 units] = S16[rows, K] . Wp[unit-tile]^T   on mma.sync (fp16 in, fp32 acc)
//   * S16 = fp16 copy of the previous state, [rows][KPs] row-major, double buffered by step parity;
//   * Wp  = recurrent weights packed once per call as [unit tile][gate][8 units][KPs] fp16, so a CTA's
//     operand block is one contiguous, 16-byte-aligned range that is copied to shared memory with cp.async
//     (rows padded by 8 halves -> conflict-free ldmatrix); the packed weights stay L2-resident (<= 5 MB);
//   * the epilogue applies BatchNorm scale/shift, gate non-linearities, dropout mask and the state update on
//     fp32 state kept in a [rows][H] buffer, and writes the saved tensors / fp16 operands channel-major.
// Backward: carry_h[rows, 8 units] = G16[rows, NG*K] . UTp[unit-tile]^T  (K chunked per gate), then the
// pointwise backward of the step, which produces the next G16.
#include "pk_common.cuh"
#include "pk_kernels.h"

#include <algorithm>
#include <cstdlib>

namespace pk {

namespace {

constexpr int kStepThreads = 128;  // 4 warps = 4 m16 row tiles = 64 rows per CTA
constexpr int kRowsPerCta = 64;

// ------------------------------------------------------------------------------------
// weight packing (once per layer call)
// ------------------------------------------------------------------------------------
// forward: Wp[ut][g][uu][k] = U[(g*H + 8*ut + uu)][k]      (zero padded)
__global__ void pack_fwd_kernel(const float* __restrict__ U, int NG, int H, int KPs, __half* __restrict__ Wp) {
  const int NU = (H + 7) / 8;
  const long long total = static_cast<long long>(NU) * NG * 8 * KPs;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(e % KPs);
    const long long r = e / KPs;
    const int uu = static_cast<int>(r % 8);
    const int g = static_cast<int>((r / 8) % NG);
    const int ut = static_cast<int>(r / (8 * NG));
    const int u = 8 * ut + uu;
    float v = 0.f;
    if (u < H && k < H) v = U[(static_cast<long long>(g) * H + u) * H + k];
    Wp[e] = f16_sat(v);
  }
}
// backward: UTp[ut][g][uu][j] = U[(g*H + j)][8*ut + uu]     (contraction over j for every gate g)
__global__ void pack_bwd_kernel(const float* __restrict__ U, int NG, int H, int KPs, __half* __restrict__ UTp) {
  const int NU = (H + 7) / 8;
  const long long total = static_cast<long long>(NU) * NG * 8 * KPs;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(e % KPs);
    const long long r = e / KPs;
    const int uu = static_cast<int>(r % 8);
    const int g = static_cast<int>((r / 8) % NG);
    const int ut = static_cast<int>(r / (8 * NG));
    const int u = 8 * ut + uu;
    float v = 0.f;
    if (u < H && j < H) v = U[(static_cast<long long>(g) * H + j) * H + u];
    UTp[e] = f16_sat(v);
  }
}

enum StepMode { M_LIGRU = 0, M_LSTM = 1, M_GATES = 2, M_CAND = 3 };

struct StepFwd {
  int mode, ch0;       // epilogue kind; first gate block (of PT/scale/shift) this launch handles
  int cell, act, T, B, H, ndir, k, KT, KPs, Rp;
  const __half* S16;   // [Rp][KPs] operand (h_{k-1})
  __half* S16n;        // [Rp][KPs] next operand (h_k)
  const __half* Wp;    // packed weights
  const float* PT; long long ldp;
  const float* scale; const float* shift;
  const float* mask; float mask_scalar;
  float* Hst;          // [rows][H] fp32 state h
  float* Cst;          // [rows][H] fp32 cell state (LSTM) / update gate z handed from M_GATES to M_CAND
  float* HT; __half* HT16; __half* HP16; __half* HX16;
  float* SV0; float* SV1; float* SV2; float* SV3; float* SV4;  // liGRU: z, hc ; LSTM: f, g, i, o, c
  long long ldt;
  float* Y32; long long ldy32; __half* Y16; long long ldy16;
};

__device__ __forceinline__ void smem_copy_16(void* dst, const void* src, int bytes) {
  for (int o = threadIdx.x * 16; o < bytes; o += blockDim.x * 16)
    cp_async_16(static_cast<char*>(dst) + o, static_cast<const char*>(src) + o);
}

// bring-up instrumentation: cycle sums of CTA (0,0) / thread 0 per phase (fwd 0..4, bwd 8..12), slot 15 = enable
__device__ long long g_step_clk[16];
#define STEP_CLK(slot, tprev)                                                        \
  if (PERSIST && g_step_clk[15] && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { \
    const long long tn_ = clock64();                                                 \
    g_step_clk[slot] += tn_ - tprev;                                                 \
    tprev = tn_;                                                                     \
  }

// ---- grid-wide step barrier of the persistent variants (cooperative launch: all CTAs are co-resident) ----
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void grid_arrive(unsigned* counter) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
  }
}
__device__ __forceinline__ void grid_wait(const unsigned* counter, unsigned target) {
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while (ld_acquire_gpu(counter) < target) {
      if (clock64() - t0 > 8000000000LL) __trap();  // ~4 s: a lost CTA must not hang the device
    }
  }
  __syncthreads();
}

// Contiguous global -> shared copy on the TMA engine in 8 pieces whose order starts at a CTA-dependent piece: every
// CTA reads the SAME buffer at the same time, and marching over it in lockstep hammers a few L2 slices (measured
// 2x on the per-step copy).  Called by one thread; completion = `bytes` transaction bytes on `bar`.
__device__ __forceinline__ void bulk_copy_rotated(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  const uint32_t piece = ((bytes / 8 + 15) / 16) * 16;
  const uint32_t np = (bytes + piece - 1) / piece;
  const uint32_t rot = blockIdx.x % np;
  for (uint32_t i = 0; i < np; ++i) {
    uint32_t j = i + rot;
    if (j >= np) j -= np;
    const uint32_t off = j * piece;
    const uint32_t len = min(piece, bytes - off);
    bulk_load_1d(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, len, bar);
  }
}

struct StepBars {
  uint64_t* bar;     // [2] mbarriers in shared memory (operand buffers)
  uint32_t ph[2];    // their phase parities
};

// One time step of the forward recurrence for this CTA's (unit tile, row block).  NG = gates handled in this
// phase.  PERSIST: the weight tile is already resident in Wsm and the step is fenced by the grid barrier.
template <int NG, bool PERSIST>
__device__ __forceinline__ void fwd_body(const StepFwd& a, __half* Ssm, __half* Wsm, unsigned* counter, unsigned target,
                                         StepBars& sb) {
  const int KPs = a.KPs;
  const int ut = blockIdx.x;
  const int r0 = blockIdx.y * kRowsPerCta;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int H = a.H, B = a.B, T = a.T;
  const int nrows = a.ndir * B;
  const int rows_here = min(kRowsPerCta, a.Rp - r0);

  // this thread's elements: rows (warp*16 + g, +8), units (8*ut + 2q, +1)
  int rr[2], rd[2], rb[2];
  long long col[2];
  bool rok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    rr[i] = r0 + warp * 16 + g + 8 * i;
    rok[i] = rr[i] < nrows;
    rd[i] = (rok[i] && rr[i] >= B) ? 1 : 0;
    rb[i] = rr[i] - rd[i] * B;
    col[i] = static_cast<long long>(rd[i] ? T - 1 - a.k : a.k) * B + rb[i];
  }
  const int u0 = 8 * ut + 2 * q;
  // projections of this step (issued before the wait so that L2 latency overlaps the operand copy)
  float pre[NG][2][2];  // [gate][row i][unit e]
#pragma unroll
  for (int gg = 0; gg < NG; ++gg)
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int u = u0 + e;
        pre[gg][i][e] = 0.f;
        if (rok[i] && u < H) {
          const int ch = (a.ch0 + gg) * H + u;
          pre[gg][i][e] = fmaf(__ldg(a.scale + ch), __ldg(a.PT + static_cast<long long>(ch) * a.ldp + col[i]),
                               __ldg(a.shift + ch));
        }
      }
  // thread-owned fp32 state / mask of this step: no dependence on other CTAs, so fetch before the barrier
  float hp_r[2][2], cs_r[2][2], mk_r[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int u = u0 + e;
      hp_r[i][e] = 0.f; cs_r[i][e] = 0.f; mk_r[i][e] = a.mask_scalar;
      if (rok[i] && u < H) {
        const long long sidx = static_cast<long long>(rr[i]) * H + u;
        if (a.k > 0) hp_r[i][e] = a.Hst[sidx];
        if (a.mode == M_CAND || (a.mode == M_LSTM && a.k > 0)) cs_r[i][e] = a.Cst[sidx];
        if (a.mask) mk_r[i][e] = __ldg(a.mask + sidx);
      }
    }
  long long tclk = clock64();
  if (PERSIST) grid_wait(counter, target);
  STEP_CLK(0, tclk)
  // operand copy on the TMA engine (deep pipelining; per-thread cp.async was latency-bound at ~12 B/clk per SM)
  if (threadIdx.x == 0) {
    const uint32_t sbytes = static_cast<uint32_t>(rows_here) * KPs * 2;
    const uint32_t wbytes = PERSIST ? 0u : static_cast<uint32_t>(NG) * 8 * KPs * 2;
    fence_proxy_async_all();
    mbar_arrive_expect_tx(&sb.bar[0], sbytes + wbytes);
    if (!PERSIST) bulk_load_1d(Wsm, a.Wp + static_cast<long long>(ut) * NG * 8 * KPs, wbytes, &sb.bar[0]);
    bulk_copy_rotated(Ssm, a.S16 + static_cast<long long>(r0) * KPs, sbytes, &sb.bar[0]);
  }
  mbar_wait(&sb.bar[0], sb.ph[0]);
  sb.ph[0] ^= 1u;
  float acc[NG][4];
#pragma unroll
  for (int gg = 0; gg < NG; ++gg) acc[gg][0] = acc[gg][1] = acc[gg][2] = acc[gg][3] = 0.f;
  // A (state rows) via ldmatrix.x4: matrices (rows 0-7,k 0-7) (rows 8-15,k 0-7) (rows 0-7,k 8-15) (rows 8-15,k 8-15)
  const uint32_t a_base = smem_u32(Ssm) + ((warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * KPs + 8 * (lane >> 4)) * 2;
  // B (weights, [n][k] rows) via ldmatrix.x2: matrices (n 0-7,k 0-7) (n 0-7,k 8-15)
  const uint32_t b_base = smem_u32(Wsm) + (((lane & 7)) * KPs + 8 * ((lane >> 3) & 1)) * 2;
#pragma unroll 5
  for (int kt = 0; kt < a.KT; ++kt) {
    uint32_t af[4];
    ldmatrix_x4(a_base + kt * 32, af[0], af[1], af[2], af[3]);
#pragma unroll
    for (int gg = 0; gg < NG; ++gg) {
      uint32_t b0, b1;
      ldmatrix_x2(b_base + (gg * 8 * KPs) * 2 + kt * 32, b0, b1);
      mma_m16n8k16_f16(acc[gg], af, b0, b1);
    }
  }

  STEP_CLK(1, tclk)
  // ---- epilogue: acc[gg] = {row g: units 2q,2q+1 ; row g+8: units 2q,2q+1}
#pragma unroll
  for (int i = 0; i < 2; ++i) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int u = u0 + e;
      if (u >= H) continue;
      float hnew = 0.f;
      const int r = rr[i];
      if (rok[i]) {
        const float m = mk_r[i][e];
        const long long sidx = static_cast<long long>(r) * H + u;
        const float hp = hp_r[i][e];
        const long long cidx = static_cast<long long>(rd[i] * H + u) * a.ldt + col[i];
        bool final_state = true;
        if (a.mode == M_LIGRU) {  // liGRU (reference :1133-1136): gate 0 = candidate, gate 1 = update
          const float zt = sigmoid_fast(pre[NG - 1][i][e] + acc[NG - 1][2 * i + e]);
          const float hc = act_fwd_fast(a.act, pre[0][i][e] + acc[0][2 * i + e]) * m;
          hnew = fmaf(zt, hp - hc, hc);
          if (a.SV0) a.SV0[cidx] = zt;
          if (a.SV1) a.SV1[cidx] = hc;
        } else if (a.mode == M_LSTM) {  // LSTM (reference :457-469): gates f, i, o, c
          const float ft = sigmoid_fast(pre[0][i][e] + acc[0][2 * i + e]);
          const float it = sigmoid_fast(pre[1 % NG][i][e] + acc[1 % NG][2 * i + e]);
          const float ot = sigmoid_fast(pre[2 % NG][i][e] + acc[2 % NG][2 * i + e]);
          const float gt = act_fwd_fast(a.act, pre[NG - 1][i][e] + acc[NG - 1][2 * i + e]) * m;
          const float cp = cs_r[i][e];
          const float ct = fmaf(it, gt, ft * cp);
          hnew = ot * act_fwd_fast(a.act, ct);
          a.Cst[sidx] = ct;
          if (a.SV0) a.SV0[cidx] = ft;
          if (a.SV1) a.SV1[cidx] = gt;
          if (a.SV2) a.SV2[cidx] = it;
          if (a.SV3) a.SV3[cidx] = ot;
          if (a.SV4) a.SV4[cidx] = ct;
        } else if (a.mode == M_GATES) {
          // GRU (:631-633) gates z, r / minimalGRU (:1293-1294) gate z: the next launch contracts (gate * h)
          const float zt = sigmoid_fast(pre[0][i][e] + acc[0][2 * i + e]);
          float gate = zt;
          if (NG >= 2) {
            gate = sigmoid_fast(pre[NG - 1][i][e] + acc[NG - 1][2 * i + e]);
            if (a.SV2) a.SV2[cidx] = gate;
          }
          a.Cst[sidx] = zt;
          if (a.SV0) a.SV0[cidx] = zt;
          hnew = gate * hp;  // operand of the candidate's recurrent product
          if (a.HX16) a.HX16[cidx] = f16_sat(hnew);
          final_state = false;
        } else {  // M_CAND: at = wh + Uh (gate*h); h = z h + (1 - z) act(at) mask   (:634-636, :1295-1297)
          const float zt = cs_r[i][e];
          const float hc = act_fwd_fast(a.act, pre[0][i][e] + acc[0][2 * i + e]) * m;
          hnew = fmaf(zt, hp - hc, hc);
          if (a.SV1) a.SV1[cidx] = hc;
        }
        if (final_state) {
          a.Hst[sidx] = hnew;
          if (a.HT) a.HT[cidx] = hnew;
          if (a.HT16) a.HT16[cidx] = f16_sat(hnew);
          if (a.HP16) a.HP16[cidx] = f16_sat(hp);
          if (a.Y32) a.Y32[col[i] * a.ldy32 + rd[i] * H + u] = hnew;
          if (a.Y16) a.Y16[col[i] * a.ldy16 + rd[i] * H + u] = f16_sat(hnew);
        }
      }
      if (r < a.Rp) a.S16n[static_cast<long long>(r) * KPs + u] = f16_sat(hnew);
    }
  }
  STEP_CLK(2, tclk)
  if (PERSIST) grid_arrive(counter);
  STEP_CLK(3, tclk)
}

// per-step launch: grid (unit tiles, row blocks of 64)
template <int NG>
__global__ void __launch_bounds__(kStepThreads) cell_fwd_step_kernel(const StepFwd a) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2];
  __half* Ssm = reinterpret_cast<__half*>(smem);                          // [64][KPs]
  __half* Wsm = Ssm + static_cast<size_t>(kRowsPerCta) * a.KPs;           // [NG*8][KPs]
  StepBars sb{bars, {0u, 0u}};
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
  __syncthreads();
  fwd_body<NG, false>(a, Ssm, Wsm, nullptr, 0u, sb);
}

// persistent launch (cooperative): weights stay in shared memory for all T steps, the fp16 state is exchanged
// through global memory (L2) behind a grid barrier.  TWO: GRU / minimalGRU (gates phase with NG blocks, then the
// candidate phase with one block).
struct PersistFwd {
  __half* S0; __half* S1;          // single-phase: double buffer; two-phase: S0 = h, S1 = gate*h
  const __half* Wmain; const __half* Wcand;
  unsigned* counter;
};
template <int NG, bool TWO>
__global__ void __launch_bounds__(kStepThreads) cell_fwd_persist_kernel(const StepFwd base, const PersistFwd e) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int KPs = base.KPs;
  __half* Ssm = reinterpret_cast<__half*>(smem);
  __half* Wsm = Ssm + static_cast<size_t>(kRowsPerCta) * KPs;             // [NG*8][KPs]
  __half* Wcs = Wsm + static_cast<size_t>(NG) * 8 * KPs;                  // [8][KPs] (TWO only)
  const unsigned ncta = gridDim.x * gridDim.y;
  __shared__ __align__(8) uint64_t bars[2];
  StepBars sb{bars, {0u, 0u}};
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
  smem_copy_16(Wsm, e.Wmain + static_cast<long long>(blockIdx.x) * NG * 8 * KPs, NG * 8 * KPs * 2);
  if (TWO) smem_copy_16(Wcs, e.Wcand + static_cast<long long>(blockIdx.x) * 8 * KPs, 8 * KPs * 2);
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  StepFwd p = base;
  for (int k = 0; k < base.T; ++k) {
    p.k = k;
    if (!TWO) {
      p.S16 = (k & 1) ? e.S1 : e.S0;
      p.S16n = (k & 1) ? e.S0 : e.S1;
      fwd_body<NG, true>(p, Ssm, Wsm, e.counter, static_cast<unsigned>(k) * ncta, sb);
    } else {
      p.mode = M_GATES; p.ch0 = 1; p.S16 = e.S0; p.S16n = e.S1;
      fwd_body<NG, true>(p, Ssm, Wsm, e.counter, static_cast<unsigned>(2 * k) * ncta, sb);
      p.mode = M_CAND; p.ch0 = 0; p.S16 = e.S1; p.S16n = e.S0;
      fwd_body<1, true>(p, Ssm, Wcs, e.counter, static_cast<unsigned>(2 * k + 1) * ncta, sb);
    }
  }
}

struct StepBwd {
  int mode, g0, NGT;   // epilogue kind; first gate block this launch writes; gate blocks in GT16 per direction
  int act, T, B, H, ndir, k, KT, KPs, Rp, first;
  int gbufs;           // chunk buffers in shared memory (2 = double buffered)
  const __half* G16;   // [NGC][Rp][KPs] scaled gradients contracted by this launch (NGC = template chunks)
  __half* G16n;        // operand written by this launch: [n_out][Rp][KPs]
  int n_out;
  const __half* UTp;   // packed transposed weights of the contracted gates
  const float* dYT;    // [ndir*H][ldt]
  const float* HT; const float* SV0; const float* SV1; const float* SV2; const float* SV3; const float* SV4;
  long long ldt;
  const float* mask; float mask_scalar;
  const float* gscale;
  float* Kh;           // [rows][H] fp32: thread-local part of the h carry
  float* Kc;           // [rows][H] fp32: LSTM dc carry / GRU: dh handed from phase A to phase B
  __half* GT16;        // [ndir][NGT*H][ldt]
};

// One reverse-time step.  NGC = gate chunks contracted in the GEMM part; the NGC weight tiles live in Wsm
// ([NGC*8][KPs]; loaded here unless PERSIST).
template <int NGC, bool PERSIST>
__device__ __forceinline__ void bwd_body(const StepBwd& a, __half* Gsm, __half* Wsm, unsigned* counter, unsigned target,
                                         StepBars& sb) {
  const int KPs = a.KPs;
  const int ut = blockIdx.x;
  const int r0 = blockIdx.y * kRowsPerCta;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int H = a.H, B = a.B, T = a.T;
  const int nrows = a.ndir * B;
  const int rows_here = min(kRowsPerCta, a.Rp - r0);
  const float s = a.gscale ? __ldg(a.gscale) : 1.f;
  const float inv_s = 1.f / s;

  // everything the pointwise part needs that does not depend on other CTAs (saved tensors, dY, own carries, mask)
  // is fetched before the barrier so that its DRAM / L2 latency hides behind the wait and the GEMM
  float pf[2][2][10];
  {
    const int u0p = 8 * ut + 2 * q;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = r0 + warp * 16 + g + 8 * i;
      const bool rok = r < nrows;
      const int d = (rok && r >= B) ? 1 : 0;
      const long long col = static_cast<long long>(d ? T - 1 - a.k : a.k) * B + (r - d * B);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int u = u0p + e;
#pragma unroll
        for (int j = 0; j < 10; ++j) pf[i][e][j] = 0.f;
        pf[i][e][7] = a.mask_scalar;
        if (rok && u < H) {
          const long long sidx = static_cast<long long>(r) * H + u;
          const long long cidx = static_cast<long long>(d * H + u) * a.ldt + col;
          const long long pidx = cidx + (d ? B : -B);
          if (a.mode != M_CAND) pf[i][e][0] = __ldg(a.dYT + cidx);
          pf[i][e][1] = __ldg(a.SV0 + cidx);
          pf[i][e][2] = __ldg(a.SV1 + cidx);
          if (a.mode == M_LSTM || (a.mode == M_CAND && a.n_out == 2)) pf[i][e][3] = __ldg(a.SV2 + cidx);
          if (a.mode == M_LSTM) {
            pf[i][e][4] = __ldg(a.SV3 + cidx);
            pf[i][e][5] = __ldg(a.SV4 + cidx);
            if (a.k > 0) pf[i][e][6] = __ldg(a.SV4 + pidx);
          } else if (a.mode != M_GATES) {
            if (a.k > 0) pf[i][e][6] = __ldg(a.HT + pidx);
          }
          if (a.mask) pf[i][e][7] = __ldg(a.mask + sidx);
          if (!a.first && (a.mode == M_LIGRU || a.mode == M_GATES)) pf[i][e][8] = a.Kh[sidx];
          if ((a.mode == M_LSTM && !a.first) || a.mode == M_CAND) pf[i][e][9] = a.Kc[sidx];
        }
      }
    }
  }
  long long tclk = clock64();
  if (PERSIST) grid_wait(counter, target);
  STEP_CLK(8, tclk)
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (!a.first || a.mode == M_CAND) {
    // sum_g G_g . U_g : contraction chunked per gate through shared memory; with two chunk buffers (a.gbufs == 2)
    // the copy of chunk g+1 overlaps the MMAs of chunk g
    float acc1[4] = {0.f, 0.f, 0.f, 0.f};
    const uint32_t a_off = ((warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * KPs + 8 * (lane >> 4)) * 2;
    const uint32_t b_base = smem_u32(Wsm) + ((lane & 7) * KPs + 8 * ((lane >> 3) & 1)) * 2;
    const size_t gbuf_elems = static_cast<size_t>(kRowsPerCta) * KPs;
    // gate chunk gg = rows r0.. of G16[gg][Rp][KPs]: one contiguous block -> TMA bulk copy (issued by thread 0)
    auto issue = [&](int gg) {
      if (threadIdx.x == 0) {
        const int buf = (a.gbufs == 2) ? (gg & 1) : 0;
        const uint32_t gbytes = static_cast<uint32_t>(rows_here) * KPs * 2;
        const uint32_t wbytes = (!PERSIST && gg == 0) ? static_cast<uint32_t>(NGC) * 8 * KPs * 2 : 0u;
        fence_proxy_async_all();
        mbar_arrive_expect_tx(&sb.bar[buf], gbytes + wbytes);
        if (wbytes) bulk_load_1d(Wsm, a.UTp + static_cast<long long>(ut) * NGC * 8 * KPs, wbytes, &sb.bar[buf]);
        bulk_copy_rotated(Gsm + buf * gbuf_elems, a.G16 + (static_cast<long long>(gg) * a.Rp + r0) * KPs, gbytes, &sb.bar[buf]);
      }
    };
    issue(0);
    for (int gg = 0; gg < NGC; ++gg) {
      const int buf = (a.gbufs == 2) ? (gg & 1) : 0;
      if (a.gbufs == 2 && gg + 1 < NGC) issue(gg + 1);
      mbar_wait(&sb.bar[buf], sb.ph[buf]);
      sb.ph[buf] ^= 1u;
      const uint32_t a_base = smem_u32(Gsm + buf * gbuf_elems) + a_off;
#pragma unroll 4
      for (int kt = 0; kt < a.KT; kt += 2) {
        uint32_t af[4], b0, b1;
        ldmatrix_x4(a_base + kt * 32, af[0], af[1], af[2], af[3]);
        ldmatrix_x2(b_base + (gg * 8 * KPs) * 2 + kt * 32, b0, b1);
        mma_m16n8k16_f16(acc, af, b0, b1);
        if (kt + 1 < a.KT) {
          ldmatrix_x4(a_base + (kt + 1) * 32, af[0], af[1], af[2], af[3]);
          ldmatrix_x2(b_base + (gg * 8 * KPs) * 2 + (kt + 1) * 32, b0, b1);
          mma_m16n8k16_f16(acc1, af, b0, b1);
        }
      }
      __syncthreads();
      if (a.gbufs != 2 && gg + 1 < NGC) issue(gg + 1);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] += acc1[i];
  }

  STEP_CLK(9, tclk)
  const int u0 = 8 * ut + 2 * q;
  const long long gate_stride = static_cast<long long>(H) * a.ldt;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = r0 + warp * 16 + g + 8 * i;
    const bool rok = r < nrows;
    const int d = (rok && r >= B) ? 1 : 0;
    const int b = r - d * B;
    const long long col = static_cast<long long>(d ? T - 1 - a.k : a.k) * B + b;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int u = u0 + e;
      if (u >= H || r >= a.Rp) continue;
      float gq[4] = {0.f, 0.f, 0.f, 0.f};
      if (rok) {
        const long long sidx = static_cast<long long>(r) * H + u;
        const long long cidx = static_cast<long long>(d * H + u) * a.ldt + col;
        const float* P = pf[i][e];
        const float m = P[7];
        const float rm = (m != 0.f) ? rcp_approx(m) : 0.f;
        const float carry = acc[2 * i + e] * inv_s;
        if (a.mode == M_LIGRU) {
          float dh = P[0];
          if (!a.first) dh += P[8] + carry;
          const float z = P[1], hc = P[2], hp = P[6];
          gq[0] = dh * (1.f - z) * m * act_bwd_from_out(a.act, hc * rm);
          gq[1] = dh * (hp - hc) * z * (1.f - z);
          a.Kh[sidx] = dh * z;
        } else if (a.mode == M_LSTM) {
          float dh = P[0];
          if (!a.first) dh += carry;
          const float f = P[1], gt = P[2], it = P[3], o = P[4], c = P[5], cp = P[6];
          const float ac = act_fwd_fast(a.act, c);
          float dc = dh * o * act_bwd_from_out(a.act, ac);
          if (!a.first) dc += P[9];
          gq[0] = dc * cp * f * (1.f - f);                        // forget gate
          gq[1] = dc * gt * it * (1.f - it);                      // input gate
          gq[2] = dh * ac * o * (1.f - o);                        // output gate
          gq[3] = dc * it * m * act_bwd_from_out(a.act, gt * rm); // candidate
          a.Kc[sidx] = dc * f;
        } else if (a.mode == M_GATES) {
          // phase A of GRU / minimalGRU: dh complete -> candidate pre-activation gradient (contracted by phase B)
          float dh = P[0];
          if (!a.first) dh += P[8] + carry;
          const float z = P[1], hc = P[2];
          gq[0] = dh * (1.f - z) * m * act_bwd_from_out(a.act, hc * rm);
          a.Kc[sidx] = dh;
        } else {
          // phase B: v = da . Uh is the gradient w.r.t. (gate * h_prev)
          const float v = carry;
          const float dh = P[9];
          const float z = P[1], hc = P[2], hp = P[6];
          if (a.n_out == 2) {  // GRU: gates z, r
            const float rt = P[3];
            gq[0] = dh * (hp - hc) * z * (1.f - z);
            gq[1] = v * hp * rt * (1.f - rt);
            a.Kh[sidx] = fmaf(dh, z, v * rt);
          } else {             // minimalGRU: z gates the recurrent operand as well
            gq[0] = (dh * (hp - hc) + v * hp) * z * (1.f - z);
            a.Kh[sidx] = (dh + v) * z;
          }
        }
#pragma unroll
        for (int gg = 0; gg < 4; ++gg)
          if (gg < a.n_out)
            a.GT16[(static_cast<long long>(d) * a.NGT + a.g0 + gg) * gate_stride + static_cast<long long>(u) * a.ldt + col] =
                f16_sat(gq[gg] * s);
      }
#pragma unroll
      for (int gg = 0; gg < 4; ++gg)
        if (gg < a.n_out) a.G16n[(static_cast<long long>(gg) * a.Rp + r) * KPs + u] = f16_sat(gq[gg] * s);
    }
  }
  STEP_CLK(10, tclk)
  if (PERSIST) grid_arrive(counter);
  STEP_CLK(11, tclk)
}

template <int NGC>
__global__ void __launch_bounds__(kStepThreads) cell_bwd_step_kernel(const StepBwd a) {
  extern __shared__ __align__(16) uint8_t smem[];
  __half* Gsm = reinterpret_cast<__half*>(smem);                          // [gbufs][64][KPs] gate chunks
  __half* Wsm = Gsm + static_cast<size_t>(a.gbufs) * kRowsPerCta * a.KPs; // [NGC*8][KPs]
  __shared__ __align__(8) uint64_t bars[2];
  StepBars sb{bars, {0u, 0u}};
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
  __syncthreads();
  bwd_body<NGC, false>(a, Gsm, Wsm, nullptr, 0u, sb);
}

struct PersistBwd {
  __half* G0; __half* G1;          // single-phase: double buffer of [Rp][NG][KPs]; two-phase: G0 = (dpz[,dpr]), G1 = da
  const __half* Wmain; const __half* Wcand;
  unsigned* counter;
  int n_main;                      // gate blocks written by the main phase (single-phase: NG; two-phase: NG-1)
};
template <int NGC, bool TWO>
__global__ void __launch_bounds__(kStepThreads) cell_bwd_persist_kernel(const StepBwd base, const PersistBwd e) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int KPs = base.KPs;
  __half* Gsm = reinterpret_cast<__half*>(smem);
  __half* Wsm = Gsm + static_cast<size_t>(base.gbufs) * kRowsPerCta * KPs;  // [NGC*8][KPs]
  __half* Wcs = Wsm + static_cast<size_t>(NGC) * 8 * KPs;                   // [8][KPs] (TWO only)
  const unsigned ncta = gridDim.x * gridDim.y;
  __shared__ __align__(8) uint64_t bars[2];
  StepBars sb{bars, {0u, 0u}};
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
  smem_copy_16(Wsm, e.Wmain + static_cast<long long>(blockIdx.x) * NGC * 8 * KPs, NGC * 8 * KPs * 2);
  if (TWO) smem_copy_16(Wcs, e.Wcand + static_cast<long long>(blockIdx.x) * 8 * KPs, 8 * KPs * 2);
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  StepBwd p = base;
  for (int k = base.T - 1; k >= 0; --k) {
    const int it = base.T - 1 - k;
    p.k = k; p.first = (it == 0) ? 1 : 0;
    if (!TWO) {
      p.G16 = ((it + 1) & 1) ? e.G1 : e.G0;
      p.G16n = (it & 1) ? e.G1 : e.G0;
      bwd_body<NGC, true>(p, Gsm, Wsm, e.counter, static_cast<unsigned>(it) * ncta, sb);
    } else {
      p.mode = M_GATES; p.g0 = 0; p.n_out = 1; p.G16 = e.G0; p.G16n = e.G1;
      bwd_body<NGC, true>(p, Gsm, Wsm, e.counter, static_cast<unsigned>(2 * it) * ncta, sb);
      p.mode = M_CAND; p.g0 = 1; p.n_out = e.n_main; p.G16 = e.G1; p.G16n = e.G0;
      bwd_body<1, true>(p, Gsm, Wcs, e.counter, static_cast<unsigned>(2 * it + 1) * ncta, sb);
    }
  }
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline int cell_gates(int cell) { return cell == CELL_LSTM ? 4 : (cell == CELL_GRU ? 3 : 2); }
inline bool two_phase(int cell) { return cell == CELL_GRU || cell == CELL_MGRU; }

struct Layout {
  int NG, KT, KPs, Rp, NU;
  size_t off_wp, off_a, off_b, off_h, off_c, off_bar, total;
};
// fwd: a = S_h (fp16 state operand, double buffered for single-phase cells), b = second buffer / S_x
// bwd: a = G16 (double buffered for single-phase cells: 2 x NG gates), b = phase-A operand (two-phase cells)
Layout make_layout(int cell, int B, int H, int ndir, bool bwd) {
  Layout L;
  L.NG = cell_gates(cell);
  L.KT = (H + 15) / 16;
  L.KPs = 16 * L.KT + 8;
  L.Rp = round_up(ndir * B, 16);
  L.NU = (H + 7) / 8;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) / 256 * 256; return r; };
  size_t wp_bytes = static_cast<size_t>(L.NU) * L.NG * 8 * L.KPs * 2;
  if (cell == CELL_LSTM)  // the cluster-persistent kernels pack one image per CTA (pk_cell_cluster.cu)
    wp_bytes = std::max(wp_bytes, static_cast<size_t>(lstm_cluster_pack_bytes(H)));
  if (cell == CELL_GRU || cell == CELL_MGRU)  // pk_cell_cluster2.cu
    wp_bytes = std::max(wp_bytes, static_cast<size_t>(gru_cluster_pack_bytes(cell, H)));
  L.off_wp = take(wp_bytes);
  L.off_h = take(static_cast<size_t>(ndir) * B * H * 4);
  L.off_c = take(static_cast<size_t>(ndir) * B * H * 4);
  const size_t op = static_cast<size_t>(L.Rp) * L.KPs * 2;
  if (!bwd) {
    L.off_a = take(op);
    L.off_b = take(op);
  } else {
    L.off_a = take(op * L.NG * 2);
    L.off_b = take(op);
  }
  L.off_bar = take(256);
  L.total = o;
  return L;
}

template <typename K>
int set_smem(K kern, size_t smem) {
  PK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  return 0;
}

// can `kern` run as ONE co-resident wave of `ctas` CTAs (needed by the grid barrier)?
template <typename K>
bool fits_one_wave(K kern, size_t smem, int ctas) {
  static int sms = 0, coop = -1;
  if (coop < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  }
  if (!coop) return false;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kStepThreads, smem) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return per_sm * sms >= ctas;
}

bool persist_enabled() {
  const char* e = getenv("PK_STEP_PERSIST");
  return !(e && e[0] == '0');
}

}  // namespace

// bring-up: enable / read-and-reset the phase clocks (not part of the C ABI)
extern "C" int pk_debug_step_clocks(int enable, long long* out16) {
  long long h[16];
  if (cudaMemcpyFromSymbol(h, g_step_clk, sizeof(h)) != cudaSuccess) return 1;
  if (out16) for (int i = 0; i < 16; ++i) out16[i] = h[i];
  for (int i = 0; i < 16; ++i) h[i] = 0;
  h[15] = enable;
  return cudaMemcpyToSymbol(g_step_clk, h, sizeof(h)) != cudaSuccess;
}

long long cell_step_workspace_bytes(int cell, int T, int B, int H, int ndir, int backward) {
  (void)T;
  return static_cast<long long>(make_layout(cell, B, H, ndir, backward != 0).total);
}

int cell_step_launches(int cell, int T, int B, int H, int ndir, int backward) {
  if (lstm_cluster_usable(cell, H) || gru_cluster_usable(cell, H)) return 2;  // pack + one cluster-persistent kernel
  const Layout L = make_layout(cell, B, H, ndir, backward != 0);
  const bool tp = two_phase(cell);
  const int packs = tp ? 2 : 1;
  if (persist_enabled()) {  // the cooperative variant is used whenever the grid fits one wave (148 SMs x >= 1 CTA)
    const int ctas = L.NU * ((L.Rp + kRowsPerCta - 1) / kRowsPerCta);
    if (ctas <= 148) return packs + 1;
  }
  return packs + (tp ? 2 : 1) * T;
}

static bool cell_supported(int cell) {
  return cell == CELL_LIGRU || cell == CELL_LSTM || cell == CELL_GRU || cell == CELL_MGRU;
}

int cell_step_fwd(const CellStepFwdArgs& a, cudaStream_t stream) {
  PK_REQUIRE(cell_supported(a.cell), "cell_step_fwd: cell %d not implemented", a.cell);
  const Layout L = make_layout(a.cell, a.B, a.H, a.ndir, false);
  PK_REQUIRE(a.workspace && a.workspace_bytes >= static_cast<long long>(L.total),
             "cell_step_fwd: workspace too small (%lld < %zu)", a.workspace_bytes, L.total);
  char* ws = static_cast<char*>(a.workspace);
  __half* Wp = reinterpret_cast<__half*>(ws + L.off_wp);
  // cluster-persistent families first; -1 = "cannot launch here" -> the step-wise kernels below (same interface)
  if (lstm_cluster_usable(a.cell, a.H)) { const int rc = lstm_cluster_fwd(a, Wp, stream); if (rc != -1) return rc; }
  if (gru_cluster_usable(a.cell, a.H)) { const int rc = gru_cluster_fwd(a, Wp, stream); if (rc != -1) return rc; }
  __half* S[2] = {reinterpret_cast<__half*>(ws + L.off_a), reinterpret_cast<__half*>(ws + L.off_b)};
  unsigned* counter = reinterpret_cast<unsigned*>(ws + L.off_bar);
  const size_t op = static_cast<size_t>(L.Rp) * L.KPs * 2;
  PK_CHECK_CUDA(cudaMemsetAsync(S[0], 0, op, stream));
  PK_CHECK_CUDA(cudaMemsetAsync(S[1], 0, op, stream));
  PK_CHECK_CUDA(cudaMemsetAsync(counter, 0, 256, stream));
  // gate blocks are packed in the caller's order, so a two-phase cell finds its candidate block (gate 0) first
  // and the blocks contracted against h (gates 1..) right behind it
  const bool tp = two_phase(a.cell);
  __half* Wgates = Wp + static_cast<size_t>(L.NU) * 8 * L.KPs;
  if (!tp) {
    pack_fwd_kernel<<<296, 256, 0, stream>>>(a.U, L.NG, a.H, L.KPs, Wp);
  } else {
    pack_fwd_kernel<<<296, 256, 0, stream>>>(a.U, 1, a.H, L.KPs, Wp);
    pack_fwd_kernel<<<296, 256, 0, stream>>>(a.U + static_cast<size_t>(a.H) * a.H, L.NG - 1, a.H, L.KPs, Wgates);
  }
  const int nmax = tp ? L.NG - 1 : L.NG;
  StepFwd p;
  p.cell = a.cell; p.act = a.act; p.T = a.T; p.B = a.B; p.H = a.H; p.ndir = a.ndir;
  p.KT = L.KT; p.KPs = L.KPs; p.Rp = L.Rp;
  p.PT = a.PT; p.ldp = a.ldp; p.scale = a.scale; p.shift = a.shift; p.mask = a.mask; p.mask_scalar = a.mask_scalar;
  p.Hst = reinterpret_cast<float*>(ws + L.off_h); p.Cst = reinterpret_cast<float*>(ws + L.off_c);
  p.HT = a.HT; p.HT16 = a.HT16; p.HP16 = a.HP16; p.HX16 = a.HX16;
  p.SV0 = a.SV[0]; p.SV1 = a.SV[1]; p.SV2 = a.SV[2]; p.SV3 = a.SV[3]; p.SV4 = a.SV[4];
  p.ldt = a.ldt; p.Y32 = a.Y32; p.ldy32 = a.ldy32; p.Y16 = a.Y16; p.ldy16 = a.ldy16;
  p.mode = (a.cell == CELL_LSTM) ? M_LSTM : M_LIGRU; p.ch0 = 0; p.Wp = Wp; p.k = 0; p.S16 = S[0]; p.S16n = S[1];
  const dim3 grid(L.NU, (L.Rp + kRowsPerCta - 1) / kRowsPerCta);
  const int ctas = grid.x * grid.y;

  // ---- persistent variant: one cooperative launch, weights stationary in shared memory ----
  if (persist_enabled()) {
    const size_t smem_p = (static_cast<size_t>(kRowsPerCta) + (nmax + (tp ? 1 : 0)) * 8) * L.KPs * 2;
    const void* kern = nullptr;
    bool ok = false;
    if (!tp && nmax == 2) { kern = reinterpret_cast<const void*>(cell_fwd_persist_kernel<2, false>); ok = fits_one_wave(cell_fwd_persist_kernel<2, false>, smem_p, ctas); }
    if (!tp && nmax == 4) { kern = reinterpret_cast<const void*>(cell_fwd_persist_kernel<4, false>); ok = fits_one_wave(cell_fwd_persist_kernel<4, false>, smem_p, ctas); }
    if (tp && nmax == 2) { kern = reinterpret_cast<const void*>(cell_fwd_persist_kernel<2, true>); ok = fits_one_wave(cell_fwd_persist_kernel<2, true>, smem_p, ctas); }
    if (tp && nmax == 1) { kern = reinterpret_cast<const void*>(cell_fwd_persist_kernel<1, true>); ok = fits_one_wave(cell_fwd_persist_kernel<1, true>, smem_p, ctas); }
    if (ok) {
      PersistFwd e;
      e.S0 = S[0]; e.S1 = S[1]; e.Wmain = tp ? Wgates : Wp; e.Wcand = Wp; e.counter = counter;
      void* args[] = {&p, &e};
      PK_CHECK_CUDA(cudaLaunchCooperativeKernel(kern, grid, dim3(kStepThreads), args, smem_p, stream));
      return 0;
    }
  }

  // ---- fallback: one launch per time step (any H) ----
  const size_t smem_main = (static_cast<size_t>(kRowsPerCta) + nmax * 8) * L.KPs * 2;
  const size_t smem_cand = (static_cast<size_t>(kRowsPerCta) + 8) * L.KPs * 2;
  auto kmain = (nmax == 1) ? cell_fwd_step_kernel<1> : (nmax == 2) ? cell_fwd_step_kernel<2> : cell_fwd_step_kernel<4>;
  if (set_smem(kmain, smem_main)) return 1;
  if (tp && set_smem(cell_fwd_step_kernel<1>, std::max(smem_cand, nmax == 1 ? smem_main : smem_cand))) return 1;
  for (int k = 0; k < a.T; ++k) {
    p.k = k;
    if (!tp) {
      p.S16 = S[k & 1]; p.S16n = S[(k + 1) & 1];
      kmain<<<grid, kStepThreads, smem_main, stream>>>(p);
    } else {
      // S[0] = fp16 state h, S[1] = fp16 (gate * h)
      p.mode = M_GATES; p.ch0 = 1; p.Wp = Wgates; p.S16 = S[0]; p.S16n = S[1];
      kmain<<<grid, kStepThreads, smem_main, stream>>>(p);
      p.mode = M_CAND; p.ch0 = 0; p.Wp = Wp; p.S16 = S[1]; p.S16n = S[0];
      cell_fwd_step_kernel<1><<<grid, kStepThreads, smem_cand, stream>>>(p);
    }
  }
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int cell_step_bwd(const CellStepBwdArgs& a, cudaStream_t stream) {
  PK_REQUIRE(cell_supported(a.cell), "cell_step_bwd: cell %d not implemented", a.cell);
  const Layout L = make_layout(a.cell, a.B, a.H, a.ndir, true);
  PK_REQUIRE(a.workspace && a.workspace_bytes >= static_cast<long long>(L.total),
             "cell_step_bwd: workspace too small (%lld < %zu)", a.workspace_bytes, L.total);
  char* ws = static_cast<char*>(a.workspace);
  __half* UTp = reinterpret_cast<__half*>(ws + L.off_wp);
  if (lstm_cluster_usable(a.cell, a.H)) { const int rc = lstm_cluster_bwd(a, UTp, stream); if (rc != -1) return rc; }
  if (gru_cluster_usable(a.cell, a.H)) { const int rc = gru_cluster_bwd(a, UTp, stream); if (rc != -1) return rc; }
  const size_t op = static_cast<size_t>(L.Rp) * L.KPs * 2;
  __half* Ga = reinterpret_cast<__half*>(ws + L.off_a);
  __half* Gb = reinterpret_cast<__half*>(ws + L.off_b);
  unsigned* counter = reinterpret_cast<unsigned*>(ws + L.off_bar);
  PK_CHECK_CUDA(cudaMemsetAsync(Ga, 0, op * L.NG * 2, stream));
  PK_CHECK_CUDA(cudaMemsetAsync(Gb, 0, op, stream));
  PK_CHECK_CUDA(cudaMemsetAsync(counter, 0, 256, stream));
  const bool tp = two_phase(a.cell);
  __half* UTgates = UTp + static_cast<size_t>(L.NU) * 8 * L.KPs;
  if (!tp) {
    pack_bwd_kernel<<<296, 256, 0, stream>>>(a.U, L.NG, a.H, L.KPs, UTp);
  } else {
    pack_bwd_kernel<<<296, 256, 0, stream>>>(a.U, 1, a.H, L.KPs, UTp);
    pack_bwd_kernel<<<296, 256, 0, stream>>>(a.U + static_cast<size_t>(a.H) * a.H, L.NG - 1, a.H, L.KPs, UTgates);
  }
  const int nmain = tp ? L.NG - 1 : L.NG;
  StepBwd p;
  p.act = a.act; p.T = a.T; p.B = a.B; p.H = a.H; p.ndir = a.ndir; p.NGT = L.NG;
  p.KT = L.KT; p.KPs = L.KPs; p.Rp = L.Rp;
  p.dYT = a.dYT; p.HT = a.HT;
  p.SV0 = a.SV[0]; p.SV1 = a.SV[1]; p.SV2 = a.SV[2]; p.SV3 = a.SV[3]; p.SV4 = a.SV[4];
  p.ldt = a.ldt; p.mask = a.mask; p.mask_scalar = a.mask_scalar; p.gscale = a.gscale;
  p.Kh = reinterpret_cast<float*>(ws + L.off_h); p.Kc = reinterpret_cast<float*>(ws + L.off_c);
  p.GT16 = a.GT16;
  p.mode = (a.cell == CELL_LSTM) ? M_LSTM : M_LIGRU; p.g0 = 0; p.n_out = L.NG; p.UTp = UTp; p.k = a.T - 1; p.first = 1;
  __half* G[2] = {Ga, Ga + static_cast<size_t>(L.Rp) * L.NG * L.KPs};
  p.G16 = G[1]; p.G16n = G[0];
  const dim3 grid(L.NU, (L.Rp + kRowsPerCta - 1) / kRowsPerCta);
  const int ctas = grid.x * grid.y;

  // double-buffer the gate chunks when more than one is contracted and shared memory allows it
  auto smem_for = [&](int gbufs, int wtiles) { return (static_cast<size_t>(gbufs) * kRowsPerCta + wtiles * 8) * L.KPs * 2; };
  if (persist_enabled()) {
    p.gbufs = (nmain > 1 && smem_for(2, nmain + (tp ? 1 : 0)) <= 200 * 1024) ? 2 : 1;
    const size_t smem_p = smem_for(p.gbufs, nmain + (tp ? 1 : 0));
    const void* kern = nullptr;
    bool ok = false;
    if (!tp && nmain == 2) { kern = reinterpret_cast<const void*>(cell_bwd_persist_kernel<2, false>); ok = fits_one_wave(cell_bwd_persist_kernel<2, false>, smem_p, ctas); }
    if (!tp && nmain == 4) { kern = reinterpret_cast<const void*>(cell_bwd_persist_kernel<4, false>); ok = fits_one_wave(cell_bwd_persist_kernel<4, false>, smem_p, ctas); }
    if (tp && nmain == 2) { kern = reinterpret_cast<const void*>(cell_bwd_persist_kernel<2, true>); ok = fits_one_wave(cell_bwd_persist_kernel<2, true>, smem_p, ctas); }
    if (tp && nmain == 1) { kern = reinterpret_cast<const void*>(cell_bwd_persist_kernel<1, true>); ok = fits_one_wave(cell_bwd_persist_kernel<1, true>, smem_p, ctas); }
    if (ok) {
      PersistBwd e;
      e.G0 = tp ? Ga : G[0]; e.G1 = tp ? Gb : G[1]; e.Wmain = tp ? UTgates : UTp; e.Wcand = UTp; e.counter = counter;
      e.n_main = nmain;
      void* args[] = {&p, &e};
      PK_CHECK_CUDA(cudaLaunchCooperativeKernel(kern, grid, dim3(kStepThreads), args, smem_p, stream));
      return 0;
    }
  }

  p.gbufs = (nmain > 1 && smem_for(2, nmain) <= 200 * 1024) ? 2 : 1;
  const size_t smem = smem_for(p.gbufs, nmain);
  auto kmain = (nmain == 1) ? cell_bwd_step_kernel<1> : (nmain == 2) ? cell_bwd_step_kernel<2> : cell_bwd_step_kernel<4>;
  if (set_smem(kmain, smem)) return 1;
  if (tp && set_smem(cell_bwd_step_kernel<1>, smem)) return 1;
  for (int k = a.T - 1; k >= 0; --k) {
    const int it = a.T - 1 - k;
    p.k = k; p.first = (it == 0) ? 1 : 0;
    if (!tp) {
      p.G16 = G[(it + 1) & 1];  // written by the previous iteration
      p.G16n = G[it & 1];
      kmain<<<grid, kStepThreads, smem, stream>>>(p);
    } else {
      // Ga = (dpre_z[, dpre_r]) of the previously processed step, Gb = dpre_h of this step
      p.mode = M_GATES; p.g0 = 0; p.n_out = 1; p.UTp = UTgates; p.G16 = Ga; p.G16n = Gb;
      kmain<<<grid, kStepThreads, smem, stream>>>(p);
      p.mode = M_CAND; p.g0 = 1; p.n_out = L.NG - 1; p.UTp = UTp; p.G16 = Gb; p.G16n = Ga;
      cell_bwd_step_kernel<1><<<grid, kStepThreads, smem, stream>>>(p);
    }
  }
  PK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pk
