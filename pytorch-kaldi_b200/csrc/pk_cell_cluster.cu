// pk_cell_cluster.cu — cluster-persistent LSTM recurrence (reference neural_networks.py:300-483, time loop :447-475).
//
// Replaces the step-wise LSTM path of pk_cell_step.cu (fp16 state exchanged through L2 behind a grid barrier,
// every CTA re-reading the whole [rows x K] operand each step: 10.5 / 15 us per step forward / backward at H = 550)
// for H <= 560 with the formulation of the persistent liGRU kernels, sized for four gates:
//
//   * one thread-block CLUSTER owns 16 batch rows (one m16 tile of mma.sync) for all T steps; clusters are
//     independent (2B / 16 of them);
//   * CTA c of the cluster owns the hidden units [UPC*c, UPC*c + UPC), UPC = 8 * MT, for ALL four gates; its slice
//     W_c[g][ul][k] = U_g[UPC*c + ul][k] of the recurrent weights (fp16, 178 KB at H = 550) is copied to shared
//     memory once and stays there — the SAME image serves the forward and the reverse-time kernel;
//   * forward: pre[16 rows, 4 gates x UPC] = h_{t-1}[16, K] . W_c^T on mma.sync (A = state rows via ldmatrix, B =
//     weight rows via ldmatrix), gates / cell update on fp32 state in registers, the new fp16 state slice is pushed
//     into every CTA's next-state buffer over distributed shared memory (ALL-GATHER, 16-byte rows);
//   * backward: the contraction dh_{t-1}[r, u] = sum_g sum_j dpre_g[r, j] U_g[j, u] is split along (g, j) = K
//     (K-SPLIT): a CTA multiplies the gate gradients it has just produced itself (its own units, never exchanged)
//     with its weight slice read TRANSPOSED (ldmatrix.trans on the same image) into partial sums for all H units
//     and scatters them to the owners' receive buffers (REDUCE-SCATTER; fp16 carrying the loss scale, like the
//     operands and like the K-split liGRU kernel, double buffered); the owner adds the CL partials in fp32.
//     A quarter of the DSMEM bytes of an all-gather of four gate blocks, and no second weight image.
//   * every exchange is st.async: data and complete_tx on the receiver's mbarrier in one message.  The first version
//     used st.shared::cluster + barrier.cluster per step; barrier.cluster.arrive.release compiles to MEMBAR.ALL.GPU +
//     CGA barrier and wait.acquire to an L1 invalidate (4.2 / 5.9 us per step forward / backward).
//
// Saved-tensor layout, gate order (f, i, o, c~) and the pointwise math are those of pk_cell_step.cu, so the two
// paths are interchangeable behind pk_rnn_step_fwd / pk_rnn_step_bwd (PK_LSTM_CLUSTER=0 selects the old one).
#include "pk_common.cuh"
#include "pk_kernels.h"

#include <algorithm>
#include <cstdlib>

namespace pk {

namespace {

constexpr int kRB = 16;  // batch rows per cluster
constexpr int kNG = 4;   // f, i, o, c~
constexpr int kIoS = 20;                 // floats per (tensor, unit) row of the I/O staging: 16 batch rows + 4 (bank spread)
constexpr int kIoFloats = 3 * 8 * kIoS;  // per warp: three [8 units][16 rows] fp32 tensors at a time
constexpr int kBwdStageMin = 2 * 8 * kIoS * 4;  // backward: the per-warp push stage doubles as I/O staging (two fp32 tensors)
constexpr int kMaxSmem = 232448 - 1024;  // 227 KB opt-in limit per CTA on sm_100, minus the static part (mbarriers)

struct Geom {
  int MT, UPC, CL, KT, KPs, GLS;
  size_t w_cta_bytes, smem_fwd, smem_bwd;
};

inline Geom make_geom(int H) {
  Geom g;
  g.MT = (H + 127) / 128;               // 8-unit tiles (= compute warps) per CTA
  g.UPC = 8 * g.MT;                     // hidden units per CTA
  g.CL = (H + g.UPC - 1) / g.UPC;       // CTAs per cluster
  g.KT = (g.CL * g.UPC + 15) / 16;      // k16 steps over the (padded) state vector
  g.KPs = 16 * g.KT + 8;                // row pitch in halves (+8: conflict-free ldmatrix)
  g.GLS = kNG * g.UPC + 8;              // row pitch of the local gate-gradient operand
  g.w_cta_bytes = static_cast<size_t>(kNG) * g.UPC * g.KPs * 2;
  g.smem_fwd = g.w_cta_bytes + static_cast<size_t>(2) * (2 * g.KT) * 256 + static_cast<size_t>(g.MT) * kRB * 8 * 2 +
               static_cast<size_t>(g.MT) * kIoFloats * 4;
  g.smem_bwd = g.w_cta_bytes + static_cast<size_t>(2) * g.CL * kRB * g.UPC * 2 + static_cast<size_t>(kRB) * g.GLS * 2 +
               static_cast<size_t>(g.MT) * std::max(kRB * g.UPC * 2, kBwdStageMin);
  return g;
}

// Wc[c][g][ul][k] = U[(g*H + UPC*c + ul)][k]  (fp16, zero padded): one contiguous image per CTA
__global__ void pack_cluster_kernel(const float* __restrict__ U, int H, int UPC, int CL, int KPs, __half* __restrict__ Wc) {
  const long long total = static_cast<long long>(CL) * kNG * UPC * KPs;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(e % KPs);
    long long r = e / KPs;
    const int ul = static_cast<int>(r % UPC);
    r /= UPC;
    const int g = static_cast<int>(r % kNG);
    const int c = static_cast<int>(r / kNG);
    const int u = c * UPC + ul;
    float v = 0.f;
    if (u < H && k < H) v = U[(static_cast<long long>(g) * H + u) * H + k];
    Wc[e] = f16_sat(v);
  }
}

// bring-up instrumentation: cycle sums per phase of CTA 0 / thread 0 (forward slots 0..5, backward 8..14); slot 15 = enable
__device__ long long g_cl_clk[16];
#define CL_CLK(slot)                                   \
  if (clk_on) {                                        \
    const long long tn_ = clock64();                   \
    g_cl_clk[slot] += tn_ - tclk;                      \
    tclk = tn_;                                        \
  }

template <int ACT>
__device__ __forceinline__ float actf(int act, float x) { return act_fwd_fast(ACT >= 0 ? ACT : act, x); }
template <int ACT>
__device__ __forceinline__ float dactf(int act, float y) { return act_bwd_from_out(ACT >= 0 ? ACT : act, y); }

// =====================================================================================
// forward
// =====================================================================================
struct CFwd {
  int act, T, B, H, ndir, CL, KT, KPs;
  const __half* Wc;
  const float* PT; long long ldp;
  const float* scale; const float* shift;
  const float* mask; float mask_scalar;
  float* HT; __half* HT16; __half* HP16;
  float* SV0; float* SV1; float* SV2; float* SV3; float* SV4;  // f, g (= act(c~) * mask), i, o, c
  long long ldt;
  float* Y32; long long ldy32; __half* Y16; long long ldy16;
};

template <int MT, int ACT>
__global__ void __launch_bounds__(MT * 32, 1) lstm_cluster_fwd_kernel(const CFwd a) {
  constexpr int UPC = 8 * MT;
  constexpr int NTHR = MT * 32;
  extern __shared__ __align__(128) uint8_t smem[];
  const int KPs = a.KPs;
  __half* Wsm = reinterpret_cast<__half*>(smem);                    // [4][UPC][KPs]
  // fp16 state, TILE-MAJOR [2 buffers][2 KT tiles of 8 units][16 rows][8 units]: every ldmatrix 8x8 matrix is 128
  // contiguous bytes (conflict-free without padding) and a warp's tile is ONE contiguous 256-byte run at the receiver
  // (16-byte pieces scattered over 16 rows of a row-major buffer reached a third of the DSMEM bandwidth: 2800 cycles
  // of LSU back-pressure per step in the first version)
  __half* Ssm = Wsm + static_cast<size_t>(kNG) * UPC * KPs;
  __half* stage = Ssm + static_cast<size_t>(2) * (2 * a.KT) * 128;  // [MT][16][8]
  float* iobuf = reinterpret_cast<float*>(stage + static_cast<size_t>(MT) * kRB * 8);  // [MT][kIoFloats] I/O staging
  __shared__ __align__(8) uint64_t step_bar[2];                     // one per state buffer: CL * MT * 256 bytes per fill
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const uint32_t crank = cluster_ctarank();
  const int cl = blockIdx.x / a.CL;
  const int H = a.H, B = a.B, T = a.T;
  const int nrows = a.ndir * B;
  const uint32_t tx_bytes = static_cast<uint32_t>(a.CL) * MT * kRB * 16;

  if (threadIdx.x == 0) {
    mbar_init(&step_bar[0], 1);
    mbar_init(&step_bar[1], 1);
    fence_mbar_init();
  }
  {  // weight slice of this CTA -> shared memory (stays for all T steps)
    const char* src = reinterpret_cast<const char*>(a.Wc) + static_cast<size_t>(crank) * kNG * UPC * KPs * 2;
    const int bytes = kNG * UPC * KPs * 2;
    for (int o = threadIdx.x * 16; o < bytes; o += NTHR * 16) cp_async_16(smem + o, src + o);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2 * (2 * a.KT) * 64; i += NTHR) reinterpret_cast<uint32_t*>(Ssm)[i] = 0u;  // h_{-1} = 0, both buffers
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // nobody pushes into a buffer that is still being zeroed

  // this thread's elements: rows (g, g+8) of the cluster's 16, units (ul0, ul0+1) of the CTA's UPC
  const int ul0 = warp * 8 + 2 * q;
  const int u0 = static_cast<int>(crank) * UPC + ul0;
  bool uok[2], rok[2];
  int rd[2], cstep[2];
  long long col0[2];
  uok[0] = u0 < H;
  uok[1] = u0 + 1 < H;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int rr = cl * kRB + g + 8 * i;
    rok[i] = rr < nrows;
    rd[i] = (rok[i] && rr >= B) ? 1 : 0;
    const int rb = rr - rd[i] * B;
    col0[i] = rd[i] ? static_cast<long long>(T - 1) * B + rb : rb;  // column (t*B + b) at step 0
    cstep[i] = rd[i] ? -B : B;
  }
  float sc[kNG][2], sh[kNG][2], mk[2][2], hp[2][2], cs[2][2];
#pragma unroll
  for (int gg = 0; gg < kNG; ++gg)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      sc[gg][e] = uok[e] ? __ldg(a.scale + gg * H + u0 + e) : 0.f;
      sh[gg][e] = uok[e] ? __ldg(a.shift + gg * H + u0 + e) : 0.f;
    }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      hp[i][e] = 0.f;
      cs[i][e] = 0.f;
      const int rr = cl * kRB + g + 8 * i;
      mk[i][e] = a.mask ? ((rok[i] && uok[e]) ? __ldg(a.mask + static_cast<long long>(rr) * H + u0 + e) : 0.f)
                        : a.mask_scalar;
    }
  // Global I/O runs in a second, row-vectorised view of the warp's [8 units][16 rows] tile: lane -> (unit lane / 4, rows
  // 4 (lane % 4) .. + 3).  In every channel-major tensor these four rows are 16 contiguous, 16-byte aligned bytes when
  // B % 4 == 0 (and the pitches are multiples of 4), so one LDG.128 / STG.128 moves what the MMA-fragment view needs four
  // scattered 4-byte accesses for (16 loads + 40 stores per thread and step were ~4000 of the ~7400 cycles of a
  // step); the two views are exchanged through a per-warp shared-memory tile.
  const bool vec = (B % 4 == 0) && (a.ldp % 4 == 0) && (a.ldt % 4 == 0);
  const int io_ul = lane >> 2, io_rg = lane & 3;
  const int io_u = static_cast<int>(crank) * UPC + warp * 8 + io_ul;
  const int io_r = cl * kRB + 4 * io_rg;
  const bool io_ok = vec && (io_u < H) && (io_r < nrows);
  const int io_d = (io_ok && io_r >= B) ? 1 : 0;
  const long long io_col0 = io_d ? static_cast<long long>(T - 1) * B + (io_r - B) : io_r;
  const int io_cstep = io_d ? -B : B;
  const long long io_chan = static_cast<long long>(io_d * H + io_u) * a.ldt;
  float* my_io = iobuf + warp * kIoFloats;
  float* io_w = my_io + io_ul * kIoS + 4 * io_rg;   // this lane's float4 in tensor slot 0 (row-vectorised view)
  float4 pv[kNG], vhp4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load_pre_vec = [&](int k) {
#pragma unroll
    for (int gg = 0; gg < kNG; ++gg) {
      pv[gg] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (io_ok)
        pv[gg] = __ldg(reinterpret_cast<const float4*>(a.PT + static_cast<long long>(gg * H + io_u) * a.ldp + io_col0 +
                                                       static_cast<long long>(k) * io_cstep));
    }
  };
  float pre[kNG][2][2], pnx[kNG][2][2];
  auto load_pre = [&](int k, float (&dst)[kNG][2][2]) {
#pragma unroll
    for (int gg = 0; gg < kNG; ++gg)
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          dst[gg][i][e] = 0.f;
          if (rok[i] && uok[e])
            dst[gg][i][e] = __ldg(a.PT + static_cast<long long>(gg * H + u0 + e) * a.ldp + col0[i] +
                                  static_cast<long long>(k) * cstep[i]);
        }
  };
  if (vec) load_pre_vec(0); else load_pre(0, pre);

  // A fragment of k-step kt: matrices (rows 0-7 | 8-15) x (tile 2kt | 2kt+1); 512 bytes per k-step
  const uint32_t a_off = static_cast<uint32_t>(((lane >> 4) * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * 16);
  // B fragments of TWO k-steps of one gate in one ldmatrix.x4: matrices = k offsets 0, 8, 16, 24 of the warp's 8 weight rows
  const uint32_t b_base = smem_u32(Wsm) + static_cast<uint32_t>(((warp * 8 + (lane & 7)) * KPs + 8 * (lane >> 3)) * 2);
  const uint32_t gate_bytes = static_cast<uint32_t>(UPC * KPs * 2);
  const uint32_t s_base = smem_u32(Ssm);
  const uint32_t buf_bytes = static_cast<uint32_t>(2 * a.KT * 256);
  __half* my_stage = stage + warp * kRB * 8;

  const bool clk_on = g_cl_clk[15] != 0 && blockIdx.x == 0 && threadIdx.x == 0;
  long long tclk = clock64();
  for (int k = 0; k < T; ++k) {
    const int cur = k & 1, nxt = cur ^ 1;
    if (vec) {
      // projections of THIS step: row-vectorised registers -> fragment view (two gate blocks per round)
#pragma unroll
      for (int r2 = 0; r2 < 2; ++r2) {
        *reinterpret_cast<float4*>(io_w) = pv[2 * r2];
        *reinterpret_cast<float4*>(io_w + 8 * kIoS) = pv[2 * r2 + 1];
        __syncwarp();
#pragma unroll
        for (int gl = 0; gl < 2; ++gl)
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int e = 0; e < 2; ++e) pre[2 * r2 + gl][i][e] = my_io[(gl * 8 + 2 * q + e) * kIoS + g + 8 * i];
        __syncwarp();
      }
      if (k + 1 < T) load_pre_vec(k + 1);    // lands behind the barrier wait and the MMAs
    } else if (k + 1 < T) {
      load_pre(k + 1, pnx);
    }
    CL_CLK(0)
    if (k > 0) mbar_wait(&step_bar[cur], ((k - 1) >> 1) & 1);  // h_{k-1} of every CTA has landed in buffer `cur`
    if (threadIdx.x == 0) mbar_arrive_expect_tx(&step_bar[nxt], tx_bytes);
    CL_CLK(1)
    float acc[kNG][4];
#pragma unroll
    for (int gg = 0; gg < kNG; ++gg) acc[gg][0] = acc[gg][1] = acc[gg][2] = acc[gg][3] = 0.f;
    const uint32_t a_base = s_base + cur * buf_bytes + a_off;
    // fragment loads of the NEXT pair of k-steps are issued before the MMAs of this pair (the asm statements keep
    // program order, so a load -> MMA -> load -> MMA sequence would expose the ldmatrix latency 4 x KT times per step);
    // one ldmatrix.x4 per gate brings the B fragments of both k-steps (x2 loads moved half the bytes per instruction)
    {
      struct Frag { uint32_t a0[4], a1[4], b[kNG][4]; };
      Frag f0, f1;
      auto ld2 = [&](int kt, Frag& f) {
        ldmatrix_x4(a_base + kt * 512, f.a0[0], f.a0[1], f.a0[2], f.a0[3]);
        ldmatrix_x4(a_base + kt * 512 + 512, f.a1[0], f.a1[1], f.a1[2], f.a1[3]);
#pragma unroll
        for (int gg = 0; gg < kNG; ++gg)
          ldmatrix_x4(b_base + gg * gate_bytes + kt * 32, f.b[gg][0], f.b[gg][1], f.b[gg][2], f.b[gg][3]);
      };
      auto mm2 = [&](const Frag& f) {
#pragma unroll
        for (int gg = 0; gg < kNG; ++gg) mma_m16n8k16_f16(acc[gg], f.a0, f.b[gg][0], f.b[gg][1]);
#pragma unroll
        for (int gg = 0; gg < kNG; ++gg) mma_m16n8k16_f16(acc[gg], f.a1, f.b[gg][2], f.b[gg][3]);
      };
      const int KT = a.KT;
      const int KP = KT >> 1;        // pairs of k-steps
      if (KP > 0) ld2(0, f0);
      int p2 = 0;
#pragma unroll 1
      for (; p2 + 2 <= KP; p2 += 2) {
        ld2(2 * (p2 + 1), f1);
        mm2(f0);
        if (p2 + 2 < KP) ld2(2 * (p2 + 2), f0);
        mm2(f1);
      }
      if (p2 < KP) mm2(f0);
      if (KT & 1) {                  // odd tail: one k-step with x2 loads of the B fragments
        const int kt = KT - 1;
        uint32_t fa[4];
        ldmatrix_x4(a_base + kt * 512, fa[0], fa[1], fa[2], fa[3]);
#pragma unroll
        for (int gg = 0; gg < kNG; ++gg) {
          uint32_t b0, b1;
          ldmatrix_x2(b_base + gg * gate_bytes + kt * 32, b0, b1);
          mma_m16n8k16_f16(acc[gg], fa, b0, b1);
        }
      }
    }
    CL_CLK(2)
    // ---- gates and state update (reference :457-469; gate blocks f, i, o, c~)
    float hn[2][2], vf[2][2], vg[2][2], vi[2][2], vo[2][2], vc[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        hn[i][e] = vf[i][e] = vg[i][e] = vi[i][e] = vo[i][e] = vc[i][e] = 0.f;
        if (rok[i] && uok[e]) {
          const float ft = sigmoid_fast(fmaf(sc[0][e], pre[0][i][e], sh[0][e]) + acc[0][2 * i + e]);
          const float it = sigmoid_fast(fmaf(sc[1][e], pre[1][i][e], sh[1][e]) + acc[1][2 * i + e]);
          const float ot = sigmoid_fast(fmaf(sc[2][e], pre[2][i][e], sh[2][e]) + acc[2][2 * i + e]);
          const float gt = actf<ACT>(a.act, fmaf(sc[3][e], pre[3][i][e], sh[3][e]) + acc[3][2 * i + e]) * mk[i][e];
          const float ct = fmaf(it, gt, ft * cs[i][e]);
          hn[i][e] = ot * actf<ACT>(a.act, ct);
          cs[i][e] = ct;
          vf[i][e] = ft; vg[i][e] = gt; vi[i][e] = it; vo[i][e] = ot; vc[i][e] = ct;
        }
      }
    CL_CLK(3)
    // ---- all-gather of the new fp16 state: stage the warp's [16 rows][8 units] tile, push its 16-byte rows to
    //      every CTA of the cluster (this one included)
#pragma unroll
    for (int i = 0; i < 2; ++i)
      *reinterpret_cast<uint32_t*>(my_stage + (g + 8 * i) * 8 + 2 * q) = pack_f16x2_sat(hn[i][0], hn[i][1]);
    __syncwarp();
    {
      const int row = lane & 15;
      const uint4 val = *reinterpret_cast<const uint4*>(my_stage + row * 8);
      const uint32_t laddr = s_base + nxt * buf_bytes +
                             static_cast<uint32_t>(((static_cast<int>(crank) * MT + warp) * kRB + row) * 16);
      const uint32_t lbar = smem_u32(&step_bar[nxt]);
      // data and completion (complete_tx on the receiver's mbarrier) travel in one st.async message: no fence, no
      // cluster barrier on the serial path (barrier.cluster.arrive.release compiles to MEMBAR.ALL.GPU + CGA barrier:
      // first version of this kernel, 4.2 us per step)
      for (int dst = (lane >> 4); dst < a.CL; dst += 2) st_async_v4(mapa_shared(laddr, dst), val, mapa_shared(lbar, dst));
    }
    CL_CLK(4)
    // ---- saved tensors / outputs (channel-major, natural time), in the shadow of the exchange
    if (vec) {
      const long long idx = io_chan + io_col0 + static_cast<long long>(k) * io_cstep;
      auto put3 = [&](const float (&t0)[2][2], const float (&t1)[2][2], const float (&t2)[2][2]) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float* w = my_io + (2 * q + e) * kIoS + g + 8 * i;
            w[0] = t0[i][e];
            w[8 * kIoS] = t1[i][e];
            w[16 * kIoS] = t2[i][e];
          }
        __syncwarp();
      };
      put3(vf, vg, vi);
      if (io_ok) {
        if (a.SV0) *reinterpret_cast<float4*>(a.SV0 + idx) = *reinterpret_cast<const float4*>(io_w);
        if (a.SV1) *reinterpret_cast<float4*>(a.SV1 + idx) = *reinterpret_cast<const float4*>(io_w + 8 * kIoS);
        if (a.SV2) *reinterpret_cast<float4*>(a.SV2 + idx) = *reinterpret_cast<const float4*>(io_w + 16 * kIoS);
      }
      __syncwarp();
      put3(vo, vc, hn);
      if (io_ok) {
        if (a.SV3) *reinterpret_cast<float4*>(a.SV3 + idx) = *reinterpret_cast<const float4*>(io_w);
        if (a.SV4) *reinterpret_cast<float4*>(a.SV4 + idx) = *reinterpret_cast<const float4*>(io_w + 8 * kIoS);
        const float4 h4 = *reinterpret_cast<const float4*>(io_w + 16 * kIoS);
        if (a.HT) *reinterpret_cast<float4*>(a.HT + idx) = h4;
        if (a.HT16) {
          uint2 p16;
          p16.x = pack_f16x2_sat(h4.x, h4.y);
          p16.y = pack_f16x2_sat(h4.z, h4.w);
          *reinterpret_cast<uint2*>(a.HT16 + idx) = p16;
        }
        if (a.HP16) {
          uint2 p16;
          p16.x = pack_f16x2_sat(vhp4.x, vhp4.y);
          p16.y = pack_f16x2_sat(vhp4.z, vhp4.w);
          *reinterpret_cast<uint2*>(a.HP16 + idx) = p16;
        }
        vhp4 = h4;
      }
      __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long col = col0[i] + static_cast<long long>(k) * cstep[i];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (rok[i] && uok[e]) {
          const int u = u0 + e;
          if (!vec) {
            const long long cidx = static_cast<long long>(rd[i] * H + u) * a.ldt + col;
            if (a.SV0) a.SV0[cidx] = vf[i][e];
            if (a.SV1) a.SV1[cidx] = vg[i][e];
            if (a.SV2) a.SV2[cidx] = vi[i][e];
            if (a.SV3) a.SV3[cidx] = vo[i][e];
            if (a.SV4) a.SV4[cidx] = vc[i][e];
            if (a.HT) a.HT[cidx] = hn[i][e];
            if (a.HT16) a.HT16[cidx] = f16_sat(hn[i][e]);
            if (a.HP16) a.HP16[cidx] = f16_sat(hp[i][e]);
          }
          // row-major module output / next layer's operand: for a fixed row the tile's units are contiguous
          if (a.Y32) a.Y32[col * a.ldy32 + rd[i] * H + u] = hn[i][e];
          if (a.Y16) a.Y16[col * a.ldy16 + rd[i] * H + u] = f16_sat(hn[i][e]);
        }
        hp[i][e] = hn[i][e];
      }
    }
    if (!vec) {
#pragma unroll
      for (int gg = 0; gg < kNG; ++gg)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int e = 0; e < 2; ++e) pre[gg][i][e] = pnx[gg][i][e];
    }
    CL_CLK(5)
  }
  mbar_wait(&step_bar[T & 1], ((T - 1) >> 1) & 1);  // drain the last incoming fill
  cluster_sync_all();                                // no CTA exits while peers may still address its shared memory
}

// =====================================================================================
// backward
// =====================================================================================
struct CBwd {
  int act, T, B, H, ndir, CL, KT, KPs;
  const __half* Wc;
  const float* dYT;
  const float* SV0; const float* SV1; const float* SV2; const float* SV3; const float* SV4;
  long long ldt;
  const float* mask; float mask_scalar;
  const float* gscale;
  __half* GT16;  // [ndir][4*H][ldt] fp16, scaled by *gscale
};

template <int MT, int ACT>
__global__ void __launch_bounds__(MT * 32, 1) lstm_cluster_bwd_kernel(const CBwd a) {
  constexpr int UPC = 8 * MT;
  constexpr int NTHR = MT * 32;
  constexpr int GLS = kNG * UPC + 8;
  constexpr int KS = kNG * UPC / 16;      // k16 steps of the local contraction (= 2 MT)
  constexpr int BLK = kRB * UPC;          // halves of one [16 rows][UPC units] partial block
  constexpr int NCHUNK = BLK * 2 / 16;    // its 16-byte pieces (= 2 UPC)
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t step_bar[2];
  const int KPs = a.KPs;
  __half* Wsm = reinterpret_cast<__half*>(smem);                              // [4][UPC][KPs]
  __half* recv = Wsm + static_cast<size_t>(kNG) * UPC * KPs;                  // [2][CL][16][UPC] partial dh (scaled fp16)
  __half* Gl = recv + static_cast<size_t>(2) * a.CL * BLK;                    // [16][GLS] own gate gradients (A operand)
  __half* stage = Gl + static_cast<size_t>(kRB) * GLS;                        // [MT][STG] per warp: one partial block / I/O staging
  constexpr int STG = (BLK * 2 > kBwdStageMin ? BLK * 2 : kBwdStageMin) / 2;  // halves per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const uint32_t crank = cluster_ctarank();
  const int cl = blockIdx.x / a.CL;
  const int H = a.H, B = a.B, T = a.T;
  const int nrows = a.ndir * B;
  const uint32_t tx_bytes = static_cast<uint32_t>(a.CL) * BLK * 2;

  if (threadIdx.x == 0) {
    mbar_init(&step_bar[0], 1);
    mbar_init(&step_bar[1], 1);
    fence_mbar_init();
  }
  {
    const char* src = reinterpret_cast<const char*>(a.Wc) + static_cast<size_t>(crank) * kNG * UPC * KPs * 2;
    const int bytes = kNG * UPC * KPs * 2;
    for (int o = threadIdx.x * 16; o < bytes; o += NTHR * 16) cp_async_16(smem + o, src + o);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int i = threadIdx.x; i < kRB * GLS / 2; i += NTHR) reinterpret_cast<uint32_t*>(Gl)[i] = 0u;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  cluster_sync_all();

  const int ul0 = warp * 8 + 2 * q;
  const int u0 = static_cast<int>(crank) * UPC + ul0;
  bool uok[2], rok[2];
  int rd[2], cstep[2];
  long long col0[2];
  uok[0] = u0 < H;
  uok[1] = u0 + 1 < H;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int rr = cl * kRB + g + 8 * i;
    rok[i] = rr < nrows;
    rd[i] = (rok[i] && rr >= B) ? 1 : 0;
    const int rb = rr - rd[i] * B;
    col0[i] = rd[i] ? static_cast<long long>(T - 1) * B + rb : rb;
    cstep[i] = rd[i] ? -B : B;
  }
  float mk[2][2], rm[2][2], kc[2][2], carry[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int rr = cl * kRB + g + 8 * i;
      mk[i][e] = a.mask ? ((rok[i] && uok[e]) ? __ldg(a.mask + static_cast<long long>(rr) * H + u0 + e) : 0.f)
                        : a.mask_scalar;
      rm[i][e] = (mk[i][e] != 0.f) ? rcp_approx(mk[i][e]) : 0.f;
      kc[i][e] = 0.f;
      carry[i][e] = 0.f;
    }
  const float s = a.gscale ? __ldg(a.gscale) : 1.f;
  const float inv_s = 1.f / s;

  // operands of one step: dy, f, g, i, o, c, c_prev
  float op[7][2][2], opn[7][2][2];
  auto load_ops = [&](int k, float (&dst)[7][2][2]) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
#pragma unroll
        for (int j = 0; j < 7; ++j) dst[j][i][e] = 0.f;
        if (rok[i] && uok[e]) {
          const long long cidx = static_cast<long long>(rd[i] * H + u0 + e) * a.ldt + col0[i] +
                                 static_cast<long long>(k) * cstep[i];
          dst[0][i][e] = __ldg(a.dYT + cidx);
          dst[1][i][e] = __ldg(a.SV0 + cidx);
          dst[2][i][e] = __ldg(a.SV1 + cidx);
          dst[3][i][e] = __ldg(a.SV2 + cidx);
          dst[4][i][e] = __ldg(a.SV3 + cidx);
          dst[5][i][e] = __ldg(a.SV4 + cidx);
          if (k > 0) dst[6][i][e] = __ldg(a.SV4 + cidx - cstep[i]);
        }
      }
  };
  const uint32_t w_base = smem_u32(Wsm);
  const uint32_t gl_a = smem_u32(Gl) + static_cast<uint32_t>((((lane & 7) + 8 * ((lane >> 3) & 1)) * GLS + 8 * (lane >> 4)) * 2);
  __half* my_stage = stage + warp * STG;
  const long long gate_stride = static_cast<long long>(H) * a.ldt;
  // row-vectorised I/O view (see the forward kernel): lane -> (unit lane / 4, rows 4 (lane % 4) .. + 3)
  const bool vec = (B % 4 == 0) && (a.ldt % 4 == 0);
  const int io_ul = lane >> 2, io_rg = lane & 3;
  const int io_u = static_cast<int>(crank) * UPC + warp * 8 + io_ul;
  const int io_r = cl * kRB + 4 * io_rg;
  const bool io_ok = vec && (io_u < H) && (io_r < nrows);
  const int io_d = (io_ok && io_r >= B) ? 1 : 0;
  const long long io_col0 = io_d ? static_cast<long long>(T - 1) * B + (io_r - B) : io_r;
  const int io_cstep = io_d ? -B : B;
  const long long io_chan = static_cast<long long>(io_d * H + io_u) * a.ldt;
  float* my_io = reinterpret_cast<float*>(my_stage);
  float* io_w = my_io + io_ul * kIoS + 4 * io_rg;
  float4 pv[7];
  auto load_ops_vec = [&](int k) {
#pragma unroll
    for (int j = 0; j < 7; ++j) pv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (io_ok) {
      const long long idx = io_chan + io_col0 + static_cast<long long>(k) * io_cstep;
      pv[0] = __ldg(reinterpret_cast<const float4*>(a.dYT + idx));
      pv[1] = __ldg(reinterpret_cast<const float4*>(a.SV0 + idx));
      pv[2] = __ldg(reinterpret_cast<const float4*>(a.SV1 + idx));
      pv[3] = __ldg(reinterpret_cast<const float4*>(a.SV2 + idx));
      pv[4] = __ldg(reinterpret_cast<const float4*>(a.SV3 + idx));
      pv[5] = __ldg(reinterpret_cast<const float4*>(a.SV4 + idx));
      if (k > 0) pv[6] = __ldg(reinterpret_cast<const float4*>(a.SV4 + idx - io_cstep));
    }
  };

  if (vec) load_ops_vec(T - 1); else load_ops(T - 1, op);
  const bool clk_on = g_cl_clk[15] != 0 && blockIdx.x == 0 && threadIdx.x == 0;
  long long tclk = clock64();
  for (int it = 0; it < T; ++it) {
    const int k = T - 1 - it;
    const int buf = it & 1;
    if (threadIdx.x == 0 && k > 0) mbar_arrive_expect_tx(&step_bar[buf], tx_bytes);
    if (vec) {
      // operands of THIS step: row-vectorised registers -> fragment view, two tensors per round
#pragma unroll
      for (int r2 = 0; r2 < 4; ++r2) {
        *reinterpret_cast<float4*>(io_w) = pv[2 * r2];
        if (2 * r2 + 1 < 7) *reinterpret_cast<float4*>(io_w + 8 * kIoS) = pv[2 * r2 + 1];
        __syncwarp();
#pragma unroll
        for (int jl = 0; jl < 2; ++jl)
          if (2 * r2 + jl < 7) {
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
              for (int e = 0; e < 2; ++e) op[2 * r2 + jl][i][e] = my_io[(jl * 8 + 2 * q + e) * kIoS + g + 8 * i];
          }
        __syncwarp();
      }
      if (k > 0) load_ops_vec(k - 1);
    } else if (k > 0) {
      load_ops(k - 1, opn);
    }
    CL_CLK(8)
    // ---- B: pointwise backward of step k (same algebra as pk_cell_step.cu, M_LSTM); `carry` = U^T dpre of step k+1
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long col = col0[i] + static_cast<long long>(k) * cstep[i];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float gq[kNG] = {0.f, 0.f, 0.f, 0.f};
        const bool ok = rok[i] && uok[e];
        if (ok) {
          float dh = op[0][i][e];
          if (it > 0) dh = fmaf(carry[i][e], inv_s, dh);
          const float f = op[1][i][e], gt = op[2][i][e], ig = op[3][i][e], o = op[4][i][e], c = op[5][i][e], cp = op[6][i][e];
          const float m = mk[i][e];
          const float ac = actf<ACT>(a.act, c);
          float dc = dh * o * dactf<ACT>(a.act, ac);
          if (it > 0) dc += kc[i][e];
          gq[0] = dc * cp * f * (1.f - f);                                   // forget gate
          gq[1] = dc * gt * ig * (1.f - ig);                                 // input gate
          gq[2] = dh * ac * o * (1.f - o);                                   // output gate
          gq[3] = dc * ig * m * dactf<ACT>(a.act, gt * rm[i][e]);            // candidate
          kc[i][e] = dc * f;
        }
#pragma unroll
        for (int gg = 0; gg < kNG; ++gg) {
          const __half hv = f16_sat(gq[gg] * s);
          Gl[(g + 8 * i) * GLS + gg * UPC + ul0 + e] = hv;
          if (vec)
            my_stage[(gg * 8 + 2 * q + e) * kIoS + g + 8 * i] = hv;
          else if (ok)
            a.GT16[(static_cast<long long>(rd[i]) * kNG + gg) * gate_stride + static_cast<long long>(u0 + e) * a.ldt + col] = hv;
        }
      }
    }
    if (vec) {  // gate gradients of the step: 4 rows x 2 bytes per (gate, unit) = one 8-byte store
      __syncwarp();
      if (io_ok) {
        const long long cidx = static_cast<long long>(io_u) * a.ldt + io_col0 + static_cast<long long>(k) * io_cstep;
#pragma unroll
        for (int gg = 0; gg < kNG; ++gg)
          *reinterpret_cast<uint2*>(a.GT16 + (static_cast<long long>(io_d) * kNG + gg) * gate_stride + cidx) =
              *reinterpret_cast<const uint2*>(my_stage + (gg * 8 + io_ul) * kIoS + 4 * io_rg);
      }
      __syncwarp();
    }
    CL_CLK(9)
    // ---- C / D: partial dh_{k-1} for every owner CTA (K-split over this CTA's gate rows), reduce-scatter
    if (k > 0) {
      __syncthreads();
      uint32_t af[KS][4];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) ldmatrix_x4(gl_a + ks * 32, af[ks][0], af[ks][1], af[ks][2], af[ks][3]);
      const uint32_t rbase = smem_u32(recv + (static_cast<size_t>(buf) * a.CL + crank) * BLK);
      const uint32_t lbar = smem_u32(&step_bar[buf]);
      for (int d = warp; d < a.CL; d += MT) {
        float acc[MT][4];
#pragma unroll
        for (int t = 0; t < MT; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
        const uint32_t b_base = w_base + static_cast<uint32_t>(((lane & 15) * KPs + d * UPC) * 2);
        {  // B fragments of k-step ks+1 are loaded before the MMAs of ks (see the forward kernel)
          uint32_t f0[MT][2], f1[MT][2];
          auto ldb = [&](int ks, uint32_t (&f)[MT][2]) {  // one x4 per pair of n-tiles (lanes 16-31: the second tile)
#pragma unroll
            for (int t = 0; t + 1 < MT; t += 2)
              ldmatrix_x4_trans(b_base + static_cast<uint32_t>((ks * 16 * KPs + 8 * t) * 2) + (lane >> 4) * 16, f[t][0], f[t][1],
                                f[t + 1][0], f[t + 1][1]);
            if (MT & 1)
              ldmatrix_x2_trans(b_base + static_cast<uint32_t>((ks * 16 * KPs + 8 * (MT - 1)) * 2), f[MT - 1][0], f[MT - 1][1]);
          };
          ldb(0, f0);
#pragma unroll
          for (int ks = 0; ks < KS; ks += 2) {  // KS = 2 MT is even
            ldb(ks + 1, f1);
#pragma unroll
            for (int t = 0; t < MT; ++t) mma_m16n8k16_f16(acc[t], af[ks], f0[t][0], f0[t][1]);
            if (ks + 2 < KS) ldb(ks + 2, f0);
#pragma unroll
            for (int t = 0; t < MT; ++t) mma_m16n8k16_f16(acc[t], af[ks + 1], f1[t][0], f1[t][1]);
          }
        }
        // the [16 rows][UPC] block of owner d: staged as fp16 (it carries the loss scale like the operands), pushed
        // as one contiguous run of 16-byte pieces; data + complete_tx on the owner's mbarrier in one message
        __syncwarp();  // the previous round's pieces have been read
#pragma unroll
        for (int t = 0; t < MT; ++t) {
          *reinterpret_cast<uint32_t*>(my_stage + g * UPC + 8 * t + 2 * q) = pack_f16x2_sat(acc[t][0], acc[t][1]);
          *reinterpret_cast<uint32_t*>(my_stage + (g + 8) * UPC + 8 * t + 2 * q) = pack_f16x2_sat(acc[t][2], acc[t][3]);
        }
        __syncwarp();
        const uint32_t daddr = mapa_shared(rbase, d), dbar = mapa_shared(lbar, d);
        for (int j = lane; j < NCHUNK; j += 32)
          st_async_v4(daddr + j * 16, *reinterpret_cast<const uint4*>(my_stage + j * 8), dbar);
      }
      CL_CLK(10)
      // ---- A: the CL partial blocks for this CTA's units have landed -> carry into step k-1
      mbar_wait(&step_bar[buf], (it >> 1) & 1);
      CL_CLK(11)
      const __half* rb = recv + static_cast<size_t>(buf) * a.CL * BLK;
#pragma unroll
      for (int i = 0; i < 2; ++i) carry[i][0] = carry[i][1] = 0.f;
      for (int src = 0; src < a.CL; ++src) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float2 v = __half22float2(*reinterpret_cast<const __half2*>(rb + (static_cast<size_t>(src) * kRB + g + 8 * i) * UPC + ul0));
          carry[i][0] += v.x;
          carry[i][1] += v.y;
        }
      }
      CL_CLK(12)
    }
    if (!vec) {
#pragma unroll
      for (int j = 0; j < 7; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int e = 0; e < 2; ++e) op[j][i][e] = opn[j][i][e];
    }
  }
  cluster_sync_all();  // no CTA exits while peers may still address its shared memory
}

// =====================================================================================
// host side
// =====================================================================================
template <typename Args, void (*Kern)(const Args)>
int launch_cluster(const Args& a, int cluster, int nclusters, int threads, size_t smem, cudaStream_t stream) {
  static PerDeviceOnce once;
  const cudaError_t err = once.run([&] {
    cudaError_t e = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(Kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    return e;
  });
  PK_CHECK_CUDA(err);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters * cluster, 1, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // a cluster of this size with this much shared memory per CTA may not be schedulable on every part (GPC sizes,
  // MIG slices): report -1 and let the caller run the step-wise family instead of failing the layer
  const cudaError_t le = cudaLaunchKernelEx(&cfg, Kern, a);
  if (le != cudaSuccess) {
    cudaGetLastError();
    set_last_error("cluster-persistent kernel launch failed (%s): falling back to the step-wise kernels", cudaGetErrorString(le));
    return -1;
  }
  return 0;
}

#define PK_CL_LAUNCH(ARGS, KERN, MT, SMEM)                                                                  \
  (p.act == ACT_TANH ? launch_cluster<ARGS, KERN<MT, ACT_TANH>>(p, G.CL, nclusters, MT * 32, SMEM, stream)  \
                     : launch_cluster<ARGS, KERN<MT, -1>>(p, G.CL, nclusters, MT * 32, SMEM, stream))

constexpr int kDefaultOn = 1;  // verified on B200 against the step-wise family and the full-size reference fixture (profiles/r2_lstm_cluster.txt)

}  // namespace

// bring-up: enable / read-and-reset the phase clocks (not part of the C ABI)
extern "C" int pk_debug_cluster_clocks(int enable, long long* out16) {
  long long h[16];
  if (cudaMemcpyFromSymbol(h, g_cl_clk, sizeof(h)) != cudaSuccess) return 1;
  if (out16) for (int i = 0; i < 16; ++i) out16[i] = h[i];
  for (int i = 0; i < 16; ++i) h[i] = 0;
  h[15] = enable;
  return cudaMemcpyToSymbol(g_cl_clk, h, sizeof(h)) != cudaSuccess;
}

// Can the cluster-persistent kernels run this LSTM layer?  (PK_LSTM_CLUSTER=0 / 1 overrides the default.)
bool lstm_cluster_usable(int cell, int H) {
  if (cell != CELL_LSTM) return false;
  const char* e = getenv("PK_LSTM_CLUSTER");
  const bool on = e ? (e[0] != '0') : (kDefaultOn != 0);
  if (!on || H < 1) return false;
  const Geom G = make_geom(H);
  return G.MT <= 5 && G.CL <= 16 && G.CL >= G.MT && G.smem_fwd <= static_cast<size_t>(kMaxSmem) &&
         G.smem_bwd <= static_cast<size_t>(kMaxSmem);
}

long long lstm_cluster_pack_bytes(int H) {
  const Geom G = make_geom(H);
  return static_cast<long long>(G.CL) * static_cast<long long>(G.w_cta_bytes);
}

int lstm_cluster_fwd(const CellStepFwdArgs& a, __half* Wc, cudaStream_t stream) {
  const Geom G = make_geom(a.H);
  pack_cluster_kernel<<<296, 256, 0, stream>>>(a.U, a.H, G.UPC, G.CL, G.KPs, Wc);
  PK_CHECK_CUDA(cudaGetLastError());
  CFwd p;
  p.act = a.act; p.T = a.T; p.B = a.B; p.H = a.H; p.ndir = a.ndir; p.CL = G.CL; p.KT = G.KT; p.KPs = G.KPs;
  p.Wc = Wc; p.PT = a.PT; p.ldp = a.ldp; p.scale = a.scale; p.shift = a.shift; p.mask = a.mask; p.mask_scalar = a.mask_scalar;
  p.HT = a.HT; p.HT16 = a.HT16; p.HP16 = a.HP16;
  p.SV0 = a.SV[0]; p.SV1 = a.SV[1]; p.SV2 = a.SV[2]; p.SV3 = a.SV[3]; p.SV4 = a.SV[4];
  p.ldt = a.ldt; p.Y32 = a.Y32; p.ldy32 = a.ldy32; p.Y16 = a.Y16; p.ldy16 = a.ldy16;
  const int nclusters = (a.ndir * a.B + kRB - 1) / kRB;
  switch (G.MT) {
    case 1: return PK_CL_LAUNCH(CFwd, lstm_cluster_fwd_kernel, 1, G.smem_fwd);
    case 2: return PK_CL_LAUNCH(CFwd, lstm_cluster_fwd_kernel, 2, G.smem_fwd);
    case 3: return PK_CL_LAUNCH(CFwd, lstm_cluster_fwd_kernel, 3, G.smem_fwd);
    case 4: return PK_CL_LAUNCH(CFwd, lstm_cluster_fwd_kernel, 4, G.smem_fwd);
    default: return PK_CL_LAUNCH(CFwd, lstm_cluster_fwd_kernel, 5, G.smem_fwd);
  }
}

int lstm_cluster_bwd(const CellStepBwdArgs& a, __half* Wc, cudaStream_t stream) {
  const Geom G = make_geom(a.H);
  pack_cluster_kernel<<<296, 256, 0, stream>>>(a.U, a.H, G.UPC, G.CL, G.KPs, Wc);
  PK_CHECK_CUDA(cudaGetLastError());
  CBwd p;
  p.act = a.act; p.T = a.T; p.B = a.B; p.H = a.H; p.ndir = a.ndir; p.CL = G.CL; p.KT = G.KT; p.KPs = G.KPs;
  p.Wc = Wc; p.dYT = a.dYT;
  p.SV0 = a.SV[0]; p.SV1 = a.SV[1]; p.SV2 = a.SV[2]; p.SV3 = a.SV[3]; p.SV4 = a.SV[4];
  p.ldt = a.ldt; p.mask = a.mask; p.mask_scalar = a.mask_scalar; p.gscale = a.gscale; p.GT16 = a.GT16;
  const int nclusters = (a.ndir * a.B + kRB - 1) / kRB;
  switch (G.MT) {
    case 1: return PK_CL_LAUNCH(CBwd, lstm_cluster_bwd_kernel, 1, G.smem_bwd);
    case 2: return PK_CL_LAUNCH(CBwd, lstm_cluster_bwd_kernel, 2, G.smem_bwd);
    case 3: return PK_CL_LAUNCH(CBwd, lstm_cluster_bwd_kernel, 3, G.smem_bwd);
    case 4: return PK_CL_LAUNCH(CBwd, lstm_cluster_bwd_kernel, 4, G.smem_bwd);
    default: return PK_CL_LAUNCH(CBwd, lstm_cluster_bwd_kernel, 5, G.smem_bwd);
  }
}

}  // namespace pk
