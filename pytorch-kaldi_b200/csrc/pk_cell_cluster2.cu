// pk_cell_cluster2.cu — cluster-persistent GRU / minimalGRU recurrences (reference neural_networks.py:486-654 GRU, time
// loop :623-640; :1158-1316 minimalGRU, time loop :1287-1300).
//
// Same decomposition as the LSTM kernels of pk_cell_cluster.cu (16 batch rows per thread-block cluster, CTA c owns the
// hidden units [UPC*c, UPC*c + UPC) for all gate blocks, its fp16 weight slice stationary in shared memory, exchanges
// as st.async messages that complete on the receiver's mbarrier) — but these cells contract the candidate's recurrent
// matrix with (gate * h_{t-1}), so a time step is TWO dependent products and two exchanges:
//
//   forward   1: [z (, r)] = sigmoid(P + U_z|r h_{t-1})        -> x = (r | z) * h_{t-1} for the own units, all-gather x
//             2: hc = act(P_h + U_h x) * mask, h_t = z h_{t-1} + (1 - z) hc                 -> all-gather h_t
//   backward  A: da = dh (1 - z) mask act'(.)                  -> K-split partial U_h^T da, reduce-scatter -> v = d(x)
//             B: dpz (, dpr) from dh, v                        -> K-split partial U_z|r^T [dpz; dpr], reduce-scatter
//                dh_{t-1} = dh z + v (r | z) + the sum of the received partials
//
// Gate blocks (PT / scale / shift / U rows / GT16): (h, z, r) for GRU, (h, z) for minimalGRU; saved tensors sv0 = z,
// sv1 = masked candidate, sv2 = r, HX16 = fp16 x — exactly the interface of the step-wise kernels (pk_cell_step.cu), so
// the families are interchangeable behind pk_rnn_step_fwd / pk_rnn_step_bwd (PK_GRU_CLUSTER=0 / 1).
#include "pk_common.cuh"
#include "pk_kernels.h"

#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace pk {

namespace {

constexpr int kRB = 16;                   // batch rows per cluster
constexpr int kMaxSmem = 232448 - 1024;   // 227 KB opt-in limit per CTA, minus the static part (mbarriers)

struct Geom2 {
  int NGT, MT, UPC, CL, KT, KPs, KA, KB, GAS, GBS;
  size_t w_cta_bytes, smem_fwd, smem_bwd;
};

inline Geom2 make_geom2(int cell, int H) {
  Geom2 g;
  g.NGT = (cell == CELL_GRU) ? 3 : 2;
  g.MT = (H + 127) / 128;
  g.UPC = 8 * g.MT;
  g.CL = (H + g.UPC - 1) / g.UPC;
  g.KT = (g.CL * g.UPC + 15) / 16;
  g.KPs = 16 * g.KT + 8;
  g.KA = (g.UPC + 15) / 16;                     // k16 steps of phase A (one gate block)
  g.KB = ((g.NGT - 1) * g.UPC + 15) / 16;       // k16 steps of phase B (z [, r])
  g.GAS = 16 * g.KA + 8;
  g.GBS = 16 * g.KB + 8;
  g.w_cta_bytes = static_cast<size_t>(g.NGT) * g.UPC * g.KPs * 2;
  g.smem_fwd = g.w_cta_bytes + static_cast<size_t>(3) * (2 * g.KT) * 256 + static_cast<size_t>(g.MT) * kRB * 8 * 2;
  g.smem_bwd = g.w_cta_bytes + static_cast<size_t>(2) * g.CL * kRB * g.UPC * 2 + static_cast<size_t>(kRB) * (g.GAS + g.GBS) * 2 +
               static_cast<size_t>(g.MT) * kRB * g.UPC * 2;
  return g;
}

// Wc[c][g][ul][k] = U[(g*H + UPC*c + ul)][k]  (fp16, zero padded): one contiguous image per CTA
__global__ void pack_cluster2_kernel(const float* __restrict__ U, int NG, int H, int UPC, int CL, int KPs,
                                     __half* __restrict__ Wc) {
  const long long total = static_cast<long long>(CL) * NG * UPC * KPs;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(e % KPs);
    long long r = e / KPs;
    const int ul = static_cast<int>(r % UPC);
    r /= UPC;
    const int g = static_cast<int>(r % NG);
    const int c = static_cast<int>(r / NG);
    const int u = c * UPC + ul;
    float v = 0.f;
    if (u < H && k < H) v = U[(static_cast<long long>(g) * H + u) * H + k];
    Wc[e] = f16_sat(v);
  }
}

template <int ACT>
__device__ __forceinline__ float actf(int act, float x) { return act_fwd_fast(ACT >= 0 ? ACT : act, x); }
template <int ACT>
__device__ __forceinline__ float dactf(int act, float y) { return act_bwd_from_out(ACT >= 0 ? ACT : act, y); }

// =====================================================================================
// forward
// =====================================================================================
struct GFwd {
  int act, T, B, H, ndir, CL, KT, KPs;
  const __half* Wc;
  const float* PT; long long ldp;
  const float* scale; const float* shift;
  const float* mask; float mask_scalar;
  float* HT; __half* HT16; __half* HP16; __half* HX16;
  float* SV0; float* SV1; float* SV2;  // z, masked candidate, r (GRU)
  long long ldt;
  float* Y32; long long ldy32; __half* Y16; long long ldy16;
};

template <int MT, int NGT, int ACT>
__global__ void __launch_bounds__(MT * 32, 1) gru_cluster_fwd_kernel(const GFwd a) {
  constexpr int UPC = 8 * MT;
  constexpr int NTHR = MT * 32;
  constexpr int NG1 = NGT - 1;  // gate blocks of phase 1: z (, r)
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[3];  // [0], [1]: the two h buffers; [2]: the x buffer
  const int KPs = a.KPs;
  __half* Wsm = reinterpret_cast<__half*>(smem);                    // [NGT][UPC][KPs]
  // operand buffers are TILE-MAJOR [2 KT tiles of 8 units][16 rows][8 units] (see pk_cell_cluster.cu): conflict-free
  // ldmatrix without padding, and a warp's tile lands as one contiguous 256-byte run at every receiver
  const int TB = 2 * a.KT * 128;                                    // halves per buffer
  __half* Ssm = Wsm + static_cast<size_t>(NGT) * UPC * KPs;         // [2] h, double buffered
  __half* Xsm = Ssm + static_cast<size_t>(2) * TB;                  // x = gate * h
  __half* stage = Xsm + static_cast<size_t>(TB);                    // [MT][16][8]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const uint32_t crank = cluster_ctarank();
  const int cl = blockIdx.x / a.CL;
  const int H = a.H, B = a.B, T = a.T;
  const int nrows = a.ndir * B;
  const uint32_t tx_bytes = static_cast<uint32_t>(a.CL) * MT * kRB * 16;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_mbar_init();
  }
  {
    const char* src = reinterpret_cast<const char*>(a.Wc) + static_cast<size_t>(crank) * NGT * UPC * KPs * 2;
    const int bytes = NGT * UPC * KPs * 2;
    for (int o = threadIdx.x * 16; o < bytes; o += NTHR * 16) cp_async_16(smem + o, src + o);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 3 * TB / 2; i += NTHR) reinterpret_cast<uint32_t*>(Ssm)[i] = 0u;  // h_{-1} = 0, x = 0
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
  float sc[NGT][2], sh[NGT][2], mk[2][2], hp[2][2];
#pragma unroll
  for (int gg = 0; gg < NGT; ++gg)
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
      const int rr = cl * kRB + g + 8 * i;
      mk[i][e] = a.mask ? ((rok[i] && uok[e]) ? __ldg(a.mask + static_cast<long long>(rr) * H + u0 + e) : 0.f)
                        : a.mask_scalar;
    }
  float pre[NGT][2][2], pnx[NGT][2][2];
  auto load_pre = [&](int k, float (&dst)[NGT][2][2]) {
#pragma unroll
    for (int gg = 0; gg < NGT; ++gg)
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
  load_pre(0, pre);

  const uint32_t a_off = static_cast<uint32_t>(((lane >> 4) * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * 16);
  const uint32_t b_base = smem_u32(Wsm) + static_cast<uint32_t>(((warp * 8 + (lane & 7)) * KPs + 8 * ((lane >> 3) & 1)) * 2);
  const uint32_t gate_bytes = static_cast<uint32_t>(UPC * KPs * 2);
  const uint32_t s_base = smem_u32(Ssm), x_base = smem_u32(Xsm);
  const uint32_t buf_bytes = static_cast<uint32_t>(TB * 2);
  __half* my_stage = stage + warp * kRB * 8;
  const int KT = a.KT;
  const int prow = lane & 15;
  const uint32_t tile_off = static_cast<uint32_t>(((static_cast<int>(crank) * MT + warp) * kRB + prow) * 16);

  // NB gate blocks starting at block b0 times the [16 x K] operand at a_base (fragment loads one k-step ahead)
  auto product = [&](uint32_t a_base, int b0, auto& acc, auto nb_tag) {
    constexpr int NB = decltype(nb_tag)::value;
    uint32_t fa0[4], fb0[NB][2], fa1[4], fb1[NB][2];
    auto ldk = [&](int kt, uint32_t (&fa)[4], uint32_t (&fb)[NB][2]) {
      ldmatrix_x4(a_base + kt * 512, fa[0], fa[1], fa[2], fa[3]);
#pragma unroll
      for (int gg = 0; gg < NB; ++gg) ldmatrix_x2(b_base + (b0 + gg) * gate_bytes + kt * 32, fb[gg][0], fb[gg][1]);
    };
    auto mmk = [&](const uint32_t (&fa)[4], const uint32_t (&fb)[NB][2]) {
#pragma unroll
      for (int gg = 0; gg < NB; ++gg) mma_m16n8k16_f16(acc[gg], fa, fb[gg][0], fb[gg][1]);
    };
    ldk(0, fa0, fb0);
    int kt = 0;
#pragma unroll 2
    for (; kt + 2 <= KT; kt += 2) {
      ldk(kt + 1, fa1, fb1);
      mmk(fa0, fb0);
      if (kt + 2 < KT) ldk(kt + 2, fa0, fb0);
      mmk(fa1, fb1);
    }
    if (kt < KT) mmk(fa0, fb0);
  };
  // stage the warp's [16 rows][8 units] fp16 tile and push its 16-byte rows into `dst_base` of every CTA
  auto push_tile = [&](const float (&v)[2][2], uint32_t dst_base, uint64_t* bar) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
      *reinterpret_cast<uint32_t*>(my_stage + (g + 8 * i) * 8 + 2 * q) = pack_f16x2_sat(v[i][0], v[i][1]);
    __syncwarp();
    const uint4 val = *reinterpret_cast<const uint4*>(my_stage + prow * 8);
    const uint32_t laddr = dst_base + tile_off;
    const uint32_t lbar = smem_u32(bar);
    for (int dst = (lane >> 4); dst < a.CL; dst += 2) st_async_v4(mapa_shared(laddr, dst), val, mapa_shared(lbar, dst));
    __syncwarp();
  };

  for (int k = 0; k < T; ++k) {
    const int cur = k & 1, nxt = cur ^ 1;
    if (k + 1 < T) load_pre(k + 1, pnx);
    if (k > 0) mbar_wait(&bars[cur], ((k - 1) >> 1) & 1);  // h_{k-1} of every CTA has landed
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(&bars[2], tx_bytes);
      mbar_arrive_expect_tx(&bars[nxt], tx_bytes);
    }
    // ---- phase 1: update (and reset) gate from h_{k-1}; x = gate * h_{k-1} (reference :631-634 / :1293-1295)
    float acc1[NG1][4];
#pragma unroll
    for (int gg = 0; gg < NG1; ++gg) acc1[gg][0] = acc1[gg][1] = acc1[gg][2] = acc1[gg][3] = 0.f;
    product(s_base + cur * buf_bytes + a_off, 1, acc1, std::integral_constant<int, NG1>{});
    float zt[2][2], rt[2][2], xv[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        zt[i][e] = rt[i][e] = xv[i][e] = 0.f;
        if (rok[i] && uok[e]) {
          zt[i][e] = sigmoid_fast(fmaf(sc[1][e], pre[1][i][e], sh[1][e]) + acc1[0][2 * i + e]);
          float gate = zt[i][e];
          if (NGT == 3) {
            rt[i][e] = sigmoid_fast(fmaf(sc[NGT - 1][e], pre[NGT - 1][i][e], sh[NGT - 1][e]) + acc1[NG1 - 1][2 * i + e]);
            gate = rt[i][e];
          }
          xv[i][e] = gate * hp[i][e];
        }
      }
    push_tile(xv, x_base, &bars[2]);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long col = col0[i] + static_cast<long long>(k) * cstep[i];
#pragma unroll
      for (int e = 0; e < 2; ++e)
        if (rok[i] && uok[e]) {
          const long long cidx = static_cast<long long>(rd[i] * H + u0 + e) * a.ldt + col;
          if (a.SV0) a.SV0[cidx] = zt[i][e];
          if (NGT == 3 && a.SV2) a.SV2[cidx] = rt[i][e];
          if (a.HX16) a.HX16[cidx] = f16_sat(xv[i][e]);
        }
    }
    mbar_wait(&bars[2], k & 1);  // x of every CTA has landed
    // ---- phase 2: candidate from x, new state (reference :634-636 / :1295-1297)
    float acc2[1][4];
    acc2[0][0] = acc2[0][1] = acc2[0][2] = acc2[0][3] = 0.f;
    product(x_base + a_off, 0, acc2, std::integral_constant<int, 1>{});
    float hn[2][2], hcv[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        hn[i][e] = hcv[i][e] = 0.f;
        if (rok[i] && uok[e]) {
          const float hc = actf<ACT>(a.act, fmaf(sc[0][e], pre[0][i][e], sh[0][e]) + acc2[0][2 * i + e]) * mk[i][e];
          hcv[i][e] = hc;
          hn[i][e] = fmaf(zt[i][e], hp[i][e] - hc, hc);
        }
      }
    push_tile(hn, s_base + nxt * buf_bytes, &bars[nxt]);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long col = col0[i] + static_cast<long long>(k) * cstep[i];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (rok[i] && uok[e]) {
          const int u = u0 + e;
          const long long cidx = static_cast<long long>(rd[i] * H + u) * a.ldt + col;
          if (a.SV1) a.SV1[cidx] = hcv[i][e];
          if (a.HT) a.HT[cidx] = hn[i][e];
          if (a.HT16) a.HT16[cidx] = f16_sat(hn[i][e]);
          if (a.HP16) a.HP16[cidx] = f16_sat(hp[i][e]);
          if (a.Y32) a.Y32[col * a.ldy32 + rd[i] * H + u] = hn[i][e];
          if (a.Y16) a.Y16[col * a.ldy16 + rd[i] * H + u] = f16_sat(hn[i][e]);
        }
        hp[i][e] = hn[i][e];
      }
    }
#pragma unroll
    for (int gg = 0; gg < NGT; ++gg)
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int e = 0; e < 2; ++e) pre[gg][i][e] = pnx[gg][i][e];
  }
  mbar_wait(&bars[T & 1], ((T - 1) >> 1) & 1);  // drain the last incoming fill
  cluster_sync_all();
}

// =====================================================================================
// backward
// =====================================================================================
struct GBwd {
  int act, T, B, H, ndir, CL, KT, KPs;
  const __half* Wc;
  const float* dYT; const float* HT;
  const float* SV0; const float* SV1; const float* SV2;
  long long ldt;
  const float* mask; float mask_scalar;
  const float* gscale;
  __half* GT16;  // [ndir][NGT*H][ldt] fp16 scaled: blocks (da, dpz [, dpr])
};

template <int MT, int NGT, int ACT>
__global__ void __launch_bounds__(MT * 32, 1) gru_cluster_bwd_kernel(const GBwd a) {
  constexpr int UPC = 8 * MT;
  constexpr int NTHR = MT * 32;
  constexpr int NG1 = NGT - 1;
  constexpr int KA = (UPC + 15) / 16, KB = (NG1 * UPC + 15) / 16;
  constexpr int GAS = 16 * KA + 8, GBS = 16 * KB + 8;
  constexpr int BLK = kRB * UPC;
  constexpr int NCHUNK = BLK * 2 / 16;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2];  // [0]: phase-A partials (v), [1]: phase-B partials (carry)
  const int KPs = a.KPs;
  __half* Wsm = reinterpret_cast<__half*>(smem);                       // [NGT][UPC][KPs]
  __half* recvA = Wsm + static_cast<size_t>(NGT) * UPC * KPs;          // [CL][16][UPC] partial U_h^T da      (scaled fp16)
  __half* recvB = recvA + static_cast<size_t>(a.CL) * BLK;             // [CL][16][UPC] partial U_z|r^T dp    (scaled fp16)
  __half* GA = recvB + static_cast<size_t>(a.CL) * BLK;                // [16][GAS] own da
  __half* GB = GA + static_cast<size_t>(kRB) * GAS;                    // [16][GBS] own (dpz [, dpr])
  __half* stage = GB + static_cast<size_t>(kRB) * GBS;                 // [MT][16][UPC]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const uint32_t crank = cluster_ctarank();
  const int cl = blockIdx.x / a.CL;
  const int H = a.H, B = a.B, T = a.T;
  const int nrows = a.ndir * B;
  const uint32_t tx_bytes = static_cast<uint32_t>(a.CL) * BLK * 2;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  {
    const char* src = reinterpret_cast<const char*>(a.Wc) + static_cast<size_t>(crank) * NGT * UPC * KPs * 2;
    const int bytes = NGT * UPC * KPs * 2;
    for (int o = threadIdx.x * 16; o < bytes; o += NTHR * 16) cp_async_16(smem + o, src + o);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int i = threadIdx.x; i < kRB * (GAS + GBS) / 2; i += NTHR) reinterpret_cast<uint32_t*>(GA)[i] = 0u;  // incl. the k padding
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
  float mk[2][2], rm[2][2], kh[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int rr = cl * kRB + g + 8 * i;
      mk[i][e] = a.mask ? ((rok[i] && uok[e]) ? __ldg(a.mask + static_cast<long long>(rr) * H + u0 + e) : 0.f)
                        : a.mask_scalar;
      rm[i][e] = (mk[i][e] != 0.f) ? rcp_approx(mk[i][e]) : 0.f;
      kh[i][e] = 0.f;
    }
  const float s = a.gscale ? __ldg(a.gscale) : 1.f;
  const float inv_s = 1.f / s;

  // operands of one step: dy, z, hc, r, h_prev
  float op[5][2][2], opn[5][2][2];
  auto load_ops = [&](int k, float (&dst)[5][2][2]) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
#pragma unroll
        for (int j = 0; j < 5; ++j) dst[j][i][e] = 0.f;
        if (rok[i] && uok[e]) {
          const long long cidx = static_cast<long long>(rd[i] * H + u0 + e) * a.ldt + col0[i] +
                                 static_cast<long long>(k) * cstep[i];
          dst[0][i][e] = __ldg(a.dYT + cidx);
          dst[1][i][e] = __ldg(a.SV0 + cidx);
          dst[2][i][e] = __ldg(a.SV1 + cidx);
          if (NGT == 3) dst[3][i][e] = __ldg(a.SV2 + cidx);
          if (k > 0) dst[4][i][e] = __ldg(a.HT + cidx - cstep[i]);
        }
      }
  };
  load_ops(T - 1, op);

  const uint32_t w_base = smem_u32(Wsm);
  const uint32_t a_lane = static_cast<uint32_t>(((lane & 7) + 8 * ((lane >> 3) & 1)));
  const uint32_t ga_a = smem_u32(GA) + (a_lane * GAS + 8 * (lane >> 4)) * 2;
  const uint32_t gb_a = smem_u32(GB) + (a_lane * GBS + 8 * (lane >> 4)) * 2;
  __half* my_stage = stage + warp * BLK;
  const long long gate_stride = static_cast<long long>(H) * a.ldt;

  // partial[16 rows][UPC units of owner d] = G_local[16][NK*16] . W_c[row0 + k][d*UPC + n] for every owner d, pushed
  // to the owners' receive buffer `rbuf` (complete_tx on their `bar`).  Weight rows beyond `nvalid` are padding of the
  // k range: the A columns there are zero, the row index is clamped so that only finite values are multiplied.
  auto scatter = [&](uint32_t g_a, auto nk_tag, int row0, int nvalid, __half* rbuf, uint64_t* bar) {
    constexpr int NK = decltype(nk_tag)::value;
    uint32_t af[NK][4];
#pragma unroll
    for (int ks = 0; ks < NK; ++ks) ldmatrix_x4(g_a + ks * 32, af[ks][0], af[ks][1], af[ks][2], af[ks][3]);
    const uint32_t rbase = smem_u32(rbuf + static_cast<size_t>(crank) * BLK);
    const uint32_t lbar = smem_u32(bar);
    for (int d = warp; d < a.CL; d += MT) {
      float acc[MT][4];
#pragma unroll
      for (int t = 0; t < MT; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < NK; ++ks) {
        const int kr = min(ks * 16 + (lane & 15), nvalid - 1);
        const uint32_t b_row = w_base + static_cast<uint32_t>(((row0 + kr) * KPs + d * UPC) * 2);
        uint32_t f[MT][2];
#pragma unroll
        for (int t = 0; t + 1 < MT; t += 2)  // one x4 per pair of n-tiles (lanes 16-31: the second tile)
          ldmatrix_x4_trans(b_row + 16 * t + (lane >> 4) * 16, f[t][0], f[t][1], f[t + 1][0], f[t + 1][1]);
        if (MT & 1) ldmatrix_x2_trans(b_row + 16 * (MT - 1), f[MT - 1][0], f[MT - 1][1]);
#pragma unroll
        for (int t = 0; t < MT; ++t) mma_m16n8k16_f16(acc[t], af[ks], f[t][0], f[t][1]);
      }
      __syncwarp();
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
  };
  auto gather = [&](const __half* rbuf, float (&out)[2][2]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) out[i][0] = out[i][1] = 0.f;
    for (int src = 0; src < a.CL; ++src) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float2 v = __half22float2(*reinterpret_cast<const __half2*>(rbuf + (static_cast<size_t>(src) * kRB + g + 8 * i) * UPC + ul0));
        out[i][0] += v.x;
        out[i][1] += v.y;
      }
    }
  };

  float carry[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int it = 0; it < T; ++it) {
    const int k = T - 1 - it;
    if (threadIdx.x == 0 && k > 0) {
      mbar_arrive_expect_tx(&bars[0], tx_bytes);
      mbar_arrive_expect_tx(&bars[1], tx_bytes);
    }
    if (k > 0) load_ops(k - 1, opn);
    // ---- phase A: dh complete -> candidate pre-activation gradient (reference autograd of :634-636 / :1295-1297)
    float dh[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long col = col0[i] + static_cast<long long>(k) * cstep[i];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float da = 0.f;
        dh[i][e] = 0.f;
        const bool ok = rok[i] && uok[e];
        if (ok) {
          float d = op[0][i][e];
          if (it > 0) d += kh[i][e] + carry[i][e] * inv_s;
          dh[i][e] = d;
          const float z = op[1][i][e], hc = op[2][i][e];
          da = d * (1.f - z) * mk[i][e] * dactf<ACT>(a.act, hc * rm[i][e]);
        }
        const __half hv = f16_sat(da * s);
        GA[(g + 8 * i) * GAS + ul0 + e] = hv;
        if (ok) a.GT16[(static_cast<long long>(rd[i]) * NGT) * gate_stride + static_cast<long long>(u0 + e) * a.ldt + col] = hv;
      }
    }
    float v[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    if (k > 0) {
      __syncthreads();
      scatter(ga_a, std::integral_constant<int, KA>{}, 0, UPC, recvA, &bars[0]);
      mbar_wait(&bars[0], it & 1);
      gather(recvA, v);  // d(loss) / d(x = gate * h_{k-1}) for the own units, still scaled
    }
    // ---- phase B: gate pre-activation gradients, local part of the carry
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long col = col0[i] + static_cast<long long>(k) * cstep[i];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float gz = 0.f, gr = 0.f;
        const bool ok = rok[i] && uok[e];
        if (ok) {
          const float d = dh[i][e], vv = v[i][e] * inv_s;
          const float z = op[1][i][e], hc = op[2][i][e], hpv = op[4][i][e];
          if (NGT == 3) {  // GRU: x = r * h
            const float r = op[3][i][e];
            gz = d * (hpv - hc) * z * (1.f - z);
            gr = vv * hpv * r * (1.f - r);
            kh[i][e] = fmaf(d, z, vv * r);
          } else {         // minimalGRU: x = z * h
            gz = (d * (hpv - hc) + vv * hpv) * z * (1.f - z);
            kh[i][e] = (d + vv) * z;
          }
        }
        const __half hz = f16_sat(gz * s);
        GB[(g + 8 * i) * GBS + ul0 + e] = hz;
        if (ok) a.GT16[(static_cast<long long>(rd[i]) * NGT + 1) * gate_stride + static_cast<long long>(u0 + e) * a.ldt + col] = hz;
        if (NGT == 3) {
          const __half hr = f16_sat(gr * s);
          GB[(g + 8 * i) * GBS + UPC + ul0 + e] = hr;
          if (ok) a.GT16[(static_cast<long long>(rd[i]) * NGT + 2) * gate_stride + static_cast<long long>(u0 + e) * a.ldt + col] = hr;
        }
      }
    }
    if (k > 0) {
      __syncthreads();
      scatter(gb_a, std::integral_constant<int, KB>{}, UPC, NG1 * UPC, recvB, &bars[1]);
      mbar_wait(&bars[1], it & 1);
      gather(recvB, carry);
    }
#pragma unroll
    for (int j = 0; j < 5; ++j)
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int e = 0; e < 2; ++e) op[j][i][e] = opn[j][i][e];
  }
  cluster_sync_all();
}

// =====================================================================================
// host side
// =====================================================================================
template <typename Args, void (*Kern)(const Args)>
int launch_cluster2(const Args& a, int cluster, int nclusters, int threads, size_t smem, cudaStream_t stream) {
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

// activation: relu (the shipped GRU recipes) and tanh get their own instantiation, the rest is a runtime switch
#define PK_G_ACT(ARGS, KERN, MT, NGT, SMEM)                                                                          \
  (p.act == ACT_RELU   ? launch_cluster2<ARGS, KERN<MT, NGT, ACT_RELU>>(p, G.CL, nclusters, MT * 32, SMEM, stream)  \
   : p.act == ACT_TANH ? launch_cluster2<ARGS, KERN<MT, NGT, ACT_TANH>>(p, G.CL, nclusters, MT * 32, SMEM, stream)  \
                       : launch_cluster2<ARGS, KERN<MT, NGT, -1>>(p, G.CL, nclusters, MT * 32, SMEM, stream))
#define PK_G_NGT(ARGS, KERN, MT, SMEM) (G.NGT == 3 ? PK_G_ACT(ARGS, KERN, MT, 3, SMEM) : PK_G_ACT(ARGS, KERN, MT, 2, SMEM))
#define PK_G_MT(ARGS, KERN, SMEM)                          \
  switch (G.MT) {                                          \
    case 1: return PK_G_NGT(ARGS, KERN, 1, SMEM);          \
    case 2: return PK_G_NGT(ARGS, KERN, 2, SMEM);          \
    case 3: return PK_G_NGT(ARGS, KERN, 3, SMEM);          \
    case 4: return PK_G_NGT(ARGS, KERN, 4, SMEM);          \
    default: return PK_G_NGT(ARGS, KERN, 5, SMEM);         \
  }

constexpr int kDefaultOn2 = 1;  // verified on B200 against the step-wise family (profiles/r2_gru_cluster.txt)

}  // namespace

bool gru_cluster_usable(int cell, int H) {
  if (cell != CELL_GRU && cell != CELL_MGRU) return false;
  const char* e = getenv("PK_GRU_CLUSTER");
  const bool on = e ? (e[0] != '0') : (kDefaultOn2 != 0);
  if (!on || H < 1) return false;
  const Geom2 G = make_geom2(cell, H);
  return G.MT <= 5 && G.CL <= 16 && G.CL >= G.MT && G.smem_fwd <= static_cast<size_t>(kMaxSmem) &&
         G.smem_bwd <= static_cast<size_t>(kMaxSmem);
}

long long gru_cluster_pack_bytes(int cell, int H) {
  const Geom2 G = make_geom2(cell, H);
  return static_cast<long long>(G.CL) * static_cast<long long>(G.w_cta_bytes);
}

int gru_cluster_fwd(const CellStepFwdArgs& a, __half* Wc, cudaStream_t stream) {
  const Geom2 G = make_geom2(a.cell, a.H);
  pack_cluster2_kernel<<<296, 256, 0, stream>>>(a.U, G.NGT, a.H, G.UPC, G.CL, G.KPs, Wc);
  PK_CHECK_CUDA(cudaGetLastError());
  GFwd p;
  p.act = a.act; p.T = a.T; p.B = a.B; p.H = a.H; p.ndir = a.ndir; p.CL = G.CL; p.KT = G.KT; p.KPs = G.KPs;
  p.Wc = Wc; p.PT = a.PT; p.ldp = a.ldp; p.scale = a.scale; p.shift = a.shift; p.mask = a.mask; p.mask_scalar = a.mask_scalar;
  p.HT = a.HT; p.HT16 = a.HT16; p.HP16 = a.HP16; p.HX16 = a.HX16;
  p.SV0 = a.SV[0]; p.SV1 = a.SV[1]; p.SV2 = a.SV[2];
  p.ldt = a.ldt; p.Y32 = a.Y32; p.ldy32 = a.ldy32; p.Y16 = a.Y16; p.ldy16 = a.ldy16;
  const int nclusters = (a.ndir * a.B + kRB - 1) / kRB;
  PK_G_MT(GFwd, gru_cluster_fwd_kernel, G.smem_fwd)
}

int gru_cluster_bwd(const CellStepBwdArgs& a, __half* Wc, cudaStream_t stream) {
  const Geom2 G = make_geom2(a.cell, a.H);
  pack_cluster2_kernel<<<296, 256, 0, stream>>>(a.U, G.NGT, a.H, G.UPC, G.CL, G.KPs, Wc);
  PK_CHECK_CUDA(cudaGetLastError());
  GBwd p;
  p.act = a.act; p.T = a.T; p.B = a.B; p.H = a.H; p.ndir = a.ndir; p.CL = G.CL; p.KT = G.KT; p.KPs = G.KPs;
  p.Wc = Wc; p.dYT = a.dYT; p.HT = a.HT;
  p.SV0 = a.SV[0]; p.SV1 = a.SV[1]; p.SV2 = a.SV[2];
  p.ldt = a.ldt; p.mask = a.mask; p.mask_scalar = a.mask_scalar; p.gscale = a.gscale; p.GT16 = a.GT16;
  const int nclusters = (a.ndir * a.B + kRB - 1) / kRB;
  PK_G_MT(GBwd, gru_cluster_bwd_kernel, G.smem_bwd)
}

}  // namespace pk
