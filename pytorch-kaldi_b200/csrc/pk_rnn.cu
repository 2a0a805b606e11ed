// pk_rnn.cu — persistent, weight-stationary recurrent kernels (liGRU) for sm_100a.
//
// The reference runs `for k in range(T)` in Python with ~11 launches per step
// (neural_networks.py:1130-1141).  Here one launch covers the whole sequence:
//
//   * a thread-block CLUSTER of CL CTAs (8..16) owns 8 rows of the direction-stacked batch; the
//     bidirectional layer is just 2B independent rows (reference :1095-1097 stacks x and
//     flip(x) on the batch axis and shares the weights), so 2B/8 clusters run concurrently;
//   * inside a cluster the hidden units are sliced across the CTAs; every warp keeps its
//     [16 x H] slice of the recurrent matrix U in REGISTERS as mma.sync A-fragments (fp16,
//     loaded once), so the only per-step operand traffic is the hidden state itself;
//   * each step: h_{t-1} (fp16 copy, 8 rows x H) is read from shared memory with ldmatrix,
//     the [gates x 8 rows] products run on the tensor cores with fp32 accumulation, the gate
//     non-linearities / dropout mask / BatchNorm affine are applied in registers on the fp32
//     state, and the new fp16 slice is pushed to every CTA of the cluster through distributed
//     shared memory;
//   * step synchronisation (SY=1, default): the push is `st.async` — each 16-byte store carries
//     its own completion (complete_tx on the TARGET CTA's mbarrier), consumers just wait on
//     their local mbarrier: no fence, no cluster barrier, so the global-memory traffic of the
//     step (saved activations out, next projections in) is never waited for.
//     (SY=0 keeps the first implementation — plain st.shared::cluster + one
//     barrier.cluster per step — whose release fence compiles to MEMBAR.ALL.GPU and stalls on
//     the outstanding global stores; kept for A/B measurement.)
//   * the per-step global inputs (projections / saved activations) are prefetched one step ahead
//     with cp.async straight into shared memory: no registers held across the MMA phase and no
//     scoreboard stall on L2 latency;
//   * flip / stack / cat of the reference (:1144-1150, :1962-1970) disappear into indexing:
//     direction-1 rows read time T-1-k and write their outputs at natural time.
//
// The backward kernel walks the same recurrence in reverse with U^T stationary in registers.
#include "pk_common.cuh"
#include "pk_kernels.h"

#include <cstdlib>
#include <mutex>

namespace pk {

namespace {

constexpr int kRows = 8;  // batch rows per cluster (the n8 of m16n8k16)

__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int SY>
__device__ __forceinline__ void step_wait(uint64_t* bar, uint32_t parity) {
  // the payload is shared memory written by st.async and published by complete_tx on this very
  // barrier: a CTA-scope acquire suffices (and avoids the CCTL.IVALL a cluster-scope acquire emits)
  mbar_wait(bar, parity);
}

// =====================================================================================
// forward
// =====================================================================================
template <int KT, int MT, int CL>
struct FwdSmem {
  static constexpr int HS = CL * 8 * MT + 8;  // halves per staged state row
  __half h16[2][kRows][HS];
  __half stage[MT][kRows][8];
  float pre[2][MT * 32][4];  // cp.async landing zone: {ph0, ph1, pz0, pz1} per thread, double buffered
  uint64_t mbar[2];
};

template <int KT, int MT, int CL, int SY>
__global__ void __launch_bounds__(MT * 32, 1) ligru_fwd_kernel(const RecFwdArgs a) {
  using S = FwdSmem<KT, MT, CL>;
  constexpr int HS = S::HS;
  static_assert((CL * MT) % 2 == 0, "row pitch must be an odd multiple of 16 bytes (ldmatrix conflict-free)");
  static_assert(CL * 8 * MT >= 16 * KT, "unit slots must cover the K range");
  constexpr uint32_t kTxBytes = CL * MT * 128;  // bytes every CTA receives per step
  extern __shared__ __align__(16) uint8_t smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2;
  const int q = lane & 3;
  const uint32_t crank = cluster_ctarank();
  const int cl = blockIdx.x / CL;
  const int H = a.H, B = a.B, T = a.T;
  const int nrows = a.ndir * B;
  const int ubase = crank * (8 * MT) + warp * 8;  // first unit of this warp's 8-unit tile
  const int u = ubase + g;
  const bool u_ok = u < H;
  const bool do_store = !(a.dbg & 1);
  const bool do_load = !(a.dbg & 2);

  // ---- recurrent weights -> A fragments (row g = candidate gate "h", row g+8 = update gate "z")
  uint32_t A[KT][4];
  {
    const float* Uh = a.U + static_cast<long long>(u) * H;
    const float* Uz = a.U + static_cast<long long>(H + u) * H;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      const int k0 = kt * 16 + 2 * q;
      float h00 = 0.f, h01 = 0.f, h10 = 0.f, h11 = 0.f, z00 = 0.f, z01 = 0.f, z10 = 0.f, z11 = 0.f;
      if (u_ok) {
        if (k0 < H) { h00 = __ldg(Uh + k0); z00 = __ldg(Uz + k0); }
        if (k0 + 1 < H) { h01 = __ldg(Uh + k0 + 1); z01 = __ldg(Uz + k0 + 1); }
        if (k0 + 8 < H) { h10 = __ldg(Uh + k0 + 8); z10 = __ldg(Uz + k0 + 8); }
        if (k0 + 9 < H) { h11 = __ldg(Uh + k0 + 9); z11 = __ldg(Uz + k0 + 9); }
      }
      A[kt][0] = pack_f16x2_sat(h00, h01);
      A[kt][1] = pack_f16x2_sat(z00, z01);
      A[kt][2] = pack_f16x2_sat(h10, h11);
      A[kt][3] = pack_f16x2_sat(z10, z11);
    }
  }

  // ---- zero both state buffers (h_0 = 0, reference :1096), init barriers, then cluster-wide sync
  for (int i = threadIdx.x; i < 2 * kRows * HS / 2; i += blockDim.x)
    reinterpret_cast<uint32_t*>(&sm.h16[0][0][0])[i] = 0u;
  for (int i = threadIdx.x; i < 2 * MT * 32 * 4; i += blockDim.x) (&sm.pre[0][0][0])[i] = 0.f;
  if (SY && threadIdx.x == 0) {
    mbar_init(&sm.mbar[0], 1);
    mbar_init(&sm.mbar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  cluster_sync_all();

  // ---- per-thread row bookkeeping: this thread owns (unit u, rows 2q and 2q+1)
  int rb[2], rd[2];
  bool rok[2];
  float msk[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = cl * kRows + 2 * q + i;
    rok[i] = (r < nrows) && u_ok;
    const int rr = r < nrows ? r : 0;
    rd[i] = (rr >= B) ? 1 : 0;
    rb[i] = rr - rd[i] * B;
    msk[i] = a.mask ? ((rok[i]) ? __ldg(a.mask + static_cast<long long>(rr) * H + u) : 0.f)
                    : a.mask_scalar;
  }
  float sc_h = 0.f, sh_h = 0.f, sc_z = 0.f, sh_z = 0.f;
  if (u_ok) {
    sc_h = __ldg(a.scale + u);
    sh_h = __ldg(a.shift + u);
    sc_z = __ldg(a.scale + H + u);
    sh_z = __ldg(a.shift + H + u);
  }
  const float* Ph = a.PT + static_cast<long long>(u_ok ? u : 0) * a.ldp;
  const float* Pz = a.PT + static_cast<long long>(u_ok ? H + u : 0) * a.ldp;
  float* mypre0 = &sm.pre[0][threadIdx.x][0];
  float* mypre1 = &sm.pre[1][threadIdx.x][0];

  // prefetch step 0's projections
#pragma unroll
  for (int i = 0; i < 2; ++i)
    if (rok[i] && do_load) {
      const long long col = static_cast<long long>(rd[i] ? T - 1 : 0) * B + rb[i];
      cp_async_f32(mypre0 + i, Ph + col);
      cp_async_f32(mypre0 + 2 + i, Pz + col);
    }
  cp_async_commit();

  float hprev[2] = {0.f, 0.f};

  // ldmatrix lane addressing: matrix (lane>>3) row (lane&7): &h16[buf][lane&7][k0 + 8*(lane>>3)]
  const uint32_t ldm_off = static_cast<uint32_t>(((lane & 7) * HS + 8 * (lane >> 3)) * 2);
  const uint32_t ldm_off2 = static_cast<uint32_t>(((lane & 7) * HS + 8 * ((lane >> 3) & 1)) * 2);
  const uint32_t h16_base = smem_u32(&sm.h16[0][0][0]);
  constexpr uint32_t kBufBytes = kRows * HS * 2;

  for (int k = 0; k < T; ++k) {
    const int cur = k & 1, nxt = cur ^ 1;
    if (SY) {
      if (k > 0) step_wait<SY>(&sm.mbar[cur], ((k - 1) >> 1) & 1);  // h_{k-1} has landed from all peers
      if (threadIdx.x == 0) mbar_arrive_expect_tx(&sm.mbar[nxt], kTxBytes);  // arm the fill of this step
    }
    // ---------------- U * h_{k-1} on the tensor cores ----------------
    float acc[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f; }
    const uint32_t bufa = h16_base + cur * kBufBytes;
#pragma unroll
    for (int kt = 0; kt + 1 < KT; kt += 2) {
      uint32_t b0, b1, b2, b3;
      ldmatrix_x4(bufa + ldm_off + kt * 32, b0, b1, b2, b3);
      mma_m16n8k16_f16(acc[kt & 3], A[kt], b0, b1);
      mma_m16n8k16_f16(acc[(kt + 1) & 3], A[kt + 1], b2, b3);
    }
    if (KT & 1) {
      uint32_t b0, b1;
      ldmatrix_x2(bufa + ldm_off2 + (KT - 1) * 32, b0, b1);
      mma_m16n8k16_f16(acc[(KT - 1) & 3], A[KT - 1], b0, b1);
    }
    float ch[2], cz[2];
    ch[0] = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
    ch[1] = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
    cz[0] = (acc[0][2] + acc[1][2]) + (acc[2][2] + acc[3][2]);
    cz[1] = (acc[0][3] + acc[1][3]) + (acc[2][3] + acc[3][3]);

    // ---------------- gates (reference :1133-1136) ----------------
    cp_async_wait_all();  // this step's projections (issued one step ago) are in smem
    const float* pre = cur ? mypre1 : mypre0;
    float hn[2], zz[2], hcv[2], hold[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float zt = sigmoidf_(fmaf(sc_z, pre[2 + i], sh_z) + cz[i]);
      const float at = fmaf(sc_h, pre[i], sh_h) + ch[i];
      const float hc = act_fwd(a.act, at) * msk[i];
      float h = zt * hprev[i] + (1.f - zt) * hc;
      if (!rok[i]) h = 0.f;
      hn[i] = h; zz[i] = zt; hcv[i] = hc;
      hold[i] = hprev[i];
      hprev[i] = h;
      if (SY) sm.stage[warp][2 * q + i][g] = f16_sat(h);
      else sm.h16[nxt][2 * q + i][u] = f16_sat(h);
    }
    __syncwarp();
    // ---------------- push the warp's 8x8 fp16 tile to the CTAs of the cluster ----------------
    {
      const int n = lane & 7;
      const uint32_t laddr = smem_u32(&sm.h16[nxt][n][ubase]);
      if (SY) {
        const uint4 val = *reinterpret_cast<const uint4*>(&sm.stage[warp][n][0]);
        const uint32_t lbar = smem_u32(&sm.mbar[nxt]);
#pragma unroll
        for (int dst = (lane >> 3); dst < CL; dst += 4)
          st_async_v4(mapa_shared(laddr, dst), val, mapa_shared(lbar, dst));
      } else {
        const uint4 val = *reinterpret_cast<const uint4*>(&sm.h16[nxt][n][ubase]);
#pragma unroll
        for (int dst = (lane >> 3); dst < CL; dst += 4)
          if (dst != static_cast<int>(crank)) st_cluster_v4(mapa_shared(laddr, dst), val);
      }
    }
    if (!SY) cluster_arrive_release();

    // ---------------- off the critical path: next step's projections + global stores ----------
    {
      float* npre = nxt ? mypre1 : mypre0;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (rok[i]) {
          const int t = rd[i] ? (T - 1 - k) : k;
          const long long col = static_cast<long long>(t) * B + rb[i];
          if (k + 1 < T && do_load) {
            const long long ncol = col + (rd[i] ? -B : B);
            cp_async_f32(npre + i, Ph + ncol);
            cp_async_f32(npre + 2 + i, Pz + ncol);
          }
          if (do_store) {
            const long long ch_idx = static_cast<long long>(rd[i] * H + u) * a.ldt + col;
            if (a.HT) a.HT[ch_idx] = hn[i];
            if (a.HT16) a.HT16[ch_idx] = f16_sat(hn[i]);
            if (a.HP16) a.HP16[ch_idx] = f16_sat(hold[i]);
            if (a.ZT) a.ZT[ch_idx] = zz[i];
            if (a.HCT) a.HCT[ch_idx] = hcv[i];
            if (a.Y32) a.Y32[col * a.ldy32 + rd[i] * H + u] = hn[i];
            if (a.Y16) a.Y16[col * a.ldy16 + rd[i] * H + u] = f16_sat(hn[i]);
          }
        }
      }
      cp_async_commit();
    }
    if (!SY) cluster_wait_acquire();
    else __syncwarp();
  }
  cp_async_wait_all();
  // drain the last incoming fill, then: no CTA may exit while peers can still write into it
  if (SY) step_wait<SY>(&sm.mbar[T & 1], ((T - 1) >> 1) & 1);
  cluster_sync_all();
}

// =====================================================================================
// backward
// =====================================================================================
// One step k (time running backwards):
//   dh   = dY_k + z_{k+1} . dh_{k+1} + U^T [da_{k+1}; dpz_{k+1}]            (carry)
//   dz   = dh . (h_{k-1} - hc_k);   dhc = dh . (1 - z_k)
//   da   = dhc . mask . act'(.)      dpz = dz . z_k (1 - z_k)
// Thread ownership: warp (mt, half) keeps the 16-unit x (half-gate K range) slice of U^T in
// registers; after the MMA the two warps of a pair exchange partial sums through smem so that
// warp (mt,0) finishes units g and warp (mt,1) units g+8 of the tile.
template <int KT, int MT, int CL>
struct BwdSmem {
  static constexpr int MT16 = (MT + 1) / 2;
  static constexpr int NW = MT16 * 2;
  static constexpr int KP = CL * 8 * MT;
  static constexpr int GS = 2 * KP + 8;
  __half g16[2][kRows][GS];
  __half stage[NW][2][kRows][8];
  float xbuf[2][MT16][2][32][2];  // pair exchange, double-buffered by step parity
  float pre[2][NW * 32][8];       // cp.async landing zone: {dy0,dy1,z0,z1,hc0,hc1,hp0,hp1}
  uint64_t mbar[2];
};

template <int KT, int MT, int CL, int SY>
__global__ void __launch_bounds__(((MT + 1) / 2) * 64, 1) ligru_bwd_kernel(const RecBwdArgs a) {
  using S = BwdSmem<KT, MT, CL>;
  constexpr int KP = S::KP;   // unit-slot stride between the two gates in the staged vector
  constexpr int GS = S::GS;   // halves per staged row
  static_assert((2 * CL * MT) % 2 == 0, "row pitch");
  static_assert(KP >= 16 * KT, "unit slots must cover the K range");
  constexpr int UPC = 8 * MT;  // units owned per CTA
  constexpr uint32_t kTxBytes = CL * MT * 256;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2;
  const int q = lane & 3;
  const int mt = warp >> 1;
  const int half = warp & 1;
  const uint32_t crank = cluster_ctarank();
  const int cl = blockIdx.x / CL;
  const int H = a.H, B = a.B, T = a.T;
  const int nrows = a.ndir * B;
  const int cta_ubase = crank * UPC;
  const bool do_store = !(a.dbg & 1);
  const bool do_load = !(a.dbg & 2);

  // ---- U^T slice -> A fragments.  A[row = unit][col = j] = Ug[j][unit], Ug = Uh (half 0) / Uz (half 1)
  uint32_t A[KT][4];
  {
    const int slot_lo = mt * 16 + g, slot_hi = slot_lo + 8;
    const int u_lo = cta_ubase + slot_lo, u_hi = cta_ubase + slot_hi;
    const bool ok_lo = (slot_lo < UPC) && (u_lo < H);
    const bool ok_hi = (slot_hi < UPC) && (u_hi < H);
    const float* Ug = a.U + static_cast<long long>(half) * H * H;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      const int j0 = kt * 16 + 2 * q;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
      // v layout: {lo[j0], lo[j0+1], hi[j0], hi[j0+1], lo[j0+8], lo[j0+9], hi[j0+8], hi[j0+9]}
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = j0 + (e & 1) + ((e >> 1) ? 8 : 0);
        if (j < H) {
          if (ok_lo) v[(e >> 1) * 4 + (e & 1)] = __ldg(Ug + static_cast<long long>(j) * H + u_lo);
          if (ok_hi) v[(e >> 1) * 4 + 2 + (e & 1)] = __ldg(Ug + static_cast<long long>(j) * H + u_hi);
        }
      }
      A[kt][0] = pack_f16x2_sat(v[0], v[1]);
      A[kt][1] = pack_f16x2_sat(v[2], v[3]);
      A[kt][2] = pack_f16x2_sat(v[4], v[5]);
      A[kt][3] = pack_f16x2_sat(v[6], v[7]);
    }
  }

  for (int i = threadIdx.x; i < 2 * kRows * GS / 2; i += blockDim.x)
    reinterpret_cast<uint32_t*>(&sm.g16[0][0][0])[i] = 0u;
  for (int i = threadIdx.x; i < 2 * S::NW * 32 * 8; i += blockDim.x) (&sm.pre[0][0][0])[i] = 0.f;
  if (SY && threadIdx.x == 0) {
    mbar_init(&sm.mbar[0], 1);
    mbar_init(&sm.mbar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  cluster_sync_all();

  // ---- element ownership of this thread: unit = tile unit (8*half + g), rows 2q, 2q+1
  const int slot = mt * 16 + 8 * half + g;
  const int u = cta_ubase + slot;
  const bool u_ok = (slot < UPC) && (u < H);
  const int wslot0 = mt * 16 + 8 * half;  // first slot of this warp's 8-unit group
  const bool warp_ok = wslot0 < UPC;      // the last tile may be half empty (MT odd)
  int rb[2], rd[2];
  bool rok[2];
  float msk[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = cl * kRows + 2 * q + i;
    rok[i] = (r < nrows) && u_ok;
    const int rr = r < nrows ? r : 0;
    rd[i] = (rr >= B) ? 1 : 0;
    rb[i] = rr - rd[i] * B;
    msk[i] = a.mask ? (rok[i] ? __ldg(a.mask + static_cast<long long>(rr) * H + u) : 0.f)
                    : a.mask_scalar;
  }
  const float s = a.gscale ? __ldg(a.gscale) : 1.f;
  const float inv_s = 1.f / s;

  float carry[2] = {0.f, 0.f};
  float* mypre0 = &sm.pre[0][threadIdx.x][0];
  float* mypre1 = &sm.pre[1][threadIdx.x][0];
  // prefetch of step k's operands into pre[k & 1]
  auto prefetch_step = [&](int k) {
    float* dst = (k & 1) ? mypre1 : mypre0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (rok[i] && do_load) {
        const int t = rd[i] ? (T - 1 - k) : k;
        const long long col = static_cast<long long>(t) * B + rb[i];
        const long long idx = static_cast<long long>(rd[i] * H + u) * a.ldt + col;
        cp_async_f32(dst + i, a.dYT + idx);
        cp_async_f32(dst + 2 + i, a.ZT + idx);
        cp_async_f32(dst + 4 + i, a.HCT + idx);
        if (k > 0) cp_async_f32(dst + 6 + i, a.HT + idx + (rd[i] ? B : -B));
      }
    }
    cp_async_commit();
  };
  prefetch_step(T - 1);

  const uint32_t ldm_off = static_cast<uint32_t>(((lane & 7) * GS + 8 * (lane >> 3)) * 2);
  const uint32_t ldm_off2 = static_cast<uint32_t>(((lane & 7) * GS + 8 * ((lane >> 3) & 1)) * 2);
  const uint32_t g16_base = smem_u32(&sm.g16[0][0][0]);
  constexpr uint32_t kBufBytes = kRows * GS * 2;
  const long long gate_stride = static_cast<long long>(H) * a.ldt;         // da -> dpz rows
  const long long dir_stride = 2 * gate_stride;                            // direction blocks of GT

  for (int k = T - 1; k >= 0; --k) {
    const int buf = k & 1;
    if (SY && threadIdx.x == 0) mbar_arrive_expect_tx(&sm.mbar[buf], kTxBytes);
    // ---------------- phase A: pointwise backward of step k ----------------
    cp_async_wait_all();
    const float* pre = buf ? mypre1 : mypre0;
    float da[2], dpz[2], keep[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float zz = pre[2 + i], hc = pre[4 + i];
      const float hp = (k > 0) ? pre[6 + i] : 0.f;
      const float dh = pre[i] + carry[i];
      const float dzv = dh * (hp - hc);
      const float dhc = dh * (1.f - zz);
      const float m = msk[i];
      const float y = (m != 0.f) ? hc / m : 0.f;
      float dav = dhc * m * act_bwd_from_out(a.act, y);
      float dpzv = dzv * zz * (1.f - zz);
      if (!rok[i]) { dav = 0.f; dpzv = 0.f; }
      da[i] = dav; dpz[i] = dpzv;
      keep[i] = dh * zz;
      if (warp_ok) {
        if (SY) {
          sm.stage[warp][0][2 * q + i][g] = f16_sat(dav * s);
          sm.stage[warp][1][2 * q + i][g] = f16_sat(dpzv * s);
        } else {
          sm.g16[buf][2 * q + i][slot + cta_ubase] = f16_sat(dav * s);
          sm.g16[buf][2 * q + i][KP + slot + cta_ubase] = f16_sat(dpzv * s);
        }
      }
    }
    __syncwarp();
    if (warp_ok) {
      // 16 chunks (8 rows x 2 gates) of 16 bytes, each to every peer: lane -> chunk (lane&15), peer parity (lane>>4)
      const int n = lane & 7;
      const int gate = (lane >> 3) & 1;
      const uint32_t laddr = smem_u32(&sm.g16[buf][n][gate * KP + cta_ubase + wslot0]);
      if (SY) {
        const uint4 val = *reinterpret_cast<const uint4*>(&sm.stage[warp][gate][n][0]);
        const uint32_t lbar = smem_u32(&sm.mbar[buf]);
#pragma unroll
        for (int dst = (lane >> 4); dst < CL; dst += 2)
          st_async_v4(mapa_shared(laddr, dst), val, mapa_shared(lbar, dst));
      } else {
        const uint4 val = *reinterpret_cast<const uint4*>(&sm.g16[buf][n][gate * KP + cta_ubase + wslot0]);
#pragma unroll
        for (int dst = (lane >> 4); dst < CL; dst += 2)
          if (dst != static_cast<int>(crank)) st_cluster_v4(mapa_shared(laddr, dst), val);
      }
    }
    if (!SY) cluster_arrive_release();
    // prefetch of step k-1 + global stores of this step while the exchange completes
    if (k > 0) prefetch_step(k - 1);
    if (do_store) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (rok[i]) {
          const int t = rd[i] ? (T - 1 - k) : k;
          const long long col = static_cast<long long>(t) * B + rb[i];
          const long long idx = rd[i] * dir_stride + static_cast<long long>(u) * a.ldt + col;
          a.GT[idx] = da[i];
          a.GT[idx + gate_stride] = dpz[i];
          if (a.GT16) {
            a.GT16[idx] = f16_sat(da[i] * s);
            a.GT16[idx + gate_stride] = f16_sat(dpz[i] * s);
          }
        }
      }
    }
    if (SY) step_wait<SY>(&sm.mbar[buf], ((T - 1 - k) >> 1) & 1);
    else cluster_wait_acquire();

    // ---------------- phase B: U^T [da; dpz] for the carry into step k-1 ----------------
    if (k > 0) {
      float acc[4][4];
#pragma unroll
      for (int c = 0; c < 4; ++c) { acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f; }
      const uint32_t bufa = g16_base + buf * kBufBytes + half * (KP * 2);
#pragma unroll
      for (int kt = 0; kt + 1 < KT; kt += 2) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(bufa + ldm_off + kt * 32, b0, b1, b2, b3);
        mma_m16n8k16_f16(acc[kt & 3], A[kt], b0, b1);
        mma_m16n8k16_f16(acc[(kt + 1) & 3], A[kt + 1], b2, b3);
      }
      if (KT & 1) {
        uint32_t b0, b1;
        ldmatrix_x2(bufa + ldm_off2 + (KT - 1) * 32, b0, b1);
        mma_m16n8k16_f16(acc[(KT - 1) & 3], A[KT - 1], b0, b1);
      }
      float c4[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) c4[e] = (acc[0][e] + acc[1][e]) + (acc[2][e] + acc[3][e]);
      // warp half 0 keeps units g (c4[0..1]) and ships c4[2..3]; half 1 keeps g+8 and ships c4[0..1]
      sm.xbuf[buf][mt][half][lane][0] = half ? c4[0] : c4[2];
      sm.xbuf[buf][mt][half][lane][1] = half ? c4[1] : c4[3];
      asm volatile("bar.sync %0, 64;" ::"r"(mt + 1) : "memory");
      const float o0 = sm.xbuf[buf][mt][half ^ 1][lane][0];
      const float o1 = sm.xbuf[buf][mt][half ^ 1][lane][1];
      const float m0 = half ? c4[2] : c4[0];
      const float m1 = half ? c4[3] : c4[1];
      carry[0] = keep[0] + (m0 + o0) * inv_s;
      carry[1] = keep[1] + (m1 + o1) * inv_s;
    }
  }
  cp_async_wait_all();
  cluster_sync_all();
}

template <typename Args, void (*Kern)(const Args)>
int launch_rec(const Args& a, int cluster, int nclusters, int threads, size_t smem, cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [&] {
    err = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (err == cudaSuccess && cluster > 8)
      err = cudaFuncSetAttribute(Kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
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
  PK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, Kern, a));
  return 0;
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
int pick_cluster(int requested, int H) {
  if (requested > 0) return requested;
  static int env = env_int("PK_REC_CLUSTER", 0);  // tuning knob for bring-up
  if (env > 0) return env;
  // measured on B200 (profiles/): 10 CTAs x 7 warps keeps the 35 k-tile weight slice in registers
  // without spills (<= 8 warps/CTA -> 255 regs) and gave the shortest step
  return H > 512 ? 10 : 8;
}
int pick_sync(int requested) {  // 1 = st.async + mbarrier (default), 0 = barrier.cluster
  if (requested == 0 || requested == 1) return requested;
  static int env = env_int("PK_REC_SYNC", 1);
  return env ? 1 : 0;
}

#define PK_FWD(KT, MT, CL, SY)                                                                             \
  return launch_rec<RecFwdArgs, ligru_fwd_kernel<KT, MT, CL, SY>>(a, CL, nclusters, (MT) * 32,               \
                                                                  sizeof(FwdSmem<KT, MT, CL>), stream)
#define PK_BWD(KT, MT, CL, SY)                                                                             \
  return launch_rec<RecBwdArgs, ligru_bwd_kernel<KT, MT, CL, SY>>(a, CL, nclusters, (((MT) + 1) / 2) * 64,   \
                                                                  sizeof(BwdSmem<KT, MT, CL>), stream)

// (k-tiles, 8-unit tiles per CTA, cluster size) per hidden-size class; CL * 8 * MT >= H
#define PK_DISPATCH(MACRO)                                                                     \
  if (sy == 0) { /* barrier.cluster variants: cluster of 8 only */                             \
    if (H <= 256) { MACRO(16, 4, 8, 0); }                                                      \
    if (H <= 384) { MACRO(24, 6, 8, 0); }                                                      \
    if (H <= 512) { MACRO(32, 8, 8, 0); }                                                      \
    MACRO(35, 9, 8, 0);                                                                        \
  }                                                                                            \
  if (H <= 256) { MACRO(16, 4, 8, 1); }                                                        \
  if (H <= 384) { MACRO(24, 6, 8, 1); }                                                        \
  if (H <= 512) { MACRO(32, 8, 8, 1); }                                                        \
  if (cl == 9) { MACRO(35, 8, 9, 1); }                                                         \
  if (cl == 10) { MACRO(35, 7, 10, 1); }                                                       \
  if (cl == 12) { MACRO(35, 6, 12, 1); }                                                       \
  if (cl == 16) { MACRO(35, 5, 16, 1); }                                                       \
  MACRO(35, 9, 8, 1);

int fwd_dispatch(const RecFwdArgs& a, int cl, int sy, int nclusters, cudaStream_t stream) {
  const int H = a.H;
  PK_DISPATCH(PK_FWD)
}
int bwd_dispatch(const RecBwdArgs& a, int cl, int sy, int nclusters, cudaStream_t stream) {
  const int H = a.H;
  PK_DISPATCH(PK_BWD)
}

}  // namespace

int ligru_fwd(const RecFwdArgs& a, cudaStream_t stream) {
  PK_REQUIRE(a.T > 0 && a.B > 0 && a.H > 0, "ligru_fwd: empty problem");
  PK_REQUIRE(a.ndir == 1 || a.ndir == 2, "ligru_fwd: ndir must be 1 or 2");
  // 0 = auto: the faster kernel for this hidden size as measured on B200 (profiles/): register-stationary mma.sync
  // (ws) up to H = 560, tcgen05 with TMEM-stationary weights beyond; 3 = tcgen05, 2 = ws, 1 = legacy (pk_rnn.cu)
  static const int env_mode = env_int("PK_REC_MODE", 0);
  int mode = a.legacy ? a.legacy : env_mode;
  if (mode == 0) mode = a.H <= 560 ? 2 : 3;
  if (mode == 3 && a.cluster == 0 && a.sync != 0) return ligru_fwd_tc(a, stream);
  PK_REQUIRE(a.H <= 560, "ligru_fwd: hidden size %d > 560 not supported by the register-resident kernel", a.H);
  if (mode != 1 && a.cluster == 0 && a.sync != 0) return ligru_fwd_ws(a, stream);
  const int nclusters = (a.ndir * a.B + kRows - 1) / kRows;
  return fwd_dispatch(a, pick_cluster(a.cluster, a.H), pick_sync(a.sync), nclusters, stream);
}

int ligru_bwd(const RecBwdArgs& a, cudaStream_t stream) {
  PK_REQUIRE(a.T > 0 && a.B > 0 && a.H > 0, "ligru_bwd: empty problem");
  PK_REQUIRE(a.ndir == 1 || a.ndir == 2, "ligru_bwd: ndir must be 1 or 2");
  static const int env_mode = env_int("PK_REC_MODE", 0);
  int mode = a.legacy ? a.legacy : env_mode;
  if (mode == 0) mode = a.H <= 560 ? 2 : 3;
  if (mode == 3 && a.cluster == 0 && a.sync != 0) return ligru_bwd_tc(a, stream);
  PK_REQUIRE(a.H <= 560, "ligru_bwd: hidden size %d > 560 not supported by the register-resident kernel", a.H);
  if (mode != 1 && a.cluster == 0 && a.sync != 0) {
    PK_REQUIRE(a.GT16 != nullptr, "ligru_bwd: the warp-specialised kernel writes GT16 (required)");
    return ligru_bwd_ws(a, stream);
  }
  PK_REQUIRE(a.GT != nullptr, "ligru_bwd: legacy kernels need the fp32 GT buffer");
  const int nclusters = (a.ndir * a.B + kRows - 1) / kRows;
  return bwd_dispatch(a, pick_cluster(a.cluster, a.H), pick_sync(a.sync), nclusters, stream);
}

}  // namespace pk
