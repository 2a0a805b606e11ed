// pk_rnn_ws.cu — warp-specialised persistent liGRU kernels on mma.sync (default for H <= 560).
//
// Cluster of CL CTAs x 8 batch rows, recurrent weights stationary in registers as mma.sync fragments, asynchronous
// DSMEM exchange completed on the receiver's mbarrier (the non-specialised first version, pk_rnn.cu, was removed in
// round 2), with the refinements driven by the ncu profiles under profiles/:
//
//  1. WARP SPECIALISATION.  The per-step global-memory work no longer runs on the warps that sit on the
//     serial critical path:
//       compute warps (2 warpgroups, setmaxnreg.inc 224): wait -> ldmatrix + HMMA -> gates -> push;
//                     step inputs are read from / outputs written to small shared-memory RINGS
//       I/O warps (1 warpgroup, setmaxnreg.dec 56): global --cp.async--> in-ring (RI-1 steps ahead),
//                     out-ring --> global
//     (before: HMMA ~21 % and peer-wait ~9 % of a warp's step; the rest was address arithmetic, 12 scattered
//      stores and prefetch issue per thread per step)
//  2. TILE-MAJOR STAGING.  The staged fp16 vector is stored as [8-unit tile][row][8 units]: every ldmatrix
//     8x8 matrix is 128 contiguous bytes (conflict-free without padding) and a warp's tile rows are the
//     16-byte st.async messages.  (One 128-byte cp.async.bulk per (warp, peer) was measured too: it needs a
//     proxy fence and the compiler serialises the uniform-operand UBLKCPs -> 780 vs ~250 cycles per step.)
//  3. TRANSIT SHADOW.  Ring hand-offs (this step's outputs, next step's inputs) happen after the push,
//     while the data is in flight; gate math uses MUFU-based sigmoid/tanh (phase clocks: 668 -> ~150 cycles).
#include "pk_common.cuh"
#include "pk_kernels.h"

#include <cstdlib>
#include <mutex>

namespace pk {

namespace {

constexpr int kRows = 8;
constexpr int RI = 4;  // input ring depth (prefetch distance + 1)
constexpr int RO = 4;  // output ring depth
constexpr int kComputeWarps = 8;  // two warpgroups
constexpr int kIoWarps = 4;       // one warpgroup
constexpr int kThreads = (kComputeWarps + kIoWarps) * 32;
constexpr int NIO = kIoWarps * 32;

__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// the mbarrier receives one arrival when all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 128-byte tile: local shared memory -> shared memory of CTA `dst` (+ complete_tx on ITS mbarrier)
__device__ __forceinline__ void bulk_push_128(uint32_t dst_cluster_addr, uint32_t src_local_addr,
                                              uint32_t dst_cluster_mbar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];" ::"r"(
          dst_cluster_addr),
      "r"(src_local_addr), "r"(dst_cluster_mbar)
      : "memory");
}

// =====================================================================================
// forward
// =====================================================================================
template <int KT, int MT, int CL>
struct FwdWs {
  static constexpr int NT = CL * MT;  // 8-unit tiles of the whole layer
  static constexpr int UPC = 8 * MT;
  __half h16[2][NT][kRows][8];          // tile-major staged state, double buffered by step parity
  __half stage[MT][kRows][8];           // per-warp 8x8 tile, re-read as 16-byte rows for the push
  float inr[RI][2][UPC][kRows];         // [slot][gate h,z][unit][row]
  float outr[RO][3][UPC][kRows];        // [slot][h, z, hc][unit][row]
  uint64_t step_bar[2];
  uint64_t in_full[RI], in_empty[RI], out_full[RO], out_empty[RO];
};

template <int KT, int MT, int CL, int ACT>  // ACT: a per-element activation switch is an indirect branch (BRX) on the serial path
__global__ void __launch_bounds__(kThreads, 1) ligru_fwd_ws_kernel(const RecFwdArgs a) {
  using S = FwdWs<KT, MT, CL>;
  constexpr int NT = S::NT, UPC = S::UPC;
  static_assert(8 * NT >= 16 * KT, "unit tiles must cover the K range");
  static_assert(MT <= kComputeWarps, "one compute warp per 8-unit tile");
  constexpr uint32_t kTxBytes = NT * 128;  // every CTA receives the whole staged vector each step
  extern __shared__ __align__(128) uint8_t smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int cl = blockIdx.x / CL;
  const int H = a.H, B = a.B, T = a.T;
  const int nrows = a.ndir * B;
  const int cta_ubase = crank * UPC;

  for (int i = threadIdx.x; i < 2 * NT * kRows * 8 / 2; i += blockDim.x)
    reinterpret_cast<uint32_t*>(&sm.h16[0][0][0][0])[i] = 0u;
  for (int i = threadIdx.x; i < RI * 2 * UPC * kRows; i += blockDim.x) (&sm.inr[0][0][0][0])[i] = 0.f;
  if (threadIdx.x == 0) {
    mbar_init(&sm.step_bar[0], 1);
    mbar_init(&sm.step_bar[1], 1);
    for (int s = 0; s < RI; ++s) { mbar_init(&sm.in_full[s], NIO); mbar_init(&sm.in_empty[s], MT); }
    for (int s = 0; s < RO; ++s) { mbar_init(&sm.out_full[s], MT); mbar_init(&sm.out_empty[s], kIoWarps); }
    fence_mbar_init();
  }
  __syncthreads();
  cluster_sync_all();

  if (warp < kComputeWarps) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;" ::: "memory");
    if (warp < MT) {
      // ================= compute warps =================
      const int g = lane >> 2, q = lane & 3;
      const int ul = warp * 8 + g;  // local unit
      const int u = cta_ubase + ul;
      const bool u_ok = u < H;
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
      bool rok[2];
      float msk[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = cl * kRows + 2 * q + i;
        rok[i] = (r < nrows) && u_ok;
        msk[i] = a.mask ? (rok[i] ? __ldg(a.mask + static_cast<long long>(r) * H + u) : 0.f) : a.mask_scalar;
      }
      float sc_h = 0.f, sh_h = 0.f, sc_z = 0.f, sh_z = 0.f;
      if (u_ok) {
        sc_h = __ldg(a.scale + u); sh_h = __ldg(a.shift + u);
        sc_z = __ldg(a.scale + H + u); sh_z = __ldg(a.shift + H + u);
      }
      float hprev[2] = {0.f, 0.f};
      // ldmatrix: matrix m (lane>>3) = tile 2*kt + m, row lane&7, 16 bytes per row
      const uint32_t ldm_off = static_cast<uint32_t>(((lane >> 3) * 8 + (lane & 7)) * 16);
      const uint32_t ldm_off2 = static_cast<uint32_t>((((lane >> 3) & 1) * 8 + (lane & 7)) * 16);
      const uint32_t h16_base = smem_u32(&sm.h16[0][0][0][0]);
      constexpr uint32_t kBufBytes = NT * kRows * 16;
      const int tg = crank * MT + warp;  // global tile id of this warp's 8 units
      const bool z0 = a.force_z0 != 0;

      const bool clk_on = a.dbg_clk != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
      long long tsum[6] = {0, 0, 0, 0, 0, 0};
      // projections of step 0 (later steps are fetched in the shadow of the DSMEM transit)
      mbar_wait(&sm.in_full[0], 0);
      float2 ph = *reinterpret_cast<const float2*>(&sm.inr[0][0][ul][2 * q]);
      float2 pz = *reinterpret_cast<const float2*>(&sm.inr[0][1][ul][2 * q]);
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.in_empty[0]);
      for (int k = 0; k < T; ++k) {
        const int cur = k & 1, nxt = cur ^ 1;
        long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
        if (clk_on) t0 = clock64();
        if (k > 0) mbar_wait(&sm.step_bar[cur], ((k - 1) >> 1) & 1);
        if (threadIdx.x == 0) mbar_arrive_expect_tx(&sm.step_bar[nxt], kTxBytes);
        if (clk_on) t1 = clock64();
        float acc[4][4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f; }
        const uint32_t bufa = h16_base + cur * kBufBytes;
#pragma unroll
        for (int kt = 0; kt + 1 < KT; kt += 2) {
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4(bufa + ldm_off + kt * 256, b0, b1, b2, b3);  // 2 tiles (256 bytes) per k-tile
          mma_m16n8k16_f16(acc[kt & 3], A[kt], b0, b1);
          mma_m16n8k16_f16(acc[(kt + 1) & 3], A[kt + 1], b2, b3);
        }
        if (KT & 1) {
          uint32_t b0, b1;
          ldmatrix_x2(bufa + ldm_off2 + (KT - 1) * 256, b0, b1);
          mma_m16n8k16_f16(acc[(KT - 1) & 3], A[KT - 1], b0, b1);
        }
        const float ch0 = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
        const float ch1 = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
        const float cz0 = (acc[0][2] + acc[1][2]) + (acc[2][2] + acc[3][2]);
        const float cz1 = (acc[0][3] + acc[1][3]) + (acc[2][3] + acc[3][3]);
        if (clk_on) { t2 = clock64(); if (cz1 == 123.456f) t2 = 0; }

        // ---- gates (reference :1133-1136)
        float hn[2], zz[2], hcv[2];
        {
          const float zt = z0 ? 0.f : sigmoid_fast(fmaf(sc_z, pz.x, sh_z) + cz0);
          const float hc = act_fwd_fast(ACT, fmaf(sc_h, ph.x, sh_h) + ch0) * msk[0];
          float h = fmaf(zt, hprev[0] - hc, hc);
          if (!rok[0]) h = 0.f;
          hn[0] = h; zz[0] = zt; hcv[0] = hc; hprev[0] = h;
        }
        {
          const float zt = z0 ? 0.f : sigmoid_fast(fmaf(sc_z, pz.y, sh_z) + cz1);
          const float hc = act_fwd_fast(ACT, fmaf(sc_h, ph.y, sh_h) + ch1) * msk[1];
          float h = fmaf(zt, hprev[1] - hc, hc);
          if (!rok[1]) h = 0.f;
          hn[1] = h; zz[1] = zt; hcv[1] = hc; hprev[1] = h;
        }
        if (clk_on) { t3 = clock64(); if (hn[1] == 123.456f) t3 = 0; }
        // ---- stage the warp's 8x8 fp16 tile and push its 8 rows (16 bytes each) to every CTA: data and
        //      completion (complete_tx on the receiver's mbarrier) travel in one st.async message
        sm.stage[warp][2 * q][g] = f16_sat(hn[0]);
        sm.stage[warp][2 * q + 1][g] = f16_sat(hn[1]);
        __syncwarp();
        {
          const int n = lane & 7;
          const uint4 val = *reinterpret_cast<const uint4*>(&sm.stage[warp][n][0]);
          const uint32_t laddr = smem_u32(&sm.h16[nxt][tg][n][0]);
          const uint32_t lbar = smem_u32(&sm.step_bar[nxt]);
#pragma unroll
          for (int dst = (lane >> 3); dst < CL; dst += 4)
            st_async_v4(mapa_shared(laddr, dst), val, mapa_shared(lbar, dst));
        }
        if (clk_on) t4 = clock64();
        // ---- in the shadow of the DSMEM transit: outputs -> I/O warps, next step's projections <- ring
        const int so = k % RO;
        if (k >= RO) mbar_wait(&sm.out_empty[so], ((k / RO) - 1) & 1);
        *reinterpret_cast<float2*>(&sm.outr[so][0][ul][2 * q]) = make_float2(hn[0], hn[1]);
        *reinterpret_cast<float2*>(&sm.outr[so][1][ul][2 * q]) = make_float2(zz[0], zz[1]);
        *reinterpret_cast<float2*>(&sm.outr[so][2][ul][2 * q]) = make_float2(hcv[0], hcv[1]);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.out_full[so]);
        if (k + 1 < T) {
          const int si = (k + 1) % RI;
          mbar_wait(&sm.in_full[si], ((k + 1) / RI) & 1);
          ph = *reinterpret_cast<const float2*>(&sm.inr[si][0][ul][2 * q]);
          pz = *reinterpret_cast<const float2*>(&sm.inr[si][1][ul][2 * q]);
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.in_empty[si]);
        }
        if (clk_on) {
          const long long t5 = clock64();
          tsum[0] += t1 - t0; tsum[1] += t2 - t1; tsum[2] += t3 - t2; tsum[3] += t4 - t3; tsum[4] += t5 - t4;
        }
      }
      if (clk_on)
        for (int i = 0; i < 6; ++i) a.dbg_clk[i] = tsum[i];
      mbar_wait(&sm.step_bar[T & 1], ((T - 1) >> 1) & 1);  // drain the last incoming fill
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
    // ================= I/O warps (128 threads) =================
    // element e = tid + 128*j -> (unit = e / 8, row = e % 8); per-element bookkeeping is fixed over time
    const int tid = threadIdx.x - kComputeWarps * 32;
    constexpr int NE = (UPC * kRows + NIO - 1) / NIO;
    int colv[NE], cstep[NE];
    long long chan[NE], pch[NE];
    float hp[NE];
#pragma unroll
    for (int j = 0; j < NE; ++j) {
      const int e = tid + NIO * j;
      const int ul = e >> 3, r = e & 7;
      const int u = cta_ubase + ul;
      const int rr = cl * kRows + r;
      const bool ok = (e < UPC * kRows) && (u < H) && (rr < nrows);
      const int d = (ok && rr >= B) ? 1 : 0;
      const int b = rr - d * B;
      colv[j] = ok ? (d ? (T - 1) * B + b : b) : -1;  // column at step 0
      cstep[j] = d ? -B : B;
      chan[j] = static_cast<long long>(d * H + u) * a.ldt;
      pch[j] = static_cast<long long>(u) * a.ldp;
      hp[j] = 0.f;
    }
    const long long gate_z = static_cast<long long>(H) * a.ldp;
    const bool do_store = !(a.dbg & 1);
    const bool do_load = !(a.dbg & 2);
    // fast path: groups of 4 consecutive rows are contiguous and 16-byte aligned in every channel-major
    // array when B % 4 == 0 -> one thread moves (unit, 4 rows) with 16-byte cp.async / float4 stores
    const bool vec = (B % 4 == 0) && (a.ldp % 4 == 0) && (a.ldt % 4 == 0);
    const int vul = tid >> 1, vr0 = (tid & 1) * 4;  // this thread's (unit, first row) in the fast path
    const int vu = cta_ubase + vul;
    const int vrr = cl * kRows + vr0;
    const bool vok = (vul < UPC) && (vu < H) && (vrr < nrows);
    const int vd = (vok && vrr >= B) ? 1 : 0;
    const int vb = vrr - vd * B;
    const int vcstep = vd ? -B : B;
    const long long vcol0 = vd ? static_cast<long long>(T - 1) * B + vb : vb;
    const long long vchan = static_cast<long long>(vd * H + vu) * a.ldt;
    const long long vpch = static_cast<long long>(vu) * a.ldp;
    float4 vhp = make_float4(0.f, 0.f, 0.f, 0.f);

    auto issue_load = [&](int kl) {  // projections of step kl -> in-ring
      const int s = kl % RI;
      if (kl >= RI) mbar_wait(&sm.in_empty[s], ((kl / RI) - 1) & 1);
      if (do_load && vec) {
        if (vok) {
          const long long col = vcol0 + static_cast<long long>(kl) * vcstep;
          cp_async_16(&sm.inr[s][0][vul][vr0], a.PT + vpch + col);
          cp_async_16(&sm.inr[s][1][vul][vr0], a.PT + vpch + gate_z + col);
        }
      } else if (do_load) {
#pragma unroll
        for (int j = 0; j < NE; ++j) {
          if (colv[j] >= 0) {
            const int e = tid + NIO * j;
            const long long col = colv[j] + static_cast<long long>(kl) * cstep[j];
            cp_async_f32(&sm.inr[s][0][e >> 3][e & 7], a.PT + pch[j] + col);
            cp_async_f32(&sm.inr[s][1][e >> 3][e & 7], a.PT + pch[j] + gate_z + col);
          }
        }
      }
      cp_async_arrive_noinc(&sm.in_full[s]);
    };
    for (int kl = 0; kl < RI - 1 && kl < T; ++kl) issue_load(kl);
    for (int k = 0; k < T; ++k) {
      if (k + RI - 1 < T) issue_load(k + RI - 1);
      const int s = k % RO;
      mbar_wait(&sm.out_full[s], (k / RO) & 1);
      if (do_store && vec) {
        if (vok) {
          const long long idx = vchan + vcol0 + static_cast<long long>(k) * vcstep;
          const float4 h4 = *reinterpret_cast<const float4*>(&sm.outr[s][0][vul][vr0]);
          if (a.HT) *reinterpret_cast<float4*>(a.HT + idx) = h4;
          if (a.ZT) *reinterpret_cast<float4*>(a.ZT + idx) = *reinterpret_cast<const float4*>(&sm.outr[s][1][vul][vr0]);
          if (a.HCT) *reinterpret_cast<float4*>(a.HCT + idx) = *reinterpret_cast<const float4*>(&sm.outr[s][2][vul][vr0]);
          if (a.HT16) {
            uint2 pk16;
            pk16.x = pack_f16x2_sat(h4.x, h4.y);
            pk16.y = pack_f16x2_sat(h4.z, h4.w);
            *reinterpret_cast<uint2*>(a.HT16 + idx) = pk16;
          }
          if (a.HP16) {
            uint2 pk16;
            pk16.x = pack_f16x2_sat(vhp.x, vhp.y);
            pk16.y = pack_f16x2_sat(vhp.z, vhp.w);
            *reinterpret_cast<uint2*>(a.HP16 + idx) = pk16;
          }
          vhp = h4;
          if (a.Y32 || a.Y16) {
            const long long col = vcol0 + static_cast<long long>(k) * vcstep;
            const float hv[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (a.Y32) a.Y32[(col + i) * a.ldy32 + vd * H + vu] = hv[i];
              if (a.Y16) a.Y16[(col + i) * a.ldy16 + vd * H + vu] = f16_sat(hv[i]);
            }
          }
        }
      } else if (do_store) {
#pragma unroll
        for (int j = 0; j < NE; ++j) {
          if (colv[j] >= 0) {
            const int e = tid + NIO * j;
            const float h = sm.outr[s][0][e >> 3][e & 7];
            const long long idx = chan[j] + colv[j] + static_cast<long long>(k) * cstep[j];
            if (a.HT) a.HT[idx] = h;
            if (a.ZT) a.ZT[idx] = sm.outr[s][1][e >> 3][e & 7];
            if (a.HCT) a.HCT[idx] = sm.outr[s][2][e >> 3][e & 7];
            if (a.HT16) a.HT16[idx] = f16_sat(h);
            if (a.HP16) a.HP16[idx] = f16_sat(hp[j]);
            hp[j] = h;
          }
        }
        // row-major module output: for a fixed row the CTA's units are contiguous -> threads run along units
        if (a.Y32 || a.Y16) {
          for (int f = tid; f < UPC * kRows; f += NIO) {
            const int r = f / UPC, ul = f - r * UPC;
            const int rr = cl * kRows + r;
            const int u = cta_ubase + ul;
            if (rr < nrows && u < H) {
              const int d = rr >= B ? 1 : 0;
              const int b = rr - d * B;
              const long long col = static_cast<long long>(d ? T - 1 - k : k) * B + b;
              const float h = sm.outr[s][0][ul][r];
              if (a.Y32) a.Y32[col * a.ldy32 + d * H + u] = h;
              if (a.Y16) a.Y16[col * a.ldy16 + d * H + u] = f16_sat(h);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.out_empty[s]);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  // no CTA may exit while peers can still write into its shared memory
  cluster_sync_all();
}

// =====================================================================================
// backward
// =====================================================================================
// I/O warps of the backward kernels (128 threads): global --cp.async--> in-ring (RI-1 steps ahead), out-ring --> global.
template <class S, int UPC>
__device__ __forceinline__ void bwd_io_warps(S& sm, const RecBwdArgs& a, int cta_ubase, int cl, int lane) {
  const int H = a.H, B = a.B, T = a.T;
  const int nrows = a.ndir * B;
  // ================= I/O warps (128 threads) =================
  const int tid = threadIdx.x - kComputeWarps * 32;
  constexpr int NE = (UPC * kRows + NIO - 1) / NIO;
  int colv[NE], cstep[NE];
  long long chan[NE];
#pragma unroll
  for (int j = 0; j < NE; ++j) {
    const int e = tid + NIO * j;
    const int ul = e >> 3, r = e & 7;
    const int u = cta_ubase + ul;
    const int rr = cl * kRows + r;
    const bool ok = (e < UPC * kRows) && (u < H) && (rr < nrows);
    const int d = (ok && rr >= B) ? 1 : 0;
    const int b = rr - d * B;
    colv[j] = ok ? (d ? (T - 1) * B + b : b) : -1;  // column at step index 0
    cstep[j] = d ? -B : B;
    chan[j] = static_cast<long long>(d * H + u) * a.ldt;
  }
  const long long gate_stride = static_cast<long long>(H) * a.ldt;
  const long long dir_stride = 2 * gate_stride;
  const bool do_store = !(a.dbg & 1);
  const bool do_load = !(a.dbg & 2);
  const bool vec = (B % 4 == 0) && (a.ldt % 4 == 0);  // see the forward kernel
  const int vul = tid >> 1, vr0 = (tid & 1) * 4;
  const int vu = cta_ubase + vul;
  const int vrr = cl * kRows + vr0;
  const bool vok = (vul < UPC) && (vu < H) && (vrr < nrows);
  const int vd = (vok && vrr >= B) ? 1 : 0;
  const int vb = vrr - vd * B;
  const int vcstep = vd ? -B : B;
  const long long vcol0 = vd ? static_cast<long long>(T - 1) * B + vb : vb;
  const long long vchan = static_cast<long long>(vd * H + vu) * a.ldt;

  auto issue_load = [&](int it) {  // operands of step k = T-1-it -> in-ring slot it % RI
    const int k = T - 1 - it;
    const int s = it % RI;
    if (it >= RI) mbar_wait(&sm.in_empty[s], ((it / RI) - 1) & 1);
    if (do_load && vec) {
      if (vok) {
        const long long idx = vchan + vcol0 + static_cast<long long>(k) * vcstep;
        cp_async_16(&sm.inr[s][0][vul][vr0], a.dYT + idx);
        cp_async_16(&sm.inr[s][1][vul][vr0], a.ZT + idx);
        cp_async_16(&sm.inr[s][2][vul][vr0], a.HCT + idx);
        if (k > 0) cp_async_16(&sm.inr[s][3][vul][vr0], a.HT + idx - vcstep);
      }
    } else if (do_load) {
#pragma unroll
      for (int j = 0; j < NE; ++j) {
        if (colv[j] >= 0) {
          const int e = tid + NIO * j;
          const long long idx = chan[j] + colv[j] + static_cast<long long>(k) * cstep[j];
          cp_async_f32(&sm.inr[s][0][e >> 3][e & 7], a.dYT + idx);
          cp_async_f32(&sm.inr[s][1][e >> 3][e & 7], a.ZT + idx);
          cp_async_f32(&sm.inr[s][2][e >> 3][e & 7], a.HCT + idx);
          if (k > 0) cp_async_f32(&sm.inr[s][3][e >> 3][e & 7], a.HT + idx - cstep[j]);
        }
      }
    }
    cp_async_arrive_noinc(&sm.in_full[s]);
  };
  for (int it = 0; it < RI - 1 && it < T; ++it) issue_load(it);
  for (int it = 0; it < T; ++it) {
    if (it + RI - 1 < T) issue_load(it + RI - 1);
    const int k = T - 1 - it;
    const int s = it % RO;
    mbar_wait(&sm.out_full[s], (it / RO) & 1);
    if (do_store && vec) {
      if (vok) {
        const long long idx = vd * dir_stride + static_cast<long long>(vu) * a.ldt + vcol0 +
                              static_cast<long long>(k) * vcstep;
        *reinterpret_cast<uint2*>(a.GT16 + idx) = *reinterpret_cast<const uint2*>(&sm.outr[s][0][vul][vr0]);
        *reinterpret_cast<uint2*>(a.GT16 + idx + gate_stride) =
            *reinterpret_cast<const uint2*>(&sm.outr[s][1][vul][vr0]);
      }
    } else if (do_store) {
#pragma unroll
      for (int j = 0; j < NE; ++j) {
        if (colv[j] >= 0) {
          const int e = tid + NIO * j;
          const int ul = e >> 3, r = e & 7;
          const int u = cta_ubase + ul;
          const int d = cstep[j] < 0 ? 1 : 0;
          const long long col = colv[j] + static_cast<long long>(k) * cstep[j];
          const long long idx = d * dir_stride + static_cast<long long>(u) * a.ldt + col;
          a.GT16[idx] = sm.outr[s][0][ul][r];
          a.GT16[idx + gate_stride] = sm.outr[s][1][ul][r];
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.out_empty[s]);
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
}

template <int KT, int MT, int CL>
struct BwdWs {
  static constexpr int MT16 = (MT + 1) / 2;
  static constexpr int NWC = MT16 * 2;  // active compute warps
  static constexpr int NT = CL * MT;
  static constexpr int UPC = 8 * MT;
  __half g16[2][2][NT][kRows][8];          // [buffer][gate da,dpz][tile][row][8 units]
  __half stage[NWC][2][kRows][8];
  float xbuf[2][MT16][2][32][2];
  float inr[RI][4][UPC][kRows];            // [slot][dy, z, hc, hprev][unit][row]
  __half outr[RO][2][UPC][kRows];          // [slot][da, dpz][unit][row]  (scaled fp16)
  uint64_t step_bar[2];
  uint64_t in_full[RI], in_empty[RI], out_full[RO], out_empty[RO];
};

template <int KT, int MT, int CL, int ACT>
__global__ void __launch_bounds__(kThreads, 1) ligru_bwd_ws_kernel(const RecBwdArgs a) {
  using S = BwdWs<KT, MT, CL>;
  constexpr int MT16 = S::MT16, NWC = S::NWC, NT = S::NT, UPC = S::UPC;
  static_assert(8 * NT >= 16 * KT, "unit tiles must cover the K range");
  static_assert(NWC <= kComputeWarps, "too many unit tiles per CTA");
  constexpr uint32_t kTxBytes = 2 * NT * 128;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int cl = blockIdx.x / CL;
  const int H = a.H, B = a.B, T = a.T;
  const int nrows = a.ndir * B;
  const int cta_ubase = crank * UPC;

  for (int i = threadIdx.x; i < 2 * 2 * NT * kRows * 8 / 2; i += blockDim.x)
    reinterpret_cast<uint32_t*>(&sm.g16[0][0][0][0][0])[i] = 0u;
  for (int i = threadIdx.x; i < RI * 4 * UPC * kRows; i += blockDim.x) (&sm.inr[0][0][0][0])[i] = 0.f;
  if (threadIdx.x == 0) {
    mbar_init(&sm.step_bar[0], 1);
    mbar_init(&sm.step_bar[1], 1);
    for (int s = 0; s < RI; ++s) { mbar_init(&sm.in_full[s], NIO); mbar_init(&sm.in_empty[s], NWC); }
    for (int s = 0; s < RO; ++s) { mbar_init(&sm.out_full[s], NWC); mbar_init(&sm.out_empty[s], kIoWarps); }
    fence_mbar_init();
  }
  __syncthreads();
  cluster_sync_all();

  if (warp < kComputeWarps) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;" ::: "memory");
    if (warp < NWC) {
      // ================= compute warps =================
      const int g = lane >> 2, q = lane & 3;
      const int mt = warp >> 1, half = warp & 1;
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
      const int slot = mt * 16 + 8 * half + g;  // element ownership: unit slot, rows 2q, 2q+1
      const int u = cta_ubase + slot;
      const bool u_ok = (slot < UPC) && (u < H);
      const int wtile = 2 * mt + half;          // local 8-unit tile of this warp's elements
      const bool warp_ok = wtile < MT;          // the last 16-unit tile may be half empty (MT odd)
      const int sl = warp_ok ? slot : 0;
      const int tg = crank * MT + wtile;        // global tile id
      bool rok[2];
      float msk[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = cl * kRows + 2 * q + i;
        rok[i] = (r < nrows) && u_ok;
        msk[i] = a.mask ? (rok[i] ? __ldg(a.mask + static_cast<long long>(r) * H + u) : 0.f) : a.mask_scalar;
      }
      const float s = a.gscale ? __ldg(a.gscale) : 1.f;
      const float inv_s = 1.f / s;
      float carry[2] = {0.f, 0.f};
      const uint32_t ldm_off = static_cast<uint32_t>(((lane >> 3) * 8 + (lane & 7)) * 16);
      const uint32_t ldm_off2 = static_cast<uint32_t>((((lane >> 3) & 1) * 8 + (lane & 7)) * 16);
      const uint32_t g16_base = smem_u32(&sm.g16[0][0][0][0][0]);
      constexpr uint32_t kGateBytes = NT * kRows * 16;
      constexpr uint32_t kBufBytes = 2 * kGateBytes;

      const bool clk_on = a.dbg_clk != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
      long long tsum[6] = {0, 0, 0, 0, 0, 0};
      // operands of the first processed step (later ones are fetched in the shadow of the transit)
      mbar_wait(&sm.in_full[0], 0);
      float2 dy = *reinterpret_cast<const float2*>(&sm.inr[0][0][sl][2 * q]);
      float2 zz = *reinterpret_cast<const float2*>(&sm.inr[0][1][sl][2 * q]);
      float2 hc = *reinterpret_cast<const float2*>(&sm.inr[0][2][sl][2 * q]);
      float2 hp = *reinterpret_cast<const float2*>(&sm.inr[0][3][sl][2 * q]);
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.in_empty[0]);
      // masks are 0/1 in training; eval uses the scalar (1-p): y = hc / m via one reciprocal
      const float rm0 = (msk[0] != 0.f) ? __frcp_rn(msk[0]) : 0.f;
      const float rm1 = (msk[1] != 0.f) ? __frcp_rn(msk[1]) : 0.f;
      for (int k = T - 1; k >= 0; --k) {
        const int it = T - 1 - k;
        const int buf = k & 1;
        long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
        if (clk_on) t0 = clock64();
        if (threadIdx.x == 0) mbar_arrive_expect_tx(&sm.step_bar[buf], kTxBytes);
        // ---------------- phase A: pointwise backward of step k ----------------
        if (k == 0) hp = make_float2(0.f, 0.f);
        float keep[2];
        __half2 da16, dz16;
        {
          const float dh0 = dy.x + carry[0], dh1 = dy.y + carry[1];
          float da0 = dh0 * (1.f - zz.x) * msk[0] * act_bwd_from_out(ACT, hc.x * rm0);
          float da1 = dh1 * (1.f - zz.y) * msk[1] * act_bwd_from_out(ACT, hc.y * rm1);
          float dz0 = dh0 * (hp.x - hc.x) * zz.x * (1.f - zz.x);
          float dz1 = dh1 * (hp.y - hc.y) * zz.y * (1.f - zz.y);
          if (!rok[0]) { da0 = 0.f; dz0 = 0.f; }
          if (!rok[1]) { da1 = 0.f; dz1 = 0.f; }
          keep[0] = dh0 * zz.x;
          keep[1] = dh1 * zz.y;
          da16 = __halves2half2(f16_sat(da0 * s), f16_sat(da1 * s));
          dz16 = __halves2half2(f16_sat(dz0 * s), f16_sat(dz1 * s));
        }
        if (clk_on) t1 = clock64();
        if (warp_ok) {
          sm.stage[warp][0][2 * q][g] = __low2half(da16);
          sm.stage[warp][0][2 * q + 1][g] = __high2half(da16);
          sm.stage[warp][1][2 * q][g] = __low2half(dz16);
          sm.stage[warp][1][2 * q + 1][g] = __high2half(dz16);
        }
        __syncwarp();
        if (warp_ok) {  // 16 rows of 16 bytes (8 rows x 2 gates), each to every CTA of the cluster
          const int n = lane & 7;
          const int gate = (lane >> 3) & 1;
          const uint4 val = *reinterpret_cast<const uint4*>(&sm.stage[warp][gate][n][0]);
          const uint32_t laddr = smem_u32(&sm.g16[buf][gate][tg][n][0]);
          const uint32_t lbar = smem_u32(&sm.step_bar[buf]);
#pragma unroll
          for (int dst = (lane >> 4); dst < CL; dst += 2)
            st_async_v4(mapa_shared(laddr, dst), val, mapa_shared(lbar, dst));
        }
        if (clk_on) t2 = clock64();
        // ---- in the shadow of the transit: outputs -> I/O warps, next step's operands <- ring
        const int so = it % RO;
        if (it >= RO) mbar_wait(&sm.out_empty[so], ((it / RO) - 1) & 1);
        if (warp_ok) {
          *reinterpret_cast<__half2*>(&sm.outr[so][0][slot][2 * q]) = da16;
          *reinterpret_cast<__half2*>(&sm.outr[so][1][slot][2 * q]) = dz16;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.out_full[so]);
        if (k > 0) {
          const int si = (it + 1) % RI;
          mbar_wait(&sm.in_full[si], ((it + 1) / RI) & 1);
          dy = *reinterpret_cast<const float2*>(&sm.inr[si][0][sl][2 * q]);
          zz = *reinterpret_cast<const float2*>(&sm.inr[si][1][sl][2 * q]);
          hc = *reinterpret_cast<const float2*>(&sm.inr[si][2][sl][2 * q]);
          hp = *reinterpret_cast<const float2*>(&sm.inr[si][3][sl][2 * q]);
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.in_empty[si]);
        }
        if (clk_on) t3 = clock64();
        mbar_wait(&sm.step_bar[buf], (it >> 1) & 1);
        if (clk_on) t4 = clock64();

        // ---------------- phase B: U^T [da; dpz] for the carry into step k-1 ----------------
        if (k > 0) {
          float acc[4][4];
#pragma unroll
          for (int c = 0; c < 4; ++c) { acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f; }
          const uint32_t bufa = g16_base + buf * kBufBytes + half * kGateBytes;
#pragma unroll
          for (int kt = 0; kt + 1 < KT; kt += 2) {
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4(bufa + ldm_off + kt * 256, b0, b1, b2, b3);
            mma_m16n8k16_f16(acc[kt & 3], A[kt], b0, b1);
            mma_m16n8k16_f16(acc[(kt + 1) & 3], A[kt + 1], b2, b3);
          }
          if (KT & 1) {
            uint32_t b0, b1;
            ldmatrix_x2(bufa + ldm_off2 + (KT - 1) * 256, b0, b1);
            mma_m16n8k16_f16(acc[(KT - 1) & 3], A[KT - 1], b0, b1);
          }
          float c4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) c4[e] = (acc[0][e] + acc[1][e]) + (acc[2][e] + acc[3][e]);
          sm.xbuf[buf][mt][half][lane][0] = half ? c4[0] : c4[2];
          sm.xbuf[buf][mt][half][lane][1] = half ? c4[1] : c4[3];
          asm volatile("bar.sync %0, 64;" ::"r"(mt + 1) : "memory");
          const float o0 = sm.xbuf[buf][mt][half ^ 1][lane][0];
          const float o1 = sm.xbuf[buf][mt][half ^ 1][lane][1];
          carry[0] = keep[0] + ((half ? c4[2] : c4[0]) + o0) * inv_s;
          carry[1] = keep[1] + ((half ? c4[3] : c4[1]) + o1) * inv_s;
        }
        if (clk_on && k > 0) {
          const long long t5 = clock64();
          tsum[0] += t1 - t0; tsum[1] += t2 - t1; tsum[2] += t3 - t2; tsum[3] += t4 - t3; tsum[4] += t5 - t4;
        }
      }
      if (clk_on)
        for (int i = 0; i < 6; ++i) a.dbg_clk[i] = tsum[i];
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
    bwd_io_warps<S, UPC>(sm, a, cta_ubase, cl, lane);
  }
  cluster_sync_all();
}

#include "pk_rnn_ks.inc"

template <typename Args, void (*Kern)(const Args)>
int launch_ws(const Args& a, int cluster, int nclusters, size_t smem, cudaStream_t stream) {
  static PerDeviceOnce once;
  const cudaError_t err = once.run([&] {
    cudaError_t e = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess && cluster > 8) e = cudaFuncSetAttribute(Kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    return e;
  });
  PK_CHECK_CUDA(err);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters * cluster, 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
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

#define PK_WS_ACT(ARGS, KERN, SMEM, KT, MT, CL)                                                              \
  switch (a.act) {                                                                                         \
    case ACT_RELU: return launch_ws<ARGS, KERN<KT, MT, CL, ACT_RELU>>(a, CL, nclusters, SMEM, stream);     \
    case ACT_TANH: return launch_ws<ARGS, KERN<KT, MT, CL, ACT_TANH>>(a, CL, nclusters, SMEM, stream);     \
    case ACT_SIGMOID: return launch_ws<ARGS, KERN<KT, MT, CL, ACT_SIGMOID>>(a, CL, nclusters, SMEM, stream); \
    case ACT_LEAKY_RELU: return launch_ws<ARGS, KERN<KT, MT, CL, ACT_LEAKY_RELU>>(a, CL, nclusters, SMEM, stream); \
    case ACT_ELU: return launch_ws<ARGS, KERN<KT, MT, CL, ACT_ELU>>(a, CL, nclusters, SMEM, stream);       \
    default: return launch_ws<ARGS, KERN<KT, MT, CL, ACT_LINEAR>>(a, CL, nclusters, SMEM, stream);         \
  }
#define PK_FWD_WS(KT, MT, CL) PK_WS_ACT(RecFwdArgs, ligru_fwd_ws_kernel, sizeof(FwdWs<KT, MT, CL>) + 128, KT, MT, CL)
#define PK_BWD_WS(KT, MT, CL) PK_WS_ACT(RecBwdArgs, ligru_bwd_ws_kernel, sizeof(BwdWs<KT, MT, CL>) + 128, KT, MT, CL)

#define PK_KS_ACT(MT, CL)                                                                                    \
  {                                                                                                          \
    constexpr size_t smem = sizeof(BwdKs<MT, CL>) + 128;                                                       \
    switch (a.act) {                                                                                         \
      case ACT_RELU: return launch_ws<RecBwdArgs, ligru_bwd_ks_kernel<MT, CL, ACT_RELU>>(a, CL, nclusters, smem, stream); \
      case ACT_TANH: return launch_ws<RecBwdArgs, ligru_bwd_ks_kernel<MT, CL, ACT_TANH>>(a, CL, nclusters, smem, stream); \
      case ACT_SIGMOID: return launch_ws<RecBwdArgs, ligru_bwd_ks_kernel<MT, CL, ACT_SIGMOID>>(a, CL, nclusters, smem, stream); \
      case ACT_LEAKY_RELU: return launch_ws<RecBwdArgs, ligru_bwd_ks_kernel<MT, CL, ACT_LEAKY_RELU>>(a, CL, nclusters, smem, stream); \
      case ACT_ELU: return launch_ws<RecBwdArgs, ligru_bwd_ks_kernel<MT, CL, ACT_ELU>>(a, CL, nclusters, smem, stream); \
      default: return launch_ws<RecBwdArgs, ligru_bwd_ks_kernel<MT, CL, ACT_LINEAR>>(a, CL, nclusters, smem, stream); \
    }                                                                                                        \
  }

long long* g_dbg_clk = nullptr;

}  // namespace

void set_debug_clock_buffer(long long* dev_ptr) {
  g_dbg_clk = dev_ptr;
  set_debug_clock_buffer_tc(dev_ptr);
}

// (k-tiles, 8-unit tiles per CTA, cluster size): CL * 8 * MT >= 16 * KT >= H.  Measured (round 2): a 14-CTA cluster with 5
// tiles per CTA (shorter HMMA chain per SM) is 1.8x SLOWER (1.89 vs 1.04 us per step) — large clusters lose on the
// DSMEM exchange / placement, so 10 x 7 tiles stays.
int ligru_fwd_ws(const RecFwdArgs& a_in, cudaStream_t stream) {
  RecFwdArgs a = a_in;
  a.dbg_clk = g_dbg_clk;
  const int nclusters = (a.ndir * a.B + kRows - 1) / kRows;
  const int H = a.H;
  if (H <= 256) { PK_FWD_WS(16, 4, 8); }
  if (H <= 384) { PK_FWD_WS(24, 6, 8); }
  if (H <= 512) { PK_FWD_WS(32, 7, 10); }
  PK_FWD_WS(35, 7, 10);
}
int ligru_bwd_ws(const RecBwdArgs& a_in, cudaStream_t stream) {
  RecBwdArgs a = a_in;
  a.dbg_clk = g_dbg_clk;
  const int nclusters = (a.ndir * a.B + kRows - 1) / kRows;
  const int H = a.H;
  // formulation of the exchange: dbg bit 4 forces the all-gather kernel, bit 5 the K-split one; default K-split
  // (1.20 vs 1.33 us per step at H = 550, profiles/r2_selftest_ksplit.log), PK_BWD_KS=0 in the environment flips it
  static const int env_ks = [] { const char* e = getenv("PK_BWD_KS"); return e ? atoi(e) : 1; }();
  const bool ksplit = (a.dbg & 32) ? true : ((a.dbg & 16) ? false : env_ks != 0);
  if (ksplit) {
    if (H <= 256) PK_KS_ACT(4, 8)
    if (H <= 384) PK_KS_ACT(6, 8)
    PK_KS_ACT(7, 10)
  }
  if (H <= 256) { PK_BWD_WS(16, 4, 8); }
  if (H <= 384) { PK_BWD_WS(24, 6, 8); }
  if (H <= 512) { PK_BWD_WS(32, 7, 10); }
  PK_BWD_WS(35, 7, 10);
}

}  // namespace pk
