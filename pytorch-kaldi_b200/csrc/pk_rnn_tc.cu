// pk_rnn_tc.cu — persistent liGRU / RNN recurrence on tcgen05 tensor cores (round-2 default).
//
// Replaces the `for k in range(T)` loop of the reference (neural_networks.py:1130-1141 liGRU, :1438-1447 RNN), its
// flip / stack / cat shuffles (:1095-1097, :1144-1150) and, in reverse time, the autograd tape behind it.
//
// Decomposition (measured numbers: profiles/r2_microbench.txt):
//   * a thread-block CLUSTER owns 8 of the 2B independent batch rows; CTA c of the cluster owns hidden units
//     [64c, 64c+64) for both gates -> M = 128 gate rows per CTA, CL = ceil(H / 64) CTAs (9 for H = 550);
//   * WEIGHTS STATIONARY IN TENSOR MEMORY: the CTA's [128 x K] fp16 slice of U = [Uh; Uz] is written once into
//     TMEM (lane = gate row, two K-elements per 32-bit column, 32 columns per 64-unit chunk) and is the A operand
//     of `tcgen05.mma.cta_group::1.kind::f16` (TS form) for all T steps.  With A in shared memory every MMA
//     re-reads 4 KB of weights (57 cycles per instruction at N = 16, measured); from TMEM the same instruction
//     costs ~17 cycles;
//   * the fp16 state is the B operand: [16 rows x 64 units] per source CTA in the 128-byte-swizzled K-major
//     layout (rows 8..15 are zero padding: M = 128 needs N >= 16), double buffered by step parity.  After its gate
//     math each CTA st.async-pushes its 8 x 64 block (16-byte messages, complete_tx on the RECEIVER's per-source
//     mbarrier) to every CTA of the cluster;
//   * the MMA-issuing thread waits per SOURCE chunk (its own chunk first) and fires that chunk's 4 MMAs as soon
//     as it has landed, so the tensor work runs inside the DSMEM transit (the exchange is bandwidth-bound:
//     ~400 + bytes / 20.5 cycles per step, measured) and only the last chunk's MMAs are exposed;
//   * the fp32 accumulator [128 x 16] lives in TMEM (double buffered); four epilogue warps read it with
//     tcgen05.ld, swap half of their rows through shared memory so that every thread has BOTH gates of one unit
//     for 4 rows, and do sigmoid / activation / mask / convex update on fp32 state kept in registers;
//   * global memory is touched only by four I/O warps: projections (forward) or saved tensors (backward) are
//     prefetched into a shared-memory ring with 16-byte cp.async, outputs are drained from a ring in the
//     consumer layouts (channel-major fp32 saved tensors, fp16 operand copies, row-major module output).
//
// Backward (reverse time) keeps U^T stationary the same way: lanes 0..63 hold Uh[:, unit]^T, lanes 64..127
// Uz[:, unit]^T; the exchanged operand is [da; dpz] (rows 0..7 = da, rows 8..15 = dpz of the same 8 batch rows,
// so N = 16 carries no padding), and dh_{t-1} = keep + (lanes 0..63, columns 0..7) + (lanes 64..127, columns 8..15).
#include "pk_common.cuh"
#include "pk_kernels.h"

#include <cstdlib>
#include <mutex>

namespace pk {

namespace {

constexpr int kRows = 8;    // real batch rows per cluster
constexpr int kUPC = 64;    // hidden units per CTA
constexpr int kMaxCL = 16;  // H <= 1024
constexpr int kTmemCL = 15; // chunks whose weights fit tensor memory: 32 accumulator columns + 32 * 15 = 512; the
                            // 16th chunk (units 960..1023 as K) stays in shared memory and is multiplied in SS form
constexpr int RI = 4;       // input ring depth (prefetch distance + 1)
constexpr int RO = 4;       // output ring depth
constexpr int kIoWarps = 4;
constexpr int NIO = kIoWarps * 32;
constexpr int kEpiWarp0 = 2;  // warps 2..5: epilogue (TMEM lane group = warp % 4)
constexpr int kIoWarp0 = 6;   // warps 6..9: global-memory I/O
constexpr int kThreads = 320;
constexpr uint32_t kACol = 32;          // first TMEM column of the stationary weights
constexpr uint32_t kChunkBytes = 2048;  // 16 rows x 128 bytes: one source CTA's 64 units
constexpr uint32_t kIdesc = umma_idesc(0, 128, 16);
constexpr int kTrace0 = 200;  // bring-up: steps [kTrace0, kTrace0 + 8) of CTA 0 are time-stamped into dbg_clk[8..]
constexpr uint32_t kFwdTx = kRows * 128;      // bytes one source CTA delivers per step (8 rows x 64 units fp16)
constexpr uint32_t kBwdTx = 2 * kRows * 128;  // backward: two gates

__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// bounded waits: a protocol bug must trap (the launch fails with an error) instead of hanging the device
__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t parity) {
  uint32_t polls = 0;
  long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++polls & 0x3ffu) == 0) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000LL) __trap();
    }
  }
}
__device__ __forceinline__ void tc_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t polls = 0;
  long long t0 = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if ((++polls & 0x3ffu) == 0) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000LL) __trap();
    }
  }
}

// spinning variant (mbarrier.test_wait never suspends the thread): used by the MMA issuer, whose wake-up latency
// is on the serial critical path of every step
__device__ __forceinline__ void tc_spin_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t polls = 0;
  long long t0 = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if ((++polls & 0xfffu) == 0) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000LL) __trap();
    }
  }
}

// Source chunks are grouped by ARRIVAL ORDER.  Every CTA pushes to destinations crank, crank+1, ... (mod CL), so
// the block of source s reaches destination d in slot i = (d - s) mod CL; slots [g * gsz, (g+1) * gsz) share one
// mbarrier (kGroups barriers per buffer instead of one per source: a barrier hand-off costs ~100 cycles, a chunk's
// MMAs ~70).
constexpr int kDefaultGroups = 1;
constexpr int kGroups = 3;  // at most; the launch picks 1..3 (RecArgs::groups)
__device__ __forceinline__ long long gtime_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ int group_size(int CL, int ng) { return (CL + ng - 1) / ng; }

// ---- the MMA-issuing thread, one step: per arrival group: wait -> re-arm -> that group's MMAs (chunk order =
// arrival order); finally commit to acc_full.
template <bool TAIL>  // TAIL: H > 960, chunk 15's weights are multiplied from shared memory
__device__ __forceinline__ void mma_step(uint32_t opnd_addr, uint64_t* grp_bar_buf, uint64_t* acc_full, uint32_t tmem_base,
                                         uint32_t acc_col, int CL, int gsz, bool tail4, uint32_t crank, bool wait_data, uint32_t parity,
                                         uint32_t tx_bytes, bool proxy_fence, bool spin, uint32_t wtail_addr, long long* trace, long long* gt = nullptr) {
  if (trace) trace[0] = clock64();
  // Running operand addresses (chunk c = crank, crank-1, ... with wrap-around): per MMA only two uniform adds remain.
  // Shared-memory descriptor: low word = (address >> 4) | LBO field, high word constant (SBO 1024 B, version 1,
  // 128-byte swizzle) — advancing the address by x bytes adds x >> 4 to the low word.
  constexpr uint32_t kDescHi = static_cast<uint32_t>((1024u >> 4) | (1u << 14) | (2u << 29));
  const uint32_t d_tmem = tmem_base + acc_col;
  int c = static_cast<int>(crank);
  uint32_t a_addr = tmem_base + kACol + crank * 32u;
  uint32_t b_lo = (((opnd_addr + crank * kChunkBytes) & 0x3FFFFu) >> 4) | (1u << 16);
  uint32_t first = 0;  // the very first MMA of the step overwrites the accumulator
  int i = 0;
  for (int g = 0; i < CL; ++g) {
    const int i_end = min(CL, i + gsz);
    if (wait_data) {
      if (spin) tc_spin_cluster(&grp_bar_buf[g], parity); else tc_wait_cluster(&grp_bar_buf[g], parity);
      if (proxy_fence) fence_proxy_async_smem();
      tc_fence_after();
    }
    const uint32_t rearm = tx_bytes * static_cast<uint32_t>(i_end - i);
    if (trace) trace[g == 0 ? 1 : 2] = clock64();
    if (gt) gt[2] = gtime_ns();
    for (; i < i_end; ++i) {
      const uint64_t hi = static_cast<uint64_t>(kDescHi) << 32;
      if (TAIL && c == kTmemCL) {  // H > 960: this chunk's weights live in shared memory (SS form, ~57 cycles per MMA)
        const uint32_t w_lo = ((wtail_addr & 0x3FFFFu) >> 4) | (1u << 16);
        umma_f16(d_tmem, hi | w_lo, hi | b_lo, kIdesc, first);
        umma_f16(d_tmem, hi | (w_lo + 2), hi | (b_lo + 2), kIdesc, 1u);
        umma_f16(d_tmem, hi | (w_lo + 4), hi | (b_lo + 4), kIdesc, 1u);
        umma_f16(d_tmem, hi | (w_lo + 6), hi | (b_lo + 6), kIdesc, 1u);
        first = 1u;
        --c;
        a_addr -= 32u;
        b_lo -= (kChunkBytes >> 4);
        continue;
      }
      // branch-free body (a branch per MMA serialises the descriptor chains); k-steps beyond H multiply zero weights
      umma_f16_ts(d_tmem, a_addr, hi | b_lo, kIdesc, first);
      umma_f16_ts(d_tmem, a_addr + 8, hi | (b_lo + 2), kIdesc, 1u);
      umma_f16_ts(d_tmem, a_addr + 16, hi | (b_lo + 4), kIdesc, 1u);
      if (c != CL - 1 || tail4) umma_f16_ts(d_tmem, a_addr + 24, hi | (b_lo + 6), kIdesc, 1u);  // last chunk: k-steps beyond H are zero
      first = 1u;
      if (c == 0) {
        c = CL - 1;
        a_addr += static_cast<uint32_t>(CL - 1) * 32u;
        b_lo += static_cast<uint32_t>(CL - 1) * (kChunkBytes >> 4);
      } else {
        --c;
        a_addr -= 32u;
        b_lo -= (kChunkBytes >> 4);
      }
    }
    if (wait_data) mbar_arrive_expect_tx(&grp_bar_buf[g], rearm);  // re-arm for this buffer's next use (two steps on)
  }
  umma_commit(acc_full);
  if (trace) trace[3] = clock64();
}

// =====================================================================================
// forward
// =====================================================================================
struct FwdTc {
  uint8_t hbuf[2][kMaxCL][kChunkBytes];  // B operand: [buffer][source CTA][16 rows x 128 B, 128B-swizzled]
  uint8_t wtail[128 * 128];              // A tile of chunk 15 (H > 960 only): [128 gate rows x 64 K] fp16, 128B-swizzled
  float inr[RI][2][kUPC][kRows];         // [slot][gate h,z][unit][row]
  float outr[RO][3][kUPC][kRows];        // [slot][h, z, hc][unit][row]
  uint8_t stage[kRows * 128];            // image of this CTA's block of the operand (8 swizzled rows of 128 B)
  uint64_t src_bar[2][kGroups];          // [buffer][arrival group]
  uint64_t acc_full[2];
  uint64_t in_full[RI], in_empty[RI], out_full[RO], out_empty[RO];
  uint32_t tmem_slot;
};

template <int ACT, bool TAIL>  // activation id: the per-element switch would be an indirect branch (BRX) on the serial path
__global__ void __launch_bounds__(kThreads, 1) ligru_fwd_tc_kernel(const RecFwdArgs a, const int CL) {
  extern __shared__ uint8_t smem_raw[];
  FwdTc& sm = *reinterpret_cast<FwdTc*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int cl = blockIdx.x / CL;
  const int H = a.H, B = a.B, T = a.T;
  const int nrows = a.ndir * B;
  const int cta_ubase = crank * kUPC;
  const int gsz = group_size(CL, a.groups);  // source chunks per arrival-group barrier

  for (int i = threadIdx.x; i < static_cast<int>(sizeof(sm.hbuf) / 16); i += blockDim.x)
    reinterpret_cast<uint4*>(&sm.hbuf[0][0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < RI * 2 * kUPC * kRows; i += blockDim.x) (&sm.inr[0][0][0][0])[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      for (int g = 0; g < kGroups; ++g) mbar_init(&sm.src_bar[b][g], 1);
      mbar_init(&sm.acc_full[b], 1);
    }
    for (int s = 0; s < RI; ++s) { mbar_init(&sm.in_full[s], NIO); mbar_init(&sm.in_empty[s], 4); }
    for (int s = 0; s < RO; ++s) { mbar_init(&sm.out_full[s], 4); mbar_init(&sm.out_empty[s], kIoWarps); }
    fence_mbar_init();
    // arm every per-source barrier for its first use (buffer 1: step 1, buffer 0: step 2)
    for (int b = 0; b < 2; ++b)
      for (int g = 0; g * gsz < CL; ++g)
        mbar_arrive_expect_tx(&sm.src_bar[b][g], kFwdTx * static_cast<uint32_t>(min(CL, (g + 1) * gsz) - g * gsz));
  }
  fence_proxy_async_smem();  // the zero-filled operand buffers are read by the tensor core (async proxy) at step 0
  if (warp == 1) tmem_alloc<512>(&sm.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_slot;

  // ---- stationary weights -> tensor memory (epilogue warps).  Gate rows are interleaved per 32-lane group so that
  //      both gates of a unit live in ONE warp: lane 32g + l holds gate (l >> 4) of local unit 16g + (l & 15)
  if (warp >= kEpiWarp0 && warp < kIoWarp0) {
    const int lg = warp & 3;
    const int u = cta_ubase + lg * 16 + (lane & 15);
    const bool u_ok = u < H;
    const float* Urow = a.U + (static_cast<long long>(lane >> 4) * H + (u_ok ? u : 0)) * H;
    const int KP = min(CL, kTmemCL) * kUPC;
    if (TAIL) {  // K = 960..1023 of this thread's gate row -> the shared-memory A tile (row = TMEM lane)
      const int m = lg * 32 + lane;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = kTmemCL * kUPC + j * 8 + 2 * e;
          w[e] = pack_f16x2_sat((u_ok && k < H) ? __ldg(Urow + k) : 0.f, (u_ok && k + 1 < H) ? __ldg(Urow + k + 1) : 0.f);
        }
        *reinterpret_cast<uint4*>(&sm.wtail[(m >> 3) * 1024 + (m & 7) * 128 + ((j ^ (m & 7)) << 4)]) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      fence_proxy_async_smem();
    }
    for (int k0 = 0; k0 < KP; k0 += 16) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = k0 + 2 * j;
        const float x0 = (u_ok && k < H) ? __ldg(Urow + k) : 0.f;
        const float x1 = (u_ok && k + 1 < H) ? __ldg(Urow + k + 1) : 0.f;
        v[j] = pack_f16x2_sat(x0, x1);
      }
      tmem_st_32x8(tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + kACol + static_cast<uint32_t>(k0 >> 1), v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers and buffers are initialised before any peer pushes
  tc_fence_after();

  if (warp == 0) {
    // ================= MMA issuer =================
    if (elect_one()) {
      const bool pf = !(a.dbg & 4);
      for (int k = 0; k < T; ++k) {
        const int cur = k & 1;
        long long* tr = (a.dbg_clk && blockIdx.x == 0 && k >= kTrace0 && k < kTrace0 + 8) ? a.dbg_clk + 8 + (k - kTrace0) * 16 : nullptr;
        long long* gt = (a.dbg_clk && blockIdx.x < CL && k >= kTrace0 && k < kTrace0 + 8) ? a.dbg_clk + 136 + blockIdx.x * 64 + (k - kTrace0) * 8 : nullptr;
        mma_step<TAIL>(smem_u32(&sm.hbuf[cur][0][0]), &sm.src_bar[cur][0], &sm.acc_full[cur], tmem_base, static_cast<uint32_t>(cur * 16), CL,
                 gsz, (H - (CL - 1) * kUPC) > 48, crank, k > 0, static_cast<uint32_t>(((k - 1) >> 1) & 1), kFwdTx, pf, !(a.dbg & 8), smem_u32(&sm.wtail[0]), tr, gt);
        if (gt) gt[3] = gtime_ns();
      }
    }
    __syncwarp();
  } else if (warp >= kEpiWarp0 && warp < kIoWarp0) {
    // ================= epilogue warps: gates (each warp is self-contained: no block-level barrier) =================
    const int lg = warp & 3;           // TMEM lane group this warp may read
    const int ew = warp - kEpiWarp0;   // staging buffer of this warp
    const int half = lane >> 4;        // 0: this lane's TMEM row is the candidate gate, 1: the update gate
    const int ul = lg * 16 + (lane & 15);  // local unit
    const int u = cta_ubase + ul;
    const bool u_ok = u < H;
    const int r0 = half * 4;           // this thread finishes rows r0 .. r0+3 of its unit
    float msk[4];
    bool rok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = cl * kRows + r0 + i;
      rok[i] = (r < nrows) && u_ok;
      msk[i] = a.mask ? (rok[i] ? __ldg(a.mask + static_cast<long long>(r) * H + u) : 0.f) : a.mask_scalar;
    }
    float sc_h = 0.f, sh_h = 0.f, sc_z = 0.f, sh_z = 0.f;
    if (u_ok) {
      sc_h = __ldg(a.scale + u); sh_h = __ldg(a.shift + u);
      sc_z = __ldg(a.scale + H + u); sh_z = __ldg(a.shift + H + u);
    }
    float hprev[4] = {0.f, 0.f, 0.f, 0.f};
    const bool z0 = a.force_z0 != 0;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    // The block is staged as its final shared-memory image and pushed as WHOLE 128-byte rows: one st.async
    // instruction = 32 lanes x 16 B = 4 complete rows for one destination (scattered 16-byte pieces reach only
    // ~1/3 of the DSMEM bandwidth, measured).  Instruction q of the CTA (q < 2 CL): slot q >> 1, rows 4 (q & 1) ..;
    // warp ew issues q = ew, ew + 4, ...
    const uint32_t stage_off[4] = {static_cast<uint32_t>((r0 + 0) * 128 + ((((ul >> 3) ^ (r0 + 0)) & 7) << 4) + (ul & 7) * 2),
                                   static_cast<uint32_t>((r0 + 1) * 128 + ((((ul >> 3) ^ (r0 + 1)) & 7) << 4) + (ul & 7) * 2),
                                   static_cast<uint32_t>((r0 + 2) * 128 + ((((ul >> 3) ^ (r0 + 2)) & 7) << 4) + (ul & 7) * 2),
                                   static_cast<uint32_t>((r0 + 3) * 128 + ((((ul >> 3) ^ (r0 + 3)) & 7) << 4) + (ul & 7) * 2)};

    uint32_t gtab = 0;  // arrival group of slot i, 2 bits each (no division on the per-step path)
    for (int i = 0; i < CL; ++i) gtab |= static_cast<uint32_t>(i / gsz) << (2 * i);
    const bool clk_on = a.dbg_clk != nullptr && blockIdx.x == 0 && ew == 0 && lane == 0;
    const bool clk3 = a.dbg_clk != nullptr && blockIdx.x == 0 && ew == 3 && lane == 0;
    long long tsum[6] = {0, 0, 0, 0, 0, 0};
    tc_wait(&sm.in_full[0], 0);
    float4 ph = *reinterpret_cast<const float4*>(&sm.inr[0][0][ul][r0]);
    float4 pz = *reinterpret_cast<const float4*>(&sm.inr[0][1][ul][r0]);
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.in_empty[0]);
    for (int k = 0; k < T; ++k) {
      const int cur = k & 1, nxt = cur ^ 1;
      long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
      if (clk_on) t0 = clock64();
      tc_wait(&sm.acc_full[cur], static_cast<uint32_t>((k >> 1) & 1));
      tc_fence_after();
      if (clk_on) t1 = clock64();
      if (clk3 && k >= kTrace0 && k < kTrace0 + 8) a.dbg_clk[8 + (k - kTrace0) * 16 + 9] = clock64();
      uint32_t v[8];
      tmem_ld_32x8(lane_addr + static_cast<uint32_t>(cur * 16), v);
      tmem_ld_wait();
      tc_fence_before();
      // lane l and lane l ^ 16 hold the two gates of one unit: swap the rows the partner finishes
      float ah[4], az[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float mine = __uint_as_float(half ? v[4 + i] : v[i]);      // own gate, own rows
        const float give = __uint_as_float(half ? v[i] : v[4 + i]);      // own gate, partner's rows
        const float got = __shfl_xor_sync(0xffffffffu, give, 16);        // partner's gate, own rows
        ah[i] = half ? got : mine;
        az[i] = half ? mine : got;
      }
      if (clk_on) t2 = clock64();
      // ---- gates (reference :1133-1136)
      const float phv[4] = {ph.x, ph.y, ph.z, ph.w}, pzv[4] = {pz.x, pz.y, pz.z, pz.w};
      float hn[4], zz[4], hcv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float zt = z0 ? 0.f : sigmoid_fast(fmaf(sc_z, pzv[i], sh_z) + az[i]);
        const float hc = act_fwd_fast(ACT, fmaf(sc_h, phv[i], sh_h) + ah[i]) * msk[i];
        float h = fmaf(zt, hprev[i] - hc, hc);
        if (!rok[i]) h = 0.f;
        hn[i] = h; zz[i] = zt; hcv[i] = hc; hprev[i] = h;
        *reinterpret_cast<__half*>(&sm.stage[stage_off[i]]) = f16_sat(h);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps: the block image is complete
      if (clk_on) t3 = clock64();
      long long* gt = (a.dbg_clk && blockIdx.x < CL && ew == 0 && lane == 0 && k >= kTrace0 && k < kTrace0 + 8) ? a.dbg_clk + 136 + blockIdx.x * 64 + (k - kTrace0) * 8 : nullptr;
      if (gt) gt[0] = gtime_ns();
      // ---- push the warp's 8 x 16 block to every CTA of the cluster (data + complete_tx in one message)
      if (k + 1 < T) {
        const uint4 v0 = *reinterpret_cast<const uint4*>(&sm.stage[lane * 16]);        // rows 0..3
        const uint4 v1 = *reinterpret_cast<const uint4*>(&sm.stage[512 + lane * 16]);  // rows 4..7
        const uint32_t laddr = smem_u32(&sm.hbuf[nxt][0][0]) + crank * kChunkBytes + static_cast<uint32_t>(lane * 16);
        const uint32_t lbar0 = smem_u32(&sm.src_bar[nxt][0]);
        for (int q = ew; q < 2 * CL; q += 4) {
          const int i = q >> 1;  // slot i: destination crank + i -> lands in its arrival group
          int dst = static_cast<int>(crank) + i;
          if (dst >= CL) dst -= CL;
          st_async_v4(mapa_shared(laddr + static_cast<uint32_t>((q & 1) * 512), dst), (q & 1) ? v1 : v0,
                      mapa_shared(lbar0 + ((gtab >> (2 * i)) & 3u) * 8u, dst));
        }
      }
      if (clk_on) t4 = clock64();
      if (gt) gt[1] = gtime_ns();
      // ---- in the shadow of the transit: outputs -> I/O warps, next step's projections <- ring
      const int so = k % RO;
      if (k >= RO) tc_wait(&sm.out_empty[so], static_cast<uint32_t>(((k / RO) - 1) & 1));
      *reinterpret_cast<float4*>(&sm.outr[so][0][ul][r0]) = make_float4(hn[0], hn[1], hn[2], hn[3]);
      *reinterpret_cast<float4*>(&sm.outr[so][1][ul][r0]) = make_float4(zz[0], zz[1], zz[2], zz[3]);
      *reinterpret_cast<float4*>(&sm.outr[so][2][ul][r0]) = make_float4(hcv[0], hcv[1], hcv[2], hcv[3]);
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.out_full[so]);
      if (k + 1 < T) {
        const int si = (k + 1) % RI;
        tc_wait(&sm.in_full[si], static_cast<uint32_t>(((k + 1) / RI) & 1));
        ph = *reinterpret_cast<const float4*>(&sm.inr[si][0][ul][r0]);
        pz = *reinterpret_cast<const float4*>(&sm.inr[si][1][ul][r0]);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.in_empty[si]);
      }
      if (clk_on) {
        const long long t5 = clock64();
        tsum[0] += t1 - t0; tsum[1] += t2 - t1; tsum[2] += t3 - t2; tsum[3] += t4 - t3; tsum[4] += t5 - t4;
        if (k >= kTrace0 && k < kTrace0 + 8) {
          long long* tr = a.dbg_clk + 8 + (k - kTrace0) * 16;
          tr[4] = t1; tr[5] = t2; tr[6] = t3; tr[7] = t4; tr[8] = t5;
        }
      }
      if (clk3 && k >= kTrace0 && k < kTrace0 + 8) a.dbg_clk[8 + (k - kTrace0) * 16 + 10] = clock64();
    }
    if (clk_on)
      for (int i = 0; i < 6; ++i) a.dbg_clk[i] = tsum[i];
  } else if (warp >= kIoWarp0) {
    // ================= I/O warps (128 threads) =================
    const int tid = threadIdx.x - kIoWarp0 * 32;
    constexpr int NE = (kUPC * kRows + NIO - 1) / NIO;
    int colv[NE], cstep[NE];
    long long chan[NE], pch[NE];
    float hp[NE];
#pragma unroll
    for (int j = 0; j < NE; ++j) {
      const int e = tid + NIO * j;
      const int ul = e >> 3, r = e & 7;
      const int u = cta_ubase + ul;
      const int rr = cl * kRows + r;
      const bool ok = (u < H) && (rr < nrows);
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
    const int vul = tid >> 1, vr0 = (tid & 1) * 4;
    const int vu = cta_ubase + vul;
    const int vrr = cl * kRows + vr0;
    const bool vok = (vu < H) && (vrr < nrows);
    const int vd = (vok && vrr >= B) ? 1 : 0;
    const int vb = vrr - vd * B;
    const int vcstep = vd ? -B : B;
    const long long vcol0 = vd ? static_cast<long long>(T - 1) * B + vb : vb;
    const long long vchan = static_cast<long long>(vd * H + vu) * a.ldt;
    const long long vpch = static_cast<long long>(vu) * a.ldp;
    float4 vhp = make_float4(0.f, 0.f, 0.f, 0.f);

    auto issue_load = [&](int kl) {  // projections of step kl -> in-ring
      const int s = kl % RI;
      if (kl >= RI) tc_wait(&sm.in_empty[s], static_cast<uint32_t>(((kl / RI) - 1) & 1));
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
      tc_wait(&sm.out_full[s], static_cast<uint32_t>((k / RO) & 1));
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
          for (int f = tid; f < kUPC * kRows; f += NIO) {
            const int r = f / kUPC, ul = f - r * kUPC;
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
  // no CTA may exit (or free its tensor memory) while peers can still write into its shared memory
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// =====================================================================================
// backward
// =====================================================================================
struct BwdTc {
  uint8_t gbuf[2][kMaxCL][kChunkBytes];  // B operand: rows 0..7 = da, rows 8..15 = dpz of the cluster's 8 batch rows
  uint8_t wtail[128 * 128];              // A tile of chunk 15 (H > 960 only)
  float inr[RI][4][kUPC][kRows];         // [slot][dy, z, hc, hprev][unit][row]
  __half outr[RO][2][kUPC][kRows];       // [slot][da, dpz][unit][row]  (scaled fp16)
  uint8_t stage[2 * kRows * 128];        // image of this CTA's block: rows 0..7 = da, rows 8..15 = dpz (swizzled)
  uint64_t src_bar[2][kGroups];          // [buffer][arrival group]
  uint64_t acc_full[2];
  uint64_t in_full[RI], in_empty[RI], out_full[RO], out_empty[RO];
  uint32_t tmem_slot;
};

template <int ACT, bool TAIL>
__global__ void __launch_bounds__(kThreads, 1) ligru_bwd_tc_kernel(const RecBwdArgs a, const int CL) {
  extern __shared__ uint8_t smem_raw[];
  BwdTc& sm = *reinterpret_cast<BwdTc*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int cl = blockIdx.x / CL;
  const int H = a.H, B = a.B, T = a.T;
  const int nrows = a.ndir * B;
  const int cta_ubase = crank * kUPC;
  const int gsz = group_size(CL, a.groups);  // source chunks per arrival-group barrier

  for (int i = threadIdx.x; i < static_cast<int>(sizeof(sm.gbuf) / 16); i += blockDim.x)
    reinterpret_cast<uint4*>(&sm.gbuf[0][0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < RI * 4 * kUPC * kRows; i += blockDim.x) (&sm.inr[0][0][0][0])[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      for (int g = 0; g < kGroups; ++g) mbar_init(&sm.src_bar[b][g], 1);
      mbar_init(&sm.acc_full[b], 1);
    }
    for (int s = 0; s < RI; ++s) { mbar_init(&sm.in_full[s], NIO); mbar_init(&sm.in_empty[s], 4); }
    for (int s = 0; s < RO; ++s) { mbar_init(&sm.out_full[s], 4); mbar_init(&sm.out_empty[s], kIoWarps); }
    fence_mbar_init();
    for (int b = 0; b < 2; ++b)
      for (int g = 0; g * gsz < CL; ++g)
        mbar_arrive_expect_tx(&sm.src_bar[b][g], kBwdTx * static_cast<uint32_t>(min(CL, (g + 1) * gsz) - g * gsz));
  }
  fence_proxy_async_smem();
  if (warp == 1) tmem_alloc<512>(&sm.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_slot;

  // ---- stationary U^T -> tensor memory: lane 32g + l, unit = 16g + (l & 15): A[.][j] = Uh[j][unit] (l < 16) or
  //      Uz[j][unit] (l >= 16) — both halves of a unit's sum live in one warp
  if (warp >= kEpiWarp0 && warp < kIoWarp0) {
    const int lg = warp & 3;
    const int u = cta_ubase + lg * 16 + (lane & 15);
    const bool u_ok = u < H;
    const float* Ucol = a.U + static_cast<long long>(lane >> 4) * H * H + (u_ok ? u : 0);
    const int KP = min(CL, kTmemCL) * kUPC;
    if (TAIL) {
      const int m = lg * 32 + lane;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = kTmemCL * kUPC + j * 8 + 2 * e;
          w[e] = pack_f16x2_sat((u_ok && k < H) ? __ldg(Ucol + static_cast<long long>(k) * H) : 0.f,
                                (u_ok && k + 1 < H) ? __ldg(Ucol + static_cast<long long>(k + 1) * H) : 0.f);
        }
        *reinterpret_cast<uint4*>(&sm.wtail[(m >> 3) * 1024 + (m & 7) * 128 + ((j ^ (m & 7)) << 4)]) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      fence_proxy_async_smem();
    }
    for (int k0 = 0; k0 < KP; k0 += 16) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = k0 + 2 * j;
        const float x0 = (u_ok && k < H) ? __ldg(Ucol + static_cast<long long>(k) * H) : 0.f;
        const float x1 = (u_ok && k + 1 < H) ? __ldg(Ucol + static_cast<long long>(k + 1) * H) : 0.f;
        v[j] = pack_f16x2_sat(x0, x1);
      }
      tmem_st_32x8(tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + kACol + static_cast<uint32_t>(k0 >> 1), v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();

  if (warp == 0) {
    // ================= MMA issuer: one product per step k = T-1 .. 1 (the carry into step k-1) =================
    if (elect_one()) {
      const bool pf = !(a.dbg & 4);
      for (int k = T - 1; k > 0; --k) {
        const int it = T - 1 - k;
        const int buf = k & 1;
        mma_step<TAIL>(smem_u32(&sm.gbuf[buf][0][0]), &sm.src_bar[buf][0], &sm.acc_full[buf], tmem_base, static_cast<uint32_t>(buf * 16), CL,
                 gsz, (H - (CL - 1) * kUPC) > 48, crank, true, static_cast<uint32_t>((it >> 1) & 1), kBwdTx, pf, !(a.dbg & 8), smem_u32(&sm.wtail[0]), nullptr);
      }
    }
    __syncwarp();
  } else if (warp >= kEpiWarp0 && warp < kIoWarp0) {
    // ================= epilogue warps: pointwise backward (self-contained warps) =================
    const int lg = warp & 3;
    const int ew = warp - kEpiWarp0;
    const int half = lane >> 4;
    const int ul = lg * 16 + (lane & 15);
    const int u = cta_ubase + ul;
    const bool u_ok = u < H;
    const int r0 = half * 4;
    float msk[4], rmk[4];
    bool rok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = cl * kRows + r0 + i;
      rok[i] = (r < nrows) && u_ok;
      msk[i] = a.mask ? (rok[i] ? __ldg(a.mask + static_cast<long long>(r) * H + u) : 0.f) : a.mask_scalar;
      rmk[i] = (msk[i] != 0.f) ? __frcp_rn(msk[i]) : 0.f;  // masks are 0/1 in training; eval: scalar (1-p)
    }
    const float s = a.gscale ? __ldg(a.gscale) : 1.f;
    const float inv_s = 1.f / s;
    float carry[4] = {0.f, 0.f, 0.f, 0.f};
    // lanes < 16 own the Uh^T partial sums (valid in accumulator columns 0..7), lanes >= 16 the Uz^T ones (8..15)
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    // staged block image + whole-row pushes as in the forward kernel: instruction q (q < 4 CL): slot q >> 2,
    // 512-byte quarter q & 3; warp ew issues q = ew, ew + 4, ... i.e. quarter ew of every slot
    const uint32_t stage_off[4] = {static_cast<uint32_t>((r0 + 0) * 128 + ((((ul >> 3) ^ (r0 + 0)) & 7) << 4) + (ul & 7) * 2),
                                   static_cast<uint32_t>((r0 + 1) * 128 + ((((ul >> 3) ^ (r0 + 1)) & 7) << 4) + (ul & 7) * 2),
                                   static_cast<uint32_t>((r0 + 2) * 128 + ((((ul >> 3) ^ (r0 + 2)) & 7) << 4) + (ul & 7) * 2),
                                   static_cast<uint32_t>((r0 + 3) * 128 + ((((ul >> 3) ^ (r0 + 3)) & 7) << 4) + (ul & 7) * 2)};

    uint32_t gtab = 0;  // arrival group of slot i, 2 bits each (no division on the per-step path)
    for (int i = 0; i < CL; ++i) gtab |= static_cast<uint32_t>(i / gsz) << (2 * i);
    const bool clk_on = a.dbg_clk != nullptr && blockIdx.x == 0 && ew == 0 && lane == 0;
    long long tsum[6] = {0, 0, 0, 0, 0, 0};
    tc_wait(&sm.in_full[0], 0);
    float4 dy = *reinterpret_cast<const float4*>(&sm.inr[0][0][ul][r0]);
    float4 zz = *reinterpret_cast<const float4*>(&sm.inr[0][1][ul][r0]);
    float4 hc = *reinterpret_cast<const float4*>(&sm.inr[0][2][ul][r0]);
    float4 hp = *reinterpret_cast<const float4*>(&sm.inr[0][3][ul][r0]);
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.in_empty[0]);
    for (int k = T - 1; k >= 0; --k) {
      const int it = T - 1 - k;
      const int buf = k & 1;
      long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
      if (clk_on) t0 = clock64();
      // ---------------- phase A: pointwise backward of step k ----------------
      if (k == 0) hp = make_float4(0.f, 0.f, 0.f, 0.f);
      const float dyv[4] = {dy.x, dy.y, dy.z, dy.w}, zv[4] = {zz.x, zz.y, zz.z, zz.w};
      const float hcv[4] = {hc.x, hc.y, hc.z, hc.w}, hpv[4] = {hp.x, hp.y, hp.z, hp.w};
      float keep[4];
      __half da16[4], dz16[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float dh = dyv[i] + carry[i];
        float da = dh * (1.f - zv[i]) * msk[i] * act_bwd_from_out(ACT, hcv[i] * rmk[i]);
        float dz = dh * (hpv[i] - hcv[i]) * zv[i] * (1.f - zv[i]);
        if (!rok[i]) { da = 0.f; dz = 0.f; }
        keep[i] = dh * zv[i];
        da16[i] = f16_sat(da * s);
        dz16[i] = f16_sat(dz * s);
        *reinterpret_cast<__half*>(&sm.stage[stage_off[i]]) = da16[i];
        *reinterpret_cast<__half*>(&sm.stage[1024 + stage_off[i]]) = dz16[i];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (clk_on) t1 = clock64();
      if (k > 0) {
        const uint4 val = *reinterpret_cast<const uint4*>(&sm.stage[ew * 512 + lane * 16]);
        const uint32_t laddr = smem_u32(&sm.gbuf[buf][0][0]) + crank * kChunkBytes + static_cast<uint32_t>(ew * 512 + lane * 16);
        const uint32_t lbar0 = smem_u32(&sm.src_bar[buf][0]);
        int dst = static_cast<int>(crank);
        for (int i = 0; i < CL; ++i, ++dst) {
          if (dst >= CL) dst -= CL;
          st_async_v4(mapa_shared(laddr, dst), val, mapa_shared(lbar0 + ((gtab >> (2 * i)) & 3u) * 8u, dst));
        }
      }
      if (clk_on) t2 = clock64();
      // ---- in the shadow of the transit: outputs -> I/O warps, next step's operands <- ring
      const int so = it % RO;
      if (it >= RO) tc_wait(&sm.out_empty[so], static_cast<uint32_t>(((it / RO) - 1) & 1));
      {
        uint2 p0, p1;
        p0.x = (static_cast<uint32_t>(__half_as_ushort(da16[1])) << 16) | __half_as_ushort(da16[0]);
        p0.y = (static_cast<uint32_t>(__half_as_ushort(da16[3])) << 16) | __half_as_ushort(da16[2]);
        p1.x = (static_cast<uint32_t>(__half_as_ushort(dz16[1])) << 16) | __half_as_ushort(dz16[0]);
        p1.y = (static_cast<uint32_t>(__half_as_ushort(dz16[3])) << 16) | __half_as_ushort(dz16[2]);
        *reinterpret_cast<uint2*>(&sm.outr[so][0][ul][r0]) = p0;
        *reinterpret_cast<uint2*>(&sm.outr[so][1][ul][r0]) = p1;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.out_full[so]);
      if (k > 0) {
        const int si = (it + 1) % RI;
        tc_wait(&sm.in_full[si], static_cast<uint32_t>(((it + 1) / RI) & 1));
        dy = *reinterpret_cast<const float4*>(&sm.inr[si][0][ul][r0]);
        zz = *reinterpret_cast<const float4*>(&sm.inr[si][1][ul][r0]);
        hc = *reinterpret_cast<const float4*>(&sm.inr[si][2][ul][r0]);
        hp = *reinterpret_cast<const float4*>(&sm.inr[si][3][ul][r0]);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.in_empty[si]);
      }
      if (clk_on) t3 = clock64();
      // ---------------- phase B: carry into step k-1 = keep + U^T [da; dpz] / s ----------------
      if (k > 0) {
        tc_wait(&sm.acc_full[buf], static_cast<uint32_t>((it >> 1) & 1));
        tc_fence_after();
        if (clk_on) t4 = clock64();
        // tcgen05.ld is warp-collective (one address for the warp): fetch all 16 columns, each half keeps its 8
        uint32_t w[16];
        tmem_ld_32x16(lane_addr + static_cast<uint32_t>(buf * 16), w);
        tmem_ld_wait();
        tc_fence_before();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float mine = __uint_as_float(half ? w[12 + i] : w[i]);  // own partial sum, own rows
          const float give = __uint_as_float(half ? w[8 + i] : w[4 + i]);  // own partial sum, partner's rows
          const float got = __shfl_xor_sync(0xffffffffu, give, 16);     // partner's partial sum, own rows
          carry[i] = keep[i] + (mine + got) * inv_s;
        }
      }
      if (clk_on && k > 0) {
        const long long t5 = clock64();
        tsum[0] += t1 - t0; tsum[1] += t2 - t1; tsum[2] += t3 - t2; tsum[3] += t4 - t3; tsum[4] += t5 - t4;
      }
    }
    if (clk_on)
      for (int i = 0; i < 6; ++i) a.dbg_clk[i] = tsum[i];
  } else if (warp >= kIoWarp0) {
    // ================= I/O warps (128 threads) =================
    const int tid = threadIdx.x - kIoWarp0 * 32;
    constexpr int NE = (kUPC * kRows + NIO - 1) / NIO;
    int colv[NE], cstep[NE];
    long long chan[NE];
#pragma unroll
    for (int j = 0; j < NE; ++j) {
      const int e = tid + NIO * j;
      const int ul = e >> 3, r = e & 7;
      const int u = cta_ubase + ul;
      const int rr = cl * kRows + r;
      const bool ok = (u < H) && (rr < nrows);
      const int d = (ok && rr >= B) ? 1 : 0;
      const int b = rr - d * B;
      colv[j] = ok ? (d ? (T - 1) * B + b : b) : -1;
      cstep[j] = d ? -B : B;
      chan[j] = static_cast<long long>(d * H + u) * a.ldt;
    }
    const long long gate_stride = static_cast<long long>(H) * a.ldt;
    const long long dir_stride = 2 * gate_stride;
    const bool do_store = !(a.dbg & 1);
    const bool do_load = !(a.dbg & 2);
    const bool vec = (B % 4 == 0) && (a.ldt % 4 == 0);
    const int vul = tid >> 1, vr0 = (tid & 1) * 4;
    const int vu = cta_ubase + vul;
    const int vrr = cl * kRows + vr0;
    const bool vok = (vu < H) && (vrr < nrows);
    const int vd = (vok && vrr >= B) ? 1 : 0;
    const int vb = vrr - vd * B;
    const int vcstep = vd ? -B : B;
    const long long vcol0 = vd ? static_cast<long long>(T - 1) * B + vb : vb;
    const long long vchan = static_cast<long long>(vd * H + vu) * a.ldt;

    auto issue_load = [&](int it) {  // operands of step k = T-1-it -> in-ring slot it % RI
      const int k = T - 1 - it;
      const int s = it % RI;
      if (it >= RI) tc_wait(&sm.in_empty[s], static_cast<uint32_t>(((it / RI) - 1) & 1));
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
      tc_wait(&sm.out_full[s], static_cast<uint32_t>((it / RO) & 1));
      if (do_store && vec) {
        if (vok) {
          const long long idx = vd * dir_stride + static_cast<long long>(vu) * a.ldt + vcol0 + static_cast<long long>(k) * vcstep;
          *reinterpret_cast<uint2*>(a.GT16 + idx) = *reinterpret_cast<const uint2*>(&sm.outr[s][0][vul][vr0]);
          *reinterpret_cast<uint2*>(a.GT16 + idx + gate_stride) = *reinterpret_cast<const uint2*>(&sm.outr[s][1][vul][vr0]);
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
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

template <typename Args, typename Kern>
int launch_tc(Kern kern, const Args& a, int CL, int nclusters, size_t smem, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters * CL, 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, a, CL));
  return 0;
}

long long* g_dbg_clk_tc = nullptr;

template <int ACT, bool TAIL>
cudaError_t tc_attrs_one() {
  cudaError_t err = cudaFuncSetAttribute(ligru_fwd_tc_kernel<ACT, TAIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(FwdTc) + 1024));
  if (err == cudaSuccess) err = cudaFuncSetAttribute(ligru_fwd_tc_kernel<ACT, TAIL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  if (err == cudaSuccess)
    err = cudaFuncSetAttribute(ligru_bwd_tc_kernel<ACT, TAIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(BwdTc) + 1024));
  if (err == cudaSuccess) err = cudaFuncSetAttribute(ligru_bwd_tc_kernel<ACT, TAIL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  return err;
}
template <int ACT>
cudaError_t tc_attrs_act() {
  cudaError_t e = tc_attrs_one<ACT, false>();
  return e == cudaSuccess ? tc_attrs_one<ACT, true>() : e;
}

int tc_attrs() {
  static PerDeviceOnce once;
  const cudaError_t err = once.run([] {
    cudaError_t e = tc_attrs_act<ACT_RELU>();
    if (e == cudaSuccess) e = tc_attrs_act<ACT_TANH>();
    if (e == cudaSuccess) e = tc_attrs_act<ACT_SIGMOID>();
    if (e == cudaSuccess) e = tc_attrs_act<ACT_LEAKY_RELU>();
    if (e == cudaSuccess) e = tc_attrs_act<ACT_ELU>();
    if (e == cudaSuccess) e = tc_attrs_act<ACT_LINEAR>();
    return e;
  });
  PK_CHECK_CUDA(err);
  return 0;
}

#define PK_TC_LAUNCH(KERN, A, ARGS, SMEM) \
  return (CL > kTmemCL) ? launch_tc(KERN<A, true>, ARGS, CL, nclusters, SMEM, stream) : launch_tc(KERN<A, false>, ARGS, CL, nclusters, SMEM, stream)
#define PK_TC_DISPATCH(KERN, ARGS, SMEM)                                             \
  switch (ARGS.act) {                                                                \
    case ACT_RELU: PK_TC_LAUNCH(KERN, ACT_RELU, ARGS, SMEM);                         \
    case ACT_TANH: PK_TC_LAUNCH(KERN, ACT_TANH, ARGS, SMEM);                         \
    case ACT_SIGMOID: PK_TC_LAUNCH(KERN, ACT_SIGMOID, ARGS, SMEM);                   \
    case ACT_LEAKY_RELU: PK_TC_LAUNCH(KERN, ACT_LEAKY_RELU, ARGS, SMEM);             \
    case ACT_ELU: PK_TC_LAUNCH(KERN, ACT_ELU, ARGS, SMEM);                           \
    case ACT_LINEAR: PK_TC_LAUNCH(KERN, ACT_LINEAR, ARGS, SMEM);                     \
    default: PK_REQUIRE(false, "recurrent kernel: bad activation %d", ARGS.act);     \
  }

int pick_groups(int requested) {
  static const int env = [] { const char* e = getenv("PK_TC_GROUPS"); return e ? atoi(e) : 0; }();
  int g = requested > 0 ? requested : (env > 0 ? env : kDefaultGroups);
  return g < 1 ? 1 : (g > kGroups ? kGroups : g);
}

}  // namespace

void set_debug_clock_buffer_tc(long long* dev_ptr) { g_dbg_clk_tc = dev_ptr; }

int ligru_tc_max_hidden() { return kMaxCL * kUPC; }

int ligru_fwd_tc(const RecFwdArgs& a_in, cudaStream_t stream) {
  RecFwdArgs a = a_in;
  a.dbg_clk = g_dbg_clk_tc;
  a.groups = pick_groups(a.groups);
  const int CL = (a.H + kUPC - 1) / kUPC;
  PK_REQUIRE(CL <= kMaxCL, "ligru_fwd_tc: hidden size %d > %d", a.H, kMaxCL * kUPC);
  if (int rc = tc_attrs()) return rc;
  const int nclusters = (a.ndir * a.B + kRows - 1) / kRows;
  PK_TC_DISPATCH(ligru_fwd_tc_kernel, a, sizeof(FwdTc) + 1024)
  return 0;
}

int ligru_bwd_tc(const RecBwdArgs& a_in, cudaStream_t stream) {
  RecBwdArgs a = a_in;
  a.dbg_clk = g_dbg_clk_tc;
  a.groups = pick_groups(a.groups);
  const int CL = (a.H + kUPC - 1) / kUPC;
  PK_REQUIRE(CL <= kMaxCL, "ligru_bwd_tc: hidden size %d > %d", a.H, kMaxCL * kUPC);
  PK_REQUIRE(a.GT16 != nullptr, "ligru_bwd_tc: writes GT16 (required)");
  if (int rc = tc_attrs()) return rc;
  const int nclusters = (a.ndir * a.B + kRows - 1) / kRows;
  PK_TC_DISPATCH(ligru_bwd_tc_kernel, a, sizeof(BwdTc) + 1024)
  return 0;
}

// ---- dispatch of the persistent liGRU / RNN recurrence (C ABI: pk_rnn_layer_fwd / pk_rnn_layer_bwd)
// mode 0 = auto: the faster kernel for this hidden size as measured on B200 (profiles/r2_selftest_tc_vs_ws.log):
// register-stationary mma.sync (pk_rnn_ws.cu) up to H = 560, tcgen05 with TMEM-stationary weights beyond.
namespace {
int rec_mode(int requested, int H) {
  static const int env_mode = [] { const char* e = getenv("PK_REC_MODE"); return e ? atoi(e) : 0; }();
  int mode = requested ? requested : env_mode;
  if (mode == 0) mode = H <= 560 ? 2 : 3;
  return mode;
}
}  // namespace

int ligru_fwd(const RecFwdArgs& a, cudaStream_t stream) {
  PK_REQUIRE(a.T > 0 && a.B > 0 && a.H > 0, "ligru_fwd: empty problem");
  PK_REQUIRE(a.ndir == 1 || a.ndir == 2, "ligru_fwd: ndir must be 1 or 2");
  const int mode = rec_mode(a.legacy, a.H);
  PK_REQUIRE(mode == 2 || mode == 3, "ligru_fwd: kernel variant %d was removed (PK_REC_WS / PK_REC_TC select the two that exist)", mode);
  if (mode == 3) return ligru_fwd_tc(a, stream);
  PK_REQUIRE(a.H <= 560, "ligru_fwd: hidden size %d > 560 not supported by the register-stationary kernels", a.H);
  return ligru_fwd_ws(a, stream);
}

int ligru_bwd(const RecBwdArgs& a, cudaStream_t stream) {
  PK_REQUIRE(a.T > 0 && a.B > 0 && a.H > 0, "ligru_bwd: empty problem");
  PK_REQUIRE(a.ndir == 1 || a.ndir == 2, "ligru_bwd: ndir must be 1 or 2");
  PK_REQUIRE(a.GT16 != nullptr, "ligru_bwd: the kernels write GT16 (required)");
  const int mode = rec_mode(a.legacy, a.H);
  PK_REQUIRE(mode == 2 || mode == 3, "ligru_bwd: kernel variant %d was removed (PK_REC_WS / PK_REC_TC select the two that exist)", mode);
  if (mode == 3) return ligru_bwd_tc(a, stream);
  PK_REQUIRE(a.H <= 560, "ligru_bwd: hidden size %d > 560 not supported by the register-stationary kernels", a.H);
  return ligru_bwd_ws(a, stream);
}

}  // namespace pk
