// lstm_persist.cu -- persistent-RNN kernels for the LSTM recurrences (bf16 tensor-core precision).
//
// One launch runs ALL timesteps of one recurrence (an encoder layer or a decoder cell of
// vae_definition.py:455-474,533-632).  The reference runs this loop as a Theano scan / K.rnn with one
// small gemm + elementwise ops per step; the step-streamed form in model.cu does the same with one
// tcgen05 GEMM + one pointwise launch per step.  Here the loop lives inside the kernel:
//
//   * batch rows are independent, so the batch is cut into GROUPS of 128 rows (one UMMA M tile);
//   * within a group, CTA j owns HS hidden units (all four gates of each): its slice of the recurrent
//     weights stays RESIDENT in shared memory for the whole sequence (<= 128 KB), as the B operand;
//   * per step each CTA TMA-loads the group's h_{t-1} (forward) or dG_{t+1} (backward) rows from L2 as the
//     A operand, runs the K loop with tcgen05.mma into TMEM, and its four epilogue warps (one batch row per
//     thread) apply the gate math with the cell state c (forward) / dc (backward) living in REGISTERS for
//     the whole sequence, write the stash (gates, c, h / dG) with 16-byte stores, and publish the step with
//     a release-increment of a per-(group, step) counter that the TMA-producer warps of the group's CTAs
//     acquire before loading the next step's operand.  No grid-wide barrier, no host round trip.
//
// Forward:  pre = xw_t + h_{t-1} U ;  i,f,o = gate(pre) ; g = tanh(pre) ; c' = f c + i g ; h' = o tanh(c')
// Backward: dh_t = dh_ext_t + dG_{t+1} U^T ;  dG_t = pointwise(dh_t, dc, stash_t) ; dc <- ds f      (oracle/manual_bptt.py)
#include <cuda.h>

#include <stdlib.h>

#include <algorithm>

#include "../../include/midivae.h"
#include "common.cuh"
#include "ptx.cuh"
#include "rec_persist.cuh"

namespace mvae {
namespace {

using bf16 = __nv_bfloat16;
constexpr int BM = 128, BK = 64, UMMA_K = 16, MAX_STAGES = 12, A_STAGE_BYTES = BM * BK * 2;
constexpr int EPI_WARPS = 8, PROD_LANES = 8, ROW_THREADS = 32 * EPI_WARPS;
constexpr int kThreads = 32 * (2 + EPI_WARPS);

struct RecKP {
  int n, H, G, steps, cpg, HS, group0, stages, kb_rot, prod_lanes;
  int gate_act, variant;
  unsigned* flags;   // [groups][steps + 2]
  int flag_stride;
  // forward
  const bf16* xw; bf16* hseq; bf16* cseq; bf16* gates; const bf16* c0; int ldc0;
  // backward
  const bf16* dhext; const bf16* dh_last; int ld_last; bf16* dG; bf16* dS_h; bf16* dS_c; int ldS;
  long long* trace;   // optional per-phase clock64 stamps of CTA 0 (debug / profiling)
  void* partial;      // K-split backward: bf16 exchange buffer [2][groups_total][cpg][128][H]
  bf16* hx;           // forward: h exchange buffer [2][groups_total][H/8 granules][128 rows][8]
  int groups_total;
};

#define REC_TRACE(step, point)                                                                  \
  do {                                                                                          \
    if (p.trace && blockIdx.x == 0 && (step) >= 16 && (step) < 24) p.trace[((step) - 16) * 16 + (point)] = clock64(); \
  } while (0)

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// stash layout of the persistent kernels (private to them): [slab][column/8 granules][n rows][8] -- the 32 lanes of a warp
// are 32 consecutive rows, so one 16-byte access per lane is 512 contiguous bytes per warp instruction
__device__ __forceinline__ size_t gran_off(int slab, int ngran, int gran, int n, int m) {
  return (((size_t)slab * ngran + gran) * n + m) * 8;
}
__device__ __forceinline__ void st_shared16(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void red_relaxed_add(unsigned* p, unsigned v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void wait_flag(const unsigned* p, unsigned target) {
  long long t0 = clock64();
  while (ld_acquire(p) < target) {
    if (clock64() - t0 > 4000000000LL) {   // ~2 s: a protocol bug must trap, not hang the GPU
      printf("rec_persist: flag wait timeout block %d (have %u want %u)\n", (int)blockIdx.x, ld_acquire(p), target);
      __trap();
    }
  }
}

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gate_fwd(int gate_act, float x) {
  return gate_act == MVAE_GATE_HARD_SIGMOID ? fminf(fmaxf(0.2f * x + 0.5f, 0.f), 1.f) : 0.5f * tanh_fast(0.5f * x) + 0.5f;
}
__device__ __forceinline__ float gate_bwd(int gate_act, float s) {
  return gate_act == MVAE_GATE_HARD_SIGMOID ? ((s > 0.f && s < 1.f) ? 0.2f : 0.f) : s * (1.f - s);
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

// one or two independent recurrences per launch: CTAs [0, ctas0) run recurrence 0, the rest recurrence 1
struct KsplitPair {
  RecKP p[2];
  int ctas0;
};

// FWD: A = h_{t-1} rows of the group (K = H), B = packed U slice [4*HS gate columns][H] resident, D = [128 rows][4*HS]
//      packed column n = ublock*32 + gate*8 + u8  <->  unit j*HS + ublock*8 + u8, semantic gate (i,f,g,o)
// BWD: A = dG_{t+1} rows of the group (K = 4H), B = U rows of the CTA's HS units [HS][4H] resident, D = [128 rows][HS]
//
// Warp roles (10 warps): 0 = TMA producer (PROD_LANES lanes issue K-blocks concurrently: one lane's issue costs ~350
// cycles, so a single issuer would pace the whole gather), 1 = MMA issuer + TMEM owner, 2..9 = epilogue: warp w works on
// TMEM lane quadrant w%4 (one batch row per lane) and on the 8-unit chunks ub with ub % 2 == (w-2)/4, so that every
// SM sub-partition has two epilogue warps to hide each other's latencies.
template <bool FWD, int HS>
__global__ void __launch_bounds__(kThreads, 1)
rec_persist_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b0, const __grid_constant__ CUtensorMap tma_b1,
                   const __grid_constant__ KsplitPair pp) {
  const int which = (int)blockIdx.x >= pp.ctas0 ? 1 : 0;
  const RecKP& p = pp.p[which];
  const CUtensorMap* tmb = which ? &tma_b1 : &tma_b0;
  const int bid = (int)blockIdx.x - (which ? pp.ctas0 : 0);
  constexpr int BN = FWD ? 4 * HS : HS;              // UMMA N
  constexpr int ACC_STRIDE = BN < 32 ? 32 : BN;      // the epilogue reads 32-column chunks
  constexpr int TM_COLS = 2 * ACC_STRIDE;
  constexpr int B_KB_BYTES = BN * BK * 2;            // bytes of one resident K-block of B
  constexpr int NCH = HS / 8;                        // 8-unit chunks per CTA
  constexpr int MYCH = NCH / 2;                      // chunks per epilogue thread
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tmem_full_bar[2], tmem_empty_bar[2], b_full_bar;
  __shared__ uint32_t tmem_base_slot;
  const int A_STAGES = p.stages;

  const int K = FWD ? p.H : p.G;
  const int kblocks = K / BK;
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_b = smem_base;                                   // resident weights: kblocks * B_KB_BYTES
  const uint32_t smem_a = smem_base + (uint32_t)kblocks * B_KB_BYTES;   // A ring
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = p.group0 + bid / p.cpg, j = bid % p.cpg;
  const int row0 = g * BM;
  unsigned* flags = p.flags + (size_t)g * p.flag_stride;
  const int T = p.steps;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tma_a);
    ptx::prefetch_tmap(tmb);
    for (int s = 0; s < A_STAGES; ++s) { ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1); ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(ptx::smem_u32(&tmem_full_bar[a]), 1); ptx::mbar_init(ptx::smem_u32(&tmem_empty_bar[a]), EPI_WARPS); }
    ptx::mbar_init(ptx::smem_u32(&b_full_bar), 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), TM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  // number of MMA rounds: forward T (one per step); backward T (dG_{t+1} U^T for t = T-2..0 and the final dh0)
  const int rounds = T;

  if (warp == 0) {
    // ===================== TMA producer: PROD_LANES lanes, K-block sequence number q -> lane q % PROD_LANES =====================
    if (lane == 0) {   // resident weights, once
      const uint32_t bb = ptx::smem_u32(&b_full_bar);
      ptx::mbar_arrive_expect_tx(bb, (uint32_t)kblocks * B_KB_BYTES);
      for (int kb = 0; kb < kblocks; ++kb) ptx::tma_load_2d(smem_b + kb * B_KB_BYTES, tmb, bb, kb * BK, j * BN);
    }
    if (lane < p.prod_lanes) {
      const int NL = p.prod_lanes;
      const long total = (long)rounds * kblocks;
      int waited_round = -1;
      for (long q = lane; q < total; q += NL) {
        const int r = (int)(q / kblocks), kb = (int)(q % kblocks);
        // forward round r consumes h_{r-1} = hseq slab r (slab 0 = initial state, written before the launch);
        // backward round r consumes dG slab (T-1-r), published by the epilogue of iteration r
        const int slab = FWD ? r : (T - 1 - r);
        if (waited_round != r) {
          if (lane == 0) REC_TRACE(r, 0);
          wait_flag(flags + slab, (unsigned)p.cpg);
          fence_proxy_async_global();
          waited_round = r;
          if (lane == 0) REC_TRACE(r, 1);
        }
        const int stage = (int)(q % A_STAGES);
        const uint32_t phase = (uint32_t)((q / A_STAGES) & 1);
        ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1);
        const uint32_t fb = ptx::smem_u32(&full_bar[stage]);
        ptx::mbar_arrive_expect_tx(fb, A_STAGE_BYTES);
        // CTAs of a group read the SAME rows: start each at a different K-block so they do not sweep the same L2 lines in lockstep
        const int kbr = (kb + j * p.kb_rot) % kblocks;
        if (FWD) {
          // h_{r-1} of the whole group is one contiguous 128*H*2-byte block [H/8 granules][128 rows][8] of the exchange buffer;
          // K-block kbr = granules 8*kbr .. 8*kbr+7 = 16 KB contiguous -> one bulk copy, landing as the no-swizzle K-major image
          const bf16* src = p.hx + ((size_t)(r & 1) * p.groups_total + g) * ((size_t)BM * p.H) + (size_t)kbr * 8 * BM * 8;
          ptx::bulk_load(smem_a + stage * A_STAGE_BYTES, src, A_STAGE_BYTES, fb);
        } else {
          ptx::tma_load_2d(smem_a + stage * A_STAGE_BYTES, &tma_a, fb, kbr * BK, slab * p.n + row0);
        }
        if (kb == kblocks - 1) REC_TRACE(r, 2);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN, false, false);
      ptx::mbar_wait(ptx::smem_u32(&b_full_bar), 0);
      int stage = 0; uint32_t phase = 0;
      for (int r = 0; r < rounds; ++r) {
        const int acc = r & 1; const uint32_t acc_phase = (r >> 1) & 1;
        ptx::mbar_wait(ptx::smem_u32(&tmem_empty_bar[acc]), acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
          ptx::tc_fence_after();
          if (kb == 0) REC_TRACE(r, 3);
          if (kb == kblocks - 1) REC_TRACE(r, 4);
          const uint32_t sa = smem_a + stage * A_STAGE_BYTES, sb = smem_b + ((kb + j * p.kb_rot) % kblocks) * B_KB_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // forward: A tile is [8 granules][128 rows][16 B] (no swizzle): core matrices 128 B apart along M (SBO), 2 KB apart along K (LBO)
            const uint64_t da = FWD ? ptx::umma_desc_noswz(sa + k * 4096, 2048, 128) : ptx::umma_desc_sw128(sa + k * 32, 16, 1024);
            ptx::umma_bf16(d_tmem, da, ptx::umma_desc_sw128(sb + k * 32, 16, 1024), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(ptx::smem_u32(&empty_bar[stage]));
          if (++stage == A_STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(ptx::smem_u32(&tmem_full_bar[acc]));
      }
    }
  } else {
    // ===================== epilogue warps 2..9 =====================
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int m = row0 + quad * 32 + lane;
    const bool row_ok = m < p.n;
    const bool tracer = (threadIdx.x == 64);
    const int H = p.H, G = p.G;
    const int bi = p.variant == MVAE_CELL_STANDARD ? 0 : 1, bfk = 1 - bi;   // column block of the i and f gates
    const int u0 = j * HS;
    float cst[MYCH * 8];   // forward: cell state c ; backward: dc   (this thread's units, for the whole sequence)
#pragma unroll
    for (int u = 0; u < MYCH * 8; ++u) cst[u] = 0.f;

    if (FWD) {
      if (row_ok) {
#pragma unroll
        for (int c = 0; c < MYCH; ++c) {
          const int ub = 2 * c + half;
          uint4 cv = make_uint4(0u, 0u, 0u, 0u);
          if (p.c0) cv = *reinterpret_cast<const uint4*>(p.c0 + (size_t)m * p.ldc0 + u0 + ub * 8);
          unpack8(cv, &cst[c * 8]);
          *reinterpret_cast<uint4*>(p.cseq + gran_off(0, H / 8, (u0 >> 3) + ub, p.n, m)) = cv;   // stash slab 0 = c0
        }
      }
      {   // h0 (hseq slab 0, row-major) -> exchange buffer 0, published like any other step
        bf16* hx0 = p.hx + ((size_t)0 * p.groups_total + g) * ((size_t)BM * H);
#pragma unroll
        for (int c = 0; c < MYCH; ++c) {
          const int ub = 2 * c + half;
          uint4 hv = make_uint4(0u, 0u, 0u, 0u);
          if (row_ok) hv = *reinterpret_cast<const uint4*>(p.hseq + (size_t)m * H + u0 + ub * 8);
          *reinterpret_cast<uint4*>(hx0 + ((size_t)((u0 >> 3) + ub) * BM + (quad * 32 + lane)) * 8) = hv;
        }
        epi_barrier();
        if (warp == 2 && lane == 0) { __threadfence(); fence_proxy_async_global(); red_relaxed_add(flags + 0, 1u); }
      }
      for (int t = 0; t < T; ++t) {
        const int acc = t & 1; const uint32_t acc_phase = (t >> 1) & 1;
        const size_t rowG = ((size_t)t * p.n + m) * G, rowH1 = ((size_t)(t + 1) * p.n + m) * H;
        // the input projection of this step does not depend on the recurrence: fetch it while the MMAs run
        uint4 xq[MYCH][4];
        if (row_ok) {
#pragma unroll
          for (int c = 0; c < MYCH; ++c) {
            const int ub = 2 * c + half;
            xq[c][0] = __ldg(reinterpret_cast<const uint4*>(p.xw + rowG + bi * H + u0 + ub * 8));
            xq[c][1] = __ldg(reinterpret_cast<const uint4*>(p.xw + rowG + bfk * H + u0 + ub * 8));
            xq[c][2] = __ldg(reinterpret_cast<const uint4*>(p.xw + rowG + 2 * H + u0 + ub * 8));
            xq[c][3] = __ldg(reinterpret_cast<const uint4*>(p.xw + rowG + 3 * H + u0 + ub * 8));
          }
        }
        if (tracer) REC_TRACE(t, 5);
        ptx::mbar_wait(ptx::smem_u32(&tmem_full_bar[acc]), acc_phase);
        ptx::tc_fence_after();
        if (tracer) REC_TRACE(t, 6);
        uint4 st_g[MYCH][4], st_c[MYCH], st_h[MYCH];
        bf16* hx1 = p.hx + ((size_t)((t + 1) & 1) * p.groups_total + g) * ((size_t)BM * H);
#pragma unroll
        for (int c = 0; c < MYCH; ++c) {
          const int ub = 2 * c + half;
          float v[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * ACC_STRIDE + ub * 32), v);
          float xi[8], xf[8], xg[8], xo[8], gi[8], gf[8], gg[8], go[8], cn[8], hn[8];
          unpack8(xq[c][0], xi); unpack8(xq[c][1], xf); unpack8(xq[c][2], xg); unpack8(xq[c][3], xo);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            gi[u] = gate_fwd(p.gate_act, v[u] + xi[u]);
            gf[u] = gate_fwd(p.gate_act, v[8 + u] + xf[u]);
            gg[u] = tanh_fast(v[16 + u] + xg[u]);
            go[u] = gate_fwd(p.gate_act, v[24 + u] + xo[u]);
            const float s = gf[u] * cst[c * 8 + u] + gi[u] * gg[u];
            if (p.variant == MVAE_CELL_STANDARD) { cn[u] = s; hn[u] = go[u] * tanh_fast(s); }
            else { cn[u] = tanh_fast(s); hn[u] = go[u] * cn[u]; }
            cst[c * 8 + u] = cn[u];
          }
          st_h[c] = pack8(hn);
          if (!row_ok) st_h[c] = make_uint4(0u, 0u, 0u, 0u);
          // what the other CTAs wait for: coalesced 16-byte granules of the exchange buffer
          *reinterpret_cast<uint4*>(hx1 + ((size_t)((u0 >> 3) + ub) * BM + (quad * 32 + lane)) * 8) = st_h[c];
          st_g[c][0] = pack8(gi); st_g[c][1] = pack8(gf); st_g[c][2] = pack8(gg); st_g[c][3] = pack8(go); st_c[c] = pack8(cn);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&tmem_empty_bar[acc]));
        // publish h_t: the CTA barrier orders every thread's h stores before thread 64, whose gpu-scope fence is cumulative
        if (tracer) REC_TRACE(t, 7);
        fence_proxy_async_global();
        if (tracer) REC_TRACE(t, 8);
        epi_barrier();
        if (tracer) {
          REC_TRACE(t, 9);
          __threadfence();
          REC_TRACE(t, 10);
          red_relaxed_add(flags + (t + 1), 1u);
          REC_TRACE(t, 11);
        }
        // the rest of the stash (read only by the backward pass) is written after the release: off the critical path
        if (row_ok) {
#pragma unroll
          for (int c = 0; c < MYCH; ++c) {
            const int ub = 2 * c + half;
            const int gu = (u0 >> 3) + ub;
            *reinterpret_cast<uint4*>(p.hseq + rowH1 + u0 + ub * 8) = st_h[c];      // row-major copy for the batched GEMMs
            *reinterpret_cast<uint4*>(p.cseq + gran_off(t + 1, H / 8, gu, p.n, m)) = st_c[c];
            *reinterpret_cast<uint4*>(p.gates + gran_off(t, G / 8, bi * (H / 8) + gu, p.n, m)) = st_g[c][0];
            *reinterpret_cast<uint4*>(p.gates + gran_off(t, G / 8, bfk * (H / 8) + gu, p.n, m)) = st_g[c][1];
            *reinterpret_cast<uint4*>(p.gates + gran_off(t, G / 8, 2 * (H / 8) + gu, p.n, m)) = st_g[c][2];
            *reinterpret_cast<uint4*>(p.gates + gran_off(t, G / 8, 3 * (H / 8) + gu, p.n, m)) = st_g[c][3];
          }
        }
      }
    } else {
      // backward: iteration it = 0..T: t = T-1-it is the step whose dG is produced; it == T produces dh0 / dc0 only
      for (int it = 0; it <= T; ++it) {
        const int t = T - 1 - it;
        // stash loads for step t do not depend on the recurrence
        uint4 sg[MYCH][4], sc0[MYCH], sc1[MYCH], se[MYCH];
        if (t >= 0 && row_ok) {
          const size_t rowG = ((size_t)t * p.n + m) * G, rowH0 = ((size_t)t * p.n + m) * H, rowH1 = ((size_t)(t + 1) * p.n + m) * H;
#pragma unroll
          for (int c = 0; c < MYCH; ++c) {
            const int ub = 2 * c + half;
            const int gu = (u0 >> 3) + ub;
            sg[c][0] = __ldg(reinterpret_cast<const uint4*>(p.gates + gran_off(t, G / 8, bi * (H / 8) + gu, p.n, m)));
            sg[c][1] = __ldg(reinterpret_cast<const uint4*>(p.gates + gran_off(t, G / 8, bfk * (H / 8) + gu, p.n, m)));
            sg[c][2] = __ldg(reinterpret_cast<const uint4*>(p.gates + gran_off(t, G / 8, 2 * (H / 8) + gu, p.n, m)));
            sg[c][3] = __ldg(reinterpret_cast<const uint4*>(p.gates + gran_off(t, G / 8, 3 * (H / 8) + gu, p.n, m)));
            sc0[c] = __ldg(reinterpret_cast<const uint4*>(p.cseq + gran_off(t, H / 8, gu, p.n, m)));
            sc1[c] = __ldg(reinterpret_cast<const uint4*>(p.cseq + gran_off(t + 1, H / 8, gu, p.n, m)));
            if (p.dhext) se[c] = __ldg(reinterpret_cast<const uint4*>(p.dhext + rowH0 + u0 + ub * 8));
          }
        }
        float v[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) v[u] = 0.f;
        if (it > 0) {
          const int r = it - 1;
          const int acc = r & 1; const uint32_t acc_phase = (r >> 1) & 1;
          if (tracer) REC_TRACE(r, 5);
          ptx::mbar_wait(ptx::smem_u32(&tmem_full_bar[acc]), acc_phase);
          ptx::tc_fence_after();
          if (tracer) REC_TRACE(r, 6);
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * ACC_STRIDE), v);   // columns >= HS are never used
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&tmem_empty_bar[acc]));
        }
        if (t >= 0) {
          if (row_ok) {
            const size_t rowG = ((size_t)t * p.n + m) * G;
#pragma unroll
            for (int c = 0; c < MYCH; ++c) {
              const int ub = 2 * c + half;
              float gi[8], gf[8], gg[8], go[8], c0[8], c1[8], ex[8], di[8], df[8], dg[8], dob[8];
              unpack8(sg[c][0], gi); unpack8(sg[c][1], gf); unpack8(sg[c][2], gg); unpack8(sg[c][3], go);
              unpack8(sc0[c], c0); unpack8(sc1[c], c1);
              if (p.dhext) unpack8(se[c], ex);
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                float d = half ? v[((2 * c + 1) * 8 + u) & 31] : v[(2 * c * 8 + u) & 31];
                if (p.dhext) d += ex[u];
                if (it == 0 && p.dh_last) d += __bfloat162float(p.dh_last[(size_t)m * p.ld_last + u0 + ub * 8 + u]);
                float d_o, ds;
                if (p.variant == MVAE_CELL_STANDARD) {
                  const float tc = tanh_fast(c1[u]);
                  d_o = d * tc;
                  ds = cst[c * 8 + u] + d * go[u] * (1.f - tc * tc);
                } else {
                  d_o = d * c1[u];
                  ds = (cst[c * 8 + u] + d * go[u]) * (1.f - c1[u] * c1[u]);
                }
                di[u] = ds * gg[u] * gate_bwd(p.gate_act, gi[u]);
                df[u] = ds * c0[u] * gate_bwd(p.gate_act, gf[u]);
                dg[u] = ds * gi[u] * (1.f - gg[u] * gg[u]);
                dob[u] = d_o * gate_bwd(p.gate_act, go[u]);
                cst[c * 8 + u] = ds * gf[u];
              }
              bf16* dgp = p.dG + rowG + u0 + ub * 8;
              *reinterpret_cast<uint4*>(dgp + bi * H) = pack8(di);
              *reinterpret_cast<uint4*>(dgp + bfk * H) = pack8(df);
              *reinterpret_cast<uint4*>(dgp + 2 * H) = pack8(dg);
              *reinterpret_cast<uint4*>(dgp + 3 * H) = pack8(dob);
            }
          }
          if (tracer) REC_TRACE(it, 7);
          fence_proxy_async_global();
          if (tracer) REC_TRACE(it, 8);
          epi_barrier();
          if (tracer) {
            REC_TRACE(it, 9);
            __threadfence();
            REC_TRACE(it, 10);
            red_relaxed_add(flags + t, 1u);
            REC_TRACE(it, 11);
          }
        } else if (row_ok && p.dS_h) {
          // gradients wrt the initial states (h0, c0) of a decoder cell
#pragma unroll
          for (int c = 0; c < MYCH; ++c) {
            const int ub = 2 * c + half;
            float d8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) d8[u] = half ? v[((2 * c + 1) * 8 + u) & 31] : v[(2 * c * 8 + u) & 31];
            *reinterpret_cast<uint4*>(p.dS_h + (size_t)m * p.ldS + u0 + ub * 8) = pack8(d8);
            *reinterpret_cast<uint4*>(p.dS_c + (size_t)m * p.ldS + u0 + ub * 8) = pack8(&cst[c * 8]);
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, TM_COLS);
}

// ------------------------------------------------------------------------------------------------------------------
// Backward recurrence, K-split form (H <= 512).
//
// dh_{t-1} = dG_t U^T needs, per output unit, ALL 4H columns of dG_t.  The N-split kernel above therefore gathers the
// group's whole dG_t (128 rows x 4H = 512 KB at H=512) into every CTA each step -- 4x the forward exchange, and the
// gather is what its step time is made of.  Here each CTA instead multiplies only the 64 gate columns it PRODUCES itself
// (operand A written straight from registers into the swizzled smem tile, no gather at all) with the matching 64 rows
// of U^T (resident, [H units][64] K-major), giving a partial dh for ALL H units in TMEM (H fp32 columns); the partials
// are exchanged as bf16 through L2 (128 KB written + 128 KB read per CTA per step, the forward's volume) and summed in
// fp32 by the CTA that owns the units.
//   iteration it (t = T-1-it):  [reduce partials of t+1 -> dh_t] -> gate-gradient math -> dG_t (smem A tile + global)
//                               -> 8 UMMAs (M128 x N256 x K16) -> TMEM -> bf16 partial -> global -> publish flags[t]
template <int HS>
__global__ void __launch_bounds__(kThreads, 1)
rec_bwd_ksplit_kernel(const __grid_constant__ CUtensorMap tma_b, const RecKP p) {
  static_assert(HS == 16, "one 64-column K block per CTA");
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t b_full_bar, a_full_bar, tmem_full_bar, tmem_empty_bar;
  __shared__ uint32_t tmem_base_slot;

  const int H = p.H, G = p.G, T = p.steps;
  const int NH = H <= 256 ? 1 : H / 256, NB = H / NH;              // MMA N chunks
  const uint32_t tm_cols = H <= 32 ? 32 : (H <= 64 ? 64 : (H <= 128 ? 128 : (H <= 256 ? 256 : 512)));
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_b = smem_base;                                // [H][64] bf16 K-major, H*128 bytes
  const uint32_t smem_a = smem_base + (uint32_t)H * 128;            // [128][64] bf16 K-major, 16 KB
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = p.group0 + blockIdx.x / p.cpg, j = blockIdx.x % p.cpg;
  const int row0 = g * BM;
  unsigned* flags = p.flags + (size_t)g * p.flag_stride;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tma_b);
    ptx::mbar_init(ptx::smem_u32(&b_full_bar), 1);
    ptx::mbar_init(ptx::smem_u32(&a_full_bar), ROW_THREADS);
    ptx::mbar_init(ptx::smem_u32(&tmem_full_bar), 1);
    ptx::mbar_init(ptx::smem_u32(&tmem_empty_bar), EPI_WARPS);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), tm_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t bb = ptx::smem_u32(&b_full_bar);
      ptx::mbar_arrive_expect_tx(bb, (uint32_t)H * 128);
      for (int nh = 0; nh < NH; ++nh) ptx::tma_load_2d(smem_b + nh * NB * 128, &tma_b, bb, 0, j * H + nh * NB);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(BM, NB, false, false);
      ptx::mbar_wait(ptx::smem_u32(&b_full_bar), 0);
      for (int it = 0; it < T; ++it) {
        ptx::mbar_wait(ptx::smem_u32(&a_full_bar), it & 1);                 // dG_t tile written by the row warps
        ptx::mbar_wait(ptx::smem_u32(&tmem_empty_bar), (it & 1) ^ 1);       // previous partial drained
        ptx::tc_fence_after();
        for (int nh = 0; nh < NH; ++nh)
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            ptx::umma_bf16(tmem_base + nh * NB, ptx::umma_desc_sw128(smem_a + k * 32, 16, 1024),
                           ptx::umma_desc_sw128(smem_b + nh * NB * 128 + k * 32, 16, 1024), idesc, k > 0 ? 1u : 0u);
        ptx::umma_commit(ptx::smem_u32(&tmem_full_bar));
      }
    }
  } else {
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int etid = threadIdx.x - 64;
    const int r = quad * 32 + lane, m = row0 + r;
    const bool row_ok = m < p.n;
    const bool tracer = (etid == 0);
    const int bi = p.variant == MVAE_CELL_STANDARD ? 0 : 1, bfk = 1 - bi;
    const int u0 = j * HS + half * 8;                                   // this thread's 8 units
    bf16* part = (bf16*)p.partial;                                      // [2][groups_total][cpg][128][H]
    const size_t part_cta = (size_t)BM * H, part_grp = part_cta * p.cpg, part_buf = part_grp * p.groups_total;
    float dc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) dc[u] = 0.f;

    for (int it = 0; it <= T; ++it) {
      const int t = T - 1 - it;
      uint4 sg[4], sc0, sc1, se;
      if (t >= 0 && row_ok) {
        const size_t rowG = ((size_t)t * p.n + m) * G, rowH0 = ((size_t)t * p.n + m) * H, rowH1 = ((size_t)(t + 1) * p.n + m) * H;
        const int gu = u0 >> 3;
        sg[0] = __ldg(reinterpret_cast<const uint4*>(p.gates + gran_off(t, G / 8, bi * (H / 8) + gu, p.n, m)));
        sg[1] = __ldg(reinterpret_cast<const uint4*>(p.gates + gran_off(t, G / 8, bfk * (H / 8) + gu, p.n, m)));
        sg[2] = __ldg(reinterpret_cast<const uint4*>(p.gates + gran_off(t, G / 8, 2 * (H / 8) + gu, p.n, m)));
        sg[3] = __ldg(reinterpret_cast<const uint4*>(p.gates + gran_off(t, G / 8, 3 * (H / 8) + gu, p.n, m)));
        sc0 = __ldg(reinterpret_cast<const uint4*>(p.cseq + gran_off(t, H / 8, gu, p.n, m)));
        sc1 = __ldg(reinterpret_cast<const uint4*>(p.cseq + gran_off(t + 1, H / 8, gu, p.n, m)));
        if (p.dhext) se = __ldg(reinterpret_cast<const uint4*>(p.dhext + rowH0 + u0));
      }
      // ---- dh_t = sum over the group's CTAs of their partial dG_{t+1} U^T, for this thread's 8 units
      float dh[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) dh[u] = 0.f;
      if (it > 0) {
        if (tracer) { REC_TRACE(it, 0); wait_flag(flags + (t + 1), (unsigned)p.cpg); REC_TRACE(it, 1); }
        epi_barrier();
        // exchange layout [cta][H/8 granules][128 rows][8]: a warp's 32 rows read 512 contiguous bytes per instruction
        const bf16* src = part + (size_t)((it - 1) & 1) * part_buf + (size_t)g * part_grp + ((size_t)(u0 >> 3) * BM + r) * 8;
        // software-pipelined: batch b+1 is in flight while batch b is summed (the loop is one L2 latency long, not cpg/8)
        uint4 qa[8], qb[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) if (jj < p.cpg) qa[jj] = __ldcg(reinterpret_cast<const uint4*>(src + (size_t)jj * part_cta));
        for (int j0 = 0; j0 < p.cpg; j0 += 16) {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) if (j0 + 8 + jj < p.cpg) qb[jj] = __ldcg(reinterpret_cast<const uint4*>(src + (size_t)(j0 + 8 + jj) * part_cta));
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            if (j0 + jj < p.cpg) {
              float f[8];
              unpack8(qa[jj], f);
#pragma unroll
              for (int u = 0; u < 8; ++u) dh[u] += f[u];
            }
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) if (j0 + 16 + jj < p.cpg) qa[jj] = __ldcg(reinterpret_cast<const uint4*>(src + (size_t)(j0 + 16 + jj) * part_cta));
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            if (j0 + 8 + jj < p.cpg) {
              float f[8];
              unpack8(qb[jj], f);
#pragma unroll
              for (int u = 0; u < 8; ++u) dh[u] += f[u];
            }
        }
        if (tracer) REC_TRACE(it, 2);
      }
      if (t < 0) {
        if (row_ok && p.dS_h) {
          *reinterpret_cast<uint4*>(p.dS_h + (size_t)m * p.ldS + u0) = pack8(dh);
          *reinterpret_cast<uint4*>(p.dS_c + (size_t)m * p.ldS + u0) = pack8(dc);
        }
        break;
      }
      // ---- gate-gradient math for step t
      float di[8], df[8], dg[8], dob[8];
      {
        float gi[8], gf[8], gg[8], go[8], c0[8], c1[8], ex[8];
        unpack8(sg[0], gi); unpack8(sg[1], gf); unpack8(sg[2], gg); unpack8(sg[3], go); unpack8(sc0, c0); unpack8(sc1, c1);
        if (p.dhext) unpack8(se, ex);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float d = dh[u];
          if (p.dhext) d += ex[u];
          if (it == 0 && p.dh_last) d += row_ok ? __bfloat162float(p.dh_last[(size_t)m * p.ld_last + u0 + u]) : 0.f;
          float d_o, ds;
          if (p.variant == MVAE_CELL_STANDARD) {
            const float tc = tanh_fast(c1[u]);
            d_o = d * tc;
            ds = dc[u] + d * go[u] * (1.f - tc * tc);
          } else {
            d_o = d * c1[u];
            ds = (dc[u] + d * go[u]) * (1.f - c1[u] * c1[u]);
          }
          di[u] = ds * gg[u] * gate_bwd(p.gate_act, gi[u]);
          df[u] = ds * c0[u] * gate_bwd(p.gate_act, gf[u]);
          dg[u] = ds * gi[u] * (1.f - gg[u] * gg[u]);
          dob[u] = d_o * gate_bwd(p.gate_act, go[u]);
          dc[u] = ds * gf[u];
          if (!row_ok) { di[u] = 0.f; df[u] = 0.f; dg[u] = 0.f; dob[u] = 0.f; dc[u] = 0.f; }
        }
      }
      const uint4 pi_ = pack8(di), pf_ = pack8(df), pg_ = pack8(dg), po_ = pack8(dob);
      // operand A: row r = [di(16) | df(16) | dg(16) | do(16)], 16-byte chunk (gate*2 + half), SWIZZLE_128B image
      {
        const uint32_t rowa = smem_a + r * 128;
        const int sw = r & 7;
        st_shared16(rowa + (((0 * 2 + half) ^ sw) << 4), pi_);
        st_shared16(rowa + (((1 * 2 + half) ^ sw) << 4), pf_);
        st_shared16(rowa + (((2 * 2 + half) ^ sw) << 4), pg_);
        st_shared16(rowa + (((3 * 2 + half) ^ sw) << 4), po_);
      }
      ptx::fence_proxy_async();                                     // generic-proxy smem writes -> tensor core (async proxy)
      ptx::mbar_arrive(ptx::smem_u32(&a_full_bar));
      // ---- partial dh_{t-1} for all H units: TMEM -> bf16 -> this CTA's slot of the exchange buffer
      if (tracer) REC_TRACE(it, 5);
      ptx::mbar_wait(ptx::smem_u32(&tmem_full_bar), it & 1);
      ptx::tc_fence_after();
      if (tracer) REC_TRACE(it, 6);
      {
        bf16* dst = part + (size_t)(it & 1) * part_buf + (size_t)g * part_grp + (size_t)j * part_cta;
        const int nch = H / 64;                                      // 32-column chunks in this thread's half
        for (int c = 0; c < nch; ++c) {
          float v[32];
          const int col0 = half * (H / 2) + c * 32;
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)col0, v);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) *reinterpret_cast<uint4*>(dst + ((size_t)((col0 >> 3) + q4) * BM + r) * 8) = pack8(&v[q4 * 8]);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&tmem_empty_bar));
      if (tracer) REC_TRACE(it, 7);
      epi_barrier();
      if (tracer) {
        REC_TRACE(it, 9);
        __threadfence();
        REC_TRACE(it, 10);
        red_relaxed_add(flags + t, 1u);
        REC_TRACE(it, 11);
      }
      if (row_ok) {   // dG_t for the batched weight-gradient GEMMs: row-major (16-byte pieces), off the critical path
        bf16* dgp = p.dG + ((size_t)t * p.n + m) * G + u0;
        *reinterpret_cast<uint4*>(dgp + bi * H) = pi_;
        *reinterpret_cast<uint4*>(dgp + bfk * H) = pf_;
        *reinterpret_cast<uint4*>(dgp + 2 * H) = pg_;
        *reinterpret_cast<uint4*>(dgp + 3 * H) = po_;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, tm_cols);
}

// ------------------------------------------------------------------------------------------------------------------
// K-split backward, general form: HS = 16 or 32 hidden units per CTA (K = 4*HS = 64 or 128 gate columns) and ONE or TWO
// independent recurrences per launch.  Every phase of a step is an L2 round trip, so a recurrence on its own leaves the
// SMs idle ~90 % of the time; two recurrences that do not depend on each other (e.g. the top notes cell and the velocity
// cell of the decoder) therefore share one launch: CTAs [0, ctas0) run recurrence 0, the rest recurrence 1, each with its
// own flags / exchange buffer / weights, and with HS = 32 each needs only 16 CTAs per 128-row group.
template <int HS>
__global__ void __launch_bounds__(kThreads, 1)
rec_bwd_ksplit2_kernel(const __grid_constant__ CUtensorMap tma_b0, const __grid_constant__ CUtensorMap tma_b1, const __grid_constant__ KsplitPair pp) {
  constexpr int KC = 4 * HS;               // gate columns (MMA K) owned by this CTA
  constexpr int KB = KC / BK;              // K blocks of 64
  constexpr int MYCH = HS / 16;            // 8-unit chunks per row thread
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t b_full_bar, a_full_bar, tmem_full_bar, tmem_empty_bar;
  __shared__ uint32_t tmem_base_slot;

  const int which = (int)blockIdx.x >= pp.ctas0 ? 1 : 0;
  const RecKP& p = pp.p[which];
  const CUtensorMap* tma_b = which ? &tma_b1 : &tma_b0;
  const int bid = (int)blockIdx.x - (which ? pp.ctas0 : 0);

  const int H = p.H, G = p.G, T = p.steps;
  const int NH = H <= 256 ? 1 : H / 256, NB = H / NH;              // MMA N chunks
  const uint32_t tm_cols = H <= 32 ? 32 : (H <= 64 ? 64 : (H <= 128 ? 128 : (H <= 256 ? 256 : 512)));
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_b = smem_base;                                // KB x [H][64] bf16 K-major
  const uint32_t smem_a = smem_base + (uint32_t)KB * H * 128;       // KB x [128][64] bf16 K-major
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = p.group0 + bid / p.cpg, j = bid % p.cpg;
  const int row0 = g * BM;
  unsigned* flags = p.flags + (size_t)g * p.flag_stride;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(tma_b);
    ptx::mbar_init(ptx::smem_u32(&b_full_bar), 1);
    ptx::mbar_init(ptx::smem_u32(&a_full_bar), ROW_THREADS);
    ptx::mbar_init(ptx::smem_u32(&tmem_full_bar), 1);
    ptx::mbar_init(ptx::smem_u32(&tmem_empty_bar), EPI_WARPS);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), tm_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {   // resident weights [cpg][KB][H][64]
      const uint32_t bb = ptx::smem_u32(&b_full_bar);
      ptx::mbar_arrive_expect_tx(bb, (uint32_t)KB * H * 128);
      for (int kb = 0; kb < KB; ++kb)
        for (int nh = 0; nh < NH; ++nh) ptx::tma_load_2d(smem_b + (kb * H + nh * NB) * 128, tma_b, bb, 0, (j * KB + kb) * H + nh * NB);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(BM, NB, false, false);
      ptx::mbar_wait(ptx::smem_u32(&b_full_bar), 0);
      for (int it = 0; it < T; ++it) {
        ptx::mbar_wait(ptx::smem_u32(&a_full_bar), it & 1);                 // dG_t tile written by the row warps
        ptx::mbar_wait(ptx::smem_u32(&tmem_empty_bar), (it & 1) ^ 1);       // previous partial drained
        ptx::tc_fence_after();
        for (int nh = 0; nh < NH; ++nh)
#pragma unroll
          for (int kb = 0; kb < KB; ++kb)
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              ptx::umma_bf16(tmem_base + nh * NB, ptx::umma_desc_sw128(smem_a + kb * 16384 + k * 32, 16, 1024),
                             ptx::umma_desc_sw128(smem_b + (kb * H + nh * NB) * 128 + k * 32, 16, 1024), idesc, (kb > 0 || k > 0) ? 1u : 0u);
        ptx::umma_commit(ptx::smem_u32(&tmem_full_bar));
      }
    }
  } else {
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int etid = threadIdx.x - 64;
    const int r = quad * 32 + lane, m = row0 + r;
    const bool row_ok = m < p.n;
    const bool tracer = (etid == 0);
    const int bi = p.variant == MVAE_CELL_STANDARD ? 0 : 1, bfk = 1 - bi;
    bf16* part = (bf16*)p.partial;                                      // [2][groups_total][cpg][H/8 granules][128 rows][8]
    const size_t part_cta = (size_t)BM * H, part_grp = part_cta * p.cpg, part_buf = part_grp * p.groups_total;
    float dc[MYCH * 8];
#pragma unroll
    for (int u = 0; u < MYCH * 8; ++u) dc[u] = 0.f;

    for (int it = 0; it <= T; ++it) {
      const int t = T - 1 - it;
      uint4 sg[MYCH][4], sc0[MYCH], sc1[MYCH], se[MYCH];
      if (t >= 0 && row_ok) {
#pragma unroll
        for (int c = 0; c < MYCH; ++c) {
          const int u0 = j * HS + (2 * c + half) * 8, gu = u0 >> 3;
          sg[c][0] = __ldg(reinterpret_cast<const uint4*>(p.gates + gran_off(t, G / 8, bi * (H / 8) + gu, p.n, m)));
          sg[c][1] = __ldg(reinterpret_cast<const uint4*>(p.gates + gran_off(t, G / 8, bfk * (H / 8) + gu, p.n, m)));
          sg[c][2] = __ldg(reinterpret_cast<const uint4*>(p.gates + gran_off(t, G / 8, 2 * (H / 8) + gu, p.n, m)));
          sg[c][3] = __ldg(reinterpret_cast<const uint4*>(p.gates + gran_off(t, G / 8, 3 * (H / 8) + gu, p.n, m)));
          sc0[c] = __ldg(reinterpret_cast<const uint4*>(p.cseq + gran_off(t, H / 8, gu, p.n, m)));
          sc1[c] = __ldg(reinterpret_cast<const uint4*>(p.cseq + gran_off(t + 1, H / 8, gu, p.n, m)));
          if (p.dhext) se[c] = __ldg(reinterpret_cast<const uint4*>(p.dhext + ((size_t)t * p.n + m) * H + u0));
        }
      }
      // ---- dh_t = sum over the group's CTAs of their partial dG_{t+1} U^T, for this thread's units
      float dh[MYCH * 8];
#pragma unroll
      for (int u = 0; u < MYCH * 8; ++u) dh[u] = 0.f;
      if (it > 0) {
        if (tracer) { REC_TRACE(it, 0); wait_flag(flags + (t + 1), (unsigned)p.cpg); REC_TRACE(it, 1); }
        epi_barrier();
#pragma unroll
        for (int c = 0; c < MYCH; ++c) {
          const int u0 = j * HS + (2 * c + half) * 8;
          const bf16* src = part + (size_t)((it - 1) & 1) * part_buf + (size_t)g * part_grp + ((size_t)(u0 >> 3) * BM + r) * 8;
          uint4 qa[8], qb[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) if (jj < p.cpg) qa[jj] = __ldcg(reinterpret_cast<const uint4*>(src + (size_t)jj * part_cta));
          for (int j0 = 0; j0 < p.cpg; j0 += 16) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) if (j0 + 8 + jj < p.cpg) qb[jj] = __ldcg(reinterpret_cast<const uint4*>(src + (size_t)(j0 + 8 + jj) * part_cta));
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
              if (j0 + jj < p.cpg) {
                float f[8];
                unpack8(qa[jj], f);
#pragma unroll
                for (int u = 0; u < 8; ++u) dh[c * 8 + u] += f[u];
              }
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) if (j0 + 16 + jj < p.cpg) qa[jj] = __ldcg(reinterpret_cast<const uint4*>(src + (size_t)(j0 + 16 + jj) * part_cta));
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
              if (j0 + 8 + jj < p.cpg) {
                float f[8];
                unpack8(qb[jj], f);
#pragma unroll
                for (int u = 0; u < 8; ++u) dh[c * 8 + u] += f[u];
              }
          }
        }
        if (tracer) REC_TRACE(it, 2);
      }
      if (t < 0) {
        if (row_ok && p.dS_h) {
#pragma unroll
          for (int c = 0; c < MYCH; ++c) {
            const int u0 = j * HS + (2 * c + half) * 8;
            *reinterpret_cast<uint4*>(p.dS_h + (size_t)m * p.ldS + u0) = pack8(&dh[c * 8]);
            *reinterpret_cast<uint4*>(p.dS_c + (size_t)m * p.ldS + u0) = pack8(&dc[c * 8]);
          }
        }
        break;
      }
      // ---- gate-gradient math for step t; operand A row r = [di(HS) | df(HS) | dg(HS) | do(HS)] in 16-byte chunks
      uint4 pk[MYCH][4];
#pragma unroll
      for (int c = 0; c < MYCH; ++c) {
        const int u0 = j * HS + (2 * c + half) * 8;
        float gi[8], gf[8], gg[8], go[8], c0[8], c1[8], ex[8], di[8], df[8], dg[8], dob[8];
        unpack8(sg[c][0], gi); unpack8(sg[c][1], gf); unpack8(sg[c][2], gg); unpack8(sg[c][3], go); unpack8(sc0[c], c0); unpack8(sc1[c], c1);
        if (p.dhext) unpack8(se[c], ex);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float d = dh[c * 8 + u];
          if (p.dhext) d += ex[u];
          if (it == 0 && p.dh_last) d += row_ok ? __bfloat162float(p.dh_last[(size_t)m * p.ld_last + u0 + u]) : 0.f;
          float d_o, ds;
          if (p.variant == MVAE_CELL_STANDARD) {
            const float tc = tanh_fast(c1[u]);
            d_o = d * tc;
            ds = dc[c * 8 + u] + d * go[u] * (1.f - tc * tc);
          } else {
            d_o = d * c1[u];
            ds = (dc[c * 8 + u] + d * go[u]) * (1.f - c1[u] * c1[u]);
          }
          di[u] = ds * gg[u] * gate_bwd(p.gate_act, gi[u]);
          df[u] = ds * c0[u] * gate_bwd(p.gate_act, gf[u]);
          dg[u] = ds * gi[u] * (1.f - gg[u] * gg[u]);
          dob[u] = d_o * gate_bwd(p.gate_act, go[u]);
          dc[c * 8 + u] = ds * gf[u];
          if (!row_ok) { di[u] = 0.f; df[u] = 0.f; dg[u] = 0.f; dob[u] = 0.f; dc[c * 8 + u] = 0.f; }
        }
        pk[c][0] = pack8(di); pk[c][1] = pack8(df); pk[c][2] = pack8(dg); pk[c][3] = pack8(dob);
        const int sw = r & 7;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = q * (HS / 8) + (2 * c + half);            // 16-byte chunk index along K (0 .. KC/8-1)
          st_shared16(smem_a + (chunk >> 3) * 16384 + r * 128 + (((chunk & 7) ^ sw) << 4), pk[c][q]);
        }
      }
      ptx::fence_proxy_async();                                     // generic-proxy smem writes -> tensor core (async proxy)
      ptx::mbar_arrive(ptx::smem_u32(&a_full_bar));
      // ---- partial dh_{t-1} for all H units: TMEM -> bf16 -> this CTA's slot of the exchange buffer
      if (tracer) REC_TRACE(it, 5);
      ptx::mbar_wait(ptx::smem_u32(&tmem_full_bar), it & 1);
      ptx::tc_fence_after();
      if (tracer) REC_TRACE(it, 6);
      {
        bf16* dst = part + (size_t)(it & 1) * part_buf + (size_t)g * part_grp + (size_t)j * part_cta;
        const int nch = H / 64;                                      // 32-column chunks in this thread's half
        for (int c = 0; c < nch; ++c) {
          float v[32];
          const int col0 = half * (H / 2) + c * 32;
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)col0, v);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) *reinterpret_cast<uint4*>(dst + ((size_t)((col0 >> 3) + q4) * BM + r) * 8) = pack8(&v[q4 * 8]);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&tmem_empty_bar));
      if (tracer) REC_TRACE(it, 7);
      epi_barrier();
      if (tracer) {
        REC_TRACE(it, 9);
        __threadfence();
        REC_TRACE(it, 10);
        red_relaxed_add(flags + t, 1u);
        REC_TRACE(it, 11);
      }
      if (row_ok) {   // dG_t for the batched weight-gradient GEMMs: row-major (16-byte pieces), off the critical path
#pragma unroll
        for (int c = 0; c < MYCH; ++c) {
          bf16* dgp = p.dG + ((size_t)t * p.n + m) * G + j * HS + (2 * c + half) * 8;
          *reinterpret_cast<uint4*>(dgp + bi * H) = pk[c][0];
          *reinterpret_cast<uint4*>(dgp + bfk * H) = pk[c][1];
          *reinterpret_cast<uint4*>(dgp + 2 * H) = pk[c][2];
          *reinterpret_cast<uint4*>(dgp + 3 * H) = pk[c][3];
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, tm_cols);
}

// U (H, 4H) fp32 -> packed bf16 [cpg][KB][H units][64] K-major for the general K-split backward:
// K index kk = gate*HS + uu of CTA j (K block kk/64, column kk%64) holds U[k, blk(gate)*H + j*HS + uu]
__global__ void pack_u_bwd2_kernel(const float* __restrict__ U, int ldu, bf16* __restrict__ out, int H, int HS, int variant) {
  const long total = (long)H * 4 * H;
  const int KC = 4 * HS, KB = KC / 64;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int col = (int)(e % 64);
    long rest = e / 64;
    const int k = (int)(rest % H); rest /= H;
    const int kb = (int)(rest % KB);
    const int j = (int)(rest / KB);
    const int kk = kb * 64 + col;
    const int gate = kk / HS, uu = kk % HS;
    int blk = gate;
    if (variant != MVAE_CELL_STANDARD && gate < 2) blk = 1 - gate;
    out[e] = __float2bfloat16_rn(U[(long)k * ldu + blk * H + j * HS + uu]);
  }
}

// U (H, 4H) fp32 -> per-CTA packed bf16 [cpg][H units][64] K-major for the K-split backward:
// column kk = gate*16 + uu of CTA j's block holds U[k, blk(gate)*H + j*16 + uu]
__global__ void pack_u_bwd_kernel(const float* __restrict__ U, int ldu, bf16* __restrict__ out, int H, int variant) {
  const long total = (long)H * 4 * H;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int kk = (int)(e % 64);
    const long rk = e / 64;               // j*H + k
    const int k = (int)(rk % H), j = (int)(rk / H);
    const int gate = kk / 16, uu = kk % 16;
    int blk = gate;
    if (variant != MVAE_CELL_STANDARD && gate < 2) blk = 1 - gate;
    out[e] = __float2bfloat16_rn(U[(long)k * ldu + blk * H + j * 16 + uu]);
  }
}

// U (H, 4H) fp32 master, natural Keras column blocks -> per-CTA packed bf16 [cpg][4*HS][H], K-major:
// row n = ublock*32 + gate*8 + u8 of CTA j holds column blk(gate)*H + j*HS + ublock*8 + u8 of U.
__global__ void pack_u_kernel(const float* __restrict__ U, int ldu, bf16* __restrict__ out, int H, int HS, int variant) {
  const int BN = 4 * HS;
  const long total = (long)H * 4 * H;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int k = (int)(e % H);
    const long rn = e / H;            // j*BN + n
    const int n = (int)(rn % BN), j = (int)(rn / BN);
    const int ub = n / 32, gate = (n % 32) / 8, u8 = n % 8;
    int blk = gate;
    if (variant != MVAE_CELL_STANDARD && gate < 2) blk = 1 - gate;
    const int col = blk * H + j * HS + ub * 8 + u8;
    out[e] = __float2bfloat16_rn(U[(long)k * ldu + col]);
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MVAE_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    MVAE_REQUIRE(p != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    fn = (EncodeFn)p;
  }
  return fn;
}
CUtensorMap make_map(const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstr[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MVAE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (rec_persist)");
  return m;
}

template <bool FWD, int HS>
RecKP rec_params(const RecPersistArgs& a, int stages, int prod_lanes) {
  const int H = a.H, G = 4 * H;
  RecKP p{};
  p.n = a.n; p.H = H; p.G = G; p.steps = a.steps; p.cpg = H / HS; p.HS = HS; p.stages = stages; p.prod_lanes = prod_lanes;
  { static int rot = -1; if (rot < 0) { const char* e = getenv("MVAE_REC_ROT"); rot = e ? atoi(e) : 1; } p.kb_rot = rot; }
  p.gate_act = a.gate_act; p.variant = a.variant;
  p.flags = a.flags; p.flag_stride = a.steps + 2;
  p.xw = (const bf16*)a.xw; p.hseq = (bf16*)a.hseq; p.cseq = (bf16*)a.cseq; p.gates = (bf16*)a.gates; p.c0 = (const bf16*)a.c0; p.ldc0 = a.ldc0;
  p.dhext = (const bf16*)a.dhext; p.dh_last = (const bf16*)a.dh_last; p.ld_last = a.ld_last; p.dG = (bf16*)a.dG;
  p.dS_h = (bf16*)a.dS_h; p.dS_c = (bf16*)a.dS_c; p.ldS = a.ldS;
  p.trace = (long long*)a.trace;
  p.hx = (bf16*)a.hx; p.groups_total = (a.n + BM - 1) / BM;
  return p;
}

// b != nullptr: a second, independent recurrence of the same hidden size shares the launch (forward only)
template <bool FWD, int HS>
void launch(const RecPersistArgs& a, const RecPersistArgs* b, cudaStream_t st, int sm_count) {
  constexpr int BN = FWD ? 4 * HS : HS;
  const int H = a.H, G = 4 * H, K = FWD ? H : G, kblocks = K / BK;
  const int cpg = H / HS;
  const int groups = (a.n + BM - 1) / BM;
  const size_t b_bytes = (size_t)kblocks * BN * BK * 2;
  int stages = (int)std::min<size_t>(MAX_STAGES, (smem_max_bytes() - 1024 - b_bytes) / A_STAGE_BYTES);
  MVAE_REQUIRE(stages >= 2, "recurrent weight slice leaves no room for the operand ring");
  // every ring slot must always be refilled by the SAME producer lane (parity waits cannot tell phase k from k+2),
  // i.e. the slot count is a multiple of the lane count
  if (stages >= PROD_LANES) stages = stages / PROD_LANES * PROD_LANES;
  const int prod_lanes = stages >= PROD_LANES ? PROD_LANES : stages;
  const size_t smem = b_bytes + (size_t)stages * A_STAGE_BYTES + 1024;
  auto kern = rec_persist_kernel<FWD, HS>;
  MVAE_REQUIRE(smem <= smem_max_bytes(), "recurrent weight slice does not fit in shared memory");
  static size_t attr_smem = 0;
  if (smem > attr_smem) { MVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem = smem; }
  MVAE_REQUIRE(cpg <= sm_count, "hidden size too large for a co-resident group");
  if (FWD) MVAE_REQUIRE(a.hx != nullptr && (!b || b->hx != nullptr), "persistent forward needs the h exchange buffer");
  KsplitPair pp{};
  pp.p[0] = rec_params<FWD, HS>(a, stages, prod_lanes);
  MVAE_CUDA(cudaMemsetAsync(a.flags, 0, (size_t)groups * pp.p[0].flag_stride * sizeof(unsigned), st));
  // A (backward N-split only): dG as [steps*n, 4H]; the forward reads the exchange buffer with bulk copies
  const CUtensorMap ma = FWD ? make_map(a.hseq, H, (uint64_t)(a.steps + 1) * a.n, H, 64, BM) : make_map(a.dG, G, (uint64_t)a.steps * a.n, G, 64, BM);
  // B: forward  = packed U [cpg*4HS, H] ; backward = U shadow [H, 4H] (rows = this CTA's units)
  const CUtensorMap mb = FWD ? make_map(a.upack, H, (uint64_t)cpg * BN, H, 64, BN) : make_map(a.u_shadow, G, H, a.ldu, 64, BN);
  if (b) {
    MVAE_REQUIRE(FWD && b->H == H, "only forward recurrences of one hidden size can share a launch");
    const int gb = (b->n + BM - 1) / BM;
    MVAE_REQUIRE((groups + gb) * cpg <= sm_count, "paired recurrences must be co-resident");
    pp.p[1] = rec_params<FWD, HS>(*b, stages, prod_lanes);
    MVAE_CUDA(cudaMemsetAsync(b->flags, 0, (size_t)gb * pp.p[1].flag_stride * sizeof(unsigned), st));
    const CUtensorMap mb1 = make_map(b->upack, H, (uint64_t)cpg * BN, H, 64, BN);
    pp.p[0].group0 = 0; pp.p[1].group0 = 0; pp.ctas0 = groups * cpg;
    kern<<<(groups + gb) * cpg, kThreads, smem, st>>>(ma, mb, mb1, pp);
    count_launch();
    MVAE_CUDA(cudaGetLastError());
    return;
  }
  const int gmax = std::max(1, sm_count / cpg);   // groups that can be co-resident
  for (int g0 = 0; g0 < groups; g0 += gmax) {
    const int ng = std::min(gmax, groups - g0);
    pp.p[0].group0 = g0; pp.ctas0 = ng * cpg;
    kern<<<ng * cpg, kThreads, smem, st>>>(ma, mb, mb, pp);
    count_launch();
    MVAE_CUDA(cudaGetLastError());
  }
}

}  // namespace

size_t smem_max_bytes() { return 227 * 1024 - 1024; }   // 227 KB opt-in limit minus the kernels' static shared memory (barriers)

// Hidden units per CTA.  Smaller slices leave more shared memory for the operand ring (the per-step gather of
// h / dG from L2 is latency-bound: bytes in flight per SM set its rate) and spread a group over more SMs.
int rec_persist_hs(int H) {
  if (H % 64 != 0) return 0;
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("MVAE_REC_HS"); forced = e ? atoi(e) : 0; }
  for (int hs : {16, 32}) {
    if (forced && hs != forced) continue;
    if (H % hs) continue;
    const size_t fwd = (size_t)(H / 64) * (4 * hs) * 64 * 2, bwd = (size_t)(4 * H / 64) * hs * 64 * 2;
    if (std::max(fwd, bwd) + 4 * A_STAGE_BYTES + 1024 <= smem_max_bytes()) return hs;
  }
  return 0;
}

bool rec_persist_supported(int H, int sm_count) {
  const int hs = rec_persist_hs(H);
  return hs > 0 && H / hs <= sm_count;
}

size_t rec_persist_hx_bytes(int n, int H) { return 2 * (size_t)((n + BM - 1) / BM) * BM * H * 2; }

size_t rec_persist_flag_count(int n, int steps) { return (size_t)((n + BM - 1) / BM) * (steps + 2); }

void rec_persist_pack_u(const float* U, int ldu, void* upack, int H, int hs, int variant, cudaStream_t st) {
  if (hs == 0) hs = rec_persist_hs(H);
  MVAE_REQUIRE(hs == 16 || hs == 32, "persistent recurrence unsupported for this hidden size");
  pack_u_kernel<<<std::min(148 * 8, (int)(((long)H * 4 * H + 255) / 256)), 256, 0, st>>>(U, ldu, (bf16*)upack, H, hs, variant);
  count_launch();
  MVAE_CUDA(cudaGetLastError());
}

void rec_persist_forward(const RecPersistArgs& a, cudaStream_t st, int sm_count) {
  switch (rec_persist_hs(a.H)) {
    case 32: launch<true, 32>(a, nullptr, st, sm_count); break;
    case 16: launch<true, 16>(a, nullptr, st, sm_count); break;
    default: throw Error("persistent recurrence unsupported for this hidden size");
  }
}

// units per CTA with which TWO forward recurrences of batch n fit into one co-resident launch (0 = they do not)
int rec_persist_fwd_pair_hs(int H, int n, int sm_count) {
  if (H % 64) return 0;
  const int groups = (n + BM - 1) / BM;
  for (int hs : {16, 32}) {
    if (H % hs) continue;
    const size_t b_bytes = (size_t)(H / 64) * (4 * hs) * 64 * 2;
    if (b_bytes + 4 * A_STAGE_BYTES + 1024 > smem_max_bytes()) continue;
    if (2 * groups * (H / hs) <= sm_count) return hs;
  }
  return 0;
}

void rec_persist_forward_pair(const RecPersistArgs& a, const RecPersistArgs& b, int HS, cudaStream_t st, int sm_count) {
  if (HS == 32) launch<true, 32>(a, &b, st, sm_count);
  else launch<true, 16>(a, &b, st, sm_count);
}

bool rec_persist_ksplit_ok(int H) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("MVAE_REC_KSPLIT"); enabled = e ? atoi(e) : 1; }
  if (!enabled || rec_persist_hs(H) != 16 || H > 512) return false;
  const int NH = H <= 256 ? 1 : H / 256;
  return (H <= 256 || H % 256 == 0) && (H / NH) % 16 == 0;
}

size_t rec_persist_partial_bytes(int n, int H) {
  if (!rec_persist_ksplit_ok(H)) return 0;
  const size_t groups = (n + BM - 1) / BM;
  return 2 * groups * (size_t)(H / 16) * BM * H * 2;
}

void rec_persist_pack_u_bwd(const float* U, int ldu, void* upack_bwd, int H, int HS, int variant, cudaStream_t st) {
  MVAE_REQUIRE(HS == 16 || HS == 32, "K-split backward: 16 or 32 units per CTA");
  pack_u_bwd2_kernel<<<std::min(148 * 8, (int)(((long)H * 4 * H + 255) / 256)), 256, 0, st>>>(U, ldu, (bf16*)upack_bwd, H, HS, variant);
  count_launch();
  MVAE_CUDA(cudaGetLastError());
}

static size_t ksplit_smem(int H, int HS) { return (size_t)(4 * HS / 64) * ((size_t)H * 128 + 16384) + 1024; }

// units per CTA with which TWO recurrences of batch n fit into one co-resident launch (0 = they do not)
int rec_persist_pair_hs(int H, int n, int sm_count) {
  if (!rec_persist_ksplit_ok(H)) return 0;
  const int groups = (n + BM - 1) / BM;
  for (int hs : {16, 32}) {
    if (H % hs) continue;
    if (ksplit_smem(H, hs) > smem_max_bytes()) continue;
    if (2 * groups * (H / hs) <= sm_count) return hs;
  }
  return 0;
}

static RecKP ksplit_params(const RecPersistArgs& a, int HS, int groups) {
  RecKP p{};
  p.n = a.n; p.H = a.H; p.G = 4 * a.H; p.steps = a.steps; p.cpg = a.H / HS; p.HS = HS; p.gate_act = a.gate_act; p.variant = a.variant;
  p.flags = a.flags; p.flag_stride = a.steps + 2;
  p.gates = (bf16*)a.gates; p.cseq = (bf16*)a.cseq;
  p.dhext = (const bf16*)a.dhext; p.dh_last = (const bf16*)a.dh_last; p.ld_last = a.ld_last; p.dG = (bf16*)a.dG;
  p.dS_h = (bf16*)a.dS_h; p.dS_c = (bf16*)a.dS_c; p.ldS = a.ldS;
  p.trace = (long long*)a.trace; p.partial = a.partial; p.groups_total = groups;
  return p;
}

template <int HS>
static void launch_bwd_ksplit2(const RecPersistArgs& a, const RecPersistArgs* b, cudaStream_t st, int sm_count) {
  const int H = a.H, cpg = H / HS, KB = 4 * HS / 64;
  const int NH = H <= 256 ? 1 : H / 256, NB = H / NH;
  const size_t smem = ksplit_smem(H, HS);
  MVAE_REQUIRE(smem <= smem_max_bytes(), "K-split backward: weight slice does not fit in shared memory");
  auto kern = rec_bwd_ksplit2_kernel<HS>;
  static size_t attr_smem = 0;
  if (smem > attr_smem) { MVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem = smem; }
  MVAE_REQUIRE(a.upack_bwd && a.partial && (!b || (b->upack_bwd && b->partial && b->H == H)), "K-split backward needs packed weights and exchange buffers");
  const int ga = (a.n + BM - 1) / BM, gb = b ? (b->n + BM - 1) / BM : 0;
  KsplitPair pp{};
  pp.p[0] = ksplit_params(a, HS, ga);
  MVAE_CUDA(cudaMemsetAsync(a.flags, 0, (size_t)ga * pp.p[0].flag_stride * sizeof(unsigned), st));
  const CUtensorMap ma = make_map(a.upack_bwd, 64, (uint64_t)cpg * KB * H, 64, 64, NB);
  if (b) {
    pp.p[1] = ksplit_params(*b, HS, gb);
    MVAE_CUDA(cudaMemsetAsync(b->flags, 0, (size_t)gb * pp.p[1].flag_stride * sizeof(unsigned), st));
    const CUtensorMap mbm = make_map(b->upack_bwd, 64, (uint64_t)cpg * KB * H, 64, 64, NB);
    MVAE_REQUIRE((ga + gb) * cpg <= sm_count, "paired recurrences must be co-resident");
    pp.p[0].group0 = 0; pp.p[1].group0 = 0; pp.ctas0 = ga * cpg;
    kern<<<(ga + gb) * cpg, kThreads, smem, st>>>(ma, mbm, pp);
    count_launch();
    MVAE_CUDA(cudaGetLastError());
    return;
  }
  const int gmax = std::max(1, sm_count / cpg);
  for (int g0 = 0; g0 < ga; g0 += gmax) {
    const int ng = std::min(gmax, ga - g0);
    pp.p[0].group0 = g0; pp.ctas0 = ng * cpg;
    kern<<<ng * cpg, kThreads, smem, st>>>(ma, ma, pp);
    count_launch();
    MVAE_CUDA(cudaGetLastError());
  }
}

void rec_persist_backward_pair(const RecPersistArgs& a, const RecPersistArgs* b, int HS, cudaStream_t st, int sm_count) {
  if (HS == 32) launch_bwd_ksplit2<32>(a, b, st, sm_count);
  else launch_bwd_ksplit2<16>(a, b, st, sm_count);
}

void rec_persist_backward(const RecPersistArgs& a, cudaStream_t st, int sm_count) {
  if (rec_persist_ksplit_ok(a.H) && a.upack_bwd && a.partial) { rec_persist_backward_pair(a, nullptr, 16, st, sm_count); return; }
  switch (rec_persist_hs(a.H)) {
    case 32: launch<false, 32>(a, nullptr, st, sm_count); break;
    case 16: launch<false, 16>(a, nullptr, st, sm_count); break;
    default: throw Error("persistent recurrence unsupported for this hidden size");
  }
}

}  // namespace mvae
