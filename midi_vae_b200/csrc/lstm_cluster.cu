// lstm_cluster.cu -- cluster-resident persistent-RNN kernels for the LSTM recurrences (bf16, H = 256 / 512).
//
// Same job as lstm_persist.cu (one launch = all timesteps of one recurrence of vae_definition.py:455-474,533-632, the
// loop the reference runs as a Theano scan), but the per-step exchange of the hidden state never leaves the SMs:
//
//   * the batch is cut into groups of 64 rows; one THREAD-BLOCK CLUSTER of CS = H/32 CTAs owns a group for the whole
//     sequence.  Clusters are independent of each other (no grid-wide flag, no co-residency requirement);
//   * CTA j of the cluster owns 32 hidden units (all four gates = 128 gate columns).  Its slice of the recurrent
//     weights, [128 gate columns][H] bf16, stays resident in shared memory as the A operand (M side) of
//         gates^T [gate column, batch row] = U^T [gate column, k] * h_{t-1}^T [k, batch row]
//     so the operand that changes every step is the SMALL one (N = 64 batch rows);
//   * CTAs (2q, 2q+1) form a CTA pair: tcgen05.mma.cta_group::2 (M = 256, N = 64) reads 128 rows of A from each CTA
//     and HALF of B from each CTA, so every CTA only has to receive the h tile of 32 batch rows (32 x H bf16 = 32 KB at
//     H = 512) per step instead of the group's whole h;
//   * the epilogue (8 warps) moves the accumulator from TMEM (lane = gate column) through a swizzled fp32 scratch tile
//     into a (batch row, 8 units) ownership, applies the gate math with c in registers for the whole sequence, writes
//     the CTA's new h slice (64 rows x 32 units) into a staging tile in the consumers' operand layout and PUSHES it with
//     bulk copies (cp.async.bulk shared::cta -> shared::cluster) into the h buffers of all CS CTAs; the copies complete
//     transaction bytes on the consumers' mbarriers, so there is no flag, no fence to L2 and no polling: a step is
//     MMA -> TMEM -> gate math -> DSMEM push -> MMA;
//   * h buffers and the staging tile are double-buffered; every re-use is ordered by the data dependence itself
//     (nobody can produce h_{t+1} before having received all of h_t), see the comments at each buffer.
//
// The stash (gates, c, h) is written in the layouts of lstm_persist.cu, so either backward kernel can follow.
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

constexpr int CL_ROWS = 64;                 // batch rows per cluster (UMMA N of the pair)
constexpr int CL_HALF = 32;                 // batch rows whose operand tile one CTA of a pair holds
constexpr int CL_HS = 32;                   // hidden units per CTA
constexpr int CL_GC = 4 * CL_HS;            // gate columns per CTA = UMMA M per CTA
constexpr int CL_EPI_WARPS = 8;
constexpr int CL_THREADS = 32 * (2 + CL_EPI_WARPS);
constexpr uint32_t CL_SLICE = 4 * CL_HALF * 16;   // one pushed piece: 4 k-granules x 32 rows x 16 B = 2 KB
constexpr uint32_t CL_STAGE = 2 * CL_SLICE;       // staging tile: both row halves of the CTA's h slice
constexpr uint32_t CL_SCR = CL_ROWS * CL_GC * 4;  // fp32 scratch tile [64 rows][128 gate columns]

#define CL_TRACE(step, point)                                                                                              \
  do {                                                                                                                     \
    if (p.trace && blockIdx.x < 2 && (step) >= 16 && (step) < 24) p.trace[(blockIdx.x * 8 + (step) - 16) * 16 + (point)] = clock64(); \
  } while (0)

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool HARD>
__device__ __forceinline__ float gate_fwd(float x) {
  return HARD ? __saturatef(0.2f * x + 0.5f) : 0.5f * tanh_fast(0.5f * x) + 0.5f;
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
// stash layout shared with lstm_persist.cu: [slab][column/8 granules][n rows][8]
__device__ __forceinline__ size_t gran_off(int slab, int ngran, int gran, int n, int m) {
  return (((size_t)slab * ngran + gran) * n + m) * 8;
}
// bar.sync is the ALIGNED barrier: every lane of an arriving warp must execute it together.  Warps get here out of spin loops and lane-predicated
// blocks, so convergence is made explicit first (compute-sanitizer synccheck reported divergent arrivals without it).
__device__ __forceinline__ void named_barrier(int id, int count) {
  __syncwarp();
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// Forward.  What the first form of this kernel (round 1, U in shared memory, DSMEM pushes; deleted) taught (profiles/r1/README.md):
//   * an SS-mode MMA with N = 64 is paced by the fetch of its 4 KB A tile, so U lives in TENSOR MEMORY (A operand from
//     TMEM, 32 x (M256 N64 K16) per step) and shared memory is free for operand tiles;
//   * the DSMEM all-to-all is the slow path (6-13 B/clk/SM measured for st.async / bulk copies); the exchange goes
//     through L2 instead: every CTA stores its 4 KB h slice to a global exchange buffer and ONE multicast bulk copy per
//     row half brings it back into the operand tiles of the 8 CTAs that hold that half (complete_tx on their barriers);
//   * a step is still a latency chain (store -> multicast -> relay -> MMA -> TMEM -> gate math), so every cluster
//     carries NG = 2..3 independent 64-row groups and rotates through them: while one group's h is in flight the
//     epilogue warps work on the next group.  U is shared by the groups; each has its own accumulator and h tiles.
constexpr int CL2_MAXG = 3;
constexpr uint32_t CL2_TM_U = 64 * CL2_MAXG;   // TMEM: accumulator of group g at columns 64 g, U from column 192

struct Cluster2P {
  int n, steps, nswap, ng;
  const bf16* xw; bf16* hseq; bf16* cseq; bf16* gates; const bf16* c0; int ldc0;
  const bf16* upack;
  uint8_t* hx;        // exchange buffer [clusters][ng][2][CS][2 row halves][2 KB]
  int no_stash;       // inference: h only, no BPTT stash
  int x_mode; const bf16* xtab; const unsigned char* x_idx; int x_ld, x_shift; const bf16* x_scalar; const float* x_w; const float* x_b;
  long long* trace;
};

// Warp roles (19 warps): 0 = MMA issuer (even CTA) / relay (odd CTA), 1 / 2 = multicast warp of epilogue set 0 / 1 (warp 1 also
// owns the TMEM allocation), 3..10 = epilogue set 0, 11..18 = epilogue set 1.  Set s works on groups s, s+2: the epilogue
// is instruction-bound, so two groups are in their epilogue at the same time (4 warps per scheduler instead of 2).
constexpr int CLF_THREADS = 32 * (3 + 2 * CL_EPI_WARPS);

template <int CS, bool HARD, bool STD>
__global__ void __launch_bounds__(CLF_THREADS, 1)
rec_cluster_fwd2_kernel(const Cluster2P p) {
  constexpr int H = CS * CL_HS, G = 4 * H, KS = H / 16;
  constexpr uint32_t HBUF = (uint32_t)CL_HALF * H * 2;
  constexpr bool ALIAS = (HBUF >= CL_SCR);      // the scratch tile of an item lives in the h tile its MMA has just finished reading
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t h_full[CL2_MAXG][2], peer_ready[CL2_MAXG][2], tmem_full[CL2_MAXG];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float x_wb[2][CL_GC];   // x_mode 2: input kernel row / bias of this CTA's 128 gate columns (gate-major, semantic gate order)

  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_scr0 = smem_base;                              // !ALIAS: scratch tile of set s at smem_scr0 + s * CL_SCR
  const uint32_t smem_h0 = smem_base + (ALIAS ? 0u : 2u * CL_SCR);   // hbuf(g, b) = smem_h0 + (2 g + b) * HBUF
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int cl = (int)blockIdx.x / CS;
  const int e = (int)(rank & 1);
  const int rh = e ^ p.nswap;
  const int j = (int)rank;
  const int T = p.steps, n = p.n, ng = p.ng;
  const int row0 = cl * CL_ROWS * ng;
  const int nga = min(ng, (n - row0 + CL_ROWS - 1) / CL_ROWS);      // groups of this cluster that hold rows

  if (threadIdx.x == 0) {
    for (int g = 0; g < CL2_MAXG; ++g) {
      for (int b = 0; b < 2; ++b) { ptx::mbar_init(ptx::smem_u32(&h_full[g][b]), 1); ptx::mbar_init(ptx::smem_u32(&peer_ready[g][b]), 1); }
      ptx::mbar_init(ptx::smem_u32(&tmem_full[g]), 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc2(ptx::smem_u32(&tmem_base_slot), 512);
    ptx::tmem_relinquish2();
  }
  if (p.x_mode == 2 && threadIdx.x >= 96 && threadIdx.x < 96 + CL_GC) {
    const int cc = (int)threadIdx.x - 96, gt = cc >> 5, uu = cc & 31;
    const int blk = (gt < 2 && !STD) ? 1 - gt : gt;
    x_wb[0][cc] = p.x_w[blk * H + j * CL_HS + uu];
    x_wb[1][cc] = p.x_b[blk * H + j * CL_HS + uu];
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp >= 3 && warp < 7) {
    // U slice -> tensor memory: lane = gate column, 32-bit column c = k pair (2c, 2c+1); each thread streams its own row
    const int mrow = (warp & 3) * 32 + lane;
    const uint4* src = reinterpret_cast<const uint4*>(p.upack + ((size_t)j * CL_GC + mrow) * H);
    for (int c32 = 0; c32 < H / 64; ++c32) {
      uint32_t r[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 v = __ldg(src + c32 * 8 + i);
        r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
      }
      ptx::tmem_st_32x32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + CL2_TM_U + (uint32_t)c32 * 32u, r);
    }
    ptx::tmem_st_wait();
  }
  if (warp >= 3) {
    const int etid = (int)threadIdx.x - 96;   // 0..511 over both sets
    for (int g = 0; g < nga; ++g) {
      // initial hidden state (hseq slab 0, row-major) -> hbuf(g, 1) in operand order
      for (int idx = etid; idx < CL_HALF * (H / 8); idx += 64 * CL_EPI_WARPS) {
        const int r = idx / (H / 8), gk = idx % (H / 8);
        const int mm = row0 + g * CL_ROWS + rh * CL_HALF + r;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (mm < n) v = *reinterpret_cast<const uint4*>(p.hseq + (size_t)mm * H + gk * 8);
        ptx::st_shared_u4(smem_h0 + (2 * g + 1) * HBUF + gk * (CL_HALF * 16) + r * 16, v);
      }
    }
    ptx::fence_proxy_async();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  ptx::tc_fence_after();

  if (warp == 0) {
    if (e == 0) {
      // ===================== MMA issuer (even CTA of the pair) =====================
      // the whole warp walks the loop so that descriptors and addresses stay warp-uniform (uniform datapath); one elected lane issues
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(256, CL_ROWS, false, false);
      const uint16_t pair_mask = (uint16_t)(3u << rank);
      const uint64_t bd0 = ptx::umma_desc_noswz(smem_h0, CL_HALF * 16, 128);
      for (int t = 0; t < T; ++t) {
        const int b = (t + 1) & 1;
#pragma unroll
        for (int g = 0; g < CL2_MAXG; ++g) {
          if (g >= nga) break;
          if (t + 1 < T && lane == 0) ptx::mbar_arrive_expect_tx(ptx::smem_u32(&h_full[g][t & 1]), HBUF);
          if (t > 0) ptx::mbar_wait(ptx::smem_u32(&h_full[g][b]), (uint32_t)(((t - 1) >> 1) & 1));
          if (g == 0 && lane == 0) CL_TRACE(t, 7);
          ptx::mbar_wait(ptx::smem_u32(&peer_ready[g][b]), (uint32_t)((t >> 1) & 1));
          ptx::tc_fence_after();
          if (g == 0 && lane == 0) CL_TRACE(t, 0);
          const uint64_t bd = bd0 + (uint64_t)(((2 * g + b) * HBUF) >> 4);
          const uint32_t d_tmem = tmem_base + (uint32_t)g * 64u, a_tmem = tmem_base + CL2_TM_U;
          if (ptx::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
              ptx::umma_bf16_2cta_ts(d_tmem, a_tmem + (uint32_t)ks * 8u, bd + (uint64_t)(ks * 2 * (CL_HALF * 16) >> 4), idesc, ks > 0 ? 1u : 0u);
            ptx::umma_commit_2cta(ptx::smem_u32(&tmem_full[g]), pair_mask);
          }
          __syncwarp();
          if (g == 0 && lane == 0) CL_TRACE(t, 1);
        }
      }
    } else if (lane == 0) {
      // ===================== relay (odd CTA) =====================
      const uint32_t leader = rank - 1;
      for (int g = 0; g < nga; ++g) ptx::mbar_arrive_remote_relaxed(ptx::mapa(ptx::smem_u32(&peer_ready[g][1]), leader));
      for (int t = 0; t + 1 < T; ++t) {
        for (int g = 0; g < nga; ++g) {
          const uint32_t hb = ptx::smem_u32(&h_full[g][t & 1]);
          ptx::mbar_arrive_expect_tx(hb, HBUF);
          ptx::mbar_wait(hb, (uint32_t)((t >> 1) & 1));
          if (g == 0) CL_TRACE(t + 1, 7);
          ptx::mbar_arrive_remote_relaxed(ptx::mapa(ptx::smem_u32(&peer_ready[g][t & 1]), leader));
          if (g == 0) CL_TRACE(t + 1, 8);
        }
      }
    }
  } else if (warp <= 2) {
    // ===================== multicast warp of set s: brings this CTA's freshly stored h slice into every consumer's operand tile =====================
    const int set = warp - 1;
    for (int t = 0; t + 1 < T; ++t) {
      for (int g = set; g < nga; g += 2) {
        named_barrier(2 + 2 * set, 32 * (CL_EPI_WARPS + 1));
        if (lane < 2) {
          const uint8_t* src = p.hx + ((size_t)(((cl * ng + g) * 2 + (t & 1)) * CS + j)) * CL_STAGE + (size_t)lane * CL_SLICE;
          const uint16_t mask = (uint16_t)((((lane ^ p.nswap) & 1) ? 0xAAAAu : 0x5555u) & ((1u << CS) - 1u));
          ptx::bulk_load_multicast(smem_h0 + (2 * g + (t & 1)) * HBUF + (uint32_t)j * CL_SLICE, src, CL_SLICE, ptx::smem_u32(&h_full[g][t & 1]), mask);
        }
        if (lane == 0 && g == 0) CL_TRACE(t, 6);
      }
    }
  } else {
    // ===================== epilogue warps: set s = warps 3 + 8 s .. 10 + 8 s =====================
    const int set = (warp - 3) >> 3, ew = (warp - 3) & 7;
    // phase A ownership: TMEM lane = gate column c = gate * 32 + unit, 32 batch rows (column half ch)
    const int wq = warp & 3, ch = ew >> 2;
    const int c = wq * 32 + lane;
    // phase B ownership: batch row rr of the group, units [8 q, 8 q + 8) of the CTA's 32; a warp = 8 rows x 4 q, so that every
    // global access of the warp covers whole 64-byte (row-major tensors) or 128-byte (granule-major stash) runs
    const int rr = ew * 8 + (lane >> 2), q = lane & 3;
    const int u0 = j * CL_HS + q * 8, gu = u0 >> 3;
    constexpr int bi = STD ? 0 : 1, bfk = 1 - bi;
    const bool tracer = (ew == 0 && lane == 0);
    const uint32_t swB = (uint32_t)(rr & 7) << 4;
    float cst[2][8];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int g = set + 2 * k;
      const int m = row0 + g * CL_ROWS + rr;
      uint4 cv = make_uint4(0u, 0u, 0u, 0u);
      if (g < nga && m < n) {
        if (p.c0) cv = *reinterpret_cast<const uint4*>(p.c0 + (size_t)m * p.ldc0 + u0);
        *reinterpret_cast<uint4*>(p.cseq + gran_off(0, H / 8, gu, n, m)) = cv;   // stash slab 0 = c0
      }
      unpack8(cv, cst[k]);
    }
    for (int t = 0; t < T; ++t) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int g = set + 2 * k;
        if (g >= nga) break;
        const int m = row0 + g * CL_ROWS + rr;
        const bool row_ok = m < n;
        // the input projection does not depend on the recurrence: in flight while the MMA of this item runs
        uint4 xq[4];
        xq[0] = xq[1] = xq[2] = xq[3] = make_uint4(0u, 0u, 0u, 0u);
        float xs = 0.f;
        if (row_ok) {
          if (p.x_mode == 0) {
            const bf16* xr = p.xw + ((size_t)t * n + m) * G + u0;
            xq[0] = __ldg(reinterpret_cast<const uint4*>(xr + bi * H));
            xq[1] = __ldg(reinterpret_cast<const uint4*>(xr + bfk * H));
            xq[2] = __ldg(reinterpret_cast<const uint4*>(xr + 2 * H));
            xq[3] = __ldg(reinterpret_cast<const uint4*>(xr + 3 * H));
          } else if (p.x_mode == 1) {
            // one-hot input: x W + b is row `class` of the (64, 4H) table (row 63 = bias only = zero input)
            int cls = 63;
            if (p.x_idx && t >= p.x_shift) cls = min(63, (int)__ldg(p.x_idx + (size_t)m * p.x_ld + (t - p.x_shift)));   // an index past the table = zero input, as in the dense expansion
            const bf16* xr = p.xtab + (size_t)cls * G + u0;
            xq[0] = __ldg(reinterpret_cast<const uint4*>(xr + bi * H));
            xq[1] = __ldg(reinterpret_cast<const uint4*>(xr + bfk * H));
            xq[2] = __ldg(reinterpret_cast<const uint4*>(xr + 2 * H));
            xq[3] = __ldg(reinterpret_cast<const uint4*>(xr + 3 * H));
          } else {
            xs = __bfloat162float(p.x_scalar[((size_t)t * n + m) * p.x_ld]);
          }
        }
        ptx::mbar_wait(ptx::smem_u32(&tmem_full[g]), (uint32_t)(t & 1));
        ptx::tc_fence_after();
        if (tracer && g == 0) CL_TRACE(t, 2);
        const uint32_t scr = ALIAS ? (smem_h0 + (2 * g + ((t + 1) & 1)) * HBUF) : (smem_scr0 + (uint32_t)set * CL_SCR);
        {
          float v[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(g * 64 + ch * 32), v);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int ra = ch * 32 + i;
            ptx::st_shared_f32(scr + ra * (CL_GC * 4) + (((uint32_t)c * 4) ^ ((uint32_t)(ra & 7) << 4)), v[i]);
          }
        }
        ptx::tc_fence_before();
        named_barrier(1 + 2 * set, 32 * CL_EPI_WARPS);
        if (tracer && g == 0) CL_TRACE(t, 3);
        float pre[4][8];
#pragma unroll
        for (int gt = 0; gt < 4; ++gt) {
          const uint32_t off = (uint32_t)(gt * 32 + q * 8) * 4;
          const float4 a0 = ptx::ld_shared_f32x4(scr + rr * (CL_GC * 4) + (off ^ swB));
          const float4 a1 = ptx::ld_shared_f32x4(scr + rr * (CL_GC * 4) + ((off + 16) ^ swB));
          pre[gt][0] = a0.x; pre[gt][1] = a0.y; pre[gt][2] = a0.z; pre[gt][3] = a0.w;
          pre[gt][4] = a1.x; pre[gt][5] = a1.y; pre[gt][6] = a1.z; pre[gt][7] = a1.w;
        }
        float xv[8], gi[8], gf[8], gg[8], go[8], cn[8], hn[8];
        if (p.x_mode == 2) {
          // scalar input: x w + b from the CTA's shared copy of the kernel row and bias (fp32, no intermediate rounding)
#pragma unroll
          for (int gt = 0; gt < 4; ++gt)
#pragma unroll
            for (int u = 0; u < 8; ++u) pre[gt][u] += xs * x_wb[0][gt * 32 + q * 8 + u] + x_wb[1][gt * 32 + q * 8 + u];
        }
        unpack8(xq[0], xv);
#pragma unroll
        for (int u = 0; u < 8; ++u) gi[u] = gate_fwd<HARD>(pre[0][u] + xv[u]);
        unpack8(xq[1], xv);
#pragma unroll
        for (int u = 0; u < 8; ++u) gf[u] = gate_fwd<HARD>(pre[1][u] + xv[u]);
        unpack8(xq[2], xv);
#pragma unroll
        for (int u = 0; u < 8; ++u) gg[u] = tanh_fast(pre[2][u] + xv[u]);
        unpack8(xq[3], xv);
#pragma unroll
        for (int u = 0; u < 8; ++u) go[u] = gate_fwd<HARD>(pre[3][u] + xv[u]);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float sv = gf[u] * cst[k][u] + gi[u] * gg[u];
          if (STD) { cn[u] = sv; hn[u] = go[u] * tanh_fast(sv); }
          else { cn[u] = tanh_fast(sv); hn[u] = go[u] * cn[u]; }
          cst[k][u] = cn[u];
        }
        uint4 st_h = pack8(hn);
        if (!row_ok) st_h = make_uint4(0u, 0u, 0u, 0u);
        if (t + 1 < T) {
          // exchange slot (g, t & 1) was last read by the multicast copies of step t-2, which landed before any MMA t-1 could finish
          uint8_t* dst = p.hx + ((size_t)(((cl * ng + g) * 2 + (t & 1)) * CS + j)) * CL_STAGE + (size_t)(rr >> 5) * CL_SLICE +
                         (size_t)q * (CL_HALF * 16) + (size_t)(rr & 31) * 16;
          *reinterpret_cast<uint4*>(dst) = st_h;
          // generic store -> async-proxy read by the multicast copy; the only other traffic of this thread still in flight is the
          // stash of the previous item (the input projection above has been consumed)
          ptx::fence_proxy_async_global();
          if (tracer && g == 0) CL_TRACE(t, 4);
          named_barrier(2 + 2 * set, 32 * (CL_EPI_WARPS + 1));   // also: every thread of the set is done reading the scratch tile
          if (tracer && g == 0) CL_TRACE(t, 5);
        } else {
          named_barrier(5 + set, 32 * CL_EPI_WARPS);             // last step: only the scratch tile needs protecting
        }
        if (row_ok) *reinterpret_cast<uint4*>(p.hseq + ((size_t)(t + 1) * n + m) * H + u0) = st_h;
        if (row_ok && !p.no_stash) {   // gates and c are read only by the backward pass
          *reinterpret_cast<uint4*>(p.cseq + gran_off(t + 1, H / 8, gu, n, m)) = pack8(cn);
          *reinterpret_cast<uint4*>(p.gates + gran_off(t, G / 8, bi * (H / 8) + gu, n, m)) = pack8(gi);
          *reinterpret_cast<uint4*>(p.gates + gran_off(t, G / 8, bfk * (H / 8) + gu, n, m)) = pack8(gf);
          *reinterpret_cast<uint4*>(p.gates + gran_off(t, G / 8, 2 * (H / 8) + gu, n, m)) = pack8(gg);
          *reinterpret_cast<uint4*>(p.gates + gran_off(t, G / 8, 3 * (H / 8) + gu, n, m)) = pack8(go);
        }
        if (tracer && g == 0) CL_TRACE(t, 9);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  if (warp == 1) ptx::tmem_dealloc2(tmem_base, 512);
}

// A message whose destination is the sending CTA itself: the whole warp copies it with ordinary shared-memory accesses and completes the
// receiver barrier's transaction bytes by hand.  (cp.async.bulk.shared::cluster with the CTA's own address as destination works on the
// hardware, but compute-sanitizer memcheck rejects it -- "not located in remote CTA", reproduced by scripts/sanitizer_probe bulkx .. 0 --
// and then blocks the copy, which hangs the kernel under the tool.)
__device__ __forceinline__ void self_message(uint32_t dst, uint32_t src, uint32_t bytes, uint32_t bar, int lane) {
  for (uint32_t o = (uint32_t)lane * 16u; o < bytes; o += 32u * 16u) ptx::st_shared_u4(dst + o, ptx::ld_shared_u4(src + o));
  __threadfence_block();
  __syncwarp();
  if (lane == 0) ptx::mbar_complete_tx(bar, bytes);
}

// ------------------------------------------------------------------------------------------------------------------
// Backward.  dh_{t-1} = dG_t U^T contracts over ALL 4H gate columns, so the weight-stationary split is over K:
//   * pair q of the cluster owns hidden units [64 q, 64 q + 64) = 256 gate columns.  Within the pair, CTA e does the
//     gate-gradient math for 32 of the group's 64 batch rows (x the pair's 64 units), so the dG it produces is exactly
//     ITS half (N split) of the B operand of the pair MMA: nothing is gathered;
//   * A = U^T restricted to the pair's 256 gate columns, [H output units][256] bf16, resident in TENSOR MEMORY
//     (CTA e: units 256 mt + 128 e + lane of M tile mt); D (TMEM) = partial dh^T [units of this CTA][64 rows];
//   * reduce-scatter: the partial for (64 units of pair q', 32 rows of CTA e') is one 4 KB bf16 message; every CTA sends
//     4 H/256 messages and receives H/64 (one per pair) by bulk copies into the receiver's shared memory (complete_tx
//     on its barrier), and sums them in fp32.  Messages of step t may overtake a slow receiver by at most... anything,
//     so receivers ACK what they have read (relaxed remote arrives) and senders wait for the ACKs of the previous
//     message before they overwrite their staging tile / the receivers' buffers;
//   * NG = 1..2 independent 64-row groups rotate through the same weights to hide the exchange latency.
//   iteration it (t = T-1-it), per group:  [sum partials of t+1 -> dh_t] -> gate-gradient math -> dG_t (operand tile + global)
//                                          -> pair MMA -> TMEM -> bf16 messages -> push ; it == T: dh_{-1}, dc_{-1} -> dS
constexpr int CLB_MAXG = 2;
constexpr uint32_t CLB_GS = CL_HALF * 16 + 16;     // 528 B: k-granule stride of the dG operand tile / unit-granule stride of a message (bank padding)
constexpr uint32_t CLB_MSG = 8 * CLB_GS;           // 4224 B: 8 unit granules x 32 rows x 16 B
constexpr uint32_t CLB_BT = 32 * CLB_GS;           // 16896 B: 32 k-granules (4 gates x 64 units / 8) x 32 rows x 16 B

struct ClusterBP {
  int n, steps, nswap, ng;
  const bf16* gates; const bf16* cseq; const bf16* dhext; const bf16* dh_last; int ld_last;
  bf16* dG; bf16* dS_h; bf16* dS_c; int ldS;
  const bf16* upack;      // [CS][MT][128 units][256 gate columns of the pair]
  int l2_prefetch;        // > 0: prefetch the stash of step t - l2_prefetch into L2
  long long* trace;
};

template <bool HARD>
__device__ __forceinline__ float gate_bwd(float s) {
  return HARD ? ((s > 0.f && s < 1.f) ? 0.2f : 0.f) : s * (1.f - s);
}
__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
__device__ __forceinline__ void st_shared_u16(uint32_t addr, unsigned short v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory"); }

// Warp roles (19 warps): 0 = MMA issuer (even CTA), 1 / 2 = push warp of group 0 / 1 (warp 1 also owns the TMEM allocation),
// 3..10 = epilogue warps of group 0, 11..18 = epilogue warps of group 1.  The two groups' epilogues are instruction-bound
// (~900 instructions per thread and step), so each group gets its own 8 warps: 4 warps per scheduler instead of 2.
constexpr int CLW_SETS = 2;
constexpr int CLW_THREADS = 32 * (3 + CLW_SETS * CL_EPI_WARPS);

template <int CS, bool HARD, bool STD>
__global__ void __launch_bounds__(CLW_THREADS, 1)
rec_cluster_bwd_kernel(const ClusterBP p) {
  constexpr int H = CS * CL_HS, G = 4 * H, NP = CS / 2, MT = H / 256, ND = 4 * MT;
  constexpr uint32_t TM_U = CLB_MAXG * MT * 64;                     // TMEM: D(g, mt) at (g MT + mt) 64, U^T tile mt at TM_U + 128 mt
  constexpr uint32_t GRP = CLB_BT + (ND + NP) * CLB_MSG;            // per group: operand tile | staging (ND messages) | receive (NP messages)
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t b_ready[CLB_MAXG], tmem_full[CLB_MAXG], recv_full[CLB_MAXG], ack[CLB_MAXG];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int cl = (int)blockIdx.x / CS;
  const int e = (int)(rank & 1), q = (int)(rank >> 1);
  const int rhs = e ^ p.nswap;                                      // which 32 rows of a group this CTA does the cell math for
  const int T = p.steps, n = p.n, ng = p.ng;
  const int row0 = cl * CL_ROWS * ng;
  const int nga = min(ng, (n - row0 + CL_ROWS - 1) / CL_ROWS);

  if (threadIdx.x == 0) {
    for (int g = 0; g < CLB_MAXG; ++g) {
      ptx::mbar_init(ptx::smem_u32(&b_ready[g]), 2 * CL_EPI_WARPS);
      ptx::mbar_init(ptx::smem_u32(&tmem_full[g]), 1);
      ptx::mbar_init(ptx::smem_u32(&recv_full[g]), 1);
      ptx::mbar_init(ptx::smem_u32(&ack[g]), ND * CL_EPI_WARPS);
    }
    ptx::fence_barrier_init();
    for (int g = 0; g < CLB_MAXG; ++g) ptx::mbar_arrive_expect_tx(ptx::smem_u32(&recv_full[g]), NP * CLB_MSG);   // messages of iteration 0
  }
  if (warp == 1) {
    ptx::tmem_alloc2(ptx::smem_u32(&tmem_base_slot), 512);
    ptx::tmem_relinquish2();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp >= 3 && warp < 7) {
    // U^T tiles -> tensor memory: lane = output unit, 32-bit column c = gate-column pair (2c, 2c+1) of the pair's 256
    const int mrow = (warp & 3) * 32 + lane;
    for (int mt = 0; mt < MT; ++mt) {
      const uint4* src = reinterpret_cast<const uint4*>(p.upack + (((size_t)rank * MT + mt) * 128 + mrow) * 256);
      for (int c32 = 0; c32 < 4; ++c32) {
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 v = __ldg(src + c32 * 8 + i);
          r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
        }
        ptx::tmem_st_32x32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + TM_U + (uint32_t)(mt * 128 + c32 * 32), r);
      }
    }
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  ptx::tc_fence_after();

  if (warp == 0) {
    if (e == 0) {
      // ===================== MMA issuer (even CTA of the pair): warp-uniform loop, one elected lane issues =====================
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(256, CL_ROWS, false, false);
      const uint16_t pair_mask = (uint16_t)(3u << rank);
      const uint64_t bd0 = ptx::umma_desc_noswz(smem_base, CLB_GS, 128);
      for (int it = 0; it < T; ++it) {
#pragma unroll
        for (int g = 0; g < CLB_MAXG; ++g) {
          if (g >= nga) break;
          ptx::mbar_wait(ptx::smem_u32(&b_ready[g]), (uint32_t)(it & 1));   // both CTAs wrote (and fenced) their dG_t tiles
          ptx::tc_fence_after();
          if (g == 0 && lane == 0) CL_TRACE(it, 0);
          const uint64_t bd = bd0 + (uint64_t)((g * GRP) >> 4);
          if (ptx::elect_one()) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
              for (int ks = 0; ks < 16; ++ks)
                ptx::umma_bf16_2cta_ts(tmem_base + (uint32_t)((g * MT + mt) * 64), tmem_base + TM_U + (uint32_t)(mt * 128 + ks * 8),
                                       bd + (uint64_t)((ks * 2 * CLB_GS) >> 4), idesc, ks > 0 ? 1u : 0u);
            ptx::umma_commit_2cta(ptx::smem_u32(&tmem_full[g]), pair_mask);
          }
          __syncwarp();
          if (g == 0 && lane == 0) CL_TRACE(it, 1);
        }
      }
    }
  } else if (warp <= CLW_SETS) {
    // ===================== push warp of group g: lane x sends message x of the freshly staged partials =====================
    const int g = warp - 1;
    if (g < nga) {
      const uint32_t gb = smem_base + (uint32_t)g * GRP;
      const int xl = lane % ND;
      const int mt = xl >> 2, hp = (xl >> 1) & 1, ep = xl & 1;
      const uint32_t dest = (uint32_t)(2 * (4 * mt + 2 * e + hp) + (ep ^ p.nswap));
      const uint32_t dst = ptx::mapa(gb + CLB_BT + (ND + q) * CLB_MSG, dest), dbar = ptx::mapa(ptx::smem_u32(&recv_full[g]), dest);
      const uint32_t src = gb + CLB_BT + (uint32_t)xl * CLB_MSG;
      // which of my ND messages (if any) is addressed to this CTA itself: warp-uniform
      int self_x = -1;
      for (int x = 0; x < ND; ++x)
        if ((uint32_t)(2 * (4 * (x >> 2) + 2 * e + ((x >> 1) & 1)) + ((x & 1) ^ p.nswap)) == rank) self_x = x;
      for (int it = 0; it < T; ++it) {
        named_barrier(1 + g, 32 * (CL_EPI_WARPS + 1));
        ptx::fence_proxy_async();            // staging was written with generic stores by the epilogue warps (ordered by the barrier)
        if (lane < ND && dest != rank) ptx::bulk_copy_dsmem(dst, src, CLB_MSG, dbar);
        if (self_x >= 0) self_message(gb + CLB_BT + (ND + q) * CLB_MSG, gb + CLB_BT + (uint32_t)self_x * CLB_MSG, CLB_MSG, ptx::smem_u32(&recv_full[g]), lane);
        if (lane == 0 && g == 0) CL_TRACE(it, 6);
      }
    }
  } else {
    // ===================== epilogue warps: set s = group s =====================
    const int g = (warp - 3) >> 3, ew = (warp - 3) & 7;
    if (g < nga) {
      const uint32_t gb = smem_base + (uint32_t)g * GRP;
      // cell ownership: row r of this CTA's 32, unit granule gq of the pair's 8 (4 rows x 8 granules per warp: every global access >= 64 B contiguous)
      const int r = ew * 4 + (lane >> 3), gq = lane & 7;
      const int u0 = 64 * q + 8 * gq, gu = u0 >> 3;
      const int m = row0 + g * CL_ROWS + rhs * CL_HALF + r;
      const bool row_ok = m < n;
      // drain ownership: TMEM lane = unit 128 e + wq 32 + lane of an M tile, 32 columns = the rows of CTA `ch` of the destination pair
      const int wq = warp & 3, ch = ew >> 2;
      const int e_src = (q & 3) >> 1;                               // which CTA of every pair holds the partials for my pair's units
      constexpr int bi = STD ? 0 : 1, bfk = 1 - bi;
      const bool tracer = (ew == 0 && lane == 0);
      float dc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) dc[u] = 0.f;

      for (int it = 0; it <= T; ++it) {
        const int t = T - 1 - it;
        // stash of step t: independent of the recurrence, in flight while the partials arrive
        uint4 sg[4], sc0, sc1, sex;
        sex = make_uint4(0u, 0u, 0u, 0u);
        if (t >= 0 && row_ok) {
          sg[0] = __ldg(reinterpret_cast<const uint4*>(p.gates + gran_off(t, G / 8, bi * (H / 8) + gu, n, m)));
          sg[1] = __ldg(reinterpret_cast<const uint4*>(p.gates + gran_off(t, G / 8, bfk * (H / 8) + gu, n, m)));
          sg[2] = __ldg(reinterpret_cast<const uint4*>(p.gates + gran_off(t, G / 8, 2 * (H / 8) + gu, n, m)));
          sg[3] = __ldg(reinterpret_cast<const uint4*>(p.gates + gran_off(t, G / 8, 3 * (H / 8) + gu, n, m)));
          sc0 = __ldg(reinterpret_cast<const uint4*>(p.cseq + gran_off(t, H / 8, gu, n, m)));
          sc1 = __ldg(reinterpret_cast<const uint4*>(p.cseq + gran_off(t + 1, H / 8, gu, n, m)));
          if (p.dhext) sex = __ldg(reinterpret_cast<const uint4*>(p.dhext + ((size_t)t * n + m) * H + u0));
        }
        if (p.l2_prefetch && t >= p.l2_prefetch && row_ok) {
          // the stash streams from HBM next to the weight-gradient GEMMs of the side stream: pull the lines of a later step into L2 now
          const int tp = t - p.l2_prefetch;
          prefetch_l2(p.gates + gran_off(tp, G / 8, bi * (H / 8) + gu, n, m));
          prefetch_l2(p.gates + gran_off(tp, G / 8, bfk * (H / 8) + gu, n, m));
          prefetch_l2(p.gates + gran_off(tp, G / 8, 2 * (H / 8) + gu, n, m));
          prefetch_l2(p.gates + gran_off(tp, G / 8, 3 * (H / 8) + gu, n, m));
          prefetch_l2(p.cseq + gran_off(tp, H / 8, gu, n, m));
          if (p.dhext) prefetch_l2(p.dhext + ((size_t)tp * n + m) * H + u0);
        }
        // ---- dh_t = sum over the pairs of their partial dG_{t+1} U^T for this thread's 8 units
        float dh[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) dh[u] = 0.f;
        if (it == 0 && p.dh_last && row_ok) unpack8(__ldg(reinterpret_cast<const uint4*>(p.dh_last + (size_t)m * p.ld_last + u0)), dh);
        if (it > 0) {
          ptx::mbar_wait(ptx::smem_u32(&recv_full[g]), (uint32_t)((it - 1) & 1));
          if (tracer && g == 0) CL_TRACE(it, 2);
          if (tracer && it < T) ptx::mbar_arrive_expect_tx(ptx::smem_u32(&recv_full[g]), NP * CLB_MSG);   // arm the phase of iteration `it`
#pragma unroll
          for (int s = 0; s < NP; ++s) {
            float f[8];
            unpack8(ptx::ld_shared_u4(gb + CLB_BT + (ND + s) * CLB_MSG + (uint32_t)gq * CLB_GS + (uint32_t)r * 16), f);
#pragma unroll
            for (int u = 0; u < 8; ++u) dh[u] += f[u];
          }
          if (it < T) {
            // tell every source that its message has been read: it may overwrite its staging tile and my receive slot
            __syncwarp();
            if (lane < NP) ptx::mbar_arrive_remote_relaxed(ptx::mapa(ptx::smem_u32(&ack[g]), (uint32_t)(2 * lane + e_src)));
          }
        }
        if (t < 0) {
          if (row_ok && p.dS_h) {
            *reinterpret_cast<uint4*>(p.dS_h + (size_t)m * p.ldS + u0) = pack8(dh);
            *reinterpret_cast<uint4*>(p.dS_c + (size_t)m * p.ldS + u0) = pack8(dc);
          }
          break;
        }
        // ---- gate-gradient math for step t
        uint4 pk[4];
        if (row_ok) {
          float gi[8], gf[8], gg[8], go[8], c0[8], c1[8], ex[8], dv[8];
          unpack8(sg[0], gi); unpack8(sg[1], gf); unpack8(sg[2], gg); unpack8(sg[3], go);
          unpack8(sc0, c0); unpack8(sc1, c1); unpack8(sex, ex);
          float ds[8], d_o[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float d = dh[u] + ex[u];
            if (STD) {
              const float tc = tanh_fast(c1[u]);
              d_o[u] = d * tc;
              ds[u] = dc[u] + d * go[u] * (1.f - tc * tc);
            } else {
              d_o[u] = d * c1[u];
              ds[u] = (dc[u] + d * go[u]) * (1.f - c1[u] * c1[u]);
            }
            dc[u] = ds[u] * gf[u];
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) dv[u] = ds[u] * gg[u] * gate_bwd<HARD>(gi[u]);
          pk[0] = pack8(dv);
#pragma unroll
          for (int u = 0; u < 8; ++u) dv[u] = ds[u] * c0[u] * gate_bwd<HARD>(gf[u]);
          pk[1] = pack8(dv);
#pragma unroll
          for (int u = 0; u < 8; ++u) dv[u] = ds[u] * gi[u] * (1.f - gg[u] * gg[u]);
          pk[2] = pack8(dv);
#pragma unroll
          for (int u = 0; u < 8; ++u) dv[u] = d_o[u] * gate_bwd<HARD>(go[u]);
          pk[3] = pack8(dv);
        } else {
          pk[0] = pk[1] = pk[2] = pk[3] = make_uint4(0u, 0u, 0u, 0u);
        }
        // operand tile: k = gate * 64 + unit-in-pair -> k-granule gate * 8 + gq, this CTA's row r
#pragma unroll
        for (int gt = 0; gt < 4; ++gt) ptx::st_shared_u4(gb + (uint32_t)(gt * 8 + gq) * CLB_GS + (uint32_t)r * 16, pk[gt]);
        // nothing of this thread is in flight to global memory here except the dG stores of the previous step
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (e == 0) ptx::mbar_arrive(ptx::smem_u32(&b_ready[g]));
          else ptx::mbar_arrive_remote_relaxed(ptx::mapa(ptx::smem_u32(&b_ready[g]), rank - 1));
        }
        if (tracer && g == 0) CL_TRACE(it, 3);
        // ---- off the critical path: dG_t for the batched weight-gradient GEMMs
        if (row_ok) {
          bf16* dgp = p.dG + ((size_t)t * n + m) * G + u0;
          *reinterpret_cast<uint4*>(dgp + bi * H) = pk[0];
          *reinterpret_cast<uint4*>(dgp + bfk * H) = pk[1];
          *reinterpret_cast<uint4*>(dgp + 2 * H) = pk[2];
          *reinterpret_cast<uint4*>(dgp + 3 * H) = pk[3];
        }
        // ---- partial dh_{t-1} of this CTA's units: TMEM -> bf16 messages in the staging tile
        ptx::mbar_wait(ptx::smem_u32(&tmem_full[g]), (uint32_t)(it & 1));
        ptx::tc_fence_after();
        if (tracer && g == 0) CL_TRACE(it, 4);
        if (it > 0) ptx::mbar_wait(ptx::smem_u32(&ack[g]), (uint32_t)((it - 1) & 1));   // every receiver has read message it-1
        if (tracer && g == 0) CL_TRACE(it, 5);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          float v[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)((g * MT + mt) * 64 + ch * 32), v);
          const uint32_t msg = (uint32_t)(((mt * 2 + (wq >> 1)) * 2) + ch);
          const uint32_t base = gb + CLB_BT + msg * CLB_MSG + (uint32_t)((wq & 1) * 4 + (lane >> 3)) * CLB_GS + (uint32_t)(lane & 7) * 2;
#pragma unroll
          for (int i = 0; i < 32; ++i) st_shared_u16(base + i * 16, __bfloat16_as_ushort(__float2bfloat16_rn(v[i])));
        }
        ptx::tc_fence_before();
        named_barrier(1 + g, 32 * (CL_EPI_WARPS + 1));
        if (tracer && g == 0) CL_TRACE(it, 9);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  if (warp == 1) ptx::tmem_dealloc2(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// Backward, second form (H = 512, 16-CTA clusters): K split over QUADS instead of pairs, so that half of the exchange becomes an
// all-gather that can ride the multicast path and the DSMEM reduce-scatter shrinks from 8 x 4 KB to 4 x 4 KB per CTA and step.
//   * cell ownership as in the forward kernel: CTA `rank` does the gate-gradient math of units [32 rank, 32 rank + 32) for all 64 rows
//     of a group ((row, 8 units) per thread);
//   * quad kq = CTAs 4 kq .. 4 kq + 3 owns units [128 kq, +128) = 512 gate columns = one K quarter.  Its four dG pieces (64 rows x 128
//     gate columns each) are ALL-GATHERED inside the quad: every CTA stores its piece to a global slot and multicasts each row half
//     (8 KB) into the B-operand tiles of the two quad CTAs that hold that half; a tile is 32 rows x 512 k, double-buffered, and a
//     buffer is re-used only after both pair MMAs that read it have committed (tcgen05.commit multicast to the quad's b_free barriers);
//   * pair (mt, kq) = CTAs 4 kq + 2 mt + {0, 1}: A = U^T [256 output units of M tile mt][512 k of quarter kq] in tensor memory,
//     D = partial dh^T [128 units per CTA][64 rows], ONE M tile, 32 MMAs (M256 N64 K16) per step as before;
//   * reduce-scatter: CTA (kq, c) sends the 32-unit blocks of its D to the four CTAs of quad #c and receives one 4 KB bf16 message per
//     K quarter (from CTAs 4 s + kq): bulk copies + ACKs exactly as in the first form, with half the bytes and half the drain work.
constexpr uint32_t CLQ_BT = 64 * CL_HALF * 16;         // 32 KB: B tile [64 k-granules][32 rows][16 B]
constexpr uint32_t CLQ_PIECE = 16 * CL_HALF * 16;      // 8 KB: one CTA's dG piece for one row half (16 k-granules)
constexpr uint32_t CLQ_MGS = CL_ROWS * 16 + 16;        // 1040 B: unit-granule stride of a message (64 rows x 16 B + bank padding)
constexpr uint32_t CLQ_MSG = 4 * CLQ_MGS;              // 4160 B: 32 units x 64 rows bf16
constexpr uint32_t CLQ_GRP = 2 * CLQ_BT + 8 * CLQ_MSG; // per group: 2 B tiles | staging (4 messages) | receive (4 messages)

struct ClusterQP {
  int n, steps, nswap, ng, l2_prefetch;
  const bf16* gates; const bf16* cseq; const bf16* dhext; const bf16* dh_last; int ld_last;
  bf16* dG; bf16* dS_h; bf16* dS_c; int ldS;
  const bf16* upack;      // [16][128 output units][512 k]
  uint8_t* xbuf;          // dG exchange slots [clusters][ng][2][16 CTAs][2 row halves][8 KB]
  long long* trace;
  int stm;                // partial-dh messages staged with fragment reads + stmatrix.trans instead of per-element stores
};

template <bool HARD, bool STD>
__global__ void __launch_bounds__(CLW_THREADS, 1)
rec_cluster_bwd4_kernel(const ClusterQP p) {
  constexpr int CS = 16, H = 512, G = 4 * H;
  constexpr uint32_t TM_U = CLB_MAXG * 64;                            // TMEM: D(g) at 64 g, U^T tile from column 128
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t b_full[CLB_MAXG][2], b_free[CLB_MAXG][2], peer_ready[CLB_MAXG][2], tmem_full[CLB_MAXG], recv_full[CLB_MAXG], ack[CLB_MAXG];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int cl = (int)blockIdx.x / CS;
  const int e = (int)(rank & 1), kq = (int)(rank >> 2), c = (int)(rank & 3);
  const int T = p.steps, n = p.n, ng = p.ng;
  const int row0 = cl * CL_ROWS * ng;
  const int nga = min(ng, (n - row0 + CL_ROWS - 1) / CL_ROWS);

  if (threadIdx.x == 0) {
    for (int g = 0; g < CLB_MAXG; ++g) {
      for (int b = 0; b < 2; ++b) {
        ptx::mbar_init(ptx::smem_u32(&b_full[g][b]), 1);
        ptx::mbar_init(ptx::smem_u32(&b_free[g][b]), 2);
        ptx::mbar_init(ptx::smem_u32(&peer_ready[g][b]), 1);
      }
      ptx::mbar_init(ptx::smem_u32(&tmem_full[g]), 1);
      ptx::mbar_init(ptx::smem_u32(&recv_full[g]), 1);
      ptx::mbar_init(ptx::smem_u32(&ack[g]), 4 * CL_EPI_WARPS);
    }
    ptx::fence_barrier_init();
    for (int g = 0; g < CLB_MAXG; ++g) ptx::mbar_arrive_expect_tx(ptx::smem_u32(&recv_full[g]), 4 * CLQ_MSG);   // messages of iteration 0
  }
  if (warp == 1) {
    ptx::tmem_alloc2(ptx::smem_u32(&tmem_base_slot), 512);
    ptx::tmem_relinquish2();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp >= 3 && warp < 7) {
    // U^T tile -> tensor memory: lane = output unit 256 mt + 128 e + lane, 32-bit column = k pair of the quad's 512 k
    const int mrow = (warp & 3) * 32 + lane;
    const uint4* src = reinterpret_cast<const uint4*>(p.upack + ((size_t)rank * 128 + mrow) * 512);
    for (int c32 = 0; c32 < 8; ++c32) {
      uint32_t r[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 v = __ldg(src + c32 * 8 + i);
        r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
      }
      ptx::tmem_st_32x32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + TM_U + (uint32_t)c32 * 32u, r);
    }
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  ptx::tc_fence_after();

  if (warp == 0) {
    if (e == 0) {
      // ===================== MMA issuer (even CTA of the pair) =====================
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(256, CL_ROWS, false, false);
      const uint16_t pair_mask = (uint16_t)(3u << rank);
      const uint16_t quad_mask = (uint16_t)(0xFu << (4 * kq));
      const uint64_t bd0 = ptx::umma_desc_noswz(smem_base, CL_HALF * 16, 128);
      for (int it = 0; it < T; ++it) {
        const int b = it & 1;
#pragma unroll
        for (int g = 0; g < CLB_MAXG; ++g) {
          if (g >= nga) break;
          if (lane == 0) ptx::mbar_arrive_expect_tx(ptx::smem_u32(&b_full[g][b]), CLQ_BT);
          ptx::mbar_wait(ptx::smem_u32(&b_full[g][b]), (uint32_t)((it >> 1) & 1));          // my half of the quad's dG_t
          ptx::mbar_wait(ptx::smem_u32(&peer_ready[g][b]), (uint32_t)((it >> 1) & 1));      // the odd CTA's half
          ptx::tc_fence_after();
          if (g == 0 && lane == 0) CL_TRACE(it, 0);
          const uint64_t bd = bd0 + (uint64_t)((g * CLQ_GRP + b * CLQ_BT) >> 4);
          if (ptx::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 32; ++ks)
              ptx::umma_bf16_2cta_ts(tmem_base + (uint32_t)g * 64u, tmem_base + TM_U + (uint32_t)ks * 8u,
                                     bd + (uint64_t)((ks * 2 * (CL_HALF * 16)) >> 4), idesc, ks > 0 ? 1u : 0u);
            ptx::umma_commit_2cta(ptx::smem_u32(&tmem_full[g]), pair_mask);
            ptx::umma_commit_2cta(ptx::smem_u32(&b_free[g][b]), quad_mask);   // both B tiles of this pair may be overwritten (by anyone in the quad)
          }
          __syncwarp();
          if (g == 0 && lane == 0) CL_TRACE(it, 1);
        }
      }
    } else if (lane == 0) {
      // ===================== relay (odd CTA): "my half of dG_t landed" -> the pair's MMA issuer =====================
      const uint32_t leader = rank - 1;
      for (int it = 0; it < T; ++it) {
        const int b = it & 1;
        for (int g = 0; g < nga; ++g) {
          const uint32_t bb = ptx::smem_u32(&b_full[g][b]);
          ptx::mbar_arrive_expect_tx(bb, CLQ_BT);
          ptx::mbar_wait(bb, (uint32_t)((it >> 1) & 1));
          ptx::mbar_arrive_remote_relaxed(ptx::mapa(ptx::smem_u32(&peer_ready[g][b]), leader));
        }
      }
    }
  } else if (warp <= CLW_SETS) {
    // ===================== helper warp of group g: multicasts this CTA's dG piece, later pushes its partial messages =====================
    const int g = warp - 1;
    if (g < nga) {
      const uint32_t gb = smem_base + (uint32_t)g * CLQ_GRP;
      // push: message w goes to CTA 4 c + w (quad #c), slot kq of its receive buffer
      const int xl = lane & 3;
      const uint32_t dest = (uint32_t)(4 * c + xl);
      const uint32_t pdst = ptx::mapa(gb + 2 * CLQ_BT + (4 + kq) * CLQ_MSG, dest), pbar = ptx::mapa(ptx::smem_u32(&recv_full[g]), dest);
      const uint32_t psrc = gb + 2 * CLQ_BT + (uint32_t)xl * CLQ_MSG;
      for (int it = 0; it < T; ++it) {
        const int b = it & 1;
        named_barrier(1 + g, 32 * (CL_EPI_WARPS + 1));                      // dG piece of step t stored to the exchange slot (and fenced)
        if (it >= 2) ptx::mbar_wait(ptx::smem_u32(&b_free[g][b]), (uint32_t)((((it >> 1) - 1)) & 1));   // every MMA that read buffer b is done
        if (lane < 2) {
          const int ed = (lane ^ p.nswap) & 1;                              // parity of the quad CTAs that hold row half `lane`
          const uint16_t mask = (uint16_t)(((1u << ed) | (1u << (2 + ed))) << (4 * kq));
          const uint8_t* src = p.xbuf + ((size_t)(((cl * ng + g) * 2 + b) * CS + rank)) * (2 * CLQ_PIECE) + (size_t)lane * CLQ_PIECE;
          ptx::bulk_load_multicast(gb + b * CLQ_BT + (uint32_t)c * CLQ_PIECE, src, CLQ_PIECE, ptx::smem_u32(&b_full[g][b]), mask);
        }
        if (lane == 0 && g == 0) CL_TRACE(it, 8);
        named_barrier(3 + g, 32 * (CL_EPI_WARPS + 1));                      // partial messages staged
        ptx::fence_proxy_async();
        if (lane < 4 && dest != rank) ptx::bulk_copy_dsmem(pdst, psrc, CLQ_MSG, pbar);
        if (c == kq) self_message(gb + 2 * CLQ_BT + (4 + kq) * CLQ_MSG, gb + 2 * CLQ_BT + (uint32_t)c * CLQ_MSG, CLQ_MSG, ptx::smem_u32(&recv_full[g]), lane);   // message c of CTA (kq = c, c) stays here
        if (lane == 0 && g == 0) CL_TRACE(it, 6);
      }
    }
  } else {
    // ===================== epilogue warps: set s = group s =====================
    const int g = (warp - 3) >> 3, ew = (warp - 3) & 7;
    if (g < nga) {
      const uint32_t gb = smem_base + (uint32_t)g * CLQ_GRP;
      // cell ownership (as in the forward kernel): batch row rr, units [8 q, 8 q + 8) of this CTA's 32
      const int rr = ew * 8 + (lane >> 2), q = lane & 3;
      const int u0 = (int)rank * CL_HS + q * 8, gu = u0 >> 3;
      const int m = row0 + g * CL_ROWS + rr;
      const bool row_ok = m < n;
      // drain ownership: TMEM lane = output unit 128 (rank & 3) + 32 wq + lane -> message wq, 32 columns = batch rows ch * 32 ..
      const int wq = warp & 3, ch = ew >> 2;
      constexpr int bi = STD ? 0 : 1, bfk = 1 - bi;
      const bool tracer = (ew == 0 && lane == 0);
      float dc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) dc[u] = 0.f;

      // BPTT stash of a step (4 gate granules, c_t, dh_ext): loaded ONE STEP AHEAD -- issued right after the exchange barrier of the previous
      // iteration, in flight during its TMEM drain and the wait for the partial messages -- so that the gate-gradient math never waits for L2
      // (round 1 loaded at the top of the iteration: long_scoreboard was 45 % of the epilogue warps' stalls).  c_{t+1} of step t is c_t of t+1.
      uint4 sg[4], sc0, sc1, sex;
      sex = make_uint4(0u, 0u, 0u, 0u);
      sc0 = sc1 = sg[0] = sg[1] = sg[2] = sg[3] = make_uint4(0u, 0u, 0u, 0u);
      auto load_stash = [&](int ts) {
        sg[0] = __ldg(reinterpret_cast<const uint4*>(p.gates + gran_off(ts, G / 8, bi * (H / 8) + gu, n, m)));
        sg[1] = __ldg(reinterpret_cast<const uint4*>(p.gates + gran_off(ts, G / 8, bfk * (H / 8) + gu, n, m)));
        sg[2] = __ldg(reinterpret_cast<const uint4*>(p.gates + gran_off(ts, G / 8, 2 * (H / 8) + gu, n, m)));
        sg[3] = __ldg(reinterpret_cast<const uint4*>(p.gates + gran_off(ts, G / 8, 3 * (H / 8) + gu, n, m)));
        sc0 = __ldg(reinterpret_cast<const uint4*>(p.cseq + gran_off(ts, H / 8, gu, n, m)));
        if (p.dhext) sex = __ldg(reinterpret_cast<const uint4*>(p.dhext + ((size_t)ts * n + m) * H + u0));
      };
      if (row_ok && T > 0) {
        sc0 = __ldg(reinterpret_cast<const uint4*>(p.cseq + gran_off(T, H / 8, gu, n, m)));   // becomes c_{t+1} of the first step below
        sc1 = sc0;
        load_stash(T - 1);
      }

      for (int it = 0; it <= T; ++it) {
        const int t = T - 1 - it;
        if (p.l2_prefetch && t >= p.l2_prefetch && row_ok) {
          const int tp = t - p.l2_prefetch;
          prefetch_l2(p.gates + gran_off(tp, G / 8, bi * (H / 8) + gu, n, m));
          prefetch_l2(p.gates + gran_off(tp, G / 8, bfk * (H / 8) + gu, n, m));
          prefetch_l2(p.gates + gran_off(tp, G / 8, 2 * (H / 8) + gu, n, m));
          prefetch_l2(p.gates + gran_off(tp, G / 8, 3 * (H / 8) + gu, n, m));
          prefetch_l2(p.cseq + gran_off(tp, H / 8, gu, n, m));
          if (p.dhext) prefetch_l2(p.dhext + ((size_t)tp * n + m) * H + u0);
        }
        // ---- dh_t = sum over the K quarters of their partial dG_{t+1} U^T for this thread's 8 units
        float dh[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) dh[u] = 0.f;
        if (it == 0 && p.dh_last && row_ok) unpack8(__ldg(reinterpret_cast<const uint4*>(p.dh_last + (size_t)m * p.ld_last + u0)), dh);
        if (it > 0) {
          ptx::mbar_wait(ptx::smem_u32(&recv_full[g]), (uint32_t)((it - 1) & 1));
          if (tracer && g == 0) CL_TRACE(it, 2);
          if (tracer && it < T) ptx::mbar_arrive_expect_tx(ptx::smem_u32(&recv_full[g]), 4 * CLQ_MSG);
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            float f[8];
            unpack8(ptx::ld_shared_u4(gb + 2 * CLQ_BT + (4 + s) * CLQ_MSG + (uint32_t)q * CLQ_MGS + (uint32_t)rr * 16), f);
#pragma unroll
            for (int u = 0; u < 8; ++u) dh[u] += f[u];
          }
          if (it < T) {
            __syncwarp();
            if (lane < 4) ptx::mbar_arrive_remote_relaxed(ptx::mapa(ptx::smem_u32(&ack[g]), (uint32_t)(4 * lane + kq)));   // sources: CTAs 4 s + kq
          }
        }
        if (t < 0) {
          if (row_ok && p.dS_h) {
            *reinterpret_cast<uint4*>(p.dS_h + (size_t)m * p.ldS + u0) = pack8(dh);
            *reinterpret_cast<uint4*>(p.dS_c + (size_t)m * p.ldS + u0) = pack8(dc);
          }
          break;
        }
        // ---- gate-gradient math for step t
        uint4 pk[4];
        if (row_ok) {
          float gi[8], gf[8], gg[8], go[8], c0[8], c1[8], ex[8], dv[8];
          unpack8(sg[0], gi); unpack8(sg[1], gf); unpack8(sg[2], gg); unpack8(sg[3], go);
          unpack8(sc0, c0); unpack8(sc1, c1); unpack8(sex, ex);
          float ds[8], d_o[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float d = dh[u] + ex[u];
            if (STD) {
              const float tc = tanh_fast(c1[u]);
              d_o[u] = d * tc;
              ds[u] = dc[u] + d * go[u] * (1.f - tc * tc);
            } else {
              d_o[u] = d * c1[u];
              ds[u] = (dc[u] + d * go[u]) * (1.f - c1[u] * c1[u]);
            }
            dc[u] = ds[u] * gf[u];
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) dv[u] = ds[u] * gg[u] * gate_bwd<HARD>(gi[u]);
          pk[0] = pack8(dv);
#pragma unroll
          for (int u = 0; u < 8; ++u) dv[u] = ds[u] * c0[u] * gate_bwd<HARD>(gf[u]);
          pk[1] = pack8(dv);
#pragma unroll
          for (int u = 0; u < 8; ++u) dv[u] = ds[u] * gi[u] * (1.f - gg[u] * gg[u]);
          pk[2] = pack8(dv);
#pragma unroll
          for (int u = 0; u < 8; ++u) dv[u] = d_o[u] * gate_bwd<HARD>(go[u]);
          pk[3] = pack8(dv);
        } else {
          pk[0] = pk[1] = pk[2] = pk[3] = make_uint4(0u, 0u, 0u, 0u);
        }
        // this CTA's dG piece -> exchange slot (k within the piece = gate * 32 + unit: k-granule gate * 4 + q), one 8 KB block per row half
        {
          uint8_t* slot = p.xbuf + ((size_t)(((cl * ng + g) * 2 + (it & 1)) * CS + rank)) * (2 * CLQ_PIECE) + (size_t)(rr >> 5) * CLQ_PIECE + (size_t)(rr & 31) * 16;
#pragma unroll
          for (int gt = 0; gt < 4; ++gt) *reinterpret_cast<uint4*>(slot + (size_t)(gt * 4 + q) * (CL_HALF * 16)) = pk[gt];
        }
        ptx::fence_proxy_async_global();       // generic stores -> the multicast copy (async proxy); nothing else of this thread is in flight but old dG rows
        named_barrier(1 + g, 32 * (CL_EPI_WARPS + 1));
        if (tracer && g == 0) CL_TRACE(it, 3);
        // next step's stash (after the fence above, so that the fence never waits for these loads)
        if (row_ok && t > 0) { sc1 = sc0; load_stash(t - 1); }
        // ---- off the critical path: dG_t for the batched weight-gradient GEMMs
        if (row_ok) {
          bf16* dgp = p.dG + ((size_t)t * n + m) * G + u0;
          *reinterpret_cast<uint4*>(dgp + bi * H) = pk[0];
          *reinterpret_cast<uint4*>(dgp + bfk * H) = pk[1];
          *reinterpret_cast<uint4*>(dgp + 2 * H) = pk[2];
          *reinterpret_cast<uint4*>(dgp + 3 * H) = pk[3];
        }
        // ---- partial dh_{t-1} of this CTA's 128 output units: TMEM -> bf16 messages in the staging tile
        ptx::mbar_wait(ptx::smem_u32(&tmem_full[g]), (uint32_t)(it & 1));
        ptx::tc_fence_after();
        if (tracer && g == 0) CL_TRACE(it, 4);
        if (it > 0) ptx::mbar_wait(ptx::smem_u32(&ack[g]), (uint32_t)((it - 1) & 1));   // every receiver has read message it-1
        if (tracer && g == 0) CL_TRACE(it, 5);
        if (p.stm) {
          // accumulator-fragment reads + transposing matrix stores: 2 x tcgen05.ld (16 lanes x 32 columns), 16 packed conversions, 4 x stmatrix.x4
          // (each 8 units x 8 rows block lands as eight conflict-free 16-byte rows) instead of 32 two-byte stores per thread
          const uint32_t mbase = gb + 2 * CLQ_BT + (uint32_t)wq * CLQ_MSG + (uint32_t)(ch * 32 + (lane & 7)) * 16;
          const int mi = lane >> 3;                                   // which of the 4 matrices of an instruction this lane addresses
#pragma unroll
          for (int hl = 0; hl < 2; ++hl) {
            float v[16];
            ptx::tmem_ld_16x256b_x4(tmem_base + ((uint32_t)(wq * 32 + hl * 16) << 16) + (uint32_t)(g * 64 + ch * 32), v);
#pragma unroll
            for (int cbh = 0; cbh < 2; ++cbh) {
              // matrix mi: unit granule 2 hl + (mi & 1), rows 32 ch + 8 (2 cbh + (mi >> 1)) + j
              const uint32_t addr = mbase + (uint32_t)(2 * hl + (mi & 1)) * CLQ_MGS + (uint32_t)(8 * (2 * cbh + (mi >> 1))) * 16;
              uint32_t r[4];
#pragma unroll
              for (int x = 0; x < 4; ++x) {
                const int cb = 2 * cbh + (x >> 1), s2 = x & 1;
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4 * cb + 2 * s2], v[4 * cb + 2 * s2 + 1]);
                r[x] = *reinterpret_cast<const uint32_t*>(&h2);
              }
              ptx::stmatrix_x4_trans(addr, r[0], r[1], r[2], r[3]);
            }
          }
        } else {
          float v[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(g * 64 + ch * 32), v);
          const uint32_t base = gb + 2 * CLQ_BT + (uint32_t)wq * CLQ_MSG + (uint32_t)(lane >> 3) * CLQ_MGS + (uint32_t)(ch * 32) * 16 + (uint32_t)(lane & 7) * 2;
#pragma unroll
          for (int i = 0; i < 32; ++i) st_shared_u16(base + i * 16, __bfloat16_as_ushort(__float2bfloat16_rn(v[i])));
        }
        ptx::tc_fence_before();
        named_barrier(3 + g, 32 * (CL_EPI_WARPS + 1));
        if (tracer && g == 0) CL_TRACE(it, 9);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  if (warp == 1) ptx::tmem_dealloc2(tmem_base, 512);
}

// U (512, 2048) fp32 master -> bf16 [16 CTAs][128 lanes][512 k] for the quad backward kernel: CTA rank (kq = rank / 4, mt = (rank / 2) % 2,
// e = rank % 2), lane l: output unit 256 mt + 128 e + l; k = cs * 128 + gate * 32 + uu: U[unit, blk(gate) * 512 + 128 kq + 32 cs + uu]
__global__ void pack_u_cluster_bwd4_kernel(const float* __restrict__ U, int ldu, bf16* __restrict__ out, int variant) {
  const int H = 512;
  const long total = (long)H * 4 * H;
  for (long x = blockIdx.x * (long)blockDim.x + threadIdx.x; x < total; x += (long)gridDim.x * blockDim.x) {
    const int kk = (int)(x % 512);
    long rest = x / 512;
    const int l = (int)(rest % 128);
    const int rank = (int)(rest / 128);
    const int kq = rank >> 2, mt = (rank >> 1) & 1, e = rank & 1;
    const int cs = kk >> 7, gate = (kk >> 5) & 3, uu = kk & 31;
    int blk = gate;
    if (variant != MVAE_CELL_STANDARD && gate < 2) blk = 1 - gate;
    out[x] = __float2bfloat16_rn(U[(long)(256 * mt + 128 * e + l) * ldu + blk * H + 128 * kq + 32 * cs + uu]);
  }
}

// U (H, 4H) fp32 master -> bf16 [CS][MT][128][256] for the cluster backward kernel: CTA rank = 2 q + e, M tile mt, lane l:
// output unit 256 mt + 128 e + l; column kk = gate * 64 + uu (gate in semantic order i, f, g, o): U[unit, blk(gate) H + 64 q + uu]
__global__ void pack_u_cluster_bwd_kernel(const float* __restrict__ U, int ldu, bf16* __restrict__ out, int H, int variant) {
  const long total = (long)H * 4 * H;
  const int MT = H / 256;
  for (long x = blockIdx.x * (long)blockDim.x + threadIdx.x; x < total; x += (long)gridDim.x * blockDim.x) {
    const int kk = (int)(x % 256);
    long rest = x / 256;
    const int l = (int)(rest % 128); rest /= 128;
    const int mt = (int)(rest % MT);
    const int rank = (int)(rest / MT);
    const int q = rank >> 1, e = rank & 1;
    const int gate = kk / 64, uu = kk % 64;
    int blk = gate;
    if (variant != MVAE_CELL_STANDARD && gate < 2) blk = 1 - gate;
    out[x] = __float2bfloat16_rn(U[(long)(256 * mt + 128 * e + l) * ldu + blk * H + 64 * q + uu]);
  }
}

// U (H, 4H) fp32 master, Keras column blocks -> per-CTA bf16 [CS][128 rows][H], K-major:
// row mrow = gate * 32 + u of CTA j holds column blk(gate) * H + j * 32 + u of U (gate in semantic order i, f, g, o)
__global__ void pack_u_cluster_kernel(const float* __restrict__ U, int ldu, bf16* __restrict__ out, int H, int variant) {
  const long total = (long)H * 4 * H;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int k = (int)(e % H);
    const long rn = e / H;              // j * 128 + mrow
    const int mrow = (int)(rn % CL_GC), j = (int)(rn / CL_GC);
    const int gate = mrow / CL_HS, u = mrow % CL_HS;
    int blk = gate;
    if (variant != MVAE_CELL_STANDARD && gate < 2) blk = 1 - gate;
    out[e] = __float2bfloat16_rn(U[(long)k * ldu + blk * H + j * CL_HS + u]);
  }
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

template <int CS, bool HARD, bool STD>
void launch_fwd2(const RecPersistArgs& a, cudaStream_t st) {
  constexpr int H = CS * CL_HS;
  constexpr size_t hbuf = (size_t)CL_HALF * H * 2, scr_bytes = hbuf >= CL_SCR ? 0 : 2 * CL_SCR;
  constexpr size_t smem_max = 1024 + scr_bytes + (size_t)CL2_MAXG * 2 * hbuf;
  auto kern = rec_cluster_fwd2_kernel<CS, HARD, STD>;
  static bool configured = false;
  if (!configured) {
    MVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    if (CS > 8) MVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured = true;
  }
  const int groups = (a.n + CL_ROWS - 1) / CL_ROWS;
  int ng = env_int("MVAE_CL_NG", 0);
  if (ng <= 0) {
    // 2 groups per cluster is the fastest per step; when that needs more 16-CTA clusters than are ever co-resident (7 on a B200) a third
    // group per cluster (one warp set then serves two groups: ~1.65x per step) still beats a second wave (2x)
    ng = 2;
    if (CS == 16 && (groups + 1) / 2 > 7 && (groups + 2) / 3 <= 7) ng = 3;
    if (CS == 8 && groups <= env_int("MVAE_CL_NG1_MAX", 8)) ng = 1;   // 8-CTA clusters are plentiful: one group each has the shortest chain
  }
  ng = std::max(1, std::min(std::min(ng, CL2_MAXG), groups));
  const int clusters = (groups + ng - 1) / ng;
  const size_t smem = 1024 + scr_bytes + (size_t)ng * 2 * hbuf;
  Cluster2P p{};
  p.n = a.n; p.steps = a.steps; p.nswap = env_int("MVAE_CL_NSWAP", 0); p.ng = ng;
  p.xw = (const bf16*)a.xw;
  p.hseq = (bf16*)a.hseq; p.cseq = (bf16*)a.cseq; p.gates = (bf16*)a.gates;
  p.c0 = (const bf16*)a.c0; p.ldc0 = a.ldc0; p.no_stash = a.no_stash;
  p.upack = (const bf16*)a.upack; p.hx = (uint8_t*)a.hx; p.trace = (long long*)a.trace;
  p.x_mode = a.x_mode; p.xtab = (const bf16*)a.xtab; p.x_idx = a.x_idx; p.x_ld = a.x_ld; p.x_shift = a.x_shift;
  p.x_scalar = (const bf16*)a.x_scalar; p.x_w = a.x_w; p.x_b = a.x_b;
  MVAE_REQUIRE(p.x_mode == 0 ? p.xw != nullptr : (p.x_mode == 1 ? p.xtab != nullptr : (p.x_scalar && p.x_w && p.x_b)), "cluster forward: input projection source missing");
  MVAE_REQUIRE(p.hx != nullptr, "cluster forward: exchange buffer missing");
  MVAE_REQUIRE((size_t)clusters * ng * 2 * CS * CL_STAGE <= rec_cluster_hx_bytes(a.n, H), "cluster forward: exchange buffer too small");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(clusters * CS)); cfg.blockDim = dim3(CLF_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  static bool reported = false;
  if (!reported && env_int("MVAE_CL_VERBOSE", 0)) {
    int nc = -1;
    cudaError_t err = cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
    fprintf(stderr, "rec_cluster_fwd2<%d>: smem %zu B, ng %d, max co-resident clusters %d (%s), launching %d\n", CS, smem, ng, nc, cudaGetErrorString(err), clusters);
    reported = true;
  }
  MVAE_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  count_launch();
}


template <int CS, bool HARD, bool STD>
void launch_bwd(const RecPersistArgs& a, cudaStream_t st) {
  constexpr int H = CS * CL_HS, MT = H / 256, ND = 4 * MT, NP = CS / 2;
  constexpr size_t grp = CLB_BT + (size_t)(ND + NP) * CLB_MSG;
  constexpr size_t smem_max = 1024 + CLB_MAXG * grp;
  auto kern = rec_cluster_bwd_kernel<CS, HARD, STD>;
  static bool configured = false;
  if (!configured) {
    MVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    if (CS > 8) MVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured = true;
  }
  const int groups = (a.n + CL_ROWS - 1) / CL_ROWS;
  int ng = env_int("MVAE_CLB_NG", 0);
  if (ng <= 0) ng = (CS == 8 && groups <= env_int("MVAE_CL_NG1_MAX", 8)) ? 1 : CLB_MAXG;
  ng = std::max(1, std::min(std::min(ng, CLB_MAXG), groups));
  const int clusters = (groups + ng - 1) / ng;
  const size_t smem = 1024 + (size_t)ng * grp;
  ClusterBP p{};
  p.n = a.n; p.steps = a.steps; p.nswap = env_int("MVAE_CL_NSWAP", 0); p.ng = ng;
  p.gates = (const bf16*)a.gates; p.cseq = (const bf16*)a.cseq; p.dhext = (const bf16*)a.dhext;
  p.dh_last = (const bf16*)a.dh_last; p.ld_last = a.ld_last;
  p.dG = (bf16*)a.dG; p.dS_h = (bf16*)a.dS_h; p.dS_c = (bf16*)a.dS_c; p.ldS = a.ldS;
  p.upack = (const bf16*)a.upack_bwd; p.trace = (long long*)a.trace;
  MVAE_REQUIRE(p.upack != nullptr, "cluster backward: packed weights missing");
  p.l2_prefetch = env_int("MVAE_CLB_PREFETCH", 2);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(clusters * CS)); cfg.blockDim = dim3(CLW_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  static bool reported = false;
  if (!reported && env_int("MVAE_CL_VERBOSE", 0)) {
    int nc = -1;
    cudaError_t err = cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
    fprintf(stderr, "rec_cluster_bwd<%d>: smem %zu B, ng %d, max co-resident clusters %d (%s), launching %d\n", CS, smem, ng, nc, cudaGetErrorString(err), clusters);
    reported = true;
  }
  MVAE_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  count_launch();
}


template <bool HARD, bool STD>
void launch_bwd4(const RecPersistArgs& a, cudaStream_t st) {
  constexpr int CS = 16, H = 512;
  constexpr size_t smem_max = 1024 + (size_t)CLB_MAXG * CLQ_GRP;
  auto kern = rec_cluster_bwd4_kernel<HARD, STD>;
  static bool configured = false;
  if (!configured) {
    MVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    MVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured = true;
  }
  const int groups = (a.n + CL_ROWS - 1) / CL_ROWS;
  int ng = env_int("MVAE_CLB_NG", 0);
  if (ng <= 0) ng = CLB_MAXG;
  ng = std::max(1, std::min(std::min(ng, CLB_MAXG), groups));
  const int clusters = (groups + ng - 1) / ng;
  const size_t smem = 1024 + (size_t)ng * CLQ_GRP;
  ClusterQP p{};
  p.n = a.n; p.steps = a.steps; p.nswap = env_int("MVAE_CL_NSWAP", 0); p.ng = ng; p.l2_prefetch = env_int("MVAE_CLB_PREFETCH", 2);
  p.gates = (const bf16*)a.gates; p.cseq = (const bf16*)a.cseq; p.dhext = (const bf16*)a.dhext;
  p.dh_last = (const bf16*)a.dh_last; p.ld_last = a.ld_last;
  p.dG = (bf16*)a.dG; p.dS_h = (bf16*)a.dS_h; p.dS_c = (bf16*)a.dS_c; p.ldS = a.ldS;
  p.upack = (const bf16*)a.upack_bwd; p.trace = (long long*)a.trace; p.xbuf = (uint8_t*)a.partial;
  p.stm = env_int("MVAE_CLB_STM", 1);
  MVAE_REQUIRE(p.upack != nullptr && p.xbuf != nullptr, "cluster backward: packed weights / exchange buffer missing");
  MVAE_REQUIRE((size_t)clusters * ng * 2 * CS * 2 * CLQ_PIECE <= rec_cluster_xbuf_bytes(a.n, H), "cluster backward: exchange buffer too small");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(clusters * CS)); cfg.blockDim = dim3(CLW_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  MVAE_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  count_launch();
}

}  // namespace

// cluster kernels: H = 32 * cluster size, cluster size 8 (portable) or 16 (non-portable, one cluster per GPC)
bool rec_cluster_supported(int H) {
  static int enabled = -1;
  if (enabled < 0) enabled = env_int("MVAE_REC_CLUSTER", 1);
  return enabled && (H == 256 || H == 512);
}

// global dG exchange slots of the quad-form backward kernel: [64-row groups rounded up to whole clusters][2][H/32 CTAs][2 row halves][8 KB]
size_t rec_cluster_xbuf_bytes(int n, int H) {
  const size_t groups = (size_t)((n + CL_ROWS - 1) / CL_ROWS + CLB_MAXG), ctas = (size_t)(H / CL_HS);
  return groups * 2 * ctas * 2 * (size_t)CLQ_PIECE;
}

// global exchange buffer of the forward kernel: [64-row groups, rounded up to whole clusters][2][H/32 CTAs][4 KB]
size_t rec_cluster_hx_bytes(int n, int H) { return (size_t)((n + CL_ROWS - 1) / CL_ROWS + CL2_MAXG) * 2 * (size_t)(H / CL_HS) * CL_STAGE; }

void rec_cluster_pack_u(const float* U, int ldu, void* upack, int H, int variant, cudaStream_t st) {
  MVAE_REQUIRE(H == 256 || H == 512, "cluster recurrence: hidden size 256 or 512");
  pack_u_cluster_kernel<<<std::min(148 * 8, (int)(((long)H * 4 * H + 255) / 256)), 256, 0, st>>>(U, ldu, (bf16*)upack, H, variant);
  count_launch();
  MVAE_CUDA(cudaGetLastError());
}

// which backward form runs: the quad form exists for H = 512 only
static bool use_bwd4(int H) { return H == 512 && env_int("MVAE_CLB_V", 4) == 4; }

void rec_cluster_pack_u_bwd(const float* U, int ldu, void* upack_bwd, int H, int variant, cudaStream_t st) {
  MVAE_REQUIRE(H == 256 || H == 512, "cluster recurrence: hidden size 256 or 512");
  if (use_bwd4(H)) {
    pack_u_cluster_bwd4_kernel<<<148 * 8, 256, 0, st>>>(U, ldu, (bf16*)upack_bwd, variant);
    count_launch();
    MVAE_CUDA(cudaGetLastError());
    return;
  }
  pack_u_cluster_bwd_kernel<<<std::min(148 * 8, (int)(((long)H * 4 * H + 255) / 256)), 256, 0, st>>>(U, ldu, (bf16*)upack_bwd, H, variant);
  count_launch();
  MVAE_CUDA(cudaGetLastError());
}

// x_mode 1 table: row i < din = bf16( bf16(W[i, :]) + b ) (exactly what the bf16 GEMM of a one-hot row stores), rows din..63 = bf16(b)
__global__ void build_xtab_kernel(const bf16* __restrict__ W, int ldw, int din, const float* __restrict__ bias, bf16* __restrict__ out, int G) {
  const int total = 64 * G;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < total; x += gridDim.x * blockDim.x) {
    const int i = x / G, c = x % G;
    const float w = i < din ? __bfloat162float(W[(size_t)i * ldw + c]) : 0.f;
    out[x] = __float2bfloat16_rn(w + bias[c]);
  }
}
void rec_cluster_build_xtab(const void* W_bf16, int ldw, int din, const float* bias, void* xtab, int H, cudaStream_t st) {
  MVAE_REQUIRE(din <= 63, "one-hot input projection table: at most 63 classes");
  build_xtab_kernel<<<64, 256, 0, st>>>((const bf16*)W_bf16, ldw, din, bias, (bf16*)xtab, 4 * H);
  count_launch();
  MVAE_CUDA(cudaGetLastError());
}

bool rec_cluster_bwd_supported(int H) {
  static int enabled = -1;
  if (enabled < 0) enabled = env_int("MVAE_REC_CLUSTER_BWD", 1);
  return enabled && rec_cluster_supported(H);
}

void rec_cluster_backward(const RecPersistArgs& a, cudaStream_t st) {
  const bool hard = a.gate_act == MVAE_GATE_HARD_SIGMOID, stdc = a.variant == MVAE_CELL_STANDARD;
  MVAE_REQUIRE(a.H == 512 || a.H == 256, "cluster recurrence unsupported for this hidden size");
  if (use_bwd4(a.H)) {
    if (hard && stdc) launch_bwd4<true, true>(a, st);
    else if (hard) launch_bwd4<true, false>(a, st);
    else if (stdc) launch_bwd4<false, true>(a, st);
    else launch_bwd4<false, false>(a, st);
    return;
  }
#define MVAE_CL_BWD(CS)                                                        \
  do {                                                                         \
    if (hard && stdc) launch_bwd<CS, true, true>(a, st);                       \
    else if (hard) launch_bwd<CS, true, false>(a, st);                         \
    else if (stdc) launch_bwd<CS, false, true>(a, st);                         \
    else launch_bwd<CS, false, false>(a, st);                                  \
  } while (0)
  if (a.H == 512) MVAE_CL_BWD(16);
  else MVAE_CL_BWD(8);
#undef MVAE_CL_BWD
}

void rec_cluster_forward(const RecPersistArgs& a, cudaStream_t st) {
  const bool hard = a.gate_act == MVAE_GATE_HARD_SIGMOID, stdc = a.variant == MVAE_CELL_STANDARD;
  MVAE_REQUIRE(a.H == 512 || a.H == 256, "cluster recurrence unsupported for this hidden size");
#define MVAE_CL_FWD2(CS)                                                       \
  do {                                                                         \
    if (hard && stdc) launch_fwd2<CS, true, true>(a, st);                      \
    else if (hard) launch_fwd2<CS, true, false>(a, st);                        \
    else if (stdc) launch_fwd2<CS, false, true>(a, st);                        \
    else launch_fwd2<CS, false, false>(a, st);                                 \
  } while (0)
  if (a.H == 512) MVAE_CL_FWD2(16);
  else MVAE_CL_FWD2(8);
#undef MVAE_CL_FWD2
}

}  // namespace mvae
