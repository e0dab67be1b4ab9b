// gru_cluster.cu -- cluster-resident GRU recurrence, forward and reverse sweep, for H = 256 in bf16: the cell the reference runs by default
// (settings.py:155 cell_type = 'GRU', lstm_size = 256; Keras GRU encoders vae_definition.py:457-472, recurrentshop GRUCell decoders :535-623).
//
// A GRU step is TWO dependent products -- [z | r] = act(x W_zr + h U_zr), then hh = tanh(x W_h + (r * h) U_h) -- so the step-streamed form is
// four tiny dependent launches per step (two GEMMs, two pointwise kernels: ~44 us per step, launch- and latency-bound).  Here one launch runs
// the whole sequence:
//   * a 4-CTA thread-block cluster owns 16 or 32 batch rows for all T steps; clusters never talk to each other;
//   * CTA c owns 64 hidden units: its columns of U_zr and U_h (forward: transposed, [unit][k]) or its rows (reverse sweep: dh_{t-1} needs
//     U^T, i.e. the natural rows) stay in shared memory for the whole sequence (96 KB of bf16);
//   * the products are warp-level tensor-core MMAs (mma.sync m16n8k16, bf16 in, fp32 accumulate): the tiles are 32 rows x 64..128 columns per
//     CTA and step, far below what a tcgen05 / TMEM pipeline needs to pay for its set-up, and the step is bound by the two exchanges, not by math;
//     warp w owns units [8 w, 8 w + 8) of the CTA in every gate, so z, r, hh, h of a (row, unit) live in the same thread's registers;
//   * what the NEXT product needs from every CTA -- r * h, then h_t (forward); da_h, then [da_z | da_r] (reverse) -- is exchanged through
//     distributed shared memory: each CTA writes its 64-unit block into its own copy of the operand tile and pushes it to the three peers with one
//     bulk copy each (cp.async.bulk shared::cta -> shared::cluster, complete_tx on the receiver's mbarrier); a block is re-written only after
//     every reader's next message has arrived, which the data dependencies of the recurrence already guarantee (see the hazard notes below).
// Stash layouts are those of the step-streamed GRU path (model.cu::gru_steps_forward): xw / gates / dG (T, n, 3H) row-major [z | r | h],
// hseq (T + 1, n, H), r * h in the cseq slabs, so either path can run the other's backward.  Math: kernels.cu gru_* (oracle/manual_bptt.py).
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

constexpr int GC_H = 256, GC_CS = 4, GC_HU = GC_H / GC_CS;      // hidden size, cluster size, units per CTA
// batch rows per cluster: R = 16 MT (MT m16 tiles; template parameter): 16 while that fills at most one wave of 4-CTA clusters, else 32
constexpr int GC_THREADS = 256;                                 // 8 warps x 8 units
constexpr int GC_BLK_LD = GC_HU + 8;                            // operand-tile block: [R rows][64 units + 8 pad] bf16 (144-byte rows: conflict-free ldmatrix),
                                                                // R x 144 B = one bulk copy
constexpr int GC_BLK2_LD = 2 * GC_HU + 8;                       // reverse sweep, second exchange: [R rows][z 64 | r 64 | 8 pad]
constexpr int GC_W_LD = GC_H + 8;                               // weight rows: [unit][256 k + 8 pad] (528-byte rows)
constexpr int GC_W2_LD = 2 * GC_H + 8;                          // reverse sweep: U_zr rows [unit][512 k' + 8 pad]

struct GruP {
  int n, t0, t1, gate_act, mix;
  const bf16* U; int ldu;           // (H, 3H) bf16 shadow of the recurrent kernel [z | r | h]
  const bf16* xw;                   // (T, n, 3H) x W + b
  bf16* hseq;                       // (T + 1, n, H): slab t0 read, slabs t0 + 1 .. t1 written
  bf16* gates;                      // (T, n, 3H) [z | r | hh]
  bf16* rh;                         // (T, n, H) r * h_{t-1}
  // reverse sweep
  const bf16* dhext;                // (T, n, H) or null
  const bf16* dh_last; int ld_last; // (n, ld_last) or null: extra gradient into the last step's h
  bf16* dG;                         // (T, n, 3H) [da_z | da_r | da_h]
  bf16* dS_h; int ldS;              // (n, ldS) gradient wrt the initial state, or null
};

__device__ __forceinline__ float gate_fn(int gate_act, float x) {
  return gate_act == MVAE_GATE_HARD_SIGMOID ? fminf(fmaxf(0.2f * x + 0.5f, 0.f), 1.f) : 1.f / (1.f + __expf(-x));
}
__device__ __forceinline__ float gate_grad(int gate_act, float s) {
  return gate_act == MVAE_GATE_HARD_SIGMOID ? ((s > 0.f && s < 1.f) ? 0.2f : 0.f) : s * (1.f - s);
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_bf16(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack2(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u)); }
__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// acc[w][mt][4] += A(tile, 16 MT rows x K) * B_w^T(this warp's 8 units x K) for NW weight sets that share the A fragments, over NB k-blocks of 64
// (one block = one source CTA's units).  tile: blocks of [16 MT rows][blk_ld] bf16 `blk_bytes` apart, of each block row the k sub-range
// [koff, koff + 64) is used; wrow[w]: address of this warp's first weight row of set w, rows w_ld elements apart; the k-block kb of the tile
// meets weight elements [wk0 + kb * wkstride, + 64).
template <int NB, int NW, int MT>
__device__ __forceinline__ void mma_rows(float (&acc)[NW][MT][4], uint32_t tile, uint32_t blk_bytes, int blk_ld, int koff, const uint32_t (&wrow)[NW],
                                         int w_ld, int wk0, int wkstride, int lane) {
  // ldmatrix row addresses: A (x4: rows 0-7 / 8-15 x k 0-7 / 8-15 of an m16 x k16 tile), B (x4: this warp's 8 units x k 0-7 / 8-15 / 16-23 / 24-31)
  const uint32_t a_lane = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * blk_ld + (lane >> 4) * 8) * 2u;
  const uint32_t b_lane = (uint32_t)((lane & 7) * w_ld + (lane >> 3) * 8) * 2u;
#pragma unroll
  for (int kb = 0; kb < NB; ++kb) {
    const uint32_t ab = tile + (uint32_t)kb * blk_bytes + (uint32_t)koff * 2u + a_lane;
#pragma unroll
    for (int k2 = 0; k2 < 2; ++k2) {           // 32 k per iteration: one B ldmatrix.x4 feeds two k16 steps
      uint32_t b[NW][4];
#pragma unroll
      for (int w = 0; w < NW; ++w)
        ldsm_x4(wrow[w] + (uint32_t)(wk0 + kb * wkstride + k2 * 32) * 2u + b_lane, b[w][0], b[w][1], b[w][2], b[w][3]);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          uint32_t a0, a1, a2, a3;
          ldsm_x4(ab + (uint32_t)(mt * 16 * blk_ld + k2 * 32 + ks * 16) * 2u, a0, a1, a2, a3);
#pragma unroll
          for (int w = 0; w < NW; ++w) mma_bf16(acc[w][mt], a0, a1, a2, a3, b[w][2 * ks], b[w][2 * ks + 1]);
        }
      }
    }
  }
}

// push this CTA's block (already written to its own copy of the tile) to the three peers; one thread
__device__ __forceinline__ void push_block(uint32_t blk_addr, uint32_t bytes, uint32_t bar_addr, uint32_t rank) {
#pragma unroll
  for (uint32_t d = 1; d < GC_CS; ++d) {
    const uint32_t peer = (rank + d) % GC_CS;
    ptx::bulk_copy_dsmem(ptx::mapa(blk_addr, peer), blk_addr, bytes, ptx::mapa(bar_addr, peer));
  }
}

// Hazard notes (why single-buffered operand tiles and two mbarriers per CTA are enough).  Every step has two exchanges, X1 then X2 (forward: r*h, h;
// reverse: da_h, [da_z | da_r]); product P1 reads the X2 tile of the previous step, product P2 reads the X1 tile of this step.
//   * A peer's X1 block of step t+1 can only be produced after that peer received MY X2 block of step t, which I push after all my warps finished
//     P2 of step t (the __syncthreads in front of the push): nobody overwrites the X1 tile while I still read it.  Symmetrically a peer's X2
//     block of step t needs my X1 block of step t, pushed after all my warps finished P1 of step t: the X2 tile is not overwritten under P1.
//   * I re-write my OWN block of a tile one step after I pushed it; by then every peer has used it (its next message reached me, see above), i.e.
//     the bulk copies that read it have completed (complete_tx is signalled after the data has been written at the destination).
//   * An mbarrier phase cannot be completed early: thread 0 arms both barriers (arrive.expect_tx of the three peers' bytes) at the top of the step, and
//     a peer's bytes for the NEXT phase cannot arrive before this CTA pushed its block of the current one, which happens after every thread passed the
//     wait of the previous phase.  Bytes that land before the arming only make the transaction count transiently negative.
//   * The final cluster barrier keeps every CTA (its shared memory, its barriers) alive until all peers have seen their last message.

// ------------------------------------------------------------------------------------------------------------------ forward
// shared memory: W_zr^T [128][264] | W_h^T [64][264] | h tile [4 blocks] | r*h tile [4 blocks]
constexpr uint32_t GF_WZR = 0, GF_WH = GF_WZR + 2 * GC_HU * GC_W_LD * 2, GF_HT = GF_WH + GC_HU * GC_W_LD * 2;
constexpr uint32_t gf_total(int mt) { return GF_HT + 2 * GC_CS * (uint32_t)(16 * mt * GC_BLK_LD * 2); }

template <int MT>
__global__ void __launch_bounds__(GC_THREADS, 1) gru_cluster_fwd_kernel(const GruP p) {
  constexpr int H = GC_H, G = 3 * GC_H, GC_R = 16 * MT;
  constexpr uint32_t GC_BLK = GC_R * GC_BLK_LD * 2, GF_RHT = GF_HT + GC_CS * GC_BLK;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_rh, bar_h;
  const uint32_t sm = (ptx::smem_u32(smem_raw) + 127u) & ~127u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const uint32_t rank = ptx::cluster_ctarank();
  const int row0 = ((int)blockIdx.x / GC_CS) * GC_R;
  const int n = p.n;

  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar_rh), 1);
    ptx::mbar_init(ptx::smem_u32(&bar_h), 1);
    ptx::fence_barrier_init();
  }
  // recurrent weights of this CTA's units, transposed: W_zr^T[nn][k] = U[k][gate(nn) * H + 64 rank + nn % 64], W_h^T[nn][k] = U[k][2H + 64 rank + nn]
#pragma unroll 4
  for (int e = threadIdx.x; e < H * 3 * (GC_HU / 8); e += GC_THREADS) {
    const int k = e / (3 * (GC_HU / 8)), c = e % (3 * (GC_HU / 8)), gate = c / (GC_HU / 8), u8 = c % (GC_HU / 8);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.U + (size_t)k * p.ldu + gate * H + (int)rank * GC_HU + u8 * 8));
    const unsigned short* hv = reinterpret_cast<const unsigned short*>(&v);
    const uint32_t dst = (gate < 2 ? sm + GF_WZR + (uint32_t)((gate * GC_HU + u8 * 8) * GC_W_LD + k) * 2u : sm + GF_WH + (uint32_t)(u8 * 8 * GC_W_LD + k) * 2u);
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("st.shared.u16 [%0], %1;" ::"r"(dst + (uint32_t)(i * GC_W_LD) * 2u), "h"(hv[i]) : "memory");
  }
  // h_{t0} (all 256 units of the cluster's rows) -> every CTA's own h tile; rows past n are zero
  for (int e = threadIdx.x; e < GC_R * (H / 8); e += GC_THREADS) {
    const int r = e / (H / 8), c8 = e % (H / 8), m = row0 + r;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (m < n) v = *reinterpret_cast<const uint4*>(p.hseq + ((size_t)p.t0 * n + m) * H + c8 * 8);
    ptx::st_shared_u4(sm + GF_HT + (uint32_t)(c8 / 8) * GC_BLK + (uint32_t)(r * GC_BLK_LD + (c8 % 8) * 8) * 2u, v);
  }
  __syncthreads();
  ptx::cluster_arrive();
  ptx::cluster_wait();

  // this thread's 8 (row, unit pair) elements: m-tile mt, row half hf: row = 16 mt + 8 hf + g, units u0 + {0, 1}
  const int u0 = warp * 8 + 2 * q;                       // within the CTA
  const int gu0 = (int)rank * GC_HU + u0;                // global unit
  float hprev[MT][2][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int m = row0 + 16 * mt + 8 * hf + g;
      float2 v = make_float2(0.f, 0.f);
      if (m < n) v = unpack2(*reinterpret_cast<const uint32_t*>(p.hseq + ((size_t)p.t0 * n + m) * H + gu0));
      hprev[mt][hf][0] = v.x; hprev[mt][hf][1] = v.y;
    }
  const uint32_t my_h = sm + GF_HT + rank * GC_BLK, my_rh = sm + GF_RHT + rank * GC_BLK;
  const uint32_t wzr_rows[2] = {sm + GF_WZR + (uint32_t)(warp * 8 * GC_W_LD) * 2u, sm + GF_WZR + (uint32_t)((GC_HU + warp * 8) * GC_W_LD) * 2u};
  const uint32_t wh_rows[1] = {sm + GF_WH + (uint32_t)(warp * 8 * GC_W_LD) * 2u};

  for (int t = p.t0; t < p.t1; ++t) {
    const uint32_t par = (uint32_t)((t - p.t0) & 1);
    if (threadIdx.x == 0) {
      ptx::mbar_arrive_expect_tx(ptx::smem_u32(&bar_rh), (GC_CS - 1) * GC_BLK);
      ptx::mbar_arrive_expect_tx(ptx::smem_u32(&bar_h), (GC_CS - 1) * GC_BLK);
    }
    // input projections of this step (independent of the chain: in flight during the first product)
    uint32_t xz[MT][2], xr[MT][2], xh[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int m = row0 + 16 * mt + 8 * hf + g;
        xz[mt][hf] = xr[mt][hf] = xh[mt][hf] = 0u;
        if (m < n) {
          const bf16* xp = p.xw + ((size_t)t * n + m) * G + gu0;
          xz[mt][hf] = __ldg(reinterpret_cast<const uint32_t*>(xp));
          xr[mt][hf] = __ldg(reinterpret_cast<const uint32_t*>(xp + H));
          xh[mt][hf] = __ldg(reinterpret_cast<const uint32_t*>(xp + 2 * H));
        }
      }
    // ---- [z | r] = act(xw + h_{t-1} U_zr)
    float azr[2][MT][4];
#pragma unroll
    for (int w = 0; w < 2; ++w)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int i = 0; i < 4; ++i) azr[w][mt][i] = 0.f;
    mma_rows<GC_CS, 2, MT>(azr, sm + GF_HT, GC_BLK, GC_BLK_LD, 0, wzr_rows, GC_W_LD, 0, GC_HU, lane);
    const float (&az)[MT][4] = azr[0];
    const float (&ar)[MT][4] = azr[1];
    float zv[MT][2][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int m = row0 + 16 * mt + 8 * hf + g;
        const float2 x1 = unpack2(xz[mt][hf]), x2 = unpack2(xr[mt][hf]);
        const float z0 = gate_fn(p.gate_act, az[mt][2 * hf] + x1.x), z1 = gate_fn(p.gate_act, az[mt][2 * hf + 1] + x1.y);
        const float r0 = gate_fn(p.gate_act, ar[mt][2 * hf] + x2.x), r1 = gate_fn(p.gate_act, ar[mt][2 * hf + 1] + x2.y);
        const uint32_t zp = pack2(z0, z1), rp = pack2(r0, r1);
        // the candidate and the mix use the gates as the stash holds them (bf16), like the step-streamed kernels
        const float2 zb = unpack2(zp), rb = unpack2(rp);
        zv[mt][hf][0] = zb.x; zv[mt][hf][1] = zb.y;
        const uint32_t rhp = pack2(rb.x * hprev[mt][hf][0], rb.y * hprev[mt][hf][1]);
        st_shared_u32(my_rh + (uint32_t)((16 * mt + 8 * hf + g) * GC_BLK_LD + u0) * 2u, rhp);
        if (m < n) {
          bf16* gp = p.gates + ((size_t)t * n + m) * G + gu0;
          *reinterpret_cast<uint32_t*>(gp) = zp;
          *reinterpret_cast<uint32_t*>(gp + H) = rp;
          *reinterpret_cast<uint32_t*>(p.rh + ((size_t)t * n + m) * H + gu0) = rhp;
        }
      }
    ptx::fence_proxy_async();
    __syncthreads();                                    // own r*h block complete; every warp is done reading the h tile
    if (threadIdx.x == 0) push_block(my_rh, GC_BLK, ptx::smem_u32(&bar_rh), rank);
    ptx::mbar_wait(ptx::smem_u32(&bar_rh), par);        // the three peers' blocks have landed
    // ---- hh = tanh(xw_h + (r * h_{t-1}) U_h);  h_t = mix(z, h_{t-1}, hh)
    float ahh[1][MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) ahh[0][mt][i] = 0.f;
    mma_rows<GC_CS, 1, MT>(ahh, sm + GF_RHT, GC_BLK, GC_BLK_LD, 0, wh_rows, GC_W_LD, 0, GC_HU, lane);
    const float (&ah)[MT][4] = ahh[0];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int m = row0 + 16 * mt + 8 * hf + g;
        const float2 x3 = unpack2(xh[mt][hf]);
        const uint32_t hhp = pack2(tanhf(ah[mt][2 * hf] + x3.x), tanhf(ah[mt][2 * hf + 1] + x3.y));
        const float2 hh = unpack2(hhp);
        float hn[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float z = zv[mt][hf][i], h = hprev[mt][hf][i], c = i ? hh.y : hh.x;
          hn[i] = p.mix == 0 ? z * h + (1.f - z) * c : (1.f - z) * h + z * c;
        }
        const uint32_t hp = pack2(hn[0], hn[1]);
        const float2 hb = unpack2(hp);
        hprev[mt][hf][0] = hb.x; hprev[mt][hf][1] = hb.y;
        st_shared_u32(my_h + (uint32_t)((16 * mt + 8 * hf + g) * GC_BLK_LD + u0) * 2u, hp);
        if (m < n) {
          *reinterpret_cast<uint32_t*>(p.gates + ((size_t)t * n + m) * G + 2 * H + gu0) = hhp;
          *reinterpret_cast<uint32_t*>(p.hseq + ((size_t)(t + 1) * n + m) * H + gu0) = hp;
        }
      }
    ptx::fence_proxy_async();
    __syncthreads();                                    // own h block complete; every warp is done reading the r*h tile
    if (threadIdx.x == 0) push_block(my_h, GC_BLK, ptx::smem_u32(&bar_h), rank);
    ptx::mbar_wait(ptx::smem_u32(&bar_h), par);
  }
  // no CTA leaves while a peer's copy may still read its shared memory / write into it
  __syncthreads();
  ptx::cluster_arrive();
  ptx::cluster_wait();
}

// ------------------------------------------------------------------------------------------------------------------ reverse sweep
// shared memory: U_h rows [64][264] | U_zr rows [64][520] (k' = 128 block + 64 gate + unit) | da_h tile [4 blocks] | [da_z | da_r] tile [4 blocks]
constexpr uint32_t GB_WH = 0, GB_WZR = GB_WH + GC_HU * GC_W_LD * 2, GB_T1 = GB_WZR + GC_HU * GC_W2_LD * 2;
constexpr uint32_t gb_total(int mt) { return GB_T1 + GC_CS * (uint32_t)(16 * mt * (GC_BLK_LD + GC_BLK2_LD) * 2); }

template <int MT>
__global__ void __launch_bounds__(GC_THREADS, 1) gru_cluster_bwd_kernel(const GruP p) {
  constexpr int H = GC_H, G = 3 * GC_H, GC_R = 16 * MT;
  constexpr uint32_t GC_BLK = GC_R * GC_BLK_LD * 2, GC_BLK2 = GC_R * GC_BLK2_LD * 2, GB_T2 = GB_T1 + GC_CS * GC_BLK;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_1, bar_2;
  const uint32_t sm = (ptx::smem_u32(smem_raw) + 127u) & ~127u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const uint32_t rank = ptx::cluster_ctarank();
  const int row0 = ((int)blockIdx.x / GC_CS) * GC_R;
  const int n = p.n;

  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar_1), 1);
    ptx::mbar_init(ptx::smem_u32(&bar_2), 1);
    ptx::fence_barrier_init();
  }
  // rows of U for this CTA's units j: drh[j] = sum_k da_h[k] U[64 rank + j][2H + k];  dh[j] += sum_k' [da_z | da_r][k'] U[64 rank + j][gate H + 64 blk + u]
  for (int e = threadIdx.x; e < GC_HU * (H / 8); e += GC_THREADS) {
    const int j = e / (H / 8), c8 = e % (H / 8);
    const bf16* urow = p.U + (size_t)((int)rank * GC_HU + j) * p.ldu;
    ptx::st_shared_u4(sm + GB_WH + (uint32_t)(j * GC_W_LD + c8 * 8) * 2u, *reinterpret_cast<const uint4*>(urow + 2 * H + c8 * 8));
#pragma unroll
    for (int gate = 0; gate < 2; ++gate) {
      const int blk = c8 / 8, u8 = c8 % 8;               // unit 64 blk + 8 u8 of gate `gate`
      ptx::st_shared_u4(sm + GB_WZR + (uint32_t)(j * GC_W2_LD + blk * 2 * GC_HU + gate * GC_HU + u8 * 8) * 2u,
                        *reinterpret_cast<const uint4*>(urow + gate * H + c8 * 8));
    }
  }
  __syncthreads();
  ptx::cluster_arrive();
  ptx::cluster_wait();

  const int u0 = warp * 8 + 2 * q, gu0 = (int)rank * GC_HU + u0;
  float dh[MT][2][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) dh[mt][hf][0] = dh[mt][hf][1] = 0.f;
  const uint32_t my_1 = sm + GB_T1 + rank * GC_BLK, my_2 = sm + GB_T2 + rank * GC_BLK2;
  const uint32_t wh_rows[1] = {sm + GB_WH + (uint32_t)(warp * 8 * GC_W_LD) * 2u}, wzr_rows[1] = {sm + GB_WZR + (uint32_t)(warp * 8 * GC_W2_LD) * 2u};

  for (int t = p.t1 - 1; t >= p.t0; --t) {
    const uint32_t par = (uint32_t)((p.t1 - 1 - t) & 1);
    if (threadIdx.x == 0) {
      ptx::mbar_arrive_expect_tx(ptx::smem_u32(&bar_1), (GC_CS - 1) * GC_BLK);
      ptx::mbar_arrive_expect_tx(ptx::smem_u32(&bar_2), (GC_CS - 1) * GC_BLK2);
    }
    // ---- part 1 (gru_bwd1): dh = carried + external;  da_z, da_h, direct path
    float rv[MT][2][2], hv[MT][2][2], daz[MT][2][2];
    uint32_t dahp[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int m = row0 + 16 * mt + 8 * hf + g;
        float2 z = make_float2(0.f, 0.f), r = z, hh = z, h = z, ex = z;
        if (m < n) {
          const bf16* gp = p.gates + ((size_t)t * n + m) * G + gu0;
          z = unpack2(__ldg(reinterpret_cast<const uint32_t*>(gp)));
          r = unpack2(__ldg(reinterpret_cast<const uint32_t*>(gp + H)));
          hh = unpack2(__ldg(reinterpret_cast<const uint32_t*>(gp + 2 * H)));
          h = unpack2(__ldg(reinterpret_cast<const uint32_t*>(p.hseq + ((size_t)t * n + m) * H + gu0)));
          if (p.dhext) ex = unpack2(__ldg(reinterpret_cast<const uint32_t*>(p.dhext + ((size_t)t * n + m) * H + gu0)));
          if (p.dh_last && t == p.t1 - 1) {
            const float2 l = unpack2(__ldg(reinterpret_cast<const uint32_t*>(p.dh_last + (size_t)m * p.ld_last + gu0)));
            ex.x += l.x; ex.y += l.y;
          }
        }
        rv[mt][hf][0] = r.x; rv[mt][hf][1] = r.y; hv[mt][hf][0] = h.x; hv[mt][hf][1] = h.y;
        float dah[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float d = dh[mt][hf][i] + (i ? ex.y : ex.x), zz = i ? z.y : z.x, cc = i ? hh.y : hh.x, hp = i ? h.y : h.x;
          float dz, dhh, direct;
          if (p.mix == 0) { dz = d * (hp - cc); dhh = d * (1.f - zz); direct = d * zz; }
          else { dz = d * (cc - hp); dhh = d * zz; direct = d * (1.f - zz); }
          daz[mt][hf][i] = dz * gate_grad(p.gate_act, zz);
          dah[i] = dhh * (1.f - cc * cc);
          dh[mt][hf][i] = direct;
        }
        dahp[mt][hf] = pack2(dah[0], dah[1]);
        st_shared_u32(my_1 + (uint32_t)((16 * mt + 8 * hf + g) * GC_BLK_LD + u0) * 2u, dahp[mt][hf]);
      }
    ptx::fence_proxy_async();
    __syncthreads();                                    // own da_h block complete; every warp is done reading the [da_z | da_r] tile
    if (threadIdx.x == 0) push_block(my_1, GC_BLK, ptx::smem_u32(&bar_1), rank);
    ptx::mbar_wait(ptx::smem_u32(&bar_1), par);
    // ---- drh = da_h U_h^T;  part 2 (gru_bwd2): da_r = drh h_{t-1} act'(r);  dh += drh r
    float a1s[1][MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) a1s[0][mt][i] = 0.f;
    mma_rows<GC_CS, 1, MT>(a1s, sm + GB_T1, GC_BLK, GC_BLK_LD, 0, wh_rows, GC_W_LD, 0, GC_HU, lane);
    const float (&a1)[MT][4] = a1s[0];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int m = row0 + 16 * mt + 8 * hf + g;
        float dar[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float d = a1[mt][2 * hf + i], r = rv[mt][hf][i];
          dar[i] = d * hv[mt][hf][i] * gate_grad(p.gate_act, r);
          dh[mt][hf][i] += d * r;
        }
        const uint32_t zp = pack2(daz[mt][hf][0], daz[mt][hf][1]), rp = pack2(dar[0], dar[1]);
        const uint32_t rowa = my_2 + (uint32_t)((16 * mt + 8 * hf + g) * GC_BLK2_LD + u0) * 2u;
        st_shared_u32(rowa, zp);
        st_shared_u32(rowa + GC_HU * 2u, rp);
        if (m < n) {
          bf16* dg = p.dG + ((size_t)t * n + m) * G + gu0;
          *reinterpret_cast<uint32_t*>(dg) = zp;
          *reinterpret_cast<uint32_t*>(dg + H) = rp;
          *reinterpret_cast<uint32_t*>(dg + 2 * H) = dahp[mt][hf];
        }
      }
    ptx::fence_proxy_async();
    __syncthreads();                                    // own [da_z | da_r] block complete; every warp is done reading the da_h tile
    if (threadIdx.x == 0) push_block(my_2, GC_BLK2, ptx::smem_u32(&bar_2), rank);
    ptx::mbar_wait(ptx::smem_u32(&bar_2), par);
    // ---- dh += [da_z | da_r] U_zr^T
    float a2s[1][MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) a2s[0][mt][i] = 0.f;
    mma_rows<GC_CS, 1, MT>(a2s, sm + GB_T2, GC_BLK2, GC_BLK2_LD, 0, wzr_rows, GC_W2_LD, 0, 2 * GC_HU, lane);             // the z halves of the four blocks
    mma_rows<GC_CS, 1, MT>(a2s, sm + GB_T2, GC_BLK2, GC_BLK2_LD, GC_HU, wzr_rows, GC_W2_LD, GC_HU, 2 * GC_HU, lane);     // the r halves
    const float (&a2)[MT][4] = a2s[0];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) { dh[mt][hf][0] += a2[mt][2 * hf]; dh[mt][hf][1] += a2[mt][2 * hf + 1]; }
  }
  if (p.dS_h) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int m = row0 + 16 * mt + 8 * hf + g;
        if (m < n) *reinterpret_cast<uint32_t*>(p.dS_h + (size_t)m * p.ldS + gu0) = pack2(dh[mt][hf][0], dh[mt][hf][1]);
      }
  }
  __syncthreads();
  ptx::cluster_arrive();
  ptx::cluster_wait();
}

template <typename K>
void launch(K kern, bool& configured, const GruP& p, int rows, size_t smem, cudaStream_t st) {
  if (!configured) {
    MVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int clusters = (p.n + rows - 1) / rows;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(clusters * GC_CS)); cfg.blockDim = dim3(GC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = GC_CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  MVAE_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  count_launch();
}

// 16 rows per cluster while all clusters are co-resident in one wave (the step is a latency chain: more, thinner clusters shorten it), else 32
int rows_per_cluster(int n) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("MVAE_GRU_ROWS"); forced = e ? atoi(e) : 0; }
  if (forced == 16 || forced == 32) return forced;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return ((n + 15) / 16) * GC_CS <= sms ? 16 : 32;
}

}  // namespace

bool gru_cluster_supported(int H) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("MVAE_GRU_CLUSTER"); enabled = e ? atoi(e) : 1; }
  return enabled && H == GC_H;
}

void gru_cluster_forward(const GruClusterArgs& a, cudaStream_t st) {
  MVAE_REQUIRE(a.H == GC_H && a.ldu % 8 == 0, "cluster GRU recurrence: hidden size 256");
  GruP p{};
  p.n = a.n; p.t0 = a.t0; p.t1 = a.t1; p.gate_act = a.gate_act; p.mix = a.mix;
  p.U = (const bf16*)a.U; p.ldu = a.ldu; p.xw = (const bf16*)a.xw; p.hseq = (bf16*)a.hseq; p.gates = (bf16*)a.gates; p.rh = (bf16*)a.rh;
  static bool configured1 = false, configured2 = false;
  if (rows_per_cluster(a.n) == 16) launch(gru_cluster_fwd_kernel<1>, configured1, p, 16, gf_total(1) + 128, st);
  else launch(gru_cluster_fwd_kernel<2>, configured2, p, 32, gf_total(2) + 128, st);
}

void gru_cluster_backward(const GruClusterArgs& a, cudaStream_t st) {
  MVAE_REQUIRE(a.H == GC_H && a.ldu % 8 == 0, "cluster GRU recurrence: hidden size 256");
  GruP p{};
  p.n = a.n; p.t0 = a.t0; p.t1 = a.t1; p.gate_act = a.gate_act; p.mix = a.mix;
  p.U = (const bf16*)a.U; p.ldu = a.ldu; p.hseq = (bf16*)a.hseq; p.gates = (bf16*)a.gates;
  p.dhext = (const bf16*)a.dhext; p.dh_last = (const bf16*)a.dh_last; p.ld_last = a.ld_last; p.dG = (bf16*)a.dG; p.dS_h = (bf16*)a.dS_h; p.ldS = a.ldS;
  static bool configured1 = false, configured2 = false;
  if (rows_per_cluster(a.n) == 16) launch(gru_cluster_bwd_kernel<1>, configured1, p, 16, gb_total(1) + 128, st);
  else launch(gru_cluster_bwd_kernel<2>, configured2, p, 32, gb_total(2) + 128, st);
}

}  // namespace mvae
