// gemm_tc.cu -- bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA into 128B-swizzled shared memory, persistent over output tiles.
//
// This is the workhorse for every batchable matrix product of the MIDI-VAE step (input projections of
// the stacked layers, output Denses, all weight gradients, dX = dG W^T; SURVEY.md section 2 K1,K3,K5,K8),
// i.e. what the reference leaves to Theano's gemm under Keras Dense/LSTM (vae_definition.py:455-507,533-643).
//
//   C[M,N] = act( op(A) op(B) + bias + addend )     or     C += op(A) op(B)   (fp32 red.add, split-K)
//
// Layout cases (all row-major in global memory, no transposed copies are ever made):
//   A "K-major"  : stored [M,K]   (transA = false)     A "MN-major": stored [K,M]   (transA = true, weight grads)
//   B "K-major"  : stored [N,K]   (transB = true)      B "MN-major": stored [K,N]   (transB = false, plain weights)
// handled by the UMMA shared-memory descriptors (major-ness bits in the instruction descriptor), with the
// TMA boxes chosen so that the smem image is the canonical SWIZZLE_128B layout of the respective major-ness.
//
// CTA = 10 warps: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM allocator), warps 2..9 epilogue
// (TMEM -> registers -> global, one output row per thread, two warps per lane quadrant).  Two TMEM accumulators (2 x BN columns) let
// the epilogue of tile i overlap the MMAs of tile i+1.
//
// PAIR form (template parameter): the kernel runs as 2-CTA clusters and one `tcgen05.mma.cta_group::2` (M 256, N 256, K 16) works on a 256 x 256
// output tile: each CTA stages ITS 128 rows of A and ITS 128 columns of B (32 KB per k-block instead of 48 KB for a 128 x 256 single-CTA tile:
// 128 instead of 87 FLOP per operand byte), both CTAs' TMA loads complete transaction bytes on the LEADER's full barrier (`cp.async.bulk.tensor
// ...cta_group::2`), the leader issues the MMAs and its commits arrive on both CTAs' barriers (multicast), each CTA drains the 128 accumulator rows
// its own tensor memory holds.  The leader's producer draws the tile and hands it to the peer through distributed shared memory.
#include <cuda.h>

#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace mvae {
namespace {

using bf16 = __nv_bfloat16;

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int kEpiWarps = 8;             // two per TMEM lane quadrant: the epilogue of a K = 512 tile is as long as its MMAs
constexpr int kThreads = 32 * (2 + kEpiWarps);

struct TcParams {
  int M, N, K;
  void* C; int ldc; int c_bf16;
  const float* bias;
  const void* addend; int ldadd; int add_bf16;
  int act; int atomic_acc;
  int m_tiles, n_tiles, k_splits, kb_total, kb_per_split, ks_major, n_fast;
  int m_tiles1, M2; void* C2; int ldc2;   // m tiles >= m_tiles1 belong to the second product (A2 -> C2); m_tiles1 == m_tiles: none
  int red_v4;   // split-K accumulation with 16-byte vector reductions (red.global.add.v4.f32)
  int* sched;   // [0] next tile, [1] finished CTAs (both zero between launches that share them): dynamic tile scheduler
};

// BN = columns of B one CTA stages (PAIR: half of the tile's 2 BN columns); A_BYTES / B_BYTES / STAGE_BYTES are per CTA
template <int BN, bool PAIR>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;   // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = PAIR ? 6 : (BN == 256) ? 4 : 5;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024;   // + alignment slack
};

template <bool A_MN, bool B_MN, int BN, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const __grid_constant__ CUtensorMap tma_a2, const TcParams p) {
  using L = SmemLayout<BN, PAIR>;
  constexpr int STAGES = L::STAGES;
  constexpr int BMT = PAIR ? 2 * BM : BM, BNT = PAIR ? 2 * BN : BN;   // output tile of the CTA / the CTA pair; a CTA's accumulator = BM lanes x BNT columns
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2];
  // tiles are handed out by an atomic counter (a CTA that becomes resident late -- other kernels may hold part of the chip -- simply finds
  // nothing left); the producer warp (PAIR: of the leader CTA) draws the tile and passes it to the MMA and epilogue warps (PAIR: and to the peer
  // CTA's producer and epilogue warps, through distributed shared memory) through this ring
  constexpr int RS = 4;
  __shared__ __align__(8) uint64_t sched_full[RS], sched_empty[RS];
  __shared__ int sched_tile[RS];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B needs 1024-byte alignment
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles * p.k_splits;
  const uint32_t crank = PAIR ? ptx::cluster_ctarank() : 0u;               // PAIR: CTA 0 of the cluster is the leader
  const bool leader = crank == 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tma_a);
    ptx::prefetch_tmap(&tma_b);
    ptx::prefetch_tmap(&tma_a2);
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1); ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(ptx::smem_u32(&tmem_full_bar[a]), 1); ptx::mbar_init(ptx::smem_u32(&tmem_empty_bar[a]), (PAIR ? 2 : 1) * kEpiWarps); }
    for (int r = 0; r < RS; ++r) { ptx::mbar_init(ptx::smem_u32(&sched_full[r]), 1); ptx::mbar_init(ptx::smem_u32(&sched_empty[r]), (PAIR ? 2 : 1) * (1 + kEpiWarps)); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) { ptx::tmem_alloc2(ptx::smem_u32(&tmem_base_slot), 2 * BNT); ptx::tmem_relinquish2(); }
    else { ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), 2 * BNT); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (PAIR) { ptx::cluster_arrive(); ptx::cluster_wait(); }               // the peer's barriers exist before anything arrives on them
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  // one consumer of the tile ring (a whole warp calls; lane 0 does the barrier work): returns the tile of ring position `it`
  auto next_tile = [&](int it) -> int {
    const int rs = it % RS;
    int tile = 0;
    if (lane == 0) {
      if (PAIR && !leader) ptx::mbar_wait_cluster(ptx::smem_u32(&sched_full[rs]), (uint32_t)((it / RS) & 1));   // completed by the leader's remote arrive
      else ptx::mbar_wait(ptx::smem_u32(&sched_full[rs]), (uint32_t)((it / RS) & 1));
      tile = *reinterpret_cast<volatile int*>(&sched_tile[rs]);
      if (PAIR) ptx::mbar_arrive_remote(ptx::mapa(ptx::smem_u32(&sched_empty[rs]), 0));
      else ptx::mbar_arrive(ptx::smem_u32(&sched_empty[rs]));
    }
    return __shfl_sync(0xffffffffu, tile, 0);
  };

  if (warp == 0) {
    // ===================== TMA producer (PAIR: one per CTA, each stages its own halves) =====================
    int stage = 0; uint32_t phase = 0;
    for (int sit = 0;; ++sit) {
      int tile = 0;
      if (leader) {
        if (lane == 0) {
          const int rs = sit % RS;
          if (PAIR) ptx::mbar_wait_cluster(ptx::smem_u32(&sched_empty[rs]), (uint32_t)(((sit / RS) & 1) ^ 1));
          else ptx::mbar_wait(ptx::smem_u32(&sched_empty[rs]), (uint32_t)(((sit / RS) & 1) ^ 1));
          tile = atomicAdd(p.sched, 1);
          sched_tile[rs] = tile;
          ptx::mbar_arrive(ptx::smem_u32(&sched_full[rs]));
          if (PAIR) {
            ptx::st_cluster_u32(ptx::mapa(ptx::smem_u32(&sched_tile[rs]), 1), (uint32_t)tile);
            ptx::mbar_arrive_remote(ptx::mapa(ptx::smem_u32(&sched_full[rs]), 1));   // release at cluster scope: orders the store above
          }
        }
        tile = __shfl_sync(0xffffffffu, tile, 0);
      } else {
        tile = next_tile(sit);
      }
      if (tile >= num_tiles) break;
      if (lane != 0) continue;
      // K-split major: tiles that run at the same time share their K range, so an A / B k-block is fetched from DRAM once and hit in L2 by the other
      // output tiles of the wave (MVAE_GEMM_RASTER=0: output-tile major, the round-1 order; measured 13.1-13.3 -> 12.7 ms per cfg3 step)
      const int mnt = p.m_tiles * p.n_tiles;
      const int ks = p.ks_major ? tile / mnt : tile % p.k_splits, mn = p.ks_major ? tile % mnt : tile / p.k_splits;
      const int mt = p.n_fast ? mn / p.n_tiles : mn % p.m_tiles;
      const CUtensorMap* map_a = mt >= p.m_tiles1 ? &tma_a2 : &tma_a;
      const int m0 = (mt >= p.m_tiles1 ? mt - p.m_tiles1 : mt) * BMT + (int)crank * BM;
      const int n0 = (p.n_fast ? mn % p.n_tiles : mn / p.m_tiles) * BNT + (int)crank * BN;
      const int kb0 = ks * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1);
        // PAIR: both CTAs' loads complete their bytes on the LEADER's barrier, which expects the two stages' worth
        const uint32_t fb = PAIR ? ptx::mapa(ptx::smem_u32(&full_bar[stage]), 0) : ptx::smem_u32(&full_bar[stage]);
        if (leader) ptx::mbar_arrive_expect_tx(ptx::smem_u32(&full_bar[stage]), (PAIR ? 2 : 1) * L::STAGE_BYTES);
        const uint32_t sa = smem_base + stage * L::STAGE_BYTES, sb = sa + L::A_BYTES;
        const int k0 = kb * BK;
        if (!A_MN) {
          ptx::tma_load_2d_g<PAIR>(sa, map_a, fb, k0, m0);                        // box {64 K, 128 M}
        } else {
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) ptx::tma_load_2d_g<PAIR>(sa + j * 8192, map_a, fb, m0 + j * 64, k0);    // box {64 M, 64 K}
        }
        if (!B_MN) {
          ptx::tma_load_2d_g<PAIR>(sb, &tma_b, fb, k0, n0);                       // box {64 K, BN N}
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) ptx::tma_load_2d_g<PAIR>(sb + j * 8192, &tma_b, fb, n0 + j * 64, k0);   // box {64 N, 64 K}
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (PAIR: the leader's, for both CTAs) =====================
    if (leader) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(BMT, BNT, A_MN, B_MN);
      int stage = 0; uint32_t phase = 0;
      for (int it = 0;; ++it) {
        const int tile = next_tile(it);
        if (tile >= num_tiles) break;
        if (lane != 0) continue;
        const int ks = p.ks_major ? tile / (p.m_tiles * p.n_tiles) : tile % p.k_splits;
        const int kb0 = ks * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int acc = it & 1; const uint32_t acc_phase = (it >> 1) & 1;
        if (PAIR) ptx::mbar_wait_cluster(ptx::smem_u32(&tmem_empty_bar[acc]), acc_phase ^ 1);   // both CTAs' epilogues have drained this accumulator
        else ptx::mbar_wait(ptx::smem_u32(&tmem_empty_bar[acc]), acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BNT;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * L::STAGE_BYTES, sb = sa + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: 8-row groups 1024 B apart (SBO), advance 32 B per K=16 slice inside the 128 B swizzle row.
            // MN-major: 64-element atoms along M/N 8192 B apart (LBO), 8-row K groups 1024 B apart (SBO), 2048 B per K=16 slice.
            // PAIR: the same descriptors address each CTA's own stage (its 128 rows of A, its BN columns of B)
            const uint64_t da = A_MN ? ptx::umma_desc_sw128(sa + k * 2048, 8192, 1024) : ptx::umma_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? ptx::umma_desc_sw128(sb + k * 2048, 8192, 1024) : ptx::umma_desc_sw128(sb + k * 32, 16, 1024);
            if (PAIR) ptx::umma_bf16_2cta(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else ptx::umma_bf16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // frees the smem stage (PAIR: of both CTAs) once these MMAs retire
          if (PAIR) ptx::umma_commit_2cta(ptx::smem_u32(&empty_bar[stage]), (uint16_t)3);
          else ptx::umma_commit(ptx::smem_u32(&empty_bar[stage]));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (PAIR: of both CTAs)
        if (PAIR) ptx::umma_commit_2cta(ptx::smem_u32(&tmem_full_bar[acc]), (uint16_t)3);
        else ptx::umma_commit(ptx::smem_u32(&tmem_full_bar[acc]));
      }
    }
  } else {
    // ===================== epilogue: warps 2..9, TMEM lane quadrant = warp % 4, 32-column chunks c with c % 2 == ehalf =====================
    const int quad = warp & 3, ehalf = (warp - 2) >> 2;
    for (int it = 0;; ++it) {
      const int tile = next_tile(it);
      if (tile >= num_tiles) break;
      const int mn = p.ks_major ? tile % (p.m_tiles * p.n_tiles) : tile / p.k_splits;
      const int mt = p.n_fast ? mn / p.n_tiles : mn % p.m_tiles;
      const bool second = mt >= p.m_tiles1;
      const int m0 = (second ? mt - p.m_tiles1 : mt) * BMT + (int)crank * BM, n0 = (p.n_fast ? mn % p.n_tiles : mn / p.m_tiles) * BNT;
      const int Mlim = second ? p.M2 : p.M;
      const int acc = it & 1; const uint32_t acc_phase = (it >> 1) & 1;
      ptx::mbar_wait(ptx::smem_u32(&tmem_full_bar[acc]), acc_phase);
      ptx::tc_fence_after();
      const int m = m0 + quad * 32 + lane;
      const bool row_ok = m < Mlim;
#pragma unroll 1
      for (int c = ehalf; c < BNT / 32; c += 2) {
        const int nb = n0 + c * 32;
        if (nb >= p.N) break;                                   // warp-uniform
        float v[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BNT + c * 32), v);
        if (!row_ok) continue;
        const int nvalid = min(32, p.N - nb);
        if (p.bias) {
          if (nvalid == 32 && (reinterpret_cast<uintptr_t>(p.bias + nb) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + nb) + j);
              v[4 * j] += bv.x; v[4 * j + 1] += bv.y; v[4 * j + 2] += bv.z; v[4 * j + 3] += bv.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < nvalid) v[j] += __ldg(p.bias + nb + j);
          }
        }
        if (p.addend) {
          if (p.add_bf16) {
            const bf16* ad = (const bf16*)p.addend + (size_t)m * p.ldadd + nb;
            if (nvalid == 32 && (reinterpret_cast<uintptr_t>(ad) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 u = __ldg(reinterpret_cast<const uint4*>(ad) + j);
                const bf16* e = reinterpret_cast<const bf16*>(&u);
#pragma unroll
                for (int t = 0; t < 8; ++t) v[j * 8 + t] += __bfloat162float(e[t]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (j < nvalid) v[j] += __bfloat162float(ad[j]);
            }
          } else {
            const float* ad = (const float*)p.addend + (size_t)m * p.ldadd + nb;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < nvalid) v[j] += ad[j];
          }
        }
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
        }
        if (p.atomic_acc) {
          float* cp = second ? (float*)p.C2 + (size_t)m * p.ldc2 + nb : (float*)p.C + (size_t)m * p.ldc + nb;
          if (p.red_v4 && nvalid == 32 && (reinterpret_cast<uintptr_t>(cp) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) ptx::red_add_v4(cp + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);   // one L2 reduction per 16 bytes
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < nvalid) atomicAdd(cp + j, v[j]);
          }
        } else if (p.c_bf16) {
          bf16* cp = (bf16*)p.C + (size_t)m * p.ldc + nb;
          if (nvalid == 32 && (reinterpret_cast<uintptr_t>(cp) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 u;
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
              for (int t = 0; t < 4; ++t) h2[t] = __floats2bfloat162_rn(v[j * 8 + 2 * t], v[j * 8 + 2 * t + 1]);
              reinterpret_cast<uint4*>(cp)[j] = u;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < nvalid) cp[j] = __float2bfloat16_rn(v[j]);
          }
        } else {
          float* cp = (float*)p.C + (size_t)m * p.ldc + nb;
          if (nvalid == 32 && (reinterpret_cast<uintptr_t>(cp) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(cp)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < nvalid) cp[j] = v[j];
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) ptx::mbar_arrive_remote(ptx::mapa(ptx::smem_u32(&tmem_empty_bar[acc]), 0));   // the leader's MMA issuer waits for both CTAs
        else ptx::mbar_arrive(ptx::smem_u32(&tmem_empty_bar[acc]));
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (PAIR) { ptx::cluster_arrive(); ptx::cluster_wait(); }   // no CTA leaves (or frees tensor memory) while the peer's MMAs / commits can still touch it
  if (warp == 1) {
    if (PAIR) ptx::tmem_dealloc2(tmem_base, 2 * BNT);
    else ptx::tmem_dealloc(tmem_base, 2 * BNT);
  }
  if (threadIdx.x == 0) {   // the last CTA to leave re-arms the scheduler for the next launch that uses it (launches sharing it are stream-ordered)
    __threadfence();
    if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) { p.sched[0] = 0; p.sched[1] = 0; __threadfence(); }
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MVAE_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    MVAE_REQUIRE(p != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
    fn = (EncodeFn)p;
  }
  return fn;
}

// 2-D bf16 tensor map over a row-major matrix: `inner` contiguous elements per row, `outer` rows, row stride ld elements
CUtensorMap make_map(const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstr[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[256];
    snprintf(b, sizeof(b), "cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu ld=%llu box=%ux%u ptr=%p", (int)r, (unsigned long long)inner,
             (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer, ptr);
    throw Error(b);
  }
  return m;
}

int g_pair_mode = -1;   // -1: by shape (gemm_tc), 0: never, 1: always (self test)

template <bool A_MN, bool B_MN, int BN, bool PAIR>
void launch(const GemmArgs& g, cudaStream_t st, int sm_count, int* sched) {
  using L = SmemLayout<BN, PAIR>;
  constexpr int BMT = PAIR ? 2 * BM : BM, BNT = PAIR ? 2 * BN : BN;
  TcParams p;
  p.M = g.M; p.N = g.N; p.K = g.K; p.C = g.C; p.ldc = g.ldc; p.c_bf16 = g.c_type == DT_BF16;
  p.bias = g.bias; p.addend = g.addend; p.ldadd = g.ldadd; p.add_bf16 = g.add_type == DT_BF16; p.act = g.act;
  p.m_tiles1 = (g.M + BMT - 1) / BMT; p.m_tiles = p.m_tiles1 + (g.A2 ? (g.M2 + BMT - 1) / BMT : 0);
  p.M2 = g.M2; p.C2 = g.C2; p.ldc2 = g.ldc2;
  p.n_tiles = (g.N + BNT - 1) / BNT; p.kb_total = (g.K + BK - 1) / BK;
  const int workers = PAIR ? std::max(1, sm_count / 2) : sm_count;   // CTAs / CTA pairs the launch may occupy
  int splits = 1;
  if (g.accumulate) {   // split K until the grid fills the chip, keeping >= 8 k-blocks per split
    const int tiles = p.m_tiles * p.n_tiles;
    splits = std::max(1, std::min((2 * workers + tiles - 1) / tiles, p.kb_total / 8));
  }
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.k_splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.atomic_acc = g.accumulate ? 1 : 0;
  { static int raster = -1; if (raster < 0) { const char* e = getenv("MVAE_GEMM_RASTER"); raster = e ? atoi(e) : 2; } p.ks_major = raster;
    // un-split GEMMs (tall activations x a small weight matrix): all N tiles of an M tile back to back, so the activation rows are read from DRAM once
    // (MVAE_GEMM_RASTER=1: M fastest; 12.7-12.9 -> 12.5-12.7 ms per cfg3 step)
    p.n_fast = raster >= 2 && p.k_splits == 1; }
  { static int rv4 = -1; if (rv4 < 0) { const char* e = getenv("MVAE_GEMM_REDV4"); rv4 = e ? atoi(e) : 1; } p.red_v4 = rv4; }
  p.sched = sched;
  const CUtensorMap ma = A_MN ? make_map(g.A, g.M, g.K, g.lda, 64, 64) : make_map(g.A, g.K, g.M, g.lda, 64, BM);
  const CUtensorMap ma2 = !g.A2 ? ma : (A_MN ? make_map(g.A2, g.M2, g.K, g.lda2, 64, 64) : make_map(g.A2, g.K, g.M2, g.lda2, 64, BM));
  const CUtensorMap mb = B_MN ? make_map(g.B, g.N, g.K, g.ldb, 64, 64) : make_map(g.B, g.K, g.N, g.ldb, 64, BN);
  auto kern = gemm_tc_kernel<A_MN, B_MN, BN, PAIR>;
  static bool attr_set = false;
  if (!attr_set) {
    MVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  const int tiles = p.m_tiles * p.n_tiles * p.k_splits;
  const int grid = std::min(tiles, workers) * (PAIR ? 2 : 1);
  if (PAIR) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = L::TOTAL; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    MVAE_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, ma2, p));
  } else {
    kern<<<grid, kThreads, L::TOTAL, st>>>(ma, mb, ma2, p);
  }
  count_launch();
  MVAE_CUDA(cudaGetLastError());
}

template <int BN, bool PAIR>
void launch_major(const GemmArgs& g, cudaStream_t st, int sm_count, int* sched) {
  const bool a_mn = g.transA, b_mn = !g.transB;
  if (!a_mn && !b_mn) launch<false, false, BN, PAIR>(g, st, sm_count, sched);
  else if (!a_mn && b_mn) launch<false, true, BN, PAIR>(g, st, sm_count, sched);
  else if (a_mn && !b_mn) launch<true, false, BN, PAIR>(g, st, sm_count, sched);
  else launch<true, true, BN, PAIR>(g, st, sm_count, sched);
}

}  // namespace

bool gemm_tc_supported(const GemmArgs& g) {
  if (g.in_type != DT_BF16) return false;
  if (g.A2 && (!g.accumulate || g.M2 <= 0 || !g.C2 || (reinterpret_cast<uintptr_t>(g.A2) & 15) || (g.lda2 % 8))) return false;
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return false;
  if ((reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.B) & 15)) return false;
  if ((g.lda % 8) || (g.ldb % 8)) return false;
  if (g.accumulate && g.c_type != DT_F32) return false;
  if (g.accumulate && (g.bias || g.addend || g.act)) return false;
  return true;
}

// scheduler words of launches that are NOT given their own: one pair per (device, stream) would be needed for concurrent launches, so callers
// that run GEMMs on several streams pass their own (Model does); this default pair serves single-stream users such as the self test
static int* default_sched() {
  static int* d = nullptr;
  if (!d) { MVAE_CUDA(cudaMalloc(&d, 2 * sizeof(int))); MVAE_CUDA(cudaMemset(d, 0, 2 * sizeof(int))); }
  return d;
}

void gemm_tc(const GemmArgs& g, cudaStream_t st, int sm_count, int* sched) {
  MVAE_REQUIRE(gemm_tc_supported(g), "shape / alignment not supported by the tcgen05 GEMM");
  if (!sched) sched = default_sched();
  // 256 x 256 tiles on CTA pairs (128 FLOP per operand byte and SM) for the split-K weight gradients when both output dimensions fill them
  // (MVAE_GEMM_PAIR=0: single-CTA tiles only; 2: un-split GEMMs too -- measured: the K = 512 projection is bound by its epilogue and gets 6 % slower)
  static int pair_env = -1;
  if (pair_env < 0) { const char* e = getenv("MVAE_GEMM_PAIR"); pair_env = e ? atoi(e) : 1; }
  auto fills = [](int x) { return x >= 256 && (x % 256 == 0 || x >= 1024); };
  const bool pair = g_pair_mode >= 0 ? g_pair_mode != 0 : (pair_env && (g.accumulate || pair_env >= 2) && sm_count >= 2 && fills(g.M) && fills(g.N) && (!g.A2 || fills(g.M2)));
  if (pair) { launch_major<128, true>(g, st, sm_count, sched); return; }
  // 128x256 tiles (87 FLOP per operand byte instead of 64) when N is wide enough to fill them
  static int wide = -1;
  if (wide < 0) { const char* e = getenv("MVAE_GEMM_BN256"); wide = e ? atoi(e) : 1; }
  const bool bn256 = wide && g.N >= 256 && (g.N % 256 == 0 || g.N >= 1024);
  if (bn256) launch_major<256, false>(g, st, sm_count, sched);
  else launch_major<128, false>(g, st, sm_count, sched);
}

// ------------------------------------------------------------------------------------------------ self test
// Random bf16 operands, every layout case, ragged edges, every epilogue; the checker is the SIMT GEMM.
int gemm_tc_selftest(int device, int verbose) {
  MVAE_CUDA(cudaSetDevice(device));
  int sm = 0;
  MVAE_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device));
  cudaStream_t st;
  MVAE_CUDA(cudaStreamCreate(&st));
  struct Case { int M, N, K; bool ta, tb; int epi; };   // epi: 0 plain f32, 1 bias+tanh bf16 out, 2 addend(bf16) f32 out, 3 accumulate (split-K)
  std::vector<Case> cases;
  const int shapes[][3] = {{128, 128, 64}, {128, 128, 256}, {256, 384, 512}, {200, 61, 96}, {77, 130, 72}, {512, 2048, 512}, {61, 256, 4096}, {8, 256, 64}, {300, 16, 128}, {384, 1024, 128}, {200, 1096, 64}};
  for (auto& s : shapes)
    for (int ta = 0; ta < 2; ++ta)
      for (int tb = 0; tb < 2; ++tb)
        for (int epi = 0; epi < 4; ++epi) cases.push_back({s[0], s[1], s[2], ta != 0, tb != 0, epi});
  int failures = 0;
  uint32_t seed = 12345u;
  auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
  // every case through the single-CTA tiles (pm = 0) and through the CTA-pair form (pm = 1; also on shapes gemm_tc would not pick it for:
  // M <= 128 leaves the peer CTA entirely out of bounds)
  for (int pm = 0; pm < 2; ++pm)
  for (const Case& c : cases) {
    g_pair_mode = pm;
    const int lda = ((c.ta ? c.M : c.K) + 7) / 8 * 8, ldb = ((c.tb ? c.K : c.N) + 7) / 8 * 8, ldc = (c.N + 7) / 8 * 8;
    const size_t na = (size_t)(c.ta ? c.K : c.M) * lda, nb = (size_t)(c.tb ? c.N : c.K) * ldb, nc = (size_t)c.M * ldc;
    std::vector<bf16> ha(na), hb(nb), hadd(nc);
    std::vector<float> hbias(c.N), hc0(nc);
    for (auto& x : ha) x = __float2bfloat16(rnd());
    for (auto& x : hb) x = __float2bfloat16(rnd());
    for (auto& x : hadd) x = __float2bfloat16(rnd());
    for (auto& x : hbias) x = rnd();
    for (auto& x : hc0) x = rnd();
    bf16 *dA, *dB, *dAdd; float *dBias, *dC1, *dC2; bf16 *dCb1, *dCb2;
    MVAE_CUDA(cudaMalloc(&dA, na * 2)); MVAE_CUDA(cudaMalloc(&dB, nb * 2)); MVAE_CUDA(cudaMalloc(&dAdd, nc * 2));
    MVAE_CUDA(cudaMalloc(&dBias, c.N * 4)); MVAE_CUDA(cudaMalloc(&dC1, nc * 4)); MVAE_CUDA(cudaMalloc(&dC2, nc * 4));
    MVAE_CUDA(cudaMalloc(&dCb1, nc * 2)); MVAE_CUDA(cudaMalloc(&dCb2, nc * 2));
    MVAE_CUDA(cudaMemcpy(dA, ha.data(), na * 2, cudaMemcpyHostToDevice)); MVAE_CUDA(cudaMemcpy(dB, hb.data(), nb * 2, cudaMemcpyHostToDevice));
    MVAE_CUDA(cudaMemcpy(dAdd, hadd.data(), nc * 2, cudaMemcpyHostToDevice)); MVAE_CUDA(cudaMemcpy(dBias, hbias.data(), c.N * 4, cudaMemcpyHostToDevice));
    MVAE_CUDA(cudaMemcpy(dC1, hc0.data(), nc * 4, cudaMemcpyHostToDevice)); MVAE_CUDA(cudaMemcpy(dC2, hc0.data(), nc * 4, cudaMemcpyHostToDevice));
    MVAE_CUDA(cudaMemset(dCb1, 0, nc * 2)); MVAE_CUDA(cudaMemset(dCb2, 0, nc * 2));
    GemmArgs g; g.M = c.M; g.N = c.N; g.K = c.K; g.A = dA; g.lda = lda; g.transA = c.ta; g.B = dB; g.ldb = ldb; g.transB = c.tb; g.in_type = DT_BF16;
    g.ldc = ldc;
    if (c.epi == 1) { g.bias = dBias; g.act = 1; g.c_type = DT_BF16; }
    if (c.epi == 2) { g.addend = dAdd; g.ldadd = ldc; g.add_type = DT_BF16; }
    if (c.epi == 3) { g.accumulate = true; }
    GemmArgs g1 = g, g2 = g;
    g1.C = c.epi == 1 ? (void*)dCb1 : (void*)dC1;
    g2.C = c.epi == 1 ? (void*)dCb2 : (void*)dC2;
    gemm_tc(g1, st, sm);
    gemm_simt(g2, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { fprintf(stderr, "selftest: kernel failure %s on %s M%d N%d K%d ta%d tb%d epi%d\n", cudaGetErrorString(e), pm ? "pair" : "single", c.M, c.N, c.K, c.ta, c.tb, c.epi); g_pair_mode = -1; return 100; }
    std::vector<float> r1(nc), r2(nc);
    if (c.epi == 1) {
      std::vector<bf16> t1(nc), t2(nc);
      MVAE_CUDA(cudaMemcpy(t1.data(), dCb1, nc * 2, cudaMemcpyDeviceToHost)); MVAE_CUDA(cudaMemcpy(t2.data(), dCb2, nc * 2, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < nc; ++i) { r1[i] = __bfloat162float(t1[i]); r2[i] = __bfloat162float(t2[i]); }
    } else {
      MVAE_CUDA(cudaMemcpy(r1.data(), dC1, nc * 4, cudaMemcpyDeviceToHost)); MVAE_CUDA(cudaMemcpy(r2.data(), dC2, nc * 4, cudaMemcpyDeviceToHost));
    }
    double max_err = 0, max_ref = 0;
    for (int m = 0; m < c.M; ++m)
      for (int n = 0; n < c.N; ++n) {
        double a = r1[(size_t)m * ldc + n], b = r2[(size_t)m * ldc + n];
        max_err = std::max(max_err, fabs(a - b)); max_ref = std::max(max_ref, fabs(b));
      }
    const double tol = (c.epi == 1 ? 1e-2 : 2e-3) * std::max(1.0, max_ref);
    // independent witness: the same product in float64 on the HOST from the bf16 operands (no kernel of this library involved)
    double max_err_host = 0;
    if ((double)c.M * c.N * c.K <= (pm ? 1e8 : 6e8)) {
      std::vector<float> fa(na), fb(nb);
      for (size_t i = 0; i < na; ++i) fa[i] = __bfloat162float(ha[i]);
      for (size_t i = 0; i < nb; ++i) fb[i] = __bfloat162float(hb[i]);
      for (int m = 0; m < c.M; ++m)
        for (int n = 0; n < c.N; ++n) {
          double acc = 0;
          for (int k = 0; k < c.K; ++k) {
            const float av = c.ta ? fa[(size_t)k * lda + m] : fa[(size_t)m * lda + k];
            const float bv = c.tb ? fb[(size_t)n * ldb + k] : fb[(size_t)k * ldb + n];
            acc += (double)av * bv;
          }
          if (c.epi == 1) acc = tanh(acc + hbias[n]);
          if (c.epi == 2) acc += __bfloat162float(hadd[(size_t)m * ldc + n]);
          if (c.epi == 3) acc += hc0[(size_t)m * ldc + n];
          max_err_host = std::max(max_err_host, fabs(acc - (double)r1[(size_t)m * ldc + n]));
        }
    }
    const bool ok = max_err <= tol && max_err_host <= tol;
    if (verbose || !ok) printf("  host float64 witness: max_err=%.3e\n", max_err_host);
    if (!ok) ++failures;
    if (verbose || !ok)
      printf("gemm_tc selftest %s M=%d N=%d K=%d transA=%d transB=%d epi=%d : max_err=%.3e (ref max %.3e) %s\n", pm ? "pair" : "single", c.M, c.N, c.K,
             (int)c.ta, (int)c.tb, c.epi, max_err, max_ref, ok ? "ok" : "FAIL");
    cudaFree(dA); cudaFree(dB); cudaFree(dAdd); cudaFree(dBias); cudaFree(dC1); cudaFree(dC2); cudaFree(dCb1); cudaFree(dCb2);
  }
  g_pair_mode = -1;
  cudaStreamDestroy(st);
  printf("gemm_tc selftest: %d cases, %d failures\n", 2 * (int)cases.size(), failures);
  return failures;
}

}  // namespace mvae
