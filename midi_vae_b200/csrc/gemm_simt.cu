// gemm_simt.cu -- CUDA-core GEMM with exact fp32 accumulation.
//
// Role: (1) the fp32 "parity" precision of the MIDI-VAE path (every Dense / LSTM projection the
// reference runs through Theano gemm, vae_definition.py:455-507,533-643), where results must match
// the CPU oracle to 1e-4; (2) the checker inside mvae_selftest_gemm for the tcgen05 kernel; (3) GEMMs
// too small or too oddly shaped for a 128-row tensor-core tile (M <= 16, e.g. bias-like reductions).
#include "common.cuh"

namespace mvae {

thread_local long long g_launches = 0;

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

template <typename TI, typename TO, typename TADD>
__global__ void __launch_bounds__(256) gemm_simt_kernel(int M, int N, int K, const TI* __restrict__ A, long sam, long sak,
                                                        const TI* __restrict__ B, long sbk, long sbn, TO* __restrict__ C, int ldc,
                                                        const float* __restrict__ bias, const TADD* __restrict__ addend, int ldadd,
                                                        int act, int accumulate) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const bool a_kfast = (sak == 1), b_nfast = (sbn == 1);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < (BM * BK) / 256; ++i) {
      int e = tid + i * 256;
      int k, m;
      if (a_kfast) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
      int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < K) ? ldf<TI>(A + (long)gm * sam + (long)gk * sak) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < (BN * BK) / 256; ++i) {
      int e = tid + i * 256;
      int k, n;
      if (b_nfast) { n = e % BN; k = e / BN; } else { k = e % BK; n = e / BK; }
      int gn = n0 + n, gk = k0 + k;
      Bs[k][n] = (gn < N && gk < K) ? ldf<TI>(B + (long)gk * sbk + (long)gn * sbn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int gm = m0 + ty * TM + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int gn = n0 + tx * TN + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[gn];
      if (addend) v += ldf<TADD>(addend + (long)gm * ldadd + gn);
      if (act == 1) v = tanhf(v);
      TO* c = C + (long)gm * ldc + gn;
      if (accumulate) v += ldf<TO>(c);
      stf<TO>(c, v);
    }
  }
}

template <typename TI, typename TO, typename TADD>
void launch(const GemmArgs& g, cudaStream_t st) {
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM);
  long sam = g.transA ? 1 : g.lda, sak = g.transA ? g.lda : 1;
  long sbk = g.transB ? 1 : g.ldb, sbn = g.transB ? g.ldb : 1;
  gemm_simt_kernel<TI, TO, TADD><<<grid, 256, 0, st>>>(g.M, g.N, g.K, (const TI*)g.A, sam, sak, (const TI*)g.B, sbk, sbn, (TO*)g.C,
                                                        g.ldc, g.bias, (const TADD*)g.addend, g.ldadd, g.act, g.accumulate ? 1 : 0);
  count_launch();
  MVAE_CUDA(cudaGetLastError());
}

}  // namespace

void gemm_simt(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return;
  MVAE_REQUIRE(g.K > 0, "gemm K must be positive");
  const bool ib = g.in_type == DT_BF16, cb = g.c_type == DT_BF16, ab = g.addend && g.add_type == DT_BF16;
  using bf = __nv_bfloat16;
  if (!ib && !cb && !ab) launch<float, float, float>(g, st);
  else if (!ib && !cb && ab) launch<float, float, bf>(g, st);
  else if (ib && !cb && !ab) launch<bf, float, float>(g, st);
  else if (ib && !cb && ab) launch<bf, float, bf>(g, st);
  else if (ib && cb && !ab) launch<bf, bf, float>(g, st);
  else if (ib && cb && ab) launch<bf, bf, bf>(g, st);
  else if (!ib && cb && !ab) launch<float, bf, float>(g, st);
  else launch<float, bf, bf>(g, st);
}

}  // namespace mvae
