// common.cuh -- shared declarations of libmidivae.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdexcept>
#include <string>

namespace mvae {

enum DT : int { DT_F32 = 0, DT_BF16 = 1 };
static inline size_t dt_size(DT d) { return d == DT_F32 ? 4 : 2; }

struct Error : std::runtime_error {
  explicit Error(const std::string& s) : std::runtime_error(s) {}
};

#define MVAE_CUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      char _b[512];                                                                             \
      snprintf(_b, sizeof(_b), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      (void)cudaGetLastError(); /* clear the sticky last-error so that later launch checks do not re-report it */ \
      throw mvae::Error(_b);                                                                    \
    }                                                                                           \
  } while (0)

#define MVAE_REQUIRE(cond, msg)                                                                 \
  do {                                                                                          \
    if (!(cond)) {                                                                              \
      char _b[512];                                                                             \
      snprintf(_b, sizeof(_b), "%s:%d: requirement failed: %s (%s)", __FILE__, __LINE__, #cond, msg); \
      throw mvae::Error(_b);                                                                    \
    }                                                                                           \
  } while (0)

// typed load/store through float
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---------------------------------------------------------------------------------------------
// GEMM front door:  C[M,N] = act( op(A)[M,K] * op(B)[K,N] + bias[n] + addend[m,n] )  (+ C if accumulate)
// row-major storage; transA: A is stored [K,M]; transB: B is stored [N,K].
// ---------------------------------------------------------------------------------------------
struct GemmArgs {
  int M = 0, N = 0, K = 0;
  const void* A = nullptr; int lda = 0; bool transA = false;
  const void* B = nullptr; int ldb = 0; bool transB = false;
  DT in_type = DT_F32;
  void* C = nullptr; int ldc = 0; DT c_type = DT_F32;
  const float* bias = nullptr;
  const void* addend = nullptr; int ldadd = 0; DT add_type = DT_F32;
  int act = 0;               // 0 = identity, 1 = tanh
  bool accumulate = false;   // C += result (fp32 C only)
  // optional second product sharing B and K (accumulate GEMMs on the tcgen05 path only): C2[M2,N] += op(A2) op(B), op(A2) laid out like op(A).
  // One launch, K-split-major tile order: the two weight gradients of a layer (dU = Hprev^T dG, dW = X^T dG) read dG from DRAM once.
  const void* A2 = nullptr; int lda2 = 0; int M2 = 0; void* C2 = nullptr; int ldc2 = 0;
};

// SIMT fp32-math GEMM (inputs fp32 or bf16).  Always available; exact fp32 accumulation.
void gemm_simt(const GemmArgs& g, cudaStream_t st);
// tcgen05 / TMA / TMEM GEMM (bf16 inputs, fp32 accumulate).  Throws if the shape is unsupported.
bool gemm_tc_supported(const GemmArgs& g);
void gemm_tc(const GemmArgs& g, cudaStream_t st, int sm_count, int* sched = nullptr);   // sched: 2 zeroed ints owned by the calling stream (dynamic tile scheduler)
int gemm_tc_selftest(int device, int verbose);

// launch accounting (bench.py reports gpu_launches)
extern thread_local long long g_launches;
static inline void count_launch(int n = 1) { g_launches += n; }

}  // namespace mvae
