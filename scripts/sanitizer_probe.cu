// sanitizer_probe.cu -- minimal, known-good uses of the two constructs compute-sanitizer reports in the cluster recurrences, to tell tool
// limitations from real findings (profiles/r2/sanitizer/README.md):
//   probe "bar":  three warps of a four-warp block meet at `bar.sync 1, 96` (a named barrier with a thread count below the block size)
//   probe "bulk": a 2-CTA cluster; each CTA pushes 4160 bytes into its peer's dynamic shared memory with
//                 cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes at destination offset argv[2] (bytes), and checks them
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o scripts/sanitizer_probe scripts/sanitizer_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void bar_probe(int iters, int* out) {
  __shared__ int acc[3];
  const int warp = threadIdx.x >> 5;
  if (warp < 3) {
    for (int i = 0; i < iters; ++i) {
      if ((threadIdx.x & 31) == 0) acc[warp] = i;
      asm volatile("bar.sync 1, 96;" ::: "memory");
      if (threadIdx.x == 0) out[0] = acc[0] + acc[1] + acc[2];
      asm volatile("bar.sync 1, 96;" ::: "memory");
    }
  }
}

__global__ void barwait_probe(int iters, int* out) {
  // as bar_probe, but the fourth warp does not exit: it waits at the block-wide barrier at the end (as the idle lanes of the MMA / relay warp
  // do in the recurrence kernels) while the other three keep meeting at the named barrier
  __shared__ int acc[3];
  const int warp = threadIdx.x >> 5;
  if (warp < 3) {
    for (int i = 0; i < iters; ++i) {
      if ((threadIdx.x & 31) == 0) acc[warp] = i;
      asm volatile("bar.sync 1, 96;" ::: "memory");
      if (threadIdx.x == 0) out[0] = acc[0] + acc[1] + acc[2];
      asm volatile("bar.sync 1, 96;" ::: "memory");
    }
  }
  __syncthreads();
}

__global__ void barreg_probe(int iters, int* out) {
  // two independent warp sets (warps 0-1 and 2-3), each meeting at its own named barrier; id and count come from registers, as in the kernels
  __shared__ int acc[4];
  const int warp = threadIdx.x >> 5, set = warp >> 1;
  const int id = 1 + 2 * set, count = 64;
  for (int i = 0; i < iters; ++i) {
    if ((threadIdx.x & 31) == 0) acc[warp] = i;
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
    if ((threadIdx.x & 63) == 0) out[set] = acc[2 * set] + acc[2 * set + 1];
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
  }
}

template <bool STATIC>
__device__ __forceinline__ void bulk_body(uint32_t dst_off, int hop, int* bad);

__global__ void __cluster_dims__(2, 1, 1) bulk_probe(uint32_t dst_off, int* bad) { bulk_body<true>(dst_off, 1, bad); }
// same body, cluster shape given by a launch attribute (cudaLaunchKernelEx), any size; hop = destination rank offset (0 = the CTA itself)
__global__ void bulkx_probe(uint32_t dst_off, int hop, int* bad) { bulk_body<false>(dst_off, hop, bad); }

template <bool STATIC>
__device__ __forceinline__ void bulk_body(uint32_t dst_off, int hop, int* bad) {
  extern __shared__ __align__(128) uint8_t dyn[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t rank, csize;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  const uint32_t dest = (rank + (uint32_t)hop) % csize, src_rank = (rank + csize - (uint32_t)hop % csize) % csize;
  const uint32_t base = (smem_u32(dyn) + 127u) & ~127u;
  constexpr uint32_t BYTES = 4160;
  for (uint32_t i = threadIdx.x; i < BYTES / 4; i += blockDim.x)
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + i * 4), "r"(rank * 100000u + i) : "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BYTES) : "memory");
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    uint32_t rdst, rbar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdst) : "r"(base + dst_off), "r"(dest));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(&bar)), "r"(dest));
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rdst), "r"(base), "r"(BYTES), "r"(rbar)
                 : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < BYTES / 4; i += blockDim.x) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + dst_off + i * 4) : "memory");
    if (v != src_rank * 100000u + i) atomicAdd(bad, 1);
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int main(int argc, char** argv) {
  const char* which = argc > 1 ? argv[1] : "bar";
  int* d = nullptr;
  cudaMalloc(&d, 4);
  cudaMemset(d, 0, 4);
  if (!strcmp(which, "bar")) {
    bar_probe<<<1, 128>>>(4, d);
  } else if (!strcmp(which, "barwait")) {
    barwait_probe<<<1, 128>>>(4, d);
  } else if (!strcmp(which, "barreg")) {
    barreg_probe<<<1, 128>>>(4, d);
  } else if (!strcmp(which, "bulkx")) {
    // argv: bulkx <dst offset> <cluster size> <hop>
    const uint32_t off = argc > 2 ? (uint32_t)atoi(argv[2]) : 8192u;
    const int cs = argc > 3 ? atoi(argv[3]) : 2, hop = argc > 4 ? atoi(argv[4]) : 1;
    const size_t smem = (size_t)off + 4160 + 256;
    cudaFuncSetAttribute(bulkx_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cs > 8) cudaFuncSetAttribute(bulkx_probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, bulkx_probe, off, hop, d);
    if (le != cudaSuccess) printf("launch: %s\n", cudaGetErrorString(le));
  } else {
    const uint32_t off = argc > 2 ? (uint32_t)atoi(argv[2]) : 8192u;
    const size_t smem = (size_t)off + 4160 + 256;
    cudaFuncSetAttribute(bulk_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bulk_probe<<<2, 128, smem>>>(off, d);
  }
  cudaError_t e = cudaDeviceSynchronize();
  int h = -1;
  cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
  printf("probe %s: %s, result word %d\n", which, cudaGetErrorString(e), h);
  return e == cudaSuccess ? 0 : 1;
}
