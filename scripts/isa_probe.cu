// isa_probe.cu -- checks on the device the two layout facts the stmatrix message path of rec_cluster_bwd4_kernel rests on:
//   (1) register order of tcgen05.ld.16x256b.x4 (accumulator-fragment shape), (2) row/element mapping of stmatrix.m8n8.x4.trans.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -o scripts/isa_probe scripts/isa_probe.cu     Run: scripts/isa_probe
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../midi_vae_b200/csrc/ptx.cuh"

using namespace mvae;

__global__ void probe(int* bad) {
  __shared__ uint32_t slot;
  __shared__ __align__(128) unsigned short tile[4][32][8];   // per warp: [32 batch rows][8 units] bf16 for one unit granule
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&slot), 32); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tb = slot;
  uint32_t r[32];
  for (int c = 0; c < 32; ++c) r[c] = __float_as_uint((float)(1000 * (32 * warp + lane) + c));
  ptx::tmem_st_32x32(tb + ((uint32_t)(32 * warp) << 16), r);
  ptx::tmem_st_wait();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  // (1) fragment read
  for (int hl = 0; hl < 2; ++hl) {
    float v[16];
    ptx::tmem_ld_16x256b_x4(tb + ((uint32_t)(32 * warp + 16 * hl) << 16), v);
    for (int j = 0; j < 4; ++j)
      for (int k = 0; k < 4; ++k) {
        const int l = 32 * warp + 16 * hl + (k >> 1) * 8 + lane / 4, c = 8 * j + 2 * (lane % 4) + (k & 1);
        if (v[4 * j + k] != (float)(1000 * l + c)) atomicAdd(bad, 1);
      }
    // (2) the matrices of unit granule 2 hl + s (8 TMEM lanes) x 32 columns -> tile[warp][row = column][unit]; only granule (hl = 0, s = 0) and
    //     (hl = 1, s = 1) are stored (into the same tile, one after the other) to keep the probe small
    const int s = hl;
    uint32_t q[4];
    for (int cb = 0; cb < 4; ++cb) {
      // small exactly representable values: unit (0..7) * 32 + column (0..31)
      const float f0 = (float)((lane / 4) * 32 + 8 * cb + 2 * (lane % 4)), f1 = f0 + 1.f;
      (void)v;
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(f0, f1);
      q[cb] = *reinterpret_cast<const uint32_t*>(&h2);
    }
    // matrix m = column block m; thread 8 m + j addresses row 8 m + j
    ptx::stmatrix_x4_trans(ptx::smem_u32(&tile[warp][8 * (lane >> 3) + (lane & 7)][0]), q[0], q[1], q[2], q[3]);
    __syncwarp();
    for (int u = 0; u < 8; ++u) {
      const float got = __bfloat162float(__ushort_as_bfloat16(tile[warp][lane][u]));
      if (got != (float)(u * 32 + lane)) atomicAdd(bad + 1, 1);
    }
    __syncwarp();
    (void)s;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tb, 32);
}

int main() {
  int* d = nullptr;
  cudaMalloc(&d, 8);
  cudaMemset(d, 0, 8);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  int h[2] = {-1, -1};
  cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
  printf("isa_probe: %s; tcgen05.ld.16x256b.x4 mismatches %d, stmatrix.x4.trans mismatches %d\n", cudaGetErrorString(e), h[0], h[1]);
  return (e == cudaSuccess && h[0] == 0 && h[1] == 0) ? 0 : 1;
}
