// rec_persist.cuh -- host interface of the persistent LSTM recurrence kernels (lstm_persist.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace mvae {

struct RecPersistArgs {
  int n = 0, H = 0, steps = 0;
  int gate_act = 0, variant = 0;
  unsigned* flags = nullptr;          // rec_persist_flag_count(n, steps) counters (zeroed by the launcher)
  // forward
  const void* upack = nullptr;        // packed recurrent weights, rec_persist_pack_u
  const void* xw = nullptr;           // (steps, n, 4H) bf16 pre-activations x W + b
  void* hseq = nullptr;               // (steps+1, n, H) bf16, slab 0 = h0 (read), slabs 1.. written
  void* cseq = nullptr;               // cell-state stash, (steps+1) slabs in the kernels' private granule layout, all written
  const void* c0 = nullptr;           // (n, ldc0) bf16 initial cell state, row-major (null = zeros)
  int ldc0 = 0;
  void* gates = nullptr;              // post-activation gates stash, granule layout (written forward, read backward)
  // backward
  const void* u_shadow = nullptr;     // (H, 4H) bf16 recurrent weights, natural layout
  int ldu = 0;
  const void* dhext = nullptr;        // (steps, n, H) bf16 or null
  const void* dh_last = nullptr;      // (n, ld_last) bf16 or null: extra gradient into the last step's h
  int ld_last = 0;
  void* dG = nullptr;                 // (steps, n, 4H) bf16 written
  void* dS_h = nullptr;               // (n, ldS) bf16 gradient wrt h0 / c0 (or null)
  void* dS_c = nullptr;
  int ldS = 0;
  void* hx = nullptr;                 // forward: h exchange buffer, rec_persist_hx_bytes
  const void* upack_bwd = nullptr;    // K-split backward: packed weights, rec_persist_pack_u_bwd
  void* partial = nullptr;            // K-split backward: exchange buffer, rec_persist_partial_bytes
  void* trace = nullptr;              // optional: 8 steps x 16 clock64 stamps of CTA 0 (profiling aid)
  int no_stash = 0;                   // cluster forward: inference, do not write the gates / c stash
  // cluster forward only: input projection computed inside the kernel instead of streamed from the (steps, n, 4H) xw buffer
  int x_mode = 0;                     // 0: xw buffer; 1: one-hot input = row gather from xtab; 2: scalar input (x w + b)
  const void* xtab = nullptr;         // x_mode 1: (64, 4H) bf16: row i = W[i] + b for the one-hot class i, row 63 = b (zero input)
  const unsigned char* x_idx = nullptr;   // x_mode 1: class index of row m at step t-x_shift at x_idx[m * x_ld + t - x_shift]; null or t < x_shift: zero input
  int x_ld = 0, x_shift = 0;
  const void* x_scalar = nullptr;     // x_mode 2: bf16 scalar input of row m at step t at x_scalar[(t * n + m) * x_ld]
  const float* x_w = nullptr;         // x_mode 2: (4H) input kernel row and bias, fp32
  const float* x_b = nullptr;
};

size_t smem_max_bytes();
int rec_persist_hs(int H);                           // hidden units per CTA (0 = unsupported)
bool rec_persist_supported(int H, int sm_count);
size_t rec_persist_flag_count(int n, int steps);
size_t rec_persist_hx_bytes(int n, int H);
void rec_persist_pack_u(const float* U, int ldu, void* upack, int H, int hs, int variant, cudaStream_t st);   // hs = 0: the single-launch default
int rec_persist_fwd_pair_hs(int H, int n, int sm_count);
void rec_persist_forward_pair(const RecPersistArgs& a, const RecPersistArgs& b, int HS, cudaStream_t st, int sm_count);
bool rec_persist_ksplit_ok(int H);
size_t rec_persist_partial_bytes(int n, int H);
void rec_persist_pack_u_bwd(const float* U, int ldu, void* upack_bwd, int H, int HS, int variant, cudaStream_t st);
int rec_persist_pair_hs(int H, int n, int sm_count);
void rec_persist_backward_pair(const RecPersistArgs& a, const RecPersistArgs* b, int HS, cudaStream_t st, int sm_count);
void rec_persist_forward(const RecPersistArgs& a, cudaStream_t st, int sm_count);
void rec_persist_backward(const RecPersistArgs& a, cudaStream_t st, int sm_count);


// cluster / DSMEM kernels (lstm_cluster.cu): H = 256 or 512; same stash layouts as the kernels above
bool rec_cluster_supported(int H);
size_t rec_cluster_hx_bytes(int n, int H);
size_t rec_cluster_xbuf_bytes(int n, int H);
void rec_cluster_pack_u(const float* U, int ldu, void* upack, int H, int variant, cudaStream_t st);
void rec_cluster_forward(const RecPersistArgs& a, cudaStream_t st);
void rec_cluster_build_xtab(const void* W_bf16, int ldw, int din, const float* bias, void* xtab, int H, cudaStream_t st);
bool rec_cluster_bwd_supported(int H);
void rec_cluster_pack_u_bwd(const float* U, int ldu, void* upack_bwd, int H, int variant, cudaStream_t st);
void rec_cluster_backward(const RecPersistArgs& a, cudaStream_t st);


// cluster-resident GRU recurrence (gru_cluster.cu): H = 256, bf16; buffers in the layouts of the step-streamed GRU path (row-major, act-typed)
struct GruClusterArgs {
  int n = 0, H = 0, t0 = 0, t1 = 0;   // steps [t0, t1)
  int gate_act = 0, mix = 0;          // mix 0: h = z h' + (1 - z) hh (Keras GRU); 1: (1 - z) h' + z hh (recurrentshop GRUCell)
  const void* U = nullptr; int ldu = 0;   // (H, 3H) bf16 recurrent kernel [z | r | h]
  const void* xw = nullptr;           // (T, n, 3H) forward: x W + b
  void* hseq = nullptr;               // (T + 1, n, H): slab t0 read, slabs t0 + 1 .. t1 written (reverse sweep: read)
  void* gates = nullptr;              // (T, n, 3H) [z | r | hh]: written forward, read by the reverse sweep
  void* rh = nullptr;                 // (T, n, H) r * h_{t-1}: written forward (operand of the dU_h weight gradient)
  const void* dhext = nullptr;        // reverse sweep: (T, n, H) or null
  const void* dh_last = nullptr; int ld_last = 0;   // (n, ld_last) or null
  void* dG = nullptr;                 // (T, n, 3H) [da_z | da_r | da_h] written
  void* dS_h = nullptr; int ldS = 0;  // (n, ldS) gradient wrt the initial state, or null
};
bool gru_cluster_supported(int H);
void gru_cluster_forward(const GruClusterArgs& a, cudaStream_t st);
void gru_cluster_backward(const GruClusterArgs& a, cudaStream_t st);

}  // namespace mvae
