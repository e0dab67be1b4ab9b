// kernels.cuh -- launch wrappers of the pointwise / reduction kernels of the MIDI-VAE path.
#pragma once
#include "common.cuh"

namespace mvae {

struct CellCfg {
  int gate_act;  // MVAE_GATE_*
  int variant;   // MVAE_CELL_*
};

// metric accumulator slots (double, device)
enum { ACC_CE_NOTES = 0, ACC_ACC_NOTES, ACC_CE_INSTR, ACC_ACC_INSTR, ACC_MSE_VEL, ACC_ACC_VEL, ACC_CE_STYLE, ACC_ACC_STYLE, ACC_KL, ACC_WNZ, ACC_COUNT };

void k_expand_inputs(DT act, int n, int T, int Ti, int PD, int ID, int VD, const uint8_t* pitch, const uint8_t* target,
                     const uint8_t* instr, const float* vel, void* Xp_ext, void* Yp_ext, void* Xi_ext, void* Xv_ext, cudaStream_t st);
void k_fill_rows(DT act, void* dst, long rows, int cols, const float* bias, cudaStream_t st);
void k_copy2d(DT src_t, DT dst_t, int rows, int cols, const void* src, int lds, void* dst, int ldd, cudaStream_t st);
void k_cell_fwd(DT act, CellCfg cc, int n, int H, const float* pre, float* c_run, void* gates_t, void* cseq_t1, void* hseq_t1, cudaStream_t st);
void k_cell_bwd(DT act, CellCfg cc, int n, int H, const float* dh_run, const void* dh_ext_t, const void* dh_last, int ld_last, DT last_t,
                float* dc_run, const void* gates_t, const void* cseq_t, const void* cseq_t1, void* dG_t, cudaStream_t st);
// GRU cell (blocks [z|r|h]; mix 0 = Keras GRU, 1 = recurrentshop GRUCell as recalled): two pointwise launches per step in each direction
void k_gru_gates(DT act, int gate_act, int n, int H, const float* pre, const void* h_prev, void* gates_t, void* rh_t, cudaStream_t st);
void k_gru_out(DT act, int mix, int n, int H, const float* pre, const void* h_prev, void* gates_t, void* h_new, cudaStream_t st);
void k_gru_bwd1(DT act, int gate_act, int mix, int n, int H, float* dh_run, const void* dh_ext_t, const void* dh_last, int ld_last, DT last_t,
                const void* gates_t, const void* h_prev, void* dG_t, cudaStream_t st);
void k_gru_bwd2(DT act, int gate_act, int n, int H, const float* drh, const void* gates_t, const void* h_prev, float* dh_run, void* dG_t, cudaStream_t st);
void k_concat3(DT act, int n, int H, const void* a, const void* b, const void* c, void* u, cudaStream_t st);
void k_latent_fwd(DT act, int n, int L, int ldl, const float* mu, const float* lv, const float* eps, const float* hist, int has_hist,
                  float* z, void* q, int ldq, float beta, float m0, float s0, double* acc, cudaStream_t st);
void k_style_head(int n, int C, const float* z, int ldl, const uint8_t* style, float* probs, double* acc, cudaStream_t st);
void k_latent_bwd(DT act, int n, int L, int ldl, int C, const void* dq, int ldq, const float* mu, const float* lv, const float* eps,
                  const float* style_probs, const uint8_t* style, float beta, float m0, float s0, float style_w, void* dmu, void* dlv,
                  cudaStream_t st);
void k_count_nonzero(const float* w, long count, double* acc, cudaStream_t st);
void k_softmax_ce(DT act, int steps, int n, int D, float* logits, int ld, const uint8_t* labels, const float* w, const double* acc_wnz,
                  float loss_w, void* dlogits, int ldd, double* acc, int slot_ce, int slot_acc, cudaStream_t st);
void k_sigmoid_mse(DT act, int steps, int n, float* logits, int ld, const float* target, float loss_w, void* dlogits, int ldd, double* acc,
                   cudaStream_t st);
void k_tanh_bwd(DT act, long count, const void* dout, const void* out, void* dpre, cudaStream_t st);
// dst[c] += sum_r weight[r] * src[r,c]   (weight == nullptr: plain column sum); weight is act-typed with stride ldw
// wsum (optional, with weight): += sum_r weight[r]
void k_colsum(DT act, long rows, int cols, int ld, const void* src, const void* weight, int ldw, float* dst, cudaStream_t st, float* wsum = nullptr);
// bf16 dG (rows = T n, time-major), ONE pass: db += column sums and dW[0,:] += sum_r x[r] dG[r,:] (scalar-input recurrences).  cols % 8 == 0, ld % 8 == 0.
void k_wgrad_rows(long rows, int cols, int ld, const void* src, const void* x, int ldx, float* dW, float* db, cudaStream_t st);
// out[r,c] = x[r] * w[c] + bias[c]   (x act-typed with stride ldx; w, bias fp32; bias may be null)
void k_rank1_rows(DT act, void* out, long rows, int cols, const void* x, int ldx, const float* w, const float* bias, cudaStream_t st);
// out[r*ldo] = dot(h[r,:H], w) + b[0]
void k_rowdot(DT act, long rows, int H, const void* h, const float* w, const float* b, float* out, int ldo, cudaStream_t st);
void k_adam(long count, float* p, const float* g, float* m, float* v, float lr_t, float b1, float b2, float eps, float gscale,
            __nv_bfloat16* shadow, cudaStream_t st);
void k_f32_to_bf16(long count, const float* src, __nv_bfloat16* dst, cudaStream_t st);
void k_finalize_metrics(const double* acc, int n, int T, int Ti, float w_notes, float w_instr, float w_vel, float w_style, float* out,
                        cudaStream_t st);
void k_export_seq(int steps, int n, int D, const float* probs, int ld, float* out, cudaStream_t st);
void k_argmax_seq(int steps, int n, int D, const float* probs, int ld, uint8_t* out, cudaStream_t st);
// in place on the packed outputs: velocity override rules of process_decoder_outputs + held-note roll (held may be null)
void k_postprocess_voices(int n, int T, int voices, int silent, float thr, int scope, int do_override, const uint8_t* pitch, const uint8_t* song_start,
                          float* vel, uint8_t* held, cudaStream_t st);
void k_swap_shift(DT act, int n, int L, int ldl, const float* mu, const uint8_t* song_start, int c_from, int c_to, int has_hist, void* q,
                  int ldq, float* z_sw, cudaStream_t st);
void k_self_history(DT act, int n, int L, int ldl, const float* z, const uint8_t* song_start, const float* carry, int carry_valid, void* q, int ldq,
                    cudaStream_t st);
void k_build_q(DT act, int n, int L, const float* z, const float* hist, int has_hist, void* q, int ldq, cudaStream_t st);

}  // namespace mvae
