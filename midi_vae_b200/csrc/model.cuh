// model.cuh -- the MIDI-VAE model handle: parameter arena, workspace, and the orchestration of one
// train / eval / predict / style-transfer call as a fixed sequence of kernel launches on one stream.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../../include/midivae.h"
#include "common.cuh"
#include "kernels.cuh"
#include "rec_persist.cuh"

namespace mvae {

struct ParamT {
  std::string name;
  size_t off;
  int rows, cols, ld;
};

enum InKind { IN_NONE = 0, IN_DENSE = 1, IN_RANK1 = 2 };

// One LSTM recurrence (an encoder layer or a decoder cell) with its sequence buffers, all time-major.
struct Rec {
  std::string name;
  int steps = 0;
  int Din = 0;     // real input width (K of the input projection)
  int ldin = 0;    // padded width of the dense input rows
  int iW = -1, iU = -1, ib = -1;
  int variant = 0;
  void* xw = nullptr;     // (steps, n, 4H) act : pre-activations x W + b; reused as dG in the backward sweep
  void* hseq = nullptr;   // (steps+1, n, H) act : slot 0 = initial state
  void* cseq = nullptr;   // (steps+1, n, H) act
  void* gates = nullptr;  // (steps, n, 4H) act : post-activation gates
  void* dhext = nullptr;  // (steps, n, H) act : gradient arriving from the layer / head above
  void* upack_b = nullptr;  // (cpg, H, 64) bf16 : recurrent weights packed for the K-split persistent backward kernel
  void* upack = nullptr;  // (4H, H) bf16 : recurrent weights packed per CTA for the persistent forward kernel
  void* xtab = nullptr;   // (64, 4H) bf16 : one-hot input projection table of the cluster forward kernel
};

// one forward recurrence: its input sequence and initial states
struct FwdJob {
  Rec* r = nullptr;
  int kind = IN_NONE;
  const void* X = nullptr;
  const void* h0 = nullptr;
  const void* c0 = nullptr;
  int ld0 = 0;
  // cluster forward: the one-hot input as class indices (u8 [n][idx_ld], step t reads column t - idx_shift), instead of the dense rows X
  const unsigned char* idx = nullptr;
  int idx_ld = 0, idx_shift = 0;
  bool onehot = false;
};

// one backward recurrence: where its inputs / external gradients come from and where its outputs go
struct BwdJob {
  Rec* r = nullptr;
  int kind = IN_NONE;
  const void* X = nullptr;
  bool use_dhext = false;
  const void* dh_last = nullptr;
  int ld_last = 0;
  bool need_dx = false;
  void* dx_out = nullptr;
  void* dS_h = nullptr;
  void* dS_c = nullptr;
  int ldS = 0;
};

enum ProfClass { PC_REC_FWD = 0, PC_REC_BWD, PC_GEMM, PC_POINTWISE, PC_ADAM, PC_ALLREDUCE, PC_COUNT };

struct Model {
  mvae_config cfg{};
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int sm_count = 148;
  DT act = DT_F32;
  // dimensions
  int T = 0, H = 0, L = 0, Dp = 0, Di = 0, Ti = 0, C = 0, ne = 0, nd = 0, G = 0, NB = 0;
  bool gru = false;                    // cell_type GRU: G = 3H (blocks [z|r|h]), one state per cell, step-streamed recurrences only
  int spc = 2;                         // states per decoder cell (LSTM: h, c; GRU: h)
  int PD = 64, ID = 16, VD = 8;        // padded widths of the dense pitch / instrument / velocity rows
  int ldl = 0, Q = 0, ldq = 0, nS = 0, half = 0;
  int ld_pn = 64, ld_pi = 16, ld_pv = 8;
  // parameters
  std::vector<ParamT> ptab;
  size_t arena_n = 0;
  float *P = nullptr, *Gr = nullptr, *M1 = nullptr, *V2 = nullptr;
  __nv_bfloat16* Pb = nullptr;
  long long iterations = 0;
  int iWa = -1, iba = -1, iWe = -1, ibe = -1, iWmu = -1, ibmu = -1, iWlv = -1, iblv = -1, iWinit = -1, ibinit = -1;
  int iWy = -1, iby = -1, iWio = -1, ibio = -1, iWvo = -1, ibvo = -1;
  // recurrences
  std::vector<Rec> enc_pitch, dec_notes;
  Rec enc_instr, enc_vel, dec_instr, dec_vel;
  // workspace
  char* ws = nullptr;
  size_t ws_bytes = 0, ws_used = 0;
  // device copies of one mini-batch (host API) + expanded rolls
  uint8_t *d_pitch = nullptr, *d_target = nullptr, *d_instr = nullptr, *d_style = nullptr, *d_song_start = nullptr;
  float *d_vel = nullptr, *d_hist = nullptr, *d_eps = nullptr, *d_w = nullptr;
  void *Xp_ext = nullptr, *Yp_ext = nullptr, *Xi_ext = nullptr, *Xv_ext = nullptr;
  // per-step scratch
  float *pre = nullptr, *c_run = nullptr, *dh_run = nullptr, *dc_run = nullptr;
  void* xstep = nullptr;  // (n, 64) act : free-running decoder input
  // head
  void *u = nullptr, *a1 = nullptr, *e = nullptr, *q = nullptr, *S = nullptr, *dS = nullptr, *dSpre = nullptr, *dq = nullptr;
  void *dmu = nullptr, *dlv = nullptr, *de = nullptr, *dpre_e = nullptr, *da1 = nullptr, *dpre_a = nullptr, *du = nullptr;
  float *mu = nullptr, *lv = nullptr, *z = nullptr, *style_probs = nullptr;
  // output heads
  float *Pn = nullptr, *Pi = nullptr, *Pv = nullptr;
  void *dlog_n = nullptr, *dlog_i = nullptr, *dlog_v = nullptr;
  // metrics
  double* acc = nullptr;
  float* d_metrics = nullptr;
  // staging for the host API
  char* pin = nullptr;
  size_t pin_bytes = 0;
  float *o_y = nullptr, *o_i = nullptr, *o_v = nullptr, *o_z = nullptr;
  uint8_t *o_pitch = nullptr, *o_instr = nullptr;
  // data parallel
  void* nccl_comm = nullptr;
  int world = 1, rank = 0;
  // data parallel: the gradient arena is all-reduced in two buckets on a communication stream -- the decoder's tensors (the tail of the
  // arena; final ~4 ms before the step ends) while the encoder's reverse sweeps still run, the encoder's at the join (api.cu)
  cudaStream_t st_comm = nullptr;
  cudaEvent_t ev_dec_grads = nullptr, ev_comm = nullptr, ev_pre_comm = nullptr;
  bool overlap_allreduce = false;     // set by the fused train step only: the split API (forward_backward + apply_update) leaves Gr untouched
  bool dec_bucket_issued = false;
  int ar_buckets = 2;                 // MVAE_AR_BUCKETS=1: one all-reduce after the join (round-1 behaviour)
  // profiling
  bool profiling = false;
  struct Ev { int cls; cudaEvent_t a, b; const char* tag; };
  bool fuse_dual_wgrad = true;        // MVAE_WGRAD_DUAL=0: dU and dW of a dense-input recurrence as two GEMM launches (each streams dG)
  bool fuse_wgrad_rows = true;        // MVAE_WGRAD_ROWS=0: scalar-input recurrences take one column-sum pass per output (round 1) instead of one pass
  bool prof_detail = false;           // MVAE_TIMELINE=2: every weight-gradient launch gets its own event pair and a tag
  std::vector<Ev> evs;
  float prof_ms[PC_COUNT] = {0};
  long long prof_n[PC_COUNT] = {0};
  long long launches = 0;
  size_t h2d_bytes = 0, d2h_bytes = 0;
  cudaStream_t st = nullptr;          // stream of the call in flight
  cudaStream_t side = nullptr;        // weight-gradient GEMMs run here, next to the (SM-sparse) backward recurrences
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool use_side = false;
  bool use_gru_cluster = false;       // GRU, bf16, H = 256: gru_cluster.cu (MVAE_GRU_CLUSTER=0 or rnn_mode=streamed: step-streamed kernels)
  int* gemm_sched = nullptr;          // dynamic tile scheduler words of the GEMM kernel, one pair per stream
  int side_sms = 0;
  // independent recurrences (velocity / instrument streams) run on a branch stream next to the pitch stack: a cluster recurrence
  // occupies 4 of the chip's 7 cluster slots, so two of them overlap almost completely
  cudaStream_t st_branch = nullptr, st_saved = nullptr;
  cudaEvent_t ev_bfork = nullptr, ev_bjoin = nullptr;
  bool use_branch = false;
  int cur_slot = 0;                   // which set of exchange buffers the recurrence being issued uses
  bool fork_pending = false;          // the fork point is the launch of the next main-stream recurrence (it must win the cluster slots)
  void fork_if_pending();
  void branch_fork();
  void branch_begin();
  void branch_end();
  void branch_join();
  const unsigned char* cur_pitch = nullptr;   // device u8 rolls of the batch in flight (class indices)
  const unsigned char* cur_target = nullptr;
  bool inference_pass = false;      // forward only (style transfer / predict): the recurrences skip the BPTT stash
  bool fuse_xproj = false;          // cluster forward computes one-hot / scalar input projections itself
  const void* Y_ext_cur = nullptr;   // teacher-forcing source: Yp_ext or (target == pitch) Xp_ext
  const void* e_cur = nullptr;       // output of the last tanh Dense before the split
  bool stepwise_done = false;
  // output post-processing on the device (vae_definition.py:1156-1190): 0 = off, 1 = voice memory restarts per chunk, 2 = per song
  int post_scope = 0; float post_threshold = 0.5f; int post_override = 1; int post_voices = 4;
  uint8_t* o_held = nullptr;
  bool use_persist = false;          // persistent-RNN kernels (bf16 precision, supported hidden size)
  bool use_cluster_fwd = false;      // cluster forward recurrence kernel (lstm_cluster.cu)
  bool use_cluster_bwd = false;      // cluster backward recurrence kernel (lstm_cluster.cu)
  long long* trace_buf = nullptr;    // MVAE_REC_TRACE=1 debugging aid
  int trace_dumps = 0;
  void* rec_hx2 = nullptr;           // second recurrence of a paired launch
  void* rec_hx = nullptr;            // h exchange buffer of the persistent forward kernel
  void* rec_partial = nullptr;       // bf16 partial-dh exchange buffer of the K-split backward kernel
  unsigned* rec_flags = nullptr;
  unsigned* rec_flags2 = nullptr;    // second recurrence of a paired launch
  void* rec_partial2 = nullptr;
  bool pair_recs = true;             // MVAE_REC_PAIR=0 disables pairing two recurrences per launch     // per-(group, step) publication counters of the persistent kernels
  std::vector<void*> allocs_;

  explicit Model(const mvae_config& c, int dev);
  ~Model();

  // helpers
  const void* W(int idx) const { return act == DT_F32 ? (const void*)(P + ptab[idx].off) : (const void*)(Pb + ptab[idx].off); }
  const float* Wf(int idx) const { return P + ptab[idx].off; }
  float* Gp(int idx) const { return Gr + ptab[idx].off; }
  int ld(int idx) const { return ptab[idx].ld; }
  size_t asz() const { return dt_size(act); }
  void* slab(void* base, long idx, long elems_per_slab) const { return (char*)base + (size_t)idx * elems_per_slab * asz(); }
  CellCfg cc(int variant) const { return CellCfg{cfg.gate_act, variant}; }

  void* alloc(size_t bytes);
  int add_param(const std::string& name, int rows, int cols);
  void build_params();
  void build_workspace();
  void commit_params();
  void gemm(GemmArgs g);
  void gemm_on(GemmArgs g, cudaStream_t s, int sms);
  void prof_begin(int cls, cudaStream_t s = nullptr, const char* tag = nullptr);
  void prof_end(cudaStream_t s = nullptr);
  void rec_backward_wgrads(const BwdJob& j, int n, cudaStream_t s, int sms);
  void prof_collect();
  void dump_trace(const char* dir, const Rec& r, int nctas = 1);

  // batch plumbing
  mvae_batch upload(const mvae_batch& hb, const uint8_t* song_start = nullptr);
  void check_batch(const mvae_batch& b, bool need_style) const;

  // graph pieces
  void prepare_inputs(const mvae_batch& b, bool need_target);
  void rec_forward(Rec& r, int n, int kind, const void* X, const void* h0, const void* c0, int ld0);
  void rec_steps_forward(Rec& r, int n, int t0, int t1);
  void gru_steps_forward(Rec& r, int n, int t0, int t1);
  void gru_backward_sweep(const BwdJob& j, int n);
  void rec_forward_prepare(const FwdJob& j, int n);
  RecPersistArgs fwd_args(const FwdJob& j, int n, int slot, int hs, bool pack = true);
  void rec_forward_jobs(const FwdJob* ja, const FwdJob* jb, int n);
  RecPersistArgs bwd_args(const BwdJob& j, int n, int slot, int hs, bool pack = true);
  void rec_backward_sweep(const BwdJob* ja, const BwdJob* jb, int n);
  void rec_backward_gemms(const BwdJob& j, int n, bool tail = false);
  void rec_backward_group(std::vector<BwdJob>& stack, std::vector<BwdJob>& side, int n, bool last_group = false);
  void encoder_forward(int n);
  void head_forward(const mvae_batch& b, bool with_style_loss);
  void decoder_forward(const mvae_batch& b, int feedback);
  // opt-in: the decoder's history input is built from THIS batch's own z (shifted inside each song) instead of a separate encoder pass
  // (mvae_set_history_mode).  carry = z of the last row of the previous call, for a song that continues across two calls.
  int history_mode = 0; bool song_start_set = false; bool carry_valid = false; float* hist_carry = nullptr;
  // style-classifier mode (mvae_config::model_kind = 1): enc_pitch[] is the recurrent stack, (iWy, iby) the softmax Dense, Pn the logits
  bool cls = false, cls_scalar = false;
  void cls_forward(const mvae_batch& b, bool train);
  void cls_backward(const mvae_batch& b);
  void decoder_stepwise(int n);
  void decoder_stepwise_body(int n);
  // free-running decode = thousands of tiny dependent launches: captured once per batch size into a CUDA graph (notes chain on the step's
  // stream, velocity + instrument chains on the branch stream with their own scratch) and replayed; MVAE_STEPWISE_GRAPH=0 launches eagerly
  // step-streamed train steps (GRU cells, fp32 precision, rnn_mode=streamed) are thousands of small dependent launches on ONE stream: the whole
  // forward + backward of a mini-batch is captured once per (batch size, buffer addresses) and replayed; MVAE_STEP_GRAPH=0 launches eagerly
  struct StepGraph { int state = 0; cudaGraphExec_t exec = nullptr; long long launches = 0; };
  std::map<std::vector<size_t>, StepGraph> step_graphs;
  bool step_graph_on = true;
  void forward_backward_body(const mvae_batch& b, float* dev_metrics);
  struct StepwiseGraph { int state = 0; cudaGraphExec_t exec = nullptr; long long launches = 0; };
  std::map<int, StepwiseGraph> stepwise_graphs;
  bool stepwise_graph_on = true;
  float *pre_b = nullptr, *c_run_b = nullptr; void* xstep_b = nullptr;   // scratch of the branch chain
  void losses(const mvae_batch& b, bool train);
  void backward(const mvae_batch& b);
  void forward_backward(const mvae_batch& b, float* dev_metrics, cudaStream_t st);
  void allreduce_grads();
  void apply_update(float grad_scale, cudaStream_t st);
  void eval_step(const mvae_batch& b, float* dev_metrics, cudaStream_t st);
  void style_transfer(const mvae_batch& b, const uint8_t* song_start, int c_from, int c_to, int feedback, uint8_t* pitch_out,
                      uint8_t* instr_out, float* vel_out, cudaStream_t st);
};

// NCCL through dlopen (libnccl.so.2); no link-time dependency
int nccl_get_unique_id(void* out128);
void* nccl_comm_init(const void* id128, int world, int rank);
void nccl_comm_destroy(void* comm);
void nccl_allreduce_sum_f32(void* comm, float* buf, size_t count, cudaStream_t st);

}  // namespace mvae
