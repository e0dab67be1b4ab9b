// model.cu -- orchestration of the MIDI-VAE hot path on one B200.
//
// What the reference runs as one Keras train_function / predict_function (graph built in
// vae_definition.py:212-441, encoder :443-516, decoder :519-645, style head :730-734; driven from
// vae_training.py:804-809 and vae_evaluation.py:2180-2181,2474-2483) is laid out here as an explicit,
// fixed sequence of kernel launches on one CUDA stream:
//   rolls -> dense padded time-major inputs -> input projections (one GEMM per layer over the whole
//   sequence) -> recurrences (h U + gate math; step-streamed or persistent) -> latent head -> decoder
//   initial states (one fused GEMM for all 2*(nd+2) Denses) -> decoder recurrences -> output heads +
//   Keras losses -> reverse-time sweeps -> batched weight-gradient GEMMs -> [NCCL all-reduce] -> Adam.
// The hand-derived backward is restated (and checked against autograd) in oracle/manual_bptt.py.
#include "model.cuh"

#include <cuda.h>

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace mvae {

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

void* Model::alloc(size_t bytes) {
  void* p = nullptr;
  bytes = (bytes + 255) / 256 * 256;
  if (bytes == 0) bytes = 256;
  MVAE_CUDA(cudaMalloc(&p, bytes));
  MVAE_CUDA(cudaMemsetAsync(p, 0, bytes, stream));
  ws_used += bytes;
  allocs_.push_back(p);
  return p;
}

int Model::add_param(const std::string& name, int rows, int cols) {
  ParamT t;
  t.name = name; t.rows = rows; t.cols = cols; t.ld = round_up(cols, 8);
  t.off = arena_n;
  arena_n += (size_t)round_up(rows * t.ld, 64);
  ptab.push_back(t);
  return (int)ptab.size() - 1;
}

// Internal arena order.  Names match oracle/midivae_oracle.py::param_specs except that the 2*(nd+2)
// decoder initial-state Denses (vae_definition.py:563-568,599-604,637-642) are stored column-fused as
// dec_init/kernel (Q, nS*H) so that they are ONE GEMM; midi_vae_b200/weights.py maps both ways.
void Model::build_params() {
  auto add_rec = [&](Rec& r, const std::string& nm, int steps, int Din, int ldin, int variant, bool keras) {
    r.name = nm; r.steps = steps; r.Din = Din; r.ldin = ldin; r.variant = variant;
    r.iW = add_param(nm + "/kernel", Din, G);
    if (keras) { r.iU = add_param(nm + "/recurrent_kernel", H, G); r.ib = add_param(nm + "/bias", 1, G); }
    else { r.ib = add_param(nm + "/bias", 1, G); r.iU = add_param(nm + "/recurrent_kernel", H, G); }
  };
  enc_pitch.resize(ne);
  const std::string pre = gru ? "gru_" : "lstm_";     // Keras layer names of the two branches (vae_definition.py:457-472)
  if (cls) {   // pitch_classifier.py:89-103: GRU x num_layers -> Dense(num_classes, softmax); names as in the shipped classifier files
    for (int k = 0; k < ne; ++k)
      add_rec(enc_pitch[k], pre + std::to_string(k + 1), T, k == 0 ? (cls_scalar ? 1 : Dp) : H, k == 0 ? (cls_scalar ? VD : PD) : H, MVAE_CELL_STANDARD, true);
    iWy = add_param("dense_1/kernel", H, C); iby = add_param("dense_1/bias", 1, C);
    ld_pn = ptab[iWy].ld;
    return;
  }
  for (int k = 0; k < ne; ++k) add_rec(enc_pitch[k], pre + std::to_string(k + 1), T, k == 0 ? Dp : H, k == 0 ? PD : H, MVAE_CELL_STANDARD, true);
  add_rec(enc_instr, pre + "meta_instrument", Ti, Di, ID, MVAE_CELL_STANDARD, true);
  add_rec(enc_vel, pre + "meta_velocity", T, 1, VD, MVAE_CELL_STANDARD, true);
  iWa = add_param("extra_instrument_after_concat_layer/kernel", 3 * H, H);
  iba = add_param("extra_instrument_after_concat_layer/bias", 1, H);
  if (cfg.extra_layer) { iWe = add_param("extra_layer/kernel", H, H); ibe = add_param("extra_layer/bias", 1, H); }
  iWmu = add_param("z_mean/kernel", half, L); ibmu = add_param("z_mean/bias", 1, L);
  iWlv = add_param("z_log_var/kernel", H - half, L); iblv = add_param("z_log_var/bias", 1, L);
  iWinit = add_param("dec_init/kernel", Q, nS * H); ibinit = add_param("dec_init/bias", 1, nS * H);
  const int dvar = gru ? 1 : cfg.dec_cell_variant;      // GRU: mix 1 = recurrentshop GRUCell (encoders use mix 0 = Keras GRU)
  dec_notes.resize(nd);
  for (int k = 0; k < nd; ++k)
    add_rec(dec_notes[k], "notes/cell_" + std::to_string(k + 1), T, k == 0 ? Dp : H, k == 0 ? PD : H, dvar, false);
  iWy = add_param("notes/out/kernel", H, Dp); iby = add_param("notes/out/bias", 1, Dp);
  add_rec(dec_instr, "meta_instrument/cell", Ti, Di, ID, dvar, false);
  iWio = add_param("meta_instrument/out/kernel", H, Di); ibio = add_param("meta_instrument/out/bias", 1, Di);
  add_rec(dec_vel, "meta_velocity/cell", T, 1, VD, dvar, false);
  // (H,1) stored as the row vector (1,H): same dense bytes, and the N = 1 head is a row-dot, not a GEMM
  iWvo = add_param("meta_velocity/out/kernel", 1, H); ibvo = add_param("meta_velocity/out/bias", 1, 1);
  ld_pn = ptab[iWy].ld; ld_pi = ptab[iWio].ld; ld_pv = 8;
}

void Model::build_workspace() {
  const size_t a = asz();
  const size_t n = NB;
  auto rec_bufs = [&](Rec& r, bool need_dhext) {
    r.xw = alloc((size_t)r.steps * n * G * a);
    r.gates = alloc((size_t)r.steps * n * G * a);
    r.hseq = alloc((size_t)(r.steps + 1) * n * H * a);
    r.cseq = alloc((size_t)(r.steps + 1) * n * H * a);
    if (need_dhext) r.dhext = alloc((size_t)r.steps * n * H * a);
    if (use_persist) r.upack = alloc((size_t)G * H * 2);
    if (use_persist && rec_cluster_supported(H)) r.xtab = alloc((size_t)64 * G * 2);
    if (use_persist && (rec_persist_ksplit_ok(H) || rec_cluster_bwd_supported(H))) r.upack_b = alloc((size_t)G * H * 2);
  };
  if (use_persist && (rec_persist_ksplit_ok(H) || rec_cluster_bwd_supported(H))) {
    const size_t pb = std::max(rec_persist_partial_bytes(NB, H), rec_cluster_bwd_supported(H) ? rec_cluster_xbuf_bytes(NB, H) : (size_t)0);
    rec_partial = alloc(pb); rec_partial2 = alloc(pb);
  }
  if (use_persist) rec_flags2 = (unsigned*)alloc(rec_persist_flag_count(NB, std::max(T, Ti)) * sizeof(unsigned));
  if (use_persist) {
    const size_t hxb = std::max(rec_persist_hx_bytes(NB, H), rec_cluster_supported(H) ? rec_cluster_hx_bytes(NB, H) : (size_t)0);
    rec_hx = alloc(hxb); rec_hx2 = alloc(hxb);
  }
  if (use_persist) rec_flags = (unsigned*)alloc(rec_persist_flag_count(NB, std::max(T, Ti)) * sizeof(unsigned));
  for (int k = 0; k < ne; ++k) rec_bufs(enc_pitch[k], k < ne - 1);
  if (cls) {   // the classifier needs the stack, its input expansion, the logits and the head gradients only
    d_pitch = (uint8_t*)alloc(n * T); d_style = (uint8_t*)alloc(n); d_vel = (float*)alloc(n * T * 4);
    d_target = d_pitch; d_instr = d_pitch; d_hist = d_vel; d_eps = d_vel; d_w = d_vel; d_song_start = d_style;   // never uploaded in this mode
    if (cls_scalar) Xv_ext = alloc((size_t)(T + 1) * n * VD * a); else Xp_ext = alloc((size_t)(T + 1) * n * PD * a);
    pre = (float*)alloc(n * G * 4); c_run = (float*)alloc(n * H * 4); dh_run = (float*)alloc(n * H * 4); dc_run = (float*)alloc(n * H * 4);
    Pn = (float*)alloc(n * ld_pn * 4); dlog_n = alloc(n * ld_pn * a); du = alloc(n * H * a);
    acc = (double*)alloc(ACC_COUNT * 8); d_metrics = (float*)alloc(MVAE_NUM_METRICS * 4); o_y = (float*)alloc(n * C * 4);
    pin_bytes = n * T * 5 + n + (size_t)n * C * 4 + 4096;
    MVAE_CUDA(cudaMallocHost((void**)&pin, pin_bytes));
    return;
  }
  rec_bufs(enc_instr, false); rec_bufs(enc_vel, false);
  for (int k = 0; k < nd; ++k) rec_bufs(dec_notes[k], true);
  rec_bufs(dec_instr, true); rec_bufs(dec_vel, true);

  d_pitch = (uint8_t*)alloc(n * T); d_target = (uint8_t*)alloc(n * T); d_instr = (uint8_t*)alloc(n * Ti);
  d_style = (uint8_t*)alloc(n); d_song_start = (uint8_t*)alloc(n);
  d_vel = (float*)alloc(n * T * 4); d_hist = (float*)alloc(n * L * 4); d_eps = (float*)alloc(n * L * 4); d_w = (float*)alloc(n * T * 4);
  Xp_ext = alloc((size_t)(T + 1) * n * PD * a); Yp_ext = alloc((size_t)(T + 1) * n * PD * a);
  Xi_ext = alloc((size_t)(Ti + 1) * n * ID * a); Xv_ext = alloc((size_t)(T + 1) * n * VD * a);
  pre = (float*)alloc(n * G * 4); c_run = (float*)alloc(n * H * 4); dh_run = (float*)alloc(n * H * 4); dc_run = (float*)alloc(n * H * 4);
  xstep = alloc(n * 64 * a);
  pre_b = (float*)alloc(n * G * 4); c_run_b = (float*)alloc(n * H * 4); xstep_b = alloc(n * 64 * a);
  u = alloc(n * 3 * H * a); a1 = alloc(n * H * a); e = alloc(n * H * a); q = alloc(n * ldq * a);
  S = alloc(n * nS * H * a); dS = alloc(n * nS * H * a); dSpre = alloc(n * nS * H * a); dq = alloc(n * ldq * a);
  dmu = alloc(n * ldl * a); dlv = alloc(n * ldl * a); de = alloc(n * H * a); dpre_e = alloc(n * H * a); da1 = alloc(n * H * a);
  dpre_a = alloc(n * H * a); du = alloc(n * 3 * H * a);
  mu = (float*)alloc(n * ldl * 4); lv = (float*)alloc(n * ldl * 4); z = (float*)alloc(n * ldl * 4); style_probs = (float*)alloc(n * C * 4);
  hist_carry = (float*)alloc(ldl * 4);
  Pn = (float*)alloc((size_t)T * n * ld_pn * 4); Pi = (float*)alloc((size_t)Ti * n * ld_pi * 4); Pv = (float*)alloc((size_t)T * n * ld_pv * 4);
  dlog_n = alloc((size_t)T * n * ld_pn * a); dlog_i = alloc((size_t)Ti * n * ld_pi * a); dlog_v = alloc((size_t)T * n * ld_pv * a);
  acc = (double*)alloc(ACC_COUNT * 8); d_metrics = (float*)alloc(MVAE_NUM_METRICS * 4);
  o_y = (float*)alloc((size_t)n * T * Dp * 4); o_i = (float*)alloc((size_t)n * Ti * Di * 4); o_v = (float*)alloc((size_t)n * T * 4);
  o_z = (float*)alloc((size_t)n * L * 4 * 3); o_pitch = (uint8_t*)alloc(n * T); o_instr = (uint8_t*)alloc(n * Ti); o_held = (uint8_t*)alloc(n * T);
  // pinned staging: inputs + the largest output set
  pin_bytes = n * T * 2 + n * Ti + n * 2 + (size_t)n * T * 4 * 2 + (size_t)n * L * 4 * 3 + (size_t)n * T * Dp * 4 + (size_t)n * Ti * Di * 4 +
              (size_t)n * T * 4 + (size_t)n * C * 4 + 4096;
  MVAE_CUDA(cudaMallocHost((void**)&pin, pin_bytes));
}

Model::Model(const mvae_config& c, int dev) : cfg(c), device(dev) {
  MVAE_REQUIRE(c.input_length > 0 && c.lstm_size > 0 && c.latent_rep_size > 0, "input_length, lstm_size, latent_rep_size must be > 0 (vae_definition.py:179-182)");
  MVAE_REQUIRE(c.num_layers_encoder > 0 && c.num_layers_decoder > 0, "num_layers must be > 0 (vae_definition.py:177-178)");
  MVAE_REQUIRE(c.beta > 0, "beta must be > 0 (vae_definition.py:183)");
  MVAE_REQUIRE(c.max_batch > 0, "max_batch must be > 0");
  MVAE_REQUIRE(c.meta_instrument_length > 0, "meta_instrument_length must be > 0 (settings.py:182 uses 4; the instrument loss divides by it)");
  MVAE_REQUIRE(c.input_dim > 0 && c.input_dim <= 64, "input_dim must be in 1..64");
  MVAE_REQUIRE(c.meta_instrument_dim > 0 && c.meta_instrument_dim <= 64, "meta_instrument_dim must be in 1..64");
  MVAE_REQUIRE(c.num_composers >= 1 && c.num_composers <= c.latent_rep_size, "num_composers must be in 1..latent_rep_size");
  MVAE_REQUIRE(c.split_lstm_vector == 1, "only split_lstm_vector=True (settings.py:139) is implemented");
  MVAE_REQUIRE(c.precision == MVAE_PREC_FP32 || c.precision == MVAE_PREC_BF16, "precision");
  MVAE_REQUIRE(c.decoder_feedback >= 0 && c.decoder_feedback <= 2, "decoder_feedback");
  if (c.precision == MVAE_PREC_BF16) MVAE_REQUIRE(c.lstm_size % 16 == 0, "bf16 precision needs lstm_size % 16 == 0");
  MVAE_CUDA(cudaSetDevice(dev));
  int major = 0;
  MVAE_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  MVAE_REQUIRE(major == 10, "libmidivae.so is built for sm_100a (B200) only");
  MVAE_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  int prio_least = 0, prio_greatest = 0;
  MVAE_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  MVAE_CUDA(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, prio_greatest));   // the critical chain
  MVAE_CUDA(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, prio_least));
  MVAE_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
  MVAE_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
  MVAE_CUDA(cudaStreamCreateWithPriority(&st_branch, cudaStreamNonBlocking, prio_least));
  MVAE_CUDA(cudaEventCreateWithFlags(&ev_bfork, cudaEventDisableTiming));
  MVAE_CUDA(cudaEventCreateWithFlags(&ev_bjoin, cudaEventDisableTiming));
  act = c.precision == MVAE_PREC_FP32 ? DT_F32 : DT_BF16;
  T = c.input_length; H = c.lstm_size; L = c.latent_rep_size; Dp = c.input_dim; Di = c.meta_instrument_dim; Ti = c.meta_instrument_length;
  MVAE_REQUIRE(c.cell_type == MVAE_CELLTYPE_LSTM || c.cell_type == MVAE_CELLTYPE_GRU, "cell_type must be LSTM or GRU (vae_definition.py:457-472)");
  gru = c.cell_type == MVAE_CELLTYPE_GRU; spc = gru ? 1 : 2;
  cls = c.model_kind == 1; cls_scalar = cls && c.cls_scalar_input != 0;
  MVAE_REQUIRE(c.model_kind == 0 || c.model_kind == 1, "model_kind: 0 (VAE) or 1 (style classifier)");
  C = c.num_composers; ne = c.num_layers_encoder; nd = c.num_layers_decoder; G = (gru ? 3 : 4) * H; NB = c.max_batch;
  PD = round_up(Dp, 8); ID = round_up(Di, 8); VD = 8;
  ldl = round_up(L, 8); Q = c.history ? 2 * L : L; ldq = round_up(Q, 8); nS = spc * (nd + 2); half = H / 2;
  use_persist = !gru && act == DT_BF16 && cfg.rnn_mode != MVAE_RNN_STREAMED && rec_persist_supported(H, sm_count);
  if (cfg.rnn_mode == MVAE_RNN_PERSISTENT)
    MVAE_REQUIRE(use_persist, "rnn_mode=persistent needs the LSTM cell, bf16 precision and lstm_size % 64 == 0 with a weight slice that fits in shared memory");
  use_cluster_fwd = use_persist && rec_cluster_supported(H);
  use_cluster_bwd = use_persist && rec_cluster_bwd_supported(H);
  // the cluster recurrences occupy 16 SMs per 64..128 batch rows and leave the rest of the chip idle: the batched weight-gradient
  // GEMMs (needed only by the optimizer) run next to them on a second stream, on a grid sized for the idle SMs
  // the reference's default cell at its default size: one cluster-resident launch per recurrence instead of four launches per step
  use_gru_cluster = gru && act == DT_BF16 && cfg.rnn_mode != MVAE_RNN_STREAMED && gru_cluster_supported(H);
  { const char* e = getenv("MVAE_SIDE_STREAM"); use_side = (use_cluster_bwd || use_gru_cluster) && (e ? atoi(e) != 0 : true); }
  { const char* e = getenv("MVAE_SIDE_SMS"); side_sms = e ? atoi(e) : 0; }
  { const char* e = getenv("MVAE_FUSE_XPROJ"); fuse_xproj = use_cluster_fwd && (e ? atoi(e) != 0 : true); }
  { const char* e = getenv("MVAE_BRANCH"); use_branch = ((use_cluster_fwd && use_cluster_bwd) || use_gru_cluster) && (e ? atoi(e) != 0 : true); }
  if (side_sms <= 0) side_sms = std::max(16, sm_count - 16 * ((NB + 127) / 128));
  { const char* e = getenv("MVAE_WGRAD_DUAL"); fuse_dual_wgrad = e ? atoi(e) != 0 : true; }
  { const char* e = getenv("MVAE_WGRAD_ROWS"); fuse_wgrad_rows = e ? atoi(e) != 0 : true; }
  { const char* e = getenv("MVAE_TIMELINE"); prof_detail = e && atoi(e) >= 2; }
  { const char* e = getenv("MVAE_STEP_GRAPH"); step_graph_on = e ? atoi(e) != 0 : true; }
  { const char* e = getenv("MVAE_STEPWISE_GRAPH"); stepwise_graph_on = e ? atoi(e) != 0 : true; }
  { const char* e = getenv("MVAE_AR_BUCKETS"); ar_buckets = e ? std::max(1, atoi(e)) : 2; }
  build_params();
  P = (float*)alloc(arena_n * 4); Gr = (float*)alloc(arena_n * 4); M1 = (float*)alloc(arena_n * 4); V2 = (float*)alloc(arena_n * 4);
  if (act == DT_BF16) Pb = (__nv_bfloat16*)alloc(arena_n * 2);
  build_workspace();
  gemm_sched = (int*)alloc(8 * sizeof(int));
  { const char* e = getenv("MVAE_REC_PAIR"); pair_recs = e ? atoi(e) != 0 : true; }
  if (getenv("MVAE_REC_TRACE")) trace_buf = (long long*)alloc(512 * sizeof(long long));
  MVAE_CUDA(cudaStreamSynchronize(stream));
}

Model::~Model() {
  cudaSetDevice(device);
  if (stream) cudaStreamSynchronize(stream);
  for (auto& ev : evs) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
  for (auto& kv : stepwise_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  for (auto& kv : step_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  if (nccl_comm) nccl_comm_destroy(nccl_comm);
  if (st_comm) cudaStreamDestroy(st_comm);
  for (cudaEvent_t e : {ev_dec_grads, ev_comm, ev_pre_comm}) if (e) cudaEventDestroy(e);
  for (void* p : allocs_) cudaFree(p);
  if (pin) cudaFreeHost(pin);
  if (stream) cudaStreamDestroy(stream);
  if (side) cudaStreamDestroy(side);
  if (st_branch) cudaStreamDestroy(st_branch);
  if (ev_bfork) cudaEventDestroy(ev_bfork);
  if (ev_bjoin) cudaEventDestroy(ev_bjoin);
  if (ev_fork) cudaEventDestroy(ev_fork);
  if (ev_join) cudaEventDestroy(ev_join);
}

void Model::commit_params() {
  if (Pb) k_f32_to_bf16((long)arena_n, P, Pb, st);
}

// --------------------------------------------------------------------------------------------- profiling
void Model::prof_begin(int cls, cudaStream_t s, const char* tag) {
  if (!profiling) return;
  Ev ev; ev.cls = cls; ev.tag = tag;
  MVAE_CUDA(cudaEventCreate(&ev.a)); MVAE_CUDA(cudaEventCreate(&ev.b));
  MVAE_CUDA(cudaEventRecord(ev.a, s ? s : st));
  evs.push_back(ev);
}
void Model::prof_end(cudaStream_t s) {
  if (!profiling) return;
  MVAE_CUDA(cudaEventRecord(evs.back().b, s ? s : st));
}
void Model::prof_collect() {
  for (int i = 0; i < PC_COUNT; ++i) { prof_ms[i] = 0; prof_n[i] = 0; }
  const bool timeline = getenv("MVAE_TIMELINE") != nullptr && !evs.empty();
  static const char* cls_name[PC_COUNT] = {"rec_fwd", "rec_bwd", "gemm", "pointwise", "adam", "allreduce"};
  for (auto& ev : evs) {
    MVAE_CUDA(cudaEventSynchronize(ev.b));
    float ms = 0;
    MVAE_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
    if (timeline) {
      float t0 = 0;
      MVAE_CUDA(cudaEventElapsedTime(&t0, evs[0].a, ev.a));
      fprintf(stderr, "timeline %-9s start %8.3f ms  dur %7.3f ms  %s\n", cls_name[ev.cls], t0, ms, ev.tag ? ev.tag : "");
    }
    prof_ms[ev.cls] += ms; prof_n[ev.cls] += 1;
  }
  for (auto& ev : evs) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
  evs.clear();
}

// --------------------------------------------------------------------------------------------- branch stream
// fork: remember "now" on the main stream; begin: everything issued until end() goes to the branch stream and starts after the fork point;
// join: the main stream waits for what the branch did.  Host code stays sequential; only the stream the launches go to changes.
void Model::branch_fork() {
  if (!use_branch) return;
  fork_pending = true;
}
// called right before a recurrence kernel is launched: the branch may start once the main stream's preparation kernels are done, so that the
// main recurrence (launched next) gets its cluster slots first and the branch recurrence takes what is left
void Model::fork_if_pending() {
  if (!fork_pending) return;
  MVAE_CUDA(cudaEventRecord(ev_bfork, st));
  fork_pending = false;
}
void Model::branch_begin() {
  if (!use_branch) return;
  fork_if_pending();
  MVAE_CUDA(cudaStreamWaitEvent(st_branch, ev_bfork, 0));
  st_saved = st; st = st_branch; cur_slot = 1;
}
void Model::branch_end() {
  if (!use_branch) return;
  MVAE_CUDA(cudaEventRecord(ev_bjoin, st_branch));
  st = st_saved; cur_slot = 0;
}
void Model::branch_join() {
  if (!use_branch) return;
  MVAE_CUDA(cudaStreamWaitEvent(st, ev_bjoin, 0));
}

// MVAE_REC_TRACE=1: print the per-phase clock64 stamps CTA 0 of a persistent recurrence recorded for 8 steps
void Model::dump_trace(const char* dir, const Rec& r, int nctas) {
  if (!trace_buf || r.steps < 32) return;
  if (trace_dumps >= 16) return;
  ++trace_dumps;
  MVAE_CUDA(cudaStreamSynchronize(st));
  long long hbuf[512];
  MVAE_CUDA(cudaMemcpy(hbuf, trace_buf, sizeof(hbuf), cudaMemcpyDeviceToHost));
  MVAE_CUDA(cudaMemset(trace_buf, 0, sizeof(hbuf)));
  if (nctas > 1) {
    // cluster kernels: CTA 0 (pair leader) and CTA 1 (its partner), each on its own SM clock; cycles relative to the CTA's first stamp of step 17
    fprintf(stderr, "rec trace %s %s: points 0 mma-issue 1 commit 2 tmem-full-seen 3 scratch-written 4 push-issued(t0) 5 barrier 6 bulk-issued 7 h_full-done 8 relay-sent 9 push-issued(t255)\n",
            dir, r.name.c_str());
    for (int c = 0; c < nctas; ++c) {
      long long base = 0;
      for (int k = 0; k < 16; ++k) { const long long v = hbuf[(c * 8 + 1) * 16 + k]; if (v && (!base || v < base)) base = v; }
      for (int sidx = 1; sidx < 5; ++sidx) {
        fprintf(stderr, "  cta %d step %2d:", c, 16 + sidx);
        for (int k = 0; k < 10; ++k) fprintf(stderr, " %d:%lld", k, hbuf[(c * 8 + sidx) * 16 + k] ? hbuf[(c * 8 + sidx) * 16 + k] - base : -1);
        fprintf(stderr, "\n");
      }
    }
    return;
  }
  fprintf(stderr, "rec trace %s %s: cycles relative to the step's first stamp; points: 0 poll-start 1 flag-seen 2 tma-issued 3 first-kblock-landed "
                  "4 last-kblock-landed 5 epi-wait 6 epi-wake 7 stores-done 8 proxy-fence 9 barrier 10 threadfence 11 flag-published\n", dir, r.name.c_str());
  for (int sidx = 0; sidx < 8; ++sidx) {
    long long base = 0;
    for (int k = 0; k < 16; ++k) if (hbuf[sidx * 16 + k] && (!base || hbuf[sidx * 16 + k] < base)) base = hbuf[sidx * 16 + k];
    fprintf(stderr, "  step %2d:", 16 + sidx);
    for (int k = 0; k < 16; ++k) fprintf(stderr, " %d:%lld", k, hbuf[sidx * 16 + k] ? hbuf[sidx * 16 + k] - base : -1);
    fprintf(stderr, "  | abs0 %lld\n", base);
  }
}

// --------------------------------------------------------------------------------------------- GEMM routing
void Model::gemm(GemmArgs g) { gemm_on(g, st, sm_count); }
void Model::gemm_on(GemmArgs g, cudaStream_t s, int sms) {
  g.in_type = act;
  // one scheduler word pair per stream: launches that share a pair must be stream-ordered
  int* sched = gemm_sched + (s == side ? 2 : (s == st_branch ? 4 : 0));
  if (act == DT_BF16 && gemm_tc_supported(g)) gemm_tc(g, s, sms, sched);
  else gemm_simt(g, s);
}

// --------------------------------------------------------------------------------------------- batch plumbing
void Model::check_batch(const mvae_batch& b, bool need_style) const {
  MVAE_REQUIRE(b.n >= 1 && b.n <= NB, "mini-batch size must be in 1..max_batch");
  MVAE_REQUIRE(b.pitch && b.instr && b.velocity, "pitch, instr and velocity rolls are required");
  if (need_style) MVAE_REQUIRE(b.style != nullptr, "style classes are required for train/evaluate");
}

// host batch -> pinned staging -> device input buffers; returns the device-pointer view
mvae_batch Model::upload(const mvae_batch& hb, const uint8_t* song_start) {
  mvae_batch d{}; d.n = hb.n;
  const size_t n = hb.n;
  char* p = pin;
  auto stage = [&](const void* src, void* dst, size_t bytes) -> const void* {
    if (!src) return nullptr;
    memcpy(p, src, bytes);
    MVAE_CUDA(cudaMemcpyAsync(dst, p, bytes, cudaMemcpyHostToDevice, st));
    p += (bytes + 15) / 16 * 16;
    h2d_bytes += bytes;
    return dst;
  };
  d.pitch = (const uint8_t*)stage(hb.pitch, d_pitch, n * T);
  d.target = (const uint8_t*)stage(hb.target, d_target, n * T);
  d.instr = (const uint8_t*)stage(hb.instr, d_instr, n * Ti);
  d.velocity = (const float*)stage(hb.velocity, d_vel, n * T * 4);
  d.style = (const uint8_t*)stage(hb.style, d_style, n);
  d.history = (const float*)stage(hb.history, d_hist, n * L * 4);
  d.eps = (const float*)stage(hb.eps, d_eps, n * L * 4);
  d.w_notes = (const float*)stage(hb.w_notes, d_w, n * T * 4);
  if (song_start) stage(song_start, d_song_start, n);
  return d;
}

void Model::prepare_inputs(const mvae_batch& b, bool need_target) {
  prof_begin(PC_POINTWISE);
  const bool distinct_target = need_target && b.target && b.target != b.pitch;
  k_expand_inputs(act, b.n, T, Ti, PD, ID, VD, b.pitch, b.target, b.instr, b.velocity, Xp_ext, distinct_target ? Yp_ext : nullptr, Xi_ext, Xv_ext, st);
  Y_ext_cur = distinct_target ? Yp_ext : Xp_ext;
  cur_pitch = b.pitch; cur_target = distinct_target ? b.target : b.pitch;
  prof_end();
}

// --------------------------------------------------------------------------------------------- one recurrence, forward
// kind: IN_DENSE  X = (steps, n, ldin) act rows;  IN_RANK1  X = per-row scalar (stride VD);  IN_NONE  x == 0 (as_wired decoder)
void Model::rec_forward(Rec& r, int n, int kind, const void* X, const void* h0, const void* c0, int ld0) {
  FwdJob j; j.r = &r; j.kind = kind; j.X = X; j.h0 = h0; j.c0 = c0; j.ld0 = ld0;
  rec_forward_jobs(&j, nullptr, n);
}

// input projection over the whole sequence (xw = X W + b) and the initial-state slab
void Model::rec_forward_prepare(const FwdJob& j, int n) {
  Rec& r = *j.r;
  const long rows = (long)r.steps * n;
  const bool fused = fuse_xproj && r.steps > 8 && ((j.kind == IN_DENSE && j.onehot) || j.kind == IN_RANK1 || j.kind == IN_NONE);
  prof_begin(PC_GEMM);
  if (fused) {
    // the cluster kernel computes x W + b itself: a row gather for one-hot inputs (table built here), x w + b for the scalar stream
    if (j.kind != IN_RANK1) rec_cluster_build_xtab(W(r.iW), ld(r.iW), j.kind == IN_DENSE ? r.Din : 0, Wf(r.ib), r.xtab, H, st);
  } else if (j.kind == IN_DENSE) {
    GemmArgs g; g.M = (int)rows; g.N = G; g.K = r.Din; g.A = j.X; g.lda = r.ldin; g.B = W(r.iW); g.ldb = ld(r.iW);
    g.C = r.xw; g.ldc = G; g.c_type = act; g.bias = Wf(r.ib);
    gemm(g);
  } else if (j.kind == IN_RANK1) {
    k_rank1_rows(act, r.xw, rows, G, j.X, VD, Wf(r.iW), Wf(r.ib), st);
  } else {
    k_fill_rows(act, r.xw, rows, G, Wf(r.ib), st);
  }
  prof_end();
  prof_begin(PC_REC_FWD);
  if (j.h0) {
    k_copy2d(act, act, n, H, j.h0, j.ld0, r.hseq, H, st);
    if (!use_persist && !gru) k_copy2d(act, act, n, H, j.c0, j.ld0, r.cseq, H, st);
  } else {
    MVAE_CUDA(cudaMemsetAsync(r.hseq, 0, (size_t)n * H * asz(), st));
    if (!use_persist) MVAE_CUDA(cudaMemsetAsync(r.cseq, 0, (size_t)n * H * asz(), st));
  }
  prof_end();
}

RecPersistArgs Model::fwd_args(const FwdJob& j, int n, int slot, int hs, bool pack) {
  Rec& r = *j.r;
  if (pack) {
    if (use_cluster_fwd) rec_cluster_pack_u(Wf(r.iU), ld(r.iU), r.upack, H, r.variant, st);
    else rec_persist_pack_u(Wf(r.iU), ld(r.iU), r.upack, H, hs, r.variant, st);
  }
  RecPersistArgs a;
  a.n = n; a.H = H; a.steps = r.steps; a.gate_act = cfg.gate_act; a.variant = r.variant; a.flags = slot ? rec_flags2 : rec_flags;
  a.upack = r.upack; a.xw = r.xw; a.hseq = r.hseq; a.cseq = r.cseq; a.gates = r.gates; a.c0 = j.c0; a.ldc0 = j.ld0;
  a.hx = slot ? rec_hx2 : rec_hx;
  a.trace = slot ? nullptr : trace_buf;
  a.no_stash = inference_pass ? 1 : 0;
  if (fuse_xproj && r.steps > 8) {
    if ((j.kind == IN_DENSE && j.onehot) || j.kind == IN_NONE) {
      a.x_mode = 1; a.xtab = r.xtab;
      a.x_idx = j.kind == IN_DENSE ? j.idx : nullptr; a.x_ld = j.idx_ld; a.x_shift = j.idx_shift;
    } else if (j.kind == IN_RANK1) {
      a.x_mode = 2; a.x_scalar = j.X; a.x_ld = VD; a.x_w = Wf(r.iW); a.x_b = Wf(r.ib);
    }
  }
  return a;
}

// one or two independent recurrences: projections first, then the time loop (paired in one persistent launch when both
// run the same number of steps and fit on the chip together)
void Model::rec_forward_jobs(const FwdJob* ja, const FwdJob* jb, int n) {
  rec_forward_prepare(*ja, n);
  if (jb) rec_forward_prepare(*jb, n);
  prof_begin(PC_REC_FWD);
  if (use_cluster_fwd) {
    // clusters are independent of each other, so there is nothing to pair: each recurrence is one launch
    for (const FwdJob* j : {ja, jb}) {
      if (!j) continue;
      RecPersistArgs a = fwd_args(*j, n, cur_slot, 0);
      fork_if_pending();
      rec_cluster_forward(a, st);
      dump_trace("fwd(cluster)", *j->r, 2);
    }
  } else if (use_persist) {
    const int hs_pair = (jb && pair_recs && ja->r->steps == jb->r->steps) ? rec_persist_fwd_pair_hs(H, n, sm_count) : 0;
    if (hs_pair) {
      RecPersistArgs a = fwd_args(*ja, n, 0, hs_pair), b = fwd_args(*jb, n, 1, hs_pair);
      rec_persist_forward_pair(a, b, hs_pair, st, sm_count);
      dump_trace("fwd", *ja->r);
    } else {
      for (const FwdJob* j : {ja, jb}) {
        if (!j) continue;
        RecPersistArgs a = fwd_args(*j, n, 0, 0);
        rec_persist_forward(a, st, sm_count);
        dump_trace("fwd", *j->r);
      }
    }
  } else {
    for (const FwdJob* j : {ja, jb}) {
      if (!j) continue;
      if (j->c0 && !gru) k_copy2d(act, DT_F32, n, H, j->c0, j->ld0, c_run, H, st);
      else MVAE_CUDA(cudaMemsetAsync(c_run, 0, (size_t)n * H * 4, st));
      rec_steps_forward(*j->r, n, 0, j->r->steps);
    }
  }
  prof_end();
}

// steps [t0, t1): pre = h_{t-1} U + xw_t ; gate math  (step-streamed form)
void Model::rec_steps_forward(Rec& r, int n, int t0, int t1) {
  if (gru) { gru_steps_forward(r, n, t0, t1); return; }
  for (int t = t0; t < t1; ++t) {
    GemmArgs g; g.M = n; g.N = G; g.K = H; g.A = slab(r.hseq, t, (long)n * H); g.lda = H; g.B = W(r.iU); g.ldb = ld(r.iU);
    g.C = pre; g.ldc = G; g.c_type = DT_F32; g.addend = slab(r.xw, t, (long)n * G); g.ldadd = G; g.add_type = act;
    gemm(g);
    k_cell_fwd(act, cc(r.variant), n, H, pre, c_run, slab(r.gates, t, (long)n * G), slab(r.cseq, t + 1, (long)n * H),
               slab(r.hseq, t + 1, (long)n * H), st);
  }
}

// GRU, steps [t0, t1): two dependent products per step; the reset gate is applied BEFORE the candidate's recurrent product
// (Keras 2.0.8 GRU / recurrentshop GRUCell).  Stash: gates_t = [z|r|hh], r.cseq slab t = r * h_{t-1} (a GRU has no cell state: the buffer is free).
void Model::gru_steps_forward(Rec& r, int n, int t0, int t1) {
  const size_t a = asz();
  if (use_gru_cluster && t1 - t0 >= 2) {   // the whole range in ONE cluster-resident launch (single steps -- the free-running decoder -- stay streamed)
    GruClusterArgs g; g.n = n; g.H = H; g.t0 = t0; g.t1 = t1; g.gate_act = cfg.gate_act; g.mix = r.variant;
    g.U = W(r.iU); g.ldu = ld(r.iU); g.xw = r.xw; g.hseq = r.hseq; g.gates = r.gates; g.rh = r.cseq;
    fork_if_pending();   // the independent (velocity / instrument) recurrences may start on the branch stream now
    gru_cluster_forward(g, st);
    return;
  }
  for (int t = t0; t < t1; ++t) {
    const void* hp = slab(r.hseq, t, (long)n * H);
    const void* xw_t = slab(r.xw, t, (long)n * G);
    void* rh_t = slab(r.cseq, t, (long)n * H);
    { GemmArgs g; g.M = n; g.N = 2 * H; g.K = H; g.A = hp; g.lda = H; g.B = W(r.iU); g.ldb = ld(r.iU);
      g.C = pre; g.ldc = G; g.c_type = DT_F32; g.addend = xw_t; g.ldadd = G; g.add_type = act; gemm(g); }
    k_gru_gates(act, cfg.gate_act, n, H, pre, hp, slab(r.gates, t, (long)n * G), rh_t, st);
    { GemmArgs g; g.M = n; g.N = H; g.K = H; g.A = rh_t; g.lda = H; g.B = (const char*)W(r.iU) + (size_t)2 * H * a; g.ldb = ld(r.iU);
      g.C = pre + 2 * H; g.ldc = G; g.c_type = DT_F32; g.addend = (const char*)xw_t + (size_t)2 * H * a; g.ldadd = G; g.add_type = act; gemm(g); }
    k_gru_out(act, r.variant, n, H, pre, hp, slab(r.gates, t, (long)n * G), slab(r.hseq, t + 1, (long)n * H), st);
  }
}

// GRU reverse-time sweep (hand-derived in oracle/manual_bptt.py, checked against autograd): per step
//   [da_z | . | da_h], direct path  ->  drh = da_h U_h^T  ->  da_r, dh += drh r  ->  dh += [da_z | da_r] U_zr^T
void Model::gru_backward_sweep(const BwdJob& j, int n) {
  Rec& r = *j.r;
  const size_t a = asz();
  void* dG = r.xw;
  if (use_gru_cluster && r.steps >= 2) {
    GruClusterArgs g; g.n = n; g.H = H; g.t0 = 0; g.t1 = r.steps; g.gate_act = cfg.gate_act; g.mix = r.variant;
    g.U = W(r.iU); g.ldu = ld(r.iU); g.hseq = r.hseq; g.gates = r.gates;
    g.dhext = j.use_dhext ? r.dhext : nullptr; g.dh_last = j.dh_last; g.ld_last = j.ld_last; g.dG = dG; g.dS_h = j.dS_h; g.ldS = j.ldS;
    fork_if_pending();
    gru_cluster_backward(g, st);
    return;
  }
  float* drh = c_run;       // (n, H) fp32 scratch: a GRU has no running cell state
  MVAE_CUDA(cudaMemsetAsync(dh_run, 0, (size_t)n * H * 4, st));
  for (int t = r.steps - 1; t >= 0; --t) {
    const void* hp = slab(r.hseq, t, (long)n * H);
    const void* gt = slab(r.gates, t, (long)n * G);
    void* dG_t = slab(dG, t, (long)n * G);
    k_gru_bwd1(act, cfg.gate_act, r.variant, n, H, dh_run, j.use_dhext ? slab(r.dhext, t, (long)n * H) : nullptr, t == r.steps - 1 ? j.dh_last : nullptr,
               j.ld_last, act, gt, hp, dG_t, st);
    { GemmArgs g; g.M = n; g.N = H; g.K = H; g.A = (const char*)dG_t + (size_t)2 * H * a; g.lda = G; g.B = (const char*)W(r.iU) + (size_t)2 * H * a;
      g.ldb = ld(r.iU); g.transB = true; g.C = drh; g.ldc = H; g.c_type = DT_F32; gemm(g); }
    k_gru_bwd2(act, cfg.gate_act, n, H, drh, gt, hp, dh_run, dG_t, st);
    { GemmArgs g; g.M = n; g.N = H; g.K = 2 * H; g.A = dG_t; g.lda = G; g.B = W(r.iU); g.ldb = ld(r.iU); g.transB = true;
      g.C = dh_run; g.ldc = H; g.c_type = DT_F32; g.accumulate = true; gemm(g); }
  }
  if (j.dS_h) k_copy2d(DT_F32, act, n, H, dh_run, H, j.dS_h, j.ldS, st);
}

// --------------------------------------------------------------------------------------------- one recurrence, backward
RecPersistArgs Model::bwd_args(const BwdJob& j, int n, int slot, int hs, bool pack) {
  Rec& r = *j.r;
  RecPersistArgs a;
  a.n = n; a.H = H; a.steps = r.steps; a.gate_act = cfg.gate_act; a.variant = r.variant;
  a.flags = slot ? rec_flags2 : rec_flags;
  a.gates = r.gates; a.cseq = r.cseq; a.u_shadow = W(r.iU); a.ldu = ld(r.iU);
  a.dhext = j.use_dhext ? r.dhext : nullptr; a.dh_last = j.dh_last; a.ld_last = j.ld_last; a.dG = r.xw;
  a.dS_h = j.dS_h; a.dS_c = j.dS_c; a.ldS = j.ldS;
  if (use_cluster_bwd) {
    if (pack) rec_cluster_pack_u_bwd(Wf(r.iU), ld(r.iU), r.upack_b, H, r.variant, st);
    a.upack_bwd = r.upack_b; a.partial = slot ? rec_partial2 : rec_partial;
  } else if (r.upack_b && hs) {
    rec_persist_pack_u_bwd(Wf(r.iU), ld(r.iU), r.upack_b, H, hs, r.variant, st);
    a.upack_bwd = r.upack_b; a.partial = slot ? rec_partial2 : rec_partial;
  }
  a.trace = slot ? nullptr : trace_buf;
  return a;
}

// reverse-time sweep(s): dh_ext -> dG (into r.xw) and the gradients wrt the initial states.  Two independent recurrences
// share one persistent launch when the K-split kernel can pair them (each step is a chain of L2 round trips; the second
// recurrence fills the idle time of the first).
void Model::rec_backward_sweep(const BwdJob* ja, const BwdJob* jb, int n) {
  prof_begin(PC_REC_BWD);
  if (use_cluster_bwd) {
    for (const BwdJob* j : {ja, jb}) {
      if (!j) continue;
      RecPersistArgs a = bwd_args(*j, n, cur_slot, 0);
      fork_if_pending();
      rec_cluster_backward(a, st);
      dump_trace("bwd(cluster)", *j->r, 2);
    }
  } else if (use_persist) {
    // pairing pays only when both recurrences run the same number of steps (a 4-step instrument cell cannot fill the gaps of a T-step one)
    const int hs_pair = (jb && pair_recs && ja->r->steps == jb->r->steps) ? rec_persist_pair_hs(H, n, sm_count) : 0;
    if (hs_pair) {
      RecPersistArgs a = bwd_args(*ja, n, 0, hs_pair), b = bwd_args(*jb, n, 1, hs_pair);
      rec_persist_backward_pair(a, &b, hs_pair, st, sm_count);
      dump_trace("bwd", *ja->r);
    } else {
      for (const BwdJob* j : {ja, jb}) {
        if (!j) continue;
        RecPersistArgs a = bwd_args(*j, n, 0, rec_persist_ksplit_ok(H) ? 16 : 0);
        if (a.upack_bwd) rec_persist_backward_pair(a, nullptr, 16, st, sm_count);
        else rec_persist_backward(a, st, sm_count);
        dump_trace("bwd", *j->r);
      }
    }
  } else {
    for (const BwdJob* j : {ja, jb}) {
      if (!j) continue;
      if (gru) { gru_backward_sweep(*j, n); continue; }
      Rec& r = *j->r;
      void* dG = r.xw;
      MVAE_CUDA(cudaMemsetAsync(dh_run, 0, (size_t)n * H * 4, st));
      MVAE_CUDA(cudaMemsetAsync(dc_run, 0, (size_t)n * H * 4, st));
      for (int t = r.steps - 1; t >= 0; --t) {
        k_cell_bwd(act, cc(r.variant), n, H, dh_run, j->use_dhext ? slab(r.dhext, t, (long)n * H) : nullptr, t == r.steps - 1 ? j->dh_last : nullptr,
                   j->ld_last, act, dc_run, slab(r.gates, t, (long)n * G), slab(r.cseq, t, (long)n * H), slab(r.cseq, t + 1, (long)n * H),
                   slab(dG, t, (long)n * G), st);
        GemmArgs g; g.M = n; g.N = H; g.K = G; g.A = slab(dG, t, (long)n * G); g.lda = G; g.B = W(r.iU); g.ldb = ld(r.iU); g.transB = true;
        g.C = dh_run; g.ldc = H; g.c_type = DT_F32;
        gemm(g);
      }
      if (j->dS_h) {
        k_copy2d(DT_F32, act, n, H, dh_run, H, j->dS_h, j->ldS, st);
        k_copy2d(DT_F32, act, n, H, dc_run, H, j->dS_c, j->ldS, st);
      }
    }
  }
  prof_end();
}

// batched GEMMs after the reverse sweep of one recurrence: dx = dG W^T for the layer below (on the critical path, main stream) and
// the weight gradients dU = Hprev^T dG, dW = X^T dG, db = colsum dG (needed only by the optimizer: side stream when enabled)
void Model::rec_backward_gemms(const BwdJob& j, int n, bool tail) {
  Rec& r = *j.r;
  const long rows = (long)r.steps * n;
  void* dG = r.xw;  // the pre-activation buffer is dead after the forward sweep
  if (use_side) {
    MVAE_CUDA(cudaEventRecord(ev_fork, st));          // dG, hseq of this recurrence are final here
    MVAE_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
  }
  if (j.need_dx) {  // dx = dG W^T first: the layer below is waiting for it
    prof_begin(PC_GEMM);
    GemmArgs g; g.M = (int)rows; g.N = r.Din; g.K = G; g.A = dG; g.lda = G; g.B = W(r.iW); g.ldb = ld(r.iW); g.transB = true;
    g.C = j.dx_out; g.ldc = H; g.c_type = act;
    gemm(g);
    prof_end();
  }
  if (use_side) rec_backward_wgrads(j, n, side, tail ? sm_count : side_sms);   // tail: nothing else is left to run next to them
  else rec_backward_wgrads(j, n, st, sm_count);
}

void Model::rec_backward_wgrads(const BwdJob& j, int n, cudaStream_t s, int sms) {
  // every product accumulates into the zeroed gradient arena
  Rec& r = *j.r;
  const long rows = (long)r.steps * n;
  const void* dG = r.xw;
  const void* Xc = j.X;
  const bool det = prof_detail;
  bool dual_done = false;
  auto seg = [&](const char* tag) { if (det) { prof_end(s); prof_begin(PC_GEMM, s, tag); } };
  prof_begin(PC_GEMM, s, det ? "wgrad dU" : r.name.c_str());
  if (gru) {  // dU_zr += Hprev^T [da_z | da_r];  dU_h += (r * Hprev)^T da_h   (the r * h sequence sits in the cseq buffer)
    GemmArgs g; g.M = H; g.N = 2 * H; g.K = (int)rows; g.A = r.hseq; g.lda = H; g.transA = true; g.B = dG; g.ldb = G;
    g.C = Gp(r.iU); g.ldc = ld(r.iU); g.c_type = DT_F32; g.accumulate = true;
    gemm_on(g, s, sms);
    GemmArgs h; h.M = H; h.N = H; h.K = (int)rows; h.A = r.cseq; h.lda = H; h.transA = true;
    h.B = (const char*)dG + (size_t)2 * H * asz(); h.ldb = G; h.C = Gp(r.iU) + 2 * H; h.ldc = ld(r.iU); h.c_type = DT_F32; h.accumulate = true;
    gemm_on(h, s, sms);
  } else {  // dU += Hprev^T dG
    GemmArgs g; g.M = H; g.N = G; g.K = (int)rows; g.A = r.hseq; g.lda = H; g.transA = true; g.B = dG; g.ldb = G;
    g.C = Gp(r.iU); g.ldc = ld(r.iU); g.c_type = DT_F32; g.accumulate = true;
    if (fuse_dual_wgrad && j.kind == IN_DENSE && act == DT_BF16) {
      // ... and dW += X^T dG in the same launch: both products stream the same dG, which is then read from DRAM once (K-split-major tile order).
      // (Letting the bias gradient ride along as a third, one-row product against a column of ones was measured too: the padded 128-row tile costs
      // more GEMM time on the weight-gradient stream than the column-sum launch it removes.)
      g.A2 = Xc; g.lda2 = r.ldin; g.M2 = r.Din; g.C2 = Gp(r.iW); g.ldc2 = ld(r.iW);
      g.in_type = act;
      if (!gemm_tc_supported(g)) { g.A2 = nullptr; g.M2 = 0; g.C2 = nullptr; } else dual_done = true;
    }
    gemm_on(g, s, sms);
  }
  const bool rows_ok = fuse_wgrad_rows && act == DT_BF16 && !gru && G % 8 == 0;
  if (rows_ok && j.kind == IN_RANK1) {                  // scalar input: dW (weighted column sum) and db in ONE pass
    seg("wgrad dW(rank1)+db rows");
    k_wgrad_rows(rows, G, G, dG, Xc, VD, Gp(r.iW), Gp(r.ib), s);
    prof_end(s);
    return;
  }
  if (j.kind == IN_DENSE && !dual_done) {  // dW += X^T dG
    seg("wgrad dW");
    GemmArgs g; g.M = r.Din; g.N = G; g.K = (int)rows; g.A = Xc; g.lda = r.ldin; g.transA = true; g.B = dG; g.ldb = G;
    g.C = Gp(r.iW); g.ldc = ld(r.iW); g.c_type = DT_F32; g.accumulate = true;
    gemm_on(g, s, sms);
  } else if (j.kind == IN_RANK1) {
    seg("wgrad dW rank1 colsum");
    k_colsum(act, rows, G, G, dG, Xc, VD, Gp(r.iW), s);
  }
  seg("wgrad db colsum");
  k_colsum(act, rows, G, G, dG, nullptr, 0, Gp(r.ib), s);
  prof_end(s);
}

// a stack of layers (top first) plus independent side recurrences: pair the i-th stack layer with the i-th side recurrence
void Model::rec_backward_group(std::vector<BwdJob>& stack, std::vector<BwdJob>& side, int n, bool last_group) {
  if (use_branch) {
    // the stack (top layer first) is the critical chain; the independent recurrences run next to it on the branch stream
    branch_fork();
    rec_backward_sweep(&stack[0], nullptr, n);
    rec_backward_gemms(stack[0], n);
    branch_begin();
    // last group: by the time the side stream reaches the branch recurrences' weight gradients only the tail of the step is left: whole chip
    for (auto& j : side) { rec_backward_sweep(&j, nullptr, n); rec_backward_gemms(j, n, last_group); }
    branch_end();
    for (size_t k = 1; k < stack.size(); ++k) { rec_backward_sweep(&stack[k], nullptr, n); rec_backward_gemms(stack[k], n, last_group && k + 1 == stack.size()); }
    branch_join();
    return;
  }
  size_t si = 0;
  for (size_t k = 0; k < stack.size(); ++k) {
    const BwdJob* partner = si < side.size() ? &side[si] : nullptr;
    rec_backward_sweep(&stack[k], partner, n);
    rec_backward_gemms(stack[k], n);          // produces dhext of stack[k+1]
    if (partner) { rec_backward_gemms(*partner, n); ++si; }
  }
  for (; si < side.size(); si += 2) {
    const BwdJob* partner = si + 1 < side.size() ? &side[si + 1] : nullptr;
    rec_backward_sweep(&side[si], partner, n);
    rec_backward_gemms(side[si], n);
    if (partner) rec_backward_gemms(*partner, n);
  }
}

// --------------------------------------------------------------------------------------------- encoder (vae_definition.py:443-516)
void Model::encoder_forward(int n) {
  // the first pitch layer and the velocity stream are independent and equally long: they share a launch
  FwdJob jv; jv.r = &enc_vel; jv.kind = IN_RANK1; jv.X = slab(Xv_ext, 1, (long)n * VD);
  if (use_branch) {
    branch_fork();
    for (int k = 0; k < ne; ++k) {
      FwdJob jp; jp.r = &enc_pitch[k]; jp.kind = IN_DENSE;
      jp.X = k == 0 ? slab(Xp_ext, 1, (long)n * PD) : slab(enc_pitch[k - 1].hseq, 1, (long)n * H);
      if (k == 0) { jp.onehot = true; jp.idx = cur_pitch; jp.idx_ld = T; jp.idx_shift = 0; }
      rec_forward_jobs(&jp, nullptr, n);
      if (k == 0) {   // velocity and instrument streams: next to the pitch stack
        branch_begin();
        rec_forward_jobs(&jv, nullptr, n);
        rec_forward(enc_instr, n, IN_DENSE, slab(Xi_ext, 1, (long)n * ID), nullptr, nullptr, 0);
        branch_end();
      }
    }
    branch_join();
    return;
  }
  for (int k = 0; k < ne; ++k) {
    FwdJob jp; jp.r = &enc_pitch[k]; jp.kind = IN_DENSE;
    jp.X = k == 0 ? slab(Xp_ext, 1, (long)n * PD) : slab(enc_pitch[k - 1].hseq, 1, (long)n * H);
    rec_forward_jobs(&jp, k == 0 ? &jv : nullptr, n);
  }
  rec_forward(enc_instr, n, IN_DENSE, slab(Xi_ext, 1, (long)n * ID), nullptr, nullptr, 0);
}

// concat -> Dense(tanh) -> Dense(tanh) -> split -> z_mean / z_log_var -> KL -> z (vae_definition.py:468-515, 29-37)
void Model::head_forward(const mvae_batch& b, bool with_style_loss) {
  const int n = b.n;
  prof_begin(PC_POINTWISE);
  k_concat3(act, n, H, slab(enc_pitch[ne - 1].hseq, T, (long)n * H), slab(enc_instr.hseq, Ti, (long)n * H), slab(enc_vel.hseq, T, (long)n * H), u, st);
  prof_end();
  prof_begin(PC_GEMM);
  { GemmArgs g; g.M = n; g.N = H; g.K = 3 * H; g.A = u; g.lda = 3 * H; g.B = W(iWa); g.ldb = ld(iWa); g.C = a1; g.ldc = H; g.c_type = act;
    g.bias = Wf(iba); g.act = 1; gemm(g); }
  void* ecur = a1;
  if (cfg.extra_layer) {
    GemmArgs g; g.M = n; g.N = H; g.K = H; g.A = a1; g.lda = H; g.B = W(iWe); g.ldb = ld(iWe); g.C = e; g.ldc = H; g.c_type = act;
    g.bias = Wf(ibe); g.act = 1; gemm(g);
    ecur = e;
  }
  e_cur = ecur;
  { GemmArgs g; g.M = n; g.N = L; g.K = half; g.A = ecur; g.lda = H; g.B = W(iWmu); g.ldb = ld(iWmu); g.C = mu; g.ldc = ldl; g.c_type = DT_F32;
    g.bias = Wf(ibmu); gemm(g); }
  { GemmArgs g; g.M = n; g.N = L; g.K = H - half; g.A = (const char*)ecur + (size_t)half * asz(); g.lda = H; g.B = W(iWlv); g.ldb = ld(iWlv);
    g.C = lv; g.ldc = ldl; g.c_type = DT_F32; g.bias = Wf(iblv); gemm(g); }
  prof_end();
  prof_begin(PC_POINTWISE);
  k_latent_fwd(act, n, L, ldl, mu, lv, b.eps, b.history, cfg.history, z, q, ldq, cfg.beta, cfg.prior_mean, cfg.prior_std, acc, st);
  if (history_mode == 1 && cfg.history && with_style_loss) {   // train / evaluate only: the history is this batch's own z, a constant of the step
    k_self_history(act, n, L, ldl, z, song_start_set ? d_song_start : nullptr, hist_carry, carry_valid ? 1 : 0, q, ldq, st);
    MVAE_CUDA(cudaMemcpyAsync(hist_carry, z + (size_t)(n - 1) * ldl, (size_t)L * 4, cudaMemcpyDeviceToDevice, st));
    carry_valid = true;
  }
  k_style_head(n, C, z, ldl, with_style_loss ? b.style : nullptr, style_probs, acc, st);
  prof_end();
}

// decoder (vae_definition.py:519-645): q = [z | history] must already be in this->q
void Model::decoder_forward(const mvae_batch& b, int feedback) {
  const int n = b.n;
  prof_begin(PC_GEMM);
  { GemmArgs g; g.M = n; g.N = nS * H; g.K = Q; g.A = q; g.lda = ldq; g.B = W(iWinit); g.ldb = ld(iWinit); g.C = S; g.ldc = nS * H; g.c_type = act;
    g.bias = Wf(ibinit); g.act = 1; gemm(g); }
  prof_end();
  if (feedback == MVAE_FB_FREE_RUNNING) { decoder_stepwise(n); return; }
  const bool tf = feedback == MVAE_FB_TEACHER_FORCED;
  auto st1 = [&](int r) { return (const char*)S + (size_t)(spc * r) * H * asz(); };
  auto st2 = [&](int r) { return gru ? (const char*)nullptr : (const char*)S + (size_t)(2 * r + 1) * H * asz(); };
  FwdJob jv; jv.r = &dec_vel; jv.kind = tf ? IN_RANK1 : IN_NONE; jv.X = tf ? Xv_ext : nullptr; jv.h0 = st1(nd + 1); jv.c0 = st2(nd + 1); jv.ld0 = nS * H;
  if (use_branch) branch_fork();
  for (int k = 0; k < nd; ++k) {
    FwdJob jp; jp.r = &dec_notes[k]; jp.h0 = st1(k); jp.c0 = st2(k); jp.ld0 = nS * H;
    if (k == 0) { jp.kind = tf ? IN_DENSE : IN_NONE; jp.X = tf ? Y_ext_cur : nullptr; }
    else { jp.kind = IN_DENSE; jp.X = slab(dec_notes[k - 1].hseq, 1, (long)n * H); }
    if (k == 0 && tf) { jp.onehot = true; jp.idx = cur_target; jp.idx_ld = T; jp.idx_shift = 1; }   // x_t = y_{t-1}, x_0 = 0
    if (use_branch) {
      rec_forward_jobs(&jp, nullptr, n);
      if (k == 0) {
        branch_begin();
        rec_forward_jobs(&jv, nullptr, n);
        rec_forward(dec_instr, n, tf ? IN_DENSE : IN_NONE, tf ? Xi_ext : nullptr, st1(nd), st2(nd), nS * H);
        branch_end();
      }
    } else {
      rec_forward_jobs(&jp, k == 0 ? &jv : nullptr, n);
    }
  }
  if (use_branch) branch_join();
  else rec_forward(dec_instr, n, tf ? IN_DENSE : IN_NONE, tf ? Xi_ext : nullptr, st1(nd), st2(nd), nS * H);
  prof_begin(PC_GEMM);
  { GemmArgs g; g.M = T * n; g.N = Dp; g.K = H; g.A = slab(dec_notes[nd - 1].hseq, 1, (long)n * H); g.lda = H; g.B = W(iWy); g.ldb = ld(iWy);
    g.C = Pn; g.ldc = ld_pn; g.c_type = DT_F32; g.bias = Wf(iby); gemm(g); }
  { GemmArgs g; g.M = Ti * n; g.N = Di; g.K = H; g.A = slab(dec_instr.hseq, 1, (long)n * H); g.lda = H; g.B = W(iWio); g.ldb = ld(iWio);
    g.C = Pi; g.ldc = ld_pi; g.c_type = DT_F32; g.bias = Wf(ibio); gemm(g); }
  k_rowdot(act, (long)T * n, H, slab(dec_vel.hseq, 1, (long)n * H), Wf(iWvo), Wf(ibvo), Pv, ld_pv, st);
  prof_end();
}

// --------------------------------------------------------------------------------------------- style classifier (model_kind = 1)
// pitch_classifier.py:89-103 / velocity_classifier.py:110-118 / instrument_classifier.py:93-103: stacked recurrent layers over the (T, D) input,
// last state of the top layer -> Dense(C, softmax); loss = mean categorical cross-entropy, metric = accuracy (:100-101)
void Model::cls_forward(const mvae_batch& b, bool train) {
  const int n = b.n;
  MVAE_REQUIRE(n >= 1 && n <= NB, "mini-batch size must be in 1..max_batch");
  MVAE_REQUIRE(cls_scalar ? b.velocity != nullptr : b.pitch != nullptr, "classifier input roll is missing (pitch: class indices, or velocity: scalars)");
  MVAE_REQUIRE(!train || b.style, "classifier labels (batch.style) are required for training");
  MVAE_CUDA(cudaMemsetAsync(acc, 0, ACC_COUNT * sizeof(double), st));
  prof_begin(PC_POINTWISE);
  k_expand_inputs(act, n, T, Ti, PD, ID, VD, b.pitch, nullptr, nullptr, b.velocity, cls_scalar ? nullptr : Xp_ext, nullptr, nullptr, cls_scalar ? Xv_ext : nullptr, st);
  cur_pitch = b.pitch;
  prof_end();
  for (int k = 0; k < ne; ++k) {
    FwdJob jp; jp.r = &enc_pitch[k];
    if (k == 0) {
      jp.kind = cls_scalar ? IN_RANK1 : IN_DENSE;
      jp.X = cls_scalar ? slab(Xv_ext, 1, (long)n * VD) : slab(Xp_ext, 1, (long)n * PD);
      if (!cls_scalar) { jp.onehot = true; jp.idx = cur_pitch; jp.idx_ld = T; jp.idx_shift = 0; }
    } else {
      jp.kind = IN_DENSE; jp.X = slab(enc_pitch[k - 1].hseq, 1, (long)n * H);
    }
    rec_forward_jobs(&jp, nullptr, n);
  }
  prof_begin(PC_GEMM);
  { GemmArgs g; g.M = n; g.N = C; g.K = H; g.A = slab(enc_pitch[ne - 1].hseq, T, (long)n * H); g.lda = H; g.B = W(iWy); g.ldb = ld(iWy);
    g.C = Pn; g.ldc = ld_pn; g.c_type = DT_F32; g.bias = Wf(iby); gemm(g); }
  prof_end();
  prof_begin(PC_POINTWISE);
  k_softmax_ce(act, 1, n, C, Pn, ld_pn, b.style, nullptr, nullptr, 1.f, train ? dlog_n : nullptr, ld_pn, acc, ACC_CE_STYLE, ACC_ACC_STYLE, st);
  k_finalize_metrics(acc, n, T, Ti, 0.f, 0.f, 0.f, 1.f, d_metrics, st);
  prof_end();
}

void Model::cls_backward(const mvae_batch& b) {
  const int n = b.n;
  MVAE_CUDA(cudaMemsetAsync(Gr, 0, arena_n * 4, st));
  const void* hT = slab(enc_pitch[ne - 1].hseq, T, (long)n * H);
  prof_begin(PC_GEMM);
  { GemmArgs g; g.M = H; g.N = C; g.K = n; g.A = hT; g.lda = H; g.transA = true; g.B = dlog_n; g.ldb = ld_pn; g.C = Gp(iWy); g.ldc = ld(iWy);
    g.c_type = DT_F32; g.accumulate = true; gemm(g); }
  k_colsum(act, n, C, ld_pn, dlog_n, nullptr, 0, Gp(iby), st);
  { GemmArgs g; g.M = n; g.N = H; g.K = C; g.A = dlog_n; g.lda = ld_pn; g.B = W(iWy); g.ldb = ld(iWy); g.transB = true; g.C = du; g.ldc = H;
    g.c_type = act; gemm(g); }
  prof_end();
  std::vector<BwdJob> stack, none;
  for (int k = ne - 1; k >= 0; --k) {
    const bool top = k == ne - 1;
    BwdJob j; j.r = &enc_pitch[k];
    j.kind = (k == 0 && cls_scalar) ? IN_RANK1 : IN_DENSE;
    j.X = k == 0 ? (cls_scalar ? slab(Xv_ext, 1, (long)n * VD) : slab(Xp_ext, 1, (long)n * PD)) : slab(enc_pitch[k - 1].hseq, 1, (long)n * H);
    j.use_dhext = !top; j.dh_last = top ? du : nullptr; j.ld_last = H;
    j.need_dx = k > 0; j.dx_out = k > 0 ? enc_pitch[k - 1].dhext : nullptr;
    stack.push_back(j);
  }
  rec_backward_group(stack, none, n, true);
  if (use_side) {
    MVAE_CUDA(cudaEventRecord(ev_join, this->side));
    MVAE_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
  }
}

// free-running decode (inference only): x_t = previous prediction, one step at a time
void Model::decoder_stepwise(int n) {
  prof_begin(PC_REC_FWD);
  StepwiseGraph& sg = stepwise_graphs[n];
  if (!stepwise_graph_on) {
    decoder_stepwise_body(n);
  } else if (sg.state == 0) {
    decoder_stepwise_body(n);          // first call for this batch size: eager (kernel attributes get set outside any capture)
    sg.state = 1;
  } else if (sg.state == 1) {
    cudaGraph_t graph = nullptr;
    const long long l0 = g_launches;
    MVAE_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
    try {
      decoder_stepwise_body(n);
    } catch (...) {
      cudaStreamEndCapture(st, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    MVAE_CUDA(cudaStreamEndCapture(st, &graph));
    sg.launches = g_launches - l0;
    MVAE_CUDA(cudaGraphInstantiate(&sg.exec, graph, 0));
    cudaGraphDestroy(graph);
    sg.state = 2;
    MVAE_CUDA(cudaGraphLaunch(sg.exec, st));
  } else {
    MVAE_CUDA(cudaGraphLaunch(sg.exec, st));
    count_launch((int)sg.launches);
  }
  prof_end();
  stepwise_done = true;
}

void Model::decoder_stepwise_body(int n) {
  auto st1 = [&](int r) { return (const char*)S + (size_t)(spc * r) * H * asz(); };
  auto st2 = [&](int r) { return gru ? (const char*)nullptr : (const char*)S + (size_t)(2 * r + 1) * H * asz(); };
  auto init = [&](Rec& r, int sidx) {
    k_copy2d(act, act, n, H, st1(sidx), nS * H, r.hseq, H, st);
    if (!gru) k_copy2d(act, act, n, H, st2(sidx), nS * H, r.cseq, H, st);
  };
  auto step = [&](Rec& r, int t, int kind, const void* x, int ldx) {  // one cell step with input x (n rows) or none (t == 0 start vector = 0)
    void* xw_t = slab(r.xw, t, (long)n * G);
    if (kind == IN_DENSE) {
      GemmArgs g; g.M = n; g.N = G; g.K = r.Din; g.A = x; g.lda = ldx; g.B = W(r.iW); g.ldb = ld(r.iW); g.C = xw_t; g.ldc = G; g.c_type = act;
      g.bias = Wf(r.ib); gemm(g);
    } else if (kind == IN_RANK1) {
      k_rank1_rows(act, xw_t, n, G, x, ldx, Wf(r.iW), Wf(r.ib), st);
    } else {
      k_fill_rows(act, xw_t, n, G, Wf(r.ib), st);
    }
    if (!gru) k_copy2d(act, DT_F32, n, H, slab(r.cseq, t, (long)n * H), H, c_run, H, st);
    rec_steps_forward(r, n, t, t + 1);
  };
  // the velocity and instrument chains are independent of the notes chain: branch stream, own scratch (host code stays sequential)
  MVAE_CUDA(cudaEventRecord(ev_bfork, st));
  // notes
  for (int k = 0; k < nd; ++k) init(dec_notes[k], k);
  for (int t = 0; t < T; ++t) {
    if (t == 0) step(dec_notes[0], 0, IN_NONE, nullptr, 0);
    else {
      k_copy2d(DT_F32, act, n, PD, Pn + (size_t)(t - 1) * n * ld_pn, ld_pn, xstep, PD, st);
      step(dec_notes[0], t, IN_DENSE, xstep, PD);
    }
    for (int k = 1; k < nd; ++k) step(dec_notes[k], t, IN_DENSE, slab(dec_notes[k - 1].hseq, t + 1, (long)n * H), H);
    float* Pt = Pn + (size_t)t * n * ld_pn;
    GemmArgs g; g.M = n; g.N = Dp; g.K = H; g.A = slab(dec_notes[nd - 1].hseq, t + 1, (long)n * H); g.lda = H; g.B = W(iWy); g.ldb = ld(iWy);
    g.C = Pt; g.ldc = ld_pn; g.c_type = DT_F32; g.bias = Wf(iby); gemm(g);
    k_softmax_ce(act, 1, n, Dp, Pt, ld_pn, nullptr, nullptr, nullptr, 0.f, nullptr, ld_pn, acc, ACC_CE_NOTES, ACC_ACC_NOTES, st);
  }
  MVAE_CUDA(cudaStreamWaitEvent(st_branch, ev_bfork, 0));
  cudaStream_t st_main = st;
  st = st_branch;
  std::swap(pre, pre_b); std::swap(c_run, c_run_b); std::swap(xstep, xstep_b);
  struct Restore { Model& m; cudaStream_t s; ~Restore() { m.st = s; std::swap(m.pre, m.pre_b); std::swap(m.c_run, m.c_run_b); std::swap(m.xstep, m.xstep_b); } } restore{*this, st_main};
  // instrument
  init(dec_instr, nd);
  for (int t = 0; t < Ti; ++t) {
    if (t == 0) step(dec_instr, 0, IN_NONE, nullptr, 0);
    else {
      k_copy2d(DT_F32, act, n, ID, Pi + (size_t)(t - 1) * n * ld_pi, ld_pi, xstep, ID, st);
      step(dec_instr, t, IN_DENSE, xstep, ID);
    }
    float* Pt = Pi + (size_t)t * n * ld_pi;
    GemmArgs g; g.M = n; g.N = Di; g.K = H; g.A = slab(dec_instr.hseq, t + 1, (long)n * H); g.lda = H; g.B = W(iWio); g.ldb = ld(iWio);
    g.C = Pt; g.ldc = ld_pi; g.c_type = DT_F32; g.bias = Wf(ibio); gemm(g);
    k_softmax_ce(act, 1, n, Di, Pt, ld_pi, nullptr, nullptr, nullptr, 0.f, nullptr, ld_pi, acc, ACC_CE_INSTR, ACC_ACC_INSTR, st);
  }
  // velocity
  init(dec_vel, nd + 1);
  for (int t = 0; t < T; ++t) {
    if (t == 0) step(dec_vel, 0, IN_NONE, nullptr, 0);
    else {
      k_copy2d(DT_F32, act, n, 1, Pv + (size_t)(t - 1) * n * ld_pv, ld_pv, xstep, VD, st);
      step(dec_vel, t, IN_RANK1, xstep, VD);
    }
    float* Pt = Pv + (size_t)t * n * ld_pv;
    k_rowdot(act, n, H, slab(dec_vel.hseq, t + 1, (long)n * H), Wf(iWvo), Wf(ibvo), Pt, ld_pv, st);
    k_sigmoid_mse(act, 1, n, Pt, ld_pv, nullptr, 0.f, nullptr, ld_pv, acc, st);
  }
  MVAE_CUDA(cudaEventRecord(ev_bjoin, st_branch));
  MVAE_CUDA(cudaStreamWaitEvent(st_main, ev_bjoin, 0));
}

// softmax / sigmoid heads + Keras losses; train => also the gradients wrt the logits
void Model::losses(const mvae_batch& b, bool train) {
  const int n = b.n;
  prof_begin(PC_POINTWISE);
  k_count_nonzero(b.w_notes, (long)n * T, acc, st);
  const uint8_t* tgt = b.target ? b.target : b.pitch;
  k_softmax_ce(act, T, n, Dp, Pn, ld_pn, tgt, b.w_notes, acc + ACC_WNZ, cfg.notes_weight, train ? dlog_n : nullptr, ld_pn, acc, ACC_CE_NOTES,
               ACC_ACC_NOTES, st);
  k_softmax_ce(act, Ti, n, Di, Pi, ld_pi, b.instr, nullptr, nullptr, cfg.meta_instrument_weight, train ? dlog_i : nullptr, ld_pi, acc, ACC_CE_INSTR,
               ACC_ACC_INSTR, st);
  k_sigmoid_mse(act, T, n, Pv, ld_pv, b.velocity, cfg.meta_velocity_weight, train ? dlog_v : nullptr, ld_pv, acc, st);
  k_finalize_metrics(acc, n, T, Ti, cfg.notes_weight, cfg.meta_instrument_weight, cfg.meta_velocity_weight, cfg.composer_weight, d_metrics, st);
  prof_end();
}

// --------------------------------------------------------------------------------------------- backward
void Model::backward(const mvae_batch& b) {
  const int n = b.n;
  const bool tf = cfg.decoder_feedback == MVAE_FB_TEACHER_FORCED;
  MVAE_CUDA(cudaMemsetAsync(Gr, 0, arena_n * 4, st));
  if (use_side) {   // the side stream accumulates into the freshly zeroed gradient arena
    MVAE_CUDA(cudaEventRecord(ev_fork, st));
    MVAE_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
  }
  // weight gradients are needed only by the optimizer: they go to the side stream (when enabled), the activations' gradients stay on the
  // critical chain.  fork_side(): everything issued so far on the main stream is visible to the side stream.
  cudaStream_t ws = use_side ? side : st;
  const int wsms = use_side ? side_sms : sm_count;
  auto fork_side = [&]() {
    if (!use_side) return;
    MVAE_CUDA(cudaEventRecord(ev_fork, st));
    MVAE_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
  };
  // ---- output heads
  Rec& top = dec_notes[nd - 1];
  auto hseg = [&](const char* tag) { if (prof_detail) { prof_end(ws); prof_begin(PC_GEMM, ws, tag); } };
  prof_begin(PC_GEMM, ws, prof_detail ? "heads dWy gemm" : "heads wgrad");
  { GemmArgs g; g.M = H; g.N = Dp; g.K = T * n; g.A = slab(top.hseq, 1, (long)n * H); g.lda = H; g.transA = true; g.B = dlog_n; g.ldb = ld_pn;
    g.C = Gp(iWy); g.ldc = ld(iWy); g.c_type = DT_F32; g.accumulate = true; gemm_on(g, ws, wsms); }
  hseg("heads dby colsum");
  k_colsum(act, (long)T * n, Dp, ld_pn, dlog_n, nullptr, 0, Gp(iby), ws);
  hseg("heads instr gemm + colsum");
  { GemmArgs g; g.M = H; g.N = Di; g.K = Ti * n; g.A = slab(dec_instr.hseq, 1, (long)n * H); g.lda = H; g.transA = true; g.B = dlog_i; g.ldb = ld_pi;
    g.C = Gp(iWio); g.ldc = ld(iWio); g.c_type = DT_F32; g.accumulate = true; gemm_on(g, ws, wsms); }
  k_colsum(act, (long)Ti * n, Di, ld_pi, dlog_i, nullptr, 0, Gp(ibio), ws);
  // velocity head (N = 1): rank-1 forms instead of GEMMs
  hseg("heads vel weighted colsum + bias");
  // dWv[j] = sum_r dlog[r] h[r,j]; the bias gradient (sum_r dlog[r]) rides along instead of taking a launch of its own (0.22 ms for 256 KB)
  k_colsum(act, (long)T * n, H, H, slab(dec_vel.hseq, 1, (long)n * H), dlog_v, ld_pv, Gp(iWvo), ws, Gp(ibvo));
  prof_end(ws);
  prof_begin(PC_GEMM);
  { GemmArgs g; g.M = T * n; g.N = H; g.K = Dp; g.A = dlog_n; g.lda = ld_pn; g.B = W(iWy); g.ldb = ld(iWy); g.transB = true;
    g.C = top.dhext; g.ldc = H; g.c_type = act; gemm(g); }
  { GemmArgs g; g.M = Ti * n; g.N = H; g.K = Di; g.A = dlog_i; g.lda = ld_pi; g.B = W(iWio); g.ldb = ld(iWio); g.transB = true;
    g.C = dec_instr.dhext; g.ldc = H; g.c_type = act; gemm(g); }
  prof_end();
  prof_begin(PC_POINTWISE);
  k_rank1_rows(act, dec_vel.dhext, (long)T * n, H, dlog_v, ld_pv, Wf(iWvo), nullptr, st);              // dh[r,:] = dlog[r] * Wv
  prof_end();
  // ---- decoder recurrences (top layer first)
  auto dS1 = [&](int r) { return (char*)dS + (size_t)(spc * r) * H * asz(); };
  auto dS2 = [&](int r) { return gru ? (char*)nullptr : (char*)dS + (size_t)(2 * r + 1) * H * asz(); };
  {
    std::vector<BwdJob> stack, side;
    for (int k = nd - 1; k >= 0; --k) {
      BwdJob j; j.r = &dec_notes[k];
      j.X = k == 0 ? (tf ? Y_ext_cur : nullptr) : slab(dec_notes[k - 1].hseq, 1, (long)n * H);
      j.kind = k == 0 ? (tf ? IN_DENSE : IN_NONE) : IN_DENSE;
      j.use_dhext = true; j.need_dx = k > 0; j.dx_out = k > 0 ? dec_notes[k - 1].dhext : nullptr;
      j.dS_h = dS1(k); j.dS_c = dS2(k); j.ldS = nS * H;
      stack.push_back(j);
    }
    { BwdJob j; j.r = &dec_vel; j.kind = tf ? IN_RANK1 : IN_NONE; j.X = tf ? Xv_ext : nullptr; j.use_dhext = true;
      j.dS_h = dS1(nd + 1); j.dS_c = dS2(nd + 1); j.ldS = nS * H; side.push_back(j); }
    { BwdJob j; j.r = &dec_instr; j.kind = tf ? IN_DENSE : IN_NONE; j.X = tf ? Xi_ext : nullptr; j.use_dhext = true;
      j.dS_h = dS1(nd); j.dS_c = dS2(nd); j.ldS = nS * H; side.push_back(j); }
    rec_backward_group(stack, side, n);
  }
  // ---- initial-state Denses -> dq
  prof_begin(PC_POINTWISE);
  k_tanh_bwd(act, (long)n * nS * H, dS, S, dSpre, st);
  prof_end();
  fork_side();
  prof_begin(PC_GEMM, ws);
  { GemmArgs g; g.M = Q; g.N = nS * H; g.K = n; g.A = q; g.lda = ldq; g.transA = true; g.B = dSpre; g.ldb = nS * H;
    g.C = Gp(iWinit); g.ldc = ld(iWinit); g.c_type = DT_F32; g.accumulate = true; gemm_on(g, ws, wsms); }
  k_colsum(act, n, nS * H, nS * H, dSpre, nullptr, 0, Gp(ibinit), ws);
  prof_end(ws);
  prof_begin(PC_GEMM);
  { GemmArgs g; g.M = n; g.N = Q; g.K = nS * H; g.A = dSpre; g.lda = nS * H; g.B = W(iWinit); g.ldb = ld(iWinit); g.transB = true;
    g.C = dq; g.ldc = ldq; g.c_type = act; gemm(g); }
  prof_end();
  // every decoder-side weight gradient (arena tail from dec_init on) has been enqueued: its bucket of the data-parallel all-reduce
  // goes to the communication stream now and overlaps the encoder's reverse sweeps
  if (overlap_allreduce && world > 1 && nccl_comm && ar_buckets > 1) {
    MVAE_CUDA(cudaEventRecord(ev_dec_grads, ws));
    MVAE_CUDA(cudaStreamWaitEvent(st_comm, ev_dec_grads, 0));
    const size_t off = ptab[iWinit].off;
    prof_begin(PC_ALLREDUCE, st_comm);
    nccl_allreduce_sum_f32(nccl_comm, Gr + off, arena_n - off, st_comm);
    prof_end(st_comm);
    dec_bucket_issued = true;
  }
  // ---- style head + reparameterisation + KL
  prof_begin(PC_POINTWISE);
  k_latent_bwd(act, n, L, ldl, C, dq, ldq, mu, lv, b.eps, style_probs, b.style, cfg.beta, cfg.prior_mean, cfg.prior_std, cfg.composer_weight, dmu,
               dlv, st);
  prof_end();
  // ---- latent head Denses
  const void* e1 = e_cur;
  const void* e2 = (const char*)e_cur + (size_t)half * asz();
  fork_side();
  prof_begin(PC_GEMM, ws);
  { GemmArgs g; g.M = half; g.N = L; g.K = n; g.A = e1; g.lda = H; g.transA = true; g.B = dmu; g.ldb = ldl; g.C = Gp(iWmu); g.ldc = ld(iWmu);
    g.c_type = DT_F32; g.accumulate = true; gemm_on(g, ws, wsms); }
  k_colsum(act, n, L, ldl, dmu, nullptr, 0, Gp(ibmu), ws);
  { GemmArgs g; g.M = H - half; g.N = L; g.K = n; g.A = e2; g.lda = H; g.transA = true; g.B = dlv; g.ldb = ldl; g.C = Gp(iWlv); g.ldc = ld(iWlv);
    g.c_type = DT_F32; g.accumulate = true; gemm_on(g, ws, wsms); }
  k_colsum(act, n, L, ldl, dlv, nullptr, 0, Gp(iblv), ws);
  prof_end(ws);
  prof_begin(PC_GEMM);
  { GemmArgs g; g.M = n; g.N = half; g.K = L; g.A = dmu; g.lda = ldl; g.B = W(iWmu); g.ldb = ld(iWmu); g.transB = true; g.C = de; g.ldc = H;
    g.c_type = act; gemm(g); }
  { GemmArgs g; g.M = n; g.N = H - half; g.K = L; g.A = dlv; g.lda = ldl; g.B = W(iWlv); g.ldb = ld(iWlv); g.transB = true;
    g.C = (char*)de + (size_t)half * asz(); g.ldc = H; g.c_type = act; gemm(g); }
  prof_end();
  const void* d_a1 = de;
  if (cfg.extra_layer) {
    prof_begin(PC_POINTWISE);
    k_tanh_bwd(act, (long)n * H, de, e, dpre_e, st);
    prof_end();
    fork_side();
    prof_begin(PC_GEMM, ws);
    { GemmArgs g; g.M = H; g.N = H; g.K = n; g.A = a1; g.lda = H; g.transA = true; g.B = dpre_e; g.ldb = H; g.C = Gp(iWe); g.ldc = ld(iWe);
      g.c_type = DT_F32; g.accumulate = true; gemm_on(g, ws, wsms); }
    k_colsum(act, n, H, H, dpre_e, nullptr, 0, Gp(ibe), ws);
    prof_end(ws);
    prof_begin(PC_GEMM);
    { GemmArgs g; g.M = n; g.N = H; g.K = H; g.A = dpre_e; g.lda = H; g.B = W(iWe); g.ldb = ld(iWe); g.transB = true; g.C = da1; g.ldc = H;
      g.c_type = act; gemm(g); }
    prof_end();
    d_a1 = da1;
  }
  prof_begin(PC_POINTWISE);
  k_tanh_bwd(act, (long)n * H, d_a1, a1, dpre_a, st);
  prof_end();
  fork_side();
  prof_begin(PC_GEMM, ws);
  { GemmArgs g; g.M = 3 * H; g.N = H; g.K = n; g.A = u; g.lda = 3 * H; g.transA = true; g.B = dpre_a; g.ldb = H; g.C = Gp(iWa); g.ldc = ld(iWa);
    g.c_type = DT_F32; g.accumulate = true; gemm_on(g, ws, wsms); }
  k_colsum(act, n, H, H, dpre_a, nullptr, 0, Gp(iba), ws);
  prof_end(ws);
  prof_begin(PC_GEMM);
  { GemmArgs g; g.M = n; g.N = 3 * H; g.K = H; g.A = dpre_a; g.lda = H; g.B = W(iWa); g.ldb = ld(iWa); g.transB = true; g.C = du; g.ldc = 3 * H;
    g.c_type = act; gemm(g); }
  prof_end();
  // ---- encoder recurrences: only the last step of each top recurrence receives a gradient
  {
    std::vector<BwdJob> stack, side;
    for (int k = ne - 1; k >= 0; --k) {
      const bool is_top = k == ne - 1;
      BwdJob j; j.r = &enc_pitch[k]; j.kind = IN_DENSE;
      j.X = k == 0 ? slab(Xp_ext, 1, (long)n * PD) : slab(enc_pitch[k - 1].hseq, 1, (long)n * H);
      j.use_dhext = !is_top; j.dh_last = is_top ? du : nullptr; j.ld_last = 3 * H;
      j.need_dx = k > 0; j.dx_out = k > 0 ? enc_pitch[k - 1].dhext : nullptr;
      stack.push_back(j);
    }
    { BwdJob j; j.r = &enc_vel; j.kind = IN_RANK1; j.X = slab(Xv_ext, 1, (long)n * VD); j.dh_last = (const char*)du + (size_t)2 * H * asz(); j.ld_last = 3 * H;
      side.push_back(j); }
    { BwdJob j; j.r = &enc_instr; j.kind = IN_DENSE; j.X = slab(Xi_ext, 1, (long)n * ID); j.dh_last = (const char*)du + (size_t)H * asz(); j.ld_last = 3 * H;
      side.push_back(j); }
    rec_backward_group(stack, side, n, true);
  }
  if (use_side) {   // every weight gradient is in the arena before the all-reduce / optimizer
    MVAE_CUDA(cudaEventRecord(ev_join, this->side));
    MVAE_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
  }
}

}  // namespace mvae
