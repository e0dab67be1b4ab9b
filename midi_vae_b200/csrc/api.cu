// api.cu -- step-level entry points of the model and the extern "C" boundary declared in include/midivae.h.
#include <dlfcn.h>
#include <math.h>
#include <string.h>

#include <mutex>

#include "model.cuh"

namespace mvae {

// ------------------------------------------------------------------------------------------------ steps
// One mini-batch of autoencoder.fit (vae_training.py:804-809): forward + loss + backward; grads stay in Gr.
void Model::forward_backward(const mvae_batch& b, float* dev_metrics, cudaStream_t s) {
  st = s ? s : stream;
  check_batch(b, true);
  MVAE_REQUIRE(cfg.decoder_feedback != MVAE_FB_FREE_RUNNING, "training needs decoder_feedback as_wired or teacher_forced");
  // the cluster / persistent paths are ~120 launches on three streams: nothing to gain from a graph.  Profiling needs its events un-captured.
  if (!step_graph_on || use_persist || profiling || (overlap_allreduce && world > 1)) { forward_backward_body(b, dev_metrics); return; }
  const std::vector<size_t> key = {(size_t)b.n, (size_t)b.pitch, (size_t)b.target, (size_t)b.instr, (size_t)b.velocity, (size_t)b.style,
                                   (size_t)b.history, (size_t)b.eps, (size_t)b.w_notes, (size_t)dev_metrics,
                                   (size_t)history_mode, (size_t)song_start_set, (size_t)carry_valid};   // host-side switches baked into the launches
  if (step_graphs.size() > 64) {   // callers that keep passing fresh buffers would grow the cache without bound
    for (auto& kv : step_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    step_graphs.clear();
  }
  StepGraph& sg = step_graphs[key];
  if (sg.state == 0) {
    forward_backward_body(b, dev_metrics);      // first time: eager (kernel attributes, lazily allocated scheduler words)
    sg.state = 1;
  } else if (sg.state == 1) {
    cudaGraph_t graph = nullptr;
    const long long l0 = g_launches;
    MVAE_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
    try {
      forward_backward_body(b, dev_metrics);
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
}

// One mini-batch of autoencoder.fit (vae_training.py:804-809): forward + loss + backward; grads stay in Gr.
void Model::forward_backward_body(const mvae_batch& b, float* dev_metrics) {
  MVAE_CUDA(cudaMemsetAsync(acc, 0, ACC_COUNT * sizeof(double), st));
  prepare_inputs(b, cfg.decoder_feedback == MVAE_FB_TEACHER_FORCED);
  encoder_forward(b.n);
  head_forward(b, true);
  decoder_forward(b, cfg.decoder_feedback);
  losses(b, true);
  if (dev_metrics) MVAE_CUDA(cudaMemcpyAsync(dev_metrics, d_metrics, MVAE_NUM_METRICS * 4, cudaMemcpyDeviceToDevice, st));
  backward(b);
}

// Data parallelism: sum of the gradient arena over the ranks.  Every NCCL call of a handle goes to ONE stream (st_comm), in the
// same order on every rank.  If backward() already sent the decoder bucket, only the encoder's tensors (arena head) remain; the step's stream
// waits for the communication stream before the optimizer reads Gr.
void Model::allreduce_grads() {
  if (world <= 1 || !nccl_comm) return;
  const size_t head = dec_bucket_issued ? ptab[iWinit].off : arena_n;
  dec_bucket_issued = false;
  MVAE_CUDA(cudaEventRecord(ev_pre_comm, st));          // every gradient is final on the step's stream here (backward() joined its streams)
  MVAE_CUDA(cudaStreamWaitEvent(st_comm, ev_pre_comm, 0));
  prof_begin(PC_ALLREDUCE, st_comm);
  nccl_allreduce_sum_f32(nccl_comm, Gr, head, st_comm);
  prof_end(st_comm);
  MVAE_CUDA(cudaEventRecord(ev_comm, st_comm));
  MVAE_CUDA(cudaStreamWaitEvent(st, ev_comm, 0));
}

// keras.optimizers.Adam (vae_definition.py:174-175; Keras 2.0.8 update rule)
void Model::apply_update(float grad_scale, cudaStream_t s) {
  st = s ? s : stream;
  iterations += 1;
  const double b1 = cfg.adam_beta_1, b2 = cfg.adam_beta_2;
  const float lr_t = (float)(cfg.learning_rate * sqrt(1.0 - pow(b2, (double)iterations)) / (1.0 - pow(b1, (double)iterations)));
  prof_begin(PC_ADAM);
  k_adam((long)arena_n, P, Gr, M1, V2, lr_t, cfg.adam_beta_1, cfg.adam_beta_2, cfg.adam_epsilon, grad_scale, Pb, st);
  prof_end();
}

void Model::eval_step(const mvae_batch& b, float* dev_metrics, cudaStream_t s) {
  st = s ? s : stream;
  check_batch(b, true);
  MVAE_CUDA(cudaMemsetAsync(acc, 0, ACC_COUNT * sizeof(double), st));
  const int fb = cfg.decoder_feedback;
  prepare_inputs(b, fb == MVAE_FB_TEACHER_FORCED);
  encoder_forward(b.n);
  head_forward(b, true);
  if (fb == MVAE_FB_FREE_RUNNING) {
    // the stepwise decoder already applies softmax/sigmoid; recompute the losses on the probabilities
    MVAE_REQUIRE(false, "evaluate() with free_running feedback is not implemented; use as_wired or teacher_forced");
  }
  decoder_forward(b, fb);
  losses(b, false);
  if (dev_metrics) MVAE_CUDA(cudaMemcpyAsync(dev_metrics, d_metrics, MVAE_NUM_METRICS * 4, cudaMemcpyDeviceToDevice, st));
}

// encode -> swap -> shift history -> decode -> argmax (vae_evaluation.py:2448-2550 batched)
void Model::style_transfer(const mvae_batch& b, const uint8_t* song_start, int c_from, int c_to, int feedback, uint8_t* pitch_out,
                           uint8_t* instr_out, float* vel_out, cudaStream_t s) {
  st = s ? s : stream;
  check_batch(b, false);
  MVAE_REQUIRE(c_from >= 0 && c_from < L && c_to >= 0 && c_to < L, "latent dims to swap out of range");
  MVAE_REQUIRE(feedback == MVAE_FB_AS_WIRED || feedback == MVAE_FB_FREE_RUNNING, "style transfer decodes as_wired or free_running (no targets exist)");
  MVAE_CUDA(cudaMemsetAsync(acc, 0, ACC_COUNT * sizeof(double), st));
  inference_pass = true;    // no backward pass follows: the recurrences skip the BPTT stash
  struct Reset { bool& f; ~Reset() { f = false; } } reset{inference_pass};
  prepare_inputs(b, false);
  encoder_forward(b.n);
  mvae_batch b0 = b; b0.eps = nullptr; b0.history = nullptr;   // eps = 0: z = mu (vae_evaluation.py:482-485)
  head_forward(b0, false);
  prof_begin(PC_POINTWISE);
  k_swap_shift(act, b.n, L, ldl, mu, song_start, c_from, c_to, cfg.history, q, ldq, z, st);
  prof_end();
  stepwise_done = false;
  decoder_forward(b, feedback);
  prof_begin(PC_POINTWISE);
  if (!stepwise_done) {
    k_softmax_ce(act, T, b.n, Dp, Pn, ld_pn, nullptr, nullptr, nullptr, 0.f, nullptr, ld_pn, acc, ACC_CE_NOTES, ACC_ACC_NOTES, st);
    k_softmax_ce(act, Ti, b.n, Di, Pi, ld_pi, nullptr, nullptr, nullptr, 0.f, nullptr, ld_pi, acc, ACC_CE_INSTR, ACC_ACC_INSTR, st);
    k_sigmoid_mse(act, T, b.n, Pv, ld_pv, nullptr, 0.f, nullptr, ld_pv, acc, st);
  }
  k_argmax_seq(T, b.n, Dp, Pn, ld_pn, pitch_out, st);
  k_argmax_seq(Ti, b.n, Di, Pi, ld_pi, instr_out, st);
  k_export_seq(T, b.n, 1, Pv, ld_pv, vel_out, st);
  if (post_scope) {   // the reference's process_decoder_outputs rules on the packed outputs, before anything leaves the device
    MVAE_REQUIRE(T % post_voices == 0, "post-processing: input_length must be a multiple of max_voices (import_midi.py:249)");
    k_postprocess_voices(b.n, T, post_voices, Dp - 1, post_threshold, post_scope, post_override, pitch_out, song_start, vel_out, o_held, st);
  }
  prof_end();
}

// ------------------------------------------------------------------------------------------------ NCCL via dlopen
struct Id128 { char b[128]; };   // ncclUniqueId is passed by value
namespace {
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

void nccl_load() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.lib) return;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) throw Error(std::string("cannot dlopen libnccl.so.2: ") + dlerror());
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(lib, "ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(lib, "ncclCommDestroy");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(lib, "ncclAllReduce");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce) throw Error("libnccl.so.2 lacks required symbols");
  g_nccl.lib = lib;
}
void nccl_check(int rc, const char* what) {
  if (rc != 0) throw Error(std::string(what) + " failed: " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "nccl error"));
}
}  // namespace

int nccl_get_unique_id(void* out128) {
  nccl_load();
  nccl_check(g_nccl.GetUniqueId(out128), "ncclGetUniqueId");
  return 0;
}
void* nccl_comm_init(const void* id128, int world, int rank) {
  nccl_load();
  Id128 id;
  memcpy(&id, id128, 128);
  void* comm = nullptr;
  nccl_check(g_nccl.CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
  return comm;
}
void nccl_comm_destroy(void* comm) {
  if (g_nccl.lib && comm) g_nccl.CommDestroy(comm);
}
void nccl_allreduce_sum_f32(void* comm, float* buf, size_t count, cudaStream_t s) {
  // ncclFloat32 = 7, ncclSum = 0
  nccl_check(g_nccl.AllReduce(buf, buf, count, 7, 0, comm, s), "ncclAllReduce");
  count_launch();
}

}  // namespace mvae

// ================================================================================================= extern "C"
using mvae::Model;

struct mvae_model {
  Model* m;
};

static thread_local std::string g_create_error;

#define API_BEGIN(h)                                  \
  if (!(h) || !(h)->m) return 1;                      \
  Model& M = *(h)->m;                                 \
  try {                                               \
    cudaSetDevice(M.device);                          \
    long long _l0 = mvae::g_launches;
#define API_END()                                     \
    M.launches += mvae::g_launches - _l0;             \
    return 0;                                         \
  } catch (const std::exception& ex) {                \
    M.err = ex.what();                                \
    return 2;                                         \
  }

extern "C" {

const char* mvae_version(void) { return "midivae-b200 0.1 (sm_100a)"; }

int mvae_default_config(mvae_config* c) {
  if (!c) return 1;
  memset(c, 0, sizeof(*c));
  c->input_length = 64; c->lstm_size = 256; c->latent_rep_size = 256; c->input_dim = 61; c->meta_instrument_dim = 16;
  c->meta_instrument_length = 4; c->num_composers = 2; c->num_layers_encoder = 2; c->num_layers_decoder = 2;
  c->history = 1; c->extra_layer = 1; c->split_lstm_vector = 1;
  c->gate_act = MVAE_GATE_HARD_SIGMOID; c->dec_cell_variant = MVAE_CELL_RECURRENTSHOP_RECALLED; c->decoder_feedback = MVAE_FB_AS_WIRED;
  c->precision = MVAE_PREC_FP32; c->rnn_mode = MVAE_RNN_AUTO; c->max_batch = 256;
  c->beta = 0.1f; c->prior_mean = 0.f; c->prior_std = 1.f;
  c->notes_weight = 1.f; c->meta_instrument_weight = 0.1f; c->meta_velocity_weight = 1.f; c->composer_weight = 0.1f;
  c->learning_rate = 2e-4f; c->adam_beta_1 = 0.9f; c->adam_beta_2 = 0.999f; c->adam_epsilon = 1e-8f;
  return 0;
}

int mvae_create(const mvae_config* cfg, int device, mvae_handle* out) {
  if (!cfg || !out) { g_create_error = "null argument"; return 1; }
  try {
    Model* m = new Model(*cfg, device);
    *out = new mvae_model{m};
    return 0;
  } catch (const std::exception& ex) {
    g_create_error = ex.what();
    return 2;
  }
}

int mvae_destroy(mvae_handle h) {
  if (!h) return 1;
  delete h->m;
  delete h;
  return 0;
}

const char* mvae_last_error(mvae_handle h) { return (h && h->m) ? h->m->err.c_str() : g_create_error.c_str(); }

int mvae_param_tensor_count(mvae_handle h, int* n) { API_BEGIN(h) *n = (int)M.ptab.size(); API_END() }

int mvae_param_info_at(mvae_handle h, int index, mvae_param_info* out) {
  API_BEGIN(h)
  MVAE_REQUIRE(index >= 0 && index < (int)M.ptab.size(), "parameter index out of range");
  const auto& t = M.ptab[index];
  memset(out, 0, sizeof(*out));
  strncpy(out->name, t.name.c_str(), sizeof(out->name) - 1);
  out->offset = t.off; out->rows = t.rows; out->cols = t.cols; out->ld = t.ld;
  API_END()
}

int mvae_arena_size(mvae_handle h, size_t* n) { API_BEGIN(h) *n = M.arena_n; API_END() }

static void copy_tensor(Model& M, const float* dev_arena, int index, float* dst_host) {
  MVAE_REQUIRE(index >= 0 && index < (int)M.ptab.size(), "parameter index out of range");
  const auto& t = M.ptab[index];
  MVAE_CUDA(cudaStreamSynchronize(M.stream));
  MVAE_CUDA(cudaMemcpy2D(dst_host, (size_t)t.cols * 4, dev_arena + t.off, (size_t)t.ld * 4, (size_t)t.cols * 4, t.rows, cudaMemcpyDeviceToHost));
}

int mvae_get_param(mvae_handle h, int index, float* dst_host) { API_BEGIN(h) copy_tensor(M, M.P, index, dst_host); API_END() }
int mvae_get_grad(mvae_handle h, int index, float* dst_host) { API_BEGIN(h) copy_tensor(M, M.Gr, index, dst_host); API_END() }

int mvae_set_param(mvae_handle h, int index, const float* src_host) {
  API_BEGIN(h)
  MVAE_REQUIRE(index >= 0 && index < (int)M.ptab.size(), "parameter index out of range");
  const auto& t = M.ptab[index];
  MVAE_CUDA(cudaStreamSynchronize(M.stream));
  MVAE_CUDA(cudaMemcpy2D(M.P + t.off, (size_t)t.ld * 4, src_host, (size_t)t.cols * 4, (size_t)t.cols * 4, t.rows, cudaMemcpyHostToDevice));
  API_END()
}

int mvae_commit_params(mvae_handle h) {
  API_BEGIN(h)
  M.st = M.stream;
  M.commit_params();
  MVAE_CUDA(cudaStreamSynchronize(M.stream));
  API_END()
}

int mvae_reset_optimizer(mvae_handle h) {
  API_BEGIN(h)
  MVAE_CUDA(cudaMemsetAsync(M.M1, 0, M.arena_n * 4, M.stream));
  MVAE_CUDA(cudaMemsetAsync(M.V2, 0, M.arena_n * 4, M.stream));
  MVAE_CUDA(cudaStreamSynchronize(M.stream));
  M.iterations = 0;
  API_END()
}

int mvae_get_iterations(mvae_handle h, long long* t) { API_BEGIN(h) *t = M.iterations; API_END() }

int mvae_forward_backward(mvae_handle h, const mvae_batch* b, float* dev_metrics, void* stream) {
  API_BEGIN(h)
  MVAE_REQUIRE(b != nullptr, "batch is null");
  M.forward_backward(*b, dev_metrics, (cudaStream_t)stream);
  API_END()
}

int mvae_apply_update(mvae_handle h, float grad_scale, void* stream) {
  API_BEGIN(h)
  M.apply_update(grad_scale, (cudaStream_t)stream);
  API_END()
}

int mvae_train_step(mvae_handle h, const mvae_batch* b, float* dev_metrics, void* stream) {
  API_BEGIN(h)
  MVAE_REQUIRE(b != nullptr, "batch is null");
  M.overlap_allreduce = true;
  struct Reset { bool& f; ~Reset() { f = false; } } reset{M.overlap_allreduce};
  M.forward_backward(*b, dev_metrics, (cudaStream_t)stream);
  M.allreduce_grads();
  M.apply_update(1.0f / (float)M.world, (cudaStream_t)stream);
  API_END()
}

static void fetch_metrics(Model& M, mvae_metrics* out) {
  MVAE_CUDA(cudaMemcpyAsync(M.pin, M.d_metrics, MVAE_NUM_METRICS * 4, cudaMemcpyDeviceToHost, M.st));
  MVAE_CUDA(cudaStreamSynchronize(M.st));
  M.d2h_bytes += MVAE_NUM_METRICS * 4;
  if (out) memcpy(out->v, M.pin, MVAE_NUM_METRICS * 4);
}

int mvae_train_step_host(mvae_handle h, const mvae_batch* hb, mvae_metrics* out) {
  API_BEGIN(h)
  MVAE_REQUIRE(hb != nullptr, "batch is null");
  M.st = M.stream;
  M.check_batch(*hb, true);
  mvae_batch d = M.upload(*hb);
  M.overlap_allreduce = true;
  struct Reset { bool& f; ~Reset() { f = false; } } reset{M.overlap_allreduce};
  M.forward_backward(d, nullptr, nullptr);
  M.allreduce_grads();
  M.apply_update(1.0f / (float)M.world, nullptr);
  fetch_metrics(M, out);
  API_END()
}

int mvae_eval_step(mvae_handle h, const mvae_batch* b, float* dev_metrics, void* stream) {
  API_BEGIN(h)
  MVAE_REQUIRE(b != nullptr, "batch is null");
  M.eval_step(*b, dev_metrics, (cudaStream_t)stream);
  API_END()
}

int mvae_eval_step_host(mvae_handle h, const mvae_batch* hb, mvae_metrics* out) {
  API_BEGIN(h)
  MVAE_REQUIRE(hb != nullptr, "batch is null");
  M.st = M.stream;
  M.check_batch(*hb, true);
  mvae_batch d = M.upload(*hb);
  M.eval_step(d, nullptr, nullptr);
  fetch_metrics(M, out);
  API_END()
}

// device (rows, cols, ld) fp32 -> host dense
static void fetch_f32(Model& M, const float* dev, int rows, int cols, int ld, float* host) {
  if (!host) return;
  MVAE_CUDA(cudaStreamSynchronize(M.st));
  MVAE_CUDA(cudaMemcpy2D(host, (size_t)cols * 4, dev, (size_t)ld * 4, (size_t)cols * 4, rows, cudaMemcpyDeviceToHost));
  M.d2h_bytes += (size_t)rows * cols * 4;
}

int mvae_encode_host(mvae_handle h, const mvae_batch* hb, float* z_out, float* mu_out, float* logvar_out) {
  API_BEGIN(h)
  MVAE_REQUIRE(hb != nullptr, "batch is null");
  M.st = M.stream;
  M.check_batch(*hb, false);
  mvae_batch d = M.upload(*hb);
  MVAE_CUDA(cudaMemsetAsync(M.acc, 0, mvae::ACC_COUNT * sizeof(double), M.st));
  M.prepare_inputs(d, false);
  M.encoder_forward(d.n);
  M.head_forward(d, false);
  fetch_f32(M, M.z, d.n, M.L, M.ldl, z_out);
  fetch_f32(M, M.mu, d.n, M.L, M.ldl, mu_out);
  fetch_f32(M, M.lv, d.n, M.L, M.ldl, logvar_out);
  API_END()
}

static void export_outputs(Model& M, int n, float* y_out, float* i_out, float* v_out) {
  using namespace mvae;
  if (!M.stepwise_done) {
    k_softmax_ce(M.act, M.T, n, M.Dp, M.Pn, M.ld_pn, nullptr, nullptr, nullptr, 0.f, nullptr, M.ld_pn, M.acc, ACC_CE_NOTES, ACC_ACC_NOTES, M.st);
    k_softmax_ce(M.act, M.Ti, n, M.Di, M.Pi, M.ld_pi, nullptr, nullptr, nullptr, 0.f, nullptr, M.ld_pi, M.acc, ACC_CE_INSTR, ACC_ACC_INSTR, M.st);
    k_sigmoid_mse(M.act, M.T, n, M.Pv, M.ld_pv, nullptr, 0.f, nullptr, M.ld_pv, M.acc, M.st);
  }
  k_export_seq(M.T, n, M.Dp, M.Pn, M.ld_pn, M.o_y, M.st);
  k_export_seq(M.Ti, n, M.Di, M.Pi, M.ld_pi, M.o_i, M.st);
  k_export_seq(M.T, n, 1, M.Pv, M.ld_pv, M.o_v, M.st);
  MVAE_CUDA(cudaStreamSynchronize(M.st));
  if (y_out) { MVAE_CUDA(cudaMemcpy(y_out, M.o_y, (size_t)n * M.T * M.Dp * 4, cudaMemcpyDeviceToHost)); M.d2h_bytes += (size_t)n * M.T * M.Dp * 4; }
  if (i_out) { MVAE_CUDA(cudaMemcpy(i_out, M.o_i, (size_t)n * M.Ti * M.Di * 4, cudaMemcpyDeviceToHost)); M.d2h_bytes += (size_t)n * M.Ti * M.Di * 4; }
  if (v_out) { MVAE_CUDA(cudaMemcpy(v_out, M.o_v, (size_t)n * M.T * 4, cudaMemcpyDeviceToHost)); M.d2h_bytes += (size_t)n * M.T * 4; }
}

int mvae_decode_host(mvae_handle h, const mvae_batch* hb, const float* z, const float* history, int feedback, float* y_out, float* i_out,
                     float* v_out) {
  API_BEGIN(h)
  MVAE_REQUIRE(hb != nullptr && z != nullptr, "batch / z is null");
  MVAE_REQUIRE(hb->n >= 1 && hb->n <= M.NB, "mini-batch size must be in 1..max_batch");
  M.st = M.stream;
  const int n = hb->n;
  if (feedback == MVAE_FB_TEACHER_FORCED) MVAE_REQUIRE(hb->pitch && hb->instr && hb->velocity, "teacher_forced decode needs target rolls");
  mvae_batch hb2 = *hb;
  hb2.history = history; hb2.eps = nullptr;
  mvae_batch d = M.upload(hb2);
  // z travels through the eps staging buffer
  memcpy(M.pin + M.pin_bytes - (size_t)n * M.L * 4 - 64, z, (size_t)n * M.L * 4);
  MVAE_CUDA(cudaMemcpyAsync(M.d_eps, M.pin + M.pin_bytes - (size_t)n * M.L * 4 - 64, (size_t)n * M.L * 4, cudaMemcpyHostToDevice, M.st));
  M.h2d_bytes += (size_t)n * M.L * 4;
  MVAE_CUDA(cudaMemsetAsync(M.acc, 0, mvae::ACC_COUNT * sizeof(double), M.st));
  if (feedback == MVAE_FB_TEACHER_FORCED) M.prepare_inputs(d, true);
  mvae::k_build_q(M.act, n, M.L, M.d_eps, d.history, M.cfg.history, M.q, M.ldq, M.st);
  M.stepwise_done = false;
  M.decoder_forward(d, feedback);
  export_outputs(M, n, y_out, i_out, v_out);
  API_END()
}

int mvae_autoencode_host(mvae_handle h, const mvae_batch* hb, float* y_out, float* i_out, float* v_out, float* style_out, float* z_out) {
  API_BEGIN(h)
  MVAE_REQUIRE(hb != nullptr, "batch is null");
  M.st = M.stream;
  M.check_batch(*hb, false);
  mvae_batch d = M.upload(*hb);
  MVAE_CUDA(cudaMemsetAsync(M.acc, 0, mvae::ACC_COUNT * sizeof(double), M.st));
  const int fb = M.cfg.decoder_feedback;
  M.prepare_inputs(d, fb == MVAE_FB_TEACHER_FORCED);
  M.encoder_forward(d.n);
  M.head_forward(d, false);
  M.stepwise_done = false;
  M.decoder_forward(d, fb);
  export_outputs(M, d.n, y_out, i_out, v_out);
  fetch_f32(M, M.style_probs, d.n, M.C, M.C, style_out);
  fetch_f32(M, M.z, d.n, M.L, M.ldl, z_out);
  API_END()
}

int mvae_style_transfer(mvae_handle h, const mvae_batch* b, const uint8_t* dev_song_start, int c_from, int c_to, int feedback,
                        uint8_t* dev_pitch_out, uint8_t* dev_instr_out, float* dev_velocity_out, void* stream) {
  API_BEGIN(h)
  MVAE_REQUIRE(b && dev_pitch_out && dev_instr_out && dev_velocity_out, "null argument");
  M.style_transfer(*b, dev_song_start, c_from, c_to, feedback, dev_pitch_out, dev_instr_out, dev_velocity_out, (cudaStream_t)stream);
  API_END()
}

int mvae_style_transfer_host(mvae_handle h, const mvae_batch* hb, const uint8_t* song_start, int c_from, int c_to, int feedback,
                             uint8_t* pitch_out, uint8_t* instr_out, float* velocity_out) {
  API_BEGIN(h)
  MVAE_REQUIRE(hb && pitch_out && instr_out && velocity_out, "null argument");
  M.st = M.stream;
  M.check_batch(*hb, false);
  mvae_batch d = M.upload(*hb, song_start);
  const int n = d.n;
  M.style_transfer(d, song_start ? M.d_song_start : nullptr, c_from, c_to, feedback, M.o_pitch, M.o_instr, M.o_v, nullptr);
  MVAE_CUDA(cudaStreamSynchronize(M.st));
  MVAE_CUDA(cudaMemcpy(pitch_out, M.o_pitch, (size_t)n * M.T, cudaMemcpyDeviceToHost));
  MVAE_CUDA(cudaMemcpy(instr_out, M.o_instr, (size_t)n * M.Ti, cudaMemcpyDeviceToHost));
  MVAE_CUDA(cudaMemcpy(velocity_out, M.o_v, (size_t)n * M.T * 4, cudaMemcpyDeviceToHost));
  M.d2h_bytes += (size_t)n * M.T * 5 + (size_t)n * M.Ti;
  API_END()
}

int mvae_set_postprocess(mvae_handle h, int scope, float velocity_threshold, int override_by_velocity, int max_voices) {
  API_BEGIN(h)
  MVAE_REQUIRE(scope >= 0 && scope <= 2, "postprocess scope: 0 off, 1 per chunk, 2 per song");
  MVAE_REQUIRE(max_voices >= 1 && M.T % max_voices == 0, "max_voices must divide input_length");
  M.post_scope = scope; M.post_threshold = velocity_threshold; M.post_override = override_by_velocity != 0; M.post_voices = max_voices;
  API_END()
}

int mvae_postprocess_host(mvae_handle h, int n, const uint8_t* pitch, const uint8_t* song_start, int scope, float* velocity_inout, uint8_t* held_out) {
  API_BEGIN(h)
  MVAE_REQUIRE(pitch && velocity_inout, "null argument");
  MVAE_REQUIRE(n >= 1 && n <= M.NB, "number of chunks must be in 1..max_batch");
  MVAE_REQUIRE(scope == 1 || scope == 2, "postprocess scope: 1 per chunk, 2 per song");
  M.st = M.stream;
  const size_t nt = (size_t)n * M.T;
  MVAE_CUDA(cudaMemcpyAsync(M.o_pitch, pitch, nt, cudaMemcpyHostToDevice, M.st));
  MVAE_CUDA(cudaMemcpyAsync(M.o_v, velocity_inout, nt * 4, cudaMemcpyHostToDevice, M.st));
  if (song_start) MVAE_CUDA(cudaMemcpyAsync(M.d_song_start, song_start, (size_t)n, cudaMemcpyHostToDevice, M.st));
  M.h2d_bytes += nt * 5 + (song_start ? n : 0);
  mvae::k_postprocess_voices(n, M.T, M.post_voices, M.Dp - 1, M.post_threshold, scope, M.post_override, M.o_pitch, song_start ? M.d_song_start : nullptr,
                             M.o_v, M.o_held, M.st);
  MVAE_CUDA(cudaMemcpyAsync(velocity_inout, M.o_v, nt * 4, cudaMemcpyDeviceToHost, M.st));
  if (held_out) MVAE_CUDA(cudaMemcpyAsync(held_out, M.o_held, nt, cudaMemcpyDeviceToHost, M.st));
  MVAE_CUDA(cudaStreamSynchronize(M.st));
  M.d2h_bytes += nt * 4 + (held_out ? nt : 0);
  API_END()
}

int mvae_set_history_mode(mvae_handle h, int mode) {
  API_BEGIN(h)
  MVAE_REQUIRE(mode == 0 || mode == 1, "history mode: 0 = batch.history (the reference's separate encoder pass), 1 = from the batch's own z");
  MVAE_REQUIRE(!M.cls, "not applicable to a classifier handle");
  M.history_mode = mode; M.carry_valid = false; M.song_start_set = false;
  API_END()
}

int mvae_set_song_start_host(mvae_handle h, const uint8_t* song_start, int n) {
  API_BEGIN(h)
  MVAE_REQUIRE(!M.cls, "not applicable to a classifier handle");
  MVAE_REQUIRE(n >= 0 && n <= M.NB, "song_start: at most max_batch flags");
  if (song_start && n > 0) {
    MVAE_CUDA(cudaStreamSynchronize(M.stream));     // the staging below is a plain pageable copy: keep it out of a step in flight
    MVAE_CUDA(cudaMemcpy(M.d_song_start, song_start, (size_t)n, cudaMemcpyHostToDevice));
    M.h2d_bytes += (size_t)n;
    M.song_start_set = true;
  } else {
    M.song_start_set = false;
  }
  API_END()
}

static mvae_batch cls_upload(Model& M, const mvae_batch& hb) {
  MVAE_REQUIRE(M.cls, "this handle is not a style classifier (mvae_config::model_kind = 1)");
  MVAE_REQUIRE(hb.n >= 1 && hb.n <= M.NB, "mini-batch size must be in 1..max_batch");
  mvae_batch in{}; in.n = hb.n;
  in.pitch = M.cls_scalar ? nullptr : hb.pitch; in.velocity = M.cls_scalar ? hb.velocity : nullptr; in.style = hb.style;
  return M.upload(in);
}

int mvae_cls_train_step_host(mvae_handle h, const mvae_batch* hb, mvae_metrics* out) {
  API_BEGIN(h)
  MVAE_REQUIRE(hb != nullptr, "batch is null");
  M.st = M.stream;
  mvae_batch d = cls_upload(M, *hb);
  M.cls_forward(d, true);
  M.cls_backward(d);
  M.apply_update(1.0f, nullptr);
  fetch_metrics(M, out);
  API_END()
}

int mvae_cls_eval_step_host(mvae_handle h, const mvae_batch* hb, mvae_metrics* out, float* probs_out) {
  API_BEGIN(h)
  MVAE_REQUIRE(hb != nullptr, "batch is null");
  M.st = M.stream;
  mvae_batch d = cls_upload(M, *hb);
  M.inference_pass = true;
  struct Reset { bool& f; ~Reset() { f = false; } } reset{M.inference_pass};
  M.cls_forward(d, false);
  fetch_metrics(M, out);
  fetch_f32(M, M.Pn, d.n, M.C, M.ld_pn, probs_out);
  API_END()
}

int mvae_grad_arena(mvae_handle h, float** dev_ptr, size_t* n) { API_BEGIN(h) *dev_ptr = M.Gr; *n = M.arena_n; API_END() }
int mvae_param_arena(mvae_handle h, float** dev_ptr, size_t* n) { API_BEGIN(h) *dev_ptr = M.P; *n = M.arena_n; API_END() }

int mvae_nccl_unique_id(void* id_out) {
  try { return mvae::nccl_get_unique_id(id_out); } catch (const std::exception& ex) { g_create_error = ex.what(); return 2; }
}

int mvae_nccl_init(mvae_handle h, const void* id, int world_size, int rank) {
  API_BEGIN(h)
  MVAE_REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, "bad world_size / rank");
  if (world_size > 1) {
    M.nccl_comm = mvae::nccl_comm_init(id, world_size, rank);
    if (!M.st_comm) {
      int lo = 0, hi = 0;
      MVAE_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      MVAE_CUDA(cudaStreamCreateWithPriority(&M.st_comm, cudaStreamNonBlocking, hi));
      MVAE_CUDA(cudaEventCreateWithFlags(&M.ev_dec_grads, cudaEventDisableTiming));
      MVAE_CUDA(cudaEventCreateWithFlags(&M.ev_comm, cudaEventDisableTiming));
      MVAE_CUDA(cudaEventCreateWithFlags(&M.ev_pre_comm, cudaEventDisableTiming));
    }
  }
  M.world = world_size; M.rank = rank;
  API_END()
}

int mvae_world_size(mvae_handle h, int* n) { API_BEGIN(h) *n = M.world; API_END() }
int mvae_launch_count(mvae_handle h, long long* n) { API_BEGIN(h) *n = M.launches; API_END() }
int mvae_sync(mvae_handle h) { API_BEGIN(h) MVAE_CUDA(cudaStreamSynchronize(M.stream)); API_END() }

int mvae_set_profiling(mvae_handle h, int on) { API_BEGIN(h) M.profiling = on != 0; API_END() }

int mvae_last_kernel_ms(mvae_handle h, int which, float* ms, long long* launches) {
  API_BEGIN(h)
  MVAE_REQUIRE(which >= 0 && which < mvae::PC_COUNT, "kernel class out of range");
  if (!M.evs.empty()) M.prof_collect();
  if (ms) *ms = M.prof_ms[which];
  if (launches) *launches = M.prof_n[which];
  API_END()
}

int mvae_transfer_bytes(mvae_handle h, unsigned long long* h2d, unsigned long long* d2h, int reset) {
  API_BEGIN(h)
  if (h2d) *h2d = M.h2d_bytes;
  if (d2h) *d2h = M.d2h_bytes;
  if (reset) { M.h2d_bytes = 0; M.d2h_bytes = 0; }
  API_END()
}

int mvae_stream(mvae_handle h, void** stream) { API_BEGIN(h) *stream = (void*)M.stream; API_END() }

int mvae_selftest_gemm(int device, int verbose) {
  try { return mvae::gemm_tc_selftest(device, verbose); } catch (const std::exception& ex) { g_create_error = ex.what(); fprintf(stderr, "selftest: %s\n", ex.what()); return 2; }
}

}  // extern "C"
