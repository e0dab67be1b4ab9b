/* midivae.h -- C ABI of the B200-native MIDI-VAE hot path (libmidivae.so).
 *
 * Drop-in boundary for ONE path of brunnergino/MIDI-VAE: the three-stream recurrent VAE
 * train step and encode -> swap-style -> decode inference that the reference runs inside
 * Keras (vae_definition.py:39-441 builds it; vae_training.py:804-809 and
 * vae_evaluation.py:2180-2181,2474-2483 call it).  The reference has no FFI of its own
 * (it is pure Python over Keras/Theano); every entry point below names the Keras call it
 * replaces.  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA-runtime types in signatures
 *     (a CUDA stream is passed as void*, NULL = the handle's own stream);
 *   - the caller owns every buffer; `_host` variants take HOST pointers and perform the
 *     host<->device copies themselves (synchronous); the others take DEVICE pointers and are
 *     asynchronous on the given stream;
 *   - every function returns 0 on success, non-zero on error; mvae_last_error() explains;
 *   - one handle per device, no internal threads, re-entrant per handle;
 *   - rolls are PACKED: pitch/target u8 class index per (chunk, step) with 60 = silent
 *     (the argmax of the reference's one-hot (N,T,61) roll, import_midi.py:243-286),
 *     instrument u8 category per voice, velocity f32 (N,T), style u8 class per chunk.
 */
#ifndef MIDIVAE_H_
#define MIDIVAE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mvae_model* mvae_handle;

enum { MVAE_GATE_HARD_SIGMOID = 0, MVAE_GATE_SIGMOID = 1 };            /* Keras 2.0.8 default / north_star wording */
enum { MVAE_CELL_STANDARD = 0, MVAE_CELL_RECURRENTSHOP_RECALLED = 1 };  /* SURVEY.md appendix A.3 */
enum { MVAE_FB_AS_WIRED = 0, MVAE_FB_TEACHER_FORCED = 1, MVAE_FB_FREE_RUNNING = 2 };  /* SURVEY.md 8(a) row D-fb */
enum { MVAE_PREC_FP32 = 0, MVAE_PREC_BF16 = 1 };   /* fp32 SIMT parity path / bf16 tcgen05 tensor-core path */
/* sequence loop of a recurrence (bf16 only): STREAMED = one GEMM + gate-math launch per step; PERSISTENT = require the LSTM persistent / cluster
   kernels (error if the cell / size has none); AUTO = the fastest available form: LSTM cluster kernels (H = 256 / 512), first-generation
   persistent kernels (other H % 64 == 0), GRU cluster kernels (H = 256), else streamed */
enum { MVAE_RNN_STREAMED = 0, MVAE_RNN_PERSISTENT = 1, MVAE_RNN_AUTO = 2 };
enum { MVAE_CELLTYPE_LSTM = 0, MVAE_CELLTYPE_GRU = 1 };                 /* vae_definition.py:457-472,535,585,623; settings.py:155 ships GRU */

/* Mirrors the VAE.create(...) kwargs that are live on the hot path (vae_definition.py:40-102). */
typedef struct mvae_config {
  int input_length;            /* T  = output_length                   */
  int lstm_size;               /* H                                    */
  int latent_rep_size;         /* L                                    */
  int input_dim;               /* 61 = output_dim                      */
  int meta_instrument_dim;     /* 16                                   */
  int meta_instrument_length;  /* 4                                    */
  int num_composers;           /* 2                                    */
  int num_layers_encoder;
  int num_layers_decoder;
  int history;                 /* settings.py:145                      */
  int extra_layer;             /* settings.py:164                      */
  int split_lstm_vector;       /* settings.py:139                      */
  int gate_act;                /* MVAE_GATE_*                          */
  int dec_cell_variant;        /* MVAE_CELL_*                          */
  int decoder_feedback;        /* MVAE_FB_* used by train/eval/autoencode */
  int precision;               /* MVAE_PREC_*                          */
  int rnn_mode;                /* MVAE_RNN_* (bf16 precision only)     */
  int max_batch;               /* largest mini-batch a call may carry  */
  float beta, prior_mean, prior_std;                     /* KLDivergenceLayer, vae_definition.py:15-37 */
  float notes_weight, meta_instrument_weight, meta_velocity_weight, composer_weight; /* loss_weights, :336-397 */
  float learning_rate, adam_beta_1, adam_beta_2, adam_epsilon;  /* keras.optimizers.Adam, :174-175 */
  int cell_type;               /* MVAE_CELLTYPE_*: LSTM (north_star; cluster / persistent kernels) or GRU (the reference's shipped default;
                                  step-streamed kernels).  GRU: Keras 2.0.8 GRU encoders (blocks [z|r|h], h' = z h + (1-z) hh) and recurrentshop
                                  GRUCell decoders (as recalled: h' = (1-z) h + z hh; dec_cell_variant applies to the LSTM cells only)      */
  int model_kind;              /* 0: the three-stream VAE (everything above).  1: a STYLE CLASSIFIER (pitch_classifier.py:89-103,
                                  velocity_classifier.py:110-118, instrument_classifier.py:93-103): num_layers_encoder stacked recurrent layers of
                                  lstm_size units over a (input_length, D) sequence -> last state -> Dense(num_composers, softmax), categorical
                                  cross-entropy, Adam.  Only the mvae_cls_* entry points and the parameter / optimizer calls apply to it.       */
  int cls_scalar_input;        /* classifier: 0 = one-hot input of input_dim classes given as class indices (batch.pitch, u8 [n, input_length]:
                                  the pitch roll, or the 4 x 16 instrument matrix with input_length = 4, input_dim = 16); 1 = scalar input
                                  (batch.velocity, f32 [n, input_length]).  Labels: batch.style.                                              */
} mvae_config;

/* One mini-batch = a consecutive slice of <= batch_size chunks of a song
 * (what Keras hands its train_function inside autoencoder.fit, vae_training.py:804-809). */
typedef struct mvae_batch {
  int n;                    /* chunks in this mini-batch, 1..max_batch                               */
  const uint8_t* pitch;     /* [n,T]  X, encoder pitch roll                                          */
  const uint8_t* target;    /* [n,T]  Y, decoder target roll; NULL = pitch (vae_definition.py:926)   */
  const uint8_t* instr;     /* [n,4]  I                                                              */
  const float* velocity;    /* [n,T]  V                                                              */
  const uint8_t* style;     /* [n]    C                                                              */
  const float* history;     /* [n,L]  H; NULL = zeros (epoch 0, vae_training.py:789-790)             */
  const float* eps;         /* [n,L]  reparameterisation noise; NULL = 0 (vae_evaluation.py:482-485) */
  const float* w_notes;     /* [n,T]  temporal sample weights; NULL = ones (vae_definition.py:928-933) */
} mvae_batch;

/* hist.history / autoencoder.metrics_names order (vae_training.py:817-853) + the KL term. */
enum { MVAE_M_LOSS = 0, MVAE_M_NOTES_LOSS, MVAE_M_INSTR_LOSS, MVAE_M_VEL_LOSS, MVAE_M_STYLE_LOSS,
       MVAE_M_NOTES_ACC, MVAE_M_INSTR_ACC, MVAE_M_VEL_ACC, MVAE_M_STYLE_ACC, MVAE_M_KL, MVAE_NUM_METRICS };
typedef struct mvae_metrics { float v[MVAE_NUM_METRICS]; } mvae_metrics;

typedef struct mvae_param_info {
  char name[96];       /* e.g. "lstm_1/recurrent_kernel"                         */
  size_t offset;       /* element offset inside the flat fp32 arena              */
  int rows, cols, ld;  /* row-major, ld >= cols (pads are zero and stay zero)    */
} mvae_param_info;

/* ---- lifecycle: VAE().create(...)  (vae_training.py:47-109, vae_definition.py:40) ---- */
int mvae_default_config(mvae_config* cfg);                       /* settings.py defaults, LSTM branch */
int mvae_create(const mvae_config* cfg, int device, mvae_handle* out);
int mvae_destroy(mvae_handle h);
const char* mvae_last_error(mvae_handle h);                      /* h may be NULL: last create() error */

/* ---- weights: save_weights / load_weights (vae_training.py:120-123,966-978) ---- */
int mvae_param_tensor_count(mvae_handle h, int* n);
int mvae_param_info_at(mvae_handle h, int index, mvae_param_info* out);
int mvae_arena_size(mvae_handle h, size_t* n_floats);            /* padded flat arena incl. alignment gaps */
int mvae_get_param(mvae_handle h, int index, float* dst_host);   /* rows*cols floats, dense               */
int mvae_set_param(mvae_handle h, int index, const float* src_host);
int mvae_commit_params(mvae_handle h);                           /* refresh bf16 shadows after set_param  */
int mvae_get_grad(mvae_handle h, int index, float* dst_host);    /* gradient of the last forward_backward  */
int mvae_reset_optimizer(mvae_handle h);                         /* Adam m, v, iterations := 0            */
int mvae_get_iterations(mvae_handle h, long long* t);

/* ---- train: one mini-batch of autoencoder.fit(epochs=1) (vae_training.py:804-809) ---- */
int mvae_train_step(mvae_handle h, const mvae_batch* dev_batch, float* dev_metrics /* [MVAE_NUM_METRICS] or NULL */, void* stream);
int mvae_train_step_host(mvae_handle h, const mvae_batch* host_batch, mvae_metrics* out);
/* split form for data parallelism: grads are left in the arena returned by mvae_grad_arena */
int mvae_forward_backward(mvae_handle h, const mvae_batch* dev_batch, float* dev_metrics, void* stream);
int mvae_apply_update(mvae_handle h, float grad_scale, void* stream);   /* Keras Adam on grad*grad_scale */
int mvae_grad_arena(mvae_handle h, float** dev_ptr, size_t* n_floats);
int mvae_param_arena(mvae_handle h, float** dev_ptr, size_t* n_floats);

/* ---- evaluate: autoencoder.evaluate (vae_training.py:300) ---- */
int mvae_eval_step(mvae_handle h, const mvae_batch* dev_batch, float* dev_metrics, void* stream);
int mvae_eval_step_host(mvae_handle h, const mvae_batch* host_batch, mvae_metrics* out);

/* ---- predict: encoder / decoder / autoencoder .predict (vae_training.py:289, vae_evaluation.py:798,2199,2482) ---- */
/* z = mu + exp(logvar/2)*eps (eps NULL => z = mu); z_out, mu_out, logvar_out are [n,L], each may be NULL */
int mvae_encode_host(mvae_handle h, const mvae_batch* host_batch, float* z_out, float* mu_out, float* logvar_out);
/* decoder.predict([Y0=0, z, H, I0=0, V0=0]) -> Y [n,T,61], I [n,4,16], V [n,T]; feedback = MVAE_FB_* (teacher_forced
 * needs host_batch->target/instr/velocity, else host_batch may carry only n) */
int mvae_decode_host(mvae_handle h, const mvae_batch* host_batch, const float* z, const float* history, int feedback,
                     float* y_out, float* i_out, float* v_out);
/* autoencoder.predict -> [Y, I, V, style (n,C)] plus z */
int mvae_autoencode_host(mvae_handle h, const mvae_batch* host_batch, float* y_out, float* i_out, float* v_out,
                         float* style_out, float* z_out);

/* ---- inference: encode -> swap latent dims c_from<->c_to -> shift history -> decode -> argmax
 *      (batched form of vae_evaluation.py:2448-2550; post-processing vae_definition.py:1071-1107) ---- */
int mvae_style_transfer(mvae_handle h, const mvae_batch* dev_batch, const uint8_t* dev_song_start /* [n] or NULL */,
                        int c_from, int c_to, int feedback,
                        uint8_t* dev_pitch_out /* [n,T] */, uint8_t* dev_instr_out /* [n,4] */, float* dev_velocity_out /* [n,T] */,
                        void* stream);
int mvae_style_transfer_host(mvae_handle h, const mvae_batch* host_batch, const uint8_t* song_start,
                             int c_from, int c_to, int feedback,
                             uint8_t* pitch_out, uint8_t* instr_out, float* velocity_out);

/* ---- output post-processing on the device: process_decoder_outputs (vae_definition.py:1131-1225, sample_method 'argmax') on the packed
 *      outputs -- silent steps get velocity 0, override_sampled_pitches_based_on_velocity_info per voice (:1161-1190), held-note roll
 *      D = 0 where the final velocity is above the played-note threshold (:1214-1221).
 *      scope: 0 off (default), 1 the per-voice memory restarts at every chunk (the style-switch loop calls the reference function once per
 *      chunk, vae_evaluation.py:2483), 2 at every song start (its whole-song call sites, vae_evaluation.py:799,814).
 *      Once set, mvae_style_transfer[_host] applies it before the velocities leave the device. ---- */
int mvae_set_postprocess(mvae_handle h, int scope, float velocity_threshold /* settings.py:30: 0.5 */, int override_by_velocity, int max_voices);
/* the same rules on caller-supplied packed rolls (host): pitch u8 [n,T] (input_dim-1 = silent), velocity f32 [n,T] in place, held u8 [n,T] or NULL */
int mvae_postprocess_host(mvae_handle h, int n, const uint8_t* pitch, const uint8_t* song_start /* [n] or NULL */, int scope,
                          float* velocity_inout, uint8_t* held_out);

/* ---- history latents from the batch itself (opt-in): mode 1 makes train / eval steps build the decoder's history input from the
 *      step's own z -- H[i] = z[i-1] inside a song, 0 on a song's first chunk (flags set with mvae_set_song_start_host for the NEXT call; none =
 *      one song), row 0 continuing the previous call's last row -- instead of reading batch.history, which the reference fills from a separate
 *      encoder.predict of the song (vae_training.py:788-798).  Saves that encoder pass; differs from the reference in that the history comes from
 *      the same weights and the same epsilon draw as the step.  mode 0 (default) = batch.history. ---- */
int mvae_set_history_mode(mvae_handle h, int mode);
int mvae_set_song_start_host(mvae_handle h, const uint8_t* song_start /* [n] or NULL */, int n);

/* ---- style classifiers (model_kind = 1; the evaluators of vae_evaluation.py:110-117): one mini-batch of model.fit / evaluate / predict.
 *      metrics out: v[MVAE_M_LOSS] = v[MVAE_M_STYLE_LOSS] = mean categorical cross-entropy, v[MVAE_M_STYLE_ACC] = accuracy; probs [n, num_composers] ---- */
int mvae_cls_train_step_host(mvae_handle h, const mvae_batch* host_batch, mvae_metrics* out);
int mvae_cls_eval_step_host(mvae_handle h, const mvae_batch* host_batch, mvae_metrics* out, float* probs_out /* or NULL */);

/* ---- multi-GPU data parallelism: ONE ncclAllReduce(sum) over the gradient arena per step ---- */
int mvae_nccl_unique_id(void* id_out_128_bytes);
int mvae_nccl_init(mvae_handle h, const void* id_128_bytes, int world_size, int rank);
int mvae_world_size(mvae_handle h, int* n);

/* ---- introspection for bench / tests ---- */
int mvae_launch_count(mvae_handle h, long long* n);   /* kernels launched by this handle so far        */
int mvae_sync(mvae_handle h);
int mvae_last_kernel_ms(mvae_handle h, int which, float* ms, long long* launches); /* CUDA-event time of a kernel class in the last step */
int mvae_set_profiling(mvae_handle h, int on);          /* record CUDA events around each kernel class */
int mvae_transfer_bytes(mvae_handle h, unsigned long long* h2d, unsigned long long* d2h, int reset); /* bytes moved by the _host calls */
int mvae_stream(mvae_handle h, void** stream);          /* the handle's own cudaStream_t */
int mvae_selftest_gemm(int device, int verbose);      /* tcgen05 GEMM vs SIMT GEMM on random operands  */
const char* mvae_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MIDIVAE_H_ */
