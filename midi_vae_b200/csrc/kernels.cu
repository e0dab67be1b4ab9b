// kernels.cu -- pointwise / reduction kernels of the MIDI-VAE path (all HBM-streaming, no reuse):
// roll expansion, LSTM gate math (forward + BPTT), latent head (KL, reparameterisation, style softmax),
// Keras-semantics losses and their gradients, tanh-Dense backward, bias reductions, Keras Adam.
// Math follows SURVEY.md appendix A; the hand-derived backward is oracle/manual_bptt.py.
#include "kernels.cuh"
#include "../../include/midivae.h"

namespace mvae {
namespace {

using bf16 = __nv_bfloat16;
constexpr int TPB = 256;
static inline int nblk(long count, int cap = 148 * 16) {
  long b = (count + TPB - 1) / TPB;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

#define DISPATCH_ACT(act, ...)                       \
  do {                                               \
    if ((act) == DT_F32) { using AT = float; __VA_ARGS__; } \
    else { using AT = bf16; __VA_ARGS__; }           \
  } while (0)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum of up to 3 doubles, one atomic per block per slot
__device__ __forceinline__ void block_atomic_add3(double a, double b, double c, double* pa, double* pb, double* pc) {
  __shared__ double sh[3][TPB / 32];
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh[0][w] = a; sh[1][w] = b; sh[2][w] = c; }
  __syncthreads();
  if (w == 0) {
    int nw = blockDim.x >> 5;
    a = l < nw ? sh[0][l] : 0; b = l < nw ? sh[1][l] : 0; c = l < nw ? sh[2][l] : 0;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (l == 0) {
      if (pa) atomicAdd(pa, a);
      if (pb) atomicAdd(pb, b);
      if (pc) atomicAdd(pc, c);
    }
  }
}

__device__ __forceinline__ float gate_fn(int gate_act, float x) {
  return gate_act == MVAE_GATE_HARD_SIGMOID ? fminf(fmaxf(0.2f * x + 0.5f, 0.f), 1.f) : 1.f / (1.f + expf(-x));
}
__device__ __forceinline__ float gate_grad(int gate_act, float s) {
  return gate_act == MVAE_GATE_HARD_SIGMOID ? ((s > 0.f && s < 1.f) ? 0.2f : 0.f) : s * (1.f - s);
}

// ------------------------------------------------------------------ roll expansion
template <typename AT>
__global__ void expand_onehot_kernel(AT* ext, int steps, int n, int pad, const uint8_t* idx) {
  long total = (long)(steps + 1) * n * pad;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int c = (int)(e % pad);
    long r = e / pad;
    int b = (int)(r % n), t = (int)(r / n);
    float v = 0.f;
    if (t > 0) v = (idx[(long)b * steps + (t - 1)] == c) ? 1.f : 0.f;
    stf<AT>(ext + e, v);
  }
}
template <typename AT>
__global__ void expand_vel_kernel(AT* ext, int steps, int n, int pad, const float* vel) {
  long total = (long)(steps + 1) * n * pad;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int c = (int)(e % pad);
    long r = e / pad;
    int b = (int)(r % n), t = (int)(r / n);
    float v = 0.f;
    if (t > 0 && c == 0) v = vel[(long)b * steps + (t - 1)];
    stf<AT>(ext + e, v);
  }
}

template <typename AT>
__global__ void fill_rows_kernel(AT* dst, long rows, int cols, const float* bias) {
  long total = rows * cols;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x)
    stf<AT>(dst + e, bias[e % cols]);
}

template <typename TS, typename TD>
__global__ void copy2d_kernel(int rows, int cols, const TS* src, int lds, TD* dst, int ldd) {
  long total = (long)rows * cols;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int c = (int)(e % cols);
    long r = e / cols;
    stf<TD>(dst + r * ldd + c, ldf<TS>(src + r * lds + c));
  }
}

// ------------------------------------------------------------------ LSTM cell math
// pre (n,4H) fp32 = x W + h U + b.  Block order [i|f|g|o] (standard) or [f|i|g|o] (recalled variant).
template <typename AT>
__global__ void cell_fwd_kernel(CellCfg cc, int n, int H, const float* __restrict__ pre, float* __restrict__ c_run, AT* __restrict__ gates,
                                AT* __restrict__ cseq1, AT* __restrict__ hseq1) {
  long total = (long)n * H;
  const int bi = cc.variant == MVAE_CELL_STANDARD ? 0 : 1, bf_ = cc.variant == MVAE_CELL_STANDARD ? 1 : 0;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int j = (int)(e % H);
    long b = e / H;
    const float* pr = pre + b * 4 * H;
    float i = gate_fn(cc.gate_act, pr[bi * H + j]);
    float f = gate_fn(cc.gate_act, pr[bf_ * H + j]);
    float g = tanhf(pr[2 * H + j]);
    float o = gate_fn(cc.gate_act, pr[3 * H + j]);
    float s = f * c_run[e] + i * g;
    float cn, hn;
    if (cc.variant == MVAE_CELL_STANDARD) { cn = s; hn = o * tanhf(s); }
    else { cn = tanhf(s); hn = o * cn; }
    c_run[e] = cn;
    AT* gt = gates + b * 4 * H;
    stf<AT>(gt + bi * H + j, i); stf<AT>(gt + bf_ * H + j, f); stf<AT>(gt + 2 * H + j, g); stf<AT>(gt + 3 * H + j, o);
    stf<AT>(cseq1 + e, cn);
    stf<AT>(hseq1 + e, hn);
  }
}

template <typename AT, typename LT>
__global__ void cell_bwd_kernel(CellCfg cc, int n, int H, const float* __restrict__ dh_run, const AT* __restrict__ dh_ext,
                                const LT* __restrict__ dh_last, int ld_last, float* __restrict__ dc_run, const AT* __restrict__ gates,
                                const AT* __restrict__ cseq0, const AT* __restrict__ cseq1, AT* __restrict__ dG) {
  long total = (long)n * H;
  const int bi = cc.variant == MVAE_CELL_STANDARD ? 0 : 1, bf_ = cc.variant == MVAE_CELL_STANDARD ? 1 : 0;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int j = (int)(e % H);
    long b = e / H;
    float dh = dh_run[e];
    if (dh_ext) dh += ldf<AT>(dh_ext + e);
    if (dh_last) dh += ldf<LT>(dh_last + b * ld_last + j);
    const AT* gt = gates + b * 4 * H;
    float i = ldf<AT>(gt + bi * H + j), f = ldf<AT>(gt + bf_ * H + j), g = ldf<AT>(gt + 2 * H + j), o = ldf<AT>(gt + 3 * H + j);
    float c_prev = ldf<AT>(cseq0 + e), c_new = ldf<AT>(cseq1 + e);
    float dc = dc_run[e];
    float d_o, ds;
    if (cc.variant == MVAE_CELL_STANDARD) {
      float tc = tanhf(c_new);
      d_o = dh * tc;
      ds = dc + dh * o * (1.f - tc * tc);
    } else {
      d_o = dh * c_new;
      ds = (dc + dh * o) * (1.f - c_new * c_new);
    }
    float di = ds * g, df = ds * c_prev, dg = ds * i;
    dc_run[e] = ds * f;
    AT* dgp = dG + b * 4 * H;
    stf<AT>(dgp + bi * H + j, di * gate_grad(cc.gate_act, i));
    stf<AT>(dgp + bf_ * H + j, df * gate_grad(cc.gate_act, f));
    stf<AT>(dgp + 2 * H + j, dg * (1.f - g * g));
    stf<AT>(dgp + 3 * H + j, d_o * gate_grad(cc.gate_act, o));
  }
}

// ------------------------------------------------------------------ GRU cell math (row f-1: the reference's shipped cell type, settings.py:155)
// Blocks [z|r|h] (Keras 2.0.8 GRU, vae_definition.py:457-472; recurrentshop GRUCell, :535,585,623).  A step is TWO dependent products:
//   pre[:, 0:2H] = x W_zr + h U_zr + b_zr -> z, r;   rh = r * h;   pre[:, 2H:3H] = x W_h + rh U_h + b_h -> hh = tanh;
//   mix 0 (Keras GRU):             h' = z h + (1 - z) hh
//   mix 1 (recurrentshop GRUCell): h' = (1 - z) h + z hh          (as recalled; the shipped decoders' first step decodes only with this one)
// Stash for the reverse sweep: gates_t (n,3H) = [z|r|hh], rh_t (n,H), h sequence.
template <typename AT>
__global__ void gru_gates_kernel(int gate_act, int n, int H, const float* __restrict__ pre, const AT* __restrict__ h_prev, AT* __restrict__ gates,
                                 AT* __restrict__ rh) {
  long total = (long)n * H;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int j = (int)(e % H);
    long b = e / H;
    const float* pr = pre + b * 3 * H;
    float z = gate_fn(gate_act, pr[j]);
    float r = gate_fn(gate_act, pr[H + j]);
    AT* gt = gates + b * 3 * H;
    stf<AT>(gt + j, z); stf<AT>(gt + H + j, r);
    stf<AT>(rh + e, r * ldf<AT>(h_prev + e));
  }
}

template <typename AT>
__global__ void gru_out_kernel(int mix, int n, int H, const float* __restrict__ pre, const AT* __restrict__ h_prev, AT* __restrict__ gates,
                               AT* __restrict__ h_new) {
  long total = (long)n * H;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int j = (int)(e % H);
    long b = e / H;
    float hh = tanhf(pre[b * 3 * H + 2 * H + j]);
    AT* gt = gates + b * 3 * H;
    float z = ldf<AT>(gt + j), h = ldf<AT>(h_prev + e);
    stf<AT>(gt + 2 * H + j, hh);
    stf<AT>(h_new + e, mix == 0 ? z * h + (1.f - z) * hh : (1.f - z) * h + z * hh);
  }
}

// reverse step, part 1: dh = dh_run + dh_ext + dh_last;  dG[:, 2H:3H] = da_h = dhh (1 - hh^2);  dG[:, 0:H] = da_z;  dh_run = the direct path to h_{t-1}
template <typename AT, typename LT>
__global__ void gru_bwd1_kernel(int gate_act, int mix, int n, int H, float* __restrict__ dh_run, const AT* __restrict__ dh_ext,
                                const LT* __restrict__ dh_last, int ld_last, const AT* __restrict__ gates, const AT* __restrict__ h_prev,
                                AT* __restrict__ dG) {
  long total = (long)n * H;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int j = (int)(e % H);
    long b = e / H;
    float dh = dh_run[e];
    if (dh_ext) dh += ldf<AT>(dh_ext + e);
    if (dh_last) dh += ldf<LT>(dh_last + b * ld_last + j);
    const AT* gt = gates + b * 3 * H;
    float z = ldf<AT>(gt + j), hh = ldf<AT>(gt + 2 * H + j), h = ldf<AT>(h_prev + e);
    float dz, dhh, direct;
    if (mix == 0) { dz = dh * (h - hh); dhh = dh * (1.f - z); direct = dh * z; }
    else { dz = dh * (hh - h); dhh = dh * z; direct = dh * (1.f - z); }
    AT* dgp = dG + b * 3 * H;
    stf<AT>(dgp + j, dz * gate_grad(gate_act, z));
    stf<AT>(dgp + 2 * H + j, dhh * (1.f - hh * hh));
    dh_run[e] = direct;
  }
}

// reverse step, part 2: drh = da_h U_h^T (fp32);  da_r = drh h_{t-1} act'(r) -> dG[:, H:2H];  dh_run += drh r
template <typename AT>
__global__ void gru_bwd2_kernel(int gate_act, int n, int H, const float* __restrict__ drh, const AT* __restrict__ gates, const AT* __restrict__ h_prev,
                                float* __restrict__ dh_run, AT* __restrict__ dG) {
  long total = (long)n * H;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int j = (int)(e % H);
    long b = e / H;
    float r = ldf<AT>(gates + b * 3 * H + H + j), h = ldf<AT>(h_prev + e), d = drh[e];
    stf<AT>(dG + b * 3 * H + H + j, d * h * gate_grad(gate_act, r));
    dh_run[e] += d * r;
  }
}

template <typename AT>
__global__ void concat3_kernel(int n, int H, const AT* a, const AT* b, const AT* c, AT* u) {
  long total = (long)n * 3 * H;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int col = (int)(e % (3 * H));
    long r = e / (3 * H);
    const AT* src = col < H ? a : (col < 2 * H ? b : c);
    u[e] = src[r * H + (col % H)];
  }
}

// ------------------------------------------------------------------ latent head
template <typename AT>
__global__ void latent_fwd_kernel(int n, int L, int ldl, const float* mu, const float* lv, const float* eps, const float* hist,
                                  int has_hist, float* z, AT* q, int ldq, float beta, float m0, float s0, double* acc) {
  long total = (long)n * L;
  double kl = 0.0;
  const float two_ln_s0 = 2.f * logf(s0), inv_var = 1.f / (s0 * s0);
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int j = (int)(e % L);
    long b = e / L;
    float m = mu[b * ldl + j], l = lv[b * ldl + j];
    float ep = eps ? eps[b * L + j] : 0.f;
    float zz = m + expf(0.5f * l) * ep;
    z[b * ldl + j] = zz;
    stf<AT>(q + b * ldq + j, zz);
    if (has_hist) stf<AT>(q + b * ldq + L + j, hist ? hist[b * L + j] : 0.f);
    kl += (double)(1.f + l - two_ln_s0 - ((m - m0) * (m - m0) + expf(l)) * inv_var);
  }
  block_atomic_add3(-0.5 * (double)beta * kl, 0, 0, acc + ACC_KL, nullptr, nullptr);
}

__global__ void style_head_kernel(int n, int C, const float* z, int ldl, const uint8_t* style, float* probs, double* acc) {
  double ce = 0, ac = 0;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += gridDim.x * blockDim.x) {
    const float* zr = z + (long)b * ldl;
    float mx = zr[0];
    for (int c = 1; c < C; ++c) mx = fmaxf(mx, zr[c]);
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum += expf(zr[c] - mx);
    int arg = 0; float best = -1.f;
    for (int c = 0; c < C; ++c) {
      float p = expf(zr[c] - mx) / sum;
      probs[(long)b * C + c] = p;
      if (p > best) { best = p; arg = c; }
    }
    if (style) {
      int y = style[b];
      float py = probs[(long)b * C + y];
      ce += (double)(-logf(fminf(fmaxf(py, 1e-7f), 1.f - 1e-7f)));
      ac += (arg == y) ? 1.0 : 0.0;
    }
  }
  block_atomic_add3(ce, ac, 0, acc + ACC_CE_STYLE, acc + ACC_ACC_STYLE, nullptr);
}

template <typename AT>
__global__ void latent_bwd_kernel(int n, int L, int ldl, int C, const AT* dq, int ldq, const float* mu, const float* lv, const float* eps,
                                  const float* sp, const uint8_t* style, float beta, float m0, float s0, float style_w, AT* dmu, AT* dlv) {
  long total = (long)n * L;
  const float inv_var = 1.f / (s0 * s0), inv_n = 1.f / n;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int j = (int)(e % L);
    long b = e / L;
    float dz = ldf<AT>(dq + b * ldq + j);
    if (j < C) {
      int y = style[b];
      float py = sp[b * C + y];
      float live = (py > 1e-7f && py < 1.f - 1e-7f) ? 1.f : 0.f;
      dz += style_w * inv_n * (sp[b * C + j] - (j == y ? 1.f : 0.f)) * live;
    }
    float m = mu[b * ldl + j], l = lv[b * ldl + j];
    float ep = eps ? eps[b * L + j] : 0.f;
    float dm = dz + beta * (m - m0) * inv_var * inv_n;
    float dl = dz * ep * 0.5f * expf(0.5f * l) + beta * (-0.5f) * (1.f - expf(l) * inv_var) * inv_n;
    stf<AT>(dmu + b * ldl + j, dm);
    stf<AT>(dlv + b * ldl + j, dl);
  }
}

// ------------------------------------------------------------------ losses
__global__ void count_nonzero_kernel(const float* w, long count, double* acc) {
  double c = 0;
  if (w) {
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < count; e += (long)gridDim.x * blockDim.x) c += (w[e] != 0.f) ? 1.0 : 0.0;
  } else if (blockIdx.x == 0 && threadIdx.x == 0) {
    c = (double)count;
  }
  block_atomic_add3(c, 0, 0, acc + ACC_WNZ, nullptr, nullptr);
}

// one warp per (t,b) row; D <= 64 classes.  Keras categorical_crossentropy on softmax outputs with the temporal
// sample-weight normaliser; gradient = loss_w * w / count_nonzero(w) * (p - onehot), zero where clipped.
template <typename AT>
__global__ void softmax_ce_kernel(int steps, int n, int D, float* logits, int ld, const uint8_t* labels, const float* w,
                                  const double* acc_wnz, float loss_w, AT* dlogits, int ldd, double* acc, int slot_ce, int slot_acc) {
  const int lane = threadIdx.x & 31;
  const long rows = (long)steps * n;
  const long warp0 = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  double ce_sum = 0, acc_sum = 0;
  const float denom = acc_wnz ? (float)(*acc_wnz) : (float)rows;
  for (long r = warp0; r < rows; r += nwarps) {
    int b = (int)(r % n), t = (int)(r / n);
    float* lr = logits + r * ld;
    float x0 = lane < D ? lr[lane] : -INFINITY, x1 = (lane + 32) < D ? lr[lane + 32] : -INFINITY;
    float mx = fmaxf(x0, x1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float e0 = lane < D ? expf(x0 - mx) : 0.f, e1 = (lane + 32) < D ? expf(x1 - mx) : 0.f;
    float sum = e0 + e1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    float p0 = e0 / sum, p1 = e1 / sum;
    if (lane < D) lr[lane] = p0;
    if (lane + 32 < D) lr[lane + 32] = p1;
    // argmax, lowest index wins ties
    float bv = p0; int bi = lane;
    if (lane + 32 < D && p1 > bv) { bv = p1; bi = lane + 32; }
    if (lane >= D) { bv = -1.f; bi = 1 << 20; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (labels) {
      int y = min((int)labels[(long)b * steps + t], 63);      // device-pointer callers are not range-checked on the host: stay inside the 64-wide row
      float py = __shfl_sync(0xffffffffu, y < 32 ? p0 : p1, y & 31);
      float wt = w ? w[(long)b * steps + t] : 1.f;
      if (lane == 0) {
        ce_sum += (double)(-logf(fminf(fmaxf(py, 1e-7f), 1.f - 1e-7f)) * wt);
        acc_sum += (bi == y) ? 1.0 : 0.0;      // Keras 2.0.8 metrics are plain means: sample weights do not enter
      }
      if (dlogits) {
        float live = (py > 1e-7f && py < 1.f - 1e-7f) ? 1.f : 0.f;
        float sc = loss_w * wt / denom * live;
        AT* dr = dlogits + r * ldd;
        if (lane < ldd) stf<AT>(dr + lane, lane < D ? sc * (p0 - (lane == y ? 1.f : 0.f)) : 0.f);
        if (lane + 32 < ldd) stf<AT>(dr + lane + 32, (lane + 32) < D ? sc * (p1 - ((lane + 32) == y ? 1.f : 0.f)) : 0.f);
      }
    }
  }
  block_atomic_add3(ce_sum, acc_sum, 0, acc + slot_ce, acc + slot_acc, nullptr);
}

template <typename AT>
__global__ void sigmoid_mse_kernel(int steps, int n, float* logits, int ld, const float* target, float loss_w, AT* dlogits, int ldd,
                                   double* acc) {
  const long rows = (long)steps * n;
  double se = 0, ac = 0;
  for (long r = blockIdx.x * (long)blockDim.x + threadIdx.x; r < rows; r += (long)gridDim.x * blockDim.x) {
    int b = (int)(r % n), t = (int)(r / n);
    float p = 1.f / (1.f + expf(-logits[r * ld]));
    logits[r * ld] = p;
    if (target) {
      float v = target[(long)b * steps + t];
      se += (double)((p - v) * (p - v));
      ac += (rintf(p) == v) ? 1.0 : 0.0;
      if (dlogits) {
        AT* dr = dlogits + r * ldd;
        stf<AT>(dr, loss_w * 2.f * (p - v) / (float)rows * p * (1.f - p));
        for (int c = 1; c < ldd; ++c) stf<AT>(dr + c, 0.f);
      }
    }
  }
  block_atomic_add3(se, ac, 0, acc + ACC_MSE_VEL, acc + ACC_ACC_VEL, nullptr);
}

template <typename AT>
__global__ void tanh_bwd_kernel(long count, const AT* dout, const AT* out, AT* dpre) {
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < count; e += (long)gridDim.x * blockDim.x) {
    float o = ldf<AT>(out + e);
    stf<AT>(dpre + e, ldf<AT>(dout + e) * (1.f - o * o));
  }
}

// column sums of a tall matrix (bias gradients, rank-1 weight gradients): each thread owns 8 consecutive columns
// (one 16-byte load per row for bf16), a block covers 256 columns x 8 row lanes; grid (ceil(cols/256), row chunks).
// These passes run next to the cluster recurrences on the few SMs those leave free, so what bounds them is bytes in flight per SM, not
// bandwidth: the vector path keeps CS_UNROLL independent 16-byte loads per thread in the air (round 2 had one: 134 MB took 0.66 ms).
// wsum (weighted form only): also accumulates sum_r weight[r] -- the bias gradient of a one-column head -- instead of a second launch.
constexpr int CS_UNROLL = 8;
template <typename AT>
__global__ void colsum_kernel(long rows, int cols, int ld, const AT* __restrict__ src, const AT* __restrict__ weight, int ldw, float* dst,
                              float* wsum) {
  __shared__ float sh[8][256 + 8];
  __shared__ float shw[8];
  const int c0 = blockIdx.x * 256 + threadIdx.x * 8;
  float s[8];
  float wacc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  const bool vec = (c0 + 8 <= cols) && (ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && sizeof(AT) == 2;
  const bool wown = wsum && weight && blockIdx.x == 0 && threadIdx.x == 0;   // one thread per row lane sums the weights
  if (c0 < cols) {
    const long stride = (long)gridDim.y * 8;
    long r = blockIdx.y * 8 + threadIdx.y;
    if (vec) {
      for (; r + (CS_UNROLL - 1) * stride < rows; r += CS_UNROLL * stride) {
        uint4 u[CS_UNROLL];
        float w[CS_UNROLL];
#pragma unroll
        for (int k = 0; k < CS_UNROLL; ++k) {
          u[k] = __ldg(reinterpret_cast<const uint4*>(src + (r + k * stride) * ld + c0));
          w[k] = weight ? ldf<AT>(weight + (r + k * stride) * ldw) : 1.f;
        }
#pragma unroll
        for (int k = 0; k < CS_UNROLL; ++k) {
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u[k]);
#pragma unroll
          for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h2[i]); s[2 * i] += w[k] * f.x; s[2 * i + 1] += w[k] * f.y; }
          wacc += w[k];
        }
      }
    }
    for (; r < rows; r += stride) {
      const float wgt = weight ? ldf<AT>(weight + r * ldw) : 1.f;
      wacc += wgt;
      if (vec) {
        uint4 u = __ldg(reinterpret_cast<const uint4*>(src + r * ld + c0));
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h2[i]); s[2 * i] += wgt * f.x; s[2 * i + 1] += wgt * f.y; }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) if (c0 + i < cols) s[i] += wgt * ldf<AT>(src + r * ld + c0 + i);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sh[threadIdx.y][threadIdx.x * 8 + i] = s[i];
  if (threadIdx.x == 0) shw[threadIdx.y] = wown ? wacc : 0.f;
  __syncthreads();
  const int t = threadIdx.y * 32 + threadIdx.x;   // 0..255: one column each
  const int c = blockIdx.x * 256 + t;
  if (c < cols) {
    float tsum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tsum += sh[i][t];
    atomicAdd(dst + c, tsum);
  }
  if (wsum && weight && blockIdx.x == 0 && t == 0) {
    float tsum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tsum += shw[i];
    atomicAdd(wsum, tsum);
  }
}

// One pass over a tall dG (rows = T n, time-major) for a recurrence whose input is a scalar (velocity): bias gradient db[c] += sum_r dG[r, c]
// and dW[0, c] += sum_r x[r] dG[r, c] (x act-typed with stride ldx) together -- round 1 read dG once per output.  Block (32, 8): a thread owns 8
// consecutive columns (one 16-byte load per row) for the rows of its lane, WR_UNROLL rows in flight.  (A one-hot variant that scattered dG rows
// into a shared-memory class table was measured 4x SLOWER than the padded tensor-core GEMM it replaced: shared-memory float adds are CAS loops
// (ATOMS.CAST.SPIN); dropped.)
constexpr int WR_UNROLL = 4;
template <typename AT>
__global__ void wgrad_rows_kernel(long rows, int cols, int ld, const AT* __restrict__ src, const AT* __restrict__ x, int ldx, float* __restrict__ dW,
                                  float* __restrict__ db) {
  __shared__ float sh[2][8][256 + 8];
  const int c0 = blockIdx.x * 256 + threadIdx.x * 8;
  float s[8], sx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; sx[i] = 0.f; }
  if (c0 < cols) {
    const long stride = (long)gridDim.y * 8;
    long r = blockIdx.y * 8 + threadIdx.y;
    for (; r + (WR_UNROLL - 1) * stride < rows; r += WR_UNROLL * stride) {
      uint4 u[WR_UNROLL];
      float xv[WR_UNROLL];
#pragma unroll
      for (int k = 0; k < WR_UNROLL; ++k) {
        u[k] = __ldg(reinterpret_cast<const uint4*>(src + (r + k * stride) * ld + c0));
        xv[k] = ldf<AT>(x + (r + k * stride) * ldx);
      }
#pragma unroll
      for (int k = 0; k < WR_UNROLL; ++k) {
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u[k]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 t2 = __bfloat1622float2(h2[i]);
          s[2 * i] += t2.x; s[2 * i + 1] += t2.y;
          sx[2 * i] += xv[k] * t2.x; sx[2 * i + 1] += xv[k] * t2.y;
        }
      }
    }
    for (; r < rows; r += stride) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + r * ld + c0));
      const float xv = ldf<AT>(x + r * ldx);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 t2 = __bfloat1622float2(h2[i]);
        s[2 * i] += t2.x; s[2 * i + 1] += t2.y;
        sx[2 * i] += xv * t2.x; sx[2 * i + 1] += xv * t2.y;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { sh[0][threadIdx.y][threadIdx.x * 8 + i] = s[i]; sh[1][threadIdx.y][threadIdx.x * 8 + i] = sx[i]; }
  __syncthreads();
  const int t = threadIdx.y * 32 + threadIdx.x;   // 0..255: one column each
  const int c = blockIdx.x * 256 + t;
  if (c < cols) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a += sh[0][i][t]; b += sh[1][i][t]; }
    atomicAdd(db + c, a);
    atomicAdd(dW + c, b);
  }
}

// out[r, c] = x[r] * w[c] + bias[c]: grid-stride over rows, each thread writes 8 consecutive columns per row
template <typename AT>
__global__ void rank1_rows_kernel(AT* __restrict__ out, long rows, int cols, const AT* __restrict__ x, int ldx, const float* __restrict__ w,
                                  const float* __restrict__ bias) {
  const int cthreads = (cols + 7) / 8;                 // threads along the columns
  const int tcol = threadIdx.x % cthreads, trow = threadIdx.x / cthreads, rpb = blockDim.x / cthreads;
  const int c0 = tcol * 8;
  if (trow >= rpb) return;
  float wv[8], bv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { wv[i] = (c0 + i < cols) ? w[c0 + i] : 0.f; bv[i] = (bias && c0 + i < cols) ? bias[c0 + i] : 0.f; }
  const bool vec = (c0 + 8 <= cols) && (cols % 8 == 0) && sizeof(AT) == 2 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  for (long r = (long)blockIdx.x * rpb + trow; r < rows; r += (long)gridDim.x * rpb) {
    const float xv = ldf<AT>(x + r * ldx);
    if (vec) {
      uint4 u;
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) h2[i] = __floats2bfloat162_rn(xv * wv[2 * i] + bv[2 * i], xv * wv[2 * i + 1] + bv[2 * i + 1]);
      *reinterpret_cast<uint4*>(out + r * cols + c0) = u;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) if (c0 + i < cols) stf<AT>(out + r * cols + c0 + i, xv * wv[i] + bv[i]);
    }
  }
}

// one warp per row
template <typename AT>
__global__ void rowdot_kernel(long rows, int H, const AT* h, const float* w, const float* b, float* out, int ldo) {
  const int lane = threadIdx.x & 31;
  const long warp0 = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long r = warp0; r < rows; r += nwarps) {
    float s = 0.f;
    for (int j = lane; j < H; j += 32) s = fmaf(ldf<AT>(h + r * H + j), w[j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[r * ldo] = s + b[0];
  }
}

// keras.optimizers.Adam 2.0.8: p -= lr_t * m / (sqrt(v) + eps), lr_t carries the bias corrections
__global__ void adam_kernel(long count, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            float lr_t, float b1, float b2, float eps, float gscale, bf16* __restrict__ shadow) {
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < count; e += (long)gridDim.x * blockDim.x) {
    float gg = g[e] * gscale;
    float mm = b1 * m[e] + (1.f - b1) * gg;
    float vv = b2 * v[e] + (1.f - b2) * gg * gg;
    float pp = p[e] - lr_t * mm / (sqrtf(vv) + eps);
    m[e] = mm; v[e] = vv; p[e] = pp;
    if (shadow) shadow[e] = __float2bfloat16_rn(pp);
  }
}

__global__ void f32_to_bf16_kernel(long count, const float* src, bf16* dst) {
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < count; e += (long)gridDim.x * blockDim.x) dst[e] = __float2bfloat16_rn(src[e]);
}

__global__ void finalize_metrics_kernel(const double* acc, int n, int T, int Ti, float w_notes, float w_instr, float w_vel, float w_style,
                                        float* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double wnz = acc[ACC_WNZ] > 0 ? acc[ACC_WNZ] : 1.0;
  double l_notes = acc[ACC_CE_NOTES] / wnz, l_instr = acc[ACC_CE_INSTR] / ((double)Ti * n), l_vel = acc[ACC_MSE_VEL] / ((double)T * n);
  double l_style = acc[ACC_CE_STYLE] / n, kl = acc[ACC_KL] / n;
  out[MVAE_M_LOSS] = (float)(w_notes * l_notes + w_instr * l_instr + w_vel * l_vel + w_style * l_style + kl);
  out[MVAE_M_NOTES_LOSS] = (float)l_notes; out[MVAE_M_INSTR_LOSS] = (float)l_instr; out[MVAE_M_VEL_LOSS] = (float)l_vel;
  out[MVAE_M_STYLE_LOSS] = (float)l_style;
  out[MVAE_M_NOTES_ACC] = (float)(acc[ACC_ACC_NOTES] / ((double)T * n));
  out[MVAE_M_INSTR_ACC] = (float)(acc[ACC_ACC_INSTR] / ((double)Ti * n));
  out[MVAE_M_VEL_ACC] = (float)(acc[ACC_ACC_VEL] / ((double)T * n));
  out[MVAE_M_STYLE_ACC] = (float)(acc[ACC_ACC_STYLE] / n);
  out[MVAE_M_KL] = (float)kl;
}

__global__ void export_seq_kernel(int steps, int n, int D, const float* probs, int ld, float* out) {
  long total = (long)steps * n * D;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int d = (int)(e % D);
    long r = e / D;
    int t = (int)(r % steps);
    long b = r / steps;
    out[e] = probs[((long)t * n + b) * ld + d];
  }
}

__global__ void argmax_seq_kernel(int steps, int n, int D, const float* probs, int ld, uint8_t* out) {
  long rows = (long)steps * n;
  for (long r = blockIdx.x * (long)blockDim.x + threadIdx.x; r < rows; r += (long)gridDim.x * blockDim.x) {
    int b = (int)(r % n), t = (int)(r / n);
    const float* pr = probs + r * ld;
    float best = pr[0]; int arg = 0;
    for (int d = 1; d < D; ++d) { float v = pr[d]; if (v > best) { best = v; arg = d; } }
    out[(long)b * steps + t] = (uint8_t)arg;
  }
}

template <typename AT>
__global__ void swap_shift_kernel(int n, int L, int ldl, const float* mu, const uint8_t* song_start, int c_from, int c_to, int has_hist,
                                  AT* q, int ldq, float* z_sw) {
  long total = (long)n * L;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int j = (int)(e % L);
    long b = e / L;
    int src = j == c_from ? c_to : (j == c_to ? c_from : j);
    float zz = mu[b * ldl + src];
    if (z_sw) z_sw[b * ldl + j] = zz;
    stf<AT>(q + b * ldq + j, zz);
    if (has_hist) {
      bool first = (b == 0) || (song_start && song_start[b]);
      stf<AT>(q + b * ldq + L + j, first ? 0.f : mu[(b - 1) * ldl + src]);
    }
  }
}

// history latents taken from the batch itself (opt-in; vae_training.py:791-798 takes them from a separate encoder.predict of the same song):
// q[b, L + j] = z[b - 1, j] inside a song, 0 on a song's first chunk; row 0 continues the previous call's last row (carry) unless it starts a song.
template <typename AT>
__global__ void self_history_kernel(int n, int L, int ldl, const float* __restrict__ z, const uint8_t* __restrict__ song_start, const float* __restrict__ carry,
                                    int carry_valid, AT* q, int ldq) {
  long total = (long)n * L;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int j = (int)(e % L);
    long b = e / L;
    const bool first = song_start && song_start[b];
    float h = 0.f;
    if (!first) h = b > 0 ? z[(b - 1) * ldl + j] : (carry_valid ? carry[j] : 0.f);
    stf<AT>(q + b * ldq + L + j, h);
  }
}

template <typename AT>
__global__ void build_q_kernel(int n, int L, const float* z, const float* hist, int has_hist, AT* q, int ldq) {
  long total = (long)n * L;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int j = (int)(e % L);
    long b = e / L;
    stf<AT>(q + b * ldq + j, z[e]);
    if (has_hist) stf<AT>(q + b * ldq + L + j, hist ? hist[e] : 0.f);
  }
}

#define LAUNCH_CHECK() do { count_launch(); MVAE_CUDA(cudaGetLastError()); } while (0)

}  // namespace

void k_expand_inputs(DT act, int n, int T, int Ti, int PD, int ID, int VD, const uint8_t* pitch, const uint8_t* target, const uint8_t* instr,
                     const float* vel, void* Xp_ext, void* Yp_ext, void* Xi_ext, void* Xv_ext, cudaStream_t st) {
  DISPATCH_ACT(act, {
    if (Xp_ext) { expand_onehot_kernel<AT><<<nblk((long)(T + 1) * n * PD), TPB, 0, st>>>((AT*)Xp_ext, T, n, PD, pitch); LAUNCH_CHECK(); }
    if (Yp_ext) { expand_onehot_kernel<AT><<<nblk((long)(T + 1) * n * PD), TPB, 0, st>>>((AT*)Yp_ext, T, n, PD, target); LAUNCH_CHECK(); }
    if (Xi_ext) { expand_onehot_kernel<AT><<<nblk((long)(Ti + 1) * n * ID), TPB, 0, st>>>((AT*)Xi_ext, Ti, n, ID, instr); LAUNCH_CHECK(); }
    if (Xv_ext) { expand_vel_kernel<AT><<<nblk((long)(T + 1) * n * VD), TPB, 0, st>>>((AT*)Xv_ext, T, n, VD, vel); LAUNCH_CHECK(); }
  });
}

void k_fill_rows(DT act, void* dst, long rows, int cols, const float* bias, cudaStream_t st) {
  DISPATCH_ACT(act, { fill_rows_kernel<AT><<<nblk(rows * cols), TPB, 0, st>>>((AT*)dst, rows, cols, bias); LAUNCH_CHECK(); });
}

void k_copy2d(DT src_t, DT dst_t, int rows, int cols, const void* src, int lds, void* dst, int ldd, cudaStream_t st) {
  int g = nblk((long)rows * cols);
  if (src_t == DT_F32 && dst_t == DT_F32) copy2d_kernel<float, float><<<g, TPB, 0, st>>>(rows, cols, (const float*)src, lds, (float*)dst, ldd);
  else if (src_t == DT_F32) copy2d_kernel<float, bf16><<<g, TPB, 0, st>>>(rows, cols, (const float*)src, lds, (bf16*)dst, ldd);
  else if (dst_t == DT_F32) copy2d_kernel<bf16, float><<<g, TPB, 0, st>>>(rows, cols, (const bf16*)src, lds, (float*)dst, ldd);
  else copy2d_kernel<bf16, bf16><<<g, TPB, 0, st>>>(rows, cols, (const bf16*)src, lds, (bf16*)dst, ldd);
  LAUNCH_CHECK();
}

void k_cell_fwd(DT act, CellCfg cc, int n, int H, const float* pre, float* c_run, void* gates_t, void* cseq_t1, void* hseq_t1, cudaStream_t st) {
  DISPATCH_ACT(act, {
    cell_fwd_kernel<AT><<<nblk((long)n * H), TPB, 0, st>>>(cc, n, H, pre, c_run, (AT*)gates_t, (AT*)cseq_t1, (AT*)hseq_t1);
    LAUNCH_CHECK();
  });
}

void k_cell_bwd(DT act, CellCfg cc, int n, int H, const float* dh_run, const void* dh_ext_t, const void* dh_last, int ld_last, DT last_t,
                float* dc_run, const void* gates_t, const void* cseq_t, const void* cseq_t1, void* dG_t, cudaStream_t st) {
  DISPATCH_ACT(act, {
    if (last_t == DT_F32)
      cell_bwd_kernel<AT, float><<<nblk((long)n * H), TPB, 0, st>>>(cc, n, H, dh_run, (const AT*)dh_ext_t, (const float*)dh_last, ld_last, dc_run,
                                                                    (const AT*)gates_t, (const AT*)cseq_t, (const AT*)cseq_t1, (AT*)dG_t);
    else
      cell_bwd_kernel<AT, bf16><<<nblk((long)n * H), TPB, 0, st>>>(cc, n, H, dh_run, (const AT*)dh_ext_t, (const bf16*)dh_last, ld_last, dc_run,
                                                                   (const AT*)gates_t, (const AT*)cseq_t, (const AT*)cseq_t1, (AT*)dG_t);
    LAUNCH_CHECK();
  });
}

void k_gru_gates(DT act, int gate_act, int n, int H, const float* pre, const void* h_prev, void* gates_t, void* rh_t, cudaStream_t st) {
  DISPATCH_ACT(act, { gru_gates_kernel<AT><<<nblk((long)n * H), TPB, 0, st>>>(gate_act, n, H, pre, (const AT*)h_prev, (AT*)gates_t, (AT*)rh_t); LAUNCH_CHECK(); });
}
void k_gru_out(DT act, int mix, int n, int H, const float* pre, const void* h_prev, void* gates_t, void* h_new, cudaStream_t st) {
  DISPATCH_ACT(act, { gru_out_kernel<AT><<<nblk((long)n * H), TPB, 0, st>>>(mix, n, H, pre, (const AT*)h_prev, (AT*)gates_t, (AT*)h_new); LAUNCH_CHECK(); });
}
void k_gru_bwd1(DT act, int gate_act, int mix, int n, int H, float* dh_run, const void* dh_ext_t, const void* dh_last, int ld_last, DT last_t,
                const void* gates_t, const void* h_prev, void* dG_t, cudaStream_t st) {
  DISPATCH_ACT(act, {
    if (last_t == DT_F32)
      gru_bwd1_kernel<AT, float><<<nblk((long)n * H), TPB, 0, st>>>(gate_act, mix, n, H, dh_run, (const AT*)dh_ext_t, (const float*)dh_last, ld_last,
                                                                    (const AT*)gates_t, (const AT*)h_prev, (AT*)dG_t);
    else
      gru_bwd1_kernel<AT, bf16><<<nblk((long)n * H), TPB, 0, st>>>(gate_act, mix, n, H, dh_run, (const AT*)dh_ext_t, (const bf16*)dh_last, ld_last,
                                                                   (const AT*)gates_t, (const AT*)h_prev, (AT*)dG_t);
    LAUNCH_CHECK();
  });
}
void k_gru_bwd2(DT act, int gate_act, int n, int H, const float* drh, const void* gates_t, const void* h_prev, float* dh_run, void* dG_t, cudaStream_t st) {
  DISPATCH_ACT(act, { gru_bwd2_kernel<AT><<<nblk((long)n * H), TPB, 0, st>>>(gate_act, n, H, drh, (const AT*)gates_t, (const AT*)h_prev, dh_run, (AT*)dG_t); LAUNCH_CHECK(); });
}

void k_concat3(DT act, int n, int H, const void* a, const void* b, const void* c, void* u, cudaStream_t st) {
  DISPATCH_ACT(act, { concat3_kernel<AT><<<nblk((long)n * 3 * H), TPB, 0, st>>>(n, H, (const AT*)a, (const AT*)b, (const AT*)c, (AT*)u); LAUNCH_CHECK(); });
}

void k_latent_fwd(DT act, int n, int L, int ldl, const float* mu, const float* lv, const float* eps, const float* hist, int has_hist, float* z,
                  void* q, int ldq, float beta, float m0, float s0, double* acc, cudaStream_t st) {
  DISPATCH_ACT(act, {
    latent_fwd_kernel<AT><<<nblk((long)n * L), TPB, 0, st>>>(n, L, ldl, mu, lv, eps, hist, has_hist, z, (AT*)q, ldq, beta, m0, s0, acc);
    LAUNCH_CHECK();
  });
}

void k_style_head(int n, int C, const float* z, int ldl, const uint8_t* style, float* probs, double* acc, cudaStream_t st) {
  style_head_kernel<<<nblk(n), TPB, 0, st>>>(n, C, z, ldl, style, probs, acc);
  LAUNCH_CHECK();
}

void k_latent_bwd(DT act, int n, int L, int ldl, int C, const void* dq, int ldq, const float* mu, const float* lv, const float* eps,
                  const float* style_probs, const uint8_t* style, float beta, float m0, float s0, float style_w, void* dmu, void* dlv,
                  cudaStream_t st) {
  DISPATCH_ACT(act, {
    latent_bwd_kernel<AT><<<nblk((long)n * L), TPB, 0, st>>>(n, L, ldl, C, (const AT*)dq, ldq, mu, lv, eps, style_probs, style, beta, m0, s0,
                                                             style_w, (AT*)dmu, (AT*)dlv);
    LAUNCH_CHECK();
  });
}

void k_count_nonzero(const float* w, long count, double* acc, cudaStream_t st) {
  count_nonzero_kernel<<<w ? nblk(count, 148) : 1, TPB, 0, st>>>(w, count, acc);
  LAUNCH_CHECK();
}

void k_softmax_ce(DT act, int steps, int n, int D, float* logits, int ld, const uint8_t* labels, const float* w, const double* acc_wnz,
                  float loss_w, void* dlogits, int ldd, double* acc, int slot_ce, int slot_acc, cudaStream_t st) {
  MVAE_REQUIRE(D <= 64 && ldd <= 64, "softmax head supports at most 64 classes");
  long rows = (long)steps * n;
  int g = nblk(rows * 32);
  DISPATCH_ACT(act, {
    softmax_ce_kernel<AT><<<g, TPB, 0, st>>>(steps, n, D, logits, ld, labels, w, acc_wnz, loss_w, (AT*)dlogits, ldd, acc, slot_ce, slot_acc);
    LAUNCH_CHECK();
  });
}

void k_sigmoid_mse(DT act, int steps, int n, float* logits, int ld, const float* target, float loss_w, void* dlogits, int ldd, double* acc,
                   cudaStream_t st) {
  DISPATCH_ACT(act, {
    sigmoid_mse_kernel<AT><<<nblk((long)steps * n), TPB, 0, st>>>(steps, n, logits, ld, target, loss_w, (AT*)dlogits, ldd, acc);
    LAUNCH_CHECK();
  });
}

void k_tanh_bwd(DT act, long count, const void* dout, const void* out, void* dpre, cudaStream_t st) {
  DISPATCH_ACT(act, { tanh_bwd_kernel<AT><<<nblk(count), TPB, 0, st>>>(count, (const AT*)dout, (const AT*)out, (AT*)dpre); LAUNCH_CHECK(); });
}

void k_rank1_rows(DT act, void* out, long rows, int cols, const void* x, int ldx, const float* w, const float* bias, cudaStream_t st) {
  MVAE_REQUIRE((cols + 7) / 8 <= 1024, "rank-1 row kernel: too many columns");
  const int cthreads = (cols + 7) / 8;
  const int threads = cthreads >= 256 ? cthreads : (256 / cthreads) * cthreads;   // whole rows per block
  const int rpb = threads / cthreads;
  long blocks = (rows + rpb - 1) / rpb;
  if (blocks > 148 * 8) blocks = 148 * 8;
  DISPATCH_ACT(act, { rank1_rows_kernel<AT><<<(int)(blocks < 1 ? 1 : blocks), threads, 0, st>>>((AT*)out, rows, cols, (const AT*)x, ldx, w, bias); LAUNCH_CHECK(); });
}

void k_rowdot(DT act, long rows, int H, const void* h, const float* w, const float* b, float* out, int ldo, cudaStream_t st) {
  DISPATCH_ACT(act, { rowdot_kernel<AT><<<nblk(rows * 32), TPB, 0, st>>>(rows, H, (const AT*)h, w, b, out, ldo); LAUNCH_CHECK(); });
}

void k_colsum(DT act, long rows, int cols, int ld, const void* src, const void* weight, int ldw, float* dst, cudaStream_t st, float* wsum) {
  long chunks = (rows + 127) / 128;
  const unsigned gx = (cols + 255) / 256;
  const long cap = (148 * 8 + gx - 1) / gx;
  dim3 grid(gx, (unsigned)(chunks < 1 ? 1 : (chunks > cap ? cap : chunks)));
  dim3 block(32, 8);
  DISPATCH_ACT(act, { colsum_kernel<AT><<<grid, block, 0, st>>>(rows, cols, ld, (const AT*)src, (const AT*)weight, ldw, dst, wsum); LAUNCH_CHECK(); });
}

void k_wgrad_rows(long rows, int cols, int ld, const void* src, const void* x, int ldx, float* dW, float* db, cudaStream_t st) {
  using AT = __nv_bfloat16;   // bf16 activations only (16-byte row pieces); callers keep the generic column sums for fp32
  long chunks = (rows + 127) / 128;
  const unsigned gx = (cols + 255) / 256;
  const long cap = (148 * 8 + gx - 1) / gx;
  dim3 grid(gx, (unsigned)(chunks < 1 ? 1 : (chunks > cap ? cap : chunks)));
  dim3 block(32, 8);
  wgrad_rows_kernel<AT><<<grid, block, 0, st>>>(rows, cols, ld, (const AT*)src, (const AT*)x, ldx, dW, db);
  LAUNCH_CHECK();
}

void k_adam(long count, float* p, const float* g, float* m, float* v, float lr_t, float b1, float b2, float eps, float gscale,
            __nv_bfloat16* shadow, cudaStream_t st) {
  adam_kernel<<<nblk(count), TPB, 0, st>>>(count, p, g, m, v, lr_t, b1, b2, eps, gscale, shadow);
  LAUNCH_CHECK();
}

void k_f32_to_bf16(long count, const float* src, __nv_bfloat16* dst, cudaStream_t st) {
  f32_to_bf16_kernel<<<nblk(count), TPB, 0, st>>>(count, src, dst);
  LAUNCH_CHECK();
}

void k_finalize_metrics(const double* acc, int n, int T, int Ti, float w_notes, float w_instr, float w_vel, float w_style, float* out,
                        cudaStream_t st) {
  finalize_metrics_kernel<<<1, 32, 0, st>>>(acc, n, T, Ti, w_notes, w_instr, w_vel, w_style, out);
  LAUNCH_CHECK();
}

// Output post-processing of the decoder (vae_definition.py:1156-1190, 1214-1221) on the packed device outputs:
// silent steps get velocity 0; override_sampled_pitches_based_on_velocity_info walks every voice (flattened step s belongs to voice
// s % max_voices) through time with a (previous pitch, previous struck velocity) memory; the held-note roll is 1 unless the final
// velocity says "struck".  One thread per (scan unit, voice): scope 1 = the memory restarts at every chunk (the reference's style-switch
// loop post-processes chunk by chunk, vae_evaluation.py:2483), scope 2 = at every song start (its whole-song call sites, :799, :814).
__global__ void postprocess_voices_kernel(int n, int T, int voices, int silent, float thr, int scope, int do_override, const uint8_t* __restrict__ pitch,
                                          const uint8_t* __restrict__ song_start, float* __restrict__ vel, uint8_t* __restrict__ held) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n * voices) return;
  const int chunk = x / voices, voice = x % voices;
  int last = chunk;                                      // scan chunks [chunk, last]
  if (scope == 2) {
    if (chunk > 0 && !(song_start && song_start[chunk])) return;       // not the first chunk of a song: scanned by the thread of its song start
    while (last + 1 < n && !(song_start && song_start[last + 1])) ++last;
  }
  int prev_pitch = -1;
  float prev_vel = 0.f;
  for (int c = chunk; c <= last; ++c) {
    for (int i = voice; i < T; i += voices) {
      const size_t s = (size_t)c * T + i;
      const int pc = pitch[s];
      const int p = pc == silent ? -1 : pc;
      float v = p < 0 ? 0.f : vel[s];                    // a silent step has velocity 0
      float out = v;
      if (do_override) {
        const bool vel_silent = v < thr;
        if (vel_silent) {
          if (p >= 0 && prev_pitch > 0 && prev_pitch != p) out = prev_vel;   // a new pitch without a struck velocity: as loud as the previous note
        } else if (p < 0) {
          out = 0.f;
        }
        prev_pitch = p;
        if (!vel_silent) prev_vel = v;
      }
      vel[s] = out;
      if (held) held[s] = out > thr ? 0 : 1;
    }
  }
}

void k_export_seq(int steps, int n, int D, const float* probs, int ld, float* out, cudaStream_t st) {
  export_seq_kernel<<<nblk((long)steps * n * D), TPB, 0, st>>>(steps, n, D, probs, ld, out);
  LAUNCH_CHECK();
}

void k_argmax_seq(int steps, int n, int D, const float* probs, int ld, uint8_t* out, cudaStream_t st) {
  argmax_seq_kernel<<<nblk((long)steps * n), TPB, 0, st>>>(steps, n, D, probs, ld, out);
  LAUNCH_CHECK();
}

void k_postprocess_voices(int n, int T, int voices, int silent, float thr, int scope, int do_override, const uint8_t* pitch, const uint8_t* song_start,
                          float* vel, uint8_t* held, cudaStream_t st) {
  postprocess_voices_kernel<<<nblk((long)n * voices), TPB, 0, st>>>(n, T, voices, silent, thr, scope, do_override, pitch, song_start, vel, held);
  LAUNCH_CHECK();
}

void k_swap_shift(DT act, int n, int L, int ldl, const float* mu, const uint8_t* song_start, int c_from, int c_to, int has_hist, void* q, int ldq,
                  float* z_sw, cudaStream_t st) {
  DISPATCH_ACT(act, {
    swap_shift_kernel<AT><<<nblk((long)n * L), TPB, 0, st>>>(n, L, ldl, mu, song_start, c_from, c_to, has_hist, (AT*)q, ldq, z_sw);
    LAUNCH_CHECK();
  });
}

void k_self_history(DT act, int n, int L, int ldl, const float* z, const uint8_t* song_start, const float* carry, int carry_valid, void* q, int ldq,
                    cudaStream_t st) {
  DISPATCH_ACT(act, { self_history_kernel<AT><<<nblk((long)n * L), TPB, 0, st>>>(n, L, ldl, z, song_start, carry, carry_valid, (AT*)q, ldq); LAUNCH_CHECK(); });
}

void k_build_q(DT act, int n, int L, const float* z, const float* hist, int has_hist, void* q, int ldq, cudaStream_t st) {
  DISPATCH_ACT(act, { build_q_kernel<AT><<<nblk((long)n * L), TPB, 0, st>>>(n, L, z, hist, has_hist, (AT*)q, ldq); LAUNCH_CHECK(); });
}

}  // namespace mvae
