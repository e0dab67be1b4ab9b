"""Hand-derived forward + backward of the MIDI-VAE train step  --  TEST INFRASTRUCTURE.

This is the *blueprint the CUDA kernels follow*: the same decomposition into
(1) dense input projections  xw = X W + b  over the whole sequence,
(2) a recurrence that only adds  h U  and applies the gate math, stashing the
    post-activation gates and the cell states,
(3) a reverse-time sweep that turns  dh_ext[t]  into  dG[t]  (pre-activation gate
    gradients) and  dh[t-1] = dG[t] U^T,
(4) batched weight gradients  dW = X^T dG,  dU = Hprev^T dG,  db = colsum(dG),
with all sequences TIME-MAJOR (T,B,.) and one-hot inputs kept dense but zero-padded,
exactly as midi_vae_b200/csrc lays them out.  It uses no autograd; tests check it
against oracle/midivae_oracle.py (autograd, fp64) so a kernel bug can be told apart
from a derivation bug.  Same import restrictions as the oracle.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

from .midivae_oracle import OracleConfig

Tensor = torch.Tensor


def _act(cfg, x):
    return torch.clamp(0.2 * x + 0.5, 0, 1) if cfg.gate_act == "hard_sigmoid" else torch.sigmoid(x)


def _dact(cfg, s):
    """derivative of the gate activation expressed through its OUTPUT s."""
    if cfg.gate_act == "hard_sigmoid":
        return 0.2 * ((s > 0) & (s < 1)).to(s.dtype)
    return s * (1 - s)


def _blocks(variant):
    """column block index of (i, f, g, o) inside the 4H pre-activation."""
    return (0, 1, 2, 3) if variant == "standard" else (1, 0, 2, 3)


def lstm_seq_fwd(cfg, xw: Tensor, U: Tensor, h0: Tensor, c0: Tensor, variant="standard"):
    """xw (T,B,4H) pre-activations incl. bias.  Returns hseq (T+1,B,H) [slot 0 = h0],
    cseq (T+1,B,H) [slot 0 = c0; recalled variant stores the post-tanh cell], gates (T,B,4H) post-activation."""
    T, B, G = xw.shape
    H = G // 4
    bi, bf, bg, bo = _blocks(variant)
    hseq = xw.new_zeros(T + 1, B, H); cseq = xw.new_zeros(T + 1, B, H); gates = xw.new_zeros(T, B, G)
    hseq[0], cseq[0] = h0, c0
    for t in range(T):
        a = xw[t] + hseq[t] @ U
        i = _act(cfg, a[:, bi * H:(bi + 1) * H]); f = _act(cfg, a[:, bf * H:(bf + 1) * H])
        g = torch.tanh(a[:, bg * H:(bg + 1) * H]); o = _act(cfg, a[:, bo * H:(bo + 1) * H])
        s = f * cseq[t] + i * g
        if variant == "standard":
            cseq[t + 1] = s; hseq[t + 1] = o * torch.tanh(s)
        else:
            cseq[t + 1] = torch.tanh(s); hseq[t + 1] = o * cseq[t + 1]
        gates[t, :, bi * H:(bi + 1) * H] = i; gates[t, :, bf * H:(bf + 1) * H] = f
        gates[t, :, bg * H:(bg + 1) * H] = g; gates[t, :, bo * H:(bo + 1) * H] = o
    return hseq, cseq, gates


def lstm_seq_bwd(cfg, dh_ext: Tensor, gates: Tensor, cseq: Tensor, U: Tensor, variant="standard"):
    """dh_ext (T,B,H): gradient arriving at h_t from above (zero where none).
    Returns dG (T,B,4H) = dLoss/d(pre-activation), dh0, dc0."""
    T, B, G = gates.shape
    H = G // 4
    bi, bf, bg, bo = _blocks(variant)
    dG = torch.zeros_like(gates)
    dh = dh_ext.new_zeros(B, H); dc = dh_ext.new_zeros(B, H)
    for t in range(T - 1, -1, -1):
        dh = dh + dh_ext[t]
        i = gates[t, :, bi * H:(bi + 1) * H]; f = gates[t, :, bf * H:(bf + 1) * H]
        g = gates[t, :, bg * H:(bg + 1) * H]; o = gates[t, :, bo * H:(bo + 1) * H]
        c_prev, c_new = cseq[t], cseq[t + 1]
        if variant == "standard":
            tc = torch.tanh(c_new)
            do = dh * tc
            ds = dc + dh * o * (1 - tc * tc)
        else:
            do = dh * c_new
            ds = (dc + dh * o) * (1 - c_new * c_new)
        di, df, dg, dc = ds * g, ds * c_prev, ds * i, ds * f
        dG[t, :, bi * H:(bi + 1) * H] = di * _dact(cfg, i)
        dG[t, :, bf * H:(bf + 1) * H] = df * _dact(cfg, f)
        dG[t, :, bg * H:(bg + 1) * H] = dg * (1 - g * g)
        dG[t, :, bo * H:(bo + 1) * H] = do * _dact(cfg, o)
        dh = dG[t] @ U.T
    return dG, dh, dc


def _pad_onehot(idx: Tensor, n: int, pad: int, dtype) -> Tensor:
    """(B,T) class indices -> time-major (T+1,B,pad) with slab 0 = zeros (the decoder start vector)."""
    B, T = idx.shape
    out = torch.zeros(T + 1, B, pad, dtype=dtype)
    out[1:].scatter_(2, idx.t().long().unsqueeze(-1), 1.0)
    return out


def train_step_manual(cfg: OracleConfig, p: Dict[str, Tensor], pitch_idx, instr_idx, velocity, style_idx,
                      hist: Tensor, eps: Tensor, w_notes: Tensor = None, tgt_idx=None):
    """Forward + backward with explicit buffers.  Returns (metrics dict, grads dict keyed like p)."""
    dt = hist.dtype
    H, L, T, Ti = cfg.H, cfg.L, cfg.T, cfg.meta_instrument_length
    Dp, Di, C = cfg.input_dim, cfg.meta_instrument_dim, cfg.num_composers
    B = pitch_idx.shape[0]
    fb = cfg.decoder_feedback
    assert fb in ("as_wired", "teacher_forced")
    var = cfg.dec_cell_variant
    tgt_idx = pitch_idx if tgt_idx is None else tgt_idx
    PD = 64   # padded pitch width used on the device
    Xp = _pad_onehot(torch.as_tensor(pitch_idx), Dp, PD, dt)       # (T+1,B,64)
    Yp = _pad_onehot(torch.as_tensor(tgt_idx), Dp, PD, dt)
    Xi = _pad_onehot(torch.as_tensor(instr_idx), Di, Di, dt)       # (Ti+1,B,16)
    Xv = torch.zeros(T + 1, B, 8, dtype=dt); Xv[1:, :, 0] = torch.as_tensor(velocity, dtype=dt).t()
    w = torch.ones(B, T, dtype=dt) if w_notes is None else torch.as_tensor(w_notes, dtype=dt)
    g: Dict[str, Tensor] = {k: torch.zeros_like(v) for k, v in p.items()}

    def padW(Wm, rows):   # zero-pad kernel rows to the padded input width
        out = Wm.new_zeros(rows, Wm.shape[1]); out[:Wm.shape[0]] = Wm; return out

    # ---------------- encoder forward
    zeros = torch.zeros(B, H, dtype=dt)
    enc = []     # per recurrence: dict(name, X (T,B,D), hseq, cseq, gates)
    x_in = Xp[1:]
    for k in range(1, cfg.num_layers_encoder + 1):
        nm = f"lstm_{k}"
        Wp = padW(p[f"{nm}/kernel"], x_in.shape[2])
        xw = x_in @ Wp + p[f"{nm}/bias"]
        hs, cs, gt = lstm_seq_fwd(cfg, xw, p[f"{nm}/recurrent_kernel"], zeros, zeros)
        enc.append(dict(name=nm, X=x_in, hseq=hs, cseq=cs, gates=gt))
        x_in = hs[1:]
    side = []
    for nm, X in (("lstm_meta_instrument", Xi[1:]), ("lstm_meta_velocity", Xv[1:])):
        Wp = padW(p[f"{nm}/kernel"], X.shape[2])
        xw = X @ Wp + p[f"{nm}/bias"]
        hs, cs, gt = lstm_seq_fwd(cfg, xw, p[f"{nm}/recurrent_kernel"], zeros, zeros)
        side.append(dict(name=nm, X=X, hseq=hs, cseq=cs, gates=gt))
    u = torch.cat([enc[-1]["hseq"][-1], side[0]["hseq"][-1], side[1]["hseq"][-1]], 1)
    a1 = torch.tanh(u @ p["extra_instrument_after_concat_layer/kernel"] + p["extra_instrument_after_concat_layer/bias"])
    e = torch.tanh(a1 @ p["extra_layer/kernel"] + p["extra_layer/bias"]) if cfg.extra_layer else a1
    half = H // 2 if cfg.split_lstm_vector else H
    e1 = e[:, :half]; e2 = e[:, half:] if cfg.split_lstm_vector else e
    mu = e1 @ p["z_mean/kernel"] + p["z_mean/bias"]
    lv = e2 @ p["z_log_var/kernel"] + p["z_log_var/bias"]
    sd = torch.exp(lv / 2)
    z = mu + sd * eps
    s0sq = cfg.prior_std ** 2
    kl = cfg.beta * (-0.5 * (1 + lv - 2 * math.log(cfg.prior_std) - ((mu - cfg.prior_mean) ** 2 + torch.exp(lv)) / s0sq).sum(1)).mean()

    # ---------------- decoder forward
    q = torch.cat([z, hist], 1) if cfg.history else z
    names = [f"notes_l{k}" for k in range(1, cfg.num_layers_decoder + 1)] + ["instr", "vel"]
    S = {}
    for nm in names:
        for j in (1, 2):
            S[(nm, j)] = torch.tanh(q @ p[f"dec_init/{nm}_s{j}/kernel"] + p[f"dec_init/{nm}_s{j}/bias"])

    def dec_in(ext, steps):   # decoder input sequence (steps,B,D)
        return ext[0:steps] if fb == "teacher_forced" else torch.zeros_like(ext[0:steps])

    dec = []
    x_in = dec_in(Yp, T)
    for k in range(1, cfg.num_layers_decoder + 1):
        cn = f"notes/cell_{k}"
        Wp = padW(p[f"{cn}/kernel"], x_in.shape[2])
        xw = x_in @ Wp + p[f"{cn}/bias"]
        hs, cs, gt = lstm_seq_fwd(cfg, xw, p[f"{cn}/recurrent_kernel"], S[(f"notes_l{k}", 1)], S[(f"notes_l{k}", 2)], var)
        dec.append(dict(name=cn, init=f"notes_l{k}", X=x_in, hseq=hs, cseq=cs, gates=gt))
        x_in = hs[1:]
    for cn, init, ext, steps in (("meta_instrument/cell", "instr", Xi, Ti), ("meta_velocity/cell", "vel", Xv, T)):
        X = dec_in(ext, steps)
        Wp = padW(p[f"{cn}/kernel"], X.shape[2])
        xw = X @ Wp + p[f"{cn}/bias"]
        hs, cs, gt = lstm_seq_fwd(cfg, xw, p[f"{cn}/recurrent_kernel"], S[(init, 1)], S[(init, 2)], var)
        dec.append(dict(name=cn, init=init, X=X, hseq=hs, cseq=cs, gates=gt))
    d_notes, d_instr, d_vel = dec[cfg.num_layers_decoder - 1], dec[-2], dec[-1]
    Pn = torch.softmax(d_notes["hseq"][1:] @ p["notes/out/kernel"] + p["notes/out/bias"], -1)             # (T,B,61)
    Pi = torch.softmax(d_instr["hseq"][1:] @ p["meta_instrument/out/kernel"] + p["meta_instrument/out/bias"], -1)  # (Ti,B,16)
    Pv = torch.sigmoid(d_vel["hseq"][1:] @ p["meta_velocity/out/kernel"] + p["meta_velocity/out/bias"])   # (T,B,1)
    Pc = torch.softmax(z[:, :C], -1)

    # ---------------- losses (Keras semantics, see oracle.losses_and_metrics)
    EPS = 1e-7
    tgt_t = torch.as_tensor(tgt_idx).t().long()                     # (T,B)
    ins_t = torch.as_tensor(instr_idx).t().long()                   # (Ti,B)
    sty = torch.as_tensor(style_idx).long()
    wt = w.t()                                                      # (T,B)
    wnorm = (wt != 0).to(dt).mean()
    py = Pn.gather(2, tgt_t.unsqueeze(-1)).squeeze(-1)
    l_notes = (-(torch.log(py.clamp(EPS, 1 - EPS))) * wt).mean() / wnorm
    pi_ = Pi.gather(2, ins_t.unsqueeze(-1)).squeeze(-1)
    l_instr = -(torch.log(pi_.clamp(EPS, 1 - EPS))).mean()
    vt = torch.as_tensor(velocity, dtype=dt).t().unsqueeze(-1)
    l_vel = ((Pv - vt) ** 2).mean()
    pc = Pc.gather(1, sty.unsqueeze(-1)).squeeze(-1)
    l_style = -(torch.log(pc.clamp(EPS, 1 - EPS))).mean()
    total = cfg.notes_weight * l_notes + cfg.meta_instrument_weight * l_instr + cfg.meta_velocity_weight * l_vel + cfg.composer_weight * l_style + kl
    metrics = {
        "loss": float(total), "decoder_loss_1": float(l_notes), "decoder_loss_2": float(l_instr), "decoder_loss_3": float(l_vel),
        "composer_decoder_loss": float(l_style),
        "decoder_acc_1": float((Pn.argmax(-1) == tgt_t).to(dt).mean()),      # Keras 2.0.8 metrics: plain mean, not sample-weighted
        "decoder_acc_2": float((Pi.argmax(-1) == ins_t).to(dt).mean()),
        "decoder_acc_3": float((torch.round(Pv) == vt).to(dt).mean()),
        "composer_decoder_acc": float((Pc.argmax(-1) == sty).to(dt).mean()), "kl": float(kl),
    }

    # ---------------- backward: output heads
    def ce_dlogits(P, tgt, ptgt, scale):
        """d/dlogits of scale * -log(clip(p_tgt)) through softmax: scale*(p - onehot), zero where clipped."""
        live = ((ptgt > EPS) & (ptgt < 1 - EPS)).to(dt).unsqueeze(-1)
        d = P.clone()
        d.scatter_add_(-1, tgt.unsqueeze(-1), -torch.ones_like(ptgt).unsqueeze(-1))
        return d * live * (scale.unsqueeze(-1) if isinstance(scale, Tensor) else scale)

    dlog_n = ce_dlogits(Pn, tgt_t, py, cfg.notes_weight * wt / (wnorm * T * B))
    dlog_i = ce_dlogits(Pi, ins_t, pi_, cfg.meta_instrument_weight / (Ti * B))
    dlog_v = cfg.meta_velocity_weight * 2 * (Pv - vt) / (T * B) * Pv * (1 - Pv)

    def out_dense_bwd(name, hs, dlog):
        Hm = hs[1:].reshape(-1, H)
        dl = dlog.reshape(-1, dlog.shape[-1])
        g[f"{name}/kernel"] += Hm.t() @ dl
        g[f"{name}/bias"] += dl.sum(0)
        return dlog @ p[f"{name}/kernel"].t()           # (steps,B,H)

    dh_notes = out_dense_bwd("notes/out", d_notes["hseq"], dlog_n)
    dh_instr = out_dense_bwd("meta_instrument/out", d_instr["hseq"], dlog_i)
    dh_vel = out_dense_bwd("meta_velocity/out", d_vel["hseq"], dlog_v)

    # ---------------- backward: decoder recurrences
    dS = {}

    def rec_bwd(r, dh_ext, variant, bias_name, need_dx):
        nm = r["name"]
        dG, dh0, dc0 = lstm_seq_bwd(cfg, dh_ext, r["gates"], r["cseq"], p[f"{nm}/recurrent_kernel"], variant)
        dGm = dG.reshape(-1, 4 * H)
        g[f"{nm}/recurrent_kernel"] += r["hseq"][:-1].reshape(-1, H).t() @ dGm
        Xm = r["X"].reshape(-1, r["X"].shape[2])
        D = p[f"{nm}/kernel"].shape[0]
        g[f"{nm}/kernel"] += (Xm.t() @ dGm)[:D]
        g[bias_name] += dGm.sum(0)
        dx = dG @ p[f"{nm}/kernel"].t() if need_dx else None
        return dx, dh0, dc0

    dh_ext = dh_notes
    for k in range(cfg.num_layers_decoder, 0, -1):
        r = dec[k - 1]
        dx, dh0, dc0 = rec_bwd(r, dh_ext, var, f"{r['name']}/bias", k > 1)
        dS[(r["init"], 1)], dS[(r["init"], 2)] = dh0, dc0
        dh_ext = dx
    for r, dh_e in ((d_instr, dh_instr), (d_vel, dh_vel)):
        _, dh0, dc0 = rec_bwd(r, dh_e, var, f"{r['name']}/bias", False)
        dS[(r["init"], 1)], dS[(r["init"], 2)] = dh0, dc0

    # ---------------- backward: init-state Denses -> dz
    dq = torch.zeros_like(q)
    for nm in names:
        for j in (1, 2):
            dpre = dS[(nm, j)] * (1 - S[(nm, j)] ** 2)
            g[f"dec_init/{nm}_s{j}/kernel"] += q.t() @ dpre
            g[f"dec_init/{nm}_s{j}/bias"] += dpre.sum(0)
            dq += dpre @ p[f"dec_init/{nm}_s{j}/kernel"].t()
    dz = dq[:, :L].clone()
    # style head
    live = ((pc > EPS) & (pc < 1 - EPS)).to(dt).unsqueeze(-1)
    dsty = Pc.clone(); dsty.scatter_add_(1, sty.unsqueeze(-1), -torch.ones(B, 1, dtype=dt))
    dz[:, :C] += cfg.composer_weight / B * dsty * live
    # reparam + KL
    dmu = dz + cfg.beta * (mu - cfg.prior_mean) / s0sq / B
    dlv = dz * eps * 0.5 * sd + cfg.beta * (-0.5) * (1 - torch.exp(lv) / s0sq) / B
    # head
    g["z_mean/kernel"] += e1.t() @ dmu; g["z_mean/bias"] += dmu.sum(0)
    g["z_log_var/kernel"] += e2.t() @ dlv; g["z_log_var/bias"] += dlv.sum(0)
    if cfg.split_lstm_vector:
        de = torch.cat([dmu @ p["z_mean/kernel"].t(), dlv @ p["z_log_var/kernel"].t()], 1)
    else:
        de = dmu @ p["z_mean/kernel"].t() + dlv @ p["z_log_var/kernel"].t()
    if cfg.extra_layer:
        dpre = de * (1 - e * e)
        g["extra_layer/kernel"] += a1.t() @ dpre; g["extra_layer/bias"] += dpre.sum(0)
        da1 = dpre @ p["extra_layer/kernel"].t()
    else:
        da1 = de
    dpre = da1 * (1 - a1 * a1)
    g["extra_instrument_after_concat_layer/kernel"] += u.t() @ dpre
    g["extra_instrument_after_concat_layer/bias"] += dpre.sum(0)
    du = dpre @ p["extra_instrument_after_concat_layer/kernel"].t()

    # ---------------- backward: encoder recurrences
    def last_only(dlast, steps):
        d = torch.zeros(steps, B, H, dtype=dt); d[-1] = dlast; return d

    dh_ext = last_only(du[:, :H], T)
    for k in range(cfg.num_layers_encoder, 0, -1):
        r = enc[k - 1]
        dx, _, _ = rec_bwd(r, dh_ext, "standard", f"{r['name']}/bias", k > 1)
        dh_ext = dx
    rec_bwd(side[0], last_only(du[:, H:2 * H], Ti), "standard", "lstm_meta_instrument/bias", False)
    rec_bwd(side[1], last_only(du[:, 2 * H:], T), "standard", "lstm_meta_velocity/bias", False)
    return metrics, g


# --------------------------------------------------------------------------------------
# GRU primitives (groundwork for SURVEY.md 8(f-1); no CUDA counterpart yet).  One step is TWO dependent products:
#   a_zr = xw[:, :2H] + h U_zr ;  z, r = act(a_zr)          a_h = xw[:, 2H:] + (r*h) U_h ;  hh = tanh(a_h)
#   mix "keep" (keras.layers.GRU 2.0.8):  h' = z*h + (1-z)*hh        mix "new" (recurrentshop GRUCell as recalled):  h' = (1-z)*h + z*hh
# The stash a reverse sweep needs is (z, r, hh) per step plus hseq; dG = [da_z | da_r | da_h] plays the role the 4H gate gradient plays for
# the LSTM: dW = X^T dG, db = colsum dG, dU_zr = Hprev^T dG[:, :2H], dU_h = (r*Hprev)^T dG[:, 2H:].
# --------------------------------------------------------------------------------------
def gru_seq_fwd(cfg, xw: Tensor, U_zr: Tensor, U_h: Tensor, h0: Tensor, mix="keep"):
    """xw (T,B,3H) pre-activations incl. bias, blocks [z|r|h].  Returns hseq (T+1,B,H) [slot 0 = h0] and gates (T,B,3H) = post-activation (z, r, hh)."""
    T, B, G = xw.shape
    H = G // 3
    hseq = xw.new_zeros(T + 1, B, H); gates = xw.new_zeros(T, B, G)
    hseq[0] = h0
    for t in range(T):
        h = hseq[t]
        a = xw[t, :, :2 * H] + h @ U_zr
        z = _act(cfg, a[:, :H]); r = _act(cfg, a[:, H:])
        hh = torch.tanh(xw[t, :, 2 * H:] + (r * h) @ U_h)
        hseq[t + 1] = z * h + (1 - z) * hh if mix == "keep" else (1 - z) * h + z * hh
        gates[t, :, :H] = z; gates[t, :, H:2 * H] = r; gates[t, :, 2 * H:] = hh
    return hseq, gates


def gru_seq_bwd(cfg, dh_ext: Tensor, gates: Tensor, hseq: Tensor, U_zr: Tensor, U_h: Tensor, mix="keep"):
    """dh_ext (T,B,H): gradient arriving at h_t from outside the recurrence (slot t = h after step t).
    Returns dG (T,B,3H) pre-activation gradients [da_z|da_r|da_h] and dh0 (B,H)."""
    T, B, G = gates.shape
    H = G // 3
    dG = gates.new_zeros(T, B, G)
    carry = gates.new_zeros(B, H)
    for t in range(T - 1, -1, -1):
        z, r, hh = gates[t, :, :H], gates[t, :, H:2 * H], gates[t, :, 2 * H:]
        h = hseq[t]
        d = dh_ext[t] + carry
        if mix == "keep":
            dz = d * (h - hh); dhh = d * (1 - z); carry = d * z
        else:
            dz = d * (hh - h); dhh = d * z; carry = d * (1 - z)
        da_h = dhh * (1 - hh * hh)
        drh = da_h @ U_h.t()                       # gradient wrt (r*h)
        da_r = drh * h * _dact(cfg, r)
        da_z = dz * _dact(cfg, z)
        carry = carry + drh * r + torch.cat([da_z, da_r], dim=1) @ U_zr.t()
        dG[t, :, :H] = da_z; dG[t, :, H:2 * H] = da_r; dG[t, :, 2 * H:] = da_h
    return dG, carry


def gru_weight_grads(X: Tensor, hseq: Tensor, gates: Tensor, dG: Tensor):
    """Batched weight gradients of one GRU recurrence from the reverse sweep's dG: X (T,B,D) inputs.  -> dW (D,3H), db (3H), dU_zr (H,2H), dU_h (H,H)."""
    T, B, G = dG.shape
    H = G // 3
    Xf, Gf = X.reshape(T * B, -1), dG.reshape(T * B, G)
    Hp = hseq[:T].reshape(T * B, H)
    R = gates[:, :, H:2 * H].reshape(T * B, H)
    return Xf.t() @ Gf, Gf.sum(0), Hp.t() @ Gf[:, :2 * H], (R * Hp).t() @ Gf[:, 2 * H:]
