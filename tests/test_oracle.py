"""CPU tests of the oracle itself: finite-difference gradient checks (fp64), the hand-derived BPTT blueprint the
kernels follow against autograd, Keras-semantics identities, and the committed golden vectors."""
import math
import os

import numpy as np
import pytest
import torch

from midi_vae_b200 import METRIC_KEYS, synth
from oracle import manual_bptt as M
from oracle import midivae_oracle as O
from tests import util


def _setup(feedback="teacher_forced", gate="hard_sigmoid", variant="standard", T=8, H=16, L=8, n=4, ne=2, nd=2):
    ecfg, ocfg = util.make_cfgs(T=T, H=H, L=L, ne=ne, nd=nd, feedback=feedback, gate=gate, variant=variant, max_batch=n)
    w = util.make_weights(ecfg, jitter=0.05)
    p = util.to_torch(w)
    r, hist, eps, sw = util.make_batch(ecfg, n, weights=True)
    return ocfg, p, r, util.oracle_inputs(ocfg, r, hist, eps, sw)


@pytest.mark.parametrize("feedback", ["as_wired", "teacher_forced", "free_running"])
@pytest.mark.parametrize("gate", ["hard_sigmoid", "sigmoid"])
def test_autograd_matches_finite_differences(feedback, gate):
    ocfg, p, r, (X, I, V, C, hist, eps, sw) = _setup(feedback, gate)
    _, g, _ = O.loss_and_grads(ocfg, p, X, I, V, C, hist, eps, sample_weight=sw)
    rng = np.random.default_rng(0)
    for name in ["lstm_1/kernel", "lstm_2/recurrent_kernel", "lstm_meta_velocity/kernel", "z_log_var/kernel", "notes/cell_2/recurrent_kernel",
                 "meta_velocity/out/kernel", "dec_init/instr_s2/kernel", "extra_layer/bias"]:
        wt = p[name]
        for _ in range(2):
            idx = tuple(int(rng.integers(0, s)) for s in wt.shape)
            old = wt[idx].item(); h = 1e-6
            wt[idx] = old + h; lp = O.evaluate_batch(ocfg, p, X, I, V, C, hist, eps, sample_weight=sw)[0]["loss"]
            wt[idx] = old - h; lm = O.evaluate_batch(ocfg, p, X, I, V, C, hist, eps, sample_weight=sw)[0]["loss"]
            wt[idx] = old
            assert abs((lp - lm) / (2 * h) - g[name][idx].item()) < 1e-7, name


@pytest.mark.parametrize("feedback", ["as_wired", "teacher_forced"])
@pytest.mark.parametrize("gate", ["hard_sigmoid", "sigmoid"])
@pytest.mark.parametrize("variant", ["standard", "recurrentshop_recalled"])
def test_manual_bptt_matches_autograd(feedback, gate, variant):
    """The derivation the CUDA kernels transcribe (time-major buffers, padded one-hots, stashed gates)."""
    ocfg, p, r, (X, I, V, C, hist, eps, sw) = _setup(feedback, gate, variant, n=5)
    m, g, _ = O.loss_and_grads(ocfg, p, X, I, V, C, hist, eps, sample_weight=sw)
    m2, g2 = M.train_step_manual(ocfg, p, r.pitch, r.instr, r.velocity, r.style, hist, eps, w_notes=sw[0])
    for k in METRIC_KEYS:
        assert abs(m[k] - m2[k]) < 1e-12, k
    for k in g:
        assert (g[k] - g2[k]).abs().max().item() <= 1e-12 * max(1.0, g[k].abs().max().item()), k


def test_single_layer_and_three_layer_stacks():
    for ne, nd in ((1, 1), (3, 1), (1, 3)):
        ocfg, p, r, (X, I, V, C, hist, eps, sw) = _setup(ne=ne, nd=nd)
        m, g, _ = O.loss_and_grads(ocfg, p, X, I, V, C, hist, eps)
        m2, g2 = M.train_step_manual(ocfg, p, r.pitch, r.instr, r.velocity, r.style, hist, eps)
        assert abs(m["loss"] - m2["loss"]) < 1e-12
        assert max((g[k] - g2[k]).abs().max().item() for k in g) < 1e-12


def test_keras_loss_identities():
    ocfg, p, r, (X, I, V, C, hist, eps, sw) = _setup()
    m, outs, (z, mu, lv) = O.evaluate_batch(ocfg, p, X, I, V, C, hist, eps)
    # total = sum w_i * loss_i + KL, and the script's KL recovery (vae_training.py:946-957)
    s = 1.0 * m["decoder_loss_1"] + 0.1 * m["decoder_loss_2"] + 1.0 * m["decoder_loss_3"] + 0.1 * m["composer_decoder_loss"]
    assert abs(m["loss"] - s - m["kl"]) < 1e-12
    # KL closed form for a standard-normal prior
    kl = 0.1 * (-0.5 * (1 + lv - mu ** 2 - lv.exp()).sum(1)).mean()
    assert abs(float(kl) - m["kl"]) < 1e-12
    # softmax outputs are normalised; CE of a one-hot target is -log p_target
    Yh = outs[0]
    assert torch.allclose(Yh.sum(-1), torch.ones_like(Yh.sum(-1)))
    ce = -(X * Yh.clamp(1e-7, 1 - 1e-7).log()).sum(-1).mean()
    assert abs(float(ce) - m["decoder_loss_1"]) < 1e-9
    # temporal sample weights: zero weights drop out of numerator AND normaliser
    w = torch.ones(X.shape[0], X.shape[1], dtype=X.dtype); w[:, ::2] = 0
    mw, _, _ = O.evaluate_batch(ocfg, p, X, I, V, C, hist, eps, sample_weight=(w, None, None, None))
    ce_kept = -(X * Yh.clamp(1e-7, 1 - 1e-7).log()).sum(-1)[:, 1::2].mean()
    assert abs(float(ce_kept) - mw["decoder_loss_1"]) < 1e-9


def test_keras_adam_first_steps():
    p = {"w": torch.tensor([1.0, -2.0, 3.0], dtype=torch.float64)}
    opt = O.KerasAdam(p, lr=1e-3)
    g = {"w": torch.tensor([0.5, -0.25, 0.0], dtype=torch.float64)}
    opt.step(p, g)
    lr_t = 1e-3 * math.sqrt(1 - 0.999) / (1 - 0.9)
    exp = torch.tensor([1.0, -2.0, 3.0], dtype=torch.float64) - lr_t * (0.1 * g["w"]) / ((0.001 * g["w"] ** 2).sqrt() + 1e-8)
    assert torch.allclose(p["w"], exp, atol=1e-15)


def test_param_inventory_matches_survey():
    """14.144 M parameters at cfg3, 3.510 M at cfg2, 0.245 M at cfg1."""
    for (T, H, L), n in (((256, 512, 256), 14.144e6), ((64, 256, 100), 3.510e6), ((16, 64, 16), 0.245e6)):
        c = O.OracleConfig(input_length=T, lstm_size=H, latent_rep_size=L)
        assert abs(O.param_count(c) - n) / n < 2e-3, (O.param_count(c), n)


def test_history_shift_and_style_swap():
    ocfg, p, r, (X, I, V, C, hist, eps, sw) = _setup(n=6)
    ss = np.array([1, 0, 0, 1, 0, 0], bool)
    out = O.style_transfer(ocfg, p, X, I, V, 0, 1, ss)
    assert torch.equal(out["z_sw"][:, 0], out["z"][:, 1]) and torch.equal(out["z_sw"][:, 1], out["z"][:, 0])
    assert torch.equal(out["z_sw"][:, 2:], out["z"][:, 2:])
    assert float(out["H_sw"][0].abs().sum()) == 0 and float(out["H_sw"][3].abs().sum()) == 0
    assert torch.equal(out["H_sw"][1], out["z_sw"][0]) and torch.equal(out["H_sw"][5], out["z_sw"][4])


def test_golden_vectors_reproduce():
    """tests/golden/cfg1_step.npz is what tests/golden/make_golden.py writes (oracle-derived, fp64)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cfg1_step.npz"))
    ecfg, ocfg = util.make_cfgs(T=16, H=64, L=16, feedback=str(g["feedback"]), variant=str(g["variant"]), max_batch=8)
    p = {k[2:]: torch.tensor(g[k], dtype=torch.float64) for k in g.files if k.startswith("w/")}
    r = synth.Rolls(g["pitch"], g["instr"], g["velocity"], g["style"])
    X, I, V, C, th, te, _ = util.oracle_inputs(ocfg, r, g["hist"], g["eps"], None)
    m, gr, _ = O.loss_and_grads(ocfg, p, X, I, V, C, th, te)
    assert np.allclose([m[k] for k in METRIC_KEYS], g["metrics"], rtol=0, atol=1e-12)
    for k, v in gr.items():
        assert np.abs(v.numpy() - g["g/" + k]).max() <= 1e-6 * max(1.0, float(v.abs().max())), k


@pytest.mark.parametrize("mix", ["keep", "new"])
@pytest.mark.parametrize("gate", ["hard_sigmoid", "sigmoid"])
def test_manual_gru_bptt_matches_autograd(mix, gate):
    """The hand-derived GRU reverse sweep (two dependent products per step; stash = z, r, hh) against autograd, for the Keras GRU mix and for
    the recalled recurrentshop GRUCell mix: blueprint of the GRU kernels of SURVEY.md 8(f-1)."""
    from oracle import manual_bptt as MB
    torch.manual_seed(3)
    T_, B_, H_, D_ = 7, 5, 12, 9
    cfg = O.OracleConfig(input_length=T_, lstm_size=H_, gate_act=gate)
    dt = torch.float64
    X = torch.randn(T_, B_, D_, dtype=dt)
    leaves = [torch.randn(D_, 3 * H_, dtype=dt) * 0.4, torch.randn(3 * H_, dtype=dt) * 0.2, torch.randn(H_, 2 * H_, dtype=dt) * 0.4,
              torch.randn(H_, H_, dtype=dt) * 0.4, torch.randn(B_, H_, dtype=dt) * 0.5]
    W, b, Uzr, Uh, h0 = [l.requires_grad_(True) for l in leaves]
    probe = torch.randn(T_, B_, H_, dtype=dt)
    act = O.hard_sigmoid if gate == "hard_sigmoid" else torch.sigmoid
    h, hs = h0, []
    xw = X @ W + b
    for t in range(T_):                     # autograd-friendly restatement of the same step (no in-place writes)
        a = xw[t, :, :2 * H_] + h @ Uzr
        z, r = act(a[:, :H_]), act(a[:, H_:])
        hh = torch.tanh(xw[t, :, 2 * H_:] + (r * h) @ Uh)
        h = z * h + (1 - z) * hh if mix == "keep" else (1 - z) * h + z * hh
        hs.append(h)
    loss = (torch.stack(hs) * probe).sum()
    gW, gb, gUzr, gUh, gh0 = torch.autograd.grad(loss, [W, b, Uzr, Uh, h0])
    with torch.no_grad():
        hseq, gates = MB.gru_seq_fwd(cfg, xw.detach(), Uzr.detach(), Uh.detach(), h0.detach(), mix)
        assert (hseq[1:] - torch.stack(hs)).abs().max() < 1e-12
    with torch.no_grad():
        dG, dh0 = MB.gru_seq_bwd(cfg, probe, gates, hseq, Uzr.detach(), Uh.detach(), mix)
        dW, db, dUzr, dUh = MB.gru_weight_grads(X, hseq, gates, dG)
    for mine, ref in ((dW, gW), (db, gb), (dUzr, gUzr), (dUh, gUh), (dh0, gh0)):
        assert (mine - ref).abs().max() < 1e-10 * max(1.0, float(ref.abs().max()))
    # and the forward is the oracle's: Keras GRU layer (keep) / recalled GRUCell (new)
    if mix == "keep":
        p = {"g/kernel": W.detach(), "g/recurrent_kernel": torch.cat([Uzr, Uh], 1).detach(), "g/bias": b.detach()}
        ref_h = O.keras_gru(cfg, p, "g", X.permute(1, 0, 2), True)
        # keras_gru starts from h = 0: compare a run from h0 = 0
        hs0, _ = MB.gru_seq_fwd(cfg, (X @ W + b).detach(), Uzr.detach(), Uh.detach(), torch.zeros(B_, H_, dtype=dt), mix)
        assert (hs0[1:].permute(1, 0, 2) - ref_h).abs().max() < 1e-12


# ------------------------------------------------------------------------------------------------ independent witnesses (PyTorch's own layers)
def test_lstm_equations_agree_with_torch_nn_lstm():
    """An implementation nobody here wrote: torch.nn.LSTM (gate blocks i, f, g, o; logistic gates) against the oracle's Keras-LSTM restatement
    with gate_act='sigmoid' (Keras blocks [i|f|c|o] are the same order).  Pins the cell equations, the block order and the bias handling;
    the Keras default hard_sigmoid differs from this only in the gate non-linearity (oracle.hard_sigmoid = clip(0.2 x + 0.5, 0, 1))."""
    torch.manual_seed(0)
    T_, B_, D_, H_ = 9, 4, 7, 10
    cfg = O.OracleConfig(input_length=T_, lstm_size=H_, gate_act="sigmoid")
    ref = torch.nn.LSTM(D_, H_, batch_first=True).double()
    p = {"l/kernel": ref.weight_ih_l0.detach().t().contiguous(), "l/recurrent_kernel": ref.weight_hh_l0.detach().t().contiguous(),
         "l/bias": (ref.bias_ih_l0 + ref.bias_hh_l0).detach()}
    x = torch.randn(B_, T_, D_, dtype=torch.float64)
    with torch.no_grad():
        want, (h_n, _) = ref(x)
        got = O.keras_lstm(cfg, p, "l", x, True)
        last = O.keras_lstm(cfg, p, "l", x, False)
    assert (got - want).abs().max() < 1e-12 and (last - h_n[0]).abs().max() < 1e-12


def test_keras_adam_agrees_with_torch_adam_up_to_the_epsilon_placement():
    """keras.optimizers.Adam (2.0.8) and torch.optim.Adam are the same recursion; they differ only in where epsilon enters
    (Keras: lr_t m / (sqrt(v) + eps) with the bias corrections folded into lr_t; torch: eps is added to sqrt(v_hat)).  With eps -> 0 the
    trajectories must coincide."""
    torch.manual_seed(1)
    w0 = torch.randn(6, 5, dtype=torch.float64)
    p = {"w": w0.clone()}
    opt = O.KerasAdam(p, lr=2e-4, epsilon=1e-300)
    wt = w0.clone().requires_grad_(True)
    topt = torch.optim.Adam([wt], lr=2e-4, betas=(0.9, 0.999), eps=1e-300)
    for step in range(25):
        g = torch.sin(torch.arange(30, dtype=torch.float64).reshape(6, 5) * (step + 1)) + 0.3
        opt.step(p, {"w": g})
        wt.grad = g.clone()
        topt.step()
    assert (p["w"] - wt.detach()).abs().max() < 1e-12


def test_crossentropy_agrees_with_torch_cross_entropy():
    """Theano-backend categorical_crossentropy on softmax outputs (renormalise, clip to [1e-7, 1 - 1e-7], -sum y log p) equals
    torch.nn.functional.cross_entropy on the logits wherever the clip is inactive."""
    torch.manual_seed(2)
    logits = torch.randn(5, 8, 61, dtype=torch.float64) * 2
    idx = torch.randint(0, 61, (5, 8))
    y = torch.nn.functional.one_hot(idx, 61).double()
    got = O.categorical_crossentropy(y, torch.softmax(logits, -1))
    want = torch.nn.functional.cross_entropy(logits.reshape(-1, 61), idx.reshape(-1), reduction="none").reshape(5, 8)
    assert (got - want).abs().max() < 1e-10
