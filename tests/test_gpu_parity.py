"""GPU parity tests: the CUDA path (through the C ABI of libmidivae.so) against the CPU oracle on the same
seeded inputs.  Tolerances: fp32 precision 1e-4 relative (north_star), bf16 precision stated per test.
Integer outputs (argmax pitch / instrument indices) are bit-exact wherever the oracle's top-2 margin exceeds the
stated tolerance."""
import numpy as np
import pytest
import torch

from midi_vae_b200 import Engine, METRIC_KEYS, synth
from oracle import midivae_oracle as O
from tests import util

pytestmark = pytest.mark.gpu

TOL32 = 1e-4


def _engine(ecfg, w):
    eng = Engine(ecfg, 0)
    eng.set_weights(w)
    return eng


def _compare_step(ecfg, ocfg, n, weights=False, tol=TOL32, grad_tol=None, seed=7):
    w = util.make_weights(ecfg)
    eng = _engine(ecfg, w)
    r, hist, eps, sw = util.make_batch(ecfg, n, seed=seed, weights=weights)
    p = util.to_torch(w)
    X, I, V, C, th, te, tsw = util.oracle_inputs(ocfg, r, hist, eps, sw)
    m_ref, g_ref, _ = O.loss_and_grads(ocfg, p, X, I, V, C, th, te, sample_weight=tsw)
    m = eng.train_on_batch(r.pitch, r.instr, r.velocity, r.style, hist, eps, sw)
    for k in METRIC_KEYS:
        assert abs(m[k] - m_ref[k]) <= tol * max(1.0, abs(m_ref[k])), (k, m[k], m_ref[k])
    g = eng.get_grads()
    gt = grad_tol or tol
    worst = max(g_ref, key=lambda k: util.rel_err(g[k], g_ref[k].numpy()))
    for k in g_ref:
        ref = g_ref[k].numpy()
        scale = max(np.abs(ref).max(), 1e-6)
        assert np.abs(g[k] - ref).max() <= gt * scale + 1e-9, (k, float(np.abs(g[k] - ref).max()), float(scale), "worst", worst)
    # one Keras-Adam step.  The optimiser is checked on the gradients the device itself produced (fp32 arena), so that the tolerance can
    # be tight on EVERY element -- including the tiny-gradient ones where Keras' epsilon placement (outside the bias-corrected sqrt,
    # i.e. an effective eps / sqrt(1 - beta_2) on the first step) differs visibly from the textbook form -- independent of the precision mode
    import torch as _t
    opt = O.KerasAdam(p, lr=ocfg.learning_rate)
    opt.step(p, {k: _t.tensor(np.asarray(g[k], np.float64)) for k in g_ref})
    w_new = eng.get_weights()
    for k in p:
        d_ref = p[k].numpy() - w[k]
        d = w_new[k] - w[k]
        ulp = np.spacing(np.abs(w[k]).astype(np.float32)).astype(np.float64)
        assert (np.abs(d - d_ref) <= 2e-3 * ocfg.learning_rate + 2 * ulp).all(), (k, float(np.abs(d - d_ref).max()))
    eng.close()
    return m, m_ref


@pytest.mark.parametrize("feedback", ["as_wired", "teacher_forced"])
@pytest.mark.parametrize("gate", ["hard_sigmoid", "sigmoid"])
def test_train_step_fp32_matches_oracle(feedback, gate):
    ecfg, ocfg = util.make_cfgs(T=16, H=64, L=16, feedback=feedback, gate=gate, max_batch=8)   # cfg1 (BASELINE.json configs[0])
    _compare_step(ecfg, ocfg, 8, weights=True)


def test_train_step_fp32_recalled_cell_and_layers():
    ecfg, ocfg = util.make_cfgs(T=8, H=32, L=12, ne=1, nd=3, feedback="teacher_forced", variant="recurrentshop_recalled", max_batch=6)
    _compare_step(ecfg, ocfg, 5)


@pytest.mark.parametrize("n", [1, 3, 8])
def test_ragged_batch_sizes(n):
    ecfg, ocfg = util.make_cfgs(T=16, H=64, L=16, feedback="teacher_forced", max_batch=8)
    _compare_step(ecfg, ocfg, n)


def test_trajectory_fp32_ten_steps():
    """10 consecutive train steps (history fed from the previous batch's z as in SURVEY 8(d)) track the oracle."""
    ecfg, ocfg = util.make_cfgs(T=16, H=64, L=16, feedback="teacher_forced", max_batch=8, lr=2e-3)
    w = util.make_weights(ecfg)
    eng = _engine(ecfg, w)
    p = util.to_torch(w)
    opt = O.KerasAdam(p, lr=ocfg.learning_rate)
    for step in range(10):
        r, hist, eps, _ = util.make_batch(ecfg, 8, seed=100 + step)
        X, I, V, C, th, te, _ = util.oracle_inputs(ocfg, r, hist, eps, None)
        m_ref, _ = O.train_on_batch(ocfg, p, opt, X, I, V, C, th, te)
        m = eng.train_on_batch(r.pitch, r.instr, r.velocity, r.style, hist, eps)
        assert abs(m["loss"] - m_ref["loss"]) <= 5e-4 * abs(m_ref["loss"]), (step, m["loss"], m_ref["loss"])
    w_new = eng.get_weights()
    for k in p:
        assert np.abs(w_new[k] - p[k].numpy()).max() <= 2e-4, k
    assert eng.iterations == 10
    eng.close()


def test_evaluate_and_predict_fp32():
    ecfg, ocfg = util.make_cfgs(T=16, H=64, L=16, feedback="teacher_forced", max_batch=8)
    w = util.make_weights(ecfg)
    eng = _engine(ecfg, w)
    p = util.to_torch(w)
    r, hist, eps, sw = util.make_batch(ecfg, 8, weights=True)
    X, I, V, C, th, te, tsw = util.oracle_inputs(ocfg, r, hist, eps, sw)
    m_ref, outs, aux = O.evaluate_batch(ocfg, p, X, I, V, C, th, te, sample_weight=tsw)
    m = eng.evaluate_batch(r.pitch, r.instr, r.velocity, r.style, hist, eps, sw)
    for k in METRIC_KEYS:
        assert abs(m[k] - m_ref[k]) <= TOL32 * max(1.0, abs(m_ref[k])), k
    z, mu, lv = eng.encode(r.pitch, r.instr, r.velocity, eps)
    assert util.rel_err(z, aux[0].numpy()) <= TOL32 and util.rel_err(mu, aux[1].numpy()) <= TOL32 and util.rel_err(lv, aux[2].numpy()) <= TOL32
    Y, Ih, Vh, S, z2 = eng.autoencode(r.pitch, r.instr, r.velocity, hist, eps)
    assert util.rel_err(Y, outs[0].numpy()) <= TOL32
    assert util.rel_err(Ih, outs[1].numpy()) <= TOL32
    assert util.rel_err(Vh, outs[2].numpy()[..., 0]) <= TOL32
    assert util.rel_err(S, outs[3].numpy()) <= TOL32
    # decoder.predict on a given z, all three feedback modes
    for fb in ("as_wired", "teacher_forced", "free_running"):
        Yr, Ir, Vr = O.decode(ocfg, p, aux[0], th, X, I, V, fb)
        Yd, Id, Vd = eng.decode(z, hist, fb, r.pitch, r.instr, r.velocity)
        assert util.rel_err(Yd, Yr.numpy()) <= 2 * TOL32, fb
        assert util.rel_err(Id, Ir.numpy()) <= 2 * TOL32, fb
        assert util.rel_err(Vd, Vr.numpy()[..., 0]) <= 2 * TOL32, fb
    eng.close()


@pytest.mark.parametrize("feedback", ["as_wired", "free_running"])
def test_style_transfer_argmax_exact(feedback):
    ecfg, ocfg = util.make_cfgs(T=16, H=64, L=16, max_batch=40)
    w = util.make_weights(ecfg, jitter=0.2)
    eng = _engine(ecfg, w)
    p = util.to_torch(w)
    songs = synth.make_songs(3, 16, seed=5, min_chunks=8, max_chunks=14)
    r = synth.concat(songs)
    X, I, V, C = [torch.tensor(a) for a in r.dense(np.float64)]
    ref = O.style_transfer(ocfg, p, X, I, V, 0, 1, r.song_start, feedback)
    P, Ii, Vv = eng.style_transfer(r.pitch, r.instr, r.velocity, 0, 1, r.song_start, feedback)
    tol = 1e-4
    safe_p = O.top2_margin(ref["Yh"]).numpy() > tol
    safe_i = O.top2_margin(ref["Ih"]).numpy() > tol
    assert safe_p.mean() > 0.9
    assert np.array_equal(P[safe_p], ref["pitch"].numpy()[safe_p])
    assert np.array_equal(Ii[safe_i], ref["instr"].numpy()[safe_i])
    assert util.rel_err(Vv, ref["Vh"].numpy()[..., 0]) <= 2e-4
    eng.close()


def test_golden_fixture_cfg1():
    """Committed oracle-derived golden vectors (tests/golden/make_golden.py) for BASELINE config[0] shapes."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cfg1_step.npz"))
    ecfg, _ = util.make_cfgs(T=16, H=64, L=16, feedback=str(g["feedback"]), variant=str(g["variant"]), max_batch=8)
    w = {k[2:]: g[k] for k in g.files if k.startswith("w/")}
    eng = _engine(ecfg, w)
    P, Ii, Vv = eng.style_transfer(g["pitch"], g["instr"], g["velocity"], 0, 1, None, "as_wired")   # before the weights move
    m = eng.train_on_batch(g["pitch"], g["instr"], g["velocity"], g["style"], g["hist"], g["eps"])
    for i, k in enumerate(METRIC_KEYS):
        assert abs(m[k] - g["metrics"][i]) <= TOL32 * max(1.0, abs(g["metrics"][i])), k
    gr = eng.get_grads()
    for k in gr:
        ref = g["g/" + k]
        assert np.abs(gr[k] - ref).max() <= TOL32 * max(np.abs(ref).max(), 1e-6) + 1e-9, k
    safe = g["st_margin"] > 1e-4
    assert np.array_equal(P[safe], g["st_pitch"][safe])
    eng.close()


@pytest.mark.parametrize("variant", ["standard", "recurrentshop_recalled"])
def test_reference_executed_fixture_cfg1(variant):
    """The CUDA path, through the reference-shaped facade and the reference's own positional lists, against vectors produced by
    EXECUTING the reference's vae_definition.py (tests/golden/make_reference_golden.py; Keras / recurrentshop restated in
    oracle/keras_shim): encoder / decoder / autoencoder predict, evaluate, three fit steps and the weights they leave behind.
    fp32 precision, 1e-4 relative (north_star's tolerance)."""
    import os
    from midi_vae_b200 import VAE, marshal
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_cfg1.npz"))
    T, H, L, n = 16, 64, 16, 8
    vae = VAE().create(input_dim=61, output_dim=61, input_length=T, output_length=T, latent_rep_size=L, lstm_size=H, activation='softmax',
                       include_composer_decoder=True, num_composers=2, composer_weight=0.1, num_layers_encoder=2, num_layers_decoder=2,
                       learning_rate=2e-4, beta=0.1, extra_layer=True, meta_instrument=True, meta_instrument_dim=16, meta_instrument_length=4,
                       meta_instrument_activation='softmax', meta_instrument_weight=0.1, meta_velocity=True, meta_velocity_length=T,
                       meta_velocity_weight=1.0, epsilon_std=0.01, teacher_force=False, max_batch=n, dec_cell_variant=variant,
                       decoder_feedback="as_wired", precision="fp32")
    ecfg, _ = util.make_cfgs(T=T, H=H, L=L, feedback="as_wired", variant=variant, max_batch=n)
    vae.engine.set_weights(util.make_weights(ecfg, seed=42, jitter=0.1))
    pre = variant + "/"
    lists = {tag: [g[f"lists/{tag}_{i}"] for i in range(sum(1 for k in g.files if k.startswith(f"lists/{tag}_")))]
             for tag in ("in", "out", "sw", "enc_in")}
    eps = g["eps"]
    z = vae.encoder.predict(lists["enc_in"], batch_size=n, eps=eps)
    assert util.rel_err(z, g[pre + "z"]) <= TOL32
    hist = lists["in"][2]
    Y, Ih, Vh = vae.decoder.predict(marshal.prepare_decoder_input(g[pre + "z"], H=hist), batch_size=n)
    for mine, name in ((Y, "Y"), (Ih, "I"), (Vh, "V")):
        assert util.rel_err(mine, g[pre + "dec_" + name]) <= 2 * TOL32, name
    outs = vae.autoencoder.predict(lists["in"], batch_size=n, eps=eps)
    for mine, name in zip(outs, ("Y", "I", "V", "C")):
        assert util.rel_err(mine, g[pre + "ae_" + name]) <= 2 * TOL32, name
    ev = vae.autoencoder.evaluate(lists["in"], lists["out"], batch_size=n, sample_weight=lists["sw"], eps=eps)
    for k, a, b in zip(vae.autoencoder.metrics_names, ev, g[pre + "evaluate"]):
        assert abs(a - b) <= TOL32 * max(1.0, abs(b)), (k, a, b)
    keys = list(g[pre + "fit_keys"])
    for step in range(3):
        h = vae.autoencoder.fit(lists["in"], lists["out"], epochs=1, batch_size=n, shuffle=False, sample_weight=lists["sw"], eps=eps)
        for k, b in zip(keys, g[pre + "fit"][step]):
            a = h.history[k][0]
            assert abs(a - b) <= TOL32 * max(1.0, abs(b)), (step, k, a, b)
    w0 = util.make_weights(ecfg, seed=42, jitter=0.1)
    w3 = vae.engine.get_weights()
    for k in w3:      # three Adam steps of ~lr each: compare the UPDATE
        d_ref = g[pre + "w3/" + k] - w0[k]
        d = w3[k] - w0[k]
        assert np.abs(d - d_ref).max() <= 0.1 * 3 * 2e-4 + 1e-7, k
    vae.engine.close()


def test_gemm_tc_selftest():
    """tcgen05 GEMM (all four major-ness cases, ragged edges, all epilogues) against the SIMT GEMM."""
    from midi_vae_b200 import _lib
    assert _lib.load().mvae_selftest_gemm(0, 0) == 0


@pytest.mark.parametrize("feedback", ["as_wired", "teacher_forced"])
def test_train_step_bf16_tolerance(feedback):
    """bf16 tensor-core precision: operands rounded to bf16, fp32 accumulation and fp32 cell state.
    Stated tolerance: metrics 2e-2 relative, gradients 6e-2 of each tensor's max (T=64 recurrent steps)."""
    ecfg, ocfg = util.make_cfgs(T=64, H=256, L=100, feedback=feedback, precision="bf16", max_batch=16)   # cfg2 shapes, small batch
    _compare_step(ecfg, ocfg, 16, tol=2e-2, grad_tol=6e-2)


@pytest.mark.parametrize("shape", [(16, 64, 16, 8), (64, 256, 100, 16), (12, 128, 32, 200), (8, 1024, 64, 40), (4, 192, 24, 130), (12, 512, 48, 70), (8, 256, 40, 150), (4, 512, 24, 65), (4, 256, 24, 1)])
@pytest.mark.parametrize("feedback,variant", [("as_wired", "standard"), ("teacher_forced", "standard"), ("teacher_forced", "recurrentshop_recalled")])
def test_persistent_rnn_matches_streamed(shape, feedback, variant):
    """The persistent-RNN kernels (U resident in SMEM, in-kernel time loop, cross-CTA flags) against the step-streamed
    form (one tcgen05 GEMM + one pointwise launch per step): same bf16 operands, so only accumulation order and the
    tanh.approx gate math differ.  Also covers a ragged 2-group batch (200 rows = 128 + 72).  H = 256 / 512 run the cluster
    kernels (lstm_cluster.cu: 8 / 16-CTA clusters, CTA-pair MMA, multicast / DSMEM exchange, in-kernel input projections), with
    ragged groups (70 = 64 + 6, 150 = 2 x 64 + 22, 65 = 64 + 1 rows), a single row and 4-step sequences."""
    T, H, L, n = shape
    res = {}
    for mode in ("streamed", "persistent"):
        ecfg, _ = util.make_cfgs(T=T, H=H, L=L, feedback=feedback, variant=variant, precision="bf16", max_batch=n, rnn_mode=mode)
        w = util.make_weights(ecfg)
        eng = _engine(ecfg, w)
        r, hist, eps, sw = util.make_batch(ecfg, n, weights=True)
        m = eng.train_on_batch(r.pitch, r.instr, r.velocity, r.style, hist, eps, sw)
        res[mode] = (m, eng.get_grads())
        eng.close()
    ms, gs = res["streamed"]; mp, gp = res["persistent"]
    for k in METRIC_KEYS:
        # accuracies are counts of argmax hits over few rows: allow a couple of near-tie flips on an untrained model
        tol_k = 0.05 if "acc" in k else 5e-3 * max(1.0, abs(ms[k]))
        assert abs(ms[k] - mp[k]) <= tol_k, (k, ms[k], mp[k])
    for k in gs:
        scale = max(np.abs(gs[k]).max(), 1e-6)
        assert np.abs(gs[k] - gp[k]).max() <= 3e-2 * scale + 1e-9, (k, float(np.abs(gs[k] - gp[k]).max()), float(scale))


def test_cluster_rnn_sigmoid_gates():
    """The logistic-sigmoid gate variant (north_star's wording; the reference default is hard_sigmoid) through the cluster kernels."""
    T, H, L, n = 16, 256, 32, 70
    res = {}
    for mode in ("streamed", "persistent"):
        ecfg, _ = util.make_cfgs(T=T, H=H, L=L, feedback="teacher_forced", gate="sigmoid", precision="bf16", max_batch=n, rnn_mode=mode)
        w = util.make_weights(ecfg)
        eng = _engine(ecfg, w)
        r, hist, eps, sw = util.make_batch(ecfg, n, weights=True)
        m = eng.train_on_batch(r.pitch, r.instr, r.velocity, r.style, hist, eps, sw)
        res[mode] = (m, eng.get_grads())
        eng.close()
    ms, gs = res["streamed"]; mp, gp = res["persistent"]
    for k in METRIC_KEYS:
        tol_k = 0.05 if "acc" in k else 5e-3 * max(1.0, abs(ms[k]))
        assert abs(ms[k] - mp[k]) <= tol_k, (k, ms[k], mp[k])
    for k in gs:
        scale = max(np.abs(gs[k]).max(), 1e-6)
        assert np.abs(gs[k] - gp[k]).max() <= 3e-2 * scale + 1e-9, (k, float(np.abs(gs[k] - gp[k]).max()), float(scale))


def test_bf16_tracks_fp32_full_cfg2():
    """BASELINE cfg2 (T64,H256,L100,B128) at full size: size-independent properties instead of the slow oracle:
    the bf16 path agrees with the fp32 CUDA path (itself oracle-checked above) and the loss decreases."""
    ecfg32, _ = util.make_cfgs(T=64, H=256, L=100, feedback="teacher_forced", precision="fp32", max_batch=128, lr=1e-3)
    ecfg16, _ = util.make_cfgs(T=64, H=256, L=100, feedback="teacher_forced", precision="bf16", max_batch=128, lr=1e-3)
    w = util.make_weights(ecfg32)
    e32, e16 = _engine(ecfg32, w), _engine(ecfg16, w)
    r, hist, eps, _ = util.make_batch(ecfg32, 128, seed=3)
    l32, l16 = [], []
    for _ in range(8):
        l32.append(e32.train_on_batch(r.pitch, r.instr, r.velocity, r.style, hist, eps)["loss"])
        l16.append(e16.train_on_batch(r.pitch, r.instr, r.velocity, r.style, hist, eps)["loss"])
    assert l32[-1] < l32[0] and l16[-1] < l16[0]
    for a, b in zip(l32, l16):
        assert abs(a - b) <= 2e-2 * abs(a), (l32, l16)
    e32.close(); e16.close()


def test_keras_facade_fit_evaluate_predict():
    """The reference-shaped surface: VAE.create(**kwargs) -> autoencoder.fit / evaluate / predict with the
    reference's positional input lists (vae_definition.py:880-1045)."""
    from midi_vae_b200 import VAE, marshal
    T, H, L = 16, 64, 16
    vae = VAE().create(input_dim=61, output_dim=61, input_length=T, output_length=T, latent_rep_size=L, lstm_size=H, activation='softmax',
                       include_composer_decoder=True, num_composers=2, composer_weight=0.1, num_layers_encoder=2, num_layers_decoder=2,
                       learning_rate=2e-4, beta=0.1, extra_layer=True, meta_instrument=True, meta_instrument_dim=16, meta_instrument_length=4,
                       meta_instrument_activation='softmax', meta_instrument_weight=0.1, meta_velocity=True, meta_velocity_length=T,
                       meta_velocity_weight=1.0, epsilon_std=0.01, max_batch=8)
    song = synth.make_song(np.random.default_rng(0), 20, T, style=1)
    X, I, V, C = song.dense(np.float64)
    H0 = np.zeros((len(song), L))
    inputs, targets, sw = marshal.prepare_autoencoder_input_and_output_list(X, X, 1, I[0], V[..., 0], H0, return_sample_weight=True)
    names = vae.autoencoder.metrics_names
    assert names[0] == "loss" and names.count("decoder_loss") == 3
    hist = vae.autoencoder.fit(inputs, targets, epochs=1, batch_size=8, shuffle=False, sample_weight=sw, verbose=False)
    assert set(["loss", "decoder_loss_1", "decoder_acc_3", "composer_decoder_acc"]).issubset(hist.history.keys())
    ev = vae.autoencoder.evaluate(inputs, targets, batch_size=8, verbose=False)
    assert len(ev) == len(names) and np.isfinite(ev).all()
    z = vae.encoder.predict(marshal.prepare_encoder_input_list(X, I[0], V[..., 0]), batch_size=8)
    assert z.shape == (20, L)
    Y, Ih, Vh = vae.decoder.predict(marshal.prepare_decoder_input(z), batch_size=8)
    assert Y.shape == (20, T, 61) and Ih.shape == (20, 4, 16) and Vh.shape == (20, T, 1)
    assert np.allclose(Y.sum(-1), 1, atol=1e-4)
    outs = vae.autoencoder.predict(inputs, batch_size=8)
    assert [o.shape for o in outs] == [(20, T, 61), (20, 4, 16), (20, T, 1), (20, 2)]
    # weights round trip
    import tempfile, os
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "autoencoderEpoch0.pickle")
        vae.autoencoder.save_weights(path)
        before = vae.autoencoder.get_weights()
        vae.autoencoder.fit(inputs, targets, epochs=1, batch_size=8, sample_weight=sw)
        vae.autoencoder.load_weights(path, by_name=False)
        after = vae.autoencoder.get_weights()
        assert all(np.array_equal(a, b) for a, b in zip(before, after))
    vae.engine.close()


def test_per_song_training_driver_cfg1():
    """BASELINE config[0]: 2 styles x 10 synthetic songs, seq_len=16, hidden=64, latent=16, batch=8, driven through the
    per-song loop of vae_training.py:728-864 (history from the encoder after epoch 0).  Epoch 0 is checked against the
    oracle running the same loop; later epochs must keep improving."""
    from midi_vae_b200 import VAE, training
    T, H, L = 16, 64, 16
    vae = VAE().create(input_dim=61, output_dim=61, input_length=T, output_length=T, latent_rep_size=L, lstm_size=H, activation='softmax',
                       include_composer_decoder=True, num_composers=2, composer_weight=0.1, num_layers_encoder=2, num_layers_decoder=2,
                       learning_rate=2e-3, beta=0.1, extra_layer=True, meta_instrument=True, meta_instrument_dim=16, meta_instrument_length=4,
                       meta_instrument_activation='softmax', meta_instrument_weight=0.1, meta_velocity=True, meta_velocity_length=T,
                       meta_velocity_weight=1.0, epsilon_std=0.0, max_batch=8, decoder_feedback="teacher_forced")
    songs = synth.make_songs(20, T, seed=1235, min_chunks=8, max_chunks=20)
    # oracle, epoch 0 (history = zeros, eps = 0)
    ecfg = vae.engine.cfg
    _, ocfg = util.make_cfgs(T=T, H=H, L=L, feedback="teacher_forced", max_batch=8, lr=2e-3)
    p = util.to_torch(vae.engine.get_weights())
    opt = O.KerasAdam(p, lr=2e-3)
    per_song = []
    for s in songs[:4]:
        X, I, V, C = [torch.tensor(a) for a in s.dense(np.float64)]
        tot = 0.0
        for a in range(0, len(s), 8):
            b = min(len(s), a + 8)
            z0 = torch.zeros(b - a, L, dtype=torch.float64)
            m, _ = O.train_on_batch(ocfg, p, opt, X[a:b], I[a:b], V[a:b], C[a:b], z0, z0)
            tot += m["loss"] * (b - a)
        per_song.append(tot / len(s))
    m0 = training.train_epoch(vae, songs[:4], epoch=0, batch_size=8)
    assert abs(m0["loss"] - np.mean(per_song)) <= 2e-3 * np.mean(per_song), (m0["loss"], np.mean(per_song))
    assert abs(m0["kl_loss"] * 0.1 - (m0["loss"] - m0["decoder_loss_1"] - 0.1 * m0["decoder_loss_2"] - m0["decoder_loss_3"] - 0.1 * m0["composer_decoder_loss"])) < 1e-6
    losses = [training.train_epoch(vae, songs, epoch=e, batch_size=8)["loss"] for e in range(1, 4)]
    assert losses[-1] < losses[0] < m0["loss"] * 1.05, (m0["loss"], losses)
    ev = training.evaluate_songs(vae, songs[:5], batch_size=8)
    assert np.isfinite(list(ev.values())).all()
    vae.engine.close()


@pytest.mark.parametrize("scope", ["chunk", "song"])
def test_postprocess_on_device_matches_reference_rules(scope):
    """the override rules of process_decoder_outputs (vae_definition.py:1156-1190) run on the device.  Checked against
    midi_vae_b200.postprocess (host numpy, pinned to the output of the reference's own function by tests/test_reference_pin.py) on random packed
    rolls that exercise every rule: silent steps with velocity, new pitches without a struck velocity, struck velocities without a pitch."""
    from midi_vae_b200 import postprocess as PP
    T, n = 16, 24
    ecfg, _ = util.make_cfgs(T=T, H=64, L=16, max_batch=n)
    eng = Engine(ecfg, 0)
    rng = np.random.default_rng(11)
    pitch = rng.integers(0, 61, size=(n, T)).astype(np.uint8)
    pitch[rng.random((n, T)) < 0.25] = 60                       # silent
    pitch[rng.random((n, T)) < 0.1] = 0                         # pitch 0: "previous_pitch > 0" is false for it
    hold = rng.random((n, T)) < 0.4
    for i in range(4, T):                                       # sustained notes per voice (same pitch as the voice's previous step)
        pitch[:, i] = np.where(hold[:, i], pitch[:, i - 4], pitch[:, i])
    vel = rng.random((n, T)).astype(np.float32)
    vel[rng.random((n, T)) < 0.3] = 0.5                         # exactly on the threshold: "not silent"
    song_start = np.zeros(n, np.uint8); song_start[[0, 5, 6, 17]] = 1
    V, D = eng.postprocess(pitch, vel, song_start, scope)
    eng.close()
    bounds = [(i, i + 1) for i in range(n)] if scope == "chunk" else list(zip([0, 5, 6, 17], [5, 6, 17, n]))
    for a, b in bounds:                                         # the host function's memory runs over whatever one call is given
        _, _, Vr, Dr = PP.process_decoder_outputs(pitch[a:b], np.zeros((b - a, 4), np.uint8), vel[a:b])
        assert np.array_equal(V[a:b].reshape(-1).astype(np.float64), Vr), (scope, a, b)
        assert np.array_equal(D[a:b].reshape(-1), Dr.astype(np.uint8)), (scope, a, b)


def test_style_transfer_with_device_postprocess():
    """mvae_set_postprocess: style_transfer hands back velocities that already went through the override rules (per chunk, as the reference's loop)."""
    from midi_vae_b200 import postprocess as PP
    ecfg, _ = util.make_cfgs(T=16, H=64, L=16, max_batch=40)
    eng = _engine(ecfg, util.make_weights(ecfg, jitter=0.2))
    r = synth.concat(synth.make_songs(3, 16, seed=5, min_chunks=8, max_chunks=14))
    P0, I0, V0 = eng.style_transfer(r.pitch, r.instr, r.velocity, 0, 1, r.song_start, "as_wired")
    eng.set_postprocess("chunk")
    P1, I1, V1 = eng.style_transfer(r.pitch, r.instr, r.velocity, 0, 1, r.song_start, "as_wired")
    eng.close()
    assert np.array_equal(P0, P1) and np.array_equal(I0, I1)
    for i in range(len(P0)):
        _, _, Vr, _ = PP.process_decoder_outputs(P0[i:i + 1], I0[i:i + 1], V0[i:i + 1])
        assert np.array_equal(V1[i].astype(np.float64), Vr), i


@pytest.mark.parametrize("precision,shape", [("fp32", (16, 64, 16)), ("bf16", (16, 256, 32))])
def test_free_running_graph_replay_is_exact(precision, shape):
    """The free-running decoder is captured into a CUDA graph on its second call for a batch size and replayed from the third on
    (Model::decoder_stepwise): eager, capturing and replaying calls must return identical outputs, on new inputs too."""
    T, H, L = shape
    ecfg, ocfg = util.make_cfgs(T=T, H=H, L=L, precision=precision, max_batch=24)
    w = util.make_weights(ecfg)
    eng = _engine(ecfg, w)
    ra = synth.concat(synth.make_songs(2, T, seed=5, min_chunks=12, max_chunks=12))
    rb = synth.concat(synth.make_songs(2, T, seed=6, min_chunks=12, max_chunks=12))
    outs_a = [eng.style_transfer(ra.pitch, ra.instr, ra.velocity, 0, 1, ra.song_start, "free_running") for _ in range(3)]   # eager, capture, replay
    n0 = eng.launch_count()
    out_b = eng.style_transfer(rb.pitch, rb.instr, rb.velocity, 0, 1, rb.song_start, "free_running")                         # replay, other inputs
    n1 = eng.launch_count()
    out_a4 = eng.style_transfer(ra.pitch, ra.instr, ra.velocity, 0, 1, ra.song_start, "free_running")
    for o in outs_a[1:] + [out_a4]:
        for x, y in zip(outs_a[0], o):
            assert np.array_equal(x, y)
    assert n1 - n0 > 4 * T            # replayed launches are counted
    if precision == "fp32":
        p = util.to_torch(w)
        X, I, V, C = [torch.tensor(a) for a in rb.dense(np.float64)]
        ref = O.style_transfer(ocfg, p, X, I, V, 0, 1, rb.song_start, "free_running")
        safe = O.top2_margin(ref["Yh"]).numpy() > 1e-4
        assert np.array_equal(out_b[0][safe], ref["pitch"].numpy()[safe])
        assert np.abs(out_b[2] - ref["Vh"].numpy()[..., 0]).max() <= 2e-4
    eng.close()


def test_packed_training_epoch_matches_oracle_on_the_same_packs():
    """SURVEY 8(f-3) opt-in packing: several whole songs per mini-batch, histories from ONE batched encoder pass per pack, shifted inside each song.
    The oracle runs the same packs (same history rule): epoch 1 (history in use) must agree, and later packed epochs must keep improving."""
    from midi_vae_b200 import VAE, training
    from midi_vae_b200.marshal import shift_history
    T, H, L, B = 16, 64, 16, 32
    vae = VAE().create(input_dim=61, output_dim=61, input_length=T, output_length=T, latent_rep_size=L, lstm_size=H, activation='softmax',
                       include_composer_decoder=True, num_composers=2, composer_weight=0.1, num_layers_encoder=2, num_layers_decoder=2,
                       learning_rate=2e-3, beta=0.1, extra_layer=True, meta_instrument=True, meta_instrument_dim=16, meta_instrument_length=4,
                       meta_instrument_activation='softmax', meta_instrument_weight=0.1, meta_velocity=True, meta_velocity_length=T,
                       meta_velocity_weight=1.0, epsilon_std=0.0, max_batch=B, decoder_feedback="teacher_forced")
    songs = synth.make_songs(12, T, seed=77, min_chunks=5, max_chunks=14)
    _, ocfg = util.make_cfgs(T=T, H=H, L=L, feedback="teacher_forced", max_batch=B, lr=2e-3)
    p = util.to_torch(vae.engine.get_weights())
    opt = O.KerasAdam(p, lr=2e-3)
    tot, seen = 0.0, 0
    for pack in training.pack_songs(songs, B):
        X, I, V, C = [torch.tensor(a) for a in pack.dense(np.float64)]
        n = len(pack)
        zeros = torch.zeros(n, L, dtype=torch.float64)
        z = O.encode(ocfg, p, X, I, V, zeros)[0].numpy()
        Hh = torch.tensor(shift_history(z, pack.song_start))
        for a in range(0, n, B):
            b = min(n, a + B)
            m, _ = O.train_on_batch(ocfg, p, opt, X[a:b], I[a:b], V[a:b], C[a:b], Hh[a:b], zeros[a:b])
            tot += m["loss"] * (b - a); seen += b - a
    m1 = training.train_epoch_packed(vae, songs, epoch=1, batch_size=B)
    assert abs(m1["loss"] - tot / seen) <= 2e-3 * tot / seen, (m1["loss"], tot / seen)
    losses = [training.train_epoch_packed(vae, songs, epoch=e, batch_size=B)["loss"] for e in range(2, 5)]
    assert losses[-1] < losses[0] < m1["loss"] * 1.05, (m1["loss"], losses)
    vae.engine.close()


def test_history_from_the_batch_itself_equals_explicit_history():
    """Opt-in fused history (mvae_set_history_mode 1): the step builds H[i] = z[i-1] (0 at song starts, row 0 carried over from the
    previous call) from its own z on the device.  Same result as handing the step that history explicitly: z from an encoder pass with the same
    weights and the same epsilon, shifted on the host."""
    from midi_vae_b200.marshal import shift_history
    ecfg, _ = util.make_cfgs(T=16, H=64, L=16, feedback="teacher_forced", max_batch=12, lr=1e-3)
    w = util.make_weights(ecfg)
    a, b = _engine(ecfg, w), _engine(ecfg, w)
    b.set_history_mode(True)
    carry = None
    for step in range(3):
        r, _, eps, _ = util.make_batch(ecfg, 12, seed=40 + step)
        ss = np.zeros(12, np.uint8); ss[[4, 9]] = 1
        if step == 0:
            ss[0] = 1                                    # the first call starts a song; later calls continue the previous call's last song
        z = a.encode(r.pitch, r.instr, r.velocity, eps)[0]
        H = shift_history(z, ss)
        if step > 0:
            H[0] = carry
        carry = z[-1].copy()
        ma = a.train_on_batch(r.pitch, r.instr, r.velocity, r.style, H.astype(np.float32), eps)
        mb = b.train_on_batch(r.pitch, r.instr, r.velocity, r.style, None, eps, song_start=ss)
        for k in METRIC_KEYS:
            assert abs(ma[k] - mb[k]) <= 2e-6 * max(1.0, abs(ma[k])), (step, k, ma[k], mb[k])
        ga, gb = a.get_grads(), b.get_grads()
        for k in ga:
            assert np.abs(ga[k] - gb[k]).max() <= 1e-5 * max(np.abs(ga[k]).max(), 1e-6) + 1e-9, (step, k)
    a.close(); b.close()
