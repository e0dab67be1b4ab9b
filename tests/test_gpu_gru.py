"""GRU branch: the reference's DEFAULT cell (settings.py:155; every shipped checkpoint) on the GPU against the CPU oracle.

Encoders are Keras 2.0.8 GRU layers (blocks [z|r|h], reset gate applied before the candidate's recurrent product, h' = z h + (1-z) hh),
decoders recurrentshop GRUCells (Dense(3H)+b on x, Dense(2H) and Dense(H) on h; h' = (1-z) h + z hh as recalled) -- the oracle's GRU branch is
pinned at 1e-9 to the reference's own graph code run at its default settings (tests/test_reference_pin.py).  The CUDA path runs them as
step-streamed recurrences (two dependent GEMMs + two pointwise launches per step and direction; the fp32 parity precision, any size) or, at the
reference's default size H = 256 in bf16, as ONE cluster-resident launch per recurrence and direction (csrc/gru_cluster.cu)."""
import os

import numpy as np
import pytest
import torch

from midi_vae_b200 import Engine, METRIC_KEYS, synth
from oracle import midivae_oracle as O
from tests import util
from tests.test_gpu_parity import _compare_step

pytestmark = pytest.mark.gpu
TOL32 = 1e-4


@pytest.mark.parametrize("feedback", ["as_wired", "teacher_forced"])
@pytest.mark.parametrize("gate", ["hard_sigmoid", "sigmoid"])
def test_gru_train_step_fp32_matches_oracle(feedback, gate):
    ecfg, ocfg = util.make_cfgs(T=16, H=64, L=16, feedback=feedback, gate=gate, max_batch=8, cell_type="GRU")
    _compare_step(ecfg, ocfg, 8, weights=True)


def test_gru_layers_and_ragged_batch_fp32():
    ecfg, ocfg = util.make_cfgs(T=8, H=32, L=12, ne=1, nd=3, feedback="teacher_forced", max_batch=6, cell_type="GRU")
    _compare_step(ecfg, ocfg, 5)
    ecfg, ocfg = util.make_cfgs(T=12, H=48, L=8, ne=3, nd=1, feedback="as_wired", max_batch=3, cell_type="GRU")
    _compare_step(ecfg, ocfg, 1)


def test_gru_reference_default_shape_bf16():
    """The reference's default shape (T64, H256, L256, settings.py:108-112) in the tensor-core precision: bf16 operands, fp32 accumulation."""
    ecfg, ocfg = util.make_cfgs(T=64, H=256, L=256, feedback="as_wired", precision="bf16", max_batch=16, cell_type="GRU")
    _compare_step(ecfg, ocfg, 16, tol=2e-2, grad_tol=6e-2)


@pytest.mark.parametrize("feedback,n", [("teacher_forced", 40), ("as_wired", 5), ("teacher_forced", 33)])
def test_gru_cluster_kernels_ragged_batches_bf16(feedback, n):
    """Cluster-resident GRU kernels (H = 256, bf16) on batches that are not a multiple of the 32 rows a cluster owns, both decoder feedbacks
    (teacher_forced exercises the dense / scalar input projections of the decoder cells), against the fp64 oracle."""
    ecfg, ocfg = util.make_cfgs(T=24, H=256, L=64, feedback=feedback, precision="bf16", max_batch=48, cell_type="GRU")
    _compare_step(ecfg, ocfg, n, tol=2e-2, grad_tol=6e-2)


def test_gru_cluster_matches_streamed_bf16():
    """The same bf16 step through the cluster-resident kernels (rnn_mode auto) and through the step-streamed kernels (rnn_mode streamed): metrics and
    every gradient tensor agree to bf16 rounding (the two paths round the candidate at different points)."""
    outs = []
    for mode in ("auto", "streamed"):
        ecfg, _ = util.make_cfgs(T=32, H=256, L=100, feedback="teacher_forced", precision="bf16", max_batch=72, cell_type="GRU", rnn_mode=mode)
        w = util.make_weights(ecfg)
        eng = Engine(ecfg, 0); eng.set_weights(w)
        r, hist, eps, _ = util.make_batch(ecfg, 72, seed=11)
        m = eng.train_on_batch(r.pitch, r.instr, r.velocity, r.style, hist, eps)
        outs.append((m, eng.get_grads(), eng.launch_count()))
        eng.close()
    (ma, ga, la), (mb, gb, lb) = outs
    assert la < lb / 4, (la, lb)          # the cluster path really ran: a fraction of the step-streamed launch count
    for k in METRIC_KEYS:
        assert abs(ma[k] - mb[k]) <= 1e-2 * max(1.0, abs(mb[k])), (k, ma[k], mb[k])
    for k in ga:
        scale = max(float(np.abs(gb[k]).max()), 1e-6)
        assert np.abs(ga[k] - gb[k]).max() <= 4e-2 * scale + 1e-9, (k, float(np.abs(ga[k] - gb[k]).max()), scale)


def test_gru_evaluate_predict_decode_fp32():
    ecfg, ocfg = util.make_cfgs(T=16, H=64, L=16, feedback="teacher_forced", max_batch=8, cell_type="GRU")
    w = util.make_weights(ecfg)
    eng = Engine(ecfg, 0); eng.set_weights(w)
    p = util.to_torch(w)
    r, hist, eps, sw = util.make_batch(ecfg, 8, weights=True)
    X, I, V, C, th, te, tsw = util.oracle_inputs(ocfg, r, hist, eps, sw)
    m_ref, outs, aux = O.evaluate_batch(ocfg, p, X, I, V, C, th, te, sample_weight=tsw)
    m = eng.evaluate_batch(r.pitch, r.instr, r.velocity, r.style, hist, eps, sw)
    for k in METRIC_KEYS:
        assert abs(m[k] - m_ref[k]) <= TOL32 * max(1.0, abs(m_ref[k])), k
    z, mu, lv = eng.encode(r.pitch, r.instr, r.velocity, eps)
    assert util.rel_err(z, aux[0].numpy()) <= TOL32 and util.rel_err(mu, aux[1].numpy()) <= TOL32 and util.rel_err(lv, aux[2].numpy()) <= TOL32
    for fb in ("as_wired", "teacher_forced", "free_running"):
        Yr, Ir, Vr = O.decode(ocfg, p, aux[0], th, X, I, V, fb)
        Yd, Id, Vd = eng.decode(z, hist, fb, r.pitch, r.instr, r.velocity)
        assert util.rel_err(Yd, Yr.numpy()) <= 2 * TOL32, fb
        assert util.rel_err(Id, Ir.numpy()) <= 2 * TOL32, fb
        assert util.rel_err(Vd, Vr.numpy()[..., 0]) <= 2 * TOL32, fb
    eng.close()


@pytest.mark.parametrize("feedback", ["as_wired", "free_running"])
def test_gru_style_transfer_argmax_exact(feedback):
    ecfg, ocfg = util.make_cfgs(T=16, H=64, L=16, max_batch=40, cell_type="GRU")
    w = util.make_weights(ecfg, jitter=0.2)
    eng = Engine(ecfg, 0); eng.set_weights(w)
    p = util.to_torch(w)
    r = synth.concat(synth.make_songs(3, 16, seed=5, min_chunks=8, max_chunks=14))
    X, I, V, C = [torch.tensor(a) for a in r.dense(np.float64)]
    ref = O.style_transfer(ocfg, p, X, I, V, 0, 1, r.song_start, feedback)
    P, Ii, Vv = eng.style_transfer(r.pitch, r.instr, r.velocity, 0, 1, r.song_start, feedback)
    safe_p = O.top2_margin(ref["Yh"]).numpy() > 1e-4
    safe_i = O.top2_margin(ref["Ih"]).numpy() > 1e-4
    assert np.array_equal(P[safe_p], ref["pitch"].numpy()[safe_p])
    assert np.array_equal(Ii[safe_i], ref["instr"].numpy()[safe_i])
    assert util.rel_err(Vv, ref["Vh"].numpy()[..., 0]) <= 2e-4
    eng.close()


def _reference_settings_kwargs(T=64, H=256, L=256, **over):
    """The 61 keyword arguments vae_training.py:47-109 passes, at settings.py's values (cell_type='GRU', teacher_force=False, ...)."""
    kw = dict(input_dim=61, output_dim=61, use_embedding=False, embedding_dim=0, input_length=T, output_length=T, latent_rep_size=L,
              vae_loss='categorical_crossentropy', optimizer='Adam', activation='softmax', lstm_activation='tanh', lstm_state_activation='tanh',
              epsilon_std=0.01, epsilon_factor=0.0, include_composer_decoder=True, num_composers=2, composer_weight=0.1, lstm_size=H, cell_type='GRU',
              num_layers_encoder=2, num_layers_decoder=2, bidirectional=False, decode=True, teacher_force=False, learning_rate=0.0002,
              split_lstm_vector=True, history=True, beta=0.1, prior_mean=0.0, prior_std=1.0, decoder_additional_input=False,
              decoder_additional_input_dim=0, extra_layer=True, meta_instrument=True, meta_instrument_dim=16, meta_instrument_length=4,
              meta_instrument_activation='softmax', meta_instrument_weight=0.1, signature_decoder=False, signature_dim=0,
              signature_activation='tanh', signature_weight=1.0, composer_decoder_at_notes_output=False, composer_decoder_at_notes_weight=1.0,
              composer_decoder_at_notes_activation='softmax', composer_decoder_at_instrument_output=False,
              composer_decoder_at_instrument_weight=1.0, composer_decoder_at_instrument_activation='softmax', meta_velocity=True,
              meta_velocity_length=T, meta_velocity_activation='sigmoid', meta_velocity_weight=1.0, meta_held_notes=False,
              meta_held_notes_length=T, meta_held_notes_activation='softmax', meta_held_notes_weight=1.0, meta_next_notes=False,
              meta_next_notes_output_length=T, meta_next_notes_weight=1.0, meta_next_notes_teacher_force=False, activation_before_splitting='tanh')
    kw.update(over)
    return kw


def test_gru_facade_with_unmodified_reference_settings_and_keras_hdf5_round_trip(tmp_path):
    """VAE().create(**the reference's own kwargs) works as is (GRU default); fit / evaluate / predict run; save_weights writes Keras-2.0.8 HDF5
    (layer groups of the shipped files) that load_weights reads back bit-exactly, sub-model files included."""
    from midi_vae_b200 import VAE, hdf5, marshal
    T, L = 64, 256
    vae = VAE().create(**_reference_settings_kwargs(), max_batch=16)
    assert vae.engine.cfg.cell_type == "GRU"
    song = synth.make_song(np.random.default_rng(0), 16, T, style=1)
    X, I, V, C = song.dense(np.float64)
    inputs, targets, sw = marshal.prepare_autoencoder_input_and_output_list(X, X, 1, I[0], V[..., 0], np.zeros((16, L)), return_sample_weight=True)
    h = vae.autoencoder.fit(inputs, targets, epochs=1, batch_size=16, shuffle=False, sample_weight=sw)
    assert np.isfinite(h.history["loss"][0])
    ev = vae.autoencoder.evaluate(inputs, targets, batch_size=16)
    assert len(ev) == 9 and np.isfinite(ev).all()
    for model, fname in ((vae.autoencoder, "autoencoderEpoch1.pickle"), (vae.encoder, "encoderEpoch1.pickle"), (vae.decoder, "decoderEpoch1.pickle")):
        path = str(tmp_path / fname)
        model.save_weights(path)
        assert open(path, "rb").read(8) == hdf5.SIGNATURE
        before = model.get_weights()
        lay = hdf5.layout(path)
        assert lay[0][1] == ("gru_1/kernel" if model is not vae.decoder else "dense_8/kernel")
        vae.autoencoder.fit(inputs, targets, epochs=1, batch_size=16, sample_weight=sw)        # move the weights
        model.load_weights(path)
        assert all(np.array_equal(a, b) for a, b in zip(before, model.get_weights()))
    with pytest.raises(ValueError):
        vae.autoencoder.fit(inputs, targets, epochs=1, batch_size=32)       # larger than max_batch: never silently re-batched
    vae.engine.close()


def _shipped(name):
    for root in ("/root/reference/models", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "models")):
        p = os.path.join(root, name)
        if os.path.exists(p):
            return p
    return None


@pytest.mark.skipif(_shipped("JvP/autoencoderEpoch440.pickle") is None, reason="shipped checkpoint not present (oracle/_ref/models is filled by __graft_entry__.build() where /root/reference exists)")
def test_gru_shipped_checkpoint_loads_and_matches_oracle():
    """models/JvP/autoencoderEpoch440.pickle (Keras 2.0.8 HDF5, GRU, T64/H256/L256) loaded positionally as the reference does
    (vae_evaluation.py:553-559), then encoder / decoder / autoencoder outputs against the oracle on the same tensors (fp32, 1e-4)."""
    from midi_vae_b200 import VAE, hdf5
    T, L, n = 64, 256, 16
    vae = VAE().create(**_reference_settings_kwargs(), max_batch=n, precision="fp32")
    path = _shipped("JvP/autoencoderEpoch440.pickle")
    vae.autoencoder.load_weights(path, by_name=False)
    w = vae.engine.get_weights()
    t = hdf5.read_weights(path)
    assert np.array_equal(w["gru_1/kernel"], dict(t["layers"]["gru_1"])["gru_1/kernel"])
    dec = dict(t["layers"]["decoder"])
    assert np.array_equal(w["notes/cell_1/recurrent_kernel_1"], dec["gru_cell_1/dense_2/kernel"]) and np.array_equal(w["notes/cell_2/recurrent_kernel_2"], dec["gru_cell_2/dense_6/kernel"])
    assert np.array_equal(w["dec_init/vel_s1/kernel"], dec["dense_19/kernel"]) and np.array_equal(w["meta_velocity/out/kernel"], dec["dense_18/kernel"])
    _, ocfg = util.make_cfgs(T=T, H=256, L=L, feedback="as_wired", max_batch=n, cell_type="GRU")
    p = util.to_torch(w)
    song = synth.make_song(np.random.default_rng(3), n, T, style=0)
    X, I, V, C = [torch.tensor(a) for a in song.dense(np.float64)]
    with torch.no_grad():
        z_ref, mu_ref, _ = O.encode(ocfg, p, X, I, V, None)
        hist = O.shift_history(z_ref, song.song_start)
        Yr, Ir, Vr = O.decode(ocfg, p, z_ref, hist, feedback="as_wired")[:3]
    z, mu, lv = vae.engine.encode(song.pitch, song.instr, song.velocity, None)
    assert util.rel_err(z, z_ref.numpy()) <= TOL32
    Y, Ih, Vh = vae.engine.decode(z_ref.numpy().astype(np.float32), hist.numpy().astype(np.float32), "as_wired")
    assert util.rel_err(Y, Yr.numpy()) <= 2 * TOL32 and util.rel_err(Ih, Ir.numpy()) <= 2 * TOL32 and util.rel_err(Vh, Vr.numpy()[..., 0]) <= 2 * TOL32
    vae.engine.close()
