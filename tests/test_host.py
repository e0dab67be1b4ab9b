"""CPU tests of the host-side mirror of the reference interface: list marshalling with the reference's positional
layouts, roll packing, the synthetic roll generator, kwarg validation, the weight inventory and the work model."""
import os
import sys

import numpy as np
import pytest

from midi_vae_b200 import EngineConfig, VAE, initial_weights, marshal, reference_param_specs, synth
from midi_vae_b200.vae import _pack_onehot
from oracle import midivae_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synthetic_rolls_have_the_reference_layout():
    song = synth.make_song(np.random.default_rng(0), 5, 64, style=1)
    assert song.pitch.shape == (5, 64) and song.pitch.dtype == np.uint8 and song.pitch.max() <= 60
    assert song.instr.shape == (5, 4) and song.instr.max() < 16
    # velocity is 0 or in [0.5, 1] (import_midi.py:269-277) and non-zero only where a note sounds
    v = song.velocity
    assert np.all((v == 0) | ((v >= 0.5) & (v <= 1.0)))
    assert np.all(song.pitch[v > 0] != synth.SILENT)
    # voice interleaving: index = step*4 + voice (import_midi.py:245-249): each voice is a slow random walk
    voice0 = song.pitch.reshape(-1, 4)[:, 0].astype(int)
    sounding = voice0[voice0 != synth.SILENT]
    assert np.abs(np.diff(sounding)).max() <= 8
    X, I, V, C = song.dense()
    assert X.shape == (5, 64, 61) and I.shape == (5, 4, 16) and V.shape == (5, 64, 1) and C.shape == (5, 2)
    assert np.all(X.sum(-1) == 1) and np.all(C[:, 1] == 1)
    assert song.song_start[0] and not song.song_start[1:].any()
    b = synth.make_batch(6, 16, seed=3)
    assert list(b.style) == [0, 1, 0, 1, 0, 1]
    assert np.array_equal(synth.make_batch(6, 16, seed=3).pitch, b.pitch)          # seeded


def test_marshal_layouts_match_the_reference_lists():
    song = synth.make_song(np.random.default_rng(1), 7, 16, style=0)
    X, I, V, C = song.dense()
    H = np.random.default_rng(2).standard_normal((7, 8))
    enc = marshal.prepare_encoder_input_list(X, I[0], V[..., 0])
    assert [a.shape for a in enc] == [(7, 16, 61), (7, 4, 16), (7, 16, 1)]                       # vae_definition.py:798-806
    ins, outs, sw = marshal.prepare_autoencoder_input_and_output_list(X, X, 0, I[0], V[..., 0], H, return_sample_weight=True)
    # [X, Y_start, H, I_start, I, V_start, V]  (vae_definition.py:924,967,984-985,996-997)
    assert [a.shape for a in ins] == [(7, 16, 61), (7, 61), (7, 8), (7, 16), (7, 4, 16), (7,), (7, 16, 1)]
    assert [a.shape for a in outs] == [(7, 16, 61), (7, 4, 16), (7, 16, 1), (7, 2)]                # :926,986,998,1031
    assert [a.shape for a in sw] == [(7, 16), (7,), (7,), (7,)]
    assert not ins[1].any() and not ins[3].any() and not ins[5].any()                               # zero start vectors
    ins_tf, _ = marshal.prepare_autoencoder_input_and_output_list(X, X, 0, I[0], V[..., 0], H, teacher_force=True)
    assert len(ins_tf) == 8 and ins_tf[2].shape == (7, 16, 61)                                      # :963-964
    dec = marshal.prepare_decoder_input(H)
    assert [a.shape for a in dec] == [(7, 61), (7, 8), (7, 8), (7, 16), (7,)]                       # :816-865
    assert not dec[2][0].any() and np.array_equal(dec[2][1:], H[:-1])                               # history = R rolled by one
    w = marshal.prepare_autoencoder_input_and_output_list(X, X, 0, I[0], V[..., 0], H, silent_weight=0.25, return_sample_weight=True)[2][0]
    assert np.all(w[song.pitch == synth.SILENT] == 0.25) and np.all(w[song.pitch != synth.SILENT] == 1)
    Hs = marshal.shift_history(H, np.array([1, 0, 0, 1, 0, 0, 0], bool))
    assert not Hs[0].any() and not Hs[3].any() and np.array_equal(Hs[4], H[3])


def test_pack_onehot_rejects_non_rolls():
    X = np.eye(61)[np.array([[3, 60, 7]])]
    assert _pack_onehot(X, 61, "x").tolist() == [[3, 60, 7]]
    bad = X.copy(); bad[0, 0, 5] = 1
    with pytest.raises(ValueError):
        _pack_onehot(bad, 61, "x")
    with pytest.raises(ValueError):
        _pack_onehot(X[..., :60], 61, "x")


def _kwargs(**over):
    kw = dict(input_dim=61, output_dim=61, input_length=16, output_length=16, latent_rep_size=16, lstm_size=64, activation='softmax',
              include_composer_decoder=True, num_composers=2, composer_weight=0.1, num_layers_encoder=2, num_layers_decoder=2, learning_rate=2e-4,
              beta=0.1, extra_layer=True, meta_instrument=True, meta_instrument_dim=16, meta_instrument_length=4, meta_instrument_activation='softmax',
              meta_instrument_weight=0.1, meta_velocity=True, meta_velocity_length=16, meta_velocity_weight=1.0, epsilon_std=0.01, max_batch=8)
    kw.update(over)
    return kw


@pytest.mark.parametrize("over", [dict(cell_type='SimpleRNN'), dict(bidirectional=True), dict(use_embedding=True), dict(meta_held_notes=True),
                                  dict(meta_next_notes=True), dict(signature_decoder=True), dict(optimizer='RMSprop'), dict(activation='sigmoid'),
                                  dict(meta_velocity=False), dict(include_composer_decoder=False), dict(split_lstm_vector=False)])
def test_out_of_scope_branches_fail_loudly(over):
    with pytest.raises(NotImplementedError):
        VAE().create(**_kwargs(**over))


def test_reference_asserts_are_kept():
    with pytest.raises(AssertionError):
        VAE().create(**_kwargs(num_layers_encoder=0))
    with pytest.raises(AssertionError):
        VAE().create(**_kwargs(beta=0))


def test_weight_inventory_matches_the_oracle_and_the_survey():
    for T, H, L in ((16, 64, 16), (64, 256, 100), (256, 512, 256)):
        ecfg = EngineConfig(input_length=T, lstm_size=H, latent_rep_size=L)
        ocfg = O.OracleConfig(input_length=T, lstm_size=H, latent_rep_size=L)
        assert [(n, tuple(s)) for n, s in reference_param_specs(ecfg)] == [(n, tuple(s)) for n, s, _ in O.param_specs(ocfg)]
    w = initial_weights(EngineConfig(input_length=16, lstm_size=64, latent_rep_size=16), 1)
    U = w["lstm_1/recurrent_kernel"]
    assert np.allclose(U @ U.T, np.eye(64), atol=1e-5)                       # orthogonal recurrent kernel (Keras LSTM default)
    assert np.all(w["lstm_1/bias"][64:128] == 1) and w["lstm_1/bias"].sum() == 64    # unit forget bias
    assert np.abs(w["notes/cell_1/kernel"]).max() <= np.sqrt(6.0 / (61 + 256)) + 1e-6    # Glorot-uniform bound
    assert not w["notes/cell_1/bias"].any()


def test_work_model_matches_baseline_md():
    sys.path.insert(0, ROOT)
    import bench
    # forward / train GFLOP per sequence
    for (T, H, L), fwd in (((16, 64, 16), 5.72e6), ((64, 256, 100), 0.2926e9), ((256, 512, 256), 4.4652e9)):
        f, rec = bench.flops_per_seq(T, H, L, "teacher_forced")
        assert abs(f - fwd) / fwd < 5e-3, (T, H, L, f)   # BASELINE.md quotes 3-4 significant digits
        assert 0 < rec < f
    f_aw, _ = bench.flops_per_seq(256, 512, 256, "as_wired")
    assert abs(3 * f_aw - 13.1997e9) / 13.1997e9 < 2e-3


def _reference_postprocess_loop(pitch, vel, thr=0.5, mv=4):
    """Literal transcription of vae_definition.py:1143-1221 (argmax sampling already applied)."""
    p = pitch.reshape(-1); V = vel.astype(np.float64).reshape(-1).copy()
    Y = np.zeros((len(p), 60))
    for i, c in enumerate(p):
        if c != 60:
            Y[i, c] = 1
    for s in range(len(V)):
        if Y[s].sum() == 0:
            V[s] = 0
    for voice in range(mv):
        previous_pitch = -1; previous_velocity = 0.0
        for i, (nv, velocity) in enumerate(zip(Y[voice::mv], V[voice::mv])):
            pitch_is_silent = nv.sum() == 0
            pitch_ = -1 if pitch_is_silent else int(np.argmax(nv))
            velocity_is_silent = velocity < thr
            if velocity_is_silent:
                if (not pitch_is_silent) and previous_pitch > 0 and previous_pitch != pitch_:
                    V[i * mv + voice] = previous_velocity
            elif pitch_is_silent:
                V[i * mv + voice] = 0
            previous_pitch = pitch_
            if not velocity_is_silent:
                previous_velocity = velocity
    D = np.ones(len(V)); D[V > thr] = 0
    return Y, V, D


def test_postprocess_matches_the_reference_loop():
    from midi_vae_b200 import postprocess
    rng = np.random.default_rng(0)
    for trial in range(10):
        r = synth.make_batch(3, 32, seed=trial)
        vel = np.where(rng.random(r.velocity.shape) < 0.3, rng.random(r.velocity.shape), r.velocity).astype(np.float32)
        pit = np.where(rng.random(r.pitch.shape) < 0.2, 60, r.pitch).astype(np.uint8)
        Y, I, V, D = postprocess.process_decoder_outputs(pit, r.instr, vel)
        Y2, V2, D2 = _reference_postprocess_loop(pit, vel)
        assert np.array_equal(Y, Y2) and np.allclose(V, V2) and np.array_equal(D, D2)
        assert I.shape == (3, 4, 16) and np.all(I.sum(-1) == 1)


def test_packed_training_epoch_host_logic():
    """SURVEY 8(f-3) opt-in packing (training.train_epoch_packed): whole songs share mini-batches, every chunk is trained exactly once, in order,
    and the history shift restarts at each song start.  A recording stand-in replaces the engine: this is host logic only."""
    from types import SimpleNamespace
    from midi_vae_b200 import training
    songs = synth.make_songs(7, 16, seed=3, min_chunks=3, max_chunks=9)
    L = 4

    class FakeEngine:
        cfg = SimpleNamespace(latent_rep_size=L, input_dim=61, notes_weight=1.0, composer_weight=0.1, meta_instrument_weight=0.1,
                              meta_velocity_weight=1.0, beta=0.1)

        def __init__(self):
            self.trained, self.hist, self.batches = [], [], []

        def encode(self, P, I, V, eps):
            z = np.repeat(P[:, :1].astype(np.float32) + 1.0, L, axis=1)      # z of a chunk = its first pitch + 1: recognisable in the history
            return z, z, z

        def train_on_batch(self, P, I, V, style, H, eps, w=None):
            self.trained.append(P.copy()); self.hist.append(H.copy()); self.batches.append(len(P))
            return {k: 1.0 for k in training.METRIC_KEYS}

    eng = FakeEngine()
    vae = SimpleNamespace(engine=eng, max_batch=16, _eps=lambda n: np.zeros((n, L), np.float32))
    out = training.train_epoch_packed(vae, songs, epoch=1, batch_size=16)
    allp = np.concatenate([s.pitch for s in songs])
    assert np.array_equal(np.concatenate(eng.trained), allp)                       # every chunk once, in song order
    assert max(eng.batches) == 16 and sum(eng.batches) == len(allp)
    H = np.concatenate(eng.hist)
    starts = np.cumsum([0] + [len(s) for s in songs[:-1]])
    assert np.all(H[starts] == 0)                                                    # H = 0 on the first chunk of every song
    inner = np.setdiff1d(np.arange(len(allp)), starts)
    assert np.array_equal(H[inner, 0], allp[inner - 1, 0].astype(np.float32) + 1.0)  # H[i] = z[i-1] inside a song
    assert out["loss"] == 1.0 and "kl_loss" in out
    # fewer optimiser steps than the per-song loop, which is the point
    assert len(eng.batches) < sum(-(-len(s) // 16) for s in songs)
    packs = training.pack_songs(songs, 16)
    assert all(len(p) >= 16 for p in packs[:-1]) and all(p.song_start[0] == 1 for p in packs)
