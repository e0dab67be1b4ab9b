"""Style classifiers (pitch_classifier.py:89-103, velocity_classifier.py:110-118, instrument_classifier.py:93-103) on the device,
against the oracle's restatement of the same Keras graph (stacked Keras GRU -> Dense softmax, categorical cross-entropy, accuracy, Adam)."""
import os

import numpy as np
import pytest
import torch

from midi_vae_b200 import synth
from midi_vae_b200.classifier import StyleClassifier, ensemble_prediction
from oracle import midivae_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _data(kind, n, T, seed=3):
    r = synth.make_batch(n, T if kind != "instrument" else 16, seed=seed)
    if kind == "pitch":
        X = np.eye(61)[r.pitch]
    elif kind == "velocity":
        X = r.velocity[..., None].astype(np.float64)
    else:
        X = np.eye(16)[r.instr]
        X[0, 3] = 0                                        # an unused voice: all-zero row
    Y = np.eye(2)[r.style]
    return X, Y


@pytest.mark.parametrize("kind,T,D", [("pitch", 16, 61), ("velocity", 16, 1), ("instrument", 4, 16)])
def test_classifier_step_matches_oracle_fp32(kind, T, D):
    H, n, lr = 64, 24, 1e-3
    clf = StyleClassifier(kind, input_dim=D, input_length=T, lstm_size=H, num_layers=2, learning_rate=lr, precision="fp32", max_batch=32, seed=5)
    w = clf.get_weights()
    rng = np.random.default_rng(9)
    w = {k: (v + 0.05 * rng.standard_normal(v.shape)).astype(np.float32) for k, v in w.items()}       # biases carry signal
    clf.set_weights(w)
    X, Y = _data(kind, n, T)
    ocfg = O.OracleConfig(input_length=T, lstm_size=H, cell_type="GRU")
    p = {k: torch.tensor(v, dtype=torch.float64) for k, v in w.items()}
    m_ref, g_ref, probs_ref = O.classifier_loss_and_grads(ocfg, p, torch.tensor(X, dtype=torch.float64), torch.tensor(Y, dtype=torch.float64), 2)
    probs = clf.predict(X, batch_size=32)
    assert np.abs(probs - probs_ref.numpy()).max() <= 1e-5
    loss, acc = clf.evaluate(X, Y, batch_size=32)
    assert abs(loss - m_ref["loss"]) <= 1e-5 * max(1.0, m_ref["loss"]) and abs(acc - m_ref["acc"]) < 1e-6
    loss_t, acc_t = clf.train_on_batch(X, Y)
    assert abs(loss_t - m_ref["loss"]) <= 1e-5 * max(1.0, m_ref["loss"])
    g = clf.get_grads()
    for k, v in g_ref.items():
        ref = v.numpy()
        assert np.abs(g[k].reshape(ref.shape) - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1e-6) + 1e-9, k
    opt = O.KerasAdam(p, lr=lr)
    opt.step(p, {k: torch.tensor(np.asarray(g[k], np.float64).reshape(p[k].shape)) for k in p})
    w_new = clf.get_weights()
    for k in p:
        assert np.abs((w_new[k] - w[k]).reshape(p[k].shape) - (p[k].numpy() - w[k].reshape(p[k].shape))).max() <= 2e-3 * lr + 1e-7, k
    clf.close()


def test_classifier_learns_a_separable_style():
    """fit() over consecutive mini-batches (shuffle=False, as the scripts call it): two 'styles' that differ in register become separable."""
    T, n = 16, 128
    rng = np.random.default_rng(1)
    style = rng.integers(0, 2, n)
    pitch = np.where(style[:, None] == 0, rng.integers(0, 25, (n, T)), rng.integers(35, 60, (n, T))).astype(np.uint8)
    clf = StyleClassifier("pitch", input_length=T, lstm_size=64, learning_rate=5e-3, precision="fp32", max_batch=32, seed=2)
    first = clf.evaluate(pitch, style, batch_size=32)
    for _ in range(12):
        h = clf.fit(pitch, style, epochs=1, batch_size=32)
    last = clf.evaluate(pitch, style, batch_size=32)
    assert last[0] < 0.5 * first[0] and last[1] >= 0.95, (first, last, h.history)
    clf.close()


@pytest.mark.parametrize("kind,fname,T,D", [("pitch", "pitch_classifier_epoch_160.pickle", 64, 61), ("velocity", "velocity_classifier_epoch_130.pickle", 64, 1),
                                            ("instrument", "instrument_classifier_epoch_170.pickle", 4, 16)])
def test_shipped_classifier_checkpoints_load_and_match_oracle(kind, fname, T, D, tmp_path):
    """The reference's trained evaluators (models/JvP/*_classifier_epoch_*.pickle, Keras HDF5): loaded as they are, the device reproduces the oracle's
    probabilities; save_weights writes a file with the same layer / weight layout that loads back identically."""
    path = os.path.join(ROOT, "oracle", "_ref", "models", "JvP", fname)
    if not os.path.exists(path):
        pytest.skip("shipped classifier checkpoint not staged (oracle/_ref is filled by __graft_entry__.build() where /root/reference exists)")
    clf = StyleClassifier(kind, input_dim=D, input_length=T, lstm_size=256, num_layers=2, precision="fp32", max_batch=32)
    clf.load_weights(path)
    w = clf.get_weights()
    X, Y = _data(kind, 20, T, seed=11)
    ocfg = O.OracleConfig(input_length=T, lstm_size=256, cell_type="GRU")
    p = {k: torch.tensor(v, dtype=torch.float64) for k, v in w.items()}
    ref = O.classifier_forward(ocfg, p, torch.tensor(X, dtype=torch.float64), 2).numpy()
    got = clf.predict(X, batch_size=32)
    assert np.abs(got - ref).max() <= 2e-5
    out = str(tmp_path / "resaved.pickle")
    clf.save_weights(out)
    from midi_vae_b200 import hdf5
    strip = lambda rows: [(a, b.split(":")[0], c) for a, b, c in rows]
    assert strip(hdf5.layout(out)) == strip(hdf5.layout(path))
    clf2 = StyleClassifier(kind, input_dim=D, input_length=T, lstm_size=256, num_layers=2, precision="fp32", max_batch=32, seed=99)
    clf2.load_weights(out)
    assert np.array_equal(clf2.predict(X, batch_size=32), got)
    clf.close(); clf2.close()


def test_ensemble_prediction_is_the_weighted_mean():
    T = 16
    r = synth.make_batch(10, T, seed=4)
    mods = [StyleClassifier("pitch", input_length=T, lstm_size=64, precision="fp32", max_batch=16, seed=1),
            StyleClassifier("instrument", input_dim=16, input_length=4, lstm_size=64, precision="fp32", max_batch=16, seed=2),
            StyleClassifier("velocity", input_length=T, lstm_size=64, precision="fp32", max_batch=16, seed=3)]
    e = ensemble_prediction(mods[0], mods[1], mods[2], r.pitch, r.instr, r.velocity, weights=(2.0, 1.0, 1.0))
    ref = (2.0 * mods[0].predict(r.pitch) + mods[1].predict(r.instr) + mods[2].predict(r.velocity)) / 4.0
    assert np.allclose(e, ref) and np.allclose(e.sum(-1), 1.0, atol=1e-5)
    for m in mods:
        m.close()
