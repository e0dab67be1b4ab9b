"""Generate REFERENCE-EXECUTED golden vectors: the reference's own, unmodified ``vae_definition.py`` is imported from
/root/reference and run -- ``VAE().create(**the kwargs vae_training.py:47-109 passes)``, ``prepare_*`` list builders,
``encoder/decoder/autoencoder.predict``, ``autoencoder.evaluate``, ``autoencoder.fit`` -- on top of the restated Keras 2.0.8 /
recurrentshop slice in oracle/keras_shim (neither library exists offline; see that directory's README for what is and is not
independent evidence).  Run where /root/reference exists:

    python tests/golden/make_reference_golden.py

Writes tests/golden/reference_cfg1.npz (LSTM branch, BASELINE configs[0] shapes: seq_len 16, hidden 64, latent 16, batch 8) and
tests/golden/reference_layout.json (layer / weight names, shapes and save order produced by the reference's graph code for
its DEFAULT GRU configuration -- compared in the tests with the shipped HDF5 checkpoints).
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MIDIVAE_REFERENCE", "/root/reference")


def import_reference():
    """vae_definition.py, unmodified, with keras / recurrentshop resolved to oracle/keras_shim."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "keras_shim"))
    sys.path.insert(0, ROOT)
    sys.path.append(REF)
    # data_class.py (plots; imports matplotlib, matplotlib2tikz, pretty_midi) is imported by vae_definition.py:11 but never used in it
    sys.modules.setdefault("data_class", types.ModuleType("data_class"))
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="mvae_ref_"))      # settings.py:60-63 creates ./pickles/<timestamp>/ at import
    try:
        import vae_definition as vd
    finally:
        os.chdir(cwd)
    return vd


def create_kwargs(vd, **over):
    """Exactly the keyword list of vae_training.py:47-109, values taken from the reference's settings module."""
    S = sys.modules["settings"]
    names = ["input_dim", "output_dim", "use_embedding", "embedding_dim", "input_length", "output_length", "vae_loss", "optimizer",
             "activation", "lstm_activation", "lstm_state_activation", "epsilon_std", "epsilon_factor", "include_composer_decoder",
             "num_composers", "composer_weight", "lstm_size", "cell_type", "num_layers_encoder", "num_layers_decoder", "bidirectional",
             "decode", "teacher_force", "learning_rate", "split_lstm_vector", "history", "beta", "prior_mean", "prior_std",
             "decoder_additional_input", "decoder_additional_input_dim", "extra_layer", "meta_instrument", "meta_instrument_dim",
             "meta_instrument_length", "meta_instrument_activation", "meta_instrument_weight", "signature_decoder", "signature_dim",
             "signature_activation", "signature_weight", "composer_decoder_at_notes_output", "composer_decoder_at_notes_weight",
             "composer_decoder_at_notes_activation", "composer_decoder_at_instrument_output", "composer_decoder_at_instrument_weight",
             "composer_decoder_at_instrument_activation", "meta_velocity", "meta_velocity_length", "meta_velocity_activation",
             "meta_velocity_weight", "meta_held_notes", "meta_held_notes_length", "meta_held_notes_activation", "meta_held_notes_weight",
             "meta_next_notes", "meta_next_notes_output_length", "meta_next_notes_weight", "meta_next_notes_teacher_force",
             "activation_before_splitting"]
    kw = {n: getattr(S, n) for n in names}
    kw["latent_rep_size"] = S.latent_dim
    kw.update(over)
    return kw


def set_module_lengths(vd, T):
    """prepare_* (vae_definition.py:770-1045) read the sequence length from the settings globals star-imported into the module."""
    vd.input_length = T
    vd.output_length = T
    vd.meta_velocity_length = T
    vd.meta_held_notes_length = T


def keras_name_map(cfg):
    """our parameter name -> (sub-model, keras layer, index inside layer.weights) for the LSTM branch."""
    m = {}
    pre = "gru" if getattr(cfg, "cell_type", "LSTM") == "GRU" else "lstm"
    enc = [f"{pre}_{k}" for k in range(1, cfg.num_layers_encoder + 1)] + [f"{pre}_meta_instrument", f"{pre}_meta_velocity"]
    for l in enc:
        for i, t in enumerate(["kernel", "recurrent_kernel", "bias"]):
            m[f"{l}/{t}"] = ("encoder", l, i)
    for l in ["extra_instrument_after_concat_layer"] + (["extra_layer"] if cfg.extra_layer else []) + ["z_mean", "z_log_var"]:
        for i, t in enumerate(["kernel", "bias"]):
            m[f"{l}/{t}"] = ("encoder", l, i)
    return m


def load_into_reference(model, cfg, w):
    """Copy our named weights into the reference-built Keras models.  Encoder layers are named by the reference; decoder tensors
    are positional: the 'decoder' model saves [init-state Denses (notes l1 s1, s2, l2 s1, s2, instr s1, s2, vel s1, s2)], then per
    RecurrentModel [cells (kernel, bias, recurrent kernel), output Dense] -- the order oracle.param_specs lists them in."""
    for k, (sub, layer, idx) in keras_name_map(cfg).items():
        l = model.encoder.get_layer(layer)
        arrs = l.get_weights()
        arrs[idx] = w[k]
        l.set_weights(arrs)
    from oracle import midivae_oracle as O
    dec_names = [n for n, _, _ in O.param_specs(cfg) if n.startswith(("dec_init/", "notes/", "meta_instrument/", "meta_velocity/"))]
    model.decoder.set_weights([w[n] for n in dec_names])
    return dec_names


VARIANTS = {
    # name: (create-kwarg overrides, oracle-config overrides, module globals read by prepare_*, batch_size)
    "tf_list": (dict(teacher_force=True), dict(), dict(teacher_force=True), 8),                       # extra ground-truth input, same computation
    "plain": (dict(history=False, extra_layer=False, num_layers_encoder=1, num_layers_decoder=1),
              dict(history=False, extra_layer=False, num_layers_encoder=1, num_layers_decoder=1), dict(history=False), 8),
    "deep": (dict(num_layers_encoder=3, num_layers_decoder=3), dict(num_layers_encoder=3, num_layers_decoder=3), dict(), 8),
    "weights": (dict(), dict(), dict(silent_weight=0.25), 3),                                          # temporal weights != 1, ragged mini-batches 3 + 3 + 2
    "gru": (dict(cell_type="GRU"), dict(cell_type="GRU"), dict(), 8),                                  # the reference's shipped default cell (settings.py:155)
}


def run_variants(vd):
    """Non-default corners of the reference's own graph / list code (vae_definition.py:262-266, 483-487, 548-551, 928-933) -> reference_cfg1_variants.npz"""
    from keras import backend as K
    import recurrentshop.cells as rs_cells
    from dataclasses import replace
    from midi_vae_b200 import synth
    from oracle import midivae_oracle as O
    from tests import util
    T, H, L, n = 16, 64, 16, 8
    out = {}
    r = synth.make_song(np.random.default_rng(4321), n, T, 0)
    rng = np.random.default_rng(4322)
    hist = (rng.standard_normal((n, L)) * 0.1).astype(np.float32)
    eps = synth.make_eps(n, L, 4321, 0.01)
    X, I_, V3, _ = r.dense(np.float64)
    for name, (kw, okw, glob, bs) in VARIANTS.items():
        _, ocfg = util.make_cfgs(T=T, H=H, L=L, feedback="as_wired", variant="standard", max_batch=n)
        ocfg = replace(ocfg, **okw)
        w = {k: v.numpy().astype(np.float32) for k, v in O.init_params(ocfg, seed=77).items()}
        jit = np.random.default_rng(78)
        w = {k: (v + 0.1 * jit.standard_normal(v.shape)).astype(np.float32) for k, v in w.items()}
        rs_cells.LSTM_VARIANT = "standard"
        K.clear_session()
        set_module_lengths(vd, T)
        saved = {k: getattr(vd, k) for k in glob}
        for k, v in glob.items():
            setattr(vd, k, v)
        try:
            model = vd.VAE()
            ckw = dict(cell_type="LSTM", input_length=T, output_length=T, lstm_size=H, latent_rep_size=L, meta_velocity_length=T,
                       meta_held_notes_length=T, meta_next_notes_output_length=T)
            ckw.update(kw)
            rs_cells.GRU_GATE_ORDER, rs_cells.GRU_MIX = "zr", "z_takes_new"
            model.create(**create_kwargs(vd, **ckw))
            load_into_reference(model, ocfg, w)
            K.set_random_normal_hook(lambda shp, mean, std: eps.astype(np.float64)[:shp[0]] * (std / 0.01) + mean)
            V = V3[..., 0]
            D = np.zeros_like(V); S = np.zeros((n, 15))
            in_list, out_list, sw = vd.prepare_autoencoder_input_and_output_list(X, X, int(r.style[0]), I_[0], V, D, S, hist.astype(np.float64), return_sample_weight=True)
            z = model.encoder.predict(vd.prepare_encoder_input_list(X, I_[0], V, D), batch_size=n)
            dec_out = model.decoder.predict(vd.prepare_decoder_input(z, int(r.style[0]), S, hist.astype(np.float64)), batch_size=n)
            # mini-batches of bs: the hook hands out eps rows from 0 for every call, so feed the batches' eps explicitly through a cursor
            cursor = {"i": 0}
            def hook(shp, mean, std):
                i = cursor["i"]; cursor["i"] = (i + shp[0]) % n
                return eps.astype(np.float64)[i:i + shp[0]] * (std / 0.01) + mean
            K.set_random_normal_hook(hook)
            ev = model.autoencoder.evaluate(in_list, out_list, batch_size=bs, sample_weight=sw, verbose=0)
            fits = []
            for _ in range(2):
                cursor["i"] = 0
                h = model.autoencoder.fit(in_list, out_list, epochs=1, batch_size=bs, shuffle=False, sample_weight=sw, verbose=0)
                fits.append([h.history[k][0] for k in sorted(h.history)])
            p = f"{name}/"
            out[p + "z"] = z
            for k, a in zip(("Y", "I", "V"), dec_out):
                out[p + "dec_" + k] = a
            out[p + "evaluate"] = np.array(ev)
            out[p + "fit_keys"] = np.array(sorted(h.history))
            out[p + "fit"] = np.array(fits)
            out[p + "n_inputs"] = np.array(len(in_list))
            out[p + "in_shapes"] = np.array([str(np.asarray(a).shape) for a in in_list])
            out[p + "sw_notes"] = np.asarray(sw[0])
            out[p + "batch_size"] = np.array(bs)
            if name in ("plain", "weights", "gru"):          # updated weights after the two epochs (kept for two variants only: fixture size)
                dec_names = [nm for nm, _, _ in O.param_specs(ocfg) if nm.startswith(("dec_init/", "notes/", "meta_instrument/", "meta_velocity/"))]
                for nm, a in zip(dec_names, model.decoder.get_weights()):
                    out[p + "w2/" + nm] = a
                for k, (sub, layer, idx) in keras_name_map(ocfg).items():
                    out[p + "w2/" + k] = model.encoder.get_layer(layer).get_weights()[idx]
        finally:
            for k, v in saved.items():
                setattr(vd, k, v)
    out["pitch"], out["instr"], out["velocity"], out["style"], out["hist"], out["eps"] = r.pitch, r.instr, r.velocity, r.style, hist, eps
    path = os.path.join(HERE, "reference_cfg1_variants.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


def run_postprocess(vd):
    """The reference's process_decoder_outputs (vae_definition.py:1131-1225, pure numpy, sample_method='argmax') on random decoder outputs
    -> reference_postprocess.npz.  No restated library is involved in this one."""
    T, n = 32, 6
    set_module_lengths(vd, T)
    rng = np.random.default_rng(2024)
    Yp = rng.random((n, T, 61)) ** 4
    Yp[..., 60] *= 3.0                                   # a good share of silent steps
    Yp /= Yp.sum(-1, keepdims=True)
    Ip = rng.random((n, 4, 16)); Ip /= Ip.sum(-1, keepdims=True)
    Vp = rng.random((n, T, 1))                           # velocities on both sides of the 0.5 played-note threshold
    Vp[rng.random((n, T, 1)) < 0.3] *= 0.4
    Y, I, V, D, N = vd.process_decoder_outputs([Yp.copy(), Ip.copy(), Vp.copy()], "argmax")
    path = os.path.join(HERE, "reference_postprocess.npz")
    np.savez_compressed(path, Yp=Yp.astype(np.float32), Ip=Ip.astype(np.float32), Vp=Vp.astype(np.float32)[..., 0],
                        Y=Y.astype(np.uint8), I=I.astype(np.uint8), V=V, D=D.astype(np.uint8), T=np.array(T))
    print(path, os.path.getsize(path), "bytes; sounding steps", int(Y.sum()), "of", Y.shape[0])


def run_training_loop(vd):
    """The reference's OWN training loop -- the source lines of vae_training.py from "# Train model" (:723) to the end of the per-epoch train
    aggregation (:959), exec'd unmodified -- over five synthetic songs for three epochs with the shim-built LSTM model: history latents (zeros in
    epoch 0, encoder.predict rolled by one chunk afterwards), one fit per song, per-song means, KL recovery.  -> reference_training_loop.npz"""
    import contextlib
    import io
    import re
    from keras import backend as K
    import recurrentshop.cells as rs_cells
    from midi_vae_b200 import synth
    from oracle import midivae_oracle as O
    from tests import util
    T, H, L, bs, epochs, n_songs = 16, 64, 16, 8, 3, 5
    src_lines = open(os.path.join(REF, "vae_training.py")).read().split("\n")
    first = next(i for i, l in enumerate(src_lines) if l.startswith("# Train model"))
    last = next(i for i, l in enumerate(src_lines) if l.startswith('    print("Total train loss: "'))
    code = "\n".join(src_lines[first:last + 1])
    assert "autoencoder.fit(input_list, output_list" in code and "H[1:] = representation_list[:-1]" in code
    ecfg, ocfg = util.make_cfgs(T=T, H=H, L=L, feedback="as_wired", variant="standard", max_batch=bs, lr=2e-3)
    w = util.make_weights(ecfg, seed=52, jitter=0.1)
    rs_cells.LSTM_VARIANT = "standard"
    K.clear_session()
    set_module_lengths(vd, T)
    model = vd.VAE()
    model.create(**create_kwargs(vd, cell_type="LSTM", input_length=T, output_length=T, lstm_size=H, latent_rep_size=L, meta_velocity_length=T,
                                 meta_held_notes_length=T, meta_next_notes_output_length=T, learning_rate=2e-3))
    load_into_reference(model, ocfg, w)
    songs = synth.make_songs(n_songs, T, seed=777, min_chunks=5, max_chunks=19)
    draws = []                                     # every K.random_normal draw, in call order (the test replays them)
    rng = np.random.default_rng(4242)

    def hook(shp, mean, std):
        e = rng.standard_normal(shp) * std + mean
        draws.append(e)
        return e
    K.set_random_normal_hook(hook)

    class _Bar:
        def __init__(self, **k): pass
        def update(self, *a): pass
    ns = {k: getattr(vd, k) for k in dir(vd) if not k.startswith("__")}           # the script's `from settings import *` names
    ns.update(np=np, vae_definition=vd, progressbar=types.SimpleNamespace(ProgressBar=_Bar), encoder=model.encoder, autoencoder=model.autoencoder,
              latent_dim=L, batch_size=bs, epochs=epochs, shuffle_train_set=False, load_previous_checkpoint=False, history=True, reset_states=True,
              train_set_size=n_songs, train_paths=[f"song{i}" for i in range(n_songs)],
              X_train=[s.dense(np.float64)[0] for s in songs], Y_train=[s.dense(np.float64)[0] for s in songs], C_train=[int(s.style[0]) for s in songs],
              I_train=[s.dense(np.float64)[1][0] for s in songs], V_train=[s.velocity.astype(np.float64) for s in songs],
              D_train=[np.zeros(s.velocity.shape) for s in songs], S_train=[np.zeros((len(s), 15)) for s in songs],
              normalized_S_train=[np.zeros((len(s), 15)) for s in songs], T_train=[120.0] * n_songs)
    for name in set(re.findall(r"\b(total_(?:train|test)_\w+_array)\b", code)):
        ns[name] = []
    with contextlib.redirect_stdout(io.StringIO()):
        exec(compile(code, "vae_training.py[train loop]", "exec"), ns)
    out = {"loss": np.array(ns["total_train_loss_array"]), "notes_acc": np.array(ns["total_train_accuracy_array"]),
           "notes_loss": np.array(ns["total_train_notes_loss_array"]), "instr_acc": np.array(ns["total_train_meta_instrument_accuracy_array"]),
           "instr_loss": np.array(ns["total_train_meta_instrument_loss_array"]), "vel_loss": np.array(ns["total_train_meta_velocity_loss_array"]),
           "style_acc": np.array(ns["total_train_composer_accuracy_array"]), "style_loss": np.array(ns["total_train_composer_loss_array"]),
           "kl": np.array(ns["total_train_kl_loss_array"]), "n_draws": np.array(len(draws)),
           "draws": np.concatenate([d.reshape(-1) for d in draws]), "draw_rows": np.array([d.shape[0] for d in draws]),
           "song_lengths": np.array([len(s) for s in songs])}
    dec_names = [nm for nm, _, _ in O.param_specs(ocfg) if nm.startswith(("dec_init/", "notes/", "meta_instrument/", "meta_velocity/"))]
    for nm, a in zip(dec_names, model.decoder.get_weights()):
        out["w/" + nm] = a
    for k, (sub, layer, idx) in keras_name_map(ocfg).items():
        out["w/" + k] = model.encoder.get_layer(layer).get_weights()[idx]
    path = os.path.join(HERE, "reference_training_loop.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes; epochs", epochs, "loss per epoch", np.round(out["loss"], 4), "draws", len(draws))


def run_style_transfer_loop(vd):
    """The reference's style-switch loop (vae_evaluation.py:2469-2483 and :2549-2550: switch the two style dimensions of every chunk's latent, decode
    it with the PREVIOUS SWITCHED latent as history, post-process) -- those source lines exec'd unmodified (the classifier / harmonicity
    bookkeeping between them left out) around the shim-built LSTM model.  -> reference_style_transfer.npz"""
    from keras import backend as K
    import recurrentshop.cells as rs_cells
    from midi_vae_b200 import synth
    from oracle import midivae_oracle as O
    from tests import util
    T, H, L, n = 16, 64, 16, 12
    lines = open(os.path.join(REF, "vae_evaluation.py")).read().split("\n")
    a = next(i for i, l in enumerate(lines) if l.strip() == "for i in range(len(encoded_representation)):")
    b = next(i for i in range(a, len(lines)) if lines[i].strip() == "D_list_switched.extend(D_switched)")
    c = next(i for i in range(b, len(lines)) if lines[i].strip() == "previous_switched_rep = switched_rep")
    body = lines[a:b + 1] + [lines[c]]
    assert any("switched_rep[C] = original_rep[C_switch]" in l for l in body) and any("prepare_decoder_input(switched_rep, C_switch, S[i], previous_switched_rep)" in l for l in body)
    indent = len(lines[a]) - len(lines[a].lstrip())
    code = "\n".join(l[indent:] for l in body)
    ecfg, ocfg = util.make_cfgs(T=T, H=H, L=L, feedback="as_wired", variant="standard", max_batch=n)
    w = util.make_weights(ecfg, seed=62, jitter=0.2)
    rs_cells.LSTM_VARIANT = "standard"
    K.clear_session()
    set_module_lengths(vd, T)
    K.set_random_normal_hook(None)
    model = vd.VAE()
    model.create(**create_kwargs(vd, cell_type="LSTM", input_length=T, output_length=T, lstm_size=H, latent_rep_size=L, meta_velocity_length=T,
                                 meta_held_notes_length=T, meta_next_notes_output_length=T, epsilon_std=0.0))     # vae_evaluation.py:482-485
    load_into_reference(model, ocfg, w)
    song = synth.make_song(np.random.default_rng(909), n, T, 0)
    X, I_, V3, _ = song.dense(np.float64)
    enc = model.encoder.predict(vd.prepare_encoder_input_list(X, I_[0], V3[..., 0], np.zeros((n, T))), batch_size=n)
    ns = dict(np=np, vae_definition=vd, decoder=model.decoder, batch_size=n, sample_method="argmax", encoded_representation=enc, C=0, C_switch=1,
              S=np.zeros((n, 15)), previous_switched_rep=np.zeros((1, L)), Y_list_switched=[], I_list_switched=[], V_list_switched=[], D_list_switched=[])
    exec(compile(code, "vae_evaluation.py[style switch loop]", "exec"), ns)
    path = os.path.join(HERE, "reference_style_transfer.npz")
    np.savez_compressed(path, pitch=song.pitch, instr=song.instr, velocity=song.velocity, encoded=enc, Y=np.asarray(ns["Y_list_switched"]).astype(np.uint8),
                        I=np.asarray(ns["I_list_switched"]).astype(np.uint8), V=np.asarray(ns["V_list_switched"]), D=np.asarray(ns["D_list_switched"]).astype(np.uint8))
    print(path, os.path.getsize(path), "bytes; sounding steps", int(np.asarray(ns["Y_list_switched"]).sum()), "of", n * T)


def run_test_loop(vd):
    """The reference's evaluation pass: the metric-name enumeration (vae_training.py:172-187) and the body of test() from its accumulators (:246)
    to the end of the per-song loop -- source lines exec'd unmodified -- over four synthetic songs.  -> reference_test_loop.npz"""
    import textwrap
    from keras import backend as K
    import recurrentshop.cells as rs_cells
    from midi_vae_b200 import synth
    from tests import util
    T, H, L, bs, n_songs = 16, 64, 16, 8, 4
    lines = open(os.path.join(REF, "vae_training.py")).read().split("\n")
    a = next(i for i, l in enumerate(lines) if l.startswith("enumerated_metric_names = []"))
    b = next(i for i in range(a, len(lines)) if lines[i].startswith("# initialize loss arrays"))
    c = next(i for i, l in enumerate(lines) if l.startswith("    total_test_loss = 0"))
    d = next(i for i in range(c, len(lines)) if lines[i].strip() == "bar.update(test_song_num+1)")
    code = "\n".join(lines[a:b]) + "\n" + textwrap.dedent("\n".join(lines[c:d + 1]))
    assert "autoencoder.evaluate(input_list, output_list" in code and "enumerated_metric_names.append" in code
    ecfg, ocfg = util.make_cfgs(T=T, H=H, L=L, feedback="as_wired", variant="standard", max_batch=bs)
    w = util.make_weights(ecfg, seed=72, jitter=0.15)
    rs_cells.LSTM_VARIANT = "standard"
    K.clear_session()
    set_module_lengths(vd, T)
    model = vd.VAE()
    model.create(**create_kwargs(vd, cell_type="LSTM", input_length=T, output_length=T, lstm_size=H, latent_rep_size=L, meta_velocity_length=T,
                                 meta_held_notes_length=T, meta_next_notes_output_length=T))
    load_into_reference(model, ocfg, w)
    songs = synth.make_songs(n_songs, T, seed=888, min_chunks=5, max_chunks=19)
    draws = []
    rng = np.random.default_rng(5151)

    def hook(shp, mean, std):
        e = rng.standard_normal(shp) * std + mean
        draws.append(e)
        return e
    K.set_random_normal_hook(hook)

    class _Bar:
        def __init__(self, **k): pass
        def update(self, *a): pass
    ns = {k: getattr(vd, k) for k in dir(vd) if not k.startswith("__")}
    ns.update(np=np, vae_definition=vd, progressbar=types.SimpleNamespace(ProgressBar=_Bar), encoder=model.encoder, autoencoder=model.autoencoder,
              latent_dim=L, batch_size=bs, history=True, reset_states=True, test_set_size=n_songs,
              X_test=[s.dense(np.float64)[0] for s in songs], Y_test=[s.dense(np.float64)[0] for s in songs], C_test=[int(s.style[0]) for s in songs],
              I_test=[s.dense(np.float64)[1][0] for s in songs], V_test=[s.velocity.astype(np.float64) for s in songs],
              D_test=[np.zeros(s.velocity.shape) for s in songs], normalized_S_test=[np.zeros((len(s), 15)) for s in songs], T_test=[120.0] * n_songs)
    exec(compile(code, "vae_training.py[test loop]", "exec"), ns)
    keys = ["total_test_loss", "total_test_notes_loss", "total_test_accuracy", "total_test_meta_instrument_loss", "total_test_meta_instrument_accuracy",
            "total_test_meta_velocity_loss", "total_test_meta_velocity_accuracy", "total_test_loss_composer", "total_test_accuracy_composer"]
    path = os.path.join(HERE, "reference_test_loop.npz")
    np.savez_compressed(path, totals=np.array([ns[k] for k in keys]), total_names=np.array(keys), enumerated=np.array(ns["enumerated_metric_names"]),
                        draws=np.concatenate([x.reshape(-1) for x in draws]), draw_rows=np.array([x.shape[0] for x in draws]),
                        song_lengths=np.array([len(s) for s in songs]))
    print(path, os.path.getsize(path), "bytes;", ns["enumerated_metric_names"])


def main():
    vd = import_reference()
    import keras
    from keras import backend as K
    import recurrentshop.cells as rs_cells
    from midi_vae_b200 import METRIC_KEYS, marshal, synth  # noqa: F401
    from oracle import midivae_oracle as O
    from tests import util

    out_layout = {}
    # ---- 1. the reference's DEFAULT configuration (GRU, T 64, H 256, L 256): names / shapes / save order
    K.clear_session()
    ref = vd.VAE()
    ref.create(**create_kwargs(vd))
    for part in ("encoder", "decoder", "autoencoder"):
        out_layout[part] = [[l, n, list(s)] for l, n, s in getattr(ref, part).weight_layout()]
    out_layout["metrics_names"] = ref.autoencoder.metrics_names
    out_layout["autoencoder_inputs"] = ref.autoencoder.input_names
    out_layout["decoder_inputs"] = ref.decoder.input_names
    out_layout["encoder_inputs"] = ref.encoder.input_names
    # the decoders' first-cell input kernels in every shipped checkpoint: still at their Glorot-uniform initialisation (max |w| = the
    # initialiser's limit, std = limit / sqrt(3)) after hundreds of epochs => their input was identically zero => decoder_feedback 'as_wired'
    from midi_vae_b200 import hdf5
    stats = {}
    for mdl in sorted(os.listdir(os.path.join(REF, "models"))):
        for f in sorted(os.listdir(os.path.join(REF, "models", mdl))):
            if f.startswith("decoderEpoch"):
                t = hdf5.read_weights(os.path.join(REF, "models", mdl, f))
                for ln in t["layer_names"]:
                    for wn, a in t["layers"][ln]:
                        if a.ndim == 2:
                            stats[f"{mdl}/{wn.replace(':0', '')}"] = [list(a.shape), float(np.abs(a).max()), float(a.std())]
    out_layout["shipped_decoder_kernel_stats"] = stats
    json.dump(out_layout, open(os.path.join(HERE, "reference_layout.json"), "w"), indent=0)

    # ---- 2. LSTM branch at cfg1 shapes, weights = the seeded test weights, run through the reference's own prepare_* + models
    T, H, L, n = 16, 64, 16, 8
    out = {}
    for variant, shim_variant in (("standard", "standard"), ("recurrentshop_recalled", "recalled")):
        ecfg, ocfg = util.make_cfgs(T=T, H=H, L=L, feedback="as_wired", variant=variant, max_batch=n)
        w = util.make_weights(ecfg, seed=42, jitter=0.1)
        # one synthetic SONG of n chunks: the reference's prepare_* take one style class and one instrument matrix per song
        r = synth.make_song(np.random.default_rng(1235), n, T, 1)
        rng = np.random.default_rng(1335)
        hist = (rng.standard_normal((n, L)) * 0.1).astype(np.float32)
        eps = synth.make_eps(n, L, 1235, 0.01)
        rs_cells.LSTM_VARIANT = shim_variant
        K.clear_session()
        set_module_lengths(vd, T)
        model = vd.VAE()
        model.create(**create_kwargs(vd, cell_type="LSTM", input_length=T, output_length=T, lstm_size=H, latent_rep_size=L,
                                     meta_velocity_length=T, meta_held_notes_length=T, meta_next_notes_output_length=T))
        load_into_reference(model, ocfg, w)
        K.set_random_normal_hook(lambda shp, mean, std: eps.astype(np.float64)[:shp[0]] * (std / 0.01) + mean)

        # the reference's per-song tensors (vae_definition.py:765-769): X/Y (N,T,61) one-hot, C int, I (4,16), V (N,T), D (N,T), S, H (N,L)
        X, I_, V3, C1 = r.dense(np.float64)
        Y = X
        I_song = I_[0]                                   # one instrument matrix per song (tiled by prepare_*)
        V = V3[..., 0]
        D = np.zeros_like(V)
        S = np.zeros((n, 15))
        C = int(r.style[0])
        in_list, out_list, sw = vd.prepare_autoencoder_input_and_output_list(X, Y, C, I_song, V, D, S, hist.astype(np.float64), return_sample_weight=True)
        enc_in = vd.prepare_encoder_input_list(X, I_song, V, D)
        z = model.encoder.predict(enc_in, batch_size=n)
        dec_in = vd.prepare_decoder_input(z, C, S, hist.astype(np.float64))
        dec_out = model.decoder.predict(dec_in, batch_size=n)
        ae_out = model.autoencoder.predict(in_list, batch_size=n)
        ev = model.autoencoder.evaluate(in_list, out_list, batch_size=n, sample_weight=sw, verbose=0)
        hist_fit = []
        for _ in range(3):
            h = model.autoencoder.fit(in_list, out_list, epochs=1, batch_size=n, shuffle=False, sample_weight=sw, verbose=0)
            hist_fit.append({k: v[0] for k, v in h.history.items()})
        g0 = None
        p = variant + "/"
        out[p + "z"] = z
        for k, a in zip(("Y", "I", "V"), dec_out):
            out[p + "dec_" + k] = a
        for k, a in zip(("Y", "I", "V", "C"), ae_out):
            out[p + "ae_" + k] = a
        out[p + "evaluate"] = np.array(ev)
        out[p + "metrics_names"] = np.array(model.autoencoder.metrics_names)
        out[p + "fit_keys"] = np.array(sorted(hist_fit[0]))
        out[p + "fit"] = np.array([[hf[k] for k in sorted(hf)] for hf in hist_fit])
        out[p + "in_shapes"] = np.array([str(np.asarray(a).shape) for a in in_list])
        out[p + "out_shapes"] = np.array([str(np.asarray(a).shape) for a in out_list])
        out[p + "sw_shapes"] = np.array([str(np.asarray(a).shape) for a in sw])
        if variant == "standard":                     # the lists themselves, as the reference's prepare_* built them
            for tag, lst in (("in", in_list), ("out", out_list), ("sw", sw), ("enc_in", enc_in), ("dec_in", dec_in)):
                for i, a in enumerate(lst):
                    out[f"lists/{tag}_{i}"] = np.asarray(a)
        # weights after the 3 Adam steps, by our names
        dec_names = [nm for nm, _, _ in O.param_specs(ocfg) if nm.startswith(("dec_init/", "notes/", "meta_instrument/", "meta_velocity/"))]
        for nm, a in zip(dec_names, model.decoder.get_weights()):
            out[p + "w3/" + nm] = a
        for k, (sub, layer, idx) in keras_name_map(ocfg).items():
            out[p + "w3/" + k] = model.encoder.get_layer(layer).get_weights()[idx]
        del g0
    out["pitch"], out["instr"], out["velocity"], out["style"], out["hist"], out["eps"] = r.pitch, r.instr, r.velocity, r.style, hist, eps
    path = os.path.join(HERE, "reference_cfg1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")
    run_variants(vd)
    run_postprocess(vd)
    run_training_loop(vd)
    run_style_transfer_loop(vd)
    run_test_loop(vd)
    print("keras shim", keras.__version__, "evaluate(standard):", dict(zip(out["standard/metrics_names"], out["standard/evaluate"])))


if __name__ == "__main__":
    main()
