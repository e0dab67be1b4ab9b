"""What IS pinned against the reference: the weight inventory.  The shipped checkpoints (models/*/*.pickle, Keras 2.0.8
HDF5) fix layer names, tensor shapes and order; they are GRU models with H = L = 256, so the comparison maps
gru -> lstm and 3 gate blocks -> 4 (a recurrentshop GRUCell stores its recurrent weights as (H,2H)+(H,H), an LSTMCell as
one (H,4H)).  The fingerprint is committed as tests/golden/checkpoint_layout.json (made by make_checkpoint_layout.py)."""
import json
import os

import pytest

from midi_vae_b200 import EngineConfig, hdf5, reference_param_specs

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = json.load(open(os.path.join(HERE, "golden", "checkpoint_layout.json")))
H = 256


def _ours():
    cfg = EngineConfig(input_length=64, lstm_size=H, latent_rep_size=256)      # models/*/params.txt: lstm_size 256, latent 256
    return reference_param_specs(cfg)


def test_all_four_shipped_models_share_one_layout():
    keys = sorted(FIX)
    assert len(keys) == 12
    for part in ("encoder", "decoder", "autoencoder"):
        layouts = [FIX[k]["layout"] for k in keys if k.endswith("/" + part)]
        assert all(l == layouts[0] for l in layouts)


def test_encoder_inventory_matches_the_checkpoints():
    ref = FIX["JvP/encoder"]["layout"]
    ours = [(n, s) for n, s in _ours() if not n.startswith(("dec_init", "notes/", "meta_"))]
    assert len(ref) == len(ours) == 20
    for (layer, wname, shape), (name, oshape) in zip(ref, ours):
        assert wname.replace("gru", "lstm") == name, (wname, name)
        expect = [4 * H if (d == 3 * H and layer.startswith("gru")) else d for d in shape]   # three GRU gate blocks -> four LSTM blocks
        assert list(oshape) == expect, (name, oshape, shape)


def test_decoder_inventory_matches_the_checkpoints():
    ref = FIX["JvP/decoder"]["layout"]
    ours = [(n, s) for n, s in _ours() if n.startswith(("dec_init", "notes/", "meta_"))]
    # initial-state Denses: GRU has one state per cell (dense_8, dense_9, dense_14, dense_19), LSTM two
    ref_init = [r for r in ref if r[0].startswith("dense_")]
    ours_init = [o for o in ours if o[0].startswith("dec_init")]
    assert len(ref_init) == 8 and len(ours_init) == 16
    assert all(tuple(r[2]) in ((512, 256), (256,)) for r in ref_init) and all(tuple(o[1]) in ((512, 256), (256,)) for o in ours_init)
    # recurrent models: notes (2 cells + Dense(61)), meta_instrument (cell + Dense(16)), meta_velocity (cell + Dense(1)), in this order
    ref_rest = [r for r in ref if not r[0].startswith("dense_")]
    groups = []
    for layer, wname, shape in ref_rest:
        if not groups or groups[-1][0] != layer:
            groups.append((layer, []))
        groups[-1][1].append((wname, tuple(shape)))
    assert [g[0] for g in groups] == ["notes", "meta_instrument", "meta_velocity"]
    ours_rest = [o for o in ours if not o[0].startswith("dec_init")]
    it = iter(ours_rest)
    for layer, tensors in groups:
        i = 0
        while i < len(tensors):
            wname, shape = tensors[i]
            if "gru_cell" in wname:      # kernel (D,3H), bias (3H), recurrent (H,2H) + (H,H)
                (kn, ks), (bn, bs), (rn, rs) = next(it), next(it), next(it)
                assert kn.endswith("/kernel") and ks == (shape[0], 4 * H)
                assert bn.endswith("/bias") and bs == (4 * H,) and tensors[i + 1][1] == (3 * H,)
                assert rn.endswith("/recurrent_kernel") and rs == (H, 4 * H) and tensors[i + 2][1] == (H, 2 * H) and tensors[i + 3][1] == (H, H)
                i += 4
            else:                        # output Dense kernel + bias: identical shapes
                (kn, ks), (bn, bs) = next(it), next(it)
                assert ks == shape and bs == tensors[i + 1][1], (kn, ks, shape)
                i += 2
    assert next(it, None) is None


def test_autoencoder_file_is_encoder_then_nested_decoder():
    ae, enc, dec = (FIX[f"JvP/{p}"]["layout"] for p in ("autoencoder", "encoder", "decoder"))
    assert [tuple(x[1:]) for x in ae[:len(enc)]] == [tuple(x[1:]) for x in enc]
    assert [x[1] for x in ae[len(enc):]] == [x[1] for x in dec] and all(x[0] == "decoder" for x in ae[len(enc):])


@pytest.mark.skipif(not os.path.exists("/root/reference/models/JvP/encoderEpoch440.pickle"), reason="reference checkout not present")
def test_hdf5_reader_reproduces_the_fixture():
    for key, entry in FIX.items():
        got = [[l, w, list(s)] for l, w, s in hdf5.layout(os.path.join("/root/reference", entry["file"]))]
        assert got == entry["layout"], key
    t = hdf5.read_weights("/root/reference/models/JvP/decoderEpoch440.pickle")
    k = dict(t["layers"]["notes"])["gru_cell_1/dense_1/kernel"]
    # the first decoder cell's input kernel still sits at its Glorot bound sqrt(6/(61+768))
    assert abs(abs(k).max() - (6.0 / (61 + 768)) ** 0.5) < 2e-3


# ---------------------------------------------------------------------------------------------------------------- round 2: GRU branch + writer
def _gru_cfg():
    return EngineConfig(input_length=64, lstm_size=H, latent_rep_size=256, cell_type="GRU")      # settings.py:108-112,155


@pytest.mark.parametrize("part", ["encoder", "decoder", "autoencoder"])
def test_gru_keras_names_equal_the_shipped_checkpoints(part):
    """keras_names.layout at the reference's default GRU settings reproduces layer names, weight names, shapes and order of the shipped files
    EXACTLY (no gru -> lstm mapping): what save_weights writes is what the reference's load_weights(by_name=False) expects."""
    from midi_vae_b200 import keras_names
    cfg = _gru_cfg()
    shapes = dict(reference_param_specs(cfg))
    ours = [[layer, kn, list(shapes[en])] for layer, ws in keras_names.layout(cfg, part) for kn, en in ws]
    assert ours == FIX[f"JvP/{part}"]["layout"]
    # every engine tensor of the part is written exactly once
    names = [en for _, ws in keras_names.layout(cfg, part) for _, en in ws]
    assert len(names) == len(set(names))
    if part == "autoencoder":
        assert sorted(names) == sorted(shapes)


@pytest.mark.skipif(not os.path.exists("/root/reference/models/JvP/encoderEpoch440.pickle"), reason="reference checkout not present")
@pytest.mark.parametrize("part", ["encoder", "decoder", "autoencoder"])
def test_gru_layer_names_including_weightless_layers(part):
    import glob
    from midi_vae_b200 import keras_names
    t = hdf5.read_weights(glob.glob(f"/root/reference/models/JvP/{part}Epoch*.pickle")[0])
    assert [l for l, _ in keras_names.layout(_gru_cfg(), part)] == t["layer_names"]


def test_lstm_keras_names_follow_the_same_construction_order():
    from midi_vae_b200 import keras_names
    cfg = EngineConfig(input_length=64, lstm_size=H, latent_rep_size=256)
    lay = dict(keras_names.layout(cfg, "decoder"))
    assert [k for k, _ in lay["notes"]] == ["lstm_cell_1/dense_1/kernel", "lstm_cell_1/dense_1/bias", "lstm_cell_1/dense_2/kernel",
                                           "lstm_cell_2/dense_3/kernel", "lstm_cell_2/dense_3/bias", "lstm_cell_2/dense_4/kernel", "dense_5/kernel", "dense_5/bias"]
    assert [l for l in lay if l.startswith("dense_")] == ["dense_6", "dense_7", "dense_8", "dense_9", "dense_13", "dense_14", "dense_18", "dense_19"]
    assert [k for k, _ in lay["meta_velocity"]] == ["lstm_cell_4/dense_15/kernel", "lstm_cell_4/dense_15/bias", "lstm_cell_4/dense_16/kernel", "dense_17/kernel", "dense_17/bias"]
    names = [en for _, ws in keras_names.layout(cfg, "autoencoder") for _, en in ws]
    assert sorted(names) == sorted(n for n, _ in reference_param_specs(cfg))


@pytest.mark.parametrize("cell_type", ["LSTM", "GRU"])
def test_engine_and_oracle_inventories_agree(cell_type):
    from oracle import midivae_oracle as O
    e = reference_param_specs(EngineConfig(input_length=16, lstm_size=64, latent_rep_size=16, cell_type=cell_type))
    o = [(n, tuple(s)) for n, s, _ in O.param_specs(O.OracleConfig(input_length=16, lstm_size=64, latent_rep_size=16, cell_type=cell_type))]
    assert [(n, tuple(s)) for n, s in e] == o


def test_hdf5_writer_round_trip(tmp_path):
    """write_weights -> read_weights: names, order, nesting, shapes and bytes survive; the file starts with the same superblock fields as the shipped ones."""
    import numpy as np
    rng = np.random.default_rng(0)
    layers = [("notes_input", []), ("gru_1", [("gru_1/kernel", rng.standard_normal((61, 24)).astype(np.float32)), ("gru_1/bias", rng.standard_normal(24).astype(np.float32))]),
              ("lambda_1", [])] + [(f"dense_{i}", [(f"dense_{i}/kernel", rng.standard_normal((3, i + 1)).astype(np.float32))]) for i in range(1, 12)] + \
             [("decoder", [("gru_cell_1/dense_1/kernel", rng.standard_normal((5, 7)).astype(np.float32)), ("gru_cell_1/dense_2/kernel", rng.standard_normal((2, 2)).astype(np.float32)),
                           ("dense_7/bias", rng.standard_normal(1).astype(np.float32))])]
    path = str(tmp_path / "w.pickle")
    hdf5.write_weights(path, layers)
    raw = open(path, "rb").read()
    assert raw[:8] == hdf5.SIGNATURE and raw[8] == 0 and raw[13] == 8 and raw[14] == 8 and raw[16:20] == bytes([4, 0, 16, 0])      # superblock v0, leaf K 4, internal K 16
    t = hdf5.read_weights(path)
    assert t["layer_names"] == [l for l, _ in layers]
    for layer, tensors in layers:
        got = t["layers"][layer]
        assert [n for n, _ in got] == [n for n, _ in tensors]
        for (_, a), (_, b) in zip(tensors, got):
            assert a.dtype == b.dtype and np.array_equal(a, b)


@pytest.mark.skipif(not os.path.exists("/root/reference/models/JvP/autoencoderEpoch440.pickle"), reason="reference checkout not present")
def test_hdf5_writer_reproduces_a_shipped_file_structure(tmp_path):
    """Re-writing a shipped checkpoint yields a file the reader parses to the identical tree (all 50 tensors bit-equal, 23 layer names)."""
    import numpy as np
    src = "/root/reference/models/JvP/autoencoderEpoch440.pickle"
    t = hdf5.read_weights(src)
    path = str(tmp_path / "copy.pickle")
    hdf5.write_weights(path, [(l, t["layers"][l]) for l in t["layer_names"]])
    t2 = hdf5.read_weights(path)
    assert t2["layer_names"] == t["layer_names"] and hdf5.layout(path) == hdf5.layout(src)
    for l in t["layer_names"]:
        for (_, a), (_, b) in zip(t["layers"][l], t2["layers"][l]):
            assert np.array_equal(a, b)
