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
