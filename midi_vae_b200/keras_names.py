"""Layer / weight names Keras 2.0.8 gives the reference's three models (vae_definition.py:212-441), so that ``save_weights`` files written
here have the layout of the reference's own checkpoints (``models/*/{encoder,decoder,autoencoder}Epoch*.pickle`` -- HDF5 despite the suffix,
vae_training.py:966-978) and the shipped files load positionally (``load_weights(by_name=False)``).

Pinned by tests/golden/checkpoint_layout.json (read out of the 12 shipped files) for the reference-default GRU graph; the LSTM branch follows the
same construction order (vae_definition.py:519-645): per decoder, Keras' global ``dense_N`` counter runs over [cells of the stack (GRUCell: 3
inner Denses, LSTMCell: 2), output Dense, initial-state Denses (one per state)] for notes, then instrument, then velocity.

``layout(cfg, part)`` -> [(layer_name, [(keras_weight_name, engine_tensor_name), ...]), ...] for EVERY layer of the model in Keras' saved order;
weightless layers (inputs, Concatenate, Lambda, KL layer) carry an empty list.  Engine tensor names are those of
``engine.reference_param_specs``.
"""
from __future__ import annotations

from typing import List, Tuple

Layer = Tuple[str, List[Tuple[str, str]]]


def _encoder_layers(cfg) -> List[Layer]:
    pre = "gru" if cfg.cell_type == "GRU" else "lstm"

    def rnn(name):
        return (name, [(f"{name}/kernel", f"{name}/kernel"), (f"{name}/recurrent_kernel", f"{name}/recurrent_kernel"), (f"{name}/bias", f"{name}/bias")])

    def dense(name):
        return (name, [(f"{name}/kernel", f"{name}/kernel"), (f"{name}/bias", f"{name}/bias")])
    ne = cfg.num_layers_encoder
    out: List[Layer] = [("notes_input", []), rnn(f"{pre}_1"), ("meta_instrument_input", [])]
    if ne == 2:                       # the shipped order (depth-sorted): gru_1, meta_instrument_input, gru_2, ...
        out.append(rnn(f"{pre}_2"))
    else:
        out[2:2] = [rnn(f"{pre}_{k}") for k in range(2, ne + 1)]
    out += [rnn(f"{pre}_meta_instrument"), ("meta_velocity_input", []), ("concatenated_instrument_and_notes_layer", []), rnn(f"{pre}_meta_velocity"),
            ("concatenated_velocity_and_rest_layer", []), dense("extra_instrument_after_concat_layer")]
    if cfg.extra_layer:
        out.append(dense("extra_layer"))
    out += [("lambda_1", []), ("lambda_2", []), dense("z_mean"), dense("z_log_var"), ("kl_layer", [])]
    return out


def _decoder_weights(cfg):
    """-> (init-state Dense layers [(layer, weights)], recurrent models [(model name, weights)]) with Keras' dense_N numbering."""
    gru = cfg.cell_type == "GRU"
    cell = "gru_cell" if gru else "lstm_cell"
    per_state = 1 if gru else 2
    n = [0]
    c = [0]

    def dense_id():
        n[0] += 1
        return f"dense_{n[0]}"

    def rs_cell(engine_cell):
        c[0] += 1
        cn = f"{cell}_{c[0]}"
        d1 = dense_id()
        w = [(f"{cn}/{d1}/kernel", f"{engine_cell}/kernel"), (f"{cn}/{d1}/bias", f"{engine_cell}/bias")]
        if gru:                        # Dense(3H, bias) on x; Dense(2H) on h for z, r; Dense(H) on r*h
            w += [(f"{cn}/{dense_id()}/kernel", f"{engine_cell}/recurrent_kernel_1"), (f"{cn}/{dense_id()}/kernel", f"{engine_cell}/recurrent_kernel_2")]
        else:                          # Dense(4H, bias) on x; Dense(4H, no bias) on h
            w += [(f"{cn}/{dense_id()}/kernel", f"{engine_cell}/recurrent_kernel")]
        return w

    inits, models = [], []

    def branch(model, cells, out_name, init_names):
        w = []
        for ec in cells:
            w += rs_cell(ec)
        d = dense_id()
        w += [(f"{d}/kernel", f"{out_name}/kernel"), (f"{d}/bias", f"{out_name}/bias")]
        for nm in init_names:
            for j in range(1, per_state + 1):
                d = dense_id()
                inits.append((d, [(f"{d}/kernel", f"dec_init/{nm}_s{j}/kernel"), (f"{d}/bias", f"dec_init/{nm}_s{j}/bias")]))
        models.append((model, w))
    nd = cfg.num_layers_decoder
    branch("notes", [f"notes/cell_{k}" for k in range(1, nd + 1)], "notes/out", [f"notes_l{k}" for k in range(1, nd + 1)])
    branch("meta_instrument", ["meta_instrument/cell"], "meta_instrument/out", ["instr"])
    branch("meta_velocity", ["meta_velocity/cell"], "meta_velocity/out", ["vel"])
    return inits, models


def _decoder_layers(cfg) -> List[Layer]:
    inits, models = _decoder_weights(cfg)
    gru = cfg.cell_type == "GRU"
    per_state = 1 if gru else 2
    nd = cfg.num_layers_decoder
    out: List[Layer] = [("encoded_input", [])]
    if cfg.history:
        out += [("history_input", []), ("concatenate_1", [])]
    out.append(("input_decoder_start", []))
    out += inits[:nd * per_state]
    out.append(("input_decoder_meta_instrument_start", []))
    out += inits[nd * per_state:(nd + 1) * per_state]
    out.append(("input_decoder_meta_velocity_start", []))
    out += inits[(nd + 1) * per_state:]
    out += models
    return out


def layout(cfg, part: str) -> List[Layer]:
    if part == "encoder":
        return _encoder_layers(cfg) + [("lambda", [])]
    if part == "decoder":
        return _decoder_layers(cfg)
    if part == "autoencoder":
        inits, models = _decoder_weights(cfg)
        nested = [w for _, ws in inits for w in ws] + [w for _, ws in models for w in ws]       # the decoder Model as ONE layer: its weights in layer order
        tail: List[Layer] = [("input_decoder_start", []), ("lambda", [])]
        if cfg.history:
            tail.append(("history_input", []))
        tail += [("input_decoder_meta_instrument_start", []), ("input_decoder_meta_velocity_start", []), ("decoder", nested), ("composer_decoder", [])]
        return _encoder_layers(cfg) + tail
    raise ValueError(part)
