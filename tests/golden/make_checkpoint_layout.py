"""Extract the structural fingerprint (layer names, weight names, shapes, file order) of the reference's shipped
Keras-HDF5 checkpoints into tests/golden/checkpoint_layout.json.  Run where /root/reference exists:

    python tests/golden/make_checkpoint_layout.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from midi_vae_b200 import hdf5  # noqa: E402

REF = "/root/reference/models"
out = {}
for model in sorted(os.listdir(REF)):
    for f in sorted(os.listdir(os.path.join(REF, model))):
        if f.startswith(("encoderEpoch", "decoderEpoch", "autoencoderEpoch")):
            part = f.split("Epoch")[0]
            out[f"{model}/{part}"] = {"file": f"models/{model}/{f}", "layout": [[l, w, list(s)] for l, w, s in hdf5.layout(os.path.join(REF, model, f))]}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "checkpoint_layout.json")
json.dump(out, open(path, "w"), indent=0)
print(path, {k: len(v["layout"]) for k, v in out.items()})
