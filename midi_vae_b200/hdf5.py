"""Minimal pure-Python reader for the Keras-2.0.8 HDF5 weight files the reference ships and writes
(``models/*/{encoder,decoder,autoencoder}Epoch*.pickle`` -- HDF5 despite the suffix; vae_training.py:966-978).

h5py is not available here, and the files only use the oldest, simplest HDF5 structures:
superblock version 0, version-1 object headers, groups as v1 B-trees + local heaps + symbol-table nodes, contiguous
little-endian float32 datasets, version-1 attributes.  Exactly those are parsed; anything else raises.

    tree = read_weights("models/JvP/encoderEpoch440.pickle")
    tree["layer_names"]                      -> ['notes_input', 'gru_1', ...]
    tree["layers"]["gru_1"]                  -> [("gru_1/kernel:0", ndarray (61, 768)), ...]   (weight_names order)
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class _File:
    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.b = f.read()
        if self.b[:8] != SIGNATURE:
            raise ValueError(f"{path}: not an HDF5 file")
        if self.b[8] != 0:
            raise NotImplementedError(f"{path}: superblock version {self.b[8]} (only 0 is handled)")
        if self.b[13] != 8 or self.b[14] != 8:
            raise NotImplementedError("only 8-byte offsets/lengths are handled")
        # superblock v0: ... base(8) freespace(8) eof(8) driver(8) then the root symbol-table entry
        root_entry = 24 + 32
        self.root_header = self.u64(root_entry + 8)

    def u8(self, o): return self.b[o]
    def u16(self, o): return struct.unpack_from("<H", self.b, o)[0]
    def u32(self, o): return struct.unpack_from("<I", self.b, o)[0]
    def u64(self, o): return struct.unpack_from("<Q", self.b, o)[0]

    # ---- object header (version 1): yields (type, offset, size) of every message, following continuations
    def messages(self, addr: int):
        if self.u8(addr) != 1:
            raise NotImplementedError(f"object header version {self.u8(addr)} at {addr}")
        nmsg = self.u16(addr + 2)
        size = self.u32(addr + 8)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            o, remaining = blocks.pop(0)
            end = o + remaining
            while o + 8 <= end and len(out) < nmsg:
                mtype, msize = self.u16(o), self.u16(o + 2)
                body = o + 8
                if mtype == 0x10:      # continuation
                    blocks.append((self.u64(body), self.u64(body + 8)))
                out.append((mtype, body, msize))
                o = body + msize
        return out

    # ---- groups
    def group_entries(self, header_addr: int) -> Dict[str, int]:
        """name -> object-header address for an old-style group (symbol-table message 0x11)."""
        for mtype, body, _ in self.messages(header_addr):
            if mtype == 0x11:
                btree, heap = self.u64(body), self.u64(body + 8)
                return self._walk_btree(btree, self._heap_data(heap))
        return {}

    def _heap_data(self, heap_addr: int) -> int:
        if self.b[heap_addr:heap_addr + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        return self.u64(heap_addr + 24)

    def _name(self, heap_data: int, off: int) -> str:
        end = self.b.index(b"\x00", heap_data + off)
        return self.b[heap_data + off:end].decode()

    def _walk_btree(self, addr: int, heap_data: int) -> Dict[str, int]:
        out: Dict[str, int] = {}
        if self.b[addr:addr + 4] != b"TREE":
            raise ValueError("bad B-tree signature")
        level, used = self.u8(addr + 5), self.u16(addr + 6)
        o = addr + 24                      # after signature/type/level/entries/left/right siblings
        for i in range(used):
            child = self.u64(o + 8)        # key(8) child(8) key(8) ...
            if level > 0:
                out.update(self._walk_btree(child, heap_data))
            else:
                out.update(self._snod(child, heap_data))
            o += 16
        return out

    def _snod(self, addr: int, heap_data: int) -> Dict[str, int]:
        if self.b[addr:addr + 4] != b"SNOD":
            raise ValueError("bad symbol-table node signature")
        n = self.u16(addr + 6)
        out = {}
        o = addr + 8
        for _ in range(n):
            out[self._name(heap_data, self.u64(o))] = self.u64(o + 8)
            o += 40
        return out

    # ---- datasets
    def dataset(self, header_addr: int) -> np.ndarray:
        shape, dtype, data_addr, data_size = None, None, None, None
        for mtype, body, _ in self.messages(header_addr):
            if mtype == 0x01:
                shape = self._dataspace(body)
            elif mtype == 0x03:
                dtype = self._datatype(body)
            elif mtype == 0x08:
                ver, cls = self.u8(body), self.u8(body + 1)
                if ver != 3 or cls != 1:
                    raise NotImplementedError(f"data layout version {ver} class {cls} (only contiguous v3)")
                data_addr, data_size = self.u64(body + 2), self.u64(body + 10)
        if shape is None or dtype is None or data_addr is None:
            raise ValueError("incomplete dataset header")
        count = int(np.prod(shape)) if shape else 1
        if data_addr == UNDEF:
            return np.zeros(shape, dtype)
        return np.frombuffer(self.b, dtype=dtype, count=count, offset=data_addr).reshape(shape).copy()

    def _dataspace(self, body: int) -> Tuple[int, ...]:
        ver, rank, flags = self.u8(body), self.u8(body + 1), self.u8(body + 2)
        o = body + (8 if ver == 1 else 4)
        return tuple(self.u64(o + 8 * i) for i in range(rank))

    def _datatype(self, body: int):
        cls = self.u8(body) & 0x0F
        size = self.u32(body + 4)
        if cls == 1 and size in (4, 8):
            return np.dtype("<f4") if size == 4 else np.dtype("<f8")
        if cls == 0 and size in (1, 2, 4, 8):
            return np.dtype(f"<i{size}")
        if cls == 3:
            return np.dtype(f"S{size}")
        raise NotImplementedError(f"datatype class {cls} size {size}")

    # ---- attributes (version 1): name -> ndarray / bytes
    def attributes(self, header_addr: int) -> Dict[str, object]:
        out = {}
        for mtype, body, _ in self.messages(header_addr):
            if mtype != 0x0C:
                continue
            if self.u8(body) != 1:
                raise NotImplementedError("attribute message version")
            nsz, tsz, ssz = self.u16(body + 2), self.u16(body + 4), self.u16(body + 6)
            pad = lambda x: (x + 7) // 8 * 8
            o = body + 8
            name = self.b[o:o + nsz].split(b"\x00")[0].decode()
            o += pad(nsz)
            tbody = o
            o += pad(tsz)
            shape = self._dataspace(o)
            o += pad(ssz)
            cls = self.u8(tbody) & 0x0F
            if cls == 9:        # variable-length string: stored in the global heap -- not needed for weight files written by Keras 2.0.8/Theano
                out[name] = None
                continue
            dtype = self._datatype(tbody)
            count = int(np.prod(shape)) if shape else 1
            arr = np.frombuffer(self.b, dtype=dtype, count=count, offset=o)
            out[name] = arr.reshape(shape).copy() if shape else arr[0]
        return out


def _strs(a) -> List[str]:
    if a is None:
        return []
    a = np.atleast_1d(a)
    return [x.decode() if isinstance(x, bytes) else str(x) for x in a.tolist()]


def read_weights(path: str) -> Dict[str, object]:
    """Keras ``save_weights`` file -> {'layer_names': [...], 'backend': ..., 'keras_version': ..., 'layers': {layer: [(weight_name, array), ...]}}.
    Nested models (the decoder inside the autoencoder file) appear as one layer whose weight names carry the inner path."""
    f = _File(path)
    root = f.root_header
    attrs = f.attributes(root)
    entries = f.group_entries(root)
    if "model_weights" in entries:                 # full-model file (model.save): weights live one level down
        root = entries["model_weights"]
        attrs = f.attributes(root)
        entries = f.group_entries(root)
    out = {"layer_names": _strs(attrs.get("layer_names")), "backend": attrs.get("backend"), "keras_version": attrs.get("keras_version"), "layers": {}}
    for layer in out["layer_names"]:
        g = entries[layer]
        names = _strs(f.attributes(g).get("weight_names"))
        tensors = []
        for wn in names:
            node = g
            for part in wn.split("/"):
                node = f.group_entries(node)[part]
            tensors.append((wn, f.dataset(node)))
        out["layers"][layer] = tensors
    return out


def layout(path: str) -> List[Tuple[str, str, Tuple[int, ...]]]:
    """[(layer, weight_name, shape), ...] in file order -- the structural fingerprint the tests pin."""
    t = read_weights(path)
    return [(layer, wn, tuple(arr.shape)) for layer in t["layer_names"] for wn, arr in t["layers"][layer]]


def load_keras_weights(model, path: str) -> None:
    """``load_weights`` on a Keras HDF5 file.  The shipped checkpoints are GRU models (settings.py:155); this build
    implements the LSTM branch, so their tensors cannot be loaded into it -- say so instead of guessing a mapping."""
    t = read_weights(path)
    shapes = [arr.shape for layer in t["layer_names"] for _, arr in t["layers"][layer]]
    want = [tuple(s) for _, s in model._specs()]
    if [tuple(s) for s in shapes] != want:
        raise NotImplementedError(
            f"{path}: Keras HDF5 weights with {len(shapes)} tensors (first shapes {shapes[:3]}) do not match this LSTM model "
            f"({len(want)} tensors, first shapes {want[:3]}); the shipped checkpoints are GRU models -- GRU cells are row f-1 (SURVEY.md 8(f))")
    flat = [arr for layer in t["layer_names"] for _, arr in t["layers"][layer]]
    model.set_weights(flat)
