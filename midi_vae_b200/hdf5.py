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


# ======================================================================================================================
# Writer: the same subset of HDF5 that Keras 2.0.8 / h5py wrote into the reference's checkpoints (vae_training.py:966-978)
# ======================================================================================================================
# Structures mirror the shipped files byte layout by byte layout (dumped with the reader above): superblock v0 (leaf K 4, internal K 16),
# v1 object headers, old-style groups (one level-0 v1 B-tree node -> symbol-table nodes of <= 8 sorted entries, local heap with a trailing free
# block), contiguous little-endian float32 datasets with the dataspace / datatype / fill-value / layout / mtime messages h5py emits,
# version-1 attributes: `layer_names` / `weight_names` as null-padded fixed-length string arrays (float64 rank-1 extent-0 when empty),
# `backend` / `keras_version` as variable-length strings in one global heap collection.
_LEAF_K, _INTERNAL_K = 4, 16
_SNOD_SIZE = 8 + 2 * _LEAF_K * 40
_TREE_SIZE = 24 + 2 * _INTERNAL_K * 8 + (2 * _INTERNAL_K + 1) * 8


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


class _Writer:
    def __init__(self):
        self.b = bytearray(96)          # superblock, filled in last
        self.gcol_addr = None
        self.gcol_items: List[bytes] = []

    def alloc(self, n: int) -> int:
        off = _pad8(len(self.b))
        self.b.extend(b"\x00" * (off - len(self.b) + n))
        return off

    def put(self, off: int, data: bytes):
        self.b[off:off + len(data)] = data

    # ---- messages
    @staticmethod
    def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
        body = body + b"\x00" * (_pad8(len(body)) - len(body))
        if len(body) >= 65536:
            raise ValueError("object-header message too large for a version-1 header (an attribute with too many / too long names)")
        return struct.pack("<HHB3x", mtype, len(body), flags) + body

    def _header(self, msgs: List[bytes]) -> int:
        data = b"".join(msgs)
        off = self.alloc(16 + len(data))
        self.put(off, struct.pack("<BxHII4x", 1, len(msgs), 1, len(data)) + data)
        return off

    @staticmethod
    def _dataspace(shape) -> bytes:
        dims = b"".join(struct.pack("<Q", d) for d in shape)
        return struct.pack("<BBB5x", 1, len(shape), 1 if shape else 0) + dims + (dims if shape else b"")     # max dims = dims

    _F32 = bytes.fromhex("11201f000400000000002000170800177f000000")
    _F64 = bytes.fromhex("11203f000800000000004000340b0034ff030000")
    _VLEN_STR = bytes.fromhex("1901000010000000" "1000000001000000" "00000800")

    def _attr(self, name: str, dtype: bytes, shape, data: bytes) -> bytes:
        nm = name.encode() + b"\x00"
        sp = self._dataspace(shape)
        body = struct.pack("<BxHHH", 1, len(nm), len(dtype), len(sp))
        body += nm + b"\x00" * (_pad8(len(nm)) - len(nm)) + dtype + b"\x00" * (_pad8(len(dtype)) - len(dtype)) + sp + b"\x00" * (_pad8(len(sp)) - len(sp)) + data
        return self._msg(0x000C, body, 4)

    def attr_strings(self, name: str, values: List[str]) -> bytes:
        if not values:                         # h5py stores an empty list as a float64 array of extent 0
            return self._attr(name, self._F64, (0,), b"")
        w = max(len(v.encode()) for v in values)
        dtype = struct.pack("<BBBBI", 0x13, 0x01, 0, 0, w)          # string class, null-padded, ASCII, fixed width
        return self._attr(name, dtype, (len(values),), b"".join(v.encode().ljust(w, b"\x00") for v in values))

    def attr_vlen(self, name: str, value: str) -> bytes:
        self.gcol_items.append(value.encode())
        ref = struct.pack("<IQI", len(value.encode()), 0, len(self.gcol_items))     # collection address patched at the end
        return self._attr(name, self._VLEN_STR, (), ref)

    # ---- datasets / groups
    def dataset(self, arr: np.ndarray, mtime: int) -> int:
        a = np.ascontiguousarray(arr, dtype="<f4")
        data = self.alloc(a.nbytes)
        self.put(data, a.tobytes())
        msgs = [self._msg(0x0001, self._dataspace(a.shape)), self._msg(0x0003, self._F32, 1), self._msg(0x0005, bytes.fromhex("0202020100000000"), 1),
                self._msg(0x0008, struct.pack("<BBQQ", 3, 1, data, a.nbytes)), self._msg(0x0012, struct.pack("<B3xI", 1, mtime))]
        return self._header(msgs)

    def group(self, children: Dict[str, Tuple[int, Tuple[int, int]]], attrs: List[bytes]) -> Tuple[int, Tuple[int, int]]:
        """children: name -> (object header address, (btree, heap) if the child is a group else None).  Returns (header address, (btree, heap))."""
        names = sorted(children, key=lambda s: s.encode())
        if len(names) > 2 * _INTERNAL_K * 2 * _LEAF_K:
            raise ValueError("too many entries for a single-node group B-tree")
        # local heap: 8 zero bytes (the empty name, B-tree key 0), the names (8-byte padded), a trailing free block
        seg = bytearray(8)
        name_off = {}
        for nm in names:
            name_off[nm] = len(seg)
            e = nm.encode() + b"\x00"
            seg += e + b"\x00" * (_pad8(len(e)) - len(e))
        free_off = len(seg)
        seg += struct.pack("<QQ", 1, 32) + b"\x00" * 16                 # free block: next = H5HL_FREE_NULL (1), size 32
        seg_addr = self.alloc(len(seg)); self.put(seg_addr, bytes(seg))
        heap = self.alloc(32)
        self.put(heap, b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), free_off, seg_addr))
        # symbol-table nodes of up to 2 * leaf K entries, in name order
        snods = []
        for i in range(0, len(names), 2 * _LEAF_K):
            chunk = names[i:i + 2 * _LEAF_K]
            node = bytearray(b"SNOD" + struct.pack("<BxH", 1, len(chunk)))
            for nm in chunk:
                addr, sub = children[nm]
                node += struct.pack("<QQII", name_off[nm], addr, 1 if sub else 0, 0) + (struct.pack("<QQ", *sub) if sub else b"\x00" * 16)
            node += b"\x00" * (_SNOD_SIZE - len(node))
            a = self.alloc(_SNOD_SIZE); self.put(a, bytes(node))
            snods.append((a, name_off[chunk[-1]]))
        tree = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF) + struct.pack("<Q", 0))
        for a, last_key in snods:
            tree += struct.pack("<QQ", a, last_key)
        tree += b"\x00" * (_TREE_SIZE - len(tree))
        bt = self.alloc(_TREE_SIZE); self.put(bt, bytes(tree))
        hdr = self._header([self._msg(0x0011, struct.pack("<QQ", bt, heap))] + attrs)
        return hdr, (bt, heap)

    def finish(self, root_hdr: int, root_sub: Tuple[int, int]) -> bytes:
        if self.gcol_items:
            size = 4096
            col = bytearray(b"GCOL" + struct.pack("<B3xQ", 1, size))
            for i, item in enumerate(self.gcol_items, 1):
                col += struct.pack("<HH4xQ", i, 0, len(item)) + item + b"\x00" * (_pad8(len(item)) - len(item))
            col += struct.pack("<HH4xQ", 0, 0, size - len(col))
            col += b"\x00" * (size - len(col))
            self.gcol_addr = self.alloc(size); self.put(self.gcol_addr, bytes(col))
            # patch the collection address into every variable-length reference (sequence length, ADDRESS, index)
            marker = self._VLEN_STR
            pos = 0
            while True:
                pos = self.b.find(marker, pos)
                if pos < 0:
                    break
                ref = pos + _pad8(len(marker)) + 8          # datatype (padded) + scalar dataspace (8 bytes) -> the 16-byte reference
                self.put(ref + 4, struct.pack("<Q", self.gcol_addr))
                pos = ref
        eof = _pad8(len(self.b))
        self.b.extend(b"\x00" * (eof - len(self.b)))
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _LEAF_K, _INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", *root_sub)
        assert len(sb) == 96
        self.put(0, sb)
        return bytes(self.b)


def write_weights(path: str, layers: List[Tuple[str, List[Tuple[str, np.ndarray]]]], backend: str = "theano", keras_version: str = "2.0.8",
                  mtime: int = 0) -> None:
    """Keras 2.0.8 ``save_weights`` layout: root attrs layer_names / backend / keras_version; one group per layer with attr weight_names
    and one dataset per weight at the (possibly nested) path its name spells, e.g. ``gru_1/gru_1/kernel``, ``decoder/gru_cell_1/dense_1/kernel``.
    ``layers`` lists EVERY layer in model order; weightless layers carry an empty list."""
    w = _Writer()

    def build(tree: dict, attrs: List[bytes]):
        children = {}
        for nm, node in tree.items():
            if isinstance(node, dict):
                children[nm] = build(node, [])
            else:
                children[nm] = (w.dataset(node, mtime), None)
        return w.group(children, attrs)

    root_children = {}
    for layer, tensors in layers:
        tree: dict = {}
        for wn, arr in tensors:
            parts = wn.split("/")
            d = tree
            for p in parts[:-1]:
                d = d.setdefault(p, {})
                if not isinstance(d, dict):
                    raise ValueError(f"weight name {wn!r} nests under a dataset")
            d[parts[-1]] = np.asarray(arr)
        root_children[layer] = build(tree, [w.attr_strings("weight_names", [wn for wn, _ in tensors])])
    root_hdr, root_sub = w.group(root_children, [w.attr_strings("layer_names", [l for l, _ in layers]), w.attr_vlen("backend", backend),
                                                 w.attr_vlen("keras_version", keras_version)])
    with open(path, "wb") as f:
        f.write(w.finish(root_hdr, root_sub))
