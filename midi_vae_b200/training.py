"""Per-song training / evaluation driver: the loop of vae_training.py:728-964 around the hot path.

Per epoch: visit every song; build its history latents (zeros in epoch 0, otherwise ``encoder.predict`` of the song
rolled by one chunk, vae_training.py:788-798); ``autoencoder.fit(epochs=1, batch_size, shuffle=False)`` on the song's
chunks (mini-batches are consecutive slices); average the per-song history values over songs; recover the KL term as
the script does, ``(loss - sum_i w_i * loss_i) / beta`` (vae_training.py:946-957).  Songs are PACKED rolls
(synth.Rolls); the per-song encoder pass and the train steps go straight to the engine, no dense one-hot tensors.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from .engine import METRIC_KEYS
from .marshal import shift_history
from .synth import Rolls, concat

_KEYS = METRIC_KEYS[:9]


def _song_history(vae, song: Rolls, epoch: int, use_history: bool) -> np.ndarray:
    n, L = len(song), vae.engine.cfg.latent_rep_size
    if not use_history or epoch == 0:                    # vae_training.py:789-790: encoder untrained in epoch 0
        return np.zeros((n, L), np.float32)
    zs = []
    step = vae.max_batch
    for a in range(0, n, step):
        b = min(n, a + step)
        zs.append(vae.engine.encode(song.pitch[a:b], song.instr[a:b], song.velocity[a:b], vae._eps(b - a))[0])
    return shift_history(np.concatenate(zs)).astype(np.float32)     # H[0] = 0, H[i] = z[i-1]  (:791-798)


def _run_song(vae, song: Rolls, H: np.ndarray, batch_size: int, train: bool, silent_weight: float = 1.0) -> np.ndarray:
    """autoencoder.fit(epochs=1) / evaluate on one song: batch-size-weighted means over consecutive mini-batches."""
    n = len(song)
    step = min(batch_size, vae.max_batch)
    w = None
    if silent_weight != 1.0:                               # vae_definition.py:930-933
        w = np.ones(song.pitch.shape, np.float32)
        w[song.pitch == vae.engine.cfg.input_dim - 1] = silent_weight
    fn = vae.engine.train_on_batch if train else vae.engine.evaluate_batch
    tot = np.zeros(len(_KEYS))
    for a in range(0, n, step):
        b = min(n, a + step)
        m = fn(song.pitch[a:b], song.instr[a:b], song.velocity[a:b], song.style[a:b], H[a:b], vae._eps(b - a),
               None if w is None else w[a:b])
        tot += np.array([m[k] for k in _KEYS]) * (b - a)
    return tot / n


def _summarise(per_song: np.ndarray, vae) -> Dict[str, float]:
    cfg = vae.engine.cfg
    out = {k: float(v) for k, v in zip(_KEYS, per_song.mean(0))}          # mean over songs of per-song means (:817,867-868)
    kl = out["loss"] - cfg.notes_weight * out["decoder_loss_1"] - cfg.composer_weight * out["composer_decoder_loss"] \
        - cfg.meta_instrument_weight * out["decoder_loss_2"] - cfg.meta_velocity_weight * out["decoder_loss_3"]
    out["kl_loss"] = kl / cfg.beta                                           # :946-957
    return out


def pack_songs(songs: Sequence[Rolls], batch_size: int) -> List[Rolls]:
    """Greedy packing of whole songs, in order, into packs of at least ``batch_size`` chunks (the last pack takes what is left).  Every pack is
    one Rolls with ``song_start`` flags, so that the history shift restarts at each song (H = 0 on a song's first chunk)."""
    packs, cur, have = [], [], 0
    for s in songs:
        ss = np.zeros(len(s), np.uint8)
        if len(s):
            ss[0] = 1
        cur.append(Rolls(s.pitch, s.instr, s.velocity, s.style, ss)); have += len(s)
        if have >= batch_size:
            packs.append(concat(cur)); cur, have = [], 0
    if cur:
        packs.append(concat(cur))
    return packs


def train_epoch_packed(vae, songs: Sequence[Rolls], epoch: int, batch_size: int = 256, history: bool = True, silent_weight: float = 1.0,
                       fused_history: bool = False) -> Dict[str, float]:
    """OPT-IN variant of train_epoch for real data: songs are 10-40 chunks long, so the reference's one-fit-per-song loop runs
    the GPU at B = 10..40.  Here several whole songs share a mini-batch: per pack, ONE batched encoder pass builds every song's history latents
    (shifted inside each song), then consecutive mini-batches of ``batch_size`` chunks are trained.

    This is NOT the reference's arithmetic, on purpose, and it says so: (i) Adam takes one step per ``batch_size`` chunks instead of one per song
    remainder, (ii) the histories of all songs of a pack come from the weights at the start of the pack (the reference re-encodes before every
    song), (iii) the returned values are chunk-weighted means over the epoch, not means over songs of per-song means.  train_epoch() is the
    reference-faithful loop.  ``fused_history=True`` goes one step further: no encoder pass at all, each step takes its history from its own z
    on the device (same weights and same epsilon draw as the step; a song that continues into the next mini-batch carries its last z over)."""
    if fused_history and history and epoch > 0:
        # no encoder pass at all: every step builds its history from its own z on the device (Engine.set_history_mode; song boundaries per step)
        vae.engine.set_history_mode(True)
        try:
            tot, seen = np.zeros(len(_KEYS)), 0
            step = min(batch_size, vae.max_batch)
            for pack in pack_songs(songs, batch_size):
                n = len(pack)
                for a in range(0, n, step):
                    b = min(n, a + step)
                    ss = pack.song_start[a:b].copy()
                    if a == 0:
                        ss[0] = 1
                    m = vae.engine.train_on_batch(pack.pitch[a:b], pack.instr[a:b], pack.velocity[a:b], pack.style[a:b], None, vae._eps(b - a), None, None, ss)
                    tot += np.array([m[k] for k in _KEYS]) * (b - a); seen += b - a
            return _summarise((tot / max(seen, 1))[None, :], vae)
        finally:
            vae.engine.set_history_mode(False)
    tot, seen = np.zeros(len(_KEYS)), 0
    for pack in pack_songs(songs, batch_size):
        n = len(pack)
        if not history or epoch == 0:
            H = np.zeros((n, vae.engine.cfg.latent_rep_size), np.float32)
        else:
            zs = [vae.engine.encode(pack.pitch[a:a + vae.max_batch], pack.instr[a:a + vae.max_batch], pack.velocity[a:a + vae.max_batch],
                                    vae._eps(min(n, a + vae.max_batch) - a))[0] for a in range(0, n, vae.max_batch)]
            H = shift_history(np.concatenate(zs), pack.song_start).astype(np.float32)
        tot += _run_song(vae, pack, H, batch_size, True, silent_weight) * n
        seen += n
    return _summarise((tot / max(seen, 1))[None, :], vae)


def train_epoch(vae, songs: Sequence[Rolls], epoch: int, batch_size: int = 256, history: bool = True, shuffle_songs: bool = False,
                rng: Optional[np.random.Generator] = None, silent_weight: float = 1.0) -> Dict[str, float]:
    order = list(range(len(songs)))
    if shuffle_songs:                                                         # vae_training.py:758-771
        (rng or np.random.default_rng(epoch)).shuffle(order)
    rows = []
    for i in order:
        H = _song_history(vae, songs[i], epoch, history)
        rows.append(_run_song(vae, songs[i], H, batch_size, True, silent_weight))
    return _summarise(np.array(rows), vae)


def evaluate_songs(vae, songs: Sequence[Rolls], batch_size: int = 256, history: bool = True) -> Dict[str, float]:
    """test() of vae_training.py:243-568 for the hot-path metrics: history always from the encoder (:284-292)."""
    rows = []
    for song in songs:
        H = _song_history(vae, song, 1, history)
        rows.append(_run_song(vae, song, H, batch_size, False))
    return _summarise(np.array(rows), vae)
