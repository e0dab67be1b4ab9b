"""Seeded synthetic MIDI rolls in the reference's tensor layout.

The reference turns MIDI files into rolls in import_midi.py:243-286: up to four
monophonic voices, 16th-note steps, voice-interleaved (``index = step*max_voices +
voice``, import_midi.py:245-249), one-hot over 60 pitches (MIDI 24..83) plus a silent
class at index 60; a velocity roll that is ``0.5 + 0.5*vel/127`` at note onsets and 0
elsewhere (import_midi.py:269-277); and one instrument category (program // 8,
midi_functions.py:22-27) per voice.  pretty_midi is not available, so rolls of the same
shape and statistics are generated here.  Rolls are kept PACKED (u8 class indices), the
layout the C ABI takes; ``dense()`` expands them to the reference's one-hot tensors.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

MAX_VOICES = 4          # settings.py:86
NUM_PITCHES = 61        # 60 pitches + silent, settings.py:147-153
SILENT = 60
NUM_INSTR = 16          # settings.py:181
VOICE_START = (45, 33, 26, 14)


@dataclass
class Rolls:
    """One batch (or one song) of packed rolls."""
    pitch: np.ndarray      # u8 (N,T)   class index per (chunk, step); SILENT = 60
    instr: np.ndarray      # u8 (N,4)   instrument category per voice
    velocity: np.ndarray   # f32 (N,T)  0 or 0.5..1 at onsets
    style: np.ndarray      # u8 (N,)    style class
    song_start: Optional[np.ndarray] = None   # bool (N,) first chunk of a song

    def __len__(self):
        return self.pitch.shape[0]

    def slice(self, a, b):
        ss = None if self.song_start is None else self.song_start[a:b]
        return Rolls(self.pitch[a:b], self.instr[a:b], self.velocity[a:b], self.style[a:b], ss)

    def dense(self, dtype=np.float64):
        """-> X (N,T,61), I (N,4,16), V (N,T,1), C (N,2) as the reference feeds Keras."""
        X = np.eye(NUM_PITCHES, dtype=dtype)[self.pitch]
        I = np.eye(NUM_INSTR, dtype=dtype)[self.instr]
        V = self.velocity.astype(dtype)[..., None]
        C = np.eye(2, dtype=dtype)[self.style]
        return X, I, V, C


def _voice_track(rng, voice: int, steps: int):
    """Random-walk melody for one voice: (pitch[steps], velocity[steps])."""
    pitch = np.full(steps, SILENT, np.uint8)
    vel = np.zeros(steps, np.float32)
    cur = int(np.clip(VOICE_START[voice] + rng.integers(-3, 4), 0, 59))
    s = 0
    while s < steps:
        dur = int(rng.choice((1, 2, 4, 8)))
        cur = int(np.clip(cur + rng.integers(-4, 5), 0, 59))
        if rng.random() >= 0.15:
            pitch[s:s + dur] = cur
            vel[s] = 0.5 + 0.5 * float(rng.integers(50, 110)) / 127.0
        s += dur
    return pitch, vel


def make_song(rng, n_chunks: int, T: int, style: int) -> Rolls:
    assert T % MAX_VOICES == 0
    steps = n_chunks * T // MAX_VOICES
    P = np.empty((steps, MAX_VOICES), np.uint8)
    Vv = np.empty((steps, MAX_VOICES), np.float32)
    for v in range(MAX_VOICES):
        P[:, v], Vv[:, v] = _voice_track(rng, v, steps)
    pitch = P.reshape(n_chunks, T)            # index = step*4 + voice
    velocity = Vv.reshape(n_chunks, T)
    instr = np.tile(rng.integers(0, NUM_INSTR, MAX_VOICES).astype(np.uint8), (n_chunks, 1))
    ss = np.zeros(n_chunks, bool); ss[0] = True
    return Rolls(pitch, instr, velocity, np.full(n_chunks, style, np.uint8), ss)


def make_batch(batch: int, T: int, seed: int = 1234) -> Rolls:
    """``batch`` independent chunks (one per synthetic song), style = b mod 2."""
    rng = np.random.default_rng(seed)
    songs = [make_song(rng, 1, T, b % 2) for b in range(batch)]
    return concat(songs)


def make_songs(n_songs: int, T: int, seed: int = 1234, min_chunks: int = 8, max_chunks: int = 40) -> List[Rolls]:
    """cfg1-style corpus: 2 styles x n_songs/2 files of min..max chunks each."""
    rng = np.random.default_rng(seed)
    return [make_song(rng, int(rng.integers(min_chunks, max_chunks + 1)), T, s % 2) for s in range(n_songs)]


def concat(parts: List[Rolls]) -> Rolls:
    ss = None if any(p.song_start is None for p in parts) else np.concatenate([p.song_start for p in parts])
    return Rolls(np.concatenate([p.pitch for p in parts]), np.concatenate([p.instr for p in parts]),
                 np.concatenate([p.velocity for p in parts]), np.concatenate([p.style for p in parts]), ss)


def make_eps(batch: int, L: int, seed: int, std: float = 0.01) -> np.ndarray:
    return (np.random.default_rng(seed + 7919).standard_normal((batch, L)) * std).astype(np.float32)
