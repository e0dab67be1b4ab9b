"""Roll I/O: MIDI file -> packed rolls and packed rolls -> MIDI file, without pretty_midi / mido.

The reference reads songs with pretty_midi (import_midi.py:13-350) and writes them with pretty_midi + mido
(midi_functions.py:57-137); neither package is installed here, so this module carries its own Standard-MIDI-File
reader / writer and restates the two reference functions on top of it:

  * ``read_smf``      -- SMF format 0 / 1 parser: tempo map, notes per (track, channel, program) instrument with the
                         pairing rule pretty_midi uses (a note-off closes every open note of its key), times in seconds;
  * ``load_rolls``    -- import_midi.load_rolls: longest constant-tempo section, tracks ordered by sounding frames,
                         16th-note grid, highest-note-first voice split (up to ``max_voices`` monophonic voices),
                         pitch crop 24..84 + silent class, velocity / held-note rolls, voice interleaving
                         ``index = step * max_voices + voice`` (:249), right padding with silence, chunks of
                         ``input_length``.  Output is PACKED (synth.Rolls): class index per step, not one-hot rows;
  * ``rolls_to_midi`` -- midi_functions.rolls_to_midi: per voice, notes are struck where the held-note roll is 0
                         (or every new pitch / every bar start without one), velocities mapped back to 0..127.

Parity note: with pretty_midi absent nothing here can be pinned to the reference's own output; tests/test_midi_io.py
checks the reader against hand-assembled SMF bytes and the pair (rolls_to_midi, load_rolls) as a round trip.
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .synth import Rolls

# settings.py:70-76,86
LOW_CROP, HIGH_CROP, NUM_NOTES = 24, 84, 128
SMALLEST_NOTE = 16
MAX_VOICES = 4
MAX_VOICES_PER_TRACK = 1
MAX_VELOCITY = 127.0
VELOCITY_THRESHOLD = 0.5
SILENT = HIGH_CROP - LOW_CROP          # class 60


# ------------------------------------------------------------------------------------------------ SMF reader
@dataclass
class Note:
    start: float
    end: float
    pitch: int
    velocity: int


@dataclass
class Instrument:
    program: int
    is_drum: bool = False
    notes: List[Note] = field(default_factory=list)


@dataclass
class Song:
    resolution: int
    instruments: List[Instrument]
    tempo_times: np.ndarray          # seconds
    tempo_bpm: np.ndarray
    other_event_times: List[float]   # time / key signatures: part of pretty_midi's get_end_time

    def get_tempo_changes(self):
        return self.tempo_times, self.tempo_bpm

    def get_end_time(self) -> float:
        ends = [n.end for i in self.instruments for n in i.notes] + list(self.other_event_times)
        return max(ends) if ends else 0.0


def _vlq(data: bytes, pos: int) -> Tuple[int, int]:
    v = 0
    while True:
        b = data[pos]; pos += 1
        v = (v << 7) | (b & 0x7F)
        if not b & 0x80:
            return v, pos


def _parse_track(data: bytes) -> List[Tuple[int, tuple]]:
    """-> [(absolute tick, event)], event = ('on'|'off', ch, pitch, vel) | ('program', ch, p) | ('tempo', us_per_quarter) | ('sig',)"""
    out, pos, tick, status = [], 0, 0, 0
    n = len(data)
    while pos < n:
        delta, pos = _vlq(data, pos)
        tick += delta
        b = data[pos]
        if b & 0x80:
            status = b; pos += 1
        elif not status:
            raise ValueError("running status without a previous status byte")
        if status == 0xFF:                                    # meta
            mtype = data[pos]; pos += 1
            ln, pos = _vlq(data, pos)
            body = data[pos:pos + ln]; pos += ln
            if mtype == 0x51 and ln == 3:
                out.append((tick, ("tempo", int.from_bytes(body, "big"))))
            elif mtype in (0x58, 0x59):
                out.append((tick, ("sig",)))
            elif mtype == 0x2F:
                break
            status = 0                                        # meta / sysex cancel running status
        elif status in (0xF0, 0xF7):                          # sysex
            ln, pos = _vlq(data, pos)
            pos += ln
            status = 0
        else:
            kind, ch = status & 0xF0, status & 0x0F
            if kind in (0x80, 0x90, 0xA0, 0xB0, 0xE0):
                d1, d2 = data[pos], data[pos + 1]; pos += 2
                if kind == 0x90 and d2 > 0:
                    out.append((tick, ("on", ch, d1, d2)))
                elif kind == 0x80 or kind == 0x90:
                    out.append((tick, ("off", ch, d1, 0)))
            elif kind in (0xC0, 0xD0):
                d1 = data[pos]; pos += 1
                if kind == 0xC0:
                    out.append((tick, ("program", ch, d1)))
            else:
                raise ValueError(f"unexpected status byte {status:#x}")
    return out


def read_smf(src) -> Song:
    """Parse a Standard MIDI File (path or bytes).  Raises ValueError on malformed input (import_midi.py:19-23 skips such files)."""
    data = src if isinstance(src, (bytes, bytearray)) else open(src, "rb").read()
    if data[:4] != b"MThd":
        raise ValueError("not a Standard MIDI File")
    hlen, fmt, ntrk, div = struct.unpack(">IHHH", data[4:14])
    if div & 0x8000:
        raise ValueError("SMPTE time division is not supported")
    pos = 8 + hlen
    tracks = []
    for _ in range(ntrk):
        if data[pos:pos + 4] != b"MTrk":
            raise ValueError("missing track chunk")
        ln = struct.unpack(">I", data[pos + 4:pos + 8])[0]
        tracks.append(_parse_track(bytes(data[pos + 8:pos + 8 + ln])))
        pos += 8 + ln
    # tempo map (all tracks contribute, as mido's merged view does): tick -> seconds, piecewise linear
    tempi = sorted((t, e[1]) for tr in tracks for t, e in tr if e[0] == "tempo")
    if not tempi or tempi[0][0] != 0:
        tempi.insert(0, (0, 500000))
    seg_tick, seg_time, seg_us = [tempi[0][0]], [0.0], [tempi[0][1]]
    for t, us in tempi[1:]:
        if t == seg_tick[-1]:
            seg_us[-1] = us
            continue
        if us == seg_us[-1]:
            continue
        seg_time.append(seg_time[-1] + (t - seg_tick[-1]) * seg_us[-1] / 1e6 / div)
        seg_tick.append(t); seg_us.append(us)
    seg_tick_a, seg_time_a, seg_us_a = np.array(seg_tick), np.array(seg_time), np.array(seg_us, np.float64)

    def to_sec(tick: int) -> float:
        i = int(np.searchsorted(seg_tick_a, tick, side="right") - 1)
        return float(seg_time_a[i] + (tick - seg_tick_a[i]) * seg_us_a[i] / 1e6 / div)

    instruments: Dict[Tuple[int, int, int], Instrument] = {}
    order: List[Tuple[int, int, int]] = []
    others: List[float] = []
    for ti, tr in enumerate(tracks):
        program = [0] * 16
        open_notes: Dict[Tuple[int, int], List[Tuple[int, int]]] = {}
        for tick, e in tr:
            if e[0] == "program":
                program[e[1]] = e[2]
            elif e[0] == "on":
                open_notes.setdefault((e[1], e[2]), []).append((tick, e[3]))
            elif e[0] == "off":
                key = (e[1], e[2])
                if key in open_notes:
                    ikey = (program[e[1]], e[1], ti)
                    if ikey not in instruments:
                        instruments[ikey] = Instrument(program[e[1]], e[1] == 9); order.append(ikey)
                    keep = []
                    for st, vel in open_notes[key]:
                        if st != tick:
                            instruments[ikey].notes.append(Note(to_sec(st), to_sec(tick), e[2], vel))
                        else:
                            keep.append((st, vel))
                    if keep:
                        open_notes[key] = keep
                    else:
                        del open_notes[key]
            elif e[0] == "sig":
                others.append(to_sec(tick))
    return Song(div, [instruments[k] for k in order], seg_time_a.copy(), 6e7 / seg_us_a, others)


def _sounding_frames(inst: Instrument, fs: float = 100.0) -> int:
    """np.count_nonzero(instrument.get_piano_roll(fs=100)) (import_midi.py:71-74); pretty_midi returns an all-zero roll for drum tracks."""
    if inst.is_drum or not inst.notes:
        return 0
    end = max(n.end for n in inst.notes)
    roll = np.zeros((128, int(fs * end) + 1), bool)
    for n in inst.notes:
        roll[n.pitch, int(n.start * fs):int(n.end * fs)] = True
    return int(roll.sum())


# ------------------------------------------------------------------------------------------------ MIDI -> rolls
@dataclass
class LoadedSong:
    rolls: Rolls                    # packed chunks: pitch (N,T) class index, instr (N,4) category, velocity (N,T), style filled by the caller
    held: np.ndarray                # (N,T) 1 where a note is held (the reference's D)
    programs: List[int]
    tempo: float


def load_rolls(src, input_length: int = 64, max_voices: int = MAX_VOICES, style: int = 0) -> Optional[LoadedSong]:
    """import_midi.load_rolls (:13-350) on the settings.py defaults; ``input_length`` counts interleaved steps (16 steps x 4 voices = 64).
    Returns None when no track has a note (the reference returns all-None)."""
    mid = src if isinstance(src, Song) else read_smf(src)
    times, bpms = mid.get_tempo_changes()
    song_start, song_end = 0.0, mid.get_end_time()
    if len(times) > 1:                                          # longest constant-tempo part (:31-52)
        best = -1.0
        full_end = song_end
        for i, t0 in enumerate(times):
            t1 = full_end if i == len(times) - 1 else times[i + 1]
            if t1 - t0 > best:
                best, song_start, song_end, tempo = t1 - t0, float(t0), float(t1), float(bpms[i])
    else:
        tempo = float(bpms[0])
    insts = []
    for inst in mid.instruments:                                # cut to that part (:57-67)
        kept = [Note(n.start - song_start, n.end - song_start, n.pitch, n.velocity) for n in inst.notes if n.start >= song_start and n.end <= song_end]
        insts.append(Instrument(inst.program, inst.is_drum, kept))
    counts = [_sounding_frames(i) for i in insts]
    insts = [insts[i] for i in np.argsort(counts)[::-1]]        # most sounding frames first (:70-76)

    fs = 1.0 / ((60.0 / tempo) * 4.0 / SMALLEST_NOTE)           # 16th notes per second (:83-88)
    total = int(math.ceil(song_end * fs))
    piano, vel, held, maxc = [], [], [], []
    for inst in insts:                                          # per track: piano roll, per-voice velocity / held rolls (:99-152)
        pr = np.zeros((total, 128), bool)
        conc = np.zeros(total, np.int64)
        onset_vel: Dict[Tuple[int, int], int] = {}
        for n in inst.notes:
            ts, te = n.start * fs, n.end * fs
            a, b = int(round(ts)), int(round(te))
            if ts - a < 10e-3 or b - a >= 1:
                pr[a:b, n.pitch] = True
                conc[a:b] += 1
                onset_vel[(a, n.pitch)] = n.velocity
        mc = int(conc.max()) if total else 0
        vr, hr = np.zeros((total, mc)), np.zeros((total, mc))
        for step in np.nonzero(pr.any(1))[0]:
            for voice, pitch in enumerate(np.nonzero(pr[step])[0][::-1]):
                if voice >= mc:
                    break
                if (step, pitch) in onset_vel:
                    vr[step, voice] = onset_vel[(step, pitch)]
                else:
                    hr[step, voice] = 1
        piano.append(pr); vel.append(vr); held.append(hr); maxc.append(mc)

    # voices per track: 1, more for the first tracks when the file has fewer sounding tracks than voices (:160-171)
    override = [MAX_VOICES_PER_TRACK] * len(maxc)
    silent_left = max_voices - sum(min(MAX_VOICES_PER_TRACK, x) if x > 0 else 0 for x in maxc[:max_voices])
    for v in range(min(max_voices, len(maxc))):
        if silent_left > 0 and maxc[v] > MAX_VOICES_PER_TRACK:
            extra = min(silent_left, maxc[v] - MAX_VOICES_PER_TRACK)
            override[v] += extra; silent_left -= extra

    chosen_p, chosen_v, chosen_h, programs = [], [], [], []
    for pr, vr, hr, inst, mc, ov in zip(piano, vel, held, insts, maxc, override):   # highest note = voice 0, ... (:186-230)
        if mc <= 0:
            continue
        for voice in range(min(mc, max(MAX_VOICES_PER_TRACK, ov))):
            if len(chosen_p) >= max_voices:
                break
            mono = np.full(total, -1, np.int64)
            for step in np.nonzero(pr.any(1))[0]:
                notes = np.nonzero(pr[step])[0][::-1]
                if len(notes) > voice:
                    mono[step] = notes[voice]
            chosen_p.append(mono); chosen_v.append(vr[:, voice]); chosen_h.append(hr[:, voice]); programs.append(inst.program)
        if len(chosen_p) == max_voices:
            break
    if not chosen_p:
        return None

    length = total * max_voices
    P = np.full(length, SILENT, np.int64)                       # interleave: index = step * max_voices + voice (:243-262)
    V = np.zeros(length); D = np.zeros(length)
    for i, (mono, vr, hr) in enumerate(zip(chosen_p, chosen_v, chosen_h)):
        inside = (mono >= LOW_CROP) & (mono < HIGH_CROP)        # pitches outside the crop become silent (:257-266)
        idx = np.arange(total) * max_voices + i
        P[idx[inside]] = mono[inside] - LOW_CROP
        struck = vr > 0
        V[idx[struck]] = VELOCITY_THRESHOLD + (vr[struck] / MAX_VELOCITY) * (1.0 - VELOCITY_THRESHOLD)   # :268-277
        D[idx] = hr                                             # :283-286
    pad = (-length) % input_length                              # right padding with silence, then chunks (:304-346)
    P = np.concatenate([P, np.full(pad, SILENT, np.int64)]).reshape(-1, input_length)
    V = np.concatenate([V, np.zeros(pad)]).reshape(-1, input_length)
    D = np.concatenate([D, np.zeros(pad)]).reshape(-1, input_length)
    cats = np.zeros(max_voices, np.int64)                       # '1hot-category': program // 8; unused voices keep category 0 ... as an all-zero
    cats[:len(programs)] = np.array(programs) // 8              # row in the reference; packed rolls cannot express "no category", see has_voice
    n = P.shape[0]
    ss = np.zeros(n, np.uint8); ss[0] = 1
    rolls = Rolls(P.astype(np.uint8), np.tile(cats.astype(np.uint8), (n, 1)), V.astype(np.float32), np.full(n, style, np.uint8), ss)
    return LoadedSong(rolls, D.astype(np.uint8), programs, tempo)


# ------------------------------------------------------------------------------------------------ rolls -> MIDI
def _vlq_bytes(v: int) -> bytes:
    out = [v & 0x7F]
    v >>= 7
    while v:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    return bytes(reversed(out))


def rolls_to_notes(pitch: np.ndarray, programs: Sequence[int], velocity: Optional[np.ndarray] = None, held: Optional[np.ndarray] = None):
    """The note list midi_functions.rolls_to_midi (:57-137) builds: per voice [(start_step, end_step, midi_pitch, velocity)].
    ``pitch`` is the flattened interleaved class-index roll (60 = silent)."""
    p = np.asarray(pitch).reshape(-1).astype(np.int64)
    nv = len(programs)
    out = []
    for voice in range(nv):
        cur = p[voice::nv]
        vroll = None
        if velocity is not None:
            vroll = np.asarray(velocity, np.float64).reshape(-1)[voice::nv].copy()
            vroll[vroll < VELOCITY_THRESHOLD] = 0
            vroll[vroll >= VELOCITY_THRESHOLD] -= 0.5
            vroll /= (1.0 - VELOCITY_THRESHOLD)
            vroll *= MAX_VELOCITY
        hroll = None if held is None else np.asarray(held).reshape(-1)[voice::nv]
        notes = []
        tracked: Optional[int] = None       # a voice is monophonic: at most one sounding pitch
        start, vel = 0, 80
        for i, c in enumerate(cur):
            now = None if c == SILENT else int(c) + LOW_CROP
            if tracked is not None:
                if hroll is not None:
                    hold = hroll[i] > 0.5 and now == tracked
                else:
                    hold = now == tracked and i % SMALLEST_NOTE != 0
                if hold:
                    now = None              # still the same note: nothing new is struck
                else:
                    notes.append((start, i, tracked, min(int(vel), int(MAX_VELOCITY)) if vroll is not None else 80))
                    tracked = None
            if now is not None:
                tracked, start = now, i
                if vroll is not None:
                    vel = int(vroll[i])
        out.append(notes)                   # a note still sounding at the end of the roll is dropped, as in the reference loop
    return out


def rolls_to_midi(pitch: np.ndarray, programs: Sequence[int], path: Optional[str], bpm: float, velocity: Optional[np.ndarray] = None,
                  held: Optional[np.ndarray] = None, resolution: int = 1000) -> bytes:
    """Write the rolls as an SMF format-1 file (one track per voice) and return its bytes; one roll step = one tick * resolution at the
    scaled tempo bpm * SMALLEST_NOTE / 4, as midi_functions.rolls_to_midi sets it up (:60, :67)."""
    step_bpm = bpm * (SMALLEST_NOTE / 4)
    us = int(round(6e7 / step_bpm))
    tracks = [b"\x00\xff\x51\x03" + us.to_bytes(3, "big") + b"\x00\xff\x58\x04\x04\x02\x18\x08" + b"\x00\xff\x2f\x00"]
    for voice, (program, notes) in enumerate(zip(programs, rolls_to_notes(pitch, programs, velocity, held))):
        ch = voice if voice < 9 else voice + 1
        ev = []
        for s, e, pit, vel in notes:
            ev.append((s * resolution, 1, bytes([0x90 | ch, pit, max(1, vel)])))
            ev.append((e * resolution, 0, bytes([0x80 | ch, pit, 0])))
        ev.sort(key=lambda x: (x[0], x[1]))
        body, last = b"\x00" + bytes([0xC0 | ch, int(program)]), 0
        for tick, _, msg in ev:
            body += _vlq_bytes(tick - last) + msg
            last = tick
        tracks.append(body + b"\x00\xff\x2f\x00")
    data = b"MThd" + struct.pack(">IHHH", 6, 1, len(tracks), resolution)
    for t in tracks:
        data += b"MTrk" + struct.pack(">I", len(t)) + t
    if path is not None:
        with open(path, "wb") as f:
            f.write(data)
    return data
