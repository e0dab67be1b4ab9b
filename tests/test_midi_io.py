"""Roll I/O: the own Standard-MIDI-File reader / writer and the restated import_midi.load_rolls /
midi_functions.rolls_to_midi.  pretty_midi is not installed, so nothing here is pinned to the reference's output: the reader is checked
against hand-assembled SMF bytes, and (rolls_to_midi, load_rolls) against each other as a round trip on synthetic songs."""
import struct

import numpy as np

from midi_vae_b200 import midi_io, synth


def _smf(tracks, div=480, fmt=1):
    data = b"MThd" + struct.pack(">IHHH", 6, fmt, len(tracks), div)
    for t in tracks:
        data += b"MTrk" + struct.pack(">I", len(t)) + t
    return data


def test_reader_tempo_map_running_status_and_note_pairing():
    t0 = (b"\x00\xff\x51\x03\x07\xa1\x20"          # tick 0: 500000 us / quarter (120 bpm)
          b"\x87\x40\xff\x51\x03\x03\xd0\x90"      # tick 960: 250000 us / quarter (240 bpm)
          b"\x00\xff\x2f\x00")
    t1 = (b"\x00\xc0\x19"                           # program 25 on channel 0
          b"\x00\x90\x3c\x64"                       # tick 0: note on 60 vel 100
          b"\x83\x60\x3e\x50"                       # tick 480: running status, note on 62 vel 80
          b"\x83\x60\x3c\x00"                       # tick 960: running status, note on 60 vel 0 = note off
          b"\x83\x60\x80\x3e\x40"                   # tick 1440: note off 62
          b"\x00\xff\x2f\x00")
    song = midi_io.read_smf(_smf([t0, t1]))
    assert np.allclose(song.tempo_times, [0.0, 1.0]) and np.allclose(song.tempo_bpm, [120.0, 240.0])
    assert len(song.instruments) == 1 and song.instruments[0].program == 25 and not song.instruments[0].is_drum
    notes = sorted((n.pitch, n.start, n.end, n.velocity) for n in song.instruments[0].notes)
    assert notes == [(60, 0.0, 1.0, 100), (62, 0.5, 1.25, 80)]          # 1440 ticks = 1 s + 480 ticks at 240 bpm
    assert song.get_end_time() == 1.25


def test_load_rolls_picks_the_longest_constant_tempo_part_and_splits_voices():
    # 120 bpm for 2 quarters, then 60 bpm: the second part is longer; a two-note chord on one track becomes two voices (fewer tracks than voices)
    t0 = b"\x00\xff\x51\x03\x07\xa1\x20" + b"\x87\x40\xff\x51\x03\x0f\x42\x40" + b"\x00\xff\x2f\x00"
    t1 = (b"\x00\xc0\x00"
          b"\x87\x40\x90\x40\x60" b"\x00\x90\x30\x50"         # tick 960 (start of the slow part): E4 vel 96 + C3 vel 80
          b"\x87\x40\x80\x40\x00" b"\x00\x80\x30\x00"         # tick 1920: both off (2 quarters at 60 bpm = 2 s)
          b"\x00\xff\x2f\x00")
    ls = midi_io.load_rolls(_smf([t0, t1]), input_length=64)
    assert ls.tempo == 60.0 and ls.programs == [0, 0]
    P, V, D = ls.rolls.pitch.reshape(-1), ls.rolls.velocity.reshape(-1), ls.held.reshape(-1)
    # 16th grid at 60 bpm = 4 steps / s; the notes last 2 s = 8 steps from step 0 of the cut part; voice 0 = highest note
    assert np.all(P[0:32:4] == 0x40 - 24) and np.all(P[1:32:4] == 0x30 - 24) and np.all(P[2::4] == 60) and np.all(P[3::4] == 60)
    assert np.isclose(V[0], 0.5 + 0.5 * 0x60 / 127) and np.isclose(V[1], 0.5 + 0.5 * 0x50 / 127) and np.all(V[4:] == 0)
    assert D[0] == 0 and np.all(D[4:32:4] == 1) and D[32] == 0
    assert ls.rolls.pitch.shape[1] == 64 and ls.rolls.song_start[0] == 1


def test_rolls_to_midi_load_rolls_round_trip():
    T = 64
    song = synth.make_songs(1, T, seed=32, min_chunks=6, max_chunks=6)[0]
    pitch = np.concatenate([song.pitch, np.full((1, T), 60, np.uint8)])          # a silent tail chunk closes every note
    vel = np.concatenate([song.velocity, np.zeros((1, T), np.float32)])
    held = ((pitch != 60) & (vel == 0)).astype(np.uint8)
    # voices must come back in the same order: the reader sorts tracks by sounding frames, so sort the voices the same way first
    flat_p, flat_v, flat_h = pitch.reshape(-1), vel.reshape(-1), held.reshape(-1)
    order = np.argsort([np.count_nonzero(flat_p[v::4] != 60) for v in range(4)], kind="stable")[::-1]
    fp, fv, fh = flat_p.copy(), flat_v.copy(), flat_h.copy()
    for new, old in enumerate(order):
        fp[new::4], fv[new::4], fh[new::4] = flat_p[old::4], flat_v[old::4], flat_h[old::4]
    programs = [0, 8, 40, 72]
    data = midi_io.rolls_to_midi(fp, programs, None, bpm=120.0, velocity=fv, held=fh)
    mid = midi_io.read_smf(data)
    mid.tempo_bpm = mid.tempo_bpm / 4.0        # rolls_to_midi writes one roll step as one quarter at 4 x the tempo (midi_functions.py:60)
    ls = midi_io.load_rolls(mid, input_length=T)
    n = min(len(ls.rolls), pitch.shape[0])
    assert n >= pitch.shape[0] - 1
    got_p, got_v, got_h = ls.rolls.pitch[:n].reshape(-1), ls.rolls.velocity[:n].reshape(-1), ls.held[:n].reshape(-1)
    m = n * T
    counts = [np.count_nonzero(fp[v::4] != 60) for v in range(4)]
    assert len(set(counts)) == 4, "test song needs distinct voice densities for an unambiguous voice order"
    assert np.array_equal(got_p, fp[:m])
    assert np.array_equal(got_h, fh[:m])
    assert np.abs(got_v - fv[:m]).max() <= 0.5 / 127 + 1e-6       # int() truncation of the velocity on the way out (as in the reference)
    assert list(ls.rolls.instr[0]) == [p // 8 for p in programs]
