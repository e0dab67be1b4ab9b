"""Output post-processing of the decoder: the reference's process_decoder_outputs
(vae_definition.py:1131-1225) for sample_method='argmax', vectorised over packed rolls.

The GPU hands back argmax class indices (mvae_style_transfer: pitch u8 [n,T] with 60 = silent, instrument u8 [n,4])
and the velocity roll f32 [n,T]; the remaining rules are cheap integer bookkeeping per voice:
  * a silent step gets velocity 0 (:1156-1159);
  * override_sampled_pitches_based_on_velocity_info (:1161-1190): per voice, walking through time,
      - velocity below the played-note threshold but a NEW pitch (different from the previous, previous > 0):
        play it as loud as the previous played note;
      - velocity above the threshold but no pitch: velocity 0;
  * D (held-note roll) is 0 where the velocity says "struck", else 1 (:1214-1221).

The per-voice "previous pitch / previous velocity" memory of the override runs over whatever is passed in ONE call: the reference's
style-switch loop (vae_evaluation.py:2469-2483) calls it once per chunk, its whole-song call sites (:799, :814) once per song.  Call this
function the same way (tests/test_reference_pin.py pins both uses to reference-executed outputs).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

MAX_VOICES = 4
VELOCITY_THRESHOLD = 0.5      # velocity_threshold_such_that_it_is_a_played_note, settings.py:30


def process_decoder_outputs(pitch_idx: np.ndarray, instr_idx: np.ndarray, velocity: np.ndarray, num_pitches: int = 61, num_instr: int = 16,
                            override_by_velocity: bool = True, threshold: float = VELOCITY_THRESHOLD,
                            max_voices: int = MAX_VOICES) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """pitch_idx (n,T) u8, instr_idx (n,4) u8, velocity (n,T) f32  ->  (Y (n*T, 60) one-hot rolls with silent = empty row,
    I (n,4,16) one-hot, V (n*T,), D (n*T,)) exactly as the reference's Y, I, V, D."""
    silent = num_pitches - 1
    p = np.asarray(pitch_idx).reshape(-1).astype(np.int64)
    V = np.asarray(velocity, np.float64).reshape(-1).copy()
    steps = p.shape[0]
    Y = np.zeros((steps, num_pitches - 1))
    sounding = p != silent
    Y[np.nonzero(sounding)[0], p[sounding]] = 1
    V[~sounding] = 0.0
    if override_by_velocity:
        for voice in range(max_voices):
            pv = np.where(sounding[voice::max_voices], p[voice::max_voices], -1)     # pitch or -1 (silent)
            vv = V[voice::max_voices]                                                 # view into V
            vel_silent = vv < threshold
            prev_pitch = np.concatenate([[-1], pv[:-1]])
            # previous_velocity = velocity of the last step (strictly before i) whose velocity was above the threshold
            idx = np.where(~vel_silent, np.arange(len(vv)), -1)
            last = np.maximum.accumulate(idx)
            last_prev = np.concatenate([[-1], last[:-1]])
            snapshot = vv.copy()                                                      # the loop reads `velocity` from the zip of the ORIGINAL roll
            prev_vel = np.where(last_prev >= 0, snapshot[np.maximum(last_prev, 0)], 0.0)
            louden = vel_silent & (pv >= 0) & (prev_pitch > 0) & (prev_pitch != pv)
            mute = (~vel_silent) & (pv < 0)
            vv[louden] = prev_vel[louden]
            vv[mute] = 0.0
    I = np.eye(num_instr)[np.asarray(instr_idx).astype(np.int64)]
    D = np.ones(steps)
    D[V > threshold] = 0
    return Y, I, V, D
