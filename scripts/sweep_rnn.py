"""Persistent-vs-streamed recurrence sweep (BASELINE config 5) and the style-transfer inference bench (config 4).

    python scripts/sweep_rnn.py sweep      # train-step ms for (T, H) x {persistent, streamed}, bf16, batch 128..512
    python scripts/sweep_rnn.py infer      # encode -> swap -> decode -> argmax, batch 1024, p50/p99 latency

Prints one JSON object per line; profiles/r1/README.md tabulates the committed run.
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from midi_vae_b200 import Engine, EngineConfig, _lib, initial_weights, synth  # noqa: E402


def train_ms(T, H, L, B, mode, steps=3, warmup=2):
    import torch
    cfg = EngineConfig(input_length=T, lstm_size=H, latent_rep_size=L, decoder_feedback="teacher_forced", precision="bf16", rnn_mode=mode, max_batch=B)
    eng = Engine(cfg, 0)
    eng.set_weights(initial_weights(cfg, 42))
    r = synth.make_batch(B, T, seed=1)
    eps = synth.make_eps(B, L, 1)
    t = dict(pitch=torch.from_numpy(r.pitch).cuda(), instr=torch.from_numpy(r.instr).cuda(), vel=torch.from_numpy(r.velocity).cuda(),
             style=torch.from_numpy(r.style).cuda(), eps=torch.from_numpy(eps).cuda())
    b = eng.device_batch(B, t["pitch"].data_ptr(), t["instr"].data_ptr(), t["vel"].data_ptr(), t["style"].data_ptr(), None, t["eps"].data_ptr())
    stream = torch.cuda.ExternalStream(eng.stream())
    for _ in range(warmup):
        eng.train_step_device(b)
    eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            eng.train_step_device(b)
        e1.record(stream)
    eng.sync()
    ms = e0.elapsed_time(e1) / steps
    eng.set_profiling(True); eng.train_step_device(b); eng.sync(); k = eng.kernel_ms(); eng.set_profiling(False)
    eng.close()
    return ms, {n: round(v[0], 3) for n, v in k.items()}


def sweep():
    for T, H, L, B in ((64, 256, 100, 128), (256, 256, 256, 512), (256, 512, 256, 512), (128, 1024, 256, 256), (1024, 1024, 256, 128)):
        row = {"T": T, "H": H, "L": L, "B": B}
        for mode in ("persistent", "streamed"):
            try:
                ms, k = train_ms(T, H, L, B, mode, steps=2 if T * H >= 2 ** 19 else 3)
                row[mode] = {"ms_per_step": round(ms, 3), "seq_per_s": round(B / ms * 1e3, 1), "rec_fwd_ms": k["rec_fwd"], "rec_bwd_ms": k["rec_bwd"], "gemm_ms": k["gemm"]}
            except _lib.MvaeError as ex:
                row[mode] = {"error": str(ex)[:160]}
        print(json.dumps(row), flush=True)


def infer():
    import torch
    for T, H, L in ((64, 256, 100), (256, 512, 256)):
        for fb in ("as_wired", "free_running"):
            B = 1024
            cfg = EngineConfig(input_length=T, lstm_size=H, latent_rep_size=L, precision="bf16", max_batch=B)
            eng = Engine(cfg, 0)
            eng.set_weights(initial_weights(cfg, 42))
            songs = synth.concat(synth.make_songs(16, T, seed=5, min_chunks=64, max_chunks=64))
            t = dict(pitch=torch.from_numpy(songs.pitch).cuda(), instr=torch.from_numpy(songs.instr).cuda(), vel=torch.from_numpy(songs.velocity).cuda(),
                     ss=torch.from_numpy(songs.song_start.astype(np.uint8)).cuda())
            op, oi, ov = torch.empty(B, T, dtype=torch.uint8, device="cuda"), torch.empty(B, 4, dtype=torch.uint8, device="cuda"), torch.empty(B, T, device="cuda")
            b = eng.device_batch(B, t["pitch"].data_ptr(), t["instr"].data_ptr(), t["vel"].data_ptr())
            stream = torch.cuda.ExternalStream(eng.stream())
            n = 20 if fb == "free_running" else 100
            lat = []
            for i in range(n + 3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    e0.record(stream)
                    eng.style_transfer_device(b, t["ss"].data_ptr(), 0, 1, fb, op.data_ptr(), oi.data_ptr(), ov.data_ptr())
                    e1.record(stream)
                eng.sync()
                if i >= 3:
                    lat.append(e0.elapsed_time(e1))
            lat = np.array(lat)
            print(json.dumps({"workload": f"cfg4 style transfer: batch {B}, seq_len={T} hidden={H} latent={L}, decoder_feedback={fb}, bf16",
                              "p50_ms": round(float(np.percentile(lat, 50)), 3), "p99_ms": round(float(np.percentile(lat, 99)), 3),
                              "seq_per_s": round(B / float(np.percentile(lat, 50)) * 1e3, 1), "runs": n}), flush=True)
            eng.close()


if __name__ == "__main__":
    {"sweep": sweep, "infer": infer}[sys.argv[1] if len(sys.argv) > 1 else "sweep"]()
