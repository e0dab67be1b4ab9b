#!/usr/bin/env python
"""bench.py -- MIDI sequences/sec of the MIDI-VAE hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|cfg1|cfg4|cfg5] [--precision bf16|fp32]
    python bench.py --impl reference ...      # the CPU oracle (the reference's Keras path cannot run here) on the host cores

Workloads = BASELINE.json configs: cfg1..cfg3 and cfg5 time one TRAIN step (vae_training.py:804-809: one mini-batch of autoencoder.fit =
forward + losses + backward + [all-reduce] + Keras-Adam); cfg3 is the headline (the default).  cfg4 times one batch-1024 style-transfer
INFERENCE call (vae_evaluation.py:2448-2550 batched: encode -> swap -> history shift -> decode -> argmax) and adds p50 / p99 latency;
cfg5 (T1024 / H1024) adds the persistent-vs-streamed recurrence comparison north_star asks for.

Contract (one JSON line on stdout from rank 0): metric/value/unit/n_gpus/steps/warmup/ms_per_step/higher_is_better/scaling/vs_baseline/
dtype/data/config/clocks/e2e/gpu_launches/roofline/cpu_baseline.  `value` is timed with inputs resident in HBM, `e2e` through the public
host-pointer API with the H2D copies of the step's inputs and the D2H read of its result inside the timed region.  The CPU arms
(`cpu_baseline`, `--impl reference`) time a BOUNDED SAMPLE of the workload's batch and say so in `cpu_baseline.sample` and `config.cpu_sample`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {   # BASELINE.json configs
    "cfg1": dict(T=16, H=64, L=16, B=8, kind="train"),
    "cfg2": dict(T=64, H=256, L=100, B=128, kind="train"),
    "cfg3": dict(T=256, H=512, L=256, B=512, kind="train"),
    "cfg4": dict(T=256, H=512, L=256, B=1024, kind="infer"),      # --cfg4-shape cfg2 runs it at T64 / H256 / L100
    "cfg5": dict(T=1024, H=1024, L=256, B=128, kind="train"),
    # not a BASELINE config: the reference's own defaults (settings.py:108-112,155: GRU cells, T64 H256 L256, batch_size 256, decoder as wired)
    "refdefault": dict(T=64, H=256, L=256, B=256, kind="train", cell="GRU", feedback="as_wired"),
}
METRIC_TRAIN = "MIDI sequences/sec (train step)"
METRIC_INFER = "MIDI sequences/sec (style-transfer inference)"


def flops_per_seq(T, H, L, feedback="teacher_forced", ne=2, nd=2, Dp=61, Di=16, Ti=4, Dv=1, cell="LSTM"):
    """Algorithmic GEMM FLOPs per sequence, forward (SURVEY.md 8(d) formula; 3 gate blocks for GRU).  Returns (total_fwd, recurrent_fwd)."""
    G = (3 if cell == "GRU" else 4) * H
    enc = 2 * G * (T * (Dp + H) + (ne - 1) * T * 2 * H + Ti * (Di + H) + T * (Dv + H))
    head = 2 * (3 * H * H + H * H + 2 * (H // 2) * L)
    init = (nd + 2) * 2 * 2 * (2 * L) * H
    tf = feedback == "teacher_forced"
    dec = 2 * G * (T * ((Dp if tf else 0) + H) + (nd - 1) * T * 2 * H + Ti * ((Di if tf else 0) + H) + T * ((Dv if tf else 0) + H))
    out = 2 * H * (T * Dp + Ti * Di + T * Dv)
    total = enc + head + init + dec + out
    rec = 2 * G * H * ((ne + 1) * T + Ti) + 2 * G * H * ((nd + 1) * T + Ti)     # the h U products only
    return total, rec


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_sustained=d.get("bf16_tflops_sustained"), bf16_burst=d.get("bf16_tflops"), hbm=d.get("hbm_gbs"), source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_sustained=1400.0, bf16_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def workload_string(name, wl, feedback, world=1):
    """The SAME string in the GPU arm and the reference arm: it names the workload, not how much of it an arm sampled."""
    if wl["kind"] == "infer":
        return (f"{name}: style-transfer inference (encode -> swap style dims -> history shift -> decode -> argmax), batch={wl['B']}/GPU = 16 synthetic songs x 64 chunks, "
                f"seq_len={wl['T']} hidden={wl['H']} latent={wl['L']}, decoder_feedback={feedback}, {wl.get('cell', 'LSTM')}, 2+2 layers, hard_sigmoid gates "
                "(CPU arms time a bounded sample of this batch: see cpu_baseline.sample)")
    return (f"{name}: train step, seq_len={wl['T']} hidden={wl['H']} latent={wl['L']} batch={wl['B']}/GPU, decoder_feedback={feedback}, "
            f"{wl.get('cell', 'LSTM')}, 2+2 layers, hard_sigmoid gates (CPU arms time a bounded sample of this batch: see cpu_baseline.sample)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1])); smax = float(f[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU oracle legs
def cpu_oracle_train_rate(wl, feedback, sample_batch, steps, warmup):
    """Reference CPU path = the fp32 PyTorch-CPU oracle, all host threads, on `sample_batch` sequences of the batch."""
    import torch
    from midi_vae_b200 import EngineConfig, initial_weights, synth
    from oracle import midivae_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    ocfg = O.OracleConfig(input_length=wl["T"], lstm_size=wl["H"], latent_rep_size=wl["L"], decoder_feedback=feedback, cell_type=wl.get("cell", "LSTM"))
    ecfg = EngineConfig(input_length=wl["T"], lstm_size=wl["H"], latent_rep_size=wl["L"], cell_type=wl.get("cell", "LSTM"))
    p = {k: torch.tensor(v, dtype=torch.float32) for k, v in initial_weights(ecfg, 42).items()}
    opt = O.KerasAdam(p, lr=2e-4)
    r = synth.make_batch(sample_batch, wl["T"], seed=1234)
    X, I, V, C = [torch.tensor(a, dtype=torch.float32) for a in r.dense(np.float32)]
    hist = torch.zeros(sample_batch, wl["L"]); eps = torch.tensor(synth.make_eps(sample_batch, wl["L"], 1234))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.train_on_batch(ocfg, p, opt, X, I, V, C, hist, eps)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.median(times))
    return sample_batch / (ms / 1e3), ms, torch.get_num_threads()


def cpu_oracle_infer_rate(wl, feedback, sample_batch, steps, warmup):
    """The oracle's batched style transfer (fp32, all host threads) on one synthetic song of `sample_batch` chunks."""
    import torch
    from midi_vae_b200 import EngineConfig, initial_weights, synth
    from oracle import midivae_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    ocfg = O.OracleConfig(input_length=wl["T"], lstm_size=wl["H"], latent_rep_size=wl["L"])
    ecfg = EngineConfig(input_length=wl["T"], lstm_size=wl["H"], latent_rep_size=wl["L"])
    p = {k: torch.tensor(v, dtype=torch.float32) for k, v in initial_weights(ecfg, 42).items()}
    song = synth.make_songs(1, wl["T"], seed=5, min_chunks=sample_batch, max_chunks=sample_batch)[0]
    X, I, V, C = [torch.tensor(a, dtype=torch.float32) for a in song.dense(np.float32)]
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.style_transfer(ocfg, p, X, I, V, 0, 1, song.song_start, feedback)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.median(times))
    return sample_batch / (ms / 1e3), ms, torch.get_num_threads()


def cpu_leg(args, wl, feedback, steps, warmup):
    sample = max(1, min(wl["B"], args.cpu_sample))
    if wl["kind"] == "infer":
        rate, ms, cores = cpu_oracle_infer_rate(wl, feedback, sample, steps, warmup)
        what = "style transfer (encode -> swap -> decode -> argmax)"
    else:
        rate, ms, cores = cpu_oracle_train_rate(wl, feedback, sample, steps, warmup)
        what = "train step"
    return {"value": rate, "unit": "sequences/s", "cores": cores, "kind": "port",
            "sample": f"{sample}-of-{wl['B']} sample: oracle fp32 torch-CPU {what} on {sample} of the {wl['B']} sequences of the workload's batch, "
                      f"median of {steps} timed steps after {warmup} warm-up ({ms:.0f} ms/step; throughput is flat in the batch size on the CPU)"}, ms


def run_reference(args, name, wl, feedback):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    cb, ms = cpu_leg(args, wl, feedback, steps, warmup)
    print(json.dumps({
        "impl": "reference", "metric": METRIC_INFER if wl["kind"] == "infer" else METRIC_TRAIN, "value": cb["value"], "unit": "sequences/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(name, wl, feedback), "cpu_sample": f"{max(1, min(wl['B'], args.cpu_sample))} of {wl['B']} sequences per step",
                   "note": "reference Keras 2.0.8 / Theano / recurrentshop stack absent and not installable offline; the CPU oracle restatement is timed (kind=port)"},
        "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------ GPU legs
class Dist:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0")); self.world = int(os.environ.get("WORLD_SIZE", "1")); self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def make_engine(D, T, H, L, B, feedback, precision, rnn_mode, cell="LSTM"):
    from midi_vae_b200 import Engine, EngineConfig, initial_weights, nccl_unique_id
    cfg = EngineConfig(input_length=T, lstm_size=H, latent_rep_size=L, decoder_feedback=feedback, precision=precision, rnn_mode=rnn_mode, max_batch=B, cell_type=cell)
    eng = Engine(cfg, D.local)
    eng.set_weights(initial_weights(cfg, 42))
    if D.world > 1:
        ids = [nccl_unique_id() if D.rank == 0 else None]
        D.dist.broadcast_object_list(ids, src=0)
        eng.nccl_init(ids[0], D.world, D.rank)
    return eng


def ncu_traffic(dom, workload, T, B, H):
    """DRAM bytes per launch of the dominant recurrence kernel.  NOT measured in this run: read from the committed `ncu --set full` capture of the
    same kernel on the same workload (a run under ncu is never a bench value, so the two cannot be the same process)."""
    for rel in (("profiles/r2/ncu_rec_cluster_cfg3.json", "profiles/r1/ncu_rec_cluster_cfg3_r1h.json") if dom == "rec_bwd" else
                ("profiles/r2/ncu_rec_cluster_cfg3.json", "profiles/r1/ncu_rec_cluster_cfg3.json")):
        f = os.path.join(ROOT, rel)
        if workload != "cfg3" or not os.path.exists(f):
            continue
        pat = "rec_cluster_bwd" if dom == "rec_bwd" else "rec_cluster_fwd"
        try:
            ks = [k for k in json.load(open(f))["kernels"] if pat in k["Kernel Name"] and k["gpu__time_duration.sum"] > 0.3]   # the T=256 launches
        except Exception:
            continue
        if ks:
            traffic = sum(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"] for k in ks) / len(ks) * 1e6      # bytes per launch
            return traffic, {"measured_in_this_run": False, "source": f"from profile: {rel} (ncu --set full, mean over the captured T={T} launches)",
                             "launches_per_step": 6, "algorithmic_stash_bytes_per_launch": T * B * H * 2 * (5 + 1 + 4),
                             "note": "algorithmic bytes = stash read (gates 4H + c H per step) + dh_ext H + dG written 4H, all bf16"}
    return None, None


def run_train(args, name, wl, D):
    torch = D.torch
    from midi_vae_b200 import synth
    rank, world = D.rank, D.world
    T, H, L, B = wl["T"], wl["H"], wl["L"], wl["B"]
    if "feedback" in wl:
        args.feedback = wl["feedback"]
    cell = wl.get("cell", "LSTM")
    eng = make_engine(D, T, H, L, B, args.feedback, args.precision, args.rnn_mode, cell)

    # synthetic rolls: NB distinct batches per rank, resident in HBM (weak scaling: B sequences per GPU)
    NB = 4
    stream = torch.cuda.ExternalStream(eng.stream())
    dev, host = [], []
    for i in range(NB):
        r = synth.make_batch(B, T, seed=1234 + 2 + 1000 * rank + i)
        eps = synth.make_eps(B, L, 1234 + i)
        hist = (np.random.default_rng(99 + i).standard_normal((B, L)) * 0.1).astype(np.float32)
        host.append((r, hist, eps))
        t = dict(pitch=torch.from_numpy(r.pitch).cuda(), instr=torch.from_numpy(r.instr).cuda(), vel=torch.from_numpy(r.velocity).cuda(),
                 style=torch.from_numpy(r.style).cuda(), hist=torch.from_numpy(hist).cuda(), eps=torch.from_numpy(eps).cuda())
        dev.append((t, eng.device_batch(B, t["pitch"].data_ptr(), t["instr"].data_ptr(), t["vel"].data_ptr(), t["style"].data_ptr(),
                                        t["hist"].data_ptr(), t["eps"].data_ptr())))
    metrics_dev = torch.zeros(10, device="cuda")
    torch.cuda.synchronize()

    # ---- device-resident timing
    for i in range(args.warmup):
        eng.train_step_device(dev[i % NB][1], metrics_dev.data_ptr())
    # the clock sampler (an nvidia-smi child process per rank) starts BEFORE the barrier: launched after it, its start-up time -- tens of milliseconds,
    # different on every rank -- sat inside the timed region of whichever rank came out first (that rank then waits in its first all-reduce)
    sampler = ClockSampler(D.local); sampler.start(); time.sleep(0.25)
    D.barrier()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for i in range(args.steps):
            eng.train_step_device(dev[i % NB][1], metrics_dev.data_ptr())
        e1.record(stream)
    D.barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1)
    ms_total = D.max_over_ranks(e0.elapsed_time(e1))
    launches = eng.launch_count() - l0
    ms_step = ms_total / args.steps
    value = world * B / (ms_step / 1e3)
    loss_last = float(metrics_dev[0].item())

    # ---- end-to-end: public host API, H2D of the step's rolls + D2H of its metrics inside the timed region
    e2e = None
    if not args.no_e2e:
        for i in range(2):
            r, hist, eps = host[i % NB]
            eng.train_on_batch(r.pitch, r.instr, r.velocity, r.style, hist, eps)
        D.barrier()
        eng.transfer_bytes(reset=True)
        k2 = max(3, args.steps // 2)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            f0.record(stream)
            for i in range(k2):
                r, hist, eps = host[i % NB]
                eng.train_on_batch(r.pitch, r.instr, r.velocity, r.style, hist, eps)
            f1.record(stream)
        D.barrier()
        ms_e2e = D.max_over_ranks(f0.elapsed_time(f1)) / k2
        h2d, d2h = eng.transfer_bytes()
        e2e = {"value": world * B / (ms_e2e / 1e3), "unit": "sequences/s", "h2d_bytes_per_step": h2d // k2, "d2h_bytes_per_step": d2h // k2,
               "ms_per_step": ms_e2e, "api": "Engine.train_on_batch (mvae_train_step_host), host numpy rolls"}

    # ---- per-kernel-class CUDA-event timing of one more step (every rank runs it -- it contains the all-reduce -- rank 0 reports)
    roof = None
    eng.set_profiling(True)
    eng.train_step_device(dev[0][1], metrics_dev.data_ptr())
    eng.sync()
    kms = eng.kernel_ms()
    eng.set_profiling(False)
    if rank == 0:
        pk = peaks()
        fwd, rec_fwd = flops_per_seq(T, H, L, args.feedback, cell=cell)
        train = 3 * fwd
        fast = args.rnn_mode != "streamed" and args.precision == "bf16" and cell == "LSTM"
        gen = "cluster-resident" if H in (256, 512) else "persistent"
        gru_fast = args.rnn_mode != "streamed" and args.precision == "bf16" and cell == "GRU" and H == 256
        classes = {
            "rec_bwd": ((("rec_cluster_bwd4_kernel" if H == 512 else "rec_cluster_bwd_kernel" if H == 256 else "rec_persist_kernel<bwd>") + f": {gen} backward recurrence "
                         "(dG*U^T per step on tcgen05 + gate-gradient math)") if fast else
                        "gru_cluster_bwd_kernel: cluster-resident GRU reverse sweep (two dependent products per step, mma.sync tiles, DSMEM exchanges)" if gru_fast
                        else "step-streamed backward recurrence", rec_fwd * B),
            "rec_fwd": ((("rec_cluster_fwd2_kernel" if H in (256, 512) else "rec_persist_kernel<fwd>") + f": {gen} forward recurrence (h*U per step on tcgen05 + gate math)")
                        if fast else "gru_cluster_fwd_kernel: cluster-resident GRU recurrence (two dependent products per step, mma.sync tiles, DSMEM exchanges)"
                        if gru_fast else "step-streamed forward recurrence", rec_fwd * B),
            "gemm": ("gemm_tc_kernel: batched tcgen05 GEMMs (input projections, heads, weight gradients)", (train - 2 * rec_fwd) * B),
        }
        dom = max(classes, key=lambda k: kms[k][0])
        dom_ms = kms[dom][0]
        peak = pk["bf16_sustained"]
        achieved = classes[dom][1] / (dom_ms / 1e3) / 1e12
        traffic, traffic_detail = ncu_traffic(dom, name, T, B, H) if dom in ("rec_bwd", "rec_fwd") else (None, None)
        roof = {"bound": "tensor", "kernel": classes[dom][0], "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_detail": traffic_detail, "peak_source": f"{pk['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                "kernel_ms_per_step": dom_ms, "kernel_flops_per_step": classes[dom][1],
                "note": "kernel classes overlap on three streams; kernel_ms_per_step is the sum of the class's launch durations (CUDA events on the launching streams)",
                "step": {"achieved": value / world * train / 1e12, "frac": value / world * train / 1e12 / peak,
                         "frac_of_burst": value / world * train / 1e12 / pk["bf16_burst"], "train_flops_per_seq": train},
                "class_ms": {k: round(v[0], 4) for k, v in kms.items()}, "class_launch_groups": {k: v[1] for k, v in kms.items()}}
    eng.close()

    # ---- cfg5: the persistent-vs-streamed comparison north_star asks for (same workload, step-streamed recurrences)
    sweep = None
    if name == "cfg5" and world == 1 and args.rnn_mode == "auto" and args.precision == "bf16":
        e2 = make_engine(D, T, H, L, B, args.feedback, args.precision, "streamed")
        b2 = dev[0][1]
        for _ in range(1):
            e2.train_step_device(b2, metrics_dev.data_ptr())
        e2.sync()
        s2 = torch.cuda.ExternalStream(e2.stream())
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ks = max(1, min(3, args.steps))
        with torch.cuda.stream(s2):
            g0.record(s2)
            for _ in range(ks):
                e2.train_step_device(b2, metrics_dev.data_ptr())
            g1.record(s2)
        e2.sync()
        ms2 = g0.elapsed_time(g1) / ks
        sweep = {"persistent_ms_per_step": ms_step, "streamed_ms_per_step": ms2, "persistent_over_streamed": ms2 / ms_step, "streamed_steps_timed": ks}
        e2.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:      # the CPU baseline is timed at N = 1 only (the other ranks would idle in a barrier)
        big = T * H >= 2 ** 19
        cpu, _ = cpu_leg(args, wl, args.feedback, 3 if big else 10, 1)

    if world > 1:
        D.dist.barrier()
    if rank == 0:
        line = {
            "metric": METRIC_TRAIN, "value": value, "unit": "sequences/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": workload_string(name, wl, args.feedback), "global_batch": B * world, "parallelism": f"dp{world}", "rnn_mode": args.rnn_mode,
                       "l2": "per-step working set (BPTT stash, several GB) >> 126 MB L2; 4 rotating input batches"},
            "loss_last_step": loss_last, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu}
        if sweep:
            line["rnn_crossover"] = sweep
        print(json.dumps(line))


def run_infer(args, name, wl, D):
    """cfg4: latency / throughput of one batch-1024 style-transfer call."""
    torch = D.torch
    from midi_vae_b200 import synth
    rank, world = D.rank, D.world
    T, H, L, B = wl["T"], wl["H"], wl["L"], wl["B"]
    fb = args.infer_feedback
    eng = make_engine(D, T, H, L, B, "as_wired", args.precision, args.rnn_mode)
    NB = 2
    stream = torch.cuda.ExternalStream(eng.stream())
    dev, host = [], []
    for i in range(NB):
        songs = synth.concat(synth.make_songs(16, T, seed=5 + 1000 * rank + i, min_chunks=64, max_chunks=64))
        ss = songs.song_start.astype(np.uint8)
        host.append((songs, ss))
        t = dict(pitch=torch.from_numpy(songs.pitch).cuda(), instr=torch.from_numpy(songs.instr).cuda(), vel=torch.from_numpy(songs.velocity).cuda(),
                 ss=torch.from_numpy(ss).cuda())
        dev.append((t, eng.device_batch(B, t["pitch"].data_ptr(), t["instr"].data_ptr(), t["vel"].data_ptr())))
    op, oi, ov = torch.empty(B, T, dtype=torch.uint8, device="cuda"), torch.empty(B, 4, dtype=torch.uint8, device="cuda"), torch.empty(B, T, device="cuda")

    def call(i):
        t, b = dev[i % NB]
        eng.style_transfer_device(b, t["ss"].data_ptr(), 0, 1, fb, op.data_ptr(), oi.data_ptr(), ov.data_ptr())

    for i in range(max(3, args.warmup)):
        call(i)
    sampler = ClockSampler(D.local); sampler.start(); time.sleep(0.25)   # before the barrier: see run_train
    D.barrier()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for i in range(args.steps):
            call(i)
        e1.record(stream)
    D.barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1)
    launches = eng.launch_count() - l0
    ms_step = D.max_over_ranks(e0.elapsed_time(e1)) / args.steps
    value = world * B / (ms_step / 1e3)
    # per-call latency distribution (one call in flight at a time)
    lat = []
    for i in range(args.latency_runs):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            a.record(stream); call(i); b.record(stream)
        eng.sync()
        lat.append(a.elapsed_time(b))
    lat = np.array(lat)
    # e2e: host rolls in, host indices out
    e2e = None
    if not args.no_e2e:
        for i in range(2):
            s, ss = host[i % NB]
            eng.style_transfer(s.pitch, s.instr, s.velocity, 0, 1, ss, fb)
        D.barrier()
        eng.transfer_bytes(reset=True)
        k2 = max(3, args.steps // 2)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            f0.record(stream)
            for i in range(k2):
                s, ss = host[i % NB]
                eng.style_transfer(s.pitch, s.instr, s.velocity, 0, 1, ss, fb)
            f1.record(stream)
        D.barrier()
        ms_e2e = D.max_over_ranks(f0.elapsed_time(f1)) / k2
        h2d, d2h = eng.transfer_bytes()
        e2e = {"value": world * B / (ms_e2e / 1e3), "unit": "sequences/s", "h2d_bytes_per_step": h2d // k2, "d2h_bytes_per_step": d2h // k2,
               "ms_per_step": ms_e2e, "api": "Engine.style_transfer (mvae_style_transfer_host), host numpy rolls in, host u8 indices + f32 velocities out"}
    eng.set_profiling(True); call(0); eng.sync(); kms = eng.kernel_ms(); eng.set_profiling(False)
    eng.close()
    roof = None
    if rank == 0:
        pk = peaks()
        fwd, rec_fwd = flops_per_seq(T, H, L, "as_wired" if fb == "as_wired" else "teacher_forced")
        peak = pk["bf16_sustained"]
        dom_ms = kms["rec_fwd"][0]
        achieved = rec_fwd * B / (dom_ms / 1e3) / 1e12
        roof = {"bound": "tensor", "kernel": "rec_cluster_fwd2_kernel (no BPTT stash): cluster-resident forward recurrence" if fb == "as_wired" else "step-streamed free-running decode",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                "peak_source": f"{pk['source']} bf16_tflops_sustained", "kernel_ms_per_step": dom_ms, "kernel_flops_per_step": rec_fwd * B,
                "step": {"achieved": value / world * fwd / 1e12, "frac": value / world * fwd / 1e12 / peak, "forward_flops_per_seq": fwd},
                "class_ms": {k: round(v[0], 4) for k, v in kms.items()}}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_leg(args, wl, fb, 3, 1)
    if world > 1:
        D.dist.barrier()
    if rank == 0:
        print(json.dumps({
            "metric": METRIC_INFER, "value": value, "unit": "sequences/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": workload_string(name, wl, fb), "global_batch": B * world, "parallelism": f"dp{world} (independent replicas, no collective)",
                       "l2": "2 rotating input batches; the forward working set (h sequences of 6 recurrences, several GB at T256/H512) >> 126 MB L2"},
            "latency_ms": {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)), "runs": int(len(lat))},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=list(WORKLOADS))
    ap.add_argument("--cfg4-shape", default="cfg3", choices=["cfg2", "cfg3"], help="layer sizes of the cfg4 inference workload (BASELINE fixes only batch = 1024)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--feedback", default="teacher_forced", choices=["teacher_forced", "as_wired"])
    ap.add_argument("--infer-feedback", default="as_wired", choices=["as_wired", "free_running"])
    ap.add_argument("--rnn-mode", default="auto", choices=["auto", "streamed", "persistent"])
    ap.add_argument("--cpu-sample", type=int, default=8, help="sequences in the CPU arms' sample of the batch")
    ap.add_argument("--latency-runs", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.workload == "cfg4" and args.cfg4_shape == "cfg2":
        wl.update(T=64, H=256, L=100)
    if args.workload == "cfg5" and args.steps > 5 and "--steps" not in " ".join(sys.argv):
        args.steps, args.warmup = 3, 3          # a cfg5 step is ~0.25 s
    if "feedback" in wl:
        args.feedback = wl["feedback"]
    if args.impl == "reference":
        run_reference(args, args.workload, wl, args.infer_feedback if wl["kind"] == "infer" else args.feedback)
        return
    D = Dist(args)
    (run_infer if wl["kind"] == "infer" else run_train)(args, args.workload, wl, D)
    D.close()


if __name__ == "__main__":
    main()
