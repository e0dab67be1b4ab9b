#!/usr/bin/env python
"""bench.py -- MIDI sequences/sec of one MIDI-VAE train step (fwd + bwd + [all-reduce] + Adam) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|cfg1] [--precision bf16|fp32]
    python bench.py --impl reference ...      # the CPU oracle (the reference's Keras path cannot run here) on the host cores

Contract (one JSON line on stdout from rank 0): metric/value/unit/n_gpus/steps/warmup/ms_per_step/higher_is_better/
scaling/vs_baseline/dtype/data/config/clocks/e2e/gpu_launches/roofline/cpu_baseline.  A step is one pass of the
hot path (vae_training.py:804-809: one mini-batch of autoencoder.fit) over one batch of synthetic rolls of the
BASELINE.json shape; `value` is timed with inputs resident in HBM, `e2e` through the public host-pointer API with
the H2D copies of the step's inputs and the D2H read of its metrics inside the timed region.
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
    "cfg1": dict(T=16, H=64, L=16, B=8),
    "cfg2": dict(T=64, H=256, L=100, B=128),
    "cfg3": dict(T=256, H=512, L=256, B=512),
}
METRIC = "MIDI sequences/sec (train step)"


def flops_per_seq(T, H, L, feedback="teacher_forced", ne=2, nd=2, Dp=61, Di=16, Ti=4, Dv=1):
    """Algorithmic GEMM FLOPs per sequence, forward (SURVEY.md 8(d) formula).  Returns (total_fwd, recurrent_fwd)."""
    G = 4 * H
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
        return dict(bf16_sustained=d.get("bf16_tflops_sustained"), bf16_burst=d.get("bf16_tflops"), hbm=d.get("hbm_gbs"), source="measured")
    return dict(bf16_sustained=1400.0, bf16_burst=1590.0, hbm=6650.0, source="fallback")


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


# ------------------------------------------------------------------------------------------------ CPU oracle leg
def cpu_oracle_rate(wl, feedback, sample_batch, steps, warmup):
    """Reference CPU path = the fp32 PyTorch-CPU oracle, all host threads, bounded sample."""
    import torch
    from midi_vae_b200 import EngineConfig, initial_weights, synth
    from oracle import midivae_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    ocfg = O.OracleConfig(input_length=wl["T"], lstm_size=wl["H"], latent_rep_size=wl["L"], decoder_feedback=feedback)
    ecfg = EngineConfig(input_length=wl["T"], lstm_size=wl["H"], latent_rep_size=wl["L"])
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


def run_reference(args, wl, feedback):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = max(1, min(wl["B"], args.cpu_sample))
    rate, ms, cores = cpu_oracle_rate(wl, feedback, sample, max(1, args.steps), max(0, min(args.warmup, 1)))
    cb = {"value": rate, "unit": "sequences/s", "cores": cores, "kind": "port",
          "sample": f"oracle fp32 torch-CPU train step on {sample} of {wl['B']} sequences of the workload, median of {max(1, args.steps)} steps"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "sequences/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: seq_len={wl['T']} hidden={wl['H']} latent={wl['L']} batch={wl['B']}/GPU, decoder_feedback={feedback}, LSTM",
                   "note": "reference Keras/Theano stack absent; CPU oracle restatement timed (kind=port)"},
        "cpu_baseline": cb, "e2e": {"value": rate, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------ GPU leg
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=list(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--feedback", default="teacher_forced", choices=["teacher_forced", "as_wired"])
    ap.add_argument("--rnn-mode", default="auto", choices=["auto", "streamed", "persistent"])
    ap.add_argument("--cpu-sample", type=int, default=8, help="sequences in the CPU-baseline sample batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl, args.feedback)
        return

    import torch
    import torch.distributed as dist
    from midi_vae_b200 import Engine, EngineConfig, initial_weights, nccl_unique_id, synth

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    T, H, L, B = wl["T"], wl["H"], wl["L"], wl["B"]
    cfg = EngineConfig(input_length=T, lstm_size=H, latent_rep_size=L, decoder_feedback=args.feedback, precision=args.precision,
                       rnn_mode=args.rnn_mode, max_batch=B)
    eng = Engine(cfg, local)
    eng.set_weights(initial_weights(cfg, 42))
    if world > 1:
        ids = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.nccl_init(ids[0], world, rank)

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

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing
    for i in range(args.warmup):
        eng.train_step_device(dev[i % NB][1], metrics_dev.data_ptr())
    barrier()
    sampler = ClockSampler(local); sampler.start(); time.sleep(0.25)
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for i in range(args.steps):
            eng.train_step_device(dev[i % NB][1], metrics_dev.data_ptr())
        e1.record(stream)
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1)
    ms_total = max_over_ranks(e0.elapsed_time(e1))
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
        barrier()
        eng.transfer_bytes(reset=True)
        k2 = max(3, args.steps // 2)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            f0.record(stream)
            for i in range(k2):
                r, hist, eps = host[i % NB]
                eng.train_on_batch(r.pitch, r.instr, r.velocity, r.style, hist, eps)
            f1.record(stream)
        barrier()
        ms_e2e = max_over_ranks(f0.elapsed_time(f1)) / k2
        h2d, d2h = eng.transfer_bytes()
        e2e = {"value": world * B / (ms_e2e / 1e3), "unit": "sequences/s", "h2d_bytes_per_step": h2d // k2, "d2h_bytes_per_step": d2h // k2,
               "ms_per_step": ms_e2e, "api": "Engine.train_on_batch (mvae_train_step_host), host numpy rolls"}

    # ---- per-kernel-class CUDA-event timing of one more step (rank 0), for the roofline of the dominant kernel
    # (every rank runs the step -- it contains the all-reduce -- rank 0 reports)
    roof = None
    eng.set_profiling(True)
    eng.train_step_device(dev[0][1], metrics_dev.data_ptr())
    eng.sync()
    kms = eng.kernel_ms()
    eng.set_profiling(False)
    if rank == 0:
        pk = peaks()
        fwd, rec_fwd = flops_per_seq(T, H, L, args.feedback)
        train = 3 * fwd
        # kernel classes and their algorithmic FLOPs per step (B sequences): the forward recurrences do h*U, the backward
        # recurrences dG*U^T (same count), everything else (input projections, heads, all weight gradients) is batched GEMM
        classes = {
            "rec_bwd": ((("rec_cluster_bwd4_kernel (H=512) / rec_cluster_bwd_kernel (H=256): cluster-resident" if H in (256, 512) else "persistent") + " backward recurrence (dG*U^T per step on tcgen05 + gate-gradient math)")
                        if args.rnn_mode != "streamed" and args.precision == "bf16" else "step-streamed backward recurrence", rec_fwd * B),
            "rec_fwd": ((("rec_cluster_fwd2_kernel: cluster-resident" if H in (256, 512) else "persistent") + " forward recurrence (h*U per step on tcgen05 + gate math)")
                        if args.rnn_mode != "streamed" and args.precision == "bf16" else "step-streamed forward recurrence", rec_fwd * B),
            "gemm": ("batched tcgen05 GEMMs (input projections, heads, weight gradients)", (train - 2 * rec_fwd) * B),
        }
        dom = max(classes, key=lambda k: kms[k][0])
        dom_ms = kms[dom][0]
        peak = pk["bf16_sustained"]
        achieved = classes[dom][1] / (dom_ms / 1e3) / 1e12
        traffic, traffic_detail = None, None
        ncu_file = os.path.join(ROOT, "profiles", "r1", "ncu_rec_cluster_cfg3_r1h.json" if dom == "rec_bwd" else "ncu_rec_cluster_cfg3.json")
        if dom in ("rec_bwd", "rec_fwd") and args.workload == "cfg3" and os.path.exists(ncu_file):
            pat = "rec_cluster_bwd" if dom == "rec_bwd" else "rec_cluster_fwd"
            ks = [k for k in json.load(open(ncu_file))["kernels"] if pat in k["Kernel Name"] and k["gpu__time_duration.sum"] > 0.3]   # the T=256 launches
            if ks:
                traffic = sum(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"] for k in ks) / len(ks) * 1e6      # bytes per launch
                traffic_detail = {"launches_per_step": 6, "algorithmic_stash_bytes_per_launch": T * B * H * 2 * (5 + 1 + 4),
                                  "source": f"profiles/r1/{os.path.basename(ncu_file)} (ncu --set full, mean over the captured T={T} launches; the stash read "
                                            "(gates 4H + c H per step), dh_ext H and the dG written 4H, all bf16, are the algorithmic bytes)"}
        roof = {"bound": "tensor", "kernel": classes[dom][0], "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_detail": traffic_detail, "peak_source": f"{pk['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                "kernel_ms_per_step": dom_ms, "kernel_flops_per_step": classes[dom][1],
                "step": {"achieved": value / world * train / 1e12, "frac": value / world * train / 1e12 / peak,
                         "frac_of_burst": value / world * train / 1e12 / pk["bf16_burst"], "train_flops_per_seq": train},
                "class_ms": {k: round(v[0], 4) for k, v in kms.items()}, "class_launch_groups": {k: v[1] for k, v in kms.items()}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:      # the CPU baseline is timed at N = 1 only (the other ranks would idle in a barrier)
        sample = max(1, min(B, args.cpu_sample))
        rate, ms, cores = cpu_oracle_rate(wl, args.feedback, sample, 10, 1)       # ~10 s of CPU work at cfg3
        cpu = {"value": rate, "unit": "sequences/s", "cores": cores, "kind": "port",
               "sample": f"oracle fp32 torch-CPU train step on {sample} of {B} sequences of the workload, median of 10 steps after 1 warm-up ({ms:.0f} ms/step)"}

    if world > 1:
        dist.barrier()
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "sequences/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: seq_len={T} hidden={H} latent={L} batch={B}/GPU (global {B * world}), decoder_feedback={args.feedback}, "
                                   f"LSTM, 2+2 layers, hard_sigmoid gates, rnn_mode={args.rnn_mode}",
                       "parallelism": f"dp{world}", "l2": "per-step working set (BPTT stash, several GB) >> 126 MB L2; 4 rotating input batches",
                       "loss_last_step": loss_last},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu}))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
