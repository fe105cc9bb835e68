#!/usr/bin/env python
"""bench.py — measures the lidbox_b200 hot path on B200 (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload logmel|xvector_train|...] [--impl reference]

One JSON line on stdout (rank 0).  `value` is device-resident throughput, `e2e` goes through the public Python
API with pinned HOST buffers (H2D + D2H inside the timed region), `roofline` describes the dominant kernel
(CUDA-event timed on the launching stream), `cpu_baseline` times the CPU oracle on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 16000
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            out = dict(FALLBACK_PEAKS)
            out.update({k: float(v) for k, v in d.items() if isinstance(v, (int, float))})
            out["source"] = "measured"
            return out
        except Exception:
            pass
    out = dict(FALLBACK_PEAKS)
    out["source"] = "fallback"
    return out


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_signals(B, N, seed, device=None, pin=False):
    """SURVEY §8(d) signal law: 0.5 sin(2 pi f_b t) + 0.05 N(0,1), f_b ~ U(100, 4000), seeded."""
    g = torch.Generator().manual_seed(seed)
    f = 100.0 + 3900.0 * torch.rand(B, 1, generator=g)
    t = torch.arange(N, dtype=torch.float32) / SR
    x = 0.5 * torch.sin(2 * np.pi * f * t) + 0.05 * torch.randn(B, N, generator=g)
    if pin:
        x = x.pin_memory()
    if device is not None:
        x = x.to(device)
    return x


# ---------------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------------
class LogmelWorkload:
    """STFT -> log-mel of B x sec utterances (BASELINE config 5 maximum: 2048 x 5 s)."""
    name = "logmel"
    metric = "log-mel frames/s (16 kHz, 25/10 ms, 512-pt FFT, 40 mel)"
    unit = "frames/s"
    dtype = "f32"

    def __init__(self, args, rank, world):
        self.B, self.sec = args.batch or 2048, args.seconds or 5
        self.N = self.sec * SR
        self.T = 1 + (self.N - 400) // 160
        self.rank, self.world = rank, world

    def config(self):
        return {"workload": "logmel %dx%ds per GPU (BASELINE config 5 max), inputs %.0f MB > L2 (no flush needed)"
                % (self.B, self.sec, self.B * self.N * 4 / 1e6), "batch_per_gpu": self.B, "seconds": self.sec,
                "frames_per_utt": self.T, "parallelism": "dp%d (independent shards, no collective)" % self.world}

    def setup(self, device):
        from lidbox_b200.features import audio
        self.audio = audio
        self.x_host = synth_signals(self.B, self.N, 1234 + self.rank, pin=True)
        self.x = self.x_host.to(device)
        self.out = torch.empty((self.B, self.T, 40), dtype=torch.float32, device=device)
        self.out_host = torch.empty((self.B, self.T, 40), dtype=torch.float32).pin_memory()

    def units_per_step(self):
        return self.B * self.T

    def step(self):
        self.audio.logmelspectrograms(self.x, SR, out=self.out)

    def launches_per_step(self):
        return 1

    def step_e2e(self):
        x = self.x_host.to(self.x.device, non_blocking=True)
        out = self.audio.logmelspectrograms(x, SR, out=self.out)
        self.out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def e2e_bytes(self):
        return self.B * self.N * 4, self.B * self.T * 40 * 4

    def roofline(self, ms_per_step, peaks):
        alg = self.B * (4 * self.N + 4 * self.T * 40)
        ach = alg / (ms_per_step * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": "logmel512_kernel<1>", "achieved": ach, "peak": peaks["hbm_gbs"],
                "peak_source": peaks["source"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                "algorithmic_bytes_per_launch": alg, "traffic": None}

    def cpu_sample(self, budget_s=15.0):
        from oracle import lidbox_oracle as O
        torch.set_num_threads(os.cpu_count())
        Bs = 64
        x = synth_signals(Bs, self.N, 99)
        O.torch_logmel(x)
        n, t0 = 0, time.perf_counter()
        while True:
            O.torch_logmel(x)
            n += 1
            dt = time.perf_counter() - t0
            if dt > budget_s or n >= 50:
                break
        return {"value": n * Bs * self.T / dt, "unit": self.unit, "cores": os.cpu_count(), "kind": "port",
                "sample": "%d iterations of %dx%ds log-mel through oracle.torch_logmel (fp32 torch-CPU restatement; "
                          "TensorFlow is not installable)" % (n, Bs, self.sec)}


WORKLOADS = {"logmel": LogmelWorkload}
try:
    from lidbox_b200.bench_workloads import EXTRA_WORKLOADS   # x-vector workloads register themselves here
    WORKLOADS.update(EXTRA_WORKLOADS)
except ImportError:
    pass
DEFAULT_WORKLOAD = os.environ.get("LBX_BENCH_WORKLOAD", "xvector_train" if "xvector_train" in WORKLOADS else "logmel")


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path (here: the oracle port, TensorFlow is not installable)."""
    if rank != 0:
        return
    wl = WORKLOADS[args.workload](args, 0, 1)
    per = max(2.0, 60.0 / max(1, args.steps + args.warmup))
    for _ in range(args.warmup):
        wl.cpu_sample(budget_s=min(per, 3.0))
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals.append(wl.cpu_sample(budget_s=per))
    dt = time.perf_counter() - t0
    cb = vals[-1]
    cb["value"] = float(np.mean([v["value"] for v in vals]))
    line = {"impl": "reference", "metric": wl.metric, "value": cb["value"], "unit": wl.unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": wl.config(), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="lidbox_b200", choices=["lidbox_b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--seconds", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: lidbox_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    from lidbox_b200 import _lib
    lib = _lib.lib()
    peaks = load_peaks()
    wl = WORKLOADS[args.workload](args, rank, world)
    wl.dist = dist
    wl.setup(device)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        wl.step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = lib.lbx_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        wl.step()
    ev1.record()
    barrier()
    launches = lib.lbx_launch_count() - n0
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps

    # dominant-kernel timing for the roofline (CUDA events on the launching stream)
    roof = wl.roofline_measure(peaks) if hasattr(wl, "roofline_measure") else wl.roofline(ms_per_step, peaks)

    # end-to-end through the public API with pinned host buffers
    for _ in range(2):
        wl.step_e2e()
    barrier()
    e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e_steps):
        wl.step_e2e()
    e1.record()
    barrier()
    e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3 * 0.0)
    if dist is not None:
        t = torch.tensor([e_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_ms = float(t.item())
    h2d, d2h = wl.e2e_bytes()

    if rank == 0:
        units = wl.units_per_step() * world
        line = {"metric": wl.metric, "value": units / (ms_per_step * 1e-3), "unit": wl.unit, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
                "config": wl.config(), "roofline": roof,
                "e2e": {"value": units / (e_ms / e_steps * 1e-3), "unit": wl.unit, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": e_ms / e_steps},
                "gpu_launches": int(launches), "clocks": clocks}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = wl.cpu_sample()
        if hasattr(wl, "extra"):
            line.update(wl.extra())
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
