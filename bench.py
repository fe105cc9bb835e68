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
    """Samples SM clock and throttle reasons through NVML every ~5 ms during the timed region (nvidia-smi's own
    polling loop is too coarse for regions of a few tens of milliseconds)."""

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self.thread = None
        self.err = None

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                     "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                     "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}
            while not self._stop.is_set():
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(0.005)
        except Exception as e:          # never take the benchmark down
            self.err = repr(e)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.err:
            out["error"] = self.err
        return out


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


def tdnn_forward_flops(T, n_out=4, F=40):
    """Algorithmic forward FLOPs of one utterance: frames 1-5 + segments (+ output layer)."""
    T2 = -(-T // 2)
    T3 = -(-T2 // 3)
    frame1 = 2 * T * 5 * F * 512
    rest = 2 * (T2 * 1536 * 512 + T3 * 1536 * 512 + T3 * 512 * 512 + T3 * 512 * 1500)
    dense = 2 * (3000 * 512 + 512 * 512 + 512 * n_out)
    return frame1 + rest + dense, frame1


def class_signals(B, N, n_classes, seed, pin=False):
    """SURVEY §8(d) cfg3/cfg4 law: label y_b = b mod C, tone frequency drawn from C disjoint bands of [100, 4000] Hz."""
    g = torch.Generator().manual_seed(seed)
    y = torch.arange(B) % n_classes
    width = 3900.0 / n_classes
    f = 100.0 + (y.float() + torch.rand(B, generator=g)) * width
    t = torch.arange(N, dtype=torch.float32) / SR
    x = 0.5 * torch.sin(2 * np.pi * f[:, None] * t) + 0.05 * torch.randn(B, N, generator=g)
    if pin:
        x = x.pin_memory()
    return x, y.to(torch.int32)


class XVectorTrainWorkload:
    """BASELINE config 3: x-vector training, 4 synthetic language labels, cross-entropy, bf16 (fp32 master weights,
    statistics, loss), Adam; per-GPU batch 256 x 2 s; data-parallel with one NCCL all-reduce of the flat gradient.
    A step = log-mel of the resident signals -> forward -> backward -> all-reduce -> Adam + bf16 weight refresh."""
    name = "xvector_train"
    metric = "x-vector training audio-sec/s (log-mel + TDNN fwd/bwd + Adam)"
    unit = "audio-sec/s"
    dtype = "bf16"
    n_classes, n_out, loss, head = 4, 4, "xent", "log_softmax"
    default_seconds = 2
    label = "BASELINE config 3"

    def __init__(self, args, rank, world):
        self.B, self.sec = args.batch or 256, args.seconds or self.default_seconds
        self.N = self.sec * SR
        self.T = 1 + (self.N - 400) // 160
        self.rank, self.world = rank, world
        self.use_graph = os.environ.get("LBX_BENCH_GRAPH", "1") != "0"
        self.pipelined = os.environ.get("LBX_BENCH_PIPELINE", "1") != "0"
        self.dist = None

    def config(self):
        return {"workload": "%s: x-vector train, batch %d x %d s per GPU, %d labels, %s, bf16, Adam; signals resident "
                            "in HBM; L2 flushed by the step itself (activations+gradients %.0f MB > 126 MB L2)"
                            % (self.label, self.B, self.sec, self.n_classes, self.loss, self._act_mb()),
                "batch_per_gpu": self.B, "global_batch": self.B * self.world, "seconds": self.sec,
                "frames_per_utt": self.T, "cuda_graph": self.use_graph,
                "feature_prefetch": "log-mel of batch i+1 runs on a second stream during step i (one log-mel + one "
                                    "training step per replay)" if self.pipelined else "inline",
                "parallelism": "dp%d" % self.world, "dp_exchange": getattr(self, "dp_exchange", None)}

    def _act_mb(self):
        return self.B * self.T * (512 * 2 * 2 + 256 * 2 * 2 + 2 * 1504 * 2 / 6 + 160) / 1e6

    def setup(self, device):
        from lidbox_b200.features import audio
        from lidbox_b200.models import xvector
        self.audio, self.device, self.xvector = audio, device, xvector
        self.x_host, y = class_signals(self.B, self.N, self.n_classes, 1234 + self.rank, pin=True)
        self.y = y.to(device)
        self.feats = torch.empty((self.B, self.T, 40), dtype=torch.float32, device=device)
        self.model = xvector.create((self.T, 40), self.n_out, precision="bf16", head=self.head, seed=0)
        self.model.configure_optimizer(lr=1e-3)
        self.pg = self.dist.group.WORLD if self.dist is not None else None
        self.dp_exchange = "none (single GPU)"
        if self.pg is not None:
            if os.environ.get("LBX_DP_SHARDED", "1") != "0":
                self.model.enable_sharded_optimizer(self.pg)
                nvls = bool(self.model._sharded.get("mc_grads"))
                self.dp_exchange = ("fused in the optimizer kernel: reduce-scatter by %s, Adam on the rank's shard, "
                                    "all-gather by %s (no NCCL call in the step)"
                                    % (("NVLS multimem.ld_reduce (in-switch sum)", "multimem.st")
                                       if nvls else ("NVLink peer loads", "peer stores")))
            else:
                self.dp_exchange = "one NCCL all-reduce of the flat fp32 gradient, then full-size Adam"
        self.kw = dict(ap_classes=self.n_classes) if self.loss == "ap" else {}
        self.pipelined = os.environ.get("LBX_BENCH_PIPELINE", "1") != "0"
        self.pipes = {}
        self.pipe = self._build_pipe(self.x_host)
        self.x = self.pipe["xs"][0]
        self.graphed = self.pipe["steps"][0] if self.use_graph else None

    def _build_pipe(self, x_host):
        """Two signal buffers (dtype of x_host: float32 or 16-bit PCM) and two feature buffers: while the model trains
        on the features of batch i, the log-mel of batch i+1 runs on a second stream (the prefetch a tf.data input
        pipeline does), and in the end-to-end path the H2D copy of batch i+2 runs on a copy stream.
        Every replay = one log-mel + one training step."""
        dev, audio, xvector = self.device, self.audio, self.xvector
        pipe = {"x_host": x_host, "xs": [x_host.to(dev), x_host.to(dev)], "i": 0, "steps": [], "e2e": None,
                "fbuf": [torch.empty((self.B, self.T, 40), dtype=torch.float32, device=dev) for _ in range(2)]}
        audio.logmelspectrograms(pipe["xs"][0], SR, out=pipe["fbuf"][0])          # prime the pipeline
        for k in range(2):
            nxt = (lambda k=k: audio.logmelspectrograms(pipe["xs"][1 - k], SR, out=pipe["fbuf"][1 - k]))
            inline = (lambda k=k: audio.logmelspectrograms(pipe["xs"][k], SR, out=pipe["fbuf"][k]))
            if self.use_graph:
                pipe["steps"].append(xvector.GraphedTrainStep(
                    self.model, pipe["fbuf"][k], self.y, loss=self.loss, process_group=self.pg,
                    pre=None if self.pipelined else inline, concurrent=nxt if self.pipelined else None, **self.kw))
            else:
                pipe["steps"].append(lambda inline=inline: self.model.train_step(inline(), self.y, loss=self.loss,
                                                                                 process_group=self.pg, **self.kw))
        return pipe

    def _features(self):
        return self.audio.logmelspectrograms(self.x, SR, out=self.feats)

    def _eager_step(self):
        return self.model.train_step(self._features(), self.y, loss=self.loss, process_group=self.pg, **self.kw)

    def units_per_step(self):
        return self.B * self.sec

    def step(self, pipe=None):
        pipe = pipe or self.pipe
        out = pipe["steps"][pipe["i"] % 2]()
        pipe["i"] += 1
        return out

    def launches_per_step(self):
        return self.graphed.kernels_per_step if self.graphed is not None else None

    def _setup_e2e(self, pipe):
        dev = self.device
        e = {"copy": torch.cuda.Stream(device=dev),
             "h2d_done": [torch.cuda.Event(), torch.cuda.Event()], "step_done": torch.cuda.Event(),
             "loss_host": [torch.empty((self.B,), dtype=torch.float32).pin_memory() for _ in range(2)]}
        cur = torch.cuda.current_stream(dev)
        torch.cuda.synchronize(dev)
        # prime: batch 0 -> xs[k0] + its features, batch 1 -> xs[1-k0]
        k0 = pipe["i"] % 2
        pipe["xs"][k0].copy_(pipe["x_host"], non_blocking=True)
        self.audio.logmelspectrograms(pipe["xs"][k0], SR, out=pipe["fbuf"][k0])
        with torch.cuda.stream(e["copy"]):
            e["copy"].wait_stream(cur)
            pipe["xs"][1 - k0].copy_(pipe["x_host"], non_blocking=True)
            e["h2d_done"][1 - k0].record(e["copy"])
        e["step_done"].record(cur)
        pipe["e2e"] = e

    def step_e2e(self, pipe=None):
        """One step of the public-API pipeline with the batch in pinned HOST memory: replay k trains on the features
        of batch i (buffer k) and extracts the features of batch i+1 from device buffer 1-k, while batch i+2 is copied
        host->device into buffer k on the copy stream; the per-sample losses of batch i are copied back."""
        pipe = pipe or self.pipe
        if pipe["e2e"] is None:
            self._setup_e2e(pipe)
        e = pipe["e2e"]
        k = pipe["i"] % 2
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(e["h2d_done"][1 - k])                  # signals of batch i+1 are on the device
        prev_done = e["step_done"]
        losses = self.step(pipe)
        e["step_done"] = torch.cuda.Event()
        e["step_done"].record(cur)
        with torch.cuda.stream(e["copy"]):
            e["copy"].wait_event(prev_done)                    # buffer k was last read by the previous replay
            pipe["xs"][k].copy_(pipe["x_host"], non_blocking=True)
            e["h2d_done"][k].record(e["copy"])
        e["loss_host"][k].copy_(losses, non_blocking=True)
        if k == 1:
            cur.synchronize()                                  # host reads the losses of the last two steps

    def e2e_pcm16(self, steps, barrier):
        """Same end-to-end pipeline fed with 16-bit PCM (the sample format of the WAV corpora; decoded inside the
        log-mel kernel exactly as read_wav does): half the host->device bytes.  Reported next to the float32 e2e."""
        pcm = torch.clamp(torch.round(self.x_host * 32768.0), -32768, 32767).to(torch.int16).pin_memory()
        pipe = self._build_pipe(pcm)
        for _ in range(4):
            self.step_e2e(pipe)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            self.step_e2e(pipe)
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / steps, self.B * self.N * 2

    def e2e_bytes(self):
        return self.B * self.N * 4, self.B * 4

    def roofline_measure(self, peaks):
        """Dominant kernel = gemm_bf16_kernel.  Every GEMM descriptor of one training step is recorded, the launches
        are replayed back-to-back from a CUDA graph (same operands, same order, nothing else in between) and timed
        with CUDA events on the launching stream; achieved = algorithmic training FLOPs of the step / that time."""
        from lidbox_b200 import ops
        fwd, f1 = tdnn_forward_flops(self.T, self.n_out)
        alg = self.B * (3 * fwd - f1)
        ops.GEMM_RECORD = []
        self._eager_step()
        torch.cuda.synchronize()
        rec, ops.GEMM_RECORD = ops.GEMM_RECORD, None
        issued = sum(2.0 * M * N * K * nt for (_d, _k, (M, N, K, nt, _l)) in rec)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            ops.replay(rec, self.device)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            ops.replay(rec, self.device)
        ms = _time_cuda(g.replay, 20)
        self.model.grads.zero_()               # the replays accumulated into the gradient buffer
        self.model._grads_clean = True
        # per-launch durations (eager, CUDA events around each launch) for the largest single launch
        ops.GEMM_TIMER = []
        self._eager_step()
        torch.cuda.synchronize()
        tim, ops.GEMM_TIMER = ops.GEMM_TIMER, None
        big = max(tim, key=lambda r: r[2] * r[3] * r[4])
        big_ms = big[0].elapsed_time(big[1])
        ach = alg / (ms * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        traffic = None
        try:        # dram__bytes_read+write per launch from the committed ncu --set full capture of the same step
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_gemm_traffic.json")))
            if tj.get("launches") == len(rec) and (self.B, self.sec) == (256, 2):
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        self._traffic = traffic
        return {"bound": "tensor", "kernel": "gemm_bf16_kernel (all %d launches of one training step, replayed "
                                             "back-to-back from a CUDA graph)" % len(rec),
                "achieved": ach, "peak": peak, "peak_source": peaks["source"] + " (sustained: timed inside a step)",
                "unit": "TFLOP/s", "frac": ach / peak, "algorithmic_flops_per_step": alg,
                "issued_flops_per_step": issued, "gemm_ms_per_step": ms, "launches": len(rec),
                "avg_launch_us": ms * 1e3 / len(rec),
                "largest_launch": {"M": big[2], "N": big[3], "K": big[4], "ms": big_ms,
                                   "tflops": 2.0 * big[2] * big[3] * big[4] / (big_ms * 1e-3) / 1e12},
                "algorithmic_flops_per_launch": alg / len(rec), "traffic": self._traffic}

    def cpu_sample(self, budget_s=20.0):
        from oracle import lidbox_oracle as O
        torch.set_num_threads(os.cpu_count())
        Bs = 32
        x, y = class_signals(Bs, self.N, self.n_classes, 99)
        params = {k: torch.tensor(v, requires_grad=True) for k, v in O.xvector_init(40, self.n_out, seed=0).items()}
        opt = torch.optim.Adam(params.values(), lr=1e-3, eps=1e-7)

        def one():
            feats = O.torch_logmel(x)
            opt.zero_grad()
            if self.loss == "ap":
                l = O.torch_ap_loss(y, O.torch_xvector_forward(params, feats, l2_normalize=True), self.n_classes)
            else:
                lp = O.torch_xvector_forward(params, feats)
                l = -lp[torch.arange(Bs), y.long()].mean()
            l.backward()
            opt.step()
        one()
        n, t0 = 0, time.perf_counter()
        while True:
            one()
            n += 1
            dt = time.perf_counter() - t0
            if dt > budget_s or n >= 200:
                break
        return {"value": n * Bs * self.sec / dt, "unit": self.unit, "cores": os.cpu_count(), "kind": "port",
                "sample": "%d training steps of batch %d x %d s through the torch-CPU fp32 restatement "
                          "(oracle.torch_logmel + torch_xvector_forward + autograd + Adam; TensorFlow is not "
                          "installable)" % (n, Bs, self.sec)}

    def extra(self):
        """Secondary BASELINE lines measured in the same run (device-timed, signals resident)."""
        out = {}
        try:
            out["also"] = {"logmel_2048x5s": bench_logmel_quick(self.device),
                           "config2_embed_64x2s_fp32": bench_embed_quick(self.device)}
        except Exception as e:      # secondary numbers must never take the headline down
            out["also"] = {"error": repr(e)}
        return out


class XVectorAPTrainWorkload(XVectorTrainWorkload):
    """BASELINE config 4: x-vector -> 64-d L2-normalised vector -> SparseAngularProximity(N=50, D=64), 3 s utterances."""
    name = "xvector_ap_train"
    metric = "x-vector + angular-proximity training audio-sec/s"
    n_classes, n_out, loss, head = 50, 64, "ap", "l2_normalize"
    default_seconds = 3
    label = "BASELINE config 4"


def _time_cuda(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_logmel_quick(device, B=2048, sec=5):
    from lidbox_b200.features import audio
    N = sec * SR
    T = 1 + (N - 400) // 160
    x = torch.randn((B, N), device=device) * 0.1
    out = torch.empty((B, T, 40), dtype=torch.float32, device=device)
    ms = _time_cuda(lambda: audio.logmelspectrograms(x, SR, out=out), 10)
    peaks = load_peaks()
    gbs = B * (4 * N + 4 * T * 40) / (ms * 1e-3) / 1e9
    # the fused kernel is FP32-issue bound, not HBM bound: 747 warp instructions per frame (ncu, profiles/README.md),
    # i.e. at most 4 IPC x 148 SMs x 1.965 GHz / 747 = 1.56 G frames/s if every scheduler issued every cycle
    issue_bound = 4 * 148 * 1.965e9 / 747.0
    return {"frames_per_s": B * T / (ms * 1e-3), "ms": ms, "hbm_GBps_algorithmic": gbs,
            "frac_of_hbm_peak": gbs / peaks["hbm_gbs"], "peak_source": peaks["source"],
            "issue_bound_frames_per_s": issue_bound, "frac_of_issue_bound": B * T / (ms * 1e-3) / issue_bound}


def bench_embed_quick(device, B=64, sec=2):
    from lidbox_b200.features import audio
    from lidbox_b200.models import xvector
    N = sec * SR
    T = 1 + (N - 400) // 160
    x = synth_signals(B, N, 1234, device=device)
    feats = torch.empty((B, T, 40), dtype=torch.float32, device=device)
    emb = xvector.as_embedding_extractor(xvector.create((T, 40), 4, precision="fp32", seed=0))

    def run():
        audio.logmelspectrograms(x, SR, out=feats)
        return emb(feats)
    ms = _time_cuda(run, 20)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    ms_graph = _time_cuda(g.replay, 50)
    fwd, _ = tdnn_forward_flops(T)
    return {"audio_sec_per_s": B * sec / (ms_graph * 1e-3), "ms_eager": ms, "ms_cuda_graph": ms_graph,
            "precision": "fp32 via bf16x3 tensor-core accumulation (3x the bf16 FLOPs)",
            "algorithmic_tflops": B * fwd / (ms_graph * 1e-3) / 1e12}


def bench_fwd_sweep(device, batches=(64, 256, 1024, 2048), seconds=(1, 2, 5)):
    """BASELINE config 5: log-mel + TDNN forward (bf16 operands, fp32 accumulation) over batch x duration; each cell is
    one CUDA-graph replay timed with CUDA events (inputs of the large cells exceed L2; small cells are L2-resident)."""
    from lidbox_b200.features import audio
    from lidbox_b200.models import xvector
    peaks = load_peaks()
    cells = []
    for sec in seconds:
        N = sec * SR
        T = 1 + (N - 400) // 160
        model = xvector.create((T, 40), 4, precision="bf16", seed=0)
        fwd_flops, _ = tdnn_forward_flops(T)
        for B in batches:
            x = synth_signals(B, N, 1234, device=device)
            feats = torch.empty((B, T, 40), dtype=torch.float32, device=device)

            def lm():
                audio.logmelspectrograms(x, SR, out=feats)

            def run():
                lm()
                return model(feats)
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                run()
            iters = 20 if B * sec <= 2048 else 5
            ms = _time_cuda(g.replay, iters)
            ms_lm = _time_cuda(lm, iters)
            cells.append({"batch": B, "seconds": sec, "ms": ms, "audio_sec_per_s": B * sec / (ms * 1e-3),
                          "logmel_ms": ms_lm, "logmel_frames_per_s": B * T / (ms_lm * 1e-3),
                          "logmel_hbm_frac": B * (4 * N + 4 * T * 40) / (ms_lm * 1e-3) / 1e9 / peaks["hbm_gbs"],
                          "tdnn_tflops": B * fwd_flops / (max(ms - ms_lm, 1e-6) * 1e-3) / 1e12,
                          "tdnn_tensor_frac": B * fwd_flops / (max(ms - ms_lm, 1e-6) * 1e-3) / 1e12 / peaks["bf16_tflops"]})
            del g, x, feats
            torch.cuda.empty_cache()
        del model
    return {"metric": "log-mel + TDNN forward sweep (BASELINE config 5)", "unit": "audio-sec/s", "n_gpus": 1,
            "dtype": "bf16", "data": "synthetic", "peak_source": peaks["source"], "cells": cells}


WORKLOADS = {"logmel": LogmelWorkload, "xvector_train": XVectorTrainWorkload,
             "xvector_ap_train": XVectorAPTrainWorkload}
DEFAULT_WORKLOAD = os.environ.get("LBX_BENCH_WORKLOAD", "xvector_train")


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path (here: the oracle port, TensorFlow is not installable)."""
    if rank != 0:
        return
    wl = WORKLOADS[args.workload](args, 0, 1)
    per = max(2.0, 60.0 / max(1, args.steps + args.warmup))
    if os.environ.get("LBX_REF_BUDGET_S"):          # tests shorten the bounded CPU sample
        per = float(os.environ["LBX_REF_BUDGET_S"])
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


def _log(msg):
    if os.environ.get("LBX_BENCH_VERBOSE"):
        sys.stderr.write("[bench rank %s] %s\n" % (os.environ.get("RANK", "0"), msg))
        sys.stderr.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="lidbox_b200", choices=["lidbox_b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS) + ["fwd_sweep"])
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
    if args.workload == "fwd_sweep":
        if rank == 0:
            print(json.dumps(bench_fwd_sweep(device)), flush=True)
        return
    wl = WORKLOADS[args.workload](args, rank, world)
    wl.dist = dist
    _log("setup")
    wl.setup(device)
    _log("setup done")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        wl.step()
    barrier()
    _log("warmup done")
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
    if launches == 0 and getattr(wl, "launches_per_step", None) and wl.launches_per_step():
        launches = wl.launches_per_step() * args.steps      # CUDA-graph replays: kernels counted at capture time
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps

    _log("timed region done")
    # dominant-kernel timing for the roofline (CUDA events on the launching stream)
    roof = wl.roofline_measure(peaks) if hasattr(wl, "roofline_measure") else wl.roofline(ms_per_step, peaks)

    _log("roofline done")
    # end-to-end through the public API with pinned host buffers
    for _ in range(4):
        wl.step_e2e()
    barrier()
    e_steps = max(10, min(args.steps, 50))
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
    _log("e2e done")
    pcm = None
    if hasattr(wl, "e2e_pcm16") and os.environ.get("LBX_BENCH_PCM16", "1") != "0":
        p_ms, p_bytes = wl.e2e_pcm16(e_steps, barrier)
        if dist is not None:
            t = torch.tensor([p_ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            p_ms = float(t.item())
        pcm = (p_ms, p_bytes)

    if rank == 0:
        units = wl.units_per_step() * world
        line = {"metric": wl.metric, "value": units / (ms_per_step * 1e-3), "unit": wl.unit, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
                "config": wl.config(), "roofline": roof,
                "e2e": {"value": units / (e_ms / e_steps * 1e-3), "unit": wl.unit, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": e_ms / e_steps},
                "gpu_launches": int(launches), "clocks": clocks}
        if pcm is not None:
            line["e2e_pcm16"] = {"value": units / (pcm[0] * 1e-3), "unit": wl.unit, "h2d_bytes_per_step": pcm[1],
                                 "d2h_bytes_per_step": d2h, "ms_per_step": pcm[0],
                                 "note": "same pipeline, batches arrive as 16-bit PCM (WAV sample format) and are "
                                         "decoded inside the log-mel kernel"}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = wl.cpu_sample()
        if hasattr(wl, "extra"):
            line.update(wl.extra())
        print(json.dumps(line), flush=True)
    if dist is not None:
        # ranks > 0 wait here while rank 0 measures the secondary lines; then leave without tearing NCCL down
        # (destroy_process_group can dead-lock while a captured CUDA graph still references the communicator)
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
