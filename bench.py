#!/usr/bin/env python
"""bench.py — measures the lidbox_b200 hot path on B200 (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload logmel|xvector_train|...] [--impl reference]

One JSON line on stdout (rank 0).  `value` is device-resident throughput, `e2e` goes through the public Python
API with pinned HOST buffers (H2D + D2H inside the timed region; the batches arrive as 16-bit PCM, the sample format
read_wav decodes, lidbox/features/audio.py:17-33 — `e2e_f32` is the same pipeline fed with float32), `roofline`
describes the dominant kernel (CUDA-event timed on the launching stream), `cpu_baseline` times the CPU oracle port on
a bounded number of full-size steps.
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

    def wait_ready(self, timeout=5.0):
        """NVML initialisation takes longer than a short timed region: block until the first sample exists."""
        t0 = time.time()
        while not self.samples and self.err is None and time.time() - t0 < timeout:
            time.sleep(0.005)

    def mark(self):
        """Start of the timed region: samples taken before this point (warm-up) are dropped."""
        self.samples, self.reasons = [], set()

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

    def config(self, reference=False):
        cfg = {"workload": "logmel %dx%ds per GPU (BASELINE config 5 max)" % (self.B, self.sec),
               "batch_per_gpu": self.B, "seconds": self.sec, "frames_per_utt": self.T,
               "parallelism": "dp%d (independent shards, no collective)" % self.world}
        if not reference:
            cfg["residency"] = "inputs %.0f MB > L2 (no flush needed)" % (self.B * self.N * 4 / 1e6)
        return cfg

    def setup(self, device):
        from lidbox_b200.features import audio
        self.audio = audio
        self.x_host = synth_signals(self.B, self.N, 1234 + self.rank, pin=True)
        self.pcm_host = to_pcm16(self.x_host).pin_memory()
        self.x = self.x_host.to(device)
        self.out = torch.empty((self.B, self.T, 40), dtype=torch.float32, device=device)
        self.out_host = torch.empty((self.B, self.T, 40), dtype=torch.float32).pin_memory()

    def units_per_step(self):
        return self.B * self.T

    def step(self):
        self.audio.logmelspectrograms(self.x, SR, out=self.out)

    def launches_per_step(self):
        return 1

    def e2e_run(self, x_host, steps, barrier):
        dev = self.x.device

        def one():
            x = x_host.to(dev, non_blocking=True)
            out = self.audio.logmelspectrograms(x, SR, out=self.out)
            self.out_host.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        for _ in range(3):
            one()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one()
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / steps, self.B * self.N * x_host.element_size(), self.B * self.T * 40 * 4

    def roofline(self, ms_per_step, peaks):
        alg = self.B * (4 * self.N + 4 * self.T * 40)
        ach = alg / (ms_per_step * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": "logmel512_kernel<1>", "achieved": ach, "peak": peaks["hbm_gbs"],
                "peak_source": peaks["source"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                "algorithmic_bytes_per_launch": alg, "traffic": None}

    def cpu_step_fn(self, Bs):
        from oracle import lidbox_oracle as O
        torch.set_num_threads(os.cpu_count())
        x = synth_signals(Bs, self.N, 99)
        return lambda: O.torch_logmel(x)

    def cpu_sample(self, budget_s=15.0):
        Bs = min(self.B, 256)
        one = self.cpu_step_fn(Bs)
        one()
        n, t0 = 0, time.perf_counter()
        while True:
            one()
            n += 1
            dt = time.perf_counter() - t0
            if (dt > budget_s and n >= 2) or n >= 50:
                break
        return {"value": n * Bs * self.T / dt, "unit": self.unit, "cores": os.cpu_count(), "kind": "port",
                "sample": "%d iterations of %dx%ds log-mel through oracle.torch_logmel (fp32 torch-CPU restatement; "
                          "TensorFlow is not installable)" % (n, Bs, self.sec), "ms_per_step": dt / n * 1e3}


E2E_LOSS_RING = 8      # end-to-end loop: per-step D2H loss copies land in a ring; the host synchronises once per ring


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


def to_pcm16(x):
    """float32 in [-1, 1) -> 16-bit PCM, the sample format of the WAV corpora (read_wav decodes it as x / 32768)."""
    return torch.clamp(torch.round(x * 32768.0), -32768, 32767).to(torch.int16)


class XVectorTrainWorkload:
    """BASELINE config 3: x-vector training, 4 synthetic language labels, cross-entropy, bf16 (fp32 master weights,
    statistics, loss), Adam; per-GPU batch 256 x 2 s; data parallel (gradient exchange fused into the optimizer kernel).
    A step = log-mel of the resident signals (bf16 rows written straight into the first frame layer's buffer) ->
    forward -> backward -> gradient exchange -> Adam + bf16 weight refresh."""
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

    def config(self, reference=False):
        cfg = {"workload": "%s: x-vector train, batch %d x %d s per GPU, %d labels, %s, bf16, Adam"
                           % (self.label, self.B, self.sec, self.n_classes, self.loss),
               "batch_per_gpu": self.B, "global_batch": self.B * self.world, "seconds": self.sec,
               "frames_per_utt": self.T, "parallelism": "dp%d" % self.world}
        if reference:
            return cfg
        cfg.update({
            "residency": "float32 signals resident in HBM; L2 flushed by the step itself (activations+gradients "
                         "%.0f MB > 126 MB L2)" % self._act_mb(),
            "cuda_graph": self.use_graph,
            "feature_handoff": "log-mel rows are written as bf16 straight into the first frame layer's buffer "
                               "(lbx_logmel_ex), no fp32 feature tensor / packing pass",
            "feature_prefetch": "log-mel of batch i+1 runs on a second stream during step i (one log-mel + one "
                                "training step per replay)" if self.pipelined else "inline",
            "tensor_launches": "5 forward + 5 data-gradient tcgen05 GEMMs, all frame-layer weight gradients in ONE "
                               "grouped stream-K launch (lbx_wgrad_grouped); dense head = lbx_head_fwd + fused "
                               "output/loss kernel + lbx_head_bwd (mma.sync, persistent)",
            "dp_exchange": getattr(self, "dp_exchange", None)})
        return cfg

    def _act_mb(self):
        return self.B * self.T * (512 * 2 * 2 + 256 * 2 * 2 + 2 * 1504 * 2 / 6 + 160) / 1e6

    def setup(self, device):
        from lidbox_b200.features import audio
        from lidbox_b200.models import xvector
        self.audio, self.device, self.xvector = audio, device, xvector
        self.x_host, y = class_signals(self.B, self.N, self.n_classes, 1234 + self.rank, pin=True)
        self.pcm_host = to_pcm16(self.x_host).pin_memory()
        self.y = y.to(device)
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
        self.sinks = [self.model.feature_sink(self.B, self.T, training=True, slot=k) for k in range(2)]
        self.pipe = self._build_pipe(self.x_host)
        self.x = self.pipe["xs"][0]
        self.graphed = self.pipe["steps"][0] if self.use_graph else None

    def _build_pipe(self, x_host, loss_to_host=False):
        """Two signal buffers (dtype of x_host: float32 or 16-bit PCM) and the two input buffers of the model: while
        the model trains on the features of batch i, the log-mel of batch i+1 is written into the other input buffer
        on a second stream (the prefetch a tf.data input pipeline does), and in the end-to-end path the H2D copy of
        batch i+2 runs on a copy stream.  Every replay = one log-mel + one training step."""
        dev, audio, xvector = self.device, self.audio, self.xvector
        pipe = {"x_host": x_host, "xs": [x_host.to(dev), x_host.to(dev)], "i": 0, "steps": [], "e2e": None,
                # end-to-end loop: the per-sample losses of every step land here through a copy node inside the graph
                "loss_host": [torch.empty((self.B,), dtype=torch.float32).pin_memory() for _ in range(E2E_LOSS_RING)]
                if (loss_to_host and self.use_graph) else None}
        audio.logmelspectrograms(pipe["xs"][0], SR, out=self.sinks[0])          # prime the pipeline
        # two alternating input buffers; the end-to-end pipe captures one graph per slot of the host-side loss ring
        # (slot j trains from input buffer j % 2 and copies its losses into pinned buffer j), so that the host finds the
        # losses of ALL steps since its last synchronisation
        for j in range(E2E_LOSS_RING if pipe["loss_host"] else 2):
            k = j % 2
            nxt = (lambda k=k: audio.logmelspectrograms(pipe["xs"][1 - k], SR, out=self.sinks[1 - k]))
            inline = (lambda k=k: audio.logmelspectrograms(pipe["xs"][k], SR, out=self.sinks[k]))
            if self.use_graph:
                pipe["steps"].append(xvector.GraphedTrainStep(
                    self.model, self.sinks[k], self.y, loss=self.loss, process_group=self.pg,
                    pre=None if self.pipelined else inline, concurrent=nxt if self.pipelined else None,
                    loss_host=pipe["loss_host"][j] if pipe["loss_host"] else None,
                    # end-to-end loop: the H2D transfer of batch i+2 into signal buffer k (free during replay k) is a
                    # branch of the same graph: a whole end-to-end step is ONE graph launch
                    copies=[(pipe["xs"][k], x_host)] if pipe["loss_host"] else None, **self.kw))
            else:
                pipe["steps"].append(lambda inline=inline: self.model.train_step(inline(), self.y, loss=self.loss,
                                                                                 process_group=self.pg, **self.kw))
        return pipe

    def _eager_step(self):
        feats = self.audio.logmelspectrograms(self.x, SR, out=self.sinks[0])
        return self.model.train_step(feats, self.y, loss=self.loss, process_group=self.pg, **self.kw)

    def units_per_step(self):
        return self.B * self.sec

    def step(self, pipe=None):
        pipe = pipe or self.pipe
        out = pipe["steps"][pipe["i"] % len(pipe["steps"])]()
        pipe["i"] += 1
        return out

    def launches_per_step(self):
        return self.graphed.kernels_per_step if self.graphed is not None else None

    def _setup_e2e(self, pipe):
        dev = self.device
        e = {"copy": torch.cuda.Stream(device=dev),
             "h2d_done": [torch.cuda.Event(), torch.cuda.Event()], "step_done": torch.cuda.Event(),
             "loss_host": [torch.empty((self.B,), dtype=torch.float32).pin_memory() for _ in range(E2E_LOSS_RING)]}
        cur = torch.cuda.current_stream(dev)
        torch.cuda.synchronize(dev)
        # prime: batch 0 -> xs[k0] + its features, batch 1 -> xs[1-k0]
        k0 = pipe["i"] % 2
        pipe["xs"][k0].copy_(pipe["x_host"], non_blocking=True)
        self.audio.logmelspectrograms(pipe["xs"][k0], SR, out=self.sinks[k0])
        with torch.cuda.stream(e["copy"]):
            e["copy"].wait_stream(cur)
            pipe["xs"][1 - k0].copy_(pipe["x_host"], non_blocking=True)
            e["h2d_done"][1 - k0].record(e["copy"])
        e["step_done"].record(cur)
        pipe["e2e"] = e

    def step_e2e(self, pipe=None):
        """One step of the public-API pipeline with the batch in pinned HOST memory: replay k trains on the features
        of batch i (input buffer k) and extracts the features of batch i+1 from device buffer 1-k, while batch i+2 is
        copied host->device into buffer k on the copy stream; the per-sample losses of batch i are copied back."""
        pipe = pipe or self.pipe
        if pipe["loss_host"] is not None:
            # CUDA-graph pipeline: host->device copy of batch i+2, log-mel of batch i+1, training step on batch i and the
            # device->host copy of its losses are all nodes of the replayed graph
            self.step(pipe)
            if (pipe["i"] - 1) % E2E_LOSS_RING == E2E_LOSS_RING - 1:
                torch.cuda.current_stream(self.device).synchronize()   # the host reads the losses that have arrived
            return
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
        slot = (pipe["i"] - 1) % E2E_LOSS_RING                 # pipe["i"] was advanced by self.step()
        if pipe["loss_host"] is None:                          # (no CUDA graph: copy the losses with a stream operation)
            e["loss_host"][slot].copy_(losses, non_blocking=True)
        # with CUDA graphs every replay copies its per-sample losses into pinned host memory itself (a copy node inside
        # the graph, GraphedTrainStep(loss_host=...)): nothing is enqueued between two replays for it
        if slot == E2E_LOSS_RING - 1:
            cur.synchronize()                                  # the host reads the losses that have arrived

    def e2e_run(self, x_host, steps, barrier):
        """Times `steps` end-to-end steps fed from the pinned host tensor x_host (int16 PCM or float32)."""
        pipe = self._build_pipe(x_host, loss_to_host=True)
        for _ in range(4):
            self.step_e2e(pipe)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            self.step_e2e(pipe)
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / steps, self.B * self.N * x_host.element_size(), self.B * 4

    def measure_exchange(self, iters=20):
        """Duration of the optimizer kernel (N = 1: Adam + bf16 refresh; N > 1: flag exchange + reduce-scatter + Adam on
        the shard + all-gather) from CUDA events around it in eager steps — the part of the step that is exposed after
        the backward pass."""
        m = self.model
        ts = []
        for _ in range(iters + 3):
            feats = self.audio.logmelspectrograms(self.x, SR, out=self.sinks[0])
            m.loss_and_grads(feats, self.y, loss=self.loss, global_batch=self.B * self.world, **self.kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if m._sharded is not None:
                m._apply_sharded()
            else:
                if self.pg is not None:
                    self.dist.all_reduce(m.grads, group=self.pg)
                m.apply_gradients()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return float(np.median(ts[3:]))

    def roofline_measure(self, peaks):
        """Dominant kernel = gemm_bf16_kernel.  Every GEMM descriptor of one training step is recorded, the launches
        are replayed back-to-back from a CUDA graph (same operands, same order, nothing else in between) and timed
        with CUDA events on the launching stream; achieved = algorithmic training FLOPs of the step / that time.
        The replay lasts milliseconds at boost clock, so the fraction is taken against the BURST peak; the fraction of
        the sustained peak is printed next to it."""
        from lidbox_b200 import ops
        fwd, f1 = tdnn_forward_flops(self.T, self.n_out)
        alg_step = self.B * (3 * fwd - f1)
        ops.GEMM_RECORD = []
        self._eager_step()
        torch.cuda.synchronize()
        rec, ops.GEMM_RECORD = ops.GEMM_RECORD, None
        # the dense head (1 % of the FLOPs) runs in the persistent mma.sync kernels lbx_head_fwd / lbx_head_bwd /
        # lbx_dense_xent_head when they are enabled: its FLOPs are then not part of the replayed tensor-core launches
        head_in_replay = any((not isinstance(shape, list)) and shape[0] == self.B for (_d, _k, shape) in rec)
        dense = 2 * (3000 * 512 + 512 * 512 + 512 * self.n_out)
        alg = alg_step if head_in_replay else self.B * (3 * (fwd - dense) - f1)
        issued = 0.0
        for (_d, _k, shape) in rec:        # one (M, N, K, n_terms, layout) per GEMM, a list of (M, N, K) per grouped launch
            issued += (sum(2.0 * M * N * K for (M, N, K) in shape) if isinstance(shape, list)
                       else 2.0 * shape[0] * shape[1] * shape[2] * shape[3])
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            ops.replay(rec, self.device)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            ops.replay(rec, self.device)
        ms = _time_cuda(g.replay, 50)
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
        traffic, traffic_src = None, None
        try:        # dram__bytes_read+write per launch from the committed ncu --set full capture of the same step
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_gemm_traffic.json")))
            if tj.get("launches") == len(rec) and (self.B, self.sec) == (256, 2):
                traffic = tj["dram_bytes_per_launch"]
                traffic_src = "profiles/r2_gemm_traffic.json (ncu --set full capture of this step; not measured live)"
        except Exception:
            pass
        return {"bound": "tensor", "kernel": "gemm_bf16_kernel + wgrad_grouped_kernel (all %d tensor-core launches of one "
                                             "training step, replayed back-to-back from a CUDA graph)" % len(rec),
                "achieved": ach, "peak": peaks["bf16_tflops"],
                "peak_source": peaks["source"] + " (burst: the GEMM replay lasts %.0f ms at boost clock)" % (ms * 53),
                "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"], "frac_burst": ach / peaks["bf16_tflops"],
                "frac_sustained": ach / peaks["bf16_tflops_sustained"], "peak_sustained": peaks["bf16_tflops_sustained"],
                "algorithmic_flops_per_step": alg_step, "algorithmic_flops_in_replay": alg,
                "issued_flops_per_step": issued, "gemm_ms_per_step": ms,
                "launches": len(rec), "avg_launch_us": ms * 1e3 / len(rec),
                "largest_launch": {"M": big[2], "N": big[3], "K": big[4], "ms": big_ms,
                                   "tflops": 2.0 * big[2] * big[3] * big[4] / (big_ms * 1e-3) / 1e12},
                "algorithmic_flops_per_launch": alg / len(rec), "traffic": traffic, "traffic_source": traffic_src}

    def cpu_step_fn(self, Bs):
        """One training step of the CPU port (oracle.torch_logmel + torch_xvector_forward + autograd + Adam, fp32,
        all host threads) on a batch of Bs utterances; returns a callable."""
        from oracle import lidbox_oracle as O
        torch.set_num_threads(os.cpu_count())
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
        return one

    def cpu_sample(self, budget_s=15.0):
        """Bounded CPU baseline: full-size steps (the configured batch) of the CPU port, as many as fit the budget
        (at least 2 after one warm-up step)."""
        one = self.cpu_step_fn(self.B)
        one()
        n, t0 = 0, time.perf_counter()
        while True:
            one()
            n += 1
            dt = time.perf_counter() - t0
            if (dt > budget_s and n >= 2) or n >= 200:
                break
        return {"value": n * self.B * self.sec / dt, "unit": self.unit, "cores": os.cpu_count(), "kind": "port",
                "sample": "%d training steps of batch %d x %d s through the torch-CPU fp32 restatement "
                          "(oracle.torch_logmel + torch_xvector_forward + autograd + Adam; TensorFlow is not "
                          "installable)" % (n, self.B, self.sec), "ms_per_step": dt / n * 1e3}

    def extra(self):
        """Secondary BASELINE lines measured in the same run (device-timed, signals resident)."""
        out = {}
        try:
            also = {"logmel_2048x5s": bench_logmel_quick(self.device),
                    "config2_embed_64x2s_fp32": bench_embed_quick(self.device)}
            if self.name == "xvector_train":
                also["config4_ap_train_256x3s"] = bench_train_quick(self.device, XVectorAPTrainWorkload)
                also["config5_fwd_256x2s"] = bench_fwd_cell(self.device, 256, 2)
                also["config5_fwd_2048x5s"] = bench_fwd_cell(self.device, 2048, 5)
            out["also"] = also
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
    pcm = to_pcm16(x)
    ms16 = _time_cuda(lambda: audio.logmelspectrograms(pcm, SR, out=out), 10)
    peaks = load_peaks()
    gbs = B * (4 * N + 4 * T * 40) / (ms * 1e-3) / 1e9
    # ncu (profiles/r2_logmel_ncu_summary.json): 429 warp instructions and 133 shared-memory wavefronts per frame; the
    # shared-memory pipe (1 wavefront / clk / SM) is the tightest per-frame bound of this FFT formulation
    smem_bound = 148 * 1.965e9 / 133.0
    return {"frames_per_s": B * T / (ms * 1e-3), "ms": ms, "ms_pcm16_input": ms16, "hbm_GBps_algorithmic": gbs,
            "frac_of_hbm_peak": gbs / peaks["hbm_gbs"], "peak_source": peaks["source"],
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": gbs / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": B * (4 * N + 4 * T * 40)},
            "smem_pipe_bound_frames_per_s": smem_bound, "frac_of_smem_pipe_bound": B * T / (ms * 1e-3) / smem_bound}


def bench_embed_quick(device, B=64, sec=2):
    """BASELINE config 2: x-vector embedding extraction, 64 x 2 s, fp32-grade (bf16x3) — signals -> log-mel (hi/lo bf16
    planes written straight into the first frame layer's buffers) -> TDNN -> segment1 embedding."""
    from lidbox_b200.features import audio
    from lidbox_b200.models import xvector
    N = sec * SR
    T = 1 + (N - 400) // 160
    x = synth_signals(B, N, 1234, device=device)
    m = xvector.create((T, 40), 4, precision="fp32", seed=0)
    emb = xvector.as_embedding_extractor(m)
    sink = m.feature_sink(B, T)

    def run():
        return emb(audio.logmelspectrograms(x, SR, out=sink))
    ms = _time_cuda(run, 20)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    ms_graph = _time_cuda(g.replay, 50)
    fwd, _ = tdnn_forward_flops(T)
    peaks = load_peaks()
    return {"audio_sec_per_s": B * sec / (ms_graph * 1e-3), "ms_eager": ms, "ms_cuda_graph": ms_graph,
            "precision": "fp32 via bf16x3 tensor-core accumulation (3x the bf16 FLOPs)",
            "algorithmic_tflops": B * fwd / (ms_graph * 1e-3) / 1e12,
            "roofline": {"bound": "tensor", "achieved": B * fwd / (ms_graph * 1e-3) / 1e12, "peak": peaks["bf16_tflops"],
                         "unit": "TFLOP/s", "frac": B * fwd / (ms_graph * 1e-3) / 1e12 / peaks["bf16_tflops"],
                         "note": "launch-latency bound: 12.6 us of tensor work in a chain of ~10 kernels"}}


def bench_train_quick(device, cls, steps=50):
    """Another training config measured the same way as the headline (CUDA-graph replays, device-resident signals)."""
    class A:
        batch, seconds = 0, 0
    wl = cls(A, 0, 1)
    wl.setup(device)
    for _ in range(5):
        wl.step()
    ms = _time_cuda(wl.step, steps, warm=0)
    fwd, f1 = tdnn_forward_flops(wl.T, wl.n_out)
    peaks = load_peaks()
    ach = wl.B * (3 * fwd - f1) / (ms * 1e-3) / 1e12
    out = {"workload": wl.config()["workload"], "ms_per_step": ms, "audio_sec_per_s": wl.B * wl.sec / (ms * 1e-3),
           "roofline": {"bound": "tensor", "scope": "whole step (log-mel + fwd + bwd + Adam)", "achieved": ach,
                        "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"]}}
    del wl
    torch.cuda.empty_cache()
    return out


def bench_fwd_cell(device, B, sec):
    """One cell of BASELINE config 5: log-mel + TDNN forward (bf16 operands, fp32 accumulation), one CUDA-graph
    replay per iteration, with the roofline of each half."""
    from lidbox_b200.features import audio
    from lidbox_b200.models import xvector
    peaks = load_peaks()
    N = sec * SR
    T = 1 + (N - 400) // 160
    model = xvector.create((T, 40), 4, precision="bf16", seed=0)
    fwd_flops, _ = tdnn_forward_flops(T)
    x = synth_signals(B, N, 1234, device=device)
    sink = model.feature_sink(B, T)

    def lm():
        return audio.logmelspectrograms(x, SR, out=sink)

    def run():
        return model(lm())
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    iters = 20 if B * sec <= 2048 else 5
    ms = _time_cuda(g.replay, iters)
    ms_lm = _time_cuda(lm, iters)
    tf = B * fwd_flops / (max(ms - ms_lm, 1e-6) * 1e-3) / 1e12
    gbs = B * (4 * N + 2 * T * 40) / (ms_lm * 1e-3) / 1e9
    out = {"batch": B, "seconds": sec, "ms": ms, "audio_sec_per_s": B * sec / (ms * 1e-3), "logmel_ms": ms_lm,
           "logmel_frames_per_s": B * T / (ms_lm * 1e-3),
           "roofline_logmel": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                               "frac": gbs / peaks["hbm_gbs"], "algorithmic_bytes": "4N in + 2*T*40 out (bf16 rows)"},
           "roofline_tdnn": {"bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                             "frac": tf / peaks["bf16_tflops"]}}
    del g, x, model, sink
    torch.cuda.empty_cache()
    return out


def bench_fwd_sweep(device, batches=(64, 256, 1024, 2048), seconds=(1, 2, 5)):
    """BASELINE config 5: log-mel + TDNN forward (bf16 operands, fp32 accumulation) over batch x duration; each cell is
    one CUDA-graph replay timed with CUDA events (inputs of the large cells exceed L2; small cells are L2-resident)."""
    peaks = load_peaks()
    cells = [bench_fwd_cell(device, B, sec) for sec in seconds for B in batches]
    return {"metric": "log-mel + TDNN forward sweep (BASELINE config 5)", "unit": "audio-sec/s", "n_gpus": 1,
            "dtype": "bf16", "data": "synthetic", "peak_source": peaks["source"], "cells": cells}


WORKLOADS = {"logmel": LogmelWorkload, "xvector_train": XVectorTrainWorkload,
             "xvector_ap_train": XVectorAPTrainWorkload}
DEFAULT_WORKLOAD = os.environ.get("LBX_BENCH_WORKLOAD", "xvector_train")


def run_reference(args, rank, world, emit):
    """--impl reference: the reference's own CPU path (here: the oracle port, TensorFlow is not installable), all host
    threads.  Every step is ONE real step of the workload on the configured batch; if K steps of that size would not
    finish within ~2.5 minutes the per-step sample is shrunk (and the line says so)."""
    if rank != 0:
        return
    wl = WORKLOADS[args.workload](args, 0, 1)
    budget = float(os.environ.get("LBX_REF_BUDGET_S", "150"))
    Bs = wl.B if wl.name != "logmel" else min(wl.B, 256)
    one = wl.cpu_step_fn(Bs)
    t0 = time.perf_counter()
    one()                                                    # first call: thread pools, allocator
    t_full = time.perf_counter() - t0
    n_total = max(1, args.steps + args.warmup)
    if t_full * n_total > budget and Bs > 8:
        Bs = max(8, int(Bs * budget / (t_full * n_total)))
        one = wl.cpu_step_fn(Bs)
        one()
    for _ in range(args.warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = time.perf_counter() - t0
    per_step_units = Bs * (wl.sec if wl.unit == "audio-sec/s" else wl.T)
    value = per_step_units * args.steps / dt
    cfg = wl.config(reference=True)
    cfg["batch_per_step"] = Bs
    cfg["impl_note"] = ("CPU port of the reference path (oracle.torch_logmel + torch_xvector_forward + autograd + Adam, "
                        "fp32); TensorFlow is not installable here")
    cb = {"value": value, "unit": wl.unit, "cores": os.cpu_count(), "kind": "port",
          "sample": "%d steps of batch %d x %d s (configured batch %d)" % (args.steps, Bs, wl.sec, wl.B)}
    line = {"impl": "reference", "metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "cpu_baseline": cb,
            "e2e": {"value": value, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's host threads (and therefore its pinned-buffer allocations) to the NUMA node of its GPU.
    Best effort: returns a short description for the bench line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if bus.startswith("0000") and len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return {"numa_node": node, "bound": False}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "bound": bool(cpus), "cpus": len(cpus)}
    except Exception as e:
        return {"bound": False, "error": repr(e)[:80]}


def dp_selfcheck(dist, device):
    """world > 1: the summed gradient of all ranks BEFORE Adam, through both exchange paths, against the gradient rank 0
    computes alone on the concatenated batch (same weights).  NCCL path: all-reduce of the flat gradient.  Fused path:
    lbx_adam_step_sharded with lr = 0 leaves the weights alone and stores m = (1 - beta1) * sum_r g_r."""
    from lidbox_b200.models import xvector
    rank, world = dist.get_rank(), dist.get_world_size()
    B, T = 4 * world, 61
    rng = np.random.default_rng(0)
    x = rng.standard_normal((B, T, 40)).astype(np.float32)
    y = np.arange(B) % 4
    per = B // world
    xs, ys = x[rank * per:(rank + 1) * per], y[rank * per:(rank + 1) * per]
    ma = xvector.create((T, 40), 4, precision="bf16", seed=5)
    ma.loss_and_grads(xs, ys, global_batch=B)
    g_nccl = ma.grads.clone()
    dist.all_reduce(g_nccl)
    mb = xvector.create((T, 40), 4, precision="bf16", seed=5)
    mb.configure_optimizer(lr=0.0)
    mb.enable_sharded_optimizer(dist.group.WORLD)
    mb.loss_and_grads(xs, ys, global_batch=B)
    mb._apply_sharded()
    sh = mb._sharded
    shards = [torch.empty_like(sh["m"]) for _ in range(world)]
    dist.all_gather(shards, sh["m"])
    g_fused = torch.cat(shards) / (1.0 - 0.9)
    out = None
    if rank == 0:
        mr = xvector.create((T, 40), 4, precision="bf16", seed=5)
        mr.loss_and_grads(x, y)
        g_ref = mr.grads
        wn = wf = 0.0
        for ly in mr.layers:
            lo, hi = ly["w_off"], ly["b_off"] + ly["ldw"]
            den = g_ref[lo:hi].abs().max().item() + 1e-30
            wn = max(wn, (g_nccl[lo:hi] - g_ref[lo:hi]).abs().max().item() / den)
            wf = max(wf, (g_fused[lo:hi] - g_ref[lo:hi]).abs().max().item() / den)
        out = {"what": "summed flat gradient of %d ranks vs one rank on the concatenated batch, worst per-layer "
                       "max|diff| / max|ref|" % world,
               "nccl_allreduce": wn, "fused_exchange": wf, "nvls": bool(sh.get("mc_grads")),
               "barrier_timeouts": int(sh["local"][3].item()), "ok": bool(wn < 2e-4 and wf < 2e-4)}
    dist.barrier()
    torch.cuda.synchronize()
    return out


def _log(msg):
    if os.environ.get("LBX_BENCH_VERBOSE"):
        sys.stderr.write("[bench rank %s] %s\n" % (os.environ.get("RANK", "0"), msg))
        sys.stderr.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
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
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner) go to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        run_reference(args, rank, world, emit)
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

    numa = bind_to_gpu_numa_node(local_rank)           # before any pinned allocation
    from lidbox_b200 import _lib
    lib = _lib.lib()
    peaks = load_peaks()
    if args.workload == "fwd_sweep":
        if rank == 0:
            emit(bench_fwd_sweep(device))
        return
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                               # NVML init overlaps the set-up and the warm-up
    dp_check = None
    if dist is not None and os.environ.get("LBX_BENCH_DP_CHECK", "1") != "0":
        dp_check = dp_selfcheck(dist, device)
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
    if rank == 0:
        sampler.wait_ready()
    n0 = lib.lbx_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    ev0.record()
    for _ in range(args.steps):
        wl.step()
    ev1.record()
    barrier()
    launches = lib.lbx_launch_count() - n0
    if launches == 0 and getattr(wl, "launches_per_step", None) and wl.launches_per_step():
        launches = wl.launches_per_step() * args.steps      # CUDA-graph replays: kernels counted at capture time
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    clocks = None
    if ms < 250.0:
        # a timed region shorter than a few NVML polls: keep the SAME workload running (untimed, the same number of
        # steps on every rank: the steps contain cross-rank barriers) until the sampler has seen it under load, so that
        # the clocks / throttle record describes this kernel mix
        for _ in range(int(400.0 / max(ms_per_step, 1e-3)) + 1):
            wl.step()
        torch.cuda.synchronize()
    if rank == 0:
        clocks = sampler.stop()
        clocks["timed_region_ms"] = ms
    if dist is not None:
        dist.barrier()

    _log("timed region done")
    # dominant-kernel timing for the roofline (CUDA events on the launching stream)
    roof = wl.roofline_measure(peaks) if hasattr(wl, "roofline_measure") else wl.roofline(ms_per_step, peaks)

    exch_us = None
    if hasattr(wl, "measure_exchange"):
        exch_us = wl.measure_exchange()
        if dist is not None:
            t = torch.tensor([exch_us], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            exch_us = float(t.item())
            wl.model.dp_health()
    _log("roofline done")
    # end-to-end through the public API with pinned host buffers: 16-bit PCM batches (primary), float32 (secondary)
    e_steps = max(10, min(args.steps, 50))

    def e2e(x_host):
        e_ms, h2d, d2h = wl.e2e_run(x_host, e_steps, barrier)
        if dist is not None:
            t = torch.tensor([e_ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        return e_ms, h2d, d2h
    p_ms, p_h2d, p_d2h = e2e(wl.pcm_host)
    _log("e2e pcm16 done")
    f_ms, f_h2d, f_d2h = (None, None, None)
    if os.environ.get("LBX_BENCH_E2E_F32", "1") != "0":
        f_ms, f_h2d, f_d2h = e2e(wl.x_host)
    _log("e2e done")

    if rank == 0:
        units = wl.units_per_step() * world
        pcie_gbs = 63.0                                         # PCIe Gen5 x16, one direction, nominal payload rate
        line = {"metric": wl.metric, "value": units / (ms_per_step * 1e-3), "unit": wl.unit, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
                "config": wl.config(), "roofline": roof,
                "e2e": {"value": units / (p_ms * 1e-3), "unit": wl.unit, "h2d_bytes_per_step": p_h2d,
                        "d2h_bytes_per_step": p_d2h, "ms_per_step": p_ms,
                        "input": "16-bit PCM batches in pinned host memory (the sample format read_wav decodes, "
                                 "lidbox/features/audio.py:17-33), decoded inside the log-mel kernel",
                        "h2d_GBps_per_gpu": p_h2d / (p_ms * 1e-3) / 1e9,
                        "pcie_frac": p_h2d / (p_ms * 1e-3) / 1e9 / pcie_gbs, "numa": numa},
                "gpu_launches": int(launches), "clocks": clocks}
        if f_ms is not None:
            line["e2e_f32"] = {"value": units / (f_ms * 1e-3), "unit": wl.unit, "h2d_bytes_per_step": f_h2d,
                               "d2h_bytes_per_step": f_d2h, "ms_per_step": f_ms,
                               "h2d_GBps_per_gpu": f_h2d / (f_ms * 1e-3) / 1e9,
                               "pcie_frac": f_h2d / (f_ms * 1e-3) / 1e9 / pcie_gbs,
                               "note": "same pipeline fed with float32 signals: PCIe-bound"}
        if dp_check is not None:
            line["dp_check"] = dp_check
        if exch_us is not None:
            line["dp_exchange_us"] = exch_us
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = wl.cpu_sample()
        if hasattr(wl, "extra"):
            line.update(wl.extra())
        emit(line)
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
