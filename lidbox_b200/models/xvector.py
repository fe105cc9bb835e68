"""Drop-in for lidbox/models/xvector.py on hand-written sm_100a kernels.

    m = create(input_shape=(T or None, F), num_outputs)     # xvector.py:46-67
    logp = m(x, training=False)                              # [B, T, F] -> [B, num_outputs] log-probabilities
    emb = as_embedding_extractor(m)(x)                       # xvector.py:70-73: pre-ReLU segment1 output [B, 512]

Layer names, Keras weight layouts (Conv1D kernel [k, C_in, C_out], Dense kernel [in, out]) and the causal/strided
semantics follow the reference; the arithmetic runs through the C-ABI (include/lidbox_b200.h):

  * every frame layer is ONE tcgen05 GEMM over an overlapping TMA view of the zero-left-padded NWC activations
    (row pitch stride*C_in, width k*C_in: no im2col) with bias + ReLU fused in the epilogue, which writes straight
    into the next layer's padded buffer;
  * activation buffers of consecutive layers share one row geometry (rows per utterance R_L = padded length of
    layer L+1), so every operand of forward, data-gradient and weight-gradient GEMMs is a flat 2-D matrix;
  * precision "fp32" (default, inference parity): bf16x3 split accumulation, fp32 statistics;
    precision "bf16" (training configs): bf16 operands, fp32 accumulation / statistics / loss / master weights.
"""
import ctypes
import math
import os

import numpy as np
import torch

from .. import _lib, ops

TIME_AXIS = 1                      # xvector.py:21
STDDEV_SQRT_MIN_CLIP = 1e-10       # xvector.py:22
_SLACK_ROWS = 8


def _ceil8(n):
    return (n + 7) // 8 * 8


BN_MOMENTUM, BN_EPSILON = 0.99, 1e-3       # Keras BatchNormalization defaults (xvector_2d.py:36)


class _LayerSpec:
    def __init__(self, kind, name, units, kernel_size=1, strides=1, activation="relu", dilation_rate=1,
                 batch_norm=False):
        self.kind, self.name, self.units = kind, name, units
        self.kernel_size, self.strides, self.activation = kernel_size, strides, activation
        self.dilation_rate, self.batch_norm = dilation_rate, batch_norm


def frame_layer(filters, kernel_size, strides, padding="causal", activation="relu", name="frame", dilation_rate=1,
                batch_norm=False):
    """xvector.py:38-39 — Conv1D(filters, kernel_size, strides, padding="causal", activation="relu").

    Two extensions beyond the reference's x-vector (both default off; the reference code has neither, its prose has):
    dilation_rate > 1 gives the dilated TDNN layer (Keras Conv1D semantics: causal left padding dilation*(k-1), tap j
    reads time t + j*dilation; like Keras, dilation needs strides == 1), and batch_norm=True appends an inference-time
    BatchNormalization AFTER the activation, the order of the only Conv+BN frame layer of the reference
    (xvector_2d.py:41-43: conv -> ReLU -> BN), with the Keras defaults momentum 0.99 / epsilon 1e-3."""
    if padding != "causal":
        raise NotImplementedError("only padding='causal' is implemented (the only mode the reference uses)")
    if int(dilation_rate) < 1:
        raise ValueError("dilation_rate must be >= 1")
    if int(dilation_rate) > 1 and int(strides) != 1:
        raise ValueError("`strides > 1` not supported in conjunction with `dilation_rate > 1`")     # Keras' own check
    if int(dilation_rate) > 1 and int(kernel_size) > 5:
        raise NotImplementedError("dilated layers support kernel_size <= 5 (one accumulating GEMM pass per tap)")
    return _LayerSpec("frame", name, filters, kernel_size, strides, activation, int(dilation_rate), bool(batch_norm))


def segment_layer(units, activation="relu", name="segment"):
    """xvector.py:42-43 — Dense(units, activation="relu")."""
    return _LayerSpec("segment", name, units, activation=activation)


class GlobalMeanStddevPooling1D:
    """xvector.py:25-35: mean and standard deviation over the time axis, concatenated -> [B, 2C]."""

    def __init__(self, name="stats_pooling"):
        self.name = name

    def __call__(self, inputs):
        x = inputs if isinstance(inputs, torch.Tensor) else torch.as_tensor(np.asarray(inputs))
        if x.dim() != 3:
            raise ValueError("expected [B, T, C], got %s" % (tuple(x.shape),))
        x = x.to(_lib.require_cuda(), torch.float32).contiguous()
        B, T, C = x.shape
        out = torch.empty((B, 2 * C), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().lbx_stats_pool_fwd(_lib.ptr(x), ops.F32, B, T, T, C, C, STDDEV_SQRT_MIN_CLIP,
                                                 _lib.ptr(out), None, None, None, _lib.stream_ptr(x.device)))
        return out


class FeatureSink:
    """Handle of the first frame layer's activation buffer for one (B, T): `features.audio.logmelspectrograms(...,
    out=sink)` writes its bf16 rows straight into it (zero-left-padded NWC layout) and the model then
    consumes the sink instead of a [B, T, F] tensor — the fp32 feature tensor and the packing pass never exist."""

    def __init__(self, model, bufs, B, T, training, slot=0):
        self.model, self.bufs, self.B, self.T, self.training, self.slot = model, bufs, B, T, training, slot
        self.shape = (B, T, model.F)
        # slot 0 is the model's own input buffer; further slots are private copies of it, so that the features of
        # batch i+1 can be produced while the step of batch i still reads its own (the weight gradient of the first
        # frame layer reads the input at the very end of the backward pass)
        if slot == 0:
            self.x0, self.x0_lo = bufs["X0_own"], bufs["X0_lo_own"]
        else:
            self.x0 = torch.zeros_like(bufs["X0_own"])
            self.x0_lo = torch.zeros_like(bufs["X0_lo_own"]) if bufs["X0_lo_own"] is not None else None

    def sink_spec(self, B, T, n_feat):
        """(hi pointer, lo pointer or None, utterance pitch, row pitch) in elements — see lbx_logmel_t."""
        m, geo = self.model, self.bufs["geo"]
        if (B, T, n_feat) != (self.B, self.T, m.F):
            raise ValueError("feature sink was created for %s, got features of shape %s" %
                             ((self.B, self.T, m.F), (B, T, n_feat)))
        off = geo.pad[0] * m.Fp * 2                            # bytes: k-1 zero rows in front of every utterance
        hi = self.x0.data_ptr() + off
        lo = self.x0_lo.data_ptr() + off if self.x0_lo is not None else None
        return hi, lo, geo.Tpad[0] * m.Fp, m.Fp


class _Geometry:
    """Row geometry shared by all activation buffers for one (T) — see DESIGN.md §Data layout."""

    def __init__(self, T, frames):
        self.T = [T]
        for f in frames:
            self.T.append(-(-self.T[-1] // f.strides))
        n = len(frames)
        prod = [1] * (n + 1)                        # prod[L] = s_{L+1} * ... * s_n   (0-based layer index L)
        for L in range(n - 1, -1, -1):
            prod[L] = prod[L + 1] * frames[L].strides
        P = self.T[n]
        for L in range(n):
            need = self.T[L] + (frames[L].kernel_size - 1) * getattr(frames[L], "dilation_rate", 1)
            P = max(P, -(-need // prod[L]))
        self.P = P
        self.Tpad = [P * prod[L] for L in range(n)]            # padded input length of layer L
        self.R = [self.Tpad[L] // frames[L].strides for L in range(n)]   # GEMM rows per utterance of layer L
        # data row offset inside layer L's input buffer = causal padding dilation * (k - 1)
        self.pad = [(f.kernel_size - 1) * getattr(f, "dilation_rate", 1) for f in frames]
        for L in range(n - 1):
            assert self.R[L] == self.Tpad[L + 1]


class XVector:
    def __init__(self, input_shape, num_outputs, channel_dropout_rate=0, name="x-vector", frames=None, segments=None,
                 precision="fp32", head="log_softmax", seed=None, device=None, output_name="outputs"):
        if len(input_shape) != 2:
            raise ValueError("input_shape must be (T or None, F)")
        self.name = name
        self.input_shape = tuple(input_shape)
        self.F = int(input_shape[1])
        self.Fp = _ceil8(self.F)
        self.num_outputs = int(num_outputs)
        if self.F < 1 or self.num_outputs < 1:
            raise ValueError("F and num_outputs must be >= 1")
        self.channel_dropout_rate = float(channel_dropout_rate)
        self.frames = frames or [frame_layer(512, 5, 1, name="frame1"), frame_layer(512, 3, 2, name="frame2"),
                                 frame_layer(512, 3, 3, name="frame3"), frame_layer(512, 1, 1, name="frame4"),
                                 frame_layer(1500, 1, 1, name="frame5")]
        self.segments = segments or [segment_layer(512, name="segment1"), segment_layer(512, name="segment2")]
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        if head not in ("log_softmax", "l2_normalize", "none"):
            raise ValueError("unknown head " + head)
        self.precision, self.head = precision, head
        self.output_name = output_name                     # "outputs" (xvector.py:64) / "output" (xvector_extended.py:40)
        self.device = _lib.require_cuda(device)
        self._dropout_seed = 0x5EED if seed is None else int(seed)
        self._dropout_counter = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._build_params(seed)
        self._bufs = {}
        self._adam = None
        self._sharded = None
        self._grads_clean = True
        self.overlap_wgrad = True
        self._side_stream = torch.cuda.Stream(device=self.device)
        self._head_sync = torch.zeros(512, dtype=torch.int32, device=self.device)   # grid-barrier state of the fused head
        self._num_sms = torch.cuda.get_device_properties(self.device).multi_processor_count

    def _head_fused(self):
        """The two segment layers run as one persistent launch each way (lbx_head_fwd / lbx_head_bwd) in bf16 precision;
        LBX_HEAD_FUSED=0 restores one tcgen05 GEMM launch per layer and direction."""
        n = len(self.frames)
        return (self.precision == "bf16" and len(self.segments) == 2 and os.environ.get("LBX_HEAD_FUSED", "1") != "0" and
                all(self.layers[n + i]["relu"] and self.layers[n + i]["N"] % 8 == 0 and self.layers[n + i]["K"] % 8 == 0
                    for i in range(2)))

    def head_health(self):
        """Raises if a grid barrier of the fused head ever timed out (synchronises the device)."""
        if int(self._head_sync[2].item()) != 0:
            raise _lib.LidboxB200Error("fused head: grid barrier time-out")

    # ------------------------------------------------------------------ parameters
    def _build_params(self, seed):
        # (name, K, N, K_real) in execution order; weights are Keras kernels flattened to [K, N]
        self.layers = []
        c_in, c_in_real = self.Fp, self.F
        for f in self.frames:
            self.layers.append(dict(name=f.name, kind="frame", K=f.kernel_size * c_in, N=f.units, k=f.kernel_size,
                                    s=f.strides, c_in=c_in, c_in_real=c_in_real, relu=f.activation == "relu",
                                    d=getattr(f, "dilation_rate", 1), bn=None))
            if getattr(f, "batch_norm", False):
                # Keras BatchNormalization state (gamma, beta, moving_mean, moving_variance) and the folded inference
                # affine y = x * scale + shift applied by the GEMM epilogue after the activation
                dev = self.device
                self.layers[-1]["bn"] = dict(gamma=torch.ones(f.units, device=dev), beta=torch.zeros(f.units, device=dev),
                                             moving_mean=torch.zeros(f.units, device=dev),
                                             moving_variance=torch.ones(f.units, device=dev),
                                             scale=torch.ones(f.units, device=dev), shift=torch.zeros(f.units, device=dev))
                self._fold_bn(self.layers[-1])
            c_in = c_in_real = f.units
        d_in = 2 * c_in
        for sgm in self.segments:
            self.layers.append(dict(name=sgm.name, kind="dense", K=d_in, N=sgm.units, relu=sgm.activation == "relu"))
            d_in = sgm.units
        self.layers.append(dict(name=self.output_name, kind="dense", K=d_in, N=self.num_outputs, relu=False))
        # one flat fp32 buffer holds every kernel as [K, ldw] (Keras layout, pitch padded to 8 columns) and every
        # bias (padded to 8); padding stays zero.  Gradients, Adam moments and the bf16 operand copy mirror it, so the
        # optimizer is one elementwise pass and the gradient is one all-reduce.
        off = 0
        for ly in self.layers:
            ly["ldw"] = _ceil8(ly["N"])
            ly["w_off"] = off
            ly["b_off"] = off + ly["K"] * ly["ldw"]
            off = ly["b_off"] + ly["ldw"]
            assert ly["K"] % 8 == 0 and ly["w_off"] % 8 == 0
        self.n_params_padded = off
        dev = self.device
        self.params = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grads = torch.zeros(off, dtype=torch.float32, device=dev)
        self.w16 = torch.zeros(off, dtype=torch.bfloat16, device=dev)      # bf16 operand copy (hi plane)
        self.w16_lo = None                                                 # residual plane, fp32 (bf16x3) mode only
        gen = torch.Generator().manual_seed(0 if seed is None else int(seed))
        for ly in self.layers:           # Keras defaults: glorot-uniform kernels, zero biases
            if ly["kind"] == "frame":
                k, ci, co = ly["k"], ly["c_in_real"], ly["N"]
                limit = math.sqrt(6.0 / (k * ci + k * co))
                w = (torch.rand((k, ci, co), generator=gen) * 2 - 1) * limit
                wp = torch.zeros((k, ly["c_in"], co))
                wp[:, :ci] = w
                self._w_view(ly)[:, :co].copy_(wp.reshape(ly["K"], co))
            else:
                limit = math.sqrt(6.0 / (ly["K"] + ly["N"]))
                self._w_view(ly)[:, :ly["N"]].copy_((torch.rand((ly["K"], ly["N"]), generator=gen) * 2 - 1) * limit)
        self._weights_dirty, self._lo_dirty = True, True

    @staticmethod
    def _fold_bn(ly):
        bn = ly["bn"]
        bn["scale"].copy_(bn["gamma"] / torch.sqrt(bn["moving_variance"] + BN_EPSILON))
        bn["shift"].copy_(bn["beta"] - bn["moving_mean"] * bn["scale"])

    def _w_view(self, ly, buf=None):
        buf = self.params if buf is None else buf
        return buf[ly["w_off"]:ly["w_off"] + ly["K"] * ly["ldw"]].view(ly["K"], ly["ldw"])

    def _b_view(self, ly):
        return self.params[ly["b_off"]:ly["b_off"] + ly["N"]]

    def count_params(self):
        return sum((ly["k"] * ly["c_in_real"] if ly["kind"] == "frame" else ly["K"]) * ly["N"] + ly["N"]
                   for ly in self.layers)

    def get_weights(self):
        """dict name/kernel|bias -> numpy array in Keras layouts."""
        self._sync_master_from_peers()
        out = {}
        for ly in self.layers:
            w = self._w_view(ly)[:, :ly["N"]].cpu()
            if ly["kind"] == "frame":
                w = w.view(ly["k"], ly["c_in"], ly["N"])[:, :ly["c_in_real"]]
            out[ly["name"] + "/kernel"] = w.numpy().copy()
            out[ly["name"] + "/bias"] = self._b_view(ly).cpu().numpy().copy()
            if ly.get("bn"):
                for key in ("gamma", "beta", "moving_mean", "moving_variance"):
                    out[ly["name"] + "_bn/" + key] = ly["bn"][key].cpu().numpy().copy()
        return out

    def set_weights(self, weights):
        """Load reference-format weights (dict as returned by get_weights; Keras layouts)."""
        for ly in self.layers:
            w = torch.as_tensor(np.asarray(weights[ly["name"] + "/kernel"]), dtype=torch.float32)
            if ly["kind"] == "frame":
                if tuple(w.shape) != (ly["k"], ly["c_in_real"], ly["N"]):
                    raise ValueError("bad kernel shape for %s: %s" % (ly["name"], tuple(w.shape)))
                wp = torch.zeros((ly["k"], ly["c_in"], ly["N"]))
                wp[:, :ly["c_in_real"]] = w
                w = wp.reshape(ly["K"], ly["N"])
            elif tuple(w.shape) != (ly["K"], ly["N"]):
                raise ValueError("bad kernel shape for %s: %s" % (ly["name"], tuple(w.shape)))
            self._w_view(ly)[:, :ly["N"]].copy_(w)
            self._b_view(ly).copy_(torch.as_tensor(np.asarray(weights[ly["name"] + "/bias"]), dtype=torch.float32))
            if ly.get("bn"):
                for key in ("gamma", "beta", "moving_mean", "moving_variance"):
                    if ly["name"] + "_bn/" + key in weights:
                        ly["bn"][key].copy_(torch.as_tensor(np.asarray(weights[ly["name"] + "_bn/" + key]),
                                                            dtype=torch.float32))
                self._fold_bn(ly)
        self._weights_dirty = self._lo_dirty = True

    def _refresh(self, need_lo):
        """(Re)build the bf16 operand copy of the parameters (and the residual plane for the bf16x3 mode)."""
        if need_lo:
            self._sync_master_from_peers()
        if need_lo and self.w16_lo is None:
            self.w16_lo = torch.zeros_like(self.w16)
            self._lo_dirty = True
        if not (self._weights_dirty or (need_lo and self._lo_dirty)):
            return
        _lib.check(_lib.lib().lbx_split_bf16(_lib.ptr(self.params), self.params.numel(), _lib.ptr(self.w16),
                                             _lib.ptr(self.w16_lo) if need_lo else None,
                                             _lib.stream_ptr(self.device)))
        self._weights_dirty = False
        self._lo_dirty = not need_lo

    # ------------------------------------------------------------------ buffers
    def _buffers(self, B, T, training):
        key = (B, T, self.precision, bool(training))
        bufs = self._bufs.pop(key, None)
        if bufs is not None:
            self._bufs[key] = bufs                 # most recently used last
            return bufs
        # least-recently-used eviction; sets captured by a CUDA graph (GraphedTrainStep) or handed out as a feature
        # sink are pinned: their raw pointers live on outside this cache
        for k in [k for k, v in self._bufs.items() if not v.get("pinned")][:max(0, len(self._bufs) - 4)]:
            del self._bufs[k]
        geo = _Geometry(T, self.frames)
        slack = max([_SLACK_ROWS] + geo.pad)         # shifted views (dilated taps) read up to pad rows past the last utterance
        dev, bf = self.device, torch.bfloat16
        split = self.precision == "fp32"
        n = len(self.frames)
        X, X_lo = [], []
        for L in range(n):
            c = self.layers[L]["c_in"]
            X.append(torch.zeros((B * geo.Tpad[L] + slack, c), dtype=bf, device=dev))
            X_lo.append(torch.zeros_like(X[-1]) if split else None)
        cn = self.layers[n - 1]["N"]
        cnp = _ceil8(cn)
        Y = torch.zeros((B * geo.R[n - 1] + slack, cnp), dtype=torch.float32 if split else bf, device=dev)
        bufs = dict(geo=geo, X=X, X_lo=X_lo, X0_own=X[0], X0_lo_own=X_lo[0], Y=Y, cn=cn, cnp=cnp,
                    pooled=torch.zeros((B, 2 * cn), dtype=torch.float32, device=dev),
                    var_raw=torch.zeros((B, cn), dtype=torch.float32, device=dev),
                    pooled_hi=torch.zeros((B, 2 * cn), dtype=bf, device=dev),
                    pooled_lo=torch.zeros((B, 2 * cn), dtype=bf, device=dev) if split else None,
                    H=[], H_lo=[], emb=torch.zeros((B, self.segments[0].units), dtype=torch.float32, device=dev),
                    logits=torch.zeros((B, self.num_outputs), dtype=torch.float32, device=dev),
                    out=torch.zeros((B, self.num_outputs), dtype=torch.float32, device=dev))
        if not split and self.segments:
            # split-K slabs of the fused head's first layer: as many as lbx_head_fwd can use (one 64 x 128 tile per SM)
            head_tiles = -(-B // 64) * -(-self.segments[0].units // 128)
            slabs = max(1, min(16, self._num_sms // head_tiles))
            bufs["head_scratch"] = torch.zeros((slabs, B, self.segments[0].units), dtype=torch.float32, device=dev)
        for sgm in self.segments:
            bufs["H"].append(torch.zeros((B, sgm.units), dtype=bf, device=dev))
            bufs["H_lo"].append(torch.zeros((B, sgm.units), dtype=bf, device=dev) if split else None)
        if training:
            # gradients w.r.t. each layer's (masked) pre-activation, in the geometry of that layer's output buffer
            bufs["dZ"] = [torch.zeros_like(X[L + 1]) for L in range(n - 1)] + [torch.zeros_like(Y)]
            bufs["gpool"] = torch.zeros((B, 2 * cn), dtype=torch.float32, device=dev)
            bufs["dH"] = [torch.zeros_like(h) for h in bufs["H"]]
            npad = _ceil8(self.num_outputs)
            bufs["dlogits"] = torch.zeros((B, npad), dtype=bf, device=dev)
            bufs["loss"] = torch.zeros((B,), dtype=torch.float32, device=dev)
            bufs["z"] = torch.zeros((B, self.num_outputs), dtype=torch.float32, device=dev)
        self._bufs[key] = bufs
        return bufs

    # ------------------------------------------------------------------ forward
    def feature_sink(self, B, T, training=False, slot=0):
        """Buffer handle for direct feature hand-off (see FeatureSink).  The buffer set stays allocated (pinned).
        slot > 0 gives an additional, independent input buffer (double buffering of the input pipeline)."""
        if training and self.channel_dropout_rate > 0:
            raise NotImplementedError("SpatialDropout1D is applied by the packing pass; feed [B,T,F] tensors instead")
        bufs = self._buffers(int(B), int(T), bool(training))
        bufs["pinned"] = True
        return FeatureSink(self, bufs, int(B), int(T), bool(training), slot)

    def _prepare_input(self, x):
        if isinstance(x, FeatureSink):
            if x.model is not self:
                raise ValueError("feature sink belongs to another model")
            return x
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(np.asarray(x))
        if x.dim() != 3 or x.shape[2] != self.F:
            raise ValueError("expected input of shape [B, T, %d], got %s" % (self.F, tuple(x.shape)))
        if x.shape[1] < 1:
            raise ValueError("empty time axis")
        return x.to(self.device, torch.float32).contiguous()

    def _forward(self, x, bufs, training, upto_embedding=False, skip_outputs=False):
        lib, st = _lib.lib(), _lib.stream_ptr(self.device)
        geo = bufs["geo"]
        B, T, _ = x.shape
        if isinstance(x, FeatureSink):
            if x.bufs is not bufs:
                raise ValueError("feature sink was created with training=%s" % x.training)
            bufs["X"][0], bufs["X_lo"][0] = x.x0, x.x0_lo       # this call (and its backward) reads the sink's buffer
        else:
            bufs["X"][0], bufs["X_lo"][0] = bufs["X0_own"], bufs["X0_lo_own"]
        split = self.precision == "fp32"
        self._refresh(need_lo=split)
        n = len(self.frames)
        rate = self.channel_dropout_rate if training else 0.0
        if not isinstance(x, FeatureSink):
            # the dropout mask is keyed on (seed, device-side call counter): the counter is advanced by a kernel, so
            # replays of a captured CUDA graph draw a fresh mask every step (a host-side counter would be baked in)
            _lib.check(lib.lbx_pack_rows_bf16(_lib.ptr(x), B, T, self.F, _lib.ptr(bufs["X"][0]),
                                              _lib.ptr(bufs["X_lo"][0]), geo.Tpad[0], geo.pad[0], self.Fp, rate,
                                              self._dropout_seed, _lib.ptr(self._dropout_counter) if rate > 0 else None,
                                              st))
            if rate > 0:
                _lib.check(lib.lbx_counter_tick(_lib.ptr(self._dropout_counter), st))
        for L in range(n):
            ly = self.layers[L]
            last = L == n - 1
            out = bufs["Y"] if last else bufs["X"][L + 1]
            out_lo = None if (last or not split) else bufs["X_lo"][L + 1]
            ldo = bufs["cnp"] if last else ly["N"]
            out_off = 0 if last else geo.pad[L + 1] * ly["N"]
            bn = ly.get("bn")
            post = dict(post_scale=bn["scale"], post_shift=bn["shift"]) if bn else {}
            if bn and training:
                raise NotImplementedError("batch_norm frame layers are inference-only (moving statistics)")
            if ly["d"] > 1:
                # dilated taps do not form one contiguous row of the padded input: one accumulating pass per tap, tap j
                # reads the rows j*d further down against rows [j*C_in, (j+1)*C_in) of the kernel
                if split:
                    raise NotImplementedError("dilated frame layers run in precision='bf16' (the bf16x3 mode already "
                                              "uses its accumulating passes for the residual planes)")
                c = ly["c_in"]
                ops.gemm(bufs["X"][L], B * geo.R[L], c, c, self.w16, c, ly["N"], ly["ldw"], out, ldo, layout=2,
                         b_off=ly["w_off"], b_map_rows=ly["K"], terms=[(0, 0, j * ly["d"], j * c) for j in range(ly["k"])],
                         out_off=out_off, bias=self._b_view(ly), relu=ly["relu"], rows_per_utt=geo.R[L],
                         valid_rows=geo.T[L + 1], **post)
                continue
            # small batches: when the 256 x 256 CTA-pair tiles would occupy less than half of the SMs (e.g. frame3 / frame4
            # at 64 x 2 s: 18 pair tiles on 74 pairs), 128 x 128 or 128 x 64 single-CTA tiles put 4 or 8 times as many CTAs
            # to work
            m_tiles = -(-(B * geo.R[L]) // 128)
            pair_units = -(-m_tiles // 2) * -(-ly["N"] // 256)
            tile_n = 64 if 8 * pair_units <= self._num_sms else (128 if 4 * pair_units <= self._num_sms else 0)
            ops.gemm(bufs["X"][L], B * geo.R[L], ly["K"], ly["s"] * ly["c_in"], self.w16, ly["K"], ly["N"], ly["ldw"],
                     out, ldo, layout=2, a_lo=bufs["X_lo"][L], b_lo=self.w16_lo if split else None, b_off=ly["w_off"],
                     out_lo=out_lo, out_off=out_off, bias=self._b_view(ly), relu=ly["relu"], rows_per_utt=geo.R[L],
                     valid_rows=geo.T[L + 1], tile_n=tile_n, **post)
        _lib.check(lib.lbx_stats_pool_fwd(_lib.ptr(bufs["Y"]), ops.F32 if split else ops.BF16, B, geo.R[n - 1],
                                          geo.T[n], bufs["cn"], bufs["cnp"], STDDEV_SQRT_MIN_CLIP,
                                          _lib.ptr(bufs["pooled"]), _lib.ptr(bufs["var_raw"]),
                                          _lib.ptr(bufs["pooled_hi"]), _lib.ptr(bufs["pooled_lo"]), st))
        a, a_lo = bufs["pooled_hi"], bufs["pooled_lo"]
        fused = self._head_fused() and not upto_embedding
        if fused:
            l1, l2 = self.layers[n], self.layers[n + 1]
            _lib.check(lib.lbx_head_fwd(_lib.ptr(a), B, l1["K"], ops._addr(self.w16, l1["w_off"]), l1["ldw"],
                                        ops._addr(self.params, l1["b_off"]), l1["N"],
                                        ops._addr(self.w16, l2["w_off"]), l2["ldw"], ops._addr(self.params, l2["b_off"]),
                                        l2["N"], _lib.ptr(bufs["H"][0]), _lib.ptr(bufs["H"][1]),
                                        _lib.ptr(bufs["head_scratch"]), bufs["head_scratch"].numel(),
                                        _lib.ptr(self._head_sync), st))
            a, a_lo = bufs["H"][1], None
        for i, sgm in enumerate(self.segments):
            if fused:
                break
            ly = self.layers[n + i]
            if i == 0 and upto_embedding:
                self._dense(bufs, a, a_lo, B, ly, relu=False, out_f32=bufs["emb"])
                return bufs["emb"]
            self._dense(bufs, a, a_lo, B, ly, relu=ly["relu"], out_hi=bufs["H"][i], out_lo=bufs["H_lo"][i])
            a, a_lo = bufs["H"][i], bufs["H_lo"][i]
        if skip_outputs:          # the fused training head computes the output layer together with the loss
            return None
        ly = self.layers[-1]
        self._dense(bufs, a, a_lo, B, ly, relu=False, out_f32=bufs["logits"])
        return bufs["logits"]

    def _dense(self, bufs, a, a_lo, B, ly, relu, out_hi=None, out_lo=None, out_f32=None):
        """Dense layer for a small number of rows: 128x64 output tiles give enough CTAs without split-K, so bias, ReLU
        and the bf16 split stay fused in the GEMM epilogue (one launch per layer)."""
        split = self.precision == "fp32"
        if out_hi is not None:
            ops.gemm(a, B, ly["K"], ly["K"], self.w16, ly["K"], ly["N"], ly["ldw"], out_hi, ly["N"], layout=2, a_lo=a_lo,
                     b_lo=self.w16_lo if split else None, b_off=ly["w_off"], out_lo=out_lo, bias=self._b_view(ly),
                     relu=relu, tile_n=64)
        if out_f32 is not None:
            # a long contraction over few rows (the embedding layer: K = 3000, x3 passes in the fp32 mode, on 8 tiles
            # took 42 us of the 195 us of BASELINE config 2): split-K over the idle SMs, fp32 partial sums added
            # atomically into the zeroed output, the bias carried by the first split
            tiles = -(-B // 128) * -(-ly["N"] // 64)
            kb = -(-ly["K"] // 64)
            ks = min(kb // 4, self._num_sms // tiles) if not relu else 1
            if ks >= 2:
                out_f32.zero_()
                ops.gemm(a, B, ly["K"], ly["K"], self.w16, ly["K"], ly["N"], ly["ldw"], out_f32, ly["N"], layout=2,
                         a_lo=a_lo, b_lo=self.w16_lo if split else None, b_off=ly["w_off"], bias=self._b_view(ly),
                         tile_n=64, k_splits=ks, epi_atomic=True)
            else:
                ops.gemm(a, B, ly["K"], ly["K"], self.w16, ly["K"], ly["N"], ly["ldw"], out_f32, ly["N"], layout=2,
                         a_lo=a_lo, b_lo=self.w16_lo if split else None, b_off=ly["w_off"], bias=self._b_view(ly),
                         relu=relu, tile_n=64)

    def __call__(self, x, training=False):
        x = self._prepare_input(x)
        B, T, _ = x.shape
        bufs = self._buffers(B, T, False)
        lib, st = _lib.lib(), _lib.stream_ptr(self.device)
        logits = self._forward(x, bufs, training)
        if self.head == "none":
            return logits.clone()
        if self.head == "l2_normalize":
            dummy = torch.zeros(B, dtype=torch.int32, device=self.device)
            z = torch.empty_like(logits)
            _lib.check(lib.lbx_ap_loss(_lib.ptr(logits), _lib.ptr(dummy), B, self.num_outputs, 1, 1.0, 1, _lib.ptr(z),
                                       None, None, None, None, 0, None, 1.0, None, st))
            return z
        _lib.check(lib.lbx_logsoftmax_xent(_lib.ptr(logits), None, B, self.num_outputs, _lib.ptr(bufs["out"]), None,
                                           None, 0, 1.0, None, st))
        return bufs["out"].clone()

    predict = __call__

    # ------------------------------------------------------------------ training
    def configure_optimizer(self, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7):
        """Adam with the Keras defaults (eps = 1e-7), fp32 master weights and moments."""
        self._adam = dict(lr=lr, beta1=beta1, beta2=beta2, eps=eps, m=torch.zeros_like(self.params),
                          v=torch.zeros_like(self.params),
                          step=torch.zeros(1, dtype=torch.int64, device=self.device),
                          lr_t=torch.zeros(1, dtype=torch.float32, device=self.device))

    def loss_and_grads(self, x, y, loss="xent", ap_classes=None, delta_weight=1.0, global_batch=None,
                       process_group=None):
        """Forward + backward of one batch in bf16 (fp32 accumulation / statistics / loss).  Fills self.grads with
        d(mean loss)/d(params) (mean over `global_batch`, default this batch) and returns the per-sample losses."""
        if self.precision != "bf16":
            raise ValueError("training runs in precision='bf16' (fp32 master weights); create the model with it")
        x = self._prepare_input(x)
        B, T, _ = x.shape
        y = torch.as_tensor(y).to(self.device, torch.int32).reshape(-1).contiguous()   # [B,1] labels are squeezed
        if y.numel() != B:
            raise ValueError("labels must have one entry per sample")
        bufs = self._buffers(B, T, True)
        geo = bufs["geo"]
        lib, st = _lib.lib(), _lib.stream_ptr(self.device)
        n = len(self.frames)
        out_ly = self.layers[-1]
        n_seg = len(self.segments)
        # few classes + cross-entropy: output layer, loss and the layer's whole backward run as ONE launch
        # (lbx_dense_xent_head) instead of GEMM + loss + 2 GEMMs (LBX_FUSED_HEAD=0 disables).  Together with the fused
        # segment layers (lbx_head_fwd / lbx_head_bwd) the dense head is 3 launches instead of 11.
        fused_head = (loss == "xent" and self.num_outputs <= 8 and n_seg >= 1 and out_ly["K"] <= 1024 and
                      out_ly["K"] * (4 * self.num_outputs + 20) + 256 <= 49152 and
                      os.environ.get("LBX_FUSED_HEAD", "1") != "0")
        cur = torch.cuda.current_stream(self.device)
        ev_zero = None
        if self._sharded is not None:
            # data-parallel step start: the previous lbx_adam_step_sharded did not wait for the peers' all-gather pushes
            # nor clear the gradient.  A side stream waits for "every peer has published" (tiny kernel, normally already
            # true), which the forward pass needs (complete weights); after it the gradient buffer is cleared while
            # the forward pass runs (first written by the loss kernel).
            sh = self._sharded
            ev0 = torch.cuda.Event()
            ev0.record(cur)
            self._side_stream.wait_event(ev0)
            with torch.cuda.stream(self._side_stream):
                _lib.check(lib.lbx_dp_wait(_lib.ptr(sh["sig"]), sh["world"], _lib.ptr(sh["epoch"]), _lib.ptr(sh["local"]),
                                           _lib.stream_ptr(self.device)))
                ev_w = torch.cuda.Event()
                ev_w.record(self._side_stream)
                self.grads.zero_()
                ev_zero = torch.cuda.Event()
                ev_zero.record(self._side_stream)
            cur.wait_event(ev_w)
            self._grads_clean = True
        logits = self._forward(x, bufs, True, skip_outputs=fused_head)
        scale = 1.0 / float(global_batch or B)
        npad = bufs["dlogits"].shape[1]
        if not self._grads_clean:
            self.grads.zero_()
        if ev_zero is not None:
            cur.wait_event(ev_zero)
        self._grads_clean = False
        g = self.grads
        if fused_head:
            below = self.layers[n + n_seg - 1]
            _lib.check(lib.lbx_dense_xent_head(
                _lib.ptr(bufs["H"][-1]), ops._addr(self.w16, out_ly["w_off"]), ops._addr(self.params, out_ly["b_off"]),
                _lib.ptr(y), B, out_ly["K"], self.num_outputs, out_ly["K"], out_ly["ldw"], scale,
                1 if below["relu"] else 0, _lib.ptr(bufs["logits"]), _lib.ptr(bufs["loss"]), _lib.ptr(bufs["dH"][-1]),
                ops._addr(g, out_ly["w_off"]), ops._addr(g, out_ly["b_off"]), ops._addr(g, below["b_off"]), st))
        elif loss == "xent":
            _lib.check(lib.lbx_logsoftmax_xent(_lib.ptr(logits), _lib.ptr(y), B, self.num_outputs, None,
                                               _lib.ptr(bufs["loss"]), _lib.ptr(bufs["dlogits"]), npad, scale,
                                               ops._addr(g, out_ly["b_off"]), st))
        elif loss == "ap":
            N = int(ap_classes or self.num_outputs)
            _lib.check(lib.lbx_ap_loss(_lib.ptr(logits), _lib.ptr(y), B, self.num_outputs, N, float(delta_weight), 1,
                                       _lib.ptr(bufs["z"]), None, _lib.ptr(bufs["loss"]), None,
                                       _lib.ptr(bufs["dlogits"]), npad, None, scale, ops._addr(g, out_ly["b_off"]),
                                       st))
        else:
            raise ValueError("loss must be 'xent' or 'ap'")

        # weight gradients are leaves of the backward graph (only the optimizer reads them): they run on a side
        # stream, concurrently with the latency-bound chain of data-gradient kernels, and are joined before Adam
        side = self._side_stream if self.overlap_wgrad else None
        side_used = [False]

        # frame-layer weight gradients: collected and issued as ONE grouped launch after the data-gradient chain
        # (lbx_wgrad_grouped: the k-blocks of all layers cut into equal ranges, one per CTA pair); LBX_WGRAD_GROUPED=0
        # restores one split-K GEMM per layer on the side stream
        dp_mode = os.environ.get("LBX_DP_MODE", "single")      # single | buckets | none (measurement only)
        grouped = [] if (os.environ.get("LBX_WGRAD_GROUPED", "1") != "0" and dp_mode != "buckets") else None

        def flush_grouped():
            if grouped:
                # largest problems first: the ranges of the line that end up shortest in wall time come last
                grouped.sort(key=lambda q: -q["rows"] * q["a_cols"] * q["b_cols"])
                for i in range(0, len(grouped), 8):
                    ops.wgrad_grouped(grouped[i:i + 8], self.device)
                del grouped[:]

        def wgrad(a, a_rows, a_cols, lda, dz, dz_cols, dz_pitch, ly, a_off=0, dz_off=0, group=False):
            if group and grouped is not None:
                grouped.append(dict(a=a, rows=a_rows, a_cols=a_cols, lda=lda, a_off=a_off, b=dz, b_cols=dz_cols,
                                    ldb=dz_pitch, b_off=dz_off, out=g, ldo=ly["ldw"], out_off=ly["w_off"]))
                return
            tiles = -(-a_cols // 128) * -(-dz_cols // 256)
            ks = max(1, min(-(-a_rows // 64), 148 // tiles))
            if side is not None:
                side_used[0] = True
                ev = torch.cuda.Event()
                ev.record(cur)
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    ops.gemm(a, a_rows, a_cols, lda, dz, a_rows, dz_cols, dz_pitch, g, ly["ldw"], layout=1,
                             a_off=a_off, b_off=dz_off, out_off=ly["w_off"], k_splits=ks, epi_atomic=True)
                return
            ops.gemm(a, a_rows, a_cols, lda, dz, a_rows, dz_cols, dz_pitch, g, ly["ldw"], layout=1, a_off=a_off,
                     b_off=dz_off, out_off=ly["w_off"], k_splits=ks, epi_atomic=True)

        # data-parallel exchange.  Default ("single"): ONE all-reduce of the flat gradient after the backward pass.
        # "buckets" all-reduces three buckets on the side stream as soon as each is complete; measured on 2 x B200 it is
        # SLOWER (0.712 vs 0.619 ms/step; no exchange: 0.572): the NCCL kernels take SMs away from the persistent
        # one-CTA-per-SM GEMMs they overlap with.  Kept for experiments (LBX_DP_MODE=buckets).

        def reduce_bucket(first_layer, last_layer):
            if process_group is None or dp_mode != "buckets":
                return
            import torch.distributed as dist
            lo = self.layers[first_layer]["w_off"]
            hi = self.layers[last_layer]["b_off"] + self.layers[last_layer]["ldw"]
            if side is not None:
                side_used[0] = True
                ev = torch.cuda.Event()
                ev.record(cur)                     # bias gradients of the bucket come from main-stream kernels
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    dist.all_reduce(g[lo:hi], group=process_group)
            else:
                dist.all_reduce(g[lo:hi], group=process_group)

        mid = n // 2
        # ---- dense head (bias gradients come fused out of the kernels that produce each dz) ----
        dz, dz_cols, dz_pitch = bufs["dlogits"], self.num_outputs, npad
        acts = [bufs["pooled_hi"]] + bufs["H"]
        first = n_seg
        if fused_head:             # the output layer is done: continue from the gradient w.r.t. the last segment layer
            first = n_seg - 1
            dz, dz_cols, dz_pitch = bufs["dH"][-1], out_ly["K"], out_ly["K"]
        head_fused = self._head_fused()
        for i in range(first, -1, -1):
            if head_fused and i == 1:
                # both segment layers' backward in one persistent launch: dH1 (+ its bias gradient), dW2, d pooled, dW1
                l1, l2 = self.layers[n], self.layers[n + 1]
                _lib.check(lib.lbx_head_bwd(_lib.ptr(dz), _lib.ptr(bufs["pooled_hi"]), _lib.ptr(bufs["H"][0]), B, l1["K"],
                                            l1["N"], l2["N"], ops._addr(self.w16, l1["w_off"]), l1["ldw"],
                                            ops._addr(self.w16, l2["w_off"]), l2["ldw"], _lib.ptr(bufs["dH"][0]),
                                            _lib.ptr(bufs["gpool"]), ops._addr(g, l1["w_off"]), ops._addr(g, l1["b_off"]),
                                            ops._addr(g, l2["w_off"]), _lib.ptr(self._head_sync), st))
                break
            ly = self.layers[n + i]
            wgrad(acts[i], B, ly["K"], ly["K"], dz, dz_cols, dz_pitch, ly)
            if i > 0:      # d hidden = (dz . W^T) masked by the ReLU of the layer below (+ its bias gradient), one launch
                below = self.layers[n + i - 1]
                ops.gemm(dz, B, dz_cols, dz_pitch, self.w16, ly["K"], ly["N"], ly["ldw"], bufs["dH"][i - 1], ly["K"],
                         b_off=ly["w_off"], mask_src=bufs["H"][i - 1] if below["relu"] else None, colsum=g,
                         colsum_off=below["b_off"], colsum_mod=ly["K"], tile_n=64)
                dz, dz_cols, dz_pitch = bufs["dH"][i - 1], ly["K"], ly["K"]
            else:          # d pooled (fp32, no mask)
                ops.gemm(dz, B, dz_cols, dz_pitch, self.w16, ly["K"], ly["N"], ly["ldw"], bufs["gpool"], ly["K"],
                         b_off=ly["w_off"], tile_n=64)
        reduce_bucket(n, len(self.layers) - 1)
        # ---- statistics pooling (+ ReLU mask and bias gradient of the last frame layer) ----
        last = self.layers[n - 1]
        if not last["relu"]:
            raise NotImplementedError("the pooling backward pass applies the ReLU mask of the last frame layer: a linear "
                                      "last frame layer is not supported in training")
        _lib.check(lib.lbx_stats_pool_bwd(_lib.ptr(bufs["Y"]), B, geo.R[n - 1], geo.T[n], bufs["cn"], bufs["cnp"],
                                          STDDEV_SQRT_MIN_CLIP, _lib.ptr(bufs["pooled"]), _lib.ptr(bufs["var_raw"]),
                                          _lib.ptr(bufs["gpool"]), _lib.ptr(bufs["dZ"][n - 1]),
                                          ops._addr(g, last["b_off"]), 0, st))
        # ---- frame layers, last to first ----
        for L in range(n - 1, -1, -1):
            ly = self.layers[L]
            rows = B * geo.R[L]
            dZ = bufs["dZ"][L]
            dz_pitch = bufs["cnp"] if L == n - 1 else ly["N"]
            dz_off = 0 if L == n - 1 else geo.pad[L + 1] * ly["N"]
            if ly["d"] > 1:
                # one weight-gradient problem per tap: dW_j += X[t + j*d]^T . dZ[t]  (rows shifted by j*d, kernel rows j*C_in..)
                c = ly["c_in"]
                for j in range(ly["k"]):
                    wgrad(bufs["X"][L], rows, c, c, dZ, ly["N"], dz_pitch,
                          dict(ldw=ly["ldw"], w_off=ly["w_off"] + j * c * ly["ldw"]), a_off=j * ly["d"] * c, dz_off=dz_off,
                          group=True)
            else:
                wgrad(bufs["X"][L], rows, ly["K"], ly["s"] * ly["c_in"], dZ, ly["N"], dz_pitch, ly, dz_off=dz_off,
                      group=True)
            if L == mid and mid > 0:
                reduce_bucket(mid, n - 1)
            if L == 0:
                reduce_bucket(0, max(mid - 1, 0) if mid > 0 else n - 1)
                break
            # data gradient through the same overlapping view, masked by the ReLU of layer L-1 (= X[L] > 0; padding
            # and junk rows of X[L] are zero, so they stay zero in dZ[L-1]); the column sums of what is written are
            # the bias gradient of layer L-1 (column n of the view is channel n mod C_in)
            k, s, c = ly["k"], ly["s"], ly["c_in"]
            below = self.layers[L - 1]
            if not below["relu"]:
                raise NotImplementedError("linear frame layers are not supported in the backward pass")
            if ly["d"] > 1:
                # dilated, stride 1: padded input time tau receives tap j of output time tau - j*d: ONE GEMM whose k
                # accumulating passes read dZ shifted by -j*d rows against kernel rows [j*C_in, (j+1)*C_in)
                ops.gemm(dZ, rows, ly["N"], dz_pitch, self.w16, c, ly["N"], ly["ldw"], bufs["dZ"][L - 1], c,
                         a_off=dz_off, b_off=ly["w_off"], b_map_rows=k * c,
                         terms=[(0, 0, -j * ly["d"], j * c) for j in range(k)], mask_src=bufs["X"][L], colsum=g,
                         colsum_off=below["b_off"], colsum_mod=c)
            elif k <= s:
                # taps tile the time axis without overlap: one GEMM writes all k*C_in columns of every view row
                ops.gemm(dZ, rows, ly["N"], dz_pitch, self.w16, k * c, ly["N"], ly["ldw"], bufs["dZ"][L - 1], s * c,
                         a_off=dz_off, b_off=ly["w_off"], mask_src=bufs["X"][L], colsum=g, colsum_off=below["b_off"],
                         colsum_mod=c)
            else:
                # gather form: padded output time tau = t*s + j receives tap j of output row t.  View row m of the
                # destination holds the s*C_in values of times m*s .. m*s + s - 1; accumulating pass i reads dZ shifted
                # by -i rows against the kernel rows of taps i*s .. i*s + s - 1 (rows past k*C_in read as zeros), so the
                # whole data gradient is ONE GEMM with N = s*C_in (round 1: one launch per residue class of tau)
                passes = -(-k // s)
                if passes > 5:
                    raise NotImplementedError("kernel_size > 5 * strides in the backward pass")
                def limit(i):
                    # pass i covers taps i*s .. min(k, (i+1)*s) - 1 = output columns [0, (that many) * C_in): tiles to
                    # the right of it would only multiply by the zero fill — skip them when tile-aligned
                    cols = (min(k, (i + 1) * s) - i * s) * c
                    return cols if (i > 0 and cols < s * c and cols % 256 == 0) else 0

                if os.environ.get("LBX_DGRAD_MERGED", "1") != "0":
                    ops.gemm(dZ, rows, ly["N"], dz_pitch, self.w16, s * c, ly["N"], ly["ldw"], bufs["dZ"][L - 1], s * c,
                             a_off=dz_off, b_off=ly["w_off"], b_map_rows=k * c,
                             terms=[(0, 0, -i, i * s * c, limit(i)) for i in range(passes)], mask_src=bufs["X"][L], colsum=g,
                             colsum_off=below["b_off"], colsum_mod=c)
                else:
                    for rho in range(s):
                        taps = list(range(rho, k, s))
                        if not taps:
                            continue
                        ops.gemm(dZ, rows, ly["N"], dz_pitch, self.w16, c, ly["N"], ly["ldw"], bufs["dZ"][L - 1], s * c,
                                 a_off=dz_off, b_off=ly["w_off"], b_map_rows=k * c,
                                 terms=[(0, 0, -i, t * c) for i, t in enumerate(taps)], out_off=rho * c,
                                 mask_src=bufs["X"][L], mask_off=rho * c, colsum=g, colsum_off=below["b_off"],
                                 colsum_mod=c)
            if self._sharded is not None and self._sharded["early_on"] and L == self._sharded["early_layer"]:
                # every gradient at flat index >= early_begin is now final on this rank (weights of layers >= L: issue
                # them now; their biases came out of the data-gradient epilogues / pooling / head kernels above)
                flush_grouped()
                if side is not None and side_used[0]:
                    ev = torch.cuda.Event()
                    ev.record(side)
                    cur.wait_event(ev)
                self._dp_early_exchange(cur)
        flush_grouped()
        if side is not None and side_used[0]:
            ev = torch.cuda.Event()
            ev.record(side)
            cur.wait_event(ev)
        if process_group is not None and dp_mode == "single":
            import torch.distributed as dist
            dist.all_reduce(g, group=process_group)
        return bufs["loss"]

    def apply_gradients(self, grad_scale=1.0):
        """Adam on the flat fp32 buffers; the same pass refreshes the bf16 operand copy and resets the gradient."""
        if self._adam is None:
            self.configure_optimizer()
        a = self._adam
        _lib.check(_lib.lib().lbx_adam_step(_lib.ptr(self.params), _lib.ptr(self.grads), _lib.ptr(a["m"]),
                                            _lib.ptr(a["v"]), self.params.numel(), a["lr"], a["beta1"], a["beta2"],
                                            a["eps"], _lib.ptr(a["step"]), _lib.ptr(a["lr_t"]), float(grad_scale),
                                            _lib.ptr(self.w16), 1, _lib.stream_ptr(self.device)))
        self._grads_clean = True
        self._weights_dirty = False        # the hi plane was refreshed by the optimizer pass
        self._lo_dirty = True

    def enable_sharded_optimizer(self, process_group):
        """Data parallel without NCCL in the step: the flat parameter / gradient / bf16 buffers move to symmetric
        (peer-mapped) memory and the optimizer step becomes lbx_adam_step_sharded — reduce-scatter by NVLink peer
        loads, Adam on this rank's 1/R shard, all-gather by peer stores, all inside one kernel.  Collective call."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        world, rank = dist.get_world_size(process_group), dist.get_rank(process_group)
        if self._adam is None:
            self.configure_optimizer()
        dev = self.device
        n = self.params.numel()
        unit = 4 * world
        n_pad = -(-n // unit) * unit
        handles, mc_ptrs = [], []

        def make(dtype, src, length):
            t = symm.empty(length, dtype=dtype, device=dev)
            t.zero_()
            if src is not None:
                t[:src.numel()].copy_(src)
            h = symm.rendezvous(t, group=process_group)
            handles.append(h)
            mc_ptrs.append(int(getattr(h, "multicast_ptr", 0) or 0))
            return t, torch.tensor(list(h.buffer_ptrs), dtype=torch.int64, device=dev)

        self.params, p_ptrs = make(torch.float32, self.params, n_pad)
        self.grads, g_ptrs = make(torch.float32, None, n_pad)
        self.w16, w_ptrs = make(torch.bfloat16, None, n_pad)
        sig, s_ptrs = make(torch.int32, None, max(2 * world, 64))
        torch.cuda.synchronize(dev)
        dist.barrier(group=process_group)                   # every rank has zeroed its pads before anyone signals
        a = self._adam
        # in-switch reduction (multimem.ld_reduce / multimem.st) pays from 4 ranks on; at 2 ranks plain peer loads are
        # faster (measured on 2 x B200: optimizer kernel 46 us vs 64 us), so it is the default only for world > 2
        nvls_default = "1" if world > 2 else "0"
        use_nvls = os.environ.get("LBX_DP_NVLS", nvls_default) != "0" and mc_ptrs[1] != 0 and mc_ptrs[2] != 0
        self._sharded = dict(world=world, rank=rank, n=n_pad, p_ptrs=p_ptrs, g_ptrs=g_ptrs, w_ptrs=w_ptrs,
                             mc_grads=mc_ptrs[1] if use_nvls else 0, mc_w16=mc_ptrs[2] if use_nvls else 0,
                             s_ptrs=s_ptrs, sig=sig, handles=handles,
                             m=torch.zeros(n_pad // world, dtype=torch.float32, device=dev),
                             v=torch.zeros(n_pad // world, dtype=torch.float32, device=dev),
                             epoch=torch.zeros(1, dtype=torch.int32, device=dev),
                             local=torch.zeros(16, dtype=torch.int32, device=dev))
        # early exchange (opt-in, LBX_DP_EARLY=1): the gradients of the later layers (second half of the frame layers,
        # pooling-side bias, dense head: ~80 % of the parameters) are complete when the data-gradient chain has passed
        # the middle frame layer.  From then on the copy engine pulls the peers' copies of this rank's shard into local
        # staging slabs while the tensor cores finish the backward pass (lbx_dp_signal / lbx_dp_wait_slot + D2D copies
        # on a side stream); the optimizer kernel reads only the late part of its shard over NVLink.
        # Measured on 2 x B200 it is correct but SLOWER (0.461 vs 0.447 ms/step): the optimizer kernel gains 6 us, and
        # cutting the grouped weight-gradient launch in two plus the extra signal kernel / stream joins cost 11 us.
        nfr = len(self.frames)
        sh = self._sharded
        sh["early_layer"] = nfr // 2
        sh["early_begin"] = self.layers[nfr // 2]["w_off"] if nfr // 2 > 0 else n_pad
        sh["early_on"] = os.environ.get("LBX_DP_EARLY", "0") == "1" and world > 1 and sh["early_begin"] < n_pad
        if sh["early_on"]:
            shard = n_pad // world
            sh["staging"] = torch.zeros((world - 1, shard), dtype=torch.float32, device=dev)
            sh["peer_grads"] = {q: handles[1].get_buffer(q, (n_pad,), torch.float32) for q in range(world) if q != rank}
            sh["dp_stream"] = torch.cuda.Stream(device=dev)
        sh["early_ev"] = None
        a["m"] = a["v"] = None                              # full-size moments are not needed any more
        self._weights_dirty = self._lo_dirty = True
        self._grads_clean = True
        self.w16_lo = None

    def _dp_early_exchange(self, cur):
        """Announce "my gradients at flat index >= early_begin are complete" and start pulling the peers' copies of this
        rank's shard with the copy engine on a side stream (see enable_sharded_optimizer)."""
        sh = self._sharded
        lib = _lib.lib()
        world, rank = sh["world"], sh["rank"]
        _lib.check(lib.lbx_dp_signal(_lib.ptr(sh["s_ptrs"]), world, rank, 2 * world, _lib.ptr(sh["epoch"]), 1,
                                     _lib.stream_ptr(self.device)))
        ev = torch.cuda.Event()
        ev.record(cur)
        s2 = sh["dp_stream"]
        s2.wait_event(ev)
        shard = sh["n"] // world
        lo, hi = max(sh["early_begin"], rank * shard), (rank + 1) * shard
        dbg = int(os.environ.get("LBX_DP_EARLY_DEBUG", "0"))      # measurement only: 1 = no copies, 2 = no wait kernel either
        with torch.cuda.stream(s2):
            if dbg < 2:
                _lib.check(lib.lbx_dp_wait_slot(_lib.ptr(sh["sig"]), world, 2 * world, _lib.ptr(sh["epoch"]), 1,
                                                _lib.ptr(sh["local"]), _lib.stream_ptr(self.device)))
            if hi > lo and dbg == 0:
                for sl in range(world - 1):
                    q = (rank + 1 + sl) % world
                    sh["staging"][sl, lo - rank * shard:hi - rank * shard].copy_(sh["peer_grads"][q][lo:hi],
                                                                                non_blocking=True)
            sh["early_ev"] = torch.cuda.Event()
            sh["early_ev"].record(s2)

    def _apply_sharded(self):
        a, sh = self._adam, self._sharded
        staging = None
        if sh.get("early_ev") is not None:          # this step's early part sits in the staging slabs
            torch.cuda.current_stream(self.device).wait_event(sh["early_ev"])
            sh["early_ev"] = None
            staging = _lib.ptr(sh["staging"])
        _lib.check(_lib.lib().lbx_adam_step_sharded(_lib.ptr(sh["p_ptrs"]), _lib.ptr(sh["g_ptrs"]),
                                                    _lib.ptr(sh["w_ptrs"]), _lib.ptr(sh["s_ptrs"]), _lib.ptr(sh["m"]),
                                                    _lib.ptr(sh["v"]), sh["n"], sh["rank"], sh["world"],
                                                    _lib.ptr(sh["epoch"]), _lib.ptr(sh["local"]), a["lr"], a["beta1"],
                                                    a["beta2"], a["eps"], _lib.ptr(a["step"]), _lib.ptr(a["lr_t"]),
                                                    1.0, 0, ctypes.c_void_p(sh["mc_grads"] or None),
                                                    ctypes.c_void_p(sh["mc_w16"] or None), staging, sh["early_begin"],
                                                    _lib.stream_ptr(self.device)))
        self._grads_clean = False          # cleared at the start of the next step, after lbx_dp_wait
        self._weights_dirty = False
        self._lo_dirty = True
        sh["master_stale"] = True

    def dp_health(self):
        """Raises if a cross-GPU barrier of the sharded optimizer ever timed out on this rank (the update of that step
        was skipped here, so the replicas have diverged: fatal).  Synchronises the device."""
        if self._sharded is not None and int(self._sharded["local"][3].item()) != 0:
            raise _lib.LidboxB200Error("data-parallel barrier time-out: a peer rank was more than the spin budget late "
                                       "(LBX_DP_SPIN_LIMIT polls of 64 ns); the optimizer step was skipped on this rank")

    def _sync_master_from_peers(self):
        """The sharded optimizer keeps the fp32 master copy of a shard current on its owner only; pull the other
        shards through the peer mappings (needed before exporting weights or building the bf16x3 residual plane).
        Every rank must call this at the same point (it ends with a barrier)."""
        sh = self._sharded
        if sh is None or not sh.get("master_stale"):
            return
        import torch.distributed as dist
        torch.cuda.synchronize(self.device)
        self.dp_health()
        dist.barrier()
        shard = sh["n"] // sh["world"]
        h = sh["handles"][0]
        for q in range(sh["world"]):
            if q != sh["rank"]:
                peer = h.get_buffer(q, (sh["n"],), torch.float32)
                self.params[q * shard:(q + 1) * shard].copy_(peer[q * shard:(q + 1) * shard])
        torch.cuda.synchronize(self.device)
        dist.barrier()
        sh["master_stale"] = False

    def train_step(self, x, y, loss="xent", process_group=None, **kw):
        """One optimisation step; with a process group the flat fp32 gradient is sum-all-reduced over NCCL before Adam."""
        world = 1
        if process_group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(process_group)
        B = x.shape[0]
        if self._sharded is not None:
            if self._sharded["world"] != world:
                raise ValueError("the sharded optimizer was enabled for a different process group")
            losses = self.loss_and_grads(x, y, loss=loss, global_batch=B * world, process_group=None, **kw)
            self._apply_sharded()
        else:
            losses = self.loss_and_grads(x, y, loss=loss, global_batch=B * world,
                                         process_group=process_group if world > 1 else None, **kw)
            self.apply_gradients()
        # the loss buffer is reused by the next step: hand out a copy (inside a graph capture the static buffer itself)
        return losses if torch.cuda.is_current_stream_capturing() else losses.clone()


class GraphedTrainStep:
    """The whole optimisation step (optionally preceded by a caller-supplied feature stage) captured once into a
    CUDA graph and replayed: the ~60 kernel launches of a step cost one host call.  Inputs are read from the static
    tensors given at construction; copy new data into them before calling."""

    def __init__(self, model, x_static, y_static, loss="xent", process_group=None, pre=None, concurrent=None,
                 warmup=3, loss_host=None, copies=None, **kw):
        """pre: optional callable producing the features inline (same stream, before the step).
        concurrent: optional callable enqueued on a second stream alongside the step and joined at its end — e.g. the
        feature extraction of the NEXT batch (input-pipeline prefetch), which is independent of this step.
        loss_host: optional pinned host tensor [B] float32: the per-sample losses of every replay are copied into it by
        a copy node INSIDE the graph (no separate stream operation between two replays); read it after a synchronise.
        copies: optional list of (dst, src) tensor pairs copied on a third branch of the graph, concurrently with the
        step — e.g. the host-to-device transfer of the batch after next from a pinned staging buffer into a signal
        buffer that nothing in this replay reads."""
        self.model, self.x, self.y = model, x_static, y_static
        lib = _lib.lib()
        aux = torch.cuda.Stream(device=model.device) if concurrent is not None else None
        cpy = torch.cuda.Stream(device=model.device) if copies else None

        def body():
            if cpy is not None:
                cur = torch.cuda.current_stream(model.device)
                evc = torch.cuda.Event()
                evc.record(cur)
                cpy.wait_event(evc)
                with torch.cuda.stream(cpy):
                    for dst, src in copies:
                        dst.copy_(src, non_blocking=True)
            if aux is not None:
                cur = torch.cuda.current_stream(model.device)
                ev = torch.cuda.Event()
                ev.record(cur)
                aux.wait_event(ev)
                with torch.cuda.stream(aux):
                    concurrent()
            feats = pre() if pre is not None else self.x
            out = model.train_step(feats, self.y, loss=loss, process_group=process_group, **kw)
            if loss_host is not None:
                loss_host.copy_(out, non_blocking=True)
            if aux is not None:
                ev2 = torch.cuda.Event()
                ev2.record(aux)
                torch.cuda.current_stream(model.device).wait_event(ev2)
            if cpy is not None:
                ev3 = torch.cuda.Event()
                ev3.record(cpy)
                torch.cuda.current_stream(model.device).wait_event(ev3)
            return out

        side = torch.cuda.Stream(device=model.device)
        side.wait_stream(torch.cuda.current_stream(model.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                body()
        torch.cuda.current_stream(model.device).wait_stream(side)
        torch.cuda.synchronize(model.device)
        # the graph holds raw pointers into the activation / gradient buffers: keep every set alive and un-evictable
        self._keep = list(model._bufs.values())
        for bufs in self._keep:
            bufs["pinned"] = True
        self.graph = torch.cuda.CUDAGraph()
        n0 = lib.lbx_launch_count()
        with torch.cuda.graph(self.graph):
            self.losses = body()
        self.kernels_per_step = int(lib.lbx_launch_count() - n0)

    def __call__(self):
        self.graph.replay()
        return self.losses


def create(input_shape, num_outputs, channel_dropout_rate=0, name="x-vector", **kwargs):
    """lidbox/models/xvector.py:46-67.  Extra keyword arguments (precision=, head=, seed=, device=) are extensions."""
    return XVector(input_shape, num_outputs, channel_dropout_rate=channel_dropout_rate, name=name, **kwargs)


def as_embedding_extractor(m):
    """xvector.py:70-73: drops segment1's activation (MUTATES m, like the reference) and returns a callable
    mapping [B, T, F] -> [B, 512] pre-ReLU segment1 outputs."""
    n = len(m.frames)
    m.segments[0].activation = None
    m.layers[n]["relu"] = False
    return _EmbeddingExtractor(m)


class _EmbeddingExtractor:
    def __init__(self, model):
        self.model = model

    def __call__(self, x, training=False):
        m = self.model
        x = m._prepare_input(x)
        bufs = m._buffers(x.shape[0], x.shape[1], False)
        return m._forward(x, bufs, training, upto_embedding=True).clone()

    predict = __call__
