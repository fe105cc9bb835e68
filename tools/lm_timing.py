"""Quick device timing of the fused log-mel kernel (fp32 / pcm16 in, fp32 / bf16-sink out)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import _time_cuda, load_peaks, SR
from lidbox_b200.features import audio
from lidbox_b200.models import xvector
dev = torch.device("cuda", 0)
res = {}
for (B, sec) in [(2048, 5), (256, 2), (64, 2)]:
    N = sec * SR
    T = 1 + (N - 400) // 160
    x = torch.randn((B, N), device=dev) * 0.1
    pcm = (x * 32768).clamp(-32768, 32767).to(torch.int16)
    out = torch.empty((B, T, 40), dtype=torch.float32, device=dev)
    m = xvector.create((T, 40), 4, precision="bf16", seed=0)
    sink = m.feature_sink(B, T, training=False)
    it = 10 if B > 256 else 50
    r = {}
    r["f32_f32_ms"] = _time_cuda(lambda: audio.logmelspectrograms(x, SR, out=out), it)
    r["i16_f32_ms"] = _time_cuda(lambda: audio.logmelspectrograms(pcm, SR, out=out), it)
    r["f32_bf16sink_ms"] = _time_cuda(lambda: audio.logmelspectrograms(x, SR, out=sink), it)
    r["i16_bf16sink_ms"] = _time_cuda(lambda: audio.logmelspectrograms(pcm, SR, out=sink), it)
    r["spectrogram_ms"] = _time_cuda(lambda: audio.spectrograms(x[:min(B, 256)], SR), 5)
    r["frames_per_s_f32"] = B * T / (r["f32_f32_ms"] * 1e-3)
    r["hbm_frac"] = B * (4 * N + 4 * T * 40) / (r["f32_f32_ms"] * 1e-3) / 1e9 / load_peaks()["hbm_gbs"]
    res["%dx%ds" % (B, sec)] = r
    del m, sink, x, pcm, out
    torch.cuda.empty_cache()
print(json.dumps(res))
