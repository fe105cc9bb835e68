"""Generates tests/golden/*.npz from the CPU oracle (oracle/lidbox_oracle.py).

The reference itself cannot be imported here (TensorFlow is not installable), so these vectors pin the ORACLE, not
TensorFlow: they make the restatement regression-proof and carry the reference's audio fixtures to the GPU box
(which has no /root/reference).  Run from the repo root in the build container:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import scipy.io.wavfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import lidbox_oracle as O  # noqa: E402

REF_AUDIO = "/root/reference/tests/audio"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    # 1. the reference's WAV fixtures (tests/test_features_audio.py:19-24): first 0.5 s of each, int16 @ 16 kHz
    names = sorted(f for f in os.listdir(REF_AUDIO) if f.endswith(".wav"))
    pcm = []
    for n in names:
        sr, x = scipy.io.wavfile.read(os.path.join(REF_AUDIO, n))
        assert sr == 16000 and x.dtype == np.int16
        pcm.append(x[:8000])
    pcm = np.stack(pcm)
    sig = (pcm.astype(np.float32) / np.float32(32768.0))
    spec = O.spectrograms(sig, 16000, dtype=np.float64)
    mel = O.linear_to_mel(spec, 16000, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "wav_fixtures.npz"), names=np.array(names), pcm=pcm,
                        logmel=O.log_eps(mel).astype(np.float32), spec_rowsum=spec.sum(axis=2).astype(np.float32),
                        db=O.power_to_db(spec.astype(np.float32))[:, ::8, ::16])

    # 2. mel table (bug-compatible _linspace) for the default configuration + two odd ones
    tabs = {}
    for (m, k, sr, lo, hi) in [(40, 257, 16000, 0.0, 8000.0), (10, 129, 8000, 125.0, 3800.0), (85, 513, 44100, 20.0, 11025.0)]:
        tabs["W_%d_%d_%d" % (m, k, sr)] = O.linear_to_mel_weight_matrix(m, k, sr, lo, hi)
    np.savez_compressed(os.path.join(OUT, "mel_tables.npz"), **tabs)

    # 3. x-vector: seeded tiny problem, fp64 oracle outputs and gradients (torch-CPU autograd twin)
    import torch
    rng = np.random.default_rng(7)
    B, T, F, n_out = 3, 37, 24, 5
    x = rng.standard_normal((B, T, F)).astype(np.float32)
    y = np.array([0, 3, 4])
    params = O.xvector_init(F, n_out, seed=11, bias_scale=0.05)
    logp = O.xvector_forward({k: v.astype(np.float64) for k, v in params.items()}, x.astype(np.float64))
    emb = O.xvector_forward({k: v.astype(np.float64) for k, v in params.items()}, x.astype(np.float64), embedding=True)
    tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in params.items()}
    lp_t = O.torch_xvector_forward(tp, torch.tensor(x, dtype=torch.float64))
    loss = -lp_t[torch.arange(B), torch.tensor(y)].mean()
    loss.backward()
    assert np.allclose(lp_t.detach().numpy(), logp, atol=1e-10)
    np.savez_compressed(os.path.join(OUT, "xvector_small.npz"), x=x, y=y, logp=logp, emb=emb, loss=loss.item(),
                        g_frame1_kernel_s8=tp["frame1/kernel"].grad.numpy()[:, :, ::8], g_frame5_bias=tp["frame5/bias"].grad.numpy(),
                        g_segment1_kernel_sum=tp["segment1/kernel"].grad.numpy().sum(axis=1),
                        g_outputs_kernel=tp["outputs/kernel"].grad.numpy())

    # 4. AP loss: seeded, fp64
    N, D, Bz = 7, 12, 9
    z = rng.standard_normal((Bz, D))
    z /= np.linalg.norm(z, axis=1, keepdims=True)
    yz = rng.integers(0, N, Bz)
    zt = torch.tensor(z, requires_grad=True)
    l = O.torch_ap_loss(torch.tensor(yz), zt, N, 1.5)
    l.backward()
    assert np.allclose(l.item(), O.ap_loss(yz, z, N, 1.5))
    np.savez_compressed(os.path.join(OUT, "ap_loss.npz"), z=z, y=yz, N=N, delta_weight=1.5,
                        per_sample=O.ap_loss_per_sample(yz, z, N, 1.5), loss=l.item(), grad=zt.grad.numpy())
    # 5. extended x-vector (xvector_extended.py:22-43): seeded tiny problem, fp64 oracle outputs
    ext = dict(frame_layers=O.XVECTOR_EXTENDED_FRAME_LAYERS, output_name="output")
    rng5 = np.random.default_rng(21)
    xe = rng5.standard_normal((3, 50, 24)).astype(np.float32)
    pe = {k: v.astype(np.float64) for k, v in O.xvector_init(24, 5, seed=12, bias_scale=0.05, **ext).items()}
    np.savez_compressed(os.path.join(OUT, "xvector_extended_small.npz"), x=xe,
                        logp=O.xvector_forward(pe, xe.astype(np.float64), **ext),
                        emb=O.xvector_forward(pe, xe.astype(np.float64), embedding=True, **ext))

    # 6. C_avg: the self-test inputs of lidbox/metrics.py:128-150 and a seeded 200 x 6 problem
    onehot = np.array([[1, 0, 0], [0, 1, 0], [0, 1, 0], [0, 1, 0], [1, 0, 0], [0, 0, 1], [0, 1, 0], [0, 0, 1]], np.float32)
    prob = np.array([[.1, .2, .9], [.9, .2, .0], [.1, .9, .0], [.2, .8, .5], [.6, .3, .1], [.1, .0, .7], [.1, .0, .7],
                     [.9, .1, .0]], np.float32)
    with np.errstate(divide="ignore"):
        pred = np.log(prob)
    thr = np.log(np.array([0.05, 0.4, 0.6, 0.95], np.float32))
    c = O.AverageDetectionCost(3, thr)
    c.update_state(onehot, pred)
    rng6 = np.random.default_rng(22)
    y6 = rng6.integers(0, 6, 200)
    s6 = rng6.standard_normal((200, 6)).astype(np.float32)
    s6[np.arange(200), y6] += 1.5
    thr6 = np.linspace(s6.min(), s6.max(), 25).astype(np.float32)
    c6 = O.SparseAverageDetectionCost(6, thr6, C_miss=1.0, C_fa=2.0, P_tar=0.3)
    c6.update_state(y6, s6)
    np.savez_compressed(os.path.join(OUT, "cavg.npz"), onehot=onehot, pred=pred, thr=thr, cavg=c.result_per_threshold(),
                        tp=c.tp, fn=c.fn, fp_pairs=c.fp_pairs, tn_pairs=c.tn_pairs,
                        y6=y6, s6=s6, thr6=thr6, cavg6=c6.result_per_threshold(), fp_pairs6=c6.fp_pairs)
    print("golden written to", OUT)


if __name__ == "__main__":
    main()
