"""Small x-vector invocation used by __graft_entry__.smoke(): fp32 embedding + one bf16 training step, each checked
against the CPU oracle (the oracle is imported here as the checker only)."""
import numpy as np
import torch


def run():
    from oracle import lidbox_oracle as O
    from .models import xvector
    rng = np.random.default_rng(0)
    B, T, F, n_out = 4, 98, 40, 4
    x = rng.standard_normal((B, T, F)).astype(np.float32)
    y = np.arange(B) % n_out
    params = O.xvector_init(F, n_out, seed=0, bias_scale=0.02)
    p64 = {k: v.astype(np.float64) for k, v in params.items()}
    m = xvector.create((T, F), n_out)
    m.set_weights(params)
    logp = m(x).cpu().numpy()
    np.testing.assert_allclose(logp, O.xvector_forward(p64, x.astype(np.float64)), rtol=1e-4, atol=1e-4)
    mt = xvector.create((T, F), n_out, precision="bf16")
    mt.set_weights(params)
    mt.configure_optimizer()
    loss = float(mt.train_step(x, y).mean())
    ref = O.sparse_xent_on_logprobs(y, O.xvector_forward(p64, x.astype(np.float64)))
    assert abs(loss - ref) < 3e-2 * max(1.0, abs(ref)), (loss, ref)
    assert torch.isfinite(mt.params).all()
    print("smoke x-vector OK: fp32 log-probs match the oracle to 1e-4; bf16 train step loss %.4f (oracle %.4f)" % (loss, ref))
