import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from oracle import lidbox_oracle as O
from lidbox_b200.models import xvector as xv
sys.path.insert(0, "tests")
from test_xvector_gpu import _oracle_grads, _nw
for (B, T, n_out, loss, head) in ((32, 198, 4, "xent", "log_softmax"), (8, 61, 64, "ap", "l2_normalize")):
    rng = np.random.default_rng(4 if loss == "xent" else 5)
    x = rng.standard_normal((B, T, 40)).astype(np.float32)
    N = 50 if loss == "ap" else n_out
    y = rng.integers(0, N, B)
    params = O.xvector_init(40, n_out, seed=3 if loss == "xent" else 4, bias_scale=0.05)
    loss_ref, g_ref = _oracle_grads(params, x, y, loss=loss, N=N, emulate_bf16=True)
    m = xv.create((T, 40), n_out, precision="bf16", head=head)
    m.set_weights(params)
    kw = dict(ap_classes=N) if loss == "ap" else {}
    per = m.loss_and_grads(x, y, loss=loss, **kw).cpu().numpy()
    print("loss", per.mean(), loss_ref)
    grads = m.grads.cpu().numpy()
    for ly in m.layers:
        gw = grads[ly["w_off"]:ly["w_off"] + ly["K"] * ly["N"]].reshape(ly["K"], ly["N"])
        gb = grads[ly["b_off"]:ly["b_off"] + ly["N"]]
        rw = g_ref[ly["name"] + "/kernel"].reshape(ly["K"], ly["N"])
        rb = g_ref[ly["name"] + "/bias"]
        for got, ref, what in ((gw, rw, "kernel"), (gb, rb, "bias")):
            cos = (got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-30)
            print(loss, ly["name"], what, "cos %.6f nw %.4f norm ratio %.4f" % (cos, _nw(got, ref), np.linalg.norm(got)/np.linalg.norm(ref)))
