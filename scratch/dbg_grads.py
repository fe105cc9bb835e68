import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from oracle import lidbox_oracle as O
from lidbox_b200.models import xvector as xv
sys.path.insert(0, "tests")
from test_xvector_gpu import _oracle_grads, _nw
for (B, T, n_out) in ((6, 37, 5), (32, 198, 4)):
    rng = np.random.default_rng(4)
    x = rng.standard_normal((B, T, 40)).astype(np.float32)
    y = rng.integers(0, n_out, B)
    params = O.xvector_init(40, n_out, seed=3, bias_scale=0.05)
    loss_ref, g_ref, per_ref = _oracle_grads(params, x, y)
    m = xv.create((T, 40), n_out, precision="bf16")
    m.set_weights(params)
    per = m.loss_and_grads(x, y).cpu().numpy()
    print("loss", per.mean(), loss_ref)
    grads = m.grads.cpu().numpy()
    for ly in m.layers:
        gw = grads[ly["w_off"]:ly["w_off"] + ly["K"] * ly["N"]].reshape(ly["K"], ly["N"])
        gb = grads[ly["b_off"]:ly["b_off"] + ly["N"]]
        rw = g_ref[ly["name"] + "/kernel"].reshape(ly["K"], ly["N"])
        rb = g_ref[ly["name"] + "/bias"]
        for got, ref, what in ((gw, rw, "kernel"), (gb, rb, "bias")):
            cos = (got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-30)
            print(B, T, ly["name"], what, "cos %.6f nw %.4f norm ratio %.4f" % (cos, _nw(got, ref), np.linalg.norm(got)/np.linalg.norm(ref)))
        if ly["name"] in ("frame1", "frame2"):
            k = ly["k"]
            gk = gw.reshape(k, -1, ly["N"]); rk = rw.reshape(k, -1, ly["N"])
            for j in range(k):
                cos = (gk[j] * rk[j]).sum() / (np.linalg.norm(gk[j]) * np.linalg.norm(rk[j]) + 1e-30)
                print("   tap", j, "cos %.6f" % cos)
