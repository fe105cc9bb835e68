import sys, numpy as np, torch
sys.path.insert(0, ".")
from oracle import lidbox_oracle as O
from lidbox_b200.models import xvector as xv

def run(layers, B=6, T=61, n_out=5, tag=""):
    frames = [xv.frame_layer(f, k, s, name=n) for n, f, k, s in layers]
    ext = dict(frame_layers=layers, output_name="outputs")
    rng = np.random.default_rng(7)
    x = rng.standard_normal((B, T, 40)).astype(np.float32)
    y = rng.integers(0, n_out, B)
    params = O.xvector_init(40, n_out, seed=6, bias_scale=0.05, **ext)
    m = xv.XVector((T, 40), n_out, frames=frames, precision="bf16")
    m.set_weights(params)
    per = m.loss_and_grads(x, y).cpu().numpy()
    grads = m.grads.cpu().numpy()
    for emu in (True, False):
        tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in params.items()}
        lp = O.torch_xvector_forward(tp, torch.tensor(x, dtype=torch.float64), emulate_bf16=emu, **ext)
        l = -lp[torch.arange(len(y)), torch.tensor(y)].mean(); l.backward()
        out = []
        for ly in m.layers:
            gw = grads[ly["w_off"]:ly["w_off"] + ly["K"] * ly["ldw"]].reshape(ly["K"], ly["ldw"])[:, :ly["N"]]
            rw = tp[ly["name"] + "/kernel"].grad.numpy().reshape(ly["K"], ly["N"])
            gb = grads[ly["b_off"]:ly["b_off"] + ly["N"]]; rb = tp[ly["name"] + "/bias"].grad.numpy()
            c = lambda a, b: (a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30)
            out.append("%s %.5f/%.5f" % (ly["name"], c(gw, rw), c(gb, rb)))
        print(tag, "emu" if emu else "f64", "loss %.5f vs %.5f |" % (per.mean(), float(l)), " ".join(out))

run(O.XVECTOR_EXTENDED_FRAME_LAYERS, tag="ext")
run((("frame1", 512, 5, 1), ("frame2", 512, 3, 4), ("frame3", 1500, 1, 1)), tag="s4")
run((("frame1", 512, 5, 1), ("frame2", 512, 1, 1), ("frame3", 512, 1, 1), ("frame4", 512, 1, 1), ("frame5", 512, 1, 1), ("frame6", 512, 1, 1), ("frame7", 512, 1, 1),("frame8", 512, 1, 1),("frame9", 512, 1, 1), ("frame10", 1500, 1, 1)), tag="deep11")
run(O.FRAME_LAYERS, tag="base")
