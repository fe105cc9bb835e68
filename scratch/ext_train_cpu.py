import sys, numpy as np, torch
sys.path.insert(0, ".")
from oracle import lidbox_oracle as O
ext = dict(frame_layers=O.XVECTOR_EXTENDED_FRAME_LAYERS, output_name="output")
rng = np.random.default_rng(8)
B, T = 32, 98
y = np.arange(B) % 4
x = (rng.standard_normal((B, T, 40)) + y[:, None, None] * 1.5).astype(np.float32)
for lr in (1e-3, 2e-4):
    params = O.xvector_init(40, 4, seed=1, **ext)
    tp = {k: torch.tensor(v, dtype=torch.float32, requires_grad=True) for k, v in params.items()}
    opt = torch.optim.Adam(tp.values(), lr=lr, eps=1e-7)
    xt, yt = torch.tensor(x), torch.tensor(y)
    ls = []
    for i in range(41):
        opt.zero_grad()
        lp = O.torch_xvector_forward(tp, xt, **ext)
        l = -lp[torch.arange(B), yt].mean(); l.backward(); opt.step(); ls.append(float(l.detach()))
    print(lr, " ".join("%.3f" % v for v in ls[::4]))
