"""Where does the end-to-end step lose its ~30 us against the device-resident step?  Same pipeline as bench.py's e2e
loop (16-bit PCM in pinned host memory), with the H2D copy and / or the loss read-back switched off."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
args = argparse.Namespace(batch=0, seconds=0)
wl = bench.XVectorTrainWorkload(args, 0, 1)
wl.setup(dev)


def run(h2d=True, d2h=True, steps=200):
    pipe = wl._build_pipe(wl.pcm_host)
    wl._setup_e2e(pipe)
    e = pipe["e2e"]

    def step():
        k = pipe["i"] % 2
        cur = torch.cuda.current_stream(dev)
        if h2d:
            cur.wait_event(e["h2d_done"][1 - k])
        prev_done = e["step_done"]
        losses = wl.step(pipe)
        e["step_done"] = torch.cuda.Event()
        e["step_done"].record(cur)
        if h2d:
            with torch.cuda.stream(e["copy"]):
                e["copy"].wait_event(prev_done)
                pipe["xs"][k].copy_(pipe["x_host"], non_blocking=True)
                e["h2d_done"][k].record(e["copy"])
        slot = (pipe["i"] - 1) % bench.E2E_LOSS_RING
        if d2h:
            e["loss_host"][slot].copy_(losses, non_blocking=True)
        if slot == bench.E2E_LOSS_RING - 1:
            cur.synchronize()
    for _ in range(8):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


res = {"h2d+d2h": run(), "h2d only": run(d2h=False), "d2h only": run(h2d=False), "neither (sync every 8 steps)": run(False, False)}
print(json.dumps(res, indent=1))
