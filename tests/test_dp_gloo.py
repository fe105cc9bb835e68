"""CPU test (-m "not gpu") of the N>1 host logic with world_size-2 gloo: the data-parallel contract used by
XVector.train_step and bench.py — every rank scales its loss gradient by 1/(global batch), the flat fp32 gradient is
sum-all-reduced once, and the result equals the single-process gradient of the whole batch.  The gradient itself is
produced by the CPU oracle here (no GPU in this container); the GPU twin of this test runs under gpurun --gpus 2."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import lidbox_oracle as O


def _flat_grad(params, x, y, global_batch):
    tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in params.items()}
    lp = O.torch_xvector_forward(tp, torch.tensor(x, dtype=torch.float64))
    loss = -lp[torch.arange(len(y)), torch.tensor(y)].sum() / global_batch      # per-rank share of the global mean
    loss.backward()
    return torch.cat([tp[k].grad.reshape(-1) for k in sorted(tp)]), float(loss.detach())


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((4, 21, 8)).astype(np.float32)
    y = np.array([0, 1, 2, 1])
    params = O.xvector_init(8, 3, seed=5, bias_scale=0.05)
    per = len(y) // world
    sl = slice(rank * per, (rank + 1) * per)                                     # rank r owns utterances [r*B/R, (r+1)*B/R)
    g, loss = _flat_grad(params, x[sl], y[sl], global_batch=len(y))
    dist.all_reduce(g)                                                           # the ONE exchange step of the path
    loss_t = torch.tensor([loss], dtype=torch.float64)
    dist.all_reduce(loss_t)
    if rank == 0:
        torch.save({"grad": g, "loss": loss_t}, os.path.join(out_dir, "dp.pt"))
    dist.destroy_process_group()


def test_two_rank_gradient_equals_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(str(tmp_path), "dp.pt"))
    rng = np.random.default_rng(0)
    x = rng.standard_normal((4, 21, 8)).astype(np.float32)
    y = np.array([0, 1, 2, 1])
    params = O.xvector_init(8, 3, seed=5, bias_scale=0.05)
    g_ref, loss_ref = _flat_grad(params, x, y, global_batch=len(y))
    assert torch.allclose(got["grad"], g_ref, atol=1e-12)
    assert abs(float(got["loss"]) - loss_ref) < 1e-12
