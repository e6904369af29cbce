"""world_size-2 gloo test of the multi-GPU host logic: sharded forecast + metric all-reduce == single process.
The per-rank forecast here is the CPU oracle (there is no GPU in this container); the sharding, the metric
accumulation and the collective are the product code under test (sttode_b200/parallel.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sttode_b200 import parallel, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _metric_sums(scene_ids):
    from oracle import sttode_oracle as O
    sd = synth.make_state_dict(1, 8, 12)
    sums = torch.zeros(3, dtype=torch.float64)
    for i in scene_ids:
        obs, gt = synth.eth_scene(i, synth.ETH_AGENT_CYCLE[i % 5])
        n = obs.shape[0]
        past, orig = O.eth_scene_normalise(obs)
        pred = O.inference(sd, past, 1, n, synth.latent(n, seed=100 + i), scene_orig=orig, fast=True)
        ade, fde = O.min_ade_fde(pred.permute(1, 0, 2, 3), gt.permute(0, 2, 1))
        sums += torch.tensor([ade * n, fde * n, n], dtype=torch.float64)
    return sums


def _worker(rank, world, port, n_scenes, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    lo, hi = parallel.shard_range(n_scenes, rank, world)
    sums = _metric_sums(range(lo, hi))
    ade, fde, n = parallel.reduce_metric_sums(sums)
    if rank == 0:
        torch.save((ade, fde, n), out)
    dist.destroy_process_group()


def test_sharded_metrics_equal_single_process(tmp_path):
    n_scenes, world = 6, 2
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(world, _free_port(), n_scenes, out), nprocs=world, join=True)
    ade, fde, n = torch.load(out)
    ref = _metric_sums(range(n_scenes))
    assert n == float(ref[2])
    assert abs(ade - float(ref[0] / ref[2])) < 1e-9 and abs(fde - float(ref[1] / ref[2])) < 1e-9


def test_shard_helpers():
    for n, w in [(10, 4), (3, 8), (16, 2), (1, 1)]:
        spans = [parallel.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    counts = [60, 10, 50, 20, 40, 30, 2, 3]
    parts = parallel.shard_scenes_by_cost(counts, 3)
    assert sorted(i for p in parts for i in p) == list(range(len(counts)))
    loads = [sum(counts[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 15
    with pytest.raises(ValueError):
        parallel.reduce_metric_sums(torch.zeros(3))


# ------------------------------------------------------------------ data-parallel training: flat gradient all-reduce
class _Tiny(torch.nn.Module):
    """Stand-in with the one structural feature that matters: a `pz_layer` that never receives a gradient."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = torch.nn.Linear(6, 5)
        self.pz_layer = torch.nn.Linear(5, 4)
        self.b = torch.nn.Linear(5, 3)
        self._x = None

    def forward(self):
        x, y = self._x
        loss = (self.b(torch.relu(self.a(x))) - y).pow(2).sum() / x.shape[0]
        return loss, float(loss.detach())


def _batch(i, n=8):
    g = torch.Generator().manual_seed(50 + i)
    return torch.randn(n, 6, generator=g), torch.randn(n, 3, generator=g)


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = _Tiny()
    red = parallel.FlatGradReducer(m)
    opt, sched = parallel.make_optimizer(m, lr=1e-2)
    losses = []
    for step in range(3):
        def stage():
            m._x = _batch(step * world + rank)
        losses.append(parallel.train_step(m, opt, red, stage)[1])
    if rank == 0:
        torch.save(({k: v.clone() for k, v in m.state_dict().items()}, red.flat.clone(), red.excluded), out)
    dist.destroy_process_group()


def test_dp_flat_gradient_allreduce_equals_big_batch(tmp_path):
    """world 2, per-rank batches: the all-reduced mean gradient equals the gradient of the mean objective over both batches,
    so 3 DP Adam steps land on the same weights as a single process that sees the concatenated batches."""
    world = 2
    out = str(tmp_path / "dp.pt")
    mp.spawn(_dp_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    sd_dp, flat, excluded = torch.load(out)
    assert excluded == ["pz_layer.weight", "pz_layer.bias"]
    m = _Tiny()
    red = parallel.FlatGradReducer(m)
    assert red.flat.numel() == sum(p.numel() for n, p in m.named_parameters() if not n.startswith("pz_layer"))
    opt, _ = parallel.make_optimizer(m, lr=1e-2)
    for step in range(3):
        xs, ys = zip(*[_batch(step * world + r) for r in range(world)])
        m._x = (torch.cat(xs), torch.cat(ys))
        red.zero()
        m.forward()[0].backward()
        red.reduce()                       # world 1: no collective
        opt.step()
    for k, v in m.state_dict().items():
        assert torch.allclose(v, sd_dp[k], rtol=1e-5, atol=1e-6), k
    assert torch.allclose(red.flat, flat, rtol=1e-4, atol=1e-6)
    assert m.pz_layer.weight.grad is None


def test_flat_reducer_guards():
    m = _Tiny()
    red = parallel.FlatGradReducer(m)
    m._x = _batch(0)
    m.forward()[0].backward()
    m.zero_grad(set_to_none=True)
    with pytest.raises(RuntimeError):
        red.reduce()
