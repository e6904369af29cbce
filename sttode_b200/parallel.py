"""Multi-GPU host logic: one process per GPU, weights replicated, scenes (ETH/UCY/SDD) or whole
reference batches (NBA — the scenes of a batch attend to each other, so a batch never splits) sharded across ranks.
There is NO data-path collective: each rank forecasts its shard independently; the only exchange is one
all_reduce(SUM) of the three metric accumulators [sum minADE, sum minFDE, agents] in float64, mirroring
`AverageMeter.update(val, n)` (utils/utils.py:29-34) / the `* batch ... / all_num` sums of test.py:532-552.
Works with any torch.distributed backend (nccl on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced [start, stop) of `n_items` (scenes or NBA batches) for `rank`."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_scenes_by_cost(agent_counts, world):
    """Cost-balanced assignment for ETH/SDD scenes of very different sizes: greedy longest-processing-time on the
    agent count (decoder work is linear in agents).  Returns a list of index lists, one per rank."""
    order = sorted(range(len(agent_counts)), key=lambda i: -agent_counts[i])
    load, out = [0] * world, [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: load[k])
        out[r].append(i)
        load[r] += agent_counts[i]
    return [sorted(x) for x in out]


def reduce_metric_sums(sums):
    """In-place all_reduce(SUM) of a float64 tensor [sum minADE, sum minFDE, agents]; returns (ADE, FDE, agents)."""
    if sums.dtype != torch.float64 or sums.numel() != 3:
        raise ValueError("sums must be a float64 tensor of 3 elements")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    n = float(sums[2])
    return float(sums[0]) / max(n, 1.0), float(sums[1]) / max(n, 1.0), n


# ---------------------------------------------------------------------------------------------------------------------
# Data-parallel training (BASELINE.json configs[4]): every rank draws its own batch (train.py:24, B = 32), so
# the global batch is 32 x world; gradients are summed over ranks in ONE all_reduce of a flat fp32 buffer and divided by
# world (the mean of the per-rank objectives).  `pz_layer` never receives a gradient (learn_prior = False, train.py:28) and is
# left out of the buffer so the collective carries no dead bytes.
# ---------------------------------------------------------------------------------------------------------------------
class FlatGradReducer:
    """Owns one contiguous fp32 gradient buffer; every participating parameter's `.grad` is a view into it, so autograd
    accumulates straight into the buffer and the data-parallel exchange is a single all_reduce(SUM) + one scale."""

    def __init__(self, module, exclude=("pz_layer",), group=None):
        self.group = group
        self.params = [(n, p) for n, p in module.named_parameters() if p.requires_grad and not any(n.startswith(e) or ("." + e) in n for e in exclude)]
        self.excluded = [n for n, p in module.named_parameters() if p.requires_grad and (n, p) not in [(a, b) for a, b in self.params]]
        if not self.params:
            raise ValueError("no parameters to reduce")
        p0 = self.params[0][1]
        if any(p.dtype != torch.float32 or p.device != p0.device for _, p in self.params):
            raise ValueError("FlatGradReducer: all parameters must be float32 on one device")
        total = sum(p.numel() for _, p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=p0.device)
        off = 0
        for _, p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        # the library's backward kernels accumulate (+=) straight into these views: no zero-fill + add per parameter per autograd node
        from . import _lib
        _lib.DIRECT_GRAD_SINKS = True

    @property
    def nbytes(self):
        return self.flat.numel() * 4

    def zero(self):
        """Replaces optimizer.zero_grad(): keeps the views alive (set_to_none would detach them from the buffer)."""
        self.flat.zero_()

    def world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def reduce(self):
        """Sum over ranks, then mean.  Call after backward(), before optimizer.step()."""
        for n, p in self.params:
            if p.grad is None or p.grad.data_ptr() < self.flat.data_ptr() or p.grad.data_ptr() >= self.flat.data_ptr() + self.nbytes:
                raise RuntimeError(f"{n}.grad no longer aliases the flat buffer (use reducer.zero(), not zero_grad(set_to_none=True))")
        w = self.world()
        if w > 1:
            if dist.get_backend(self.group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)        # 1/world applied inside the collective: one pass
            else:                                                                        # gloo (CPU tests) has no AVG
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
                self.flat.mul_(1.0 / w)
        return self.flat


class _DevMem:
    """Zero-copy torch view of library-owned device memory (torch.as_tensor reads __cuda_array_interface__)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class FusedDPAdam(FlatGradReducer):
    """Data-parallel Adam whose gradient exchange runs over NVLink / NVSwitch peer memory inside the update kernel (csrc/comm.cu,
    sttode_dp_adam_step): barrier -> every rank sums all ranks' flat gradient buffers in rank order, scales by 1/world and applies
    torch.optim.Adam's update to its weight replica -> barrier.  No NCCL call and no host synchronisation, so the WHOLE step (stage,
    forward, backward, exchange, update) is one CUDA graph at any world size.  Reference: train.py:122 (Adam, lr 1e-4), :64-66.

    Parameters become views into one flat fp32 buffer (so do their `.grad`, inherited from FlatGradReducer; `pz_layer` never receives a
    gradient and stays outside, as torch's Adam would skip it).  The gradient buffer lives in a cudaMalloc'ed arena exported through CUDA
    IPC; the 64-byte handles travel through torch.distributed (any backend) once, at construction.  `step()` does not bump the parameters'
    version counters: call `model.invalidate_packed()` before the next eval forecast (train_step / GraphedTrainStep do)."""

    def __init__(self, module, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, exclude=("pz_layer",), group=None):
        import ctypes as C
        from . import _lib
        self._lib, self.module, self.group = _lib, module, group
        lib = _lib.lib()
        params = [(n, p) for n, p in module.named_parameters() if p.requires_grad and not any(n.startswith(e) or ("." + e) in n for e in exclude)]
        self.params = params
        self.excluded = [n for n, p in module.named_parameters() if p.requires_grad and all(p is not q for _, q in params)]
        dev = params[0][1].device
        if any(p.dtype != torch.float32 or p.device != dev for _, p in params) or dev.type != "cuda":
            raise ValueError("FusedDPAdam: all parameters must be float32 on one CUDA device")
        total = sum(p.numel() for _, p in params)
        self.n = (total + 3) // 4 * 4
        self.world_size = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world_size > 1 else 0
        arena, handle = C.c_void_p(), (C.c_ubyte * 64)()
        _lib.check(lib.sttode_peer_alloc(self.n, C.byref(arena), handle))
        self.arena = arena.value
        self.flat = torch.as_tensor(_DevMem(self.arena, self.n), device=dev)              # the flat gradient buffer, in the exported arena
        self.flat_param = torch.zeros(self.n, device=dev)
        self.exp_avg, self.exp_avg_sq = torch.zeros(self.n, device=dev), torch.zeros(self.n, device=dev)
        off = 0
        with torch.no_grad():
            for _, p in params:
                k = p.numel()
                self.flat_param[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_param[off:off + k].view_as(p)
                p.grad = self.flat[off:off + k].view_as(p)
                off += k
        self.lr = torch.full((1,), float(lr), device=dev)
        self.betas, self.eps = betas, eps
        arenas = [self.arena] * self.world_size
        self._opened = []
        if self.world_size > 1:
            handles = [None] * self.world_size
            dist.all_gather_object(handles, bytes(handle), group=group)
            for r, h in enumerate(handles):
                if r == self.rank:
                    continue
                ptr = C.c_void_p()
                _lib.check(lib.sttode_peer_open((C.c_ubyte * 64).from_buffer_copy(h), C.byref(ptr)))
                arenas[r] = ptr.value
                self._opened.append(ptr.value)
        self.grads_of = torch.tensor(arenas, dtype=torch.int64, device=dev)
        self.flags_of = torch.tensor([a + self.n * 4 for a in arenas], dtype=torch.int64, device=dev)
        torch.cuda.synchronize(dev)
        if self.world_size > 1:
            dist.barrier(group=group)               # every arena is mapped everywhere before the first step touches a peer
        _lib.DIRECT_GRAD_SINKS = True

    def world(self):
        return self.world_size

    def reduce(self):
        raise RuntimeError("FusedDPAdam exchanges the gradients inside step(); there is no separate reduce()")

    def zero_grad(self, set_to_none=False):
        self.zero()

    def set_lr(self, lr):
        """StepLR (train.py:123) changes the rate between epochs: a device scalar, so captured steps see it."""
        self.lr.fill_(float(lr))

    def step(self, grad_avg=None):
        """Exchange + update on the current stream.  grad_avg (optional flat tensor) receives the averaged gradient (tests / diagnostics)."""
        L = self._lib
        L.check(L.lib().sttode_dp_adam_step(self.grads_of.data_ptr(), self.flags_of.data_ptr(), self.world_size, self.rank, L.dptr(self.flat_param),
                                            L.dptr(self.exp_avg), L.dptr(self.exp_avg_sq), L.dptr(grad_avg), self.n, L.dptr(self.lr),
                                            self.betas[0], self.betas[1], self.eps, L.stream_ptr()))

    def status(self):
        """(error word, steps taken); synchronises.  error != 0: a peer never arrived at a barrier (the step's result is invalid)."""
        import ctypes as C
        err, step = C.c_uint(), C.c_uint()
        torch.cuda.synchronize()
        self._lib.check(self._lib.lib().sttode_peer_status(self.arena, self.n, C.byref(err), C.byref(step)))
        return err.value, step.value

    def close(self):
        """Unmap the peers' arenas and free this rank's (call on every rank after a collective barrier; the parameters keep their flat storage,
        their .grad views die with the arena)."""
        lib = self._lib.lib()
        torch.cuda.synchronize()
        for _, p in self.params:
            p.grad = None
        self.flat = None
        for a in self._opened:
            lib.sttode_peer_close(a)
        self._opened = []
        if self.arena:
            lib.sttode_peer_free(self.arena)
            self.arena = None


def describe_reduce(reducer, graphed):
    """One line for bench.py's config: how the gradients of this run are exchanged."""
    w = reducer.world()
    if isinstance(reducer, FusedDPAdam):
        return (f"gradient all-reduce fused with Adam over NVLink peer memory (sttode_dp_adam_step: barrier, every rank sums the {w} flat {reducer.nbytes} B "
                f"buffers in rank order and updates its replica, barrier); no NCCL call; whole step = {'one CUDA-graph replay' if graphed else 'eager launches'}")
    if w == 1:
        return "single GPU: no exchange; whole step (stage .. Adam) is one CUDA-graph replay" if graphed else "single GPU: no exchange"
    how = "between a forward+backward graph and an Adam graph" if graphed else "eager"
    return f"one NCCL all-reduce(SUM) of the flat {reducer.nbytes} B gradient buffer per step ({how}), pz_layer excluded, 1/world folded into the same pass"


def make_optimizer(module, lr=1e-4, decay_step=10, decay_gamma=0.5, capturable=False):
    """train.py:122-123: Adam(lr) + StepLR(decay_step, decay_gamma).  capturable=True for `GraphedTrainStep`."""
    opt = torch.optim.Adam(module.parameters(), lr=lr, capturable=capturable)
    return opt, torch.optim.lr_scheduler.StepLR(opt, step_size=decay_step, gamma=decay_gamma)


def train_step(model, optimizer, reducer, stage):
    """One data-parallel step of train.py:61-66: stage() -> forward -> zero -> backward -> (all-reduce) -> Adam.
    `stage` puts this rank's batch into the model (set_data_nba / set_data).  Returns the reference's loss tuple."""
    stage()
    out = model.forward()
    reducer.zero()
    out[0].backward()
    if isinstance(reducer, FusedDPAdam):            # exchange + update in one kernel sequence over peer memory (optimizer is the reducer itself)
        reducer.step()
        model.invalidate_packed()
    else:
        reducer.reduce()
        optimizer.step()
    return out


class GraphedTrainStep:
    """One whole data-parallel training step — stage -> objective -> backward -> gradient all-reduce -> Adam — replayed from CUDA graphs
    for a fixed NBA batch shape.  The step is ~1000 small launches (502 of them libsttode_sm100 kernels); replaying a graph removes the
    Python / launch overhead that otherwise dominates it.  The optimizer must be built with `capturable=True`.

        step = GraphedTrainStep(model, optimizer, reducer, batch)      # batch: {'past_traj': [B,N,T_p,2], 'future_traj': [B,N,T_f,2]} on the GPU
        losses = step(next_batch)                                        # tensor [total, pred, recover, kl, diverse], no host sync

    World size 1: one graph holds the whole step.  World size > 1: the NCCL all-reduce stays OUTSIDE the graphs (graph 1 = stage ..
    backward, eager all-reduce of the flat buffer, graph 2 = Adam): a collective captured into a graph ties every rank's replay to
    the communicator's internal stream state, and a rank that falls out of step deadlocks the job instead of raising.
    Construction runs `warmup` (>= 1) real training steps on `batch` eagerly (allocations, optimizer state) before capturing.
    """

    def __init__(self, model, optimizer, reducer, batch, warmup=3):
        if warmup < 1:
            # the optimizer creates (and zero-fills) its state on its first step; inside the capture that fill would be replayed every step
            raise ValueError("GraphedTrainStep needs at least one eager warm-up step (it is a real training step on `batch`)")
        self.model, self.optimizer, self.reducer = model, optimizer, reducer
        dev = batch["past_traj"].device
        self.static = {k: batch[k].detach().clone().contiguous() for k in ("past_traj", "future_traj")}
        self.losses = torch.zeros(5, device=dev)
        self.fused = isinstance(reducer, FusedDPAdam)
        self.split = reducer.world() > 1 and not self.fused       # NCCL stays outside the graphs; the peer-memory exchange is captured with the rest
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._forward_backward()
                if not self.fused:
                    self.reducer.reduce()
                self._update()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import _lib
        l0 = _lib.launch_count()
        # Capture on the SAME stream the warm-up ran on: the parameters' AccumulateGrad nodes were created there and outlive the warm-up
        # graph (the model's cached stage tensors keep it alive); autograd closes backward() by making the caller's stream wait on every
        # leaf stream, which is only legal under capture when that stream is the capturing one (with direct gradient sinks nothing else
        # would pull a foreign leaf stream into the capture first).
        self.graph = torch.cuda.CUDAGraph()
        if not self.split:
            with torch.cuda.graph(self.graph, stream=side):
                self._forward_backward()
                self._update()
            self.graph_update = None
        else:
            with torch.cuda.graph(self.graph, stream=side):
                self._forward_backward()
            self.graph_update = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_update, pool=self.graph.pool(), stream=side):
                self._update()
        self.library_launches = _lib.launch_count() - l0       # libsttode_sm100 kernels recorded into the graph(s) = per replayed step

    def _forward_backward(self):
        # the tensor-core weight images (decoder, ODE blocks) must be rebuilt from the CURRENT weights inside every replay: drop the cache keys so
        # the first use in the step issues (and the capture records) each pack; later uses in the same step reuse it
        self.model.invalidate_packed()
        self.model.set_data_nba(self.static)
        total, parts = self.model.objective()
        self.reducer.zero()
        total.backward()
        self.losses.copy_(torch.cat((total.detach().reshape(1), parts)))

    def _update(self):
        if self.fused:
            self.reducer.step()
        else:
            self.optimizer.step()

    def __call__(self, batch):
        for k, v in self.static.items():
            v.copy_(batch[k], non_blocking=True)
        self.graph.replay()
        if self.split:
            self.reducer.reduce()             # eager NCCL all-reduce of the flat gradient buffer on the current stream
            self.graph_update.replay()
        # a replayed optimizer step changes the weights without bumping their version counters: the next eager forecast must repack
        self.model.invalidate_packed()
        return self.losses

    @property
    def graph_replays_per_step(self):
        return 2 if self.split else 1
