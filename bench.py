#!/usr/bin/env python
"""bench.py — forecast scenes/sec (K = 20 samples) of the STTODE forward path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config eth|nba] [--train]

Headline workload (BASELINE.json configs[1], `--config eth`, the default): ETH-UCY zara/univ-shaped batched inference — scenes of
N in {10,20,30,40,50,60} agents, obs 8 / pred 12 frames, K = 20 samples, synthetic trajectories (sttode_b200/synth.py) and seeded
synthetic weights.  `--config nba` (configs[2]): NBA shape — 11 agents, obs 10 / pred 20, whole reference batches of 128 scenes
(test.py:618; the scenes of a batch attend to each other, MHGSA L = S = 128), sharded over ranks by whole batches.
One step = one pass of the whole `STTODENet.inference` hot path (scene normalisation, past encoder with the MHGSA / tensor-ODE
block, in-kernel latent draw, K-sample decoder, min-ADE/FDE reduction) over this rank's shard.
N > 1: shards are independent (weak scaling, no data-path collective); one NCCL all-reduce of the three metric sums closes the
timed region.

Output: ONE JSON line on rank 0 (contract in the task statement): value = whole-job scenes/s with inputs resident in HBM; e2e = the
same through sttode_forecast_host (host buffers, H2D + D2H of every sample inside the timed region); e2e_metrics_only = the reference's
evaluation loop body through sttode_evaluate_host (host trajectories + ground truth in, min-over-K ADE/FDE reduced on the device, 24 bytes
out: secondary, shows the end-to-end scaling without the host-memory wall of the prediction copy); roofline for the dominant kernel from CUDA
events; cpu_baseline = the oracle port of the reference timed on the host cores (N = 1 only).  The default run also measures the
other inference config and the data-parallel training step (configs[4]) briefly and reports them under "configs".

Multi-rank discipline (round-1 failure): every collective is issued by ALL ranks in the same order; after the last one every rank
destroys its process group and ranks != 0 exit at once.  Host-only work (CPU baseline, parity, single-process extras) runs on rank 0
AFTER the group is gone and only when world == 1, so no peer ever waits inside NCCL for a CPU leg.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 20
CYCLE = (10, 20, 30, 40, 50, 60)
CONFIGS = {
    "eth": {"tp": 8, "tf": 12, "weight_seed": 1, "dataset": "eth",
            "workload": "ETH-UCY zara/univ-shape batched inference (BASELINE.json configs[1]): N in {10..60} agents/scene, obs 8 / pred 12, K=20"},
    "nba": {"tp": 10, "tf": 20, "weight_seed": 2, "dataset": "nba", "B": 128, "agents": 11,
            "workload": "NBA-shape inference (BASELINE.json configs[2]): 11 agents, obs 10 / pred 20, K=20, reference batches of 128 scenes "
                        "(MHGSA L = S = 128 across the scenes of a batch), sharded by whole batches"},
}
# dram read+write of k_dec_pair_tc per (agent, sample) pair from `ncu --set full` (one launch of 268 800 / 225 280 pairs)
DRAM_BYTES_PER_PAIR = {"eth": (85.84e6 + 14.39e6) / 268800, "nba": (72.69e6 + 12.36e6) / 225280}      # profiles/r02_ncu_full_{eth,nba}.md (eth: final persistent kernel)


# ---------------------------------------------------------------------------------- FLOP accounting
def flops_mlp(k1, n3):
    return 2 * k1 * 512 + 2 * 512 * 256 + 2 * 256 * n3


def flops_gru(T):
    return T * (2 * 6 * 32 + 2 * (32 + 96) * 288)


def issued_flops_dec_pair(tp, tf):
    """FLOPs k_dec_pair issues per (agent, sample) pair after de-duplication (csrc/decoder.cu header)."""
    return flops_mlp(32, 2 * tp) + flops_gru(tp) + flops_mlp(32, 2 * tf) + flops_mlp(128, 2 * tf)


def issued_flops_per_row(tp, S=1):
    enc = 25088 * tp + 8576                                       # front end (SURVEY §8d enc_pre)
    if S == 1:
        ode = 2 * 64 * 64 * 4 + 2 * 2 * 64 * 1024                 # L = 1: v-proj, out-proj, 2 gates, FFN
    else:
        ode = 311296 + 256 * S                                    # SURVEY §8d ode_block(S)
    dec_row = flops_gru(tp) + 2 * (2 * 224 * 512) + 2 * (2 * 128 * 512)
    return enc + ode + dec_row


def reference_flops_per_row(tp, tf, S=1):
    """Algorithmic matmul FLOPs of the reference's own op list per agent row (SURVEY §8d F_inference / R)."""
    dec_blk = 74112 * tp + 1048576 + 1024 * (tp + tf)
    return (25088 * tp + 8576) + (311296 + 256 * S) + K * 2 * dec_blk


# ---------------------------------------------------------------------------------- helpers
def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def host_threads():
    """Threads the CPU legs may use: the cores this process is allowed on (torchrun exports OMP_NUM_THREADS=1; we set it explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def make_eth_workload(n_scenes, start, tp, tf):
    from sttode_b200 import synth
    scenes = synth.eth_scenes(n_scenes, cycle=CYCLE, past_length=tp, future_length=tf, start=start)
    counts = [o.shape[0] for o, _ in scenes]
    obs = torch.cat([o for o, _ in scenes]).contiguous()
    gt = torch.cat([g for _, g in scenes]).contiguous()
    sp = torch.tensor(np.concatenate([[0], np.cumsum(counts)]), dtype=torch.int32)
    return scenes, obs, gt, sp


def make_nba_workload(n_batches, start, cfg):
    from sttode_b200 import synth
    bs = [synth.nba_batch(start + i, cfg["B"], cfg["agents"], cfg["tp"], cfg["tf"]) for i in range(n_batches)]
    past = torch.stack([b["past_traj"] for b in bs]).contiguous()          # [nb, B, N, T_p, 2]
    fut = torch.stack([b["future_traj"] for b in bs]).contiguous()
    return bs, past, fut


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)

            def pump():
                for ln in self.proc.stdout:
                    self.lines.append((time.perf_counter(), ln))
            self.thr = threading.Thread(target=pump, daemon=True)
            self.thr.start()
            t0 = time.perf_counter()
            while not self.lines and time.perf_counter() - t0 < 5.0:      # nvidia-smi takes a moment to come up
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def mark(self):
        """Start of the timed region: only samples taken after this count."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thr.join(timeout=2)
        return self.summarise(self.lines, getattr(self, "t_mark", 0.0))

    @staticmethod
    def summarise(lines, t_mark):
        """Median SM clock and throttle reasons of the samples taken inside the timed region.  A region shorter than the 20 ms sampling period may hold no
        sample at all: then the nearest ones on either side are used (the one before was taken during the warm-up steps, i.e. under the same load) and
        `window` says so."""
        parsed = []
        for ts, ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                parsed.append((ts, float(f[1]), float(f[2]), f[5:9]))
            except ValueError:
                continue
        inside = [p for p in parsed if p[0] >= t_mark]
        window = "timed region"
        if not inside and parsed:
            inside = parsed[-2:]
            window = "nearest samples before the timed region (region shorter than the 20 ms sampling period)"
        sm, smax, reasons = [p[1] for p in inside], (inside[-1][2] if inside else None), set()
        for p in inside:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ---------------------------------------------------------------------------------- CPU port of the reference (oracle/), timed legs only
def cpu_port_eth(sd, scenes, threads, budget_s=None, device=None):
    """The reference's per-scene call sequence (set_data + inference, test.py:171-187) on the oracle port.  Stops after `budget_s`
    seconds of wall time (always finishes >= 2 scenes).  -> (scenes done, seconds)."""
    from oracle import sttode_oracle as O
    torch.set_num_threads(threads)
    done, t0 = 0, time.perf_counter()
    with torch.no_grad():
        for obs, _ in scenes:
            n = obs.shape[0]
            if device is not None:
                obs = obs.to(device)                                        # test.py:175 `.cuda()` per scene
            past, orig = O.eth_scene_normalise(obs)
            z = torch.randn(n * K, 32, device=obs.device)
            out = O.inference(sd, past, 1, n, z, sample_k=K, scene_orig=orig, fast=True)
            if device is not None:
                out = out.permute(1, 0, 2, 3).cpu()                         # test.py:186-194 D2H of the predictions
            done += 1
            if budget_s is not None and done >= 2 and time.perf_counter() - t0 > budget_s:
                break
    return done, time.perf_counter() - t0


def cpu_port_nba(sd, batches, cfg, threads, budget_s=None, device=None):
    """test.py:495-526: one reference batch of B = 128 scenes per `inference` call.  -> (scenes done, seconds)."""
    from oracle import sttode_oracle as O
    torch.set_num_threads(threads)
    B, N, tp = cfg["B"], cfg["agents"], cfg["tp"]
    done, t0 = 0, time.perf_counter()
    with torch.no_grad():
        for b in batches:
            past = b["past_traj"].view(B * N, tp, 2)
            if device is not None:
                past = past.to(device)
            z = torch.randn(B * N * K, 32, device=past.device)
            out = O.inference(sd, past, B, N, z, sample_k=K, fast=True)
            if device is not None:
                out = out.cpu()
            done += B
            if budget_s is not None and time.perf_counter() - t0 > budget_s:
                break
    return done, time.perf_counter() - t0


def cpu_port(name, sd, items, threads, budget_s=None, device=None):
    if name == "eth":
        return cpu_port_eth(sd, items, threads, budget_s, device)
    return cpu_port_nba(sd, items, CONFIGS["nba"], threads, budget_s, device)


# ---------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """`--impl reference`: the reference's own CPU path for the same config.  The reference is a Python package with no installer
    and cannot travel to the GPU box, so this times the oracle port (same ATen calls in the same order), one scene per call for ETH
    (test.py:652-656), one batch of 128 per call for NBA (test.py:618).  Rank 0 alone runs; other ranks exit 0 without work."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    from oracle import sttode_oracle as O
    name = args.config
    cfg = CONFIGS[name]
    threads = host_threads()
    sd = O.make_state_dict(cfg["weight_seed"], cfg["tp"], cfg["tf"])
    if name == "eth":
        sample = args.ref_scenes
        items, _, _, _ = make_eth_workload(sample, 0, cfg["tp"], cfg["tf"])
        rows = sum(o.shape[0] for o, _ in items)
        warm = items[: max(2, sample // 8)]
    else:
        nb = max(1, args.ref_scenes // cfg["B"])
        items, _, _ = make_nba_workload(nb, 0, cfg)
        sample, rows, warm = nb * cfg["B"], nb * cfg["B"] * cfg["agents"], items[:1]
    for _ in range(args.warmup):
        cpu_port(name, sd, warm, threads)
    t = [cpu_port(name, sd, items, threads)[1] for _ in range(args.steps)]
    sec = float(np.sum(t))
    value = sample * args.steps / sec
    line = {
        "impl": "reference", "metric": "forecast scenes/sec (K=20 samples)", "value": value, "unit": "scenes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"] + ("; one scene per call like test.py:652-656" if name == "eth" else "; one batch of 128 per call like test.py:618"),
                   "scenes_per_step": sample, "agent_rows_per_step": rows},
        "cpu_baseline": {"value": value, "unit": "scenes/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} scenes x {args.steps} steps of the same workload, PyTorch-CPU oracle port of the reference, {threads} threads"},
        "e2e": {"value": value, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------- our arm: one inference config on all ranks
class Ctx:
    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.local_rank, self.world = dist_env()
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        """Last collective is behind us: tear the group down on every rank (no trailing barrier — a rank that exits early cannot strand a peer)."""
        if self.world > 1:
            torch.cuda.synchronize()
            self.dist.destroy_process_group()


def total_scenes_of(scenes_per_rank, world, steps):
    return scenes_per_rank * world * steps


def inference_leg(ctx, name, args, steps, warmup):
    """Device-resident and end-to-end timing of one inference config on every rank.  Returns (line dict, state for the rank-0 host legs)."""
    import sttode_b200._lib as L
    from sttode_b200 import parallel
    from sttode_b200.net import STTODENet, TRAJ_R2T, TRAJ_RT2
    from sttode_b200.synth import make_args, make_state_dict
    cfg = CONFIGS[name]
    tp, tf = cfg["tp"], cfg["tf"]
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    lib = L.lib()
    sd = make_state_dict(cfg["weight_seed"], tp, tf)
    model = STTODENet(make_args(cfg["dataset"], tp, tf), dev)
    model.load_state_dict(sd, strict=True)
    model.set_device(dev)
    model.eval()
    model.set_tensor_cores(not args.ffma)
    mw = model._weights()

    if name == "eth":
        S = args.scenes                                                   # scenes per rank per step
        items, traj_h, gt_h, sp_h = make_eth_workload(S, rank * S, tp, tf)          # each rank owns its own shard of scenes
        rows = traj_h.shape[0]
        sp_d = sp_h.to(dev)
        counts = (sp_d[1:] - sp_d[:-1]).long()
        rs_d = torch.repeat_interleave(torch.arange(S, device=dev, dtype=torch.int32), counts).contiguous()
        layout, gt_layout, n_scenes, Lg, agents, sp_host = TRAJ_R2T, 1, S, 1, 1, sp_h
        shard = {"scenes_per_gpu_per_step": S}
    else:
        nb = args.nba_batches                                             # whole reference batches per rank per step
        items, past_h, fut_h = make_nba_workload(nb, rank * nb, cfg)
        S = nb * cfg["B"]
        rows = S * cfg["agents"]
        traj_h, gt_h = past_h.view(rows, tp, 2), fut_h.view(rows, tf, 2)
        sp_d = rs_d = sp_host = None
        layout, gt_layout, n_scenes, Lg, agents = TRAJ_RT2, 0, 0, cfg["B"], cfg["agents"]
        shard = {"scenes_per_gpu_per_step": S, "reference_batches_per_gpu_per_step": nb, "batch": cfg["B"], "agents": cfg["agents"]}
    pairs = rows * K
    traj_d, gt_d = traj_h.to(dev), gt_h.to(dev)
    pred_d = torch.empty(rows, K, tf, 2, device=dev)
    ws = L.workspace(lib.sttode_forecast_workspace_bytes(rows, Lg, K, tp), dev)
    sums = torch.zeros(3, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step_device(seed):
        L.check(lib.sttode_forecast(L.dptr(traj_d), layout, L.dptr(sp_d), L.dptr(rs_d), n_scenes, rows, Lg, agents, None, seed, K, C.byref(mw),
                                    L.dptr(pred_d), None, None, None, L.dptr(ws), ws.numel(), stream))
        L.check(lib.sttode_min_ade_fde(L.dptr(pred_d), L.dptr(gt_d), gt_layout, rows, K, tf, L.dptr(sums), stream))

    pred_h = torch.empty(rows, K, tf, 2).pin_memory()
    traj_pin = traj_h.contiguous().pin_memory()

    def step_host(seed):
        L.check(lib.sttode_forecast_host(traj_pin.data_ptr(), layout, sp_host.data_ptr() if sp_host is not None else None, n_scenes, rows, Lg, agents,
                                         seed, K, C.byref(mw), pred_h.data_ptr(), stream))

    # ---- device-resident leg -----------------------------------------------------------------------
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0:
        sampler.start()                      # nvidia-smi needs ~0.5 s to come up: start it before the warm-up
    for i in range(warmup):
        step_device(100 + i)
    ctx.barrier()
    sums.zero_()
    lib.sttode_profile_enable(1)
    launches0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    sampler.mark()
    e0.record()
    for i in range(steps):
        step_device(1000 + i)
    ade, fde, nrow = parallel.reduce_metric_sums(sums)                 # the final metric gather: one all-reduce of 3 doubles
    e1.record()
    ctx.barrier()
    ms = e0.elapsed_time(e1)
    launches = L.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    tot_ms, n_l = C.c_double(), C.c_ulonglong()
    L.check(lib.sttode_profile_read(C.byref(tot_ms), C.byref(n_l)))
    lib.sttode_profile_enable(0)
    ms_max = ctx.max_over_ranks(ms)

    # ---- end-to-end leg (host buffers through the C ABI) ------------------------------------------------
    # timed in the library's default output mode (that is the headline e2e) and, for the record, in the other one
    # (sttode_host_output_mode: 0 = sliced D2H copies overlapped with the kernels, 1 = kernels store into the page-locked buffer)
    default_mode = lib.sttode_host_output_mode(-1)
    e2e_modes = {}
    for mode in (1 - default_mode, default_mode):
        lib.sttode_host_output_mode(mode)
        for i in range(max(1, warmup // 2)):
            step_host(200 + i)
        ctx.barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()                                                    # device clock on the stream the call uses; the call blocks until the predictions are in host memory
        for i in range(steps):
            step_host(2000 + i)
        h1.record()
        torch.cuda.synchronize()
        e2e_modes["direct" if mode else "copy"] = ctx.max_over_ranks(h0.elapsed_time(h1) * 1e-3)
        ctx.barrier()
    e2e_s = e2e_modes["direct" if default_mode else "copy"]

    # ---- evaluation loop end to end (test.py:171-208 as ONE call): host trajectories + ground truth in, min-over-K ADE/FDE reduced on the device, three
    # doubles out — the K samples never cross PCIe, so this leg does not meet the host-memory wall the prediction-returning call meets at >= 4 ranks.
    # A secondary number: `e2e` above stays the headline.  Every rank issues the same collectives whether or not its own calls failed.
    sums_h = (C.c_double * 3)()
    gt_pin = gt_h.contiguous().pin_memory()
    eval_err = None

    def step_eval(seed):
        L.check(lib.sttode_evaluate_host(traj_pin.data_ptr(), gt_pin.data_ptr(), layout, gt_layout, sp_host.data_ptr() if sp_host is not None else None,
                                         n_scenes, rows, Lg, agents, seed, K, C.byref(mw), sums_h, stream))
    try:
        for i in range(max(1, warmup // 2)):
            step_eval(300 + i)
    except Exception as e:                                                # an older prebuilt library, an allocation failure: reported, not fatal
        eval_err = f"{type(e).__name__}: {e}"[:300]
    ctx.barrier()
    v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    v0.record()
    if eval_err is None:
        try:
            for i in range(steps):
                step_eval(3000 + i)
        except Exception as e:
            eval_err = f"{type(e).__name__}: {e}"[:300]
    v1.record()
    torch.cuda.synchronize()
    eval_s = ctx.max_over_ranks(v0.elapsed_time(v1) * 1e-3)
    eval_failed = ctx.max_over_ranks(0.0 if eval_err is None else 1.0) > 0
    ctx.barrier()
    e2e_eval = {"value": None if eval_failed else total_scenes_of(S, world, steps) / eval_s, "unit": "scenes/s", "ms_per_step": None if eval_failed else 1e3 * eval_s / steps,
                "h2d_bytes_per_step": int((traj_h.numel() + gt_h.numel()) * 4 + (sp_h.numel() * 4 + rows * 4 if name == "eth" else 0)), "d2h_bytes_per_step": 24,
                "api": "sttode_evaluate_host (host trajectories + ground truth in; set_data + inference + min-over-K ADE/FDE on the device; the three AverageMeter "
                       "numerators out: the loop body of test.py:171-208 as one call)",
                "min_ade_last_step": None if eval_failed or sums_h[2] == 0 else sums_h[0] / sums_h[2],
                "min_fde_last_step": None if eval_failed or sums_h[2] == 0 else sums_h[1] / sums_h[2], "error": eval_err}
    lib.sttode_release_host_arena()

    # ---- roofline of the dominant kernel ---------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 1590.0
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1590 (B200_PROFILING.md)"
    hbm_peak = peaks.get("hbm_gbs", 6548.2)
    ker_ms = tot_ms.value / max(n_l.value, 1)
    tc = bool(model.decoder.use_tensor_cores)
    if tc and n_l.value:
        ach = issued_flops_dec_pair(tp, tf) * pairs / (ker_ms * 1e-3) / 1e12
    else:                                                               # FFMA build: no in-library event bracket; whole step as the upper bound on kernel time
        ker_ms = ms / steps
        ach = issued_flops_dec_pair(tp, tf) * pairs / (ker_ms * 1e-3) / 1e12
    bpp = DRAM_BYTES_PER_PAIR.get(name)
    algo_bytes = rows * (2048 * 4 + tp * 2 * 4) + pairs * tf * 2 * 4     # rowpre + traj in, pred out
    roofline = {"kernel": "k_dec_pair_tc" if tc else "k_dec_pair", "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": bpp * pairs if bpp else None,
                "traffic_source": "ncu --set full dram read+write per launch (profiles/), scaled by pairs" if bpp else None,
                "traffic_note": "DRAM traffic of this kernel is its algorithmic in/out: rowpre [R,2048] fp32 in (the sample-independent layer-1 partial sums written by "
                                "k_dec_row_tc) + pred out; harmless at < 1 % of HBM peak — the kernel is tensor-pipe bound",
                "algorithmic_bytes_per_launch": algo_bytes,
                "peak_source": peak_src, "kernel_ms": ker_ms, "kernel_share_of_step": ker_ms / (ms / steps),
                "hbm": {"achieved_gbs": algo_bytes / (ker_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak, "frac": algo_bytes / (ker_ms * 1e-3) / 1e9 / hbm_peak,
                        "note": "shown to document that HBM is not the limiter of this kernel (SURVEY.md §8d)"},
                "tensor_ops_issued_frac": 3 * ach / peak_tf,        # 3 tensor-core MACs per algorithmic MAC: share of the tensor pipe's sustained rate
                "flops_per_launch_issued": issued_flops_dec_pair(tp, tf) * pairs,
                "note": ("tcgen05 kind::f16 with the 3xFP16 hi/lo split (fp32-accurate; SURVEY finding 9 rules out single-pass 16/19-bit inputs): "
                         "3 tensor-core MACs per algorithmic MAC, so the attainable ceiling is peak/3; " if tc else "FP32 FFMA kernel; ") +
                        "FLOPs are the ALGORITHMIC ones this kernel issues after exact de-duplication (csrc/decoder.cu header), not the reference's op list"}
    total_scenes = S * world * steps
    line = {
        "metric": "forecast scenes/sec (K=20 samples)", "value": total_scenes / (ms_max * 1e-3), "unit": "scenes/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": dict({"workload": cfg["workload"], "agent_rows_per_gpu_per_step": rows, "sample_pairs_per_gpu_per_step": pairs,
                        "parallelism": f"{'scenes' if name == 'eth' else 'whole reference batches'} sharded over {world} GPU(s), no data-path collective; "
                                       "1 all-reduce of 3 doubles at the end",
                        "l2_policy": f"per-step working set {(ws.numel() + pred_d.numel() * 4) / 2**20:.0f} MiB > 126 MB L2 (inputs+workspace+outputs rewritten every step)",
                        "latent": "drawn in-kernel (Philox) every step, like the reference's randn_like",
                        "flops_per_step_reference_oplist": reference_flops_per_row(tp, tf, Lg) * rows,
                        "flops_per_step_issued": issued_flops_per_row(tp, Lg) * rows + issued_flops_dec_pair(tp, tf) * pairs}, **shard),
        "clocks": clocks,
        "e2e": {"value": total_scenes / e2e_s, "unit": "scenes/s",
                "h2d_bytes_per_step": int(traj_h.numel() * 4 + (sp_h.numel() * 4 + rows * 4 if name == "eth" else 0)),
                "d2h_bytes_per_step": int(pred_h.numel() * 4), "ms_per_step": 1e3 * e2e_s / steps,
                "api": "sttode_forecast_host (host trajectories in, host predictions out)",
                "output_mode": "direct" if default_mode else "copy",
                "ms_per_step_by_output_mode": {k: 1e3 * v / steps for k, v in e2e_modes.items()},
                "limiter": "host-side D2H: every rank streams its predictions to one host's memory; the device-resident value shows the GPU-side scaling" if world > 1 else None},
        "e2e_metrics_only": e2e_eval,
        "gpu_launches": int(launches),
        "roofline": roofline,
        "tc_status": model.tc_status(),
        "quality": {"min_ade": ade, "min_fde": fde, "agents": nrow, "note": "synthetic weights: parity is checked in tests/, not here"},
    }
    state = {"sd": sd, "items": items, "model": model, "cfg": cfg, "name": name}
    del traj_d, gt_d, pred_d, ws, pred_h, traj_pin, gt_pin
    return line, state


# ---------------------------------------------------------------------------------- rank-0 host legs (world == 1 only)
def cpu_baseline_leg(state, budget_s):
    name, threads = state["name"], host_threads()
    cpu_port(name, state["sd"], state["items"][:2 if name == "eth" else 1], threads)                # warm-up (allocator, thread pool)
    done, sec = cpu_port(name, state["sd"], state["items"], threads, budget_s=budget_s)
    how = "one scene per call (test.py:652-656)" if name == "eth" else "one batch of 128 scenes per call (test.py:618)"
    return {"value": done / sec, "unit": "scenes/s", "cores": threads, "kind": "port",
            "sample": f"first {done} scenes of the same workload, {how}, PyTorch-CPU oracle port of the reference, {threads} threads, {sec:.1f} s"}


def gpu_eager_baseline_leg(state, dev, budget_s):
    """Secondary context (SURVEY finding 1: the reference's real bar on a GPU box is PyTorch eager on the GPU): the same oracle port with its
    tensors on the B200 — stock ATen / cuBLAS / cuDNN kernels, one scene (ETH) or one batch (NBA) per call, H2D + D2H per call like test.py."""
    name = state["name"]
    sd = {k: v.to(dev) for k, v in state["sd"].items()}
    cpu_port(name, sd, state["items"][:3 if name == "eth" else 1], host_threads(), device=dev)
    torch.cuda.synchronize()
    done, sec = cpu_port(name, sd, state["items"], host_threads(), budget_s=budget_s, device=dev)
    return {"value": done / sec, "unit": "scenes/s", "kind": "port on torch-eager CUDA (stock ATen/cuBLAS/cuDNN kernels), same B200",
            "sample": f"{done} scenes, {sec:.1f} s, per-call H2D + D2H like test.py:175,194"}


def parity_leg(state, dev):
    """ADE/FDE delta vs the reference (second half of BASELINE.json's metric): same scenes, same injected latent, GPU path vs CPU oracle port."""
    from oracle import sttode_oracle as O
    from sttode_b200 import synth
    name, sd, model, cfg = state["name"], state["sd"], state["model"], state["cfg"]
    with torch.no_grad():
        if name == "eth":
            npar = 24
            scenes = state["items"][:npar]
            zs, ade_ref, fde_ref, n_ref = [], 0.0, 0.0, 0
            for i, (obs, gt) in enumerate(scenes):
                n = obs.shape[0]
                z = synth.latent(n, seed=500 + i)
                zs.append(z)
                past, orig = O.eth_scene_normalise(obs)
                pr = O.inference(sd, past, 1, n, z, sample_k=K, scene_orig=orig, fast=True)            # [K, n, T_f, 2]
                a, f = O.min_ade_fde(pr.permute(1, 0, 2, 3), gt.permute(0, 2, 1))
                ade_ref, fde_ref, n_ref = ade_ref + a * n, fde_ref + f * n, n_ref + n
            ade_ref, fde_ref = ade_ref / n_ref, fde_ref / n_ref
            obs_p = torch.cat([o for o, _ in scenes]).to(dev)
            gt_p = torch.cat([g for _, g in scenes])
            sp_p = torch.tensor(np.concatenate([[0], np.cumsum([o.shape[0] for o, _ in scenes])]), dtype=torch.int32)
            model.z_override = torch.cat(zs)
            pr_g = model.forecast_scenes(obs_p, sp_p).cpu()                                            # [rows, K, T_f, 2]
            model.z_override = None
            ade_g, fde_g = O.min_ade_fde(pr_g, gt_p.permute(0, 2, 1))
            what = f"{npar} scenes"
        else:
            b = state["items"][0]
            B, N, tp = cfg["B"], cfg["agents"], cfg["tp"]
            z = synth.latent(B * N, seed=501)
            pr = O.inference(sd, b["past_traj"].view(B * N, tp, 2), B, N, z, sample_k=K, fast=True)
            gt = b["future_traj"].view(B * N, cfg["tf"], 2)
            ade_ref, fde_ref = O.min_ade_fde(pr.permute(1, 0, 2, 3), gt)
            model.z_override = z
            model.set_data_nba(b)
            pr_g = model.inference(b).permute(1, 0, 2, 3).cpu()
            model.z_override = None
            ade_g, fde_g = O.min_ade_fde(pr_g, gt)
            n_ref, what = B * N, "one reference batch of 128 scenes"
    return {"sample": what, "agents": n_ref, "min_ade_reference": ade_ref, "min_ade_ours": ade_g, "min_fde_reference": fde_ref, "min_fde_ours": fde_g,
            "rel_delta_ade": abs(ade_g - ade_ref) / ade_ref, "rel_delta_fde": abs(fde_g - fde_ref) / fde_ref,
            "note": "same scenes, same injected latent z; reference = PyTorch-CPU oracle port (pinned to the reference's outputs in tests/golden); tolerance 1e-4"}


FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12        # FFMA: 148 SMs x 128 lanes x 2 FLOP at the 1965 MHz boost clock = 74.4 TFLOP/s


def latency_extras(model, sd, dev, cpu_seconds):
    """configs[0] and configs[3] (single process): one ETH scene per call the way test.py drives the model (test.py:164-208), and the SDD-shape
    sweep — end to end with one scene of N agents per call, plus the Gram-tile stress on the MHGSA module alone.  Each entry is a full line
    (metric / value / unit / roofline / cpu_baseline): at these sizes the bound is the latency of the dependent kernel chain, and the line says so."""
    from sttode_b200 import synth
    tp, tf = 8, 12
    out = {}

    def lat(fn, n=200):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return 1e3 * e0.elapsed_time(e1) / n                      # microseconds per call, device clock

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 1590.0

    def chain_roofline(n_agents, us):
        fl = (issued_flops_per_row(tp) + K * issued_flops_dec_pair(tp, tf)) * n_agents
        ach = fl / (us * 1e-6) / 1e12
        return {"bound": "latency", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": None,
                "flops_per_call_issued": fl, "flops_per_call_reference_oplist": reference_flops_per_row(tp, tf) * n_agents, "kernels_per_call": 5,
                "note": "one scene fills a fraction of one wave (rows / 128 row tiles, rows / 6 pair tiles on 148 SMs): the call is the latency of a chain of "
                        "dependent kernels (scene origin -> encoder [folded front end + ODE block] -> decoder row kernel -> decoder pair kernel, + FP32 redo "
                        "launches that exit at once), replayed as ONE CUDA graph; peak = the bf16 tensor peak the dominant kernel is measured against"}

    with torch.no_grad():
        obs7, gt7 = synth.eth_scene(3, 7)
        obs7d = obs7.to(dev)
        sp7 = torch.tensor([0, 7], dtype=torch.int32, device=dev)
        g7 = model.capture_forecast(rows=7, L=1, agents=1, n_scenes=1)
        us_graph, us_eager = lat(lambda: g7(obs7d, sp7)), lat(lambda: model.forecast_scenes(obs7d, sp7))
        scenes7 = [synth.eth_scene(3 + i, 7) for i in range(400)]
        cpu_port_eth(sd, scenes7[:3], host_threads())
        done, sec = cpu_port_eth(sd, scenes7, host_threads(), budget_s=min(3.0, cpu_seconds))
        out["eth_single_scene"] = {
            "metric": "forecast scenes/sec (K=20 samples)", "value": 1e6 / us_graph, "unit": "scenes/s", "n_gpus": 1, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[0] shape: ETH scene of N = 7 agents, ONE scene per call like test.py:164-208, obs 8 / pred 12, K=20",
                       "api": "ForecastGraph (one CUDA-graph launch per scene, inputs already on the device)"},
            "us_per_scene_cuda_graph": us_graph, "us_per_scene_eager_c_abi": us_eager, "roofline": chain_roofline(7, us_graph),
            "cpu_baseline": {"value": done / sec, "unit": "scenes/s", "cores": host_threads(), "kind": "port",
                             "sample": f"{done} scenes of N = 7, one per call, PyTorch-CPU oracle port, {sec:.1f} s ({1e3 * sec / done:.2f} ms per scene)"}}
        sweep = {}
        for n_ag in (16, 64, 256, 512):
            obs_n = (torch.rand(n_ag, 2, tp, generator=torch.Generator().manual_seed(n_ag)) * 30.0).to(dev)
            sp_n = torch.tensor([0, n_ag], dtype=torch.int32, device=dev)
            gn = model.capture_forecast(rows=n_ag, L=1, agents=1, n_scenes=1)
            us = lat(lambda: gn(obs_n, sp_n), n=50)
            sweep[f"N{n_ag}"] = {"us_per_scene_cuda_graph": us, "scenes_per_s": 1e6 / us, "agents_per_s": n_ag * 1e6 / us, "roofline": chain_roofline(n_ag, us)}
        attn = model.past_encoder.ODE_Encoder.odeblock.odefunc.layers[0].self_attn.temporal_attention_before
        for Lg in (128, 256, 512):
            xg = torch.randn(Lg, 64, 64, generator=torch.Generator().manual_seed(Lg)).to(dev)
            us_g = lat(lambda: attn(xg, xg, xg, need_weights=False), n=50)
            entries = 64 * 8 * Lg * Lg                          # (agent, head, i, j) score entries
            fl = entries * 32 + Lg * 64 * 2 * 64 * 256           # Gram 16 + AV 16 FLOP per entry; in/out projections 2*64*(192+64) per row
            ach = fl / (us_g * 1e-6) / 1e12
            sweep[f"mhgsa_L{Lg}_bsz64"] = {"us_per_call": us_g, "score_entries_per_s": entries / (us_g * 1e-6),
                                           "roofline": {"bound": "fp32+sfu", "achieved": ach, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": ach / FP32_PEAK_TFLOPS,
                                                        "traffic": None, "note": "Hyp_mhsa module alone (k_proj -> k_geo_attn -> k_proj): per score entry 32 FFMA-FLOP plus one acos "
                                                                                 "(polynomial, ~20 FP32 ops) and one exp (MUFU); the K = 8 Gram stays on FFMA by design"}}
        out["sdd_sweep"] = {"metric": "forecast scenes/sec (K=20 samples)", "unit": "scenes/s", "n_gpus": 1, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
                            "config": {"workload": "BASELINE.json configs[3]: SDD-shape large-scene sweep, one scene of N in {16..512} agents per CUDA-graph launch "
                                                   "(S stays 1: stresses rows); Gram-tile stress on the MHGSA module alone, L = S in {128, 256, 512}, bsz 64"},
                            "value": sweep["N512"]["scenes_per_s"], "sweep": sweep}
    return out


# ---------------------------------------------------------------------------------- training leg (configs[4]) on all ranks
def train_leg(ctx, args, steps, warmup):
    """STTODE training step — forward + backward through both MHGSA / tensor-ODE blocks and the decoder in libsttode_sm100, data-parallel:
    gradients summed over ranks, Adam.  NBA shape (B = 32 scenes x 11 agents per rank, obs 10 / pred 20, train.py:24), synthetic batches,
    reference default init (train.py:110 seed)."""
    import sttode_b200._lib as L
    from sttode_b200 import parallel, synth
    from sttode_b200.net import STTODENet
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    B, N, tp, tf = args.train_batch, 11, 10, 20
    torch.manual_seed(1)                                         # identical initial weights on every rank
    model = STTODENet(synth.make_args("nba", tp, tf), torch.device("cpu"))
    model.set_device(dev)
    model.train()
    graph = not args.train_eager
    fused = args.train_exchange == "fused"
    if fused:                                                    # gradient exchange over NVLink peer memory fused with Adam: no NCCL on the step
        reducer, opt = parallel.FusedDPAdam(model, lr=1e-4), None
    else:
        reducer = parallel.FlatGradReducer(model)
        opt, sched = parallel.make_optimizer(model, capturable=graph)
    torch.manual_seed(1000 + rank)                               # per-rank dropout / rsample streams
    batches = [{k: v.to(dev) for k, v in synth.nba_batch(100 + rank * 64 + i, B, N, tp, tf).items()} for i in range(8)]
    it = [0]
    if graph:
        graphed = parallel.GraphedTrainStep(model, opt, reducer, batches[0], warmup=3)

        def one_step():
            it[0] += 1
            return graphed(batches[it[0] % len(batches)])
    else:
        def stage():
            model.set_data_nba(batches[it[0] % len(batches)])
            it[0] += 1

        def one_step():
            return parallel.train_step(model, opt, reducer, stage)
    for _ in range(max(warmup, 3)):
        out = one_step()
    ctx.barrier()
    l0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    e0.record()
    for _ in range(steps):
        out = one_step()
    e1.record()
    ctx.barrier()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1)) / steps
    # replica-equality check across ranks: after the exchange every rank must hold the same averaged gradient (NCCL path) / the same weights
    # (fused path: all ranks sum the buffers in the same order, so the replicas stay bit-identical)
    spread, exch_err = 0.0, 0
    if world > 1:
        g = (reducer.flat_param if fused else reducer.flat).detach().clone()
        gmax, gmin = g.clone(), g.clone()
        ctx.dist.all_reduce(gmax, op=ctx.dist.ReduceOp.MAX)
        ctx.dist.all_reduce(gmin, op=ctx.dist.ReduceOp.MIN)
        spread = float((gmax - gmin).abs().max().item())
    if fused:
        exch_err, _ = reducer.status()
    loss = float(out[0].detach()) if not torch.is_tensor(out) else float(out[0])
    line = {"metric": "training scenes/sec (DP, forward+backward+Adam)", "value": B * world / (ms * 1e-3), "unit": "scenes/s", "n_gpus": world,
            "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "dtype": "f32", "data": "synthetic",
            "config": {"workload": "STTODE training step (BASELINE.json configs[4]), NBA shape obs 10 / pred 20", "batch_per_gpu": B, "agents": N,
                       "parallelism": f"dp{world}: {parallel.describe_reduce(reducer, graph)}"},
            "gpu_launches": int(L.launch_count() - l0) if not graph else f"{getattr(graphed, 'graph_replays_per_step', 1)} CUDA-graph replay(s) per step",
            "library_kernels_per_step": getattr(graphed, "library_launches", None) if graph else int(L.launch_count() - l0) // steps,
            "cuda_graph": graph, "backward_gemm": args.train_gemm, "last_loss": loss, "replica_spread_across_ranks": spread, "exchange": args.train_exchange, "exchange_error_word": exch_err,
            "note": "every layer forward+backward in libsttode_sm100 (ODE blocks, dense layers, decoder incl. GRU through time)"}
    state = {"model": model, "B": B, "N": N, "tp": tp, "tf": tf}
    if fused:
        ctx.barrier()                                            # nobody unmaps an arena a peer may still read
        reducer.close()
    return line, state


def train_cpu_baseline(state):
    """The oracle port of the reference's step (forward + torch autograd backward), all host threads, one timed step."""
    from oracle import sttode_oracle as O
    from sttode_b200 import synth
    B, N, tp, tf = state["B"], state["N"], state["tp"], state["tf"]
    threads = host_threads()
    torch.set_num_threads(threads)
    sd = {k: v.detach().cpu().clone() for k, v in state["model"].state_dict().items()}
    for k in sd:
        if not k.endswith("pos_encoder.pe"):
            sd[k].requires_grad_(True)
    data = synth.nba_batch(100, B, N, tp, tf)
    g = torch.Generator().manual_seed(5)
    eps_q, eps_p = torch.randn(B * N, 32, generator=g), torch.randn(B * N * 20, 32, generator=g)
    masks = ((torch.rand(B * N, tp, 64, generator=g) >= 0.1).float(), (torch.rand(B * N, tf, 64, generator=g) >= 0.1).float())
    secs = []
    for _ in range(3):
        t0 = time.perf_counter()
        out_cpu = O.training_objective(sd, data["past_traj"].view(B * N, tp, 2), data["future_traj"].view(B * N, tf, 2), B, N, eps_q, eps_p, masks=masks,
                                       tensors=True)
        out_cpu["total"].backward()
        secs.append(time.perf_counter() - t0)
    return {"value": B / min(secs[1:]), "unit": "scenes/s", "cores": threads, "kind": "port",
            "sample": f"one training step (B = {B}) of the PyTorch-CPU oracle port: forward + autograd backward, no optimizer, {min(secs[1:]):.2f} s"}


# ---------------------------------------------------------------------------------- driver
def run_ours(args):
    ctx = Ctx()
    rank, world = ctx.rank, ctx.world
    if args.train:
        line, tstate = train_leg(ctx, args, args.steps, args.warmup)
        ctx.close()
        if rank != 0:
            return
        line["cpu_baseline"] = train_cpu_baseline(tstate) if world == 1 and not args.no_extra else None
        print(json.dumps(line), flush=True)
        return
    line, state = inference_leg(ctx, args.config, args, args.steps, args.warmup)
    others, states = {}, {}
    if not args.no_extra:
        # the other inference config and the DP training step, measured briefly by ALL ranks (same collectives in the same order everywhere)
        other = "nba" if args.config == "eth" else "eth"
        torch.cuda.empty_cache()
        others[other], states[other] = inference_leg(ctx, other, args, max(3, args.steps // 2), 3)
        torch.cuda.empty_cache()
        # The training leg is secondary: it must never cost the headline line.  Every rank arms a dead-man timer; if the leg has not finished
        # in time (a stuck collective), rank 0 prints the line without it and every rank leaves with status 0 instead of waiting for the NCCL watchdog.
        def give_up():
            if rank == 0:
                others["train"] = {"error": f"training leg did not finish within {args.train_timeout:.0f} s; skipped"}
                line["cpu_baseline"], line["configs"] = None, others
                print(json.dumps(line), flush=True)
            os._exit(0)
        timer = threading.Timer(args.train_timeout, give_up)
        timer.daemon = True
        timer.start()
        try:
            others["train"], states["train"] = train_leg(ctx, args, max(5, args.steps), 3)
        except Exception as e:                                   # secondary leg: report, never lose the headline line (an asymmetric failure is caught by the timer)
            others["train"] = {"error": repr(e)[:300]}
        finally:
            timer.cancel()
    ctx.close()                                                   # the last collective is done: no rank waits on another from here on
    if rank != 0:
        return
    dev = ctx.dev
    if world == 1:
        # host legs: single process only (BASELINE: "on rank 0 at N=1 only"); bounded by wall time
        line["cpu_baseline"] = cpu_baseline_leg(state, args.cpu_seconds)
        line["ade_fde_delta_vs_reference"] = parity_leg(state, dev)
        if not args.no_extra:
            for k in ("eth", "nba"):
                if k in others and "error" not in others[k]:
                    others[k]["cpu_baseline"] = cpu_baseline_leg(states[k], max(4.0, args.cpu_seconds / 2))
                    others[k]["ade_fde_delta_vs_reference"] = parity_leg(states[k], dev)
            if "train" in states:
                others["train"]["cpu_baseline"] = train_cpu_baseline(states["train"])
            try:
                line["gpu_eager_baseline"] = gpu_eager_baseline_leg(state, dev, 6.0)
            except Exception as e:                                 # secondary context only
                line["gpu_eager_baseline"] = {"unavailable": repr(e)[:200]}
            eth_state = state if args.config == "eth" else states["eth"]
            others.update(latency_extras(eth_state["model"], eth_state["sd"], dev, args.cpu_seconds))
    else:
        line["cpu_baseline"] = None
        line["cpu_baseline_note"] = "timed at N = 1 only (task contract); see the --impl reference arm for this N"
    line["configs"] = others
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="eth", choices=["eth", "nba"], help="headline inference workload: eth = BASELINE configs[1], nba = configs[2]")
    ap.add_argument("--scenes", type=int, default=1536, help="ETH scenes per GPU per step")
    ap.add_argument("--nba-batches", type=int, default=32, help="NBA reference batches (128 scenes each) per GPU per step")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="wall-time budget of the CPU-baseline sample")
    ap.add_argument("--ref-scenes", type=int, default=96, help="scenes per step of the --impl reference arm (NBA: rounded to batches of 128)")
    ap.add_argument("--no-extra", action="store_true", help="headline config only: skip the other configs, the training leg and the latency extras")
    ap.add_argument("--train", action="store_true", help="time only the data-parallel training step (configs[4])")
    ap.add_argument("--train-eager", action="store_true", help="training step launch by launch instead of CUDA-graph replay")
    ap.add_argument("--train-timeout", type=float, default=150.0, help="dead-man timer of the secondary training leg (seconds)")
    ap.add_argument("--train-exchange", default="fused", choices=["fused", "nccl"],
                    help="DP gradient exchange: fused = sttode_dp_adam_step over NVLink peer memory (all-reduce + Adam in one kernel, whole step one CUDA graph); "
                         "nccl = one NCCL all-reduce between a forward+backward graph and an Adam graph")
    ap.add_argument("--train-gemm", default="ffma", choices=["ffma", "tc"],
                    help="backward GEMMs: ffma = FP32 kernels (parity-grade, default); tc = tcgen05 bf16x3 kernels (STTODE_TC_GEMM=1: faster, ~2e-6 per product)")
    ap.add_argument("--train-batch", type=int, default=32, help="scenes per GPU per training step (train.py:24)")
    ap.add_argument("--ffma", action="store_true", help="use the FP32-FFMA decoder tile instead of the tcgen05 one")
    args = ap.parse_args()
    if args.train_gemm == "tc":
        os.environ["STTODE_TC_GEMM"] = "1"          # read once by the library at its first GEMM
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
