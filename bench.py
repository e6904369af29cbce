#!/usr/bin/env python
"""bench.py — forecast scenes/sec (K = 20 samples) of the STTODE forward path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scenes S]

Workload (BASELINE.json configs[1]): ETH-UCY zara/univ-shaped batched inference — scenes of
N in {10,20,30,40,50,60} agents, obs 8 / pred 12 frames, K = 20 samples, synthetic trajectories
(sttode_b200/synth.py) and seeded synthetic weights.  One step = one pass of the whole
`STTODENet.inference` hot path (scene normalisation, past encoder with the MHGSA/tensor-ODE block,
in-kernel latent draw, K-sample decoder, min-ADE/FDE reduction) over S scenes per GPU.
N > 1: scenes shard across ranks (weak scaling, no data-path collective); one NCCL all-reduce of the three
metric sums closes the timed region.

Output: ONE JSON line on rank 0 (contract in the task statement): value = whole-job scenes/s with inputs
resident in HBM; e2e = the same through sttode_forecast_host (host buffers, H2D + D2H inside the timed
region); roofline for the dominant kernel (k_dec_pair) from CUDA events; cpu_baseline = the oracle port of
the reference timed on the host cores.  `--impl reference` times that CPU port alone.
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

TP, TF, K = 8, 12, 20
CYCLE = (10, 20, 30, 40, 50, 60)
WEIGHT_SEED = 1
DRAM_BYTES_PER_PAIR = 97.0e6 / 268800     # ncu --set full of the dominant kernel (profiles/r01_ncu_full_k_dec_pair_tc.md)


# ---------------------------------------------------------------------------------- FLOP accounting
def flops_mlp(k1, n3):
    return 2 * k1 * 512 + 2 * 512 * 256 + 2 * 256 * n3


def flops_gru(T):
    return T * (2 * 6 * 32 + 2 * (32 + 96) * 288)


def issued_flops_dec_pair(tp, tf):
    """FLOPs k_dec_pair issues per (agent, sample) pair after de-duplication (csrc/decoder.cu header)."""
    return flops_mlp(32, 2 * tp) + flops_gru(tp) + flops_mlp(32, 2 * tf) + flops_mlp(128, 2 * tf)


def issued_flops_per_row(tp):
    enc = 25088 * tp + 8576                                       # front end (SURVEY §8d enc_pre)
    ode = 2 * 64 * 64 * 4 + 2 * 2 * 64 * 1024                     # L = 1: v-proj, out-proj, 2 gates, FFN
    dec_row = flops_gru(tp) + 2 * (2 * 224 * 512) + 2 * (2 * 128 * 512)
    return enc + ode + dec_row


def reference_flops_per_row(tp, tf, S=1):
    """Algorithmic matmul FLOPs of the reference's own op list per agent row (SURVEY §8d F_inference / R)."""
    dec_blk = 74112 * tp + 1048576 + 1024 * (tp + tf)
    return (25088 * tp + 8576) + (311296 + 256 * S) + K * 2 * dec_blk


# ---------------------------------------------------------------------------------- helpers
def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def make_workload(n_scenes, start):
    from sttode_b200 import synth
    scenes = synth.eth_scenes(n_scenes, cycle=CYCLE, past_length=TP, future_length=TF, start=start)
    counts = [o.shape[0] for o, _ in scenes]
    obs = torch.cat([o for o, _ in scenes]).contiguous()
    gt = torch.cat([g for _, g in scenes]).contiguous()
    sp = torch.tensor(np.concatenate([[0], np.cumsum(counts)]), dtype=torch.int32)
    return scenes, obs, gt, sp


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
        sm, smax, reasons = [], None, set()
        t_mark = getattr(self, "t_mark", 0.0)
        for ts, ln in self.lines:
            if ts < t_mark:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def time_cpu_port(sd, scenes, threads):
    """The reference's per-scene call sequence (set_data + inference, test.py:171-187) on the CPU oracle port."""
    from oracle import sttode_oracle as O
    from sttode_b200 import synth
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    with torch.no_grad():
        for i, (obs, _) in enumerate(scenes):
            n = obs.shape[0]
            past, orig = O.eth_scene_normalise(obs)
            z = torch.randn(n * K, 32)
            O.inference(sd, past, 1, n, z, sample_k=K, scene_orig=orig, fast=True)
    return time.perf_counter() - t0


# ---------------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    from oracle import sttode_oracle as O
    threads = os.cpu_count() or 1
    sd = O.make_state_dict(WEIGHT_SEED, TP, TF)
    sample = args.ref_scenes
    scenes, _, _, _ = make_workload(sample, 0)
    for _ in range(args.warmup):
        time_cpu_port(sd, scenes[: max(2, sample // 8)], threads)
    t = [time_cpu_port(sd, scenes, threads) for _ in range(args.steps)]
    sec = float(np.sum(t))
    value = sample * args.steps / sec
    rows = sum(o.shape[0] for o, _ in scenes)
    line = {
        "impl": "reference", "metric": "forecast scenes/sec (K=20 samples)", "value": value, "unit": "scenes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ETH-UCY zara/univ-shape inference, N in {10..60}, obs 8 / pred 12, K=20; one scene per call like test.py",
                   "scenes_per_step": sample, "agent_rows_per_step": rows},
        "cpu_baseline": {"value": value, "unit": "scenes/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} scenes x {args.steps} steps of the same workload, PyTorch-CPU oracle port of the reference, {threads} threads"},
        "e2e": {"value": value, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist
    import sttode_b200._lib as L
    from sttode_b200 import parallel
    from sttode_b200.net import STTODENet, TRAJ_R2T
    from sttode_b200.synth import make_args, make_state_dict

    rank, local_rank, world = dist_env()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = L.lib()

    sd = make_state_dict(WEIGHT_SEED, TP, TF)
    model = STTODENet(make_args("eth", TP, TF), dev)
    model.load_state_dict(sd, strict=True)
    model.set_device(dev)
    model.eval()
    model.decoder.use_tensor_cores = not args.ffma
    mw = model._weights()

    S = args.scenes
    scenes, obs_h, gt_h, sp_h = make_workload(S, rank * S)             # each rank owns its own shard of scenes
    rows = obs_h.shape[0]
    obs_d, gt_d, sp_d = obs_h.to(dev), gt_h.to(dev), sp_h.to(dev)
    counts = (sp_d[1:] - sp_d[:-1]).long()
    rs_d = torch.repeat_interleave(torch.arange(S, device=dev, dtype=torch.int32), counts).contiguous()
    pred_d = torch.empty(rows, K, TF, 2, device=dev)
    ws = L.workspace(lib.sttode_forecast_workspace_bytes(rows, 1, K, TP), dev)
    sums = torch.zeros(3, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step_device(seed):
        L.check(lib.sttode_forecast(L.dptr(obs_d), TRAJ_R2T, L.dptr(sp_d), L.dptr(rs_d), S, rows, 1, 1, None, seed, K, C.byref(mw),
                                    L.dptr(pred_d), None, None, None, L.dptr(ws), ws.numel(), stream))
        L.check(lib.sttode_min_ade_fde(L.dptr(pred_d), L.dptr(gt_d), 1, rows, K, TF, L.dptr(sums), stream))

    pred_h = torch.empty(rows, K, TF, 2).pin_memory()
    obs_pin = obs_h.pin_memory()

    def step_host(seed):
        L.check(lib.sttode_forecast_host(obs_pin.data_ptr(), TRAJ_R2T, sp_h.data_ptr(), S, rows, 1, 1, seed, K, C.byref(mw),
                                         pred_h.data_ptr(), stream))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg -----------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                      # nvidia-smi needs ~0.5 s to come up: start it before the warm-up
    for i in range(args.warmup):
        step_device(100 + i)
    barrier()
    sums.zero_()
    lib.sttode_profile_enable(1)
    launches0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    e0.record()
    for i in range(args.steps):
        step_device(1000 + i)
    ade, fde, nrow = parallel.reduce_metric_sums(sums)                 # the final metric gather: one all-reduce of 3 doubles
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    tot_ms, n_l = C.c_double(), C.c_ulonglong()
    L.check(lib.sttode_profile_read(C.byref(tot_ms), C.byref(n_l)))
    lib.sttode_profile_enable(0)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    # ---- end-to-end leg (host buffers through the C ABI) ------------------------------------------------
    for i in range(max(1, args.warmup // 2)):
        step_host(200 + i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_host(2000 + i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 1590.0
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1590 (B200_PROFILING.md)"
    pairs = rows * K
    ker_ms = tot_ms.value / max(n_l.value, 1)
    ach = issued_flops_dec_pair(TP, TF) * pairs / (ker_ms * 1e-3) / 1e12
    tc = bool(model.decoder.use_tensor_cores)
    roofline = {"kernel": "k_dec_pair_tc" if tc else "k_dec_pair", "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": DRAM_BYTES_PER_PAIR * pairs, "traffic_source": "profiles/r01_ncu_full_k_dec_pair_tc.md: dram read+write per 268 800-pair launch, scaled by pairs",
                "peak_source": peak_src, "kernel_ms": ker_ms, "kernel_share_of_step": ker_ms / (ms / args.steps),
                "hbm": {"achieved_gbs": DRAM_BYTES_PER_PAIR * pairs / (ker_ms * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs", 6548.2),
                        "frac": DRAM_BYTES_PER_PAIR * pairs / (ker_ms * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6548.2),
                        "note": "shown to document that HBM is not the limiter of this kernel (SURVEY.md §8d)"},
                "tensor_ops_issued_frac": 3 * ach / peak_tf,        # 3 tensor-core MACs per algorithmic MAC: share of the tensor pipe's sustained rate
                "flops_per_launch_issued": issued_flops_dec_pair(TP, TF) * pairs,
                "note": ("tcgen05 kind::f16 with the 3xFP16 hi/lo split (fp32-accurate; SURVEY finding 9 rules out single-pass 16/19-bit inputs): "
                         "3 tensor-core MACs per algorithmic MAC, so the attainable ceiling is peak/3; " if tc else "FP32 FFMA kernel; ") +
                        "FLOPs are the ALGORITHMIC ones this kernel issues after exact de-duplication (csrc/decoder.cu header), not the reference's op list"}
    step_flops_ref = reference_flops_per_row(TP, TF) * rows
    step_flops_issued = issued_flops_per_row(TP) * rows + issued_flops_dec_pair(TP, TF) * pairs

    # ---- CPU baseline (bounded sample of the same workload, oracle port, all host threads) ---------------
    threads = os.cpu_count() or 1
    nsmp = args.cpu_scenes
    time_cpu_port(sd, scenes[:4], threads)
    cpu_s = time_cpu_port(sd, scenes[:nsmp], threads)
    cpu_val = nsmp / cpu_s

    # ---- ADE/FDE delta vs the reference (second half of BASELINE.json's metric): same scenes, same injected latent, GPU path vs CPU oracle port ---
    from oracle import sttode_oracle as O
    from sttode_b200 import synth as _synth
    npar = 24
    with torch.no_grad():
        zs, ade_ref, fde_ref, n_ref = [], 0.0, 0.0, 0
        for i, (obs, gt) in enumerate(scenes[:npar]):
            n = obs.shape[0]
            z = _synth.latent(n, seed=500 + i)
            zs.append(z)
            past, orig = O.eth_scene_normalise(obs)
            pr = O.inference(sd, past, 1, n, z, sample_k=K, scene_orig=orig, fast=True)            # [K, n, T_f, 2]
            a, f = O.min_ade_fde(pr.permute(1, 0, 2, 3), gt.permute(0, 2, 1))
            ade_ref, fde_ref, n_ref = ade_ref + a * n, fde_ref + f * n, n_ref + n
        obs_p = torch.cat([o for o, _ in scenes[:npar]]).to(dev)
        gt_p = torch.cat([g for _, g in scenes[:npar]])
        sp_p = torch.tensor(np.concatenate([[0], np.cumsum([o.shape[0] for o, _ in scenes[:npar]])]), dtype=torch.int32)
        model.z_override = torch.cat(zs)
        pr_g = model.forecast_scenes(obs_p, sp_p).cpu()                                            # [rows, K, T_f, 2]
        model.z_override = None
        ade_g, fde_g = O.min_ade_fde(pr_g, gt_p.permute(0, 2, 1))
    ade_ref, fde_ref = ade_ref / n_ref, fde_ref / n_ref
    parity = {"scenes": npar, "agents": n_ref, "min_ade_reference": ade_ref, "min_ade_ours": ade_g, "min_fde_reference": fde_ref, "min_fde_ours": fde_g,
              "rel_delta_ade": abs(ade_g - ade_ref) / ade_ref, "rel_delta_fde": abs(fde_g - fde_ref) / fde_ref,
              "note": "same scenes, same injected latent z; reference = PyTorch-CPU oracle port (pinned to the reference's outputs in tests/golden); tolerance 1e-4"}

    # ---- other configs of BASELINE.json, measured briefly for DESIGN.md (not the headline): rank 0 only -----------------
    extra = {}
    if not args.no_extra:
        from sttode_b200 import synth
        with torch.no_grad():
            # configs[0] shape: one ETH scene (N = 7) per call, the way test.py drives the model; eager C-ABI call vs one CUDA-graph launch
            obs7, _ = synth.eth_scene(3, 7)
            obs7 = obs7.to(dev)
            sp7 = torch.tensor([0, 7], dtype=torch.int32, device=dev)

            def lat(fn, n=200):
                for _ in range(20):
                    fn()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(n):
                    fn()
                torch.cuda.synchronize()
                return 1e6 * (time.perf_counter() - t0) / n
            g7 = model.capture_forecast(rows=7, L=1, agents=1, n_scenes=1)
            extra["eth_single_scene_N7"] = {"eager_us_per_call": lat(lambda: model.forecast_scenes(obs7, sp7)),
                                            "cuda_graph_us_per_call": lat(lambda: g7(obs7, sp7)), "kernels_per_call": 5}
            # configs[2]: NBA shape, 11 agents, obs 10 / pred 20, reference test batch B = 128 (scenes of a batch attend to each other: MHGSA L = S = 128)
            nba = STTODENet(make_args("nba", 10, 20), dev)
            nba.load_state_dict(make_state_dict(2, 10, 20), strict=True)
            nba.set_device(dev)
            nba.eval()
            nb = 16
            batches = torch.stack([synth.nba_batch(i, 128, 11, 10, 20)["past_traj"] for i in range(nb)]).to(dev)      # [nb,128,11,10,2]
            gnba = nba.capture_forecast(rows=nb * 128 * 11, L=128, agents=11, n_scenes=0, layout=0)
            us = lat(lambda: gnba(batches.view(-1, 10, 2)), n=20)
            extra["nba_B128_10_20"] = {"scenes_per_s": nb * 128 / (us * 1e-6), "batches_per_launch": nb, "us_per_launch": us,
                                       "note": "16 reference batches per launch, each batch = one MHGSA group of L = S = 128 scenes x 11 agents"}
            # configs[3]: SDD-shape large-scene sweep.  (i) end to end: one scene of N agents per call (S stays 1: stresses rows);
            # (ii) Gram-tile stress: the MHGSA module alone on [L, bsz, 64] with L = S up to 512 (one 128-query tile per (agent, head) CTA)
            sweep = {}
            for n_ag in (16, 64, 256, 512):
                obs_n = (torch.rand(n_ag, 2, TP, generator=torch.Generator().manual_seed(n_ag)) * 30.0).to(dev)
                sp_n = torch.tensor([0, n_ag], dtype=torch.int32, device=dev)
                gn = model.capture_forecast(rows=n_ag, L=1, agents=1, n_scenes=1)
                sweep[f"N{n_ag}"] = {"cuda_graph_us_per_scene": lat(lambda: gn(obs_n, sp_n), n=50)}
            attn = model.past_encoder.ODE_Encoder.odeblock.odefunc.layers[0].self_attn.temporal_attention_before
            for Lg in (128, 256, 512):
                xg = torch.randn(Lg, 64, 64, generator=torch.Generator().manual_seed(Lg)).to(dev)
                us_g = lat(lambda: attn(xg, xg, xg, need_weights=False), n=50)
                pairs_g = 64 * 8 * Lg * Lg                       # (agent, head, i, j) score entries
                sweep[f"mhgsa_L{Lg}_bsz64"] = {"us_per_call": us_g, "score_entries_per_s": pairs_g / (us_g * 1e-6)}
            extra["sdd_sweep"] = sweep

    total_scenes = S * world * args.steps
    value = total_scenes / (ms_max * 1e-3)
    line = {
        "metric": "forecast scenes/sec (K=20 samples)", "value": value, "unit": "scenes/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ETH-UCY zara/univ-shape batched inference (BASELINE.json configs[1]): N in {10..60} agents/scene, obs 8 / pred 12, K=20",
                   "scenes_per_gpu_per_step": S, "agent_rows_per_gpu_per_step": rows, "sample_pairs_per_gpu_per_step": pairs,
                   "parallelism": f"scenes sharded over {world} GPU(s), no data-path collective; 1 all-reduce of 3 doubles at the end",
                   "l2_policy": f"per-step working set {(ws.numel() + pred_d.numel() * 4) / 2**20:.0f} MiB > 126 MB L2 (inputs+workspace+outputs rewritten every step)",
                   "latent": "drawn in-kernel (Philox) every step, like the reference's randn_like",
                   "flops_per_step_reference_oplist": step_flops_ref, "flops_per_step_issued": step_flops_issued},
        "clocks": clocks,
        "e2e": {"value": total_scenes / e2e_s, "unit": "scenes/s", "h2d_bytes_per_step": int(obs_h.numel() * 4 + sp_h.numel() * 4 + rows * 4),
                "d2h_bytes_per_step": int(pred_h.numel() * 4), "ms_per_step": 1e3 * e2e_s / args.steps,
                "api": "sttode_forecast_host (host trajectories in, host predictions out)"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_val, "unit": "scenes/s", "cores": threads, "kind": "port",
                         "sample": f"first {nsmp} scenes of the same workload, one scene per call, PyTorch-CPU oracle port, {threads} threads, {cpu_s:.1f} s"},
        "tc_status": model.tc_status(),
        "ade_fde_delta_vs_reference": parity,
        "extra": extra,
        "quality": {"min_ade": ade, "min_fde": fde, "agents": nrow, "note": "synthetic weights: parity is checked in tests/, not here"},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_train(args):
    """BASELINE.json configs[4] (not the headline metric): STTODE training step — forward + backward through both MHGSA /
    tensor-ODE blocks in libsttode_sm100, data-parallel, one NCCL all-reduce of the flat 6.4 MB gradient buffer per step.
    NBA shape (B = 32 scenes x 11 agents per rank, obs 10 / pred 20, train.py:24), synthetic batches, reference default init."""
    import torch.distributed as dist
    import sttode_b200._lib as L
    from sttode_b200 import parallel, synth
    from sttode_b200.net import STTODENet
    rank, local_rank, world = dist_env()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L.lib()
    B, N, tp, tf = args.train_batch, 11, 10, 20
    torch.manual_seed(1)                                         # train.py:110: identical initial weights on every rank
    model = STTODENet(synth.make_args("nba", tp, tf), torch.device("cpu"))
    model.set_device(dev)
    model.train()
    reducer = parallel.FlatGradReducer(model)
    opt, sched = parallel.make_optimizer(model, capturable=args.train_graph)
    torch.manual_seed(1000 + rank)                               # per-rank dropout / rsample streams
    batches = [{k: v.to(dev) for k, v in synth.nba_batch(100 + rank * 64 + i, B, N, tp, tf).items()} for i in range(8)]
    it = [0]

    def stage():
        model.set_data_nba(batches[it[0] % len(batches)])
        it[0] += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    if args.train_graph:
        graphed = parallel.GraphedTrainStep(model, opt, reducer, batches[0], warmup=max(args.warmup, 3))
        launches_per_step = None

        def one_step():
            it[0] += 1
            return graphed(batches[it[0] % len(batches)])
    else:
        def one_step():
            return parallel.train_step(model, opt, reducer, stage)
    for _ in range(max(args.warmup, 3)):
        out = one_step()
    barrier()
    l0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = one_step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    cpu = None
    if rank == 0 and world == 1 and not args.no_extra:
        # CPU baseline: the oracle port of the reference's step (forward + torch autograd backward), all host threads, one timed step
        from oracle import sttode_oracle as O
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        for k in sd:
            if not k.endswith("pos_encoder.pe"):
                sd[k].requires_grad_(True)
        data = synth.nba_batch(100, B, N, tp, tf)
        g = torch.Generator().manual_seed(5)
        eps_q, eps_p = torch.randn(B * N, 32, generator=g), torch.randn(B * N * 20, 32, generator=g)
        masks = ((torch.rand(B * N, tp, 64, generator=g) >= 0.1).float(), (torch.rand(B * N, tf, 64, generator=g) >= 0.1).float())
        secs = []
        for _ in range(2):
            t0 = time.perf_counter()
            out_cpu = O.training_objective(sd, data["past_traj"].view(B * N, tp, 2), data["future_traj"].view(B * N, tf, 2), B, N, eps_q, eps_p, masks=masks,
                                           tensors=True)
            out_cpu["total"].backward()
            secs.append(time.perf_counter() - t0)
        cpu = {"value": B / secs[-1], "unit": "scenes/s", "cores": threads, "kind": "port",
               "sample": f"one training step (B = {B}) of the PyTorch-CPU oracle port: forward + autograd backward, no optimizer, {secs[-1]:.2f} s"}
    if rank == 0:
        print(json.dumps({"metric": "training scenes/sec (DP, forward+backward+Adam)", "cpu_baseline": cpu, "value": B * world / (ms * 1e-3), "unit": "scenes/s", "n_gpus": world,
                          "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                          "dtype": "f32", "data": "synthetic",
                          "config": {"workload": "STTODE training step (BASELINE.json configs[4]), NBA shape obs 10 / pred 20", "batch_per_gpu": B,
                                     "agents": N, "parallelism": f"dp{world}: one all-reduce of {reducer.nbytes} B per step (pz_layer excluded)"},
                          "gpu_launches": int(L.launch_count() - l0) if not args.train_graph else "one CUDA-graph replay per step (502 library kernels inside)",
                          "cuda_graph": bool(args.train_graph), "last_loss": float(out[0].detach()),
                          "note": "every layer forward+backward in libsttode_sm100 (ODE blocks, dense layers, decoder incl. GRU through time); torch: autograd bookkeeping, losses, Adam"}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes", type=int, default=1536, help="scenes per GPU per step")
    ap.add_argument("--cpu-scenes", type=int, default=1536, help="bounded CPU-baseline sample (scenes; ~10 s of host time at ~150 scenes/s)")
    ap.add_argument("--ref-scenes", type=int, default=96, help="scenes per step of the --impl reference arm")
    ap.add_argument("--no-extra", action="store_true", help="skip the brief single-scene-latency and NBA-shape measurements")
    ap.add_argument("--train", action="store_true", help="time the data-parallel training step (configs[4]) instead of the headline inference metric")
    ap.add_argument("--train-graph", action="store_true", help="replay the whole training step as one CUDA graph")
    ap.add_argument("--train-batch", type=int, default=32, help="scenes per GPU per training step (train.py:24)")
    ap.add_argument("--ffma", action="store_true", help="use the FP32-FFMA decoder tile instead of the tcgen05 one")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.train:
        run_train(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
