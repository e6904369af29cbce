"""bench.py's FLOP accounting — the numerator of `roofline.achieved` and of `flops_per_step_*` — checked on the CPU:
the reference op-list formula equals torch's FlopCounterMode over the oracle's inference to the digit, the table of SURVEY §8d is
reproduced, and the FLOPs the pair kernel issues after exact de-duplication are the documented 1 608 704 per (agent, sample) pair at 8/12 frames."""
import os
import sys

import pytest
import torch
from torch.utils.flop_counter import FlopCounterMode

from oracle import sttode_oracle as O
from sttode_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench      # noqa: E402


@pytest.mark.parametrize("tp,tf,B,N,mflop_per_scene", [(8, 12, 1, 7, 468.99), (8, 12, 1, 60, 4019.93), (10, 20, 32, 11, 807.35), (5, 10, 16, 11, None)])
def test_reference_flop_formula_equals_flop_counter_on_the_oracle(tp, tf, B, N, mflop_per_scene):
    sd = O.make_state_dict(1, tp, tf)
    past = torch.randn(B * N, tp, 2, generator=torch.Generator().manual_seed(0))
    with FlopCounterMode(display=False) as fc:
        O.inference(sd, past, B, N, synth.latent(B * N, seed=9))
    total = fc.get_total_flops()
    assert total == bench.reference_flops_per_row(tp, tf, B) * B * N            # S = B: the attention axis is the scene-batch axis
    if mflop_per_scene is not None:                                              # SURVEY.md §8d table, per scene
        assert abs(total / B / 1e6 - mflop_per_scene) < 0.005


def test_issued_flops_after_deduplication():
    assert bench.issued_flops_dec_pair(8, 12) == 1_608_704
    # pieces: x0-MLP and y0-MLP read the 32-dim latent slice only (K1 = 32: the 224 sample-independent layer-1 columns are in rowpre), y1-MLP the 96 + 32
    # state/latent columns; GRU = conv (2*2*3*32) + input and hidden products per step
    assert bench.flops_mlp(32, 16) + bench.flops_gru(8) + bench.flops_mlp(32, 24) + bench.flops_mlp(128, 24) == 1_608_704
    assert bench.flops_gru(1) == 2 * 6 * 32 + 2 * 32 * 288 + 2 * 96 * 288
    # whole step: issued / reference op list = 0.507 at 8/12 frames, K = 20
    rows, K = 53760, 20
    issued = bench.issued_flops_per_row(8, 1) * rows + bench.issued_flops_dec_pair(8, 12) * rows * K
    ratio = issued / (bench.reference_flops_per_row(8, 12, 1) * rows)
    assert 0.50 < ratio < 0.515, ratio


@pytest.mark.parametrize("name", ["r02_bench_final_eth.json", "r02_bench_final_nba.json", "r02_bench_n1.json", "r02_bench_n2.json", "r02_bench_n4.json", "r02_bench_n8.json"])
def test_committed_bench_lines_are_self_consistent(name):
    """The bench lines committed under profiles/ recompute: value from the shard size and the step time, the roofline from the issued FLOPs and the kernel's
    measured duration, the kernel inside the step, the end-to-end number not above the device-resident one, clean clocks."""
    import json
    with open(os.path.join(ROOT, "profiles", name)) as f:                # multi-rank captures start with NCCL's version banner: the line is the last '{' line
        d = json.loads([ln for ln in f.read().splitlines() if ln.startswith("{")][-1])
    cfg, rf = d["config"], d["roofline"]
    scenes = cfg["scenes_per_gpu_per_step"] * d["n_gpus"]
    assert abs(d["value"] - scenes / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    nba = "NBA" in cfg["workload"]
    tp, tf = (10, 20) if nba else (8, 12)
    per_launch = bench.issued_flops_dec_pair(tp, tf) * cfg["sample_pairs_per_gpu_per_step"]
    assert rf["flops_per_launch_issued"] == per_launch
    assert abs(rf["achieved"] - per_launch / (rf["kernel_ms"] * 1e-3) / 1e12) <= 1e-6 * rf["achieved"]
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and rf["bound"] == "tensor" and rf["unit"] == "TFLOP/s"
    assert 0.5 < rf["kernel_share_of_step"] < 1.0 and rf["kernel_ms"] < d["ms_per_step"]
    assert rf["frac"] < 1 / 3                                          # three tensor-core MACs per algorithmic MAC: the ceiling of the 3xFP16 split
    assert d["e2e"]["value"] <= d["value"] * 1.001 and d["e2e"]["d2h_bytes_per_step"] == cfg["sample_pairs_per_gpu_per_step"] * tf * 2 * 4
    assert d["scaling"] == "weak" and d["unit"] == "scenes/s" and d["vs_baseline"] is None and d["tc_status"] == 0 and d["gpu_launches"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_clock_sampler_summary():
    """bench.ClockSampler.summarise on synthetic nvidia-smi lines: only samples inside the timed region count; a region without any sample falls back to the
    nearest earlier ones and says so; throttle reasons are collected."""
    ok = "0, 1965, 1965, 700.0, 0x0, Not Active, Not Active, Not Active, Not Active"
    cap = "0, 1800, 1965, 990.0, 0x4, Not Active, Not Active, Not Active, Active"
    hot = "0, 1200, 1965, 990.0, 0x40, Not Active, Active, Not Active, Not Active"
    lines = [(0.00, ok), (0.02, hot), (0.04, ok), (0.06, cap), (0.08, "garbage"), (0.10, ok)]
    c = bench.ClockSampler.summarise(lines, 0.03)
    assert c == {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"], "samples": 3, "window": "timed region"}
    c = bench.ClockSampler.summarise(lines, 0.5)                      # nothing inside: the two nearest samples, flagged
    assert c["samples"] == 2 and c["sm_mhz"] == (1800 + 1965) / 2 and c["reasons"] == ["sw_power_cap"] and c["window"].startswith("nearest")
    c = bench.ClockSampler.summarise(lines, 0.0)
    assert c["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"] and c["samples"] == 5
    assert bench.ClockSampler.summarise([], 0.0) == {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "window": "timed region"}
