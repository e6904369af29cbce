"""gpurun_out/r02_launches_{eth,nba}.csv + r02_prof_{eth,nba}_raw.csv (scripts/gpu_profile_r2.sh) -> profiles/r02_launches_*.md, profiles/r02_ncu_full_*.md"""
import collections, csv, io, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
CMDS = {"eth": "python bench.py --steps 2 --warmup 3 --scenes 384 --nba-batches 8 --no-extra --cpu-seconds 1  (384 zara/univ scenes = 13 440 agent rows = 268 800 pairs per step)",
        "nba": "python bench.py --config nba --steps 2 --warmup 3 --nba-batches 8 --no-extra --cpu-seconds 1  (8 reference batches of 128 scenes = 11 264 agent rows = 225 280 pairs per step)"}


def launches(tag):
    text = open(os.path.join(ROOT, "gpurun_out", f"r02_launches_{tag}.csv")).read()
    text = text[text.index('"ID"'):]
    rows = [(r["Kernel Name"], r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1e6) for r in csv.DictReader(io.StringIO(text))
            if r["Metric Name"] == "gpu__time_duration.sum"]
    # the device-resident leg's timed steps: the LAST two groups that contain k_dec_pair_tc before the first memcpy-heavy e2e step are hard to cut out of a flat
    # list, so aggregate over the whole run and report shares among the library's own kernels
    agg = collections.OrderedDict()
    for name, grid, block, ms in rows:
        short = name.split("(")[0].replace("void ", "")
        a = agg.setdefault(short, [0, grid, block, 0.0])
        a[0] += 1
        a[3] += ms
    ours = {k: v for k, v in agg.items() if k.startswith("sttode::")}
    setup = ("k_pack_all", "k_pack_ode", "k_fold_prologue")
    tot = sum(v[3] for k, v in ours.items() if not any(s in k for s in setup))
    with open(os.path.join(OUT, f"r02_launches_{tag}.md"), "w") as f:
        f.write(f"# Round 2 — ncu launch list, {tag} config (gpu__time_duration.sum, --clock-control none)\n\nCommand: `{CMDS[tag]}`\n\n")
        f.write("Whole run (warm-up + timed steps of the device-resident leg + the end-to-end leg + the parity call), aggregated per kernel.  Per-launch times are\n"
                "cold-cache and serialised: compare SHARES.  `k_pack_all` / `k_pack_ode` / `k_fold_prologue` run once per weight version (set-up).  `k_dec_row`, `k_dec_pair`,\n"
                "`k_ode_tail`, `k_proj` are the FP32 redo launches whose CTAs exit at once when no tile left the fp16 range.\n\n")
        f.write("| kernel | launches | grid | block | total ms | ms / launch | share of the library's per-step kernels |\n|---|---|---|---|---|---|---|\n")
        for k, v in sorted(ours.items(), key=lambda kv: -kv[1][3]):
            share = "(set-up)" if any(s in k for s in setup) else f"{v[3] / tot:.4f}"
            f.write(f"| `{k.replace('sttode::', '')}` | {v[0]} | {v[1]} | {v[2]} | {v[3]:.3f} | {v[3] / v[0]:.4f} | {share} |\n")
        other = sum(v[3] for k, v in agg.items() if not k.startswith("sttode::"))
        f.write(f"\nNon-library kernels in the same run (torch fills / copies of the harness): {other:.3f} ms in {sum(v[0] for k, v in agg.items() if not k.startswith('sttode::'))} launches.\n")


KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]


def full(tag):
    path = os.path.join(ROOT, "gpurun_out", f"r02_prof_{tag}_raw.csv")
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(OUT, f"r02_ncu_full_{tag}.md"), "w") as f:
        f.write(f"# Round 2 — `ncu --set full --clock-control none`, {tag} config\n\nCommand: `{CMDS[tag]}`; one capture per kernel (cold cache, serialised).\n\n")
        for r in data:
            name = r[col["Kernel Name"]].split("(")[0]
            f.write(f"## `{name}`  grid {r[col['Grid Size']]} block {r[col['Block Size']]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in col:
                    f.write(f"| {k} | {r[col[k]]} | {units[col[k]]} |\n")
            f.write("\n")


if __name__ == "__main__":
    for t in ("eth", "nba"):
        launches(t)
        full(t)
    print("ok")
