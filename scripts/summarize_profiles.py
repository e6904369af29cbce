"""Turn the scratch ncu outputs of scripts/gpu_bench_profile.sh (gpurun_out/launches.csv, gpurun_out/prof_dec_pair.ncu-rep) into the
committed summaries under profiles/.  Run here (ncu can read reports without a GPU):  python scripts/summarize_profiles.py r01"""
import csv, io, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.path.join(ROOT, "profiles")

# ---- launch list ----------------------------------------------------------------------------------------------
rows = []
with open(os.path.join(ROOT, "gpurun_out", "launches.csv")) as f:
    text = f.read()
text = text[text.index('"ID"'):]
for r in csv.DictReader(io.StringIO(text)):
    if r["Metric Name"] == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1e6))
agg = collections.OrderedDict()
for name, grid, block, ms in rows:
    short = name.split("(")[0].replace("void ", "")
    a = agg.setdefault(short, [0, grid, block, 0.0])
    a[0] += 1
    a[3] += ms
ours = {k: v for k, v in agg.items() if k.startswith("sttode::")}
step_total = sum(v[3] for k, v in ours.items() if "k_pack_image" not in k)
with open(os.path.join(OUT, f"{tag}_launches_tc.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "grid", "block", "total_ms"])
    for k, v in agg.items():
        w.writerow([k, v[0], v[1], v[2], f"{v[3]:.6f}"])
with open(os.path.join(OUT, f"{tag}_launches_tc.md"), "w") as f:
    f.write(f"# Round 1 — ncu launch list with the tcgen05 decoder tile (gpu__time_duration.sum, --clock-control none)\n\n")
    f.write("Command: `python bench.py --steps 2 --warmup 3 --scenes 384 --cpu-scenes 4 --no-extra` (384 scenes = 13 440 agent rows = 268 800 pairs per step).\n")
    f.write("Per-launch times are cold-cache and serialised: compare SHARES. `k_pack_image` runs once per weight version (set-up, outside the timed region);\n")
    f.write("`k_dec_row` / `k_dec_pair` here are the FP32 redo launches whose CTAs exit at once when no tile left the fp16 range.\n\n")
    f.write("| kernel | launches | grid | block | total ms | share of the per-step sttode kernels |\n|---|---|---|---|---|---|\n")
    for k, v in ours.items():
        share = "(set-up)" if "k_pack_image" in k else f"{v[3] / step_total:.4f}"
        f.write(f"| `{k}` | {v[0]} | {v[1]} | {v[2]} | {v[3]:.3f} | {share} |\n")
    pair = ours.get("sttode::k_dec_pair_tc")
    if pair:
        f.write(f"\nk_dec_pair_tc share of a step under ncu: {pair[3] / step_total:.3f}; bench.py's CUDA-event share in the plain run at 1536 scenes is in the bench JSON (`roofline.kernel_share_of_step`).\n")

# ---- full capture -----------------------------------------------------------------------------------------------
rep = os.path.join(ROOT, "gpurun_out", "prof_dec_pair.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rd = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rd[0], rd[1], rd[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
kcol = hdr.index("Kernel Name")
sel = [r for r in data if "k_dec_pair_tc" in r[kcol]]
with open(os.path.join(OUT, f"{tag}_ncu_full_k_dec_pair_tc.md"), "w") as f:
    f.write("# Round 1 — `ncu --set full --clock-control none` of k_dec_pair_tc (tcgen05 3xFP16 decoder tile)\n\n")
    f.write("Command: `python bench.py --steps 2 --warmup 3 --scenes 384 --cpu-scenes 4 --no-extra`, `-k regex:k_dec_pair -s 3 -c 2` "
            f"({len(sel)} launch(es) of k_dec_pair_tc captured, 268 800 pairs each; 1 CTA/SM).\nRaw report: gpurun_out/prof_dec_pair.ncu-rep (scratch, not committed).\n\n")
    f.write("| metric | " + " | ".join(f"launch {i}" for i in range(len(sel))) + " | unit |\n|---|" + "---|" * (len(sel) + 1) + "\n")
    vals = {}
    for m in want:
        if m in hdr:
            c = hdr.index(m)
            vals[m] = [r[c] for r in sel]
            f.write(f"| {m} | " + " | ".join(vals[m]) + f" | {units[c]} |\n")
    f.write("\n(See DESIGN.md §4.3 for the reading.)\n")
print(open(os.path.join(OUT, f"{tag}_launches_tc.md")).read())
print(open(os.path.join(OUT, f"{tag}_ncu_full_k_dec_pair_tc.md")).read())
