# last check of a round: full GPU suite, then the headline bench line (and the NBA one) without the host-side extras
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 90 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
tail -n 4 gpurun_out/t_all.log | cut -c1-300
timeout 45 python bench.py --steps 10 --warmup 3 --no-extra --cpu-seconds 1 > gpurun_out/bench_final_eth.json 2> gpurun_out/bench_final_eth.err; echo "bench eth exit $?" >> gpurun_out/summary.txt
timeout 45 python bench.py --config nba --steps 5 --warmup 3 --no-extra --cpu-seconds 1 > gpurun_out/bench_final_nba.json 2> gpurun_out/bench_final_nba.err; echo "bench nba exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -2 gpurun_out/bench_final_eth.err | cut -c1-300
python - <<PY
import json
for n in ("eth", "nba"):
    try:
        d = json.load(open(f"gpurun_out/bench_final_{n}.json"))
        print(n, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "metrics_only", d.get("e2e_metrics_only"))
    except Exception as e:
        print(n, "no line:", e)
PY
