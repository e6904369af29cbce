# launch list of the ETH step (384 scenes) on the release build, then the in-kernel timeline (mode from $TLMODE, default 1)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --scenes 384 --nba-batches 8 --no-extra --cpu-seconds 1"
timeout 300 $CMD > gpurun_out/plain_eth.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_eth.csv $CMD > gpurun_out/ncu_list_eth.log 2>&1; echo "ncu list eth exit $?"
python scripts/summarize_r2.py eth-launches > /dev/null 2>&1; python - <<PY
import csv, io, collections
text = open('gpurun_out/r02_launches_eth.csv').read(); text = text[text.index('"ID"'):]
agg = collections.OrderedDict()
for r in csv.DictReader(io.StringIO(text)):
    if r["Metric Name"] != "gpu__time_duration.sum": continue
    k = r["Kernel Name"].split("(")[0].replace("void ", "").replace("sttode::", "")
    a = agg.setdefault(k, [0, 0.0, r["Grid Size"]]); a[0] += 1; a[1] += float(r["Metric Value"]) / 1e6
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]): print(f"{k:40s} n={v[0]:3d} grid={v[2]:16s} total {v[1]:8.3f} ms  per launch {v[1]/v[0]:.4f}")
PY
TLMODE=${TLMODE:-1}
STTODE_TC_TIMELINE=$TLMODE python -m sttode_b200.build --force > gpurun_out/tl_build.log 2>&1
if [ "$TLMODE" = "1" ]; then S=scripts/tc_timing.py; else S=scripts/tc_timing$TLMODE.py; fi
STTODE_TC_TIMELINE=$TLMODE timeout 300 python $S > gpurun_out/timeline$TLMODE.txt 2>&1; cat gpurun_out/timeline$TLMODE.txt | tail -12
