# final-state evidence: full GPU suite, bench line, ETH launch list + --set full of the decoder kernels, in-kernel timeline (rebuilds the library last)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
CMD="python bench.py --steps 2 --warmup 3 --scenes 384 --nba-batches 8 --no-extra --cpu-seconds 1"
timeout 300 $CMD > gpurun_out/plain_eth.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_eth.csv $CMD > gpurun_out/ncu_list_eth.log 2>&1; echo "ncu list eth exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ode_tail_tc|k_dec_row_tc|k_dec_pair_tc" -s 9 -c 3 -o /tmp/prof_eth -f $CMD > gpurun_out/ncu_full_eth.log 2>&1; echo "ncu full eth exit $?" >> gpurun_out/summary.txt
ncu -i /tmp/prof_eth.ncu-rep --page raw --csv > gpurun_out/r02_prof_eth_raw.csv 2>/dev/null
STTODE_TC_TIMELINE=1 python -m sttode_b200.build --force > gpurun_out/tl_build.log 2>&1
STTODE_TC_TIMELINE=1 timeout 300 python scripts/tc_timing.py > gpurun_out/timeline.txt 2>&1; echo "timeline exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 3 gpurun_out/t_all.log | cut -c1-300; cat gpurun_out/timeline.txt; python - <<PY
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'kernel_ms',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'])
for k,v in d['configs'].items(): print(k, v.get('value'), v.get('ms_per_step'))
PY
