# one optimisation iteration: decoder-related parity first, then everything, bench, whole-tile timeline (rebuilds the library last)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/t_*.log
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q -x -k "tc or inference or decoder" > gpurun_out/t_new.log 2>&1; echo "decoder tests exit $?" >> gpurun_out/summary.txt
tail -4 gpurun_out/t_new.log | cut -c1-300
if grep -q "exit 0" gpurun_out/summary.txt; then
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
tail -n 4 gpurun_out/t_all.log | cut -c1-300; tail -3 gpurun_out/bench.err | cut -c1-300; python - <<PY
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'kernel_ms',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'])
for k,v in d['configs'].items(): print(k, v.get('value'), v.get('ms_per_step'))
PY
fi
TLMODE=${TLMODE:-3}
STTODE_TC_TIMELINE=$TLMODE python -m sttode_b200.build --force > gpurun_out/tl_build.log 2>&1
if [ "$TLMODE" = "1" ]; then S=scripts/tc_timing.py; else S=scripts/tc_timing$TLMODE.py; fi
STTODE_TC_TIMELINE=$TLMODE timeout 300 python $S > gpurun_out/timeline$TLMODE.txt 2>&1; cat gpurun_out/timeline$TLMODE.txt | tail -12
cat gpurun_out/summary.txt
