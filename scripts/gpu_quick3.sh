# decoder-kernel change check: tensor-core self tests and decoder / inference parity first (short timeout: a protocol bug traps), then everything, then bench
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/t_*.log
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q -x -k "tc or inference or decoder" > gpurun_out/t_new.log 2>&1; echo "decoder tests exit $?" >> gpurun_out/summary.txt
tail -6 gpurun_out/t_new.log | cut -c1-300
if grep -q "exit 0" gpurun_out/summary.txt; then
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
tail -n 6 gpurun_out/t_all.log | cut -c1-300; tail -3 gpurun_out/bench.err | cut -c1-300; python - <<PY
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'kernel_ms',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'])
for k,v in d['configs'].items(): print(k, v.get('value'), v.get('ms_per_step'))
PY
fi
cat gpurun_out/summary.txt
