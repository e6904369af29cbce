# new-kernel check: targeted tests first (short timeout: a protocol bug traps, it must not hang), then everything, then bench
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/t_*.log
timeout 300 python -m pytest tests/test_gpu_parity2.py -m gpu -q -x -k "ode_block_tensor or folded or headline or nba_multi" > gpurun_out/t_new.log 2>&1; echo "new tests exit $?" >> gpurun_out/summary.txt
tail -30 gpurun_out/t_new.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 30 gpurun_out/t_all.log; tail -3 gpurun_out/bench.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'kernel_ms',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'])
for k,v in d['configs'].items(): print(k, v.get('value'), v.get('ms_per_step'))
print({k: v for k, v in d["configs"]["eth_single_scene"].items() if k.startswith("us_")})
PY
