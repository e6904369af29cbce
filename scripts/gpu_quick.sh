mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/t_*.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-scenes 16 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 3 gpurun_out/t_all.log; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'kernel_ms',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'], d['clocks'])"
