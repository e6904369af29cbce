# round 2, first GPU call: full GPU test suite, bench (both arms), smoke
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/t_*.log
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 15 gpurun_out/t_all.log; tail -5 gpurun_out/bench.err; cut -c1-600 gpurun_out/bench.json
