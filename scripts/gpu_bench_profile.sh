# full GPU test suite, bench (both arms), ncu launch list + one full capture of the dominant kernel
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/t_*.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?" >> gpurun_out/summary.txt
CMD="python bench.py --steps 2 --warmup 3 --scenes 384 --cpu-scenes 4 --no-extra"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?" >> gpurun_out/summary.txt
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dec_pair -s 3 -c 2 -o gpurun_out/prof_dec_pair -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 3 gpurun_out/t_all.log; cat gpurun_out/bench.json
