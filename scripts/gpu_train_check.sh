mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py tests/test_gpu_backward.py tests/test_gpu_sampler.py tests/test_gpu_tc.py -m gpu -q > gpurun_out/t_train.log 2>&1; echo "train tests exit $?"
tail -4 gpurun_out/t_train.log | cut -c1-300
timeout 600 python bench.py --train --steps 20 --warmup 3 --no-extra > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_train.json')); print(d['ms_per_step'], d['library_kernels_per_step'])"
STTODE_TC_RECOMPUTE=0 timeout 600 python bench.py --train --steps 20 --warmup 3 --no-extra > gpurun_out/bench_train2.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_train2.json')); print('no tc recompute:', d['ms_per_step'], d['library_kernels_per_step'])"
