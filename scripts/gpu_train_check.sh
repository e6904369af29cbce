mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py tests/test_gpu_backward.py tests/test_gpu_sampler.py -m gpu -q > gpurun_out/t_train.log 2>&1; echo "train tests exit $?"
timeout 300 python -m pytest tests/test_gpu_parity2.py -m gpu -q -k pmath > gpurun_out/t_pmath.log 2>&1; echo "pmath tests exit $?"
tail -8 gpurun_out/t_train.log | cut -c1-300; tail -8 gpurun_out/t_pmath.log | cut -c1-300
timeout 600 python bench.py --train --steps 20 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train exit $?"; tail -3 gpurun_out/bench_train.err | cut -c1-300; cut -c1-900 gpurun_out/bench_train.json
timeout 600 python bench.py --train --train-exchange nccl --steps 20 --warmup 3 --no-extra > gpurun_out/bench_train_nccl.json 2> gpurun_out/bench_train_nccl.err; echo "bench train nccl exit $?"; cut -c1-400 gpurun_out/bench_train_nccl.json
