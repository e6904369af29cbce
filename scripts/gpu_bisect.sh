mkdir -p gpurun_out
STTODE_NO_GRAD_SINKS=1 timeout 300 python -m pytest tests/test_gpu_training.py -m gpu -q -x -k "cuda_graph_training_step" > gpurun_out/t_nosinks.log 2>&1; echo "no sinks exit $?"
timeout 300 python -m pytest tests/test_gpu_training.py -m gpu -q -x -k "cuda_graph_training_step" > gpurun_out/t_sinks.log 2>&1; echo "sinks exit $?"
timeout 300 python -m pytest tests/test_gpu_training.py tests/test_gpu_backward.py tests/test_gpu_sampler.py -m gpu -q > gpurun_out/t_train.log 2>&1; echo "train tests exit $?"
tail -5 gpurun_out/t_nosinks.log; tail -5 gpurun_out/t_train.log
