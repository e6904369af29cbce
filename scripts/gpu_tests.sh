mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/gpu.txt
for grp in "pmath" "mhgsa or oblique" "ode_block or layer or solver" "inference or batched or module_level or latent or host_buffer or min_ade"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$grp" > "gpurun_out/t_${name}.log" 2>&1
  echo "group [$grp] exit $?" >> gpurun_out/summary.txt
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
