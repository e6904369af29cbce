mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for grp in "tc_selftest" "inference_eth or module_level" "inference_nba or batched or latent or host_buffer" "pmath or mhgsa or oblique or ode_block or layer or solver or min_ade"; do
  name=$(echo "$grp" | tr ' ' '_' | cut -c1-40)
  f=tests
  timeout 900 python -m pytest tests -m gpu -q -x -k "$grp" > "gpurun_out/t_${name}.log" 2>&1
  echo "group [$grp] exit $?" >> gpurun_out/summary.txt
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/t_*.log
