# ncu --set full of the named hot-path kernels (MHGSA + ODE block, NBA shape from bench.py's extra section; ODE tail / prologue / row kernel
# from the ETH-shape steps).  Only CSV summaries are kept (the .ncu-rep files would exceed the 64 MiB return limit).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --scenes 128 --cpu-scenes 2"
timeout 600 $CMD > gpurun_out/plain_hot.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none -k regex:"k_geo_attn|k_proj" -c 4 -o /tmp/prof_mhgsa -f $CMD > gpurun_out/ncu_hot1.log 2>&1
echo "ncu1 exit $?"
timeout 900 ncu --set full --clock-control none -k regex:"k_ode_tail|k_prologue|k_dec_row_tc" -s 9 -c 3 -o /tmp/prof_ode -f $CMD > gpurun_out/ncu_hot2.log 2>&1
echo "ncu2 exit $?"
ncu -i /tmp/prof_mhgsa.ncu-rep --page raw --csv > gpurun_out/prof_mhgsa_raw.csv 2>/dev/null
ncu -i /tmp/prof_ode.ncu-rep --page raw --csv > gpurun_out/prof_ode_raw.csv 2>/dev/null
ls -la gpurun_out/*.csv
