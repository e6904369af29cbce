mkdir -p gpurun_out
CMDN="python bench.py --config nba --steps 2 --warmup 3 --nba-batches 8 --no-extra --cpu-seconds 1"
timeout 600 $CMDN > gpurun_out/plain_nba.log 2>&1 || { echo "plain nba failed"; exit 1; }
timeout 900 ncu --set full --clock-control none -k regex:"k_dec_pair_tc|k_dec_row_tc" -s 4 -c 2 -o /tmp/prof_nba2 -f $CMDN > gpurun_out/ncu_full_nba2.log 2>&1; echo "ncu full nba2 exit $?"
ncu -i /tmp/prof_nba2.ncu-rep --page raw --csv > gpurun_out/r02_prof_nba2_raw.csv 2>/dev/null; ls -la gpurun_out/r02_prof_nba2_raw.csv
