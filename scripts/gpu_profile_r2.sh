# round 2 profiles: per-launch durations of one bench command (ETH headline shape and NBA), then --set full of the encoder tensor-core kernels
# and the decoder kernels.  Each ncu pass only after the same command exited 0 without ncu.  CSV summaries come back in gpurun_out/.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --scenes 384 --nba-batches 8 --no-extra --cpu-seconds 1"
CMDN="python bench.py --config nba --steps 2 --warmup 3 --scenes 384 --nba-batches 8 --no-extra --cpu-seconds 1"
timeout 600 $CMD > gpurun_out/plain_eth.log 2>&1 || { echo "plain eth failed"; tail -5 gpurun_out/plain_eth.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_eth.csv $CMD > gpurun_out/ncu_list_eth.log 2>&1; echo "ncu list eth exit $?"
timeout 600 $CMDN > gpurun_out/plain_nba.log 2>&1 || { echo "plain nba failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_nba.csv $CMDN > gpurun_out/ncu_list_nba.log 2>&1; echo "ncu list nba exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ode_tail_tc|k_dec_row_tc|k_dec_pair_tc" -s 9 -c 3 -o /tmp/prof_eth -f $CMD > gpurun_out/ncu_full_eth.log 2>&1; echo "ncu full eth exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_qkv_tc|k_geo_attn|k_ode_tail_tc" -s 9 -c 3 -o /tmp/prof_nba -f $CMDN > gpurun_out/ncu_full_nba.log 2>&1; echo "ncu full nba exit $?"
ncu -i /tmp/prof_eth.ncu-rep --page raw --csv > gpurun_out/r02_prof_eth_raw.csv 2>/dev/null
ncu -i /tmp/prof_nba.ncu-rep --page raw --csv > gpurun_out/r02_prof_nba_raw.csv 2>/dev/null
ls -la gpurun_out/r02_*
