mkdir -p gpurun_out
CMD="python bench.py --train --train-eager --steps 2 --warmup 3 --no-extra"
timeout 600 $CMD > gpurun_out/plain_train.log 2>&1 || { echo "plain train failed"; tail -5 gpurun_out/plain_train.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_train.csv $CMD > gpurun_out/ncu_list_train.log 2>&1; echo "ncu list train exit $?"
ls -la gpurun_out/r02_launches_train.csv
