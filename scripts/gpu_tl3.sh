mkdir -p gpurun_out
STTODE_TC_TIMELINE=3 python -m sttode_b200.build --force > gpurun_out/tl_build.log 2>&1
STTODE_TC_TIMELINE=3 timeout 300 python scripts/tc_timing3.py > gpurun_out/timeline3.txt 2>&1; tail -6 gpurun_out/timeline3.txt
