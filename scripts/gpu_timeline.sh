# in-kernel timeline of k_dec_pair_tc: rebuild with the stamps compiled in, run, restore nothing (the box is scratch)
mkdir -p gpurun_out
STTODE_TC_TIMELINE=1 python -m sttode_b200.build --force > gpurun_out/tl_build.log 2>&1
STTODE_TC_TIMELINE=1 timeout 300 python scripts/tc_timing.py > gpurun_out/timeline.txt 2>&1
cat gpurun_out/timeline.txt
