# weight-stream ablation on the release build, then the y0-MLP issuer timeline (STTODE_TC_TIMELINE=2 build)
mkdir -p gpurun_out
timeout 300 python scripts/tc_nocopy.py > gpurun_out/nocopy.txt 2>&1; cat gpurun_out/nocopy.txt | tail -3
STTODE_TC_TIMELINE=2 python -m sttode_b200.build --force > gpurun_out/tl_build.log 2>&1
STTODE_TC_TIMELINE=2 timeout 300 python scripts/tc_timing2.py > gpurun_out/timeline2.txt 2>&1
cat gpurun_out/timeline2.txt
