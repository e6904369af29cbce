# A/B on one box: pair-kernel phase order / ring slots (STTODE_DEC_Y_FIRST), alternating runs
mkdir -p gpurun_out
for r in 1 2; do for yf in 0 1; do
  STTODE_DEC_Y_FIRST=$yf timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --cpu-seconds 1 > gpurun_out/ab_$yf.json 2> gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_$yf.json'))
print('y_first=$yf', 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value']))
PY
done; done
for yf in 0 1; do
  STTODE_DEC_Y_FIRST=$yf timeout 300 python bench.py --config nba --steps 10 --warmup 3 --no-extra --cpu-seconds 1 > gpurun_out/abn_$yf.json 2> gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open('gpurun_out/abn_$yf.json'))
print('nba y_first=$yf', 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4))
PY
done
