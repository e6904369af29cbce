# usage: bash scripts/gpu_scale.sh N   — the driver's launch line for N ranks, both arms
N=${1:-2}
mkdir -p gpurun_out
P=29517
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/scale_ref_n$N.json 2> gpurun_out/scale_ref_n$N.err; echo "ref N=$N exit $?"
t0=$(date +%s)
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err; echo "bench N=$N exit $? in $(( $(date +%s) - t0 )) s"
tail -5 gpurun_out/scale_n$N.err
python - <<PY
import json
for f in ("gpurun_out/scale_ref_n$N.json", "gpurun_out/scale_n$N.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"])
        for k, v in d.get("configs", {}).items():
            print("   ", k, v.get("value"), v.get("ms_per_step"), v.get("e2e", {}).get("value"), v.get("replica_spread_across_ranks"), v.get("exchange_error_word"))
    except Exception as e:
        print(f, "unreadable", e)
PY
