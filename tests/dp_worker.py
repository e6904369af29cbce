"""Worker of tests/test_gpu_multi.py: launched by torch.distributed.run with one rank per GPU.  Trains the same model three data-parallel steps
twice — gradient exchange fused with Adam over NVLink peer memory (parallel.FusedDPAdam, eager and as one CUDA graph) and NCCL all-reduce +
torch.optim.Adam (parallel.FlatGradReducer) — with every random draw injected, and reports how far the replicas / the two exchanges are apart."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(out_path):
    from sttode_b200 import parallel, synth
    from sttode_b200.net import STTODENet
    from training_common import training_draws
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    dataset, tp, tf, B, N = "nba", 10, 20, 8, 11
    d = training_draws(B * N, tp, tf, seed=60 + rank)                    # every rank its own draws and its own batches
    batches = [{k: v.to(dev) for k, v in synth.nba_batch(200 + 10 * rank + i, B, N, tp, tf).items()} for i in range(3)]
    weights, losses = {}, {}
    for mode in ("nccl", "fused", "fused-graph"):
        torch.manual_seed(7)                                             # identical initial replicas
        m = STTODENet(synth.make_args(dataset, tp, tf), torch.device("cpu"))
        m.set_device(dev)
        m.train()
        m.past_encoder.dropout_mask_override, m.future_encoder.dropout_mask_override = d["mask_past"].to(dev), d["mask_future"].to(dev)
        m.eps_override = (d["eps_q"].to(dev), d["eps_p"].to(dev))
        if mode == "nccl":
            red = parallel.FlatGradReducer(m)
            opt, _ = parallel.make_optimizer(m, lr=1e-3)
        else:
            red, opt = parallel.FusedDPAdam(m, lr=1e-3), None
        ls = []
        if mode == "fused-graph":
            step = parallel.GraphedTrainStep(m, opt, red, batches[0], warmup=1)
            ls.append(float(step.losses[0]))
            for b in batches[1:]:
                ls.append(float(step(b)[0]))
        else:
            for b in batches:
                out = parallel.train_step(m, opt, red, lambda: m.set_data_nba(b))
                ls.append(float(out[0].detach()))
        torch.cuda.synchronize()
        flat = torch.cat([p.detach().reshape(-1) for n, p in m.named_parameters() if "pz_layer" not in n]).clone()
        hi, lo = flat.clone(), flat.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        weights[mode] = flat
        losses[mode] = ls
        res = {"spread": float((hi - lo).abs().max())}
        if mode != "nccl":
            err, steps = red.status()
            res.update(err=err, steps=steps)
            dist.barrier()
            red.close()
        weights[mode + "_res"] = res
    result = {"world": world, "losses": losses,
              "replica_spread": {k: weights[k + "_res"] for k in ("nccl", "fused", "fused-graph")},
              "fused_vs_nccl_mean_abs": float((weights["fused"] - weights["nccl"]).abs().mean()),
              "graph_vs_eager_mean_abs": float((weights["fused-graph"] - weights["fused"]).abs().mean()),
              "weight_mean_abs": float(weights["nccl"].abs().mean())}
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(result, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
