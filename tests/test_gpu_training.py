"""GPU parity of one whole training step (BASELINE.json configs[4]): STTODENet in train() mode — forward + backward through
both MHGSA / tensor-ODE blocks in libsttode_sm100 — against the step of the reference itself (tests/golden/training_step.npz:
loss tuple + gradient of every parameter, all random draws injected) and, at the reference's batch size, against torch
autograd over the float64 CPU oracle.  Tolerance: losses 1e-4 relative; gradients see training_common.assert_grad_close."""
import pytest
import torch

from oracle import sttode_oracle as O
from sttode_b200 import synth
from training_common import CASES, assert_grad_close, case_inputs, check_against_golden, theta_of, training_draws

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def build(dataset, tp, tf, dev, seed=3):
    from sttode_b200.net import STTODENet
    torch.manual_seed(seed)
    m = STTODENet(synth.make_args(dataset, tp, tf), torch.device("cpu"))       # CPU init = the reference's init bit for bit
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m.set_device(dev)
    return m, sd


def stage(m, name, d, data):
    dataset, tp, tf, B, N = CASES[name] if name in CASES else name
    m.past_encoder.dropout_mask_override, m.future_encoder.dropout_mask_override = d["mask_past"], d["mask_future"]
    m.eps_override = (d["eps_q"], d["eps_p"])
    if dataset == "nba":
        m.set_data_nba(data)
    else:
        m.theta_override = theta_of(d).to(m.device)
        m.set_data(None, data[0], data[1], torch.ones(N, tp), torch.ones(N, tf))


@pytest.mark.parametrize("name", ["nba", "eth"])
def test_training_step_vs_reference_golden(golden, dev, name):
    dataset, tp, tf, B, N = CASES[name]
    m, _ = build(dataset, tp, tf, dev)
    m.train()
    d, data = case_inputs(name)
    stage(m, name, d, data)
    m.zero_grad(set_to_none=True)
    total, lp, lr, lk, ld = m.forward()
    total.backward()
    check_against_golden(golden("training_step"), name, (total, lp, lr, lk, ld), [(n, p.grad) for n, p in m.named_parameters()], TOL)


def test_training_step_batch32_vs_oracle_float64(dev):
    """train.py's batch size (B = 32 NBA scenes x 11 agents, 10/20 frames): L = S = 32 in both MHGSA blocks."""
    case = ("nba", 10, 20, 32, 11)
    dataset, tp, tf, B, N = case
    m, sd = build(dataset, tp, tf, dev, seed=5)
    m.train()
    d = training_draws(B * N, tp, tf, seed=77)
    data = synth.nba_batch(4, B, N, tp, tf)
    stage(m, case, d, data)
    m.zero_grad(set_to_none=True)
    total, *parts = m.forward()
    total.backward()
    sd64 = {k: v.double() for k, v in sd.items()}
    names = [k for k in sd64 if not k.endswith("pos_encoder.pe")]
    for k in names:
        sd64[k].requires_grad_(True)
    out = O.training_objective(sd64, data["past_traj"].view(B * N, tp, 2).double(), data["future_traj"].view(B * N, tf, 2).double(), B, N,
                               d["eps_q"].double(), d["eps_p"].double(), masks=(d["mask_past"], d["mask_future"]), tensors=True)
    out["total"].backward()
    assert abs(float(total.detach()) - float(out["total"].detach())) <= TOL * abs(float(out["total"].detach()))
    for got, key in zip(parts, ("pred", "recover", "kl", "diverse")):
        assert abs(got - out[key]) <= TOL * abs(out[key]), key
    for n, p in m.named_parameters():
        ref = sd64[n].grad
        if ref is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
            continue
        assert_grad_close(p.grad.cpu().double(), ref, TOL, n)


def test_train_and_eval_forward_agree(dev):
    """eval() evaluates the objective entirely in the library (fused kernels); train() with an all-ones dropout mask (scaled
    back by 0.9) must give the same numbers through the differentiable path."""
    dataset, tp, tf, B, N = "nba", 10, 20, 6, 11
    m, _ = build(dataset, tp, tf, dev)
    d = training_draws(B * N, tp, tf, seed=9)
    d["mask_past"], d["mask_future"] = torch.full((B * N, tp, 64), 0.9), torch.full((B * N, tf, 64), 0.9)
    data = synth.nba_batch(6, B, N, tp, tf)
    m.train()
    stage(m, (dataset, tp, tf, B, N), d, data)
    a = m.forward()
    m.eval()
    stage(m, (dataset, tp, tf, B, N), d, data)
    b = m.forward()
    assert not b[0].requires_grad and a[0].requires_grad
    for x, y in zip(a, b):
        assert abs(float(x) - float(y)) <= TOL * abs(float(y))


def test_adam_steps_reduce_the_loss(dev):
    """train.py:122-123 optimiser (Adam lr 1e-4 here 1e-3 to see movement in 5 steps) over the autograd nodes: the loss on a fixed batch goes down."""
    dataset, tp, tf, B, N = "nba", 10, 20, 8, 11
    m, _ = build(dataset, tp, tf, dev)
    m.train()
    d = training_draws(B * N, tp, tf, seed=11)
    data = synth.nba_batch(8, B, N, tp, tf)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    hist = []
    for _ in range(5):
        stage(m, (dataset, tp, tf, B, N), d, data)
        total = m.forward()[0]
        opt.zero_grad()
        total.backward()
        opt.step()
        hist.append(float(total))
    assert hist[-1] < hist[0]


def test_cuda_graph_training_step_equals_eager(dev):
    """parallel.GraphedTrainStep replays stage -> objective -> backward -> (all-reduce) -> Adam as one CUDA graph; with every random draw
    injected, one warm-up step plus two replayed steps give the eager loop's losses and land on the same weights (deterministic kernels)."""
    from sttode_b200 import parallel
    dataset, tp, tf, B, N = "nba", 10, 20, 8, 11
    d = training_draws(B * N, tp, tf, seed=21)
    batches = [{k: v.to(dev) for k, v in synth.nba_batch(20 + i, B, N, tp, tf).items()} for i in range(3)]
    results = []
    for graphed in (False, True):
        m, _ = build(dataset, tp, tf, dev, seed=7)
        m.train()
        m.past_encoder.dropout_mask_override, m.future_encoder.dropout_mask_override = d["mask_past"].to(dev), d["mask_future"].to(dev)
        m.eps_override = (d["eps_q"].to(dev), d["eps_p"].to(dev))
        red = parallel.FlatGradReducer(m)
        opt, _ = parallel.make_optimizer(m, lr=1e-3, capturable=graphed)
        losses = []
        if graphed:
            step = parallel.GraphedTrainStep(m, opt, red, batches[0], warmup=1)      # one eager step on batches[0], then capture
            losses.append(step.losses.clone())
            for b in batches[1:]:
                losses.append(step(b).clone())
        else:
            for b in batches:
                out = parallel.train_step(m, opt, red, lambda: m.set_data_nba(b))
                losses.append(torch.tensor([float(out[0].detach()), *out[1:]], device=dev))
        torch.cuda.synchronize()
        results.append((torch.stack(losses).cpu(), {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}))
    (la, wa), (lb, wb) = results
    # capturable Adam evaluates its bias corrections with device tensors: same update up to fp32 rounding
    assert torch.allclose(la, lb, rtol=1e-4, atol=0), (la, lb)
    # Adam normalises every element's update to ~lr, so an element whose gradient is rounding noise may step the other way: compare in the mean
    for k in wa:
        if wa[k].is_floating_point():
            assert float((wa[k] - wb[k]).abs().mean()) < 2e-5, k


def test_fused_dp_adam_matches_torch_adam(dev):
    """parallel.FusedDPAdam (sttode_dp_adam_step: gradient exchange over peer memory fused with the update; world 1 here, so the exchange is
    the identity) against torch.optim.Adam on the same model, same batches, every random draw injected: same losses, same weights after three
    steps, eager and as ONE CUDA graph (stage .. update); the flat parameter / gradient views stay attached; the step count lives on the device."""
    from sttode_b200 import parallel
    dataset, tp, tf, B, N = "nba", 10, 20, 8, 11
    d = training_draws(B * N, tp, tf, seed=31)
    batches = [{k: v.to(dev) for k, v in synth.nba_batch(30 + i, B, N, tp, tf).items()} for i in range(3)]
    results = []
    for mode in ("torch", "fused", "fused-graph"):
        m, _ = build(dataset, tp, tf, dev, seed=7)
        m.train()
        m.past_encoder.dropout_mask_override, m.future_encoder.dropout_mask_override = d["mask_past"].to(dev), d["mask_future"].to(dev)
        m.eps_override = (d["eps_q"].to(dev), d["eps_p"].to(dev))
        losses = []
        if mode == "torch":
            red = parallel.FlatGradReducer(m)
            opt, _ = parallel.make_optimizer(m, lr=1e-3)
        else:
            red, opt = parallel.FusedDPAdam(m, lr=1e-3), None
            assert all(p.data_ptr() >= red.flat_param.data_ptr() and p.data_ptr() < red.flat_param.data_ptr() + red.n * 4 for _, p in red.params)
        if mode == "fused-graph":
            step = parallel.GraphedTrainStep(m, opt, red, batches[0], warmup=1)
            assert step.graph_replays_per_step == 1
            losses.append(step.losses.clone())
            for b in batches[1:]:
                losses.append(step(b).clone())
        else:
            for b in batches:
                out = parallel.train_step(m, opt, red, lambda: m.set_data_nba(b))
                losses.append(torch.tensor([float(out[0].detach()), *out[1:]], device=dev))
        torch.cuda.synchronize()
        if mode != "torch":
            err, steps = red.status()
            assert err == 0 and steps == 3
        results.append((torch.stack(losses).cpu(), {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}))
        if mode != "torch":
            red.close()
    (la, wa), (lb, wb), (lc, wc) = results
    assert torch.allclose(la, lb, rtol=1e-4, atol=0), (la, lb)
    assert torch.allclose(lb, lc, rtol=1e-5, atol=0), (lb, lc)
    for k in wa:
        if wa[k].is_floating_point():
            assert float((wa[k] - wb[k]).abs().mean()) < 2e-5, k
            assert float((wb[k] - wc[k]).abs().mean()) < 2e-6, k
