"""sttode_evaluate_host: the loop body of the reference's evaluation (test.py:171-208: H2D, set_data + inference, min-over-K ADE/FDE, AverageMeter
numerators) as one C-ABI call with host buffers, the K samples never leaving HBM.  Checked against (i) sttode_forecast_host with the same seed followed by
the oracle's metric on the host copy of the predictions (what test.py does with numpy), (ii) the device-buffer composition sttode_forecast +
sttode_min_ade_fde."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import sttode_oracle as O
from sttode_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def build_model(dataset, tp, tf, seed, dev):
    from oracle.ref_shim import make_args
    from sttode_b200.net import STTODENet
    m = STTODENet(make_args(dataset, tp, tf), dev)
    m.load_state_dict(O.make_state_dict(seed, tp, tf), strict=True)
    m.set_device(dev)
    return m.eval(), None


def _evaluate(L, mw, traj, gt, layout, gt_layout, sp, n_scenes, rows, Lg, agents, seed):
    sums = (C.c_double * 3)(-1.0, -1.0, -1.0)
    rc = L.lib().sttode_evaluate_host(traj.data_ptr(), gt.data_ptr(), layout, gt_layout, sp.data_ptr() if sp is not None else None, n_scenes, rows, Lg, agents,
                                      seed, 20, C.byref(mw), sums, L.stream_ptr())
    return rc, np.array(list(sums))


def test_evaluate_host_eth_matches_forecast_host_plus_oracle_metric(dev):
    import sttode_b200._lib as L
    m, _ = build_model("eth", 8, 12, 1, dev)
    scenes = synth.eth_scenes(40, cycle=(10, 20, 3, 60, 7))
    counts = [o.shape[0] for o, _ in scenes]
    obs = torch.cat([o for o, _ in scenes]).contiguous()              # [rows, 2, 8]
    gt = torch.cat([g for _, g in scenes]).contiguous()               # [rows, 2, 12] as the ETH loader yields pred_traj_gt
    sp = torch.tensor(np.concatenate([[0], np.cumsum(counts)]), dtype=torch.int32)
    rows = obs.shape[0]
    mw = m._weights()
    pred_h = torch.empty(rows, 20, 12, 2)
    L.check(L.lib().sttode_forecast_host(obs.data_ptr(), 1, sp.data_ptr(), len(counts), rows, 1, 1, 77, 20, C.byref(mw), pred_h.data_ptr(), L.stream_ptr()))
    ade, fde = O.min_ade_fde(pred_h.double(), gt.permute(0, 2, 1).double())
    for pinned in (False, True):
        a, b = (obs.pin_memory(), gt.pin_memory()) if pinned else (obs, gt)
        rc, s = _evaluate(L, mw, a, b, 1, 1, sp, len(counts), rows, 1, 1, 77)
        L.check(rc)
        assert s[2] == rows
        assert abs(s[0] / rows - ade) <= 1e-5 * ade and abs(s[1] / rows - fde) <= 1e-5 * fde, (pinned, s, ade, fde)
    # the same composition on device buffers: identical up to the order of the float64 atomics
    sums = torch.zeros(3, dtype=torch.float64, device=dev)
    m._seed = lambda: 77
    pred_d = m.forecast_scenes(obs.to(dev), sp)
    L.check(L.lib().sttode_min_ade_fde(L.dptr(pred_d), L.dptr(gt.to(dev)), 1, rows, 20, 12, L.dptr(sums), L.stream_ptr()))
    assert np.allclose(sums.cpu().numpy(), s, rtol=1e-12, atol=0)
    a_sum, f_sum, n = m.evaluate_host(obs, gt, sp, seed=77)                   # the Python mirror of the same call
    assert n == rows and np.allclose([a_sum, f_sum], s[:2], rtol=1e-12, atol=0)
    with pytest.raises(L.SttodeError):
        m.evaluate_host(obs.to(dev), gt, sp)
    with pytest.raises(L.SttodeError):
        m.evaluate_host(obs, gt[:, :, :5], sp)
    # a second call overwrites (does not accumulate) and a different seed draws a different latent
    rc, s2 = _evaluate(L, mw, obs, gt, 1, 1, sp, len(counts), rows, 1, 1, 77)
    L.check(rc)
    assert np.allclose(s2, s, rtol=1e-12, atol=0)
    rc, s3 = _evaluate(L, mw, obs, gt, 1, 1, sp, len(counts), rows, 1, 1, 78)
    L.check(rc)
    assert s3[2] == rows and abs(s3[0] - s[0]) > 1e-9 * s[0]
    assert m.tc_status() == 0


def test_evaluate_host_nba_groups(dev):
    """NBA mode: whole reference batches (L = 16 scenes x 11 agents attend to each other), ground truth [rows, T_f, 2]."""
    import sttode_b200._lib as L
    m, _ = build_model("nba", 10, 20, 2, dev)
    Lg, groups = 16, 6
    batches = [synth.nba_batch(i, Lg, 11, 10, 20) for i in range(groups)]
    past = torch.cat([b["past_traj"] for b in batches]).contiguous()
    fut = torch.cat([b["future_traj"] for b in batches]).contiguous()
    rows = groups * Lg * 11
    mw = m._weights()
    pred_h = torch.empty(rows, 20, 20, 2)
    L.check(L.lib().sttode_forecast_host(past.data_ptr(), 0, None, 0, rows, Lg, 11, 5, 20, C.byref(mw), pred_h.data_ptr(), L.stream_ptr()))
    ade, fde = O.min_ade_fde(pred_h.double(), fut.view(rows, 20, 2).double())
    rc, s = _evaluate(L, mw, past, fut, 0, 0, None, 0, rows, Lg, 11, 5)
    L.check(rc)
    assert s[2] == rows and abs(s[0] / rows - ade) <= 1e-5 * ade and abs(s[1] / rows - fde) <= 1e-5 * fde, (s, ade, fde)
    # one reference batch through the Python mirror == the same batch through the C ABI
    b0 = batches[0]
    rc, s0 = _evaluate(L, mw, b0["past_traj"], b0["future_traj"], 0, 0, None, 0, Lg * 11, Lg, 11, 9)
    L.check(rc)
    a_sum, f_sum, n = m.evaluate_host(b0["past_traj"], b0["future_traj"], seed=9)
    assert n == Lg * 11 and np.allclose([a_sum, f_sum], s0[:2], rtol=1e-12, atol=0)
    assert m.tc_status() == 0


def test_evaluate_host_error_paths(dev):
    import sttode_b200._lib as L
    m, _ = build_model("eth", 8, 12, 1, dev)
    (obs, gt), = synth.eth_scenes(1, cycle=(5,))
    obs, gt = obs.contiguous(), gt.contiguous()
    mw = m._weights()
    good = torch.tensor([0, 5], dtype=torch.int32)
    rc, _ = _evaluate(L, mw, obs, gt, 1, 2, good, 1, 5, 1, 1, 1)                       # gt_layout out of range
    assert rc == -1 and b"gt_layout" in L.load().sttode_last_error()
    for bad in ([1, 5], [0, 4], [0, 9, 5], [0, 0, 5]):                                 # malformed CSR: rejected before anything is written
        sp = torch.tensor(bad, dtype=torch.int32)
        rc, s = _evaluate(L, mw, obs, gt, 1, 1, sp, len(bad) - 1, 5, 1, 1, 1)
        assert rc == -1 and (s == -1).all(), bad
    sums = (C.c_double * 3)()
    assert L.lib().sttode_evaluate_host(obs.data_ptr(), None, 1, 1, good.data_ptr(), 1, 5, 1, 1, 1, 20, C.byref(mw), sums, L.stream_ptr()) == -1
    assert L.lib().sttode_evaluate_host(obs.data_ptr(), gt.data_ptr(), 1, 1, good.data_ptr(), 1, 5, 1, 1, 1, 20, C.byref(mw), None, L.stream_ptr()) == -1
    rc, s = _evaluate(L, mw, obs, gt, 1, 1, good, 1, 5, 1, 1, 1)                       # the library is still usable afterwards
    L.check(rc)
    assert s[2] == 5 and np.isfinite(s).all() and s[0] > 0
