"""Disk / wire formats of the reference (sttode_b200/io.py): files written here in the reference's layouts round-trip into the tensors the model
takes; when the reference tree is present its own loaders are the checker."""
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import ref_shim
from sttode_b200 import io as sio
from sttode_b200 import synth


def _eth_table(n_frames=30, peds=5, seed=0):
    rng = np.random.default_rng(seed)
    rows = []
    for p in range(peds):
        start = rng.integers(0, 6)
        stop = n_frames - rng.integers(0, 4)
        pos = rng.uniform(-5, 5, 2)
        for f in range(start, stop):
            pos = pos + rng.normal(0, 0.3, 2)
            rows.append([10.0 * f, float(p + 1), pos[0], pos[1]])
    rows.sort(key=lambda r: (r[0], r[1]))
    return np.asarray(rows)


def test_eth_csv_and_text_roundtrip(tmp_path):
    table = _eth_table()
    csv = tmp_path / "scene.csv"
    np.savetxt(csv, table.transpose(), delimiter=",")              # four rows: frame, ped, x, y
    txt = tmp_path / "scene.txt"
    with open(txt, "w") as f:
        for r in table:
            f.write("\t".join(f"{v:.6f}" for v in r) + "\n")
    a, b = sio.read_eth_csv(csv), sio.read_eth_file(txt)
    assert a.shape == table.shape and np.allclose(a, table) and np.allclose(b, table, atol=1e-6)
    wins = sio.eth_windows(a, 8, 12)
    assert len(wins) > 0
    for obs, gt, frame in wins:
        assert obs.shape[1:] == (2, 8) and gt.shape[1:] == (2, 12) and obs.shape[0] == gt.shape[0] > 1
    obs, gt, ptr = sio.scenes_to_csr(wins)
    assert ptr[0] == 0 and ptr[-1] == obs.shape[0] == gt.shape[0] and len(ptr) == len(wins) + 1
    # every window is a contiguous slice of one pedestrian's track
    o0, g0, fr = wins[0]
    full = torch.cat((o0, g0), dim=2)
    ped_rows = table[table[:, 1] == table[table[:, 0] == fr][0, 1]]
    assert full.shape[2] == 20 and ped_rows.shape[0] >= 20


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")
def test_eth_windows_equal_reference_dataset(tmp_path):
    ref_shim.install()
    from utils.dataloader import TrajectoryDataset
    d = tmp_path / "train"
    d.mkdir()
    table = _eth_table(n_frames=34, peds=7, seed=3)
    np.savetxt(d / "a.csv", table.transpose(), delimiter=",")
    ds = TrajectoryDataset(str(d) + "/", obs_len=8, pred_len=12, skip=1)
    wins = sio.eth_windows(sio.read_eth_csv(d / "a.csv"), 8, 12)
    assert len(wins) == len(ds)
    for i, (obs, gt, frame) in enumerate(wins):
        item = ds[i]
        assert torch.allclose(obs, item[0]) and torch.allclose(gt, item[1])


def test_sdd_pickle_and_nba_npy(tmp_path):
    rng = np.random.default_rng(1)
    scenes = [rng.uniform(0, 1500, (n, 20, 2)) for n in (3, 17, 1)]
    with open(tmp_path / "sdd.pkl", "wb") as f:
        pickle.dump(scenes, f)
    out = sio.read_sdd_pickle(tmp_path / "sdd.pkl", traj_scale=50.0)
    assert [o.shape[0] for o, _ in out] == [3, 17, 1]
    assert out[1][0].shape == (17, 2, 8) and out[1][1].shape == (17, 2, 12)
    assert np.allclose(out[0][0][0, :, 0].numpy(), scenes[0][0, 0] / 50.0, atol=1e-5)
    nba = rng.uniform(0, 94, (9, 30, 11, 2)).astype(np.float64)
    np.save(tmp_path / "nba.npy", nba)
    data = sio.read_nba_npy(tmp_path / "nba.npy", 10, 20)
    assert data["past_traj"].shape == (9, 11, 10, 2) and data["future_traj"].shape == (9, 11, 20, 2)
    assert np.allclose(data["past_traj"][2, 4, 3].numpy(), nba[2, 3, 4] / (94 / 28), atol=1e-5)
    sizes = [b["past_traj"].shape[0] for b in sio.nba_batches(data, 4)]
    assert sizes == [4, 4, 1]


def test_checkpoint_dict_layout(tmp_path):
    """train.py:210-213 keys; the model is rebuilt from `model_cfg` and loads with strict=True (test.py:672-678)."""
    from sttode_b200.net import STTODENet
    torch.manual_seed(2)
    m = STTODENet(synth.make_args("nba", 10, 20), torch.device("cpu"))
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    sched = torch.optim.lr_scheduler.StepLR(opt, 10, 0.5)
    path = tmp_path / "model_0005.p"
    sio.save_checkpoint(path, m, opt, sched, epoch=5)
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ckpt) == {"model_dict", "optimizer", "scheduler", "epoch", "model_cfg"} and ckpt["epoch"] == 5

    class _CpuNet(STTODENet):
        def set_device(self, device):
            self.device = device
    m2, _ = sio.load_checkpoint(path, torch.device("cpu"), model_cls=_CpuNet)
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
