"""Wire / disk formats of the reference, so real datasets and checkpoints drop in unchanged.  Host-side only.

* ETH/UCY files (`utils/dataloader.py:27-29,86-144`): csv of four rows (frame, ped, x, y) read transposed (or the classic one-observation-per-line
  text); windows of `obs_len + pred_len` consecutive frames, only pedestrians present in every frame of the window; yields `obs_traj [N,2,T_p]`,
  `pred_traj_gt [N,2,T_f]` (the layout `STTODENet.set_data` takes), plus the per-window CSR the batched `forecast_scenes` takes.
* SDD pickle (`utils/sddloader.py:47-59`): a list of arrays `[n, 20, 2]` (8 observed + 12 future), scaled by `1 / traj_scale`.
* NBA `.npy` (`data/dataloader_nba.py:35-50`): `[*, 15|30, 11, 2]` in feet, scaled by `94 / 28`, split into past / future.
* checkpoints (`train.py:210-213`): `{'model_dict', 'optimizer', 'scheduler', 'epoch', 'model_cfg'}`; `model_cfg` is the argparse Namespace
  `STTODENet(args, device)` is rebuilt from (`test.py:672-678`).
"""
import pickle

import numpy as np
import torch


def read_eth_csv(path):
    """`read_csv` (utils/dataloader.py:27-29): a comma-separated file of FOUR ROWS (frame ids, pedestrian ids, x, y), transposed to [rows, 4]."""
    return np.loadtxt(path, delimiter=",").transpose()


def read_eth_file(path, delim="\t"):
    """The classic `frame ped x y` one-observation-per-line text format (the commented-out `read_file` of utils/dataloader.py:84)."""
    data = []
    with open(path) as f:
        for line in f:
            parts = line.split() if delim in (" ", "space") else line.strip().split("\t" if delim == "tab" else delim)
            if len(parts) >= 4:
                data.append([float(v) for v in parts[:4]])
    return np.asarray(data, dtype=np.float64)


def eth_windows(data, obs_len=8, pred_len=12, skip=1, min_ped=1, traj_scale=1.0):
    """Sliding windows over the frames (utils/dataloader.py:86-144): every window of `obs_len + pred_len` consecutive frame ids, keeping the
    pedestrians observed in ALL of them (coordinates rounded to 4 decimals, divided by traj_scale), and only windows with MORE THAN
    `min_ped` such pedestrians (`num_peds_considered > min_ped`, :134).  Returns a list of
    (obs_traj [N,2,T_p], pred_traj_gt [N,2,T_f], frame id of the first predicted frame (:141))."""
    seq_len = obs_len + pred_len
    frames = np.unique(data[:, 0]).tolist()
    frame_data = [data[data[:, 0] == fr] for fr in frames]
    num_sequences = int(np.ceil((len(frames) - seq_len + 1) / skip))
    out = []
    for idx in range(0, num_sequences * skip + 1, skip):
        if idx + seq_len > len(frames):
            break
        cur = np.concatenate(frame_data[idx:idx + seq_len], axis=0)
        seqs = []
        for ped in np.unique(cur[:, 1]):
            rows = np.around(cur[cur[:, 1] == ped], decimals=4)
            pad_front, pad_end = frames.index(rows[0, 0]) - idx, frames.index(rows[-1, 0]) - idx + 1
            if pad_end - pad_front != seq_len or len(rows) != seq_len:
                continue
            seqs.append(np.transpose(rows[:, 2:4]) / traj_scale)           # [2, seq_len]
        if len(seqs) > min_ped:
            arr = torch.from_numpy(np.stack(seqs)).float()
            out.append((arr[:, :, :obs_len].contiguous(), arr[:, :, obs_len:].contiguous(), int(frames[idx + obs_len])))
    return out


def scenes_to_csr(windows):
    """Concatenate per-window observations into the batched `forecast_scenes` inputs: (obs [sum N,2,T_p], gt [sum N,2,T_f], scene_ptr int32)."""
    obs = torch.cat([w[0] for w in windows])
    gt = torch.cat([w[1] for w in windows])
    ptr = torch.tensor(np.concatenate([[0], np.cumsum([w[0].shape[0] for w in windows])]), dtype=torch.int32)
    return obs, gt, ptr


def read_sdd_pickle(path, obs_len=8, traj_scale=50.0):
    """SDD scenes (utils/sddloader.py:47-59): pickle of a list of `[n, 20, 2]` arrays -> list of (obs [n,2,8], gt [n,2,12])."""
    with open(path, "rb") as f:
        scenes = pickle.load(f)
    out = []
    for sc in scenes:
        t = torch.as_tensor(np.asarray(sc), dtype=torch.float32) / traj_scale      # [n, T, 2]
        t = t.permute(0, 2, 1).contiguous()
        out.append((t[:, :, :obs_len].contiguous(), t[:, :, obs_len:].contiguous()))
    return out


def read_nba_npy(path, obs_len=10, pred_len=20, scale=94.0 / 28.0, limit=None):
    """NBA trajectories (data/dataloader_nba.py:35-61): `[S, T, 11, 2]` in feet -> {'past_traj' [S,11,obs,2], 'future_traj' [S,11,pred,2]} in metres
    (`limit`: the reference keeps the first 32 500 training / 12 500 test sequences, :41-44)."""
    arr = np.load(path)
    if limit is not None:
        arr = arr[:limit]
    traj = torch.from_numpy(arr[:, :obs_len + pred_len] / scale).type(torch.float).permute(0, 2, 1, 3).contiguous()       # [S, 11, T, 2]
    return {"past_traj": traj[:, :, :obs_len].contiguous(), "future_traj": traj[:, :, obs_len:obs_len + pred_len].contiguous()}


def nba_batches(data, batch_size):
    """Reference batching (test.py:618: whole batches attend to each other, so a batch is the unit of work)."""
    S = data["past_traj"].shape[0]
    for i in range(0, S, batch_size):
        yield {k: v[i:i + batch_size] for k, v in data.items()}


def save_checkpoint(path, model, optimizer=None, scheduler=None, epoch=0):
    """train.py:210-213."""
    torch.save({"model_dict": model.state_dict(), "optimizer": None if optimizer is None else optimizer.state_dict(),
                "scheduler": None if scheduler is None else scheduler.state_dict(), "epoch": epoch, "model_cfg": model.args}, path)


def load_checkpoint(path, device, model_cls=None):
    """test.py:672-678: rebuild the model from the pickled `model_cfg` and load `model_dict` with strict=True.  Returns (model, checkpoint)."""
    if model_cls is None:
        from .net import STTODENet as model_cls
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    model = model_cls(ckpt["model_cfg"], device)
    model.load_state_dict(ckpt["model_dict"], strict=True)
    model.set_device(device)
    return model, ckpt
