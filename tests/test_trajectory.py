"""Appendable trajectory files and the asynchronous frame sink (reference run.py:257-291)."""
import csv
import os

import numpy as np
import pytest
import torch

from torchmd_b200.trajectory import FrameSink, LogWriter, NpyAppender


def test_appended_file_loads_like_the_reference_save(tmp_path):
    rng = np.random.default_rng(0)
    n = 37
    frames = [rng.normal(size=(n, 3)).astype(np.float32) for _ in range(5)]
    w = NpyAppender(str(tmp_path / "traj.npy"), n)
    for k, f in enumerate(frames):
        w.append(f.T)
        if k == 2:
            w.flush()
            part = np.load(tmp_path / "traj.npy")  # loadable mid-run, like the reference's periodic np.save
            assert part.shape == (n, 3, 3) and np.array_equal(part, np.stack(frames[:3], axis=2))
    w.close()
    got = np.load(tmp_path / "traj.npy")
    want = np.stack(frames, axis=2)  # run.py:271-274
    assert got.shape == want.shape == (n, 3, 5) and got.dtype == want.dtype and np.array_equal(got, want)
    with pytest.raises(ValueError):
        NpyAppender(str(tmp_path / "x.npy"), n).append(frames[0])  # (N,3) instead of (3,N)


def test_frame_sink_on_cpu_tensors(tmp_path):
    n, nrep = 23, 3
    sink = FrameSink(str(tmp_path / "out"), ".npy", n, nrep, "cpu", save_every=2)
    pos = torch.zeros(nrep, n, 3)
    want = [[] for _ in range(nrep)]
    for step in range(7):
        pos += torch.randn_like(pos)
        sink.snapshot(pos)  # the tensor is modified right after: the sink must have taken its copy
        for k in range(nrep):
            want[k].append(pos[k].numpy().copy())
    sink.close()
    for k in range(nrep):
        got = np.load(tmp_path / f"out_{k}.npy")
        assert np.array_equal(got, np.stack(want[k], axis=2))


def test_log_writer_layout(tmp_path):
    lw = LogWriter(str(tmp_path), keys=("iter", "ns", "epot", "ekin", "etot", "T"), name="monitor_0.csv")
    lw.write_row({"iter": 10, "ns": 1e-5, "epot": -1.5, "ekin": 2.5, "etot": 1.0, "T": 300.0})
    lw.f.close()
    rows = list(csv.DictReader(open(tmp_path / "monitor_0.csv")))
    assert list(rows[0].keys()) == ["iter", "ns", "epot", "ekin", "etot", "T", "t"]
    assert float(rows[0]["epot"]) == -1.5 and float(rows[0]["t"]) >= 0.0


@pytest.mark.gpu
def test_frame_sink_on_cuda(tmp_path):
    n, nrep = 2000, 2
    dev = "cuda:0"
    sink = FrameSink(str(tmp_path / "out"), ".npy", n, nrep, dev, save_every=3)
    pos = torch.zeros(nrep, n, 3, device=dev)
    want = [[] for _ in range(nrep)]
    for step in range(10):
        pos += torch.randn_like(pos)
        sink.snapshot(pos)
        pos.mul_(1.0)  # work on the compute stream right after the snapshot
        for k in range(nrep):
            want[k].append(pos[k].cpu().numpy().copy())
    sink.close()
    for k in range(nrep):
        assert np.array_equal(np.load(tmp_path / f"out_{k}.npy"), np.stack(want[k], axis=2))
