"""minimize_bfgs (reference minimizers.py:8-51) over any object with ``compute(pos, box, forces)``."""
import types

import numpy as np
import pytest
import torch

from torchmd_b200.minimizers import minimize_bfgs


class Springs:
    """E = sum k |x - x0|^2 ; forces written in place like Forces.compute."""

    def __init__(self, x0, k=2.0):
        self.x0, self.k, self.calls = x0, k, 0

    def compute(self, pos, box, forces):
        self.calls += 1
        d = pos - self.x0
        forces.copy_(-2.0 * self.k * d)
        return [float((self.k * d * d).sum())]


def test_minimiser_finds_the_minimum_and_replaces_positions(capsys):
    torch.manual_seed(0)
    x0 = torch.randn(1, 20, 3)
    system = types.SimpleNamespace(pos=x0 + 0.5 * torch.randn(1, 20, 3), box=torch.zeros(1, 3, 3), forces=torch.zeros(1, 20, 3))
    f = Springs(x0)
    start = system.pos
    minimize_bfgs(system, f, fmax=1e-4, steps=200)
    assert system.pos is not start and system.pos.shape == (1, 20, 3) and system.pos.dtype == torch.float32
    assert (system.pos - x0).abs().max() < 1e-3 and f.calls >= 3
    out = capsys.readouterr().out.splitlines()
    assert out[0].split() == ["Iter", "Epot", "fmax"] and out[1].split()[0] == "0"
    minimize_bfgs(system, f, steps=0)  # no-op
    with pytest.raises(RuntimeError):
        minimize_bfgs(types.SimpleNamespace(pos=torch.zeros(2, 4, 3), box=None, forces=None), f)
