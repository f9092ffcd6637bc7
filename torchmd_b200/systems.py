"""State container with the interface of ``torchmd.systems.System``
(reference ``torchmd/systems.py:5-98``): ``pos, vel, forces`` of shape (R, N, 3),
``box`` (R, 3, 3) with only the diagonal meaningful, ``masses`` (N, 1).

The setters accept the same layouts and raise the same ``RuntimeError`` cases as
the reference (positions (N,3)/(N,3,1)/(N,3,R); box (3,)/(3,1)/(3,R); replica
broadcast from the first entry).  Tensors are kept contiguous because the CUDA
kernels take raw pointers.
"""
import numpy as np
import torch

_FIELDS = ("box", "pos", "vel", "forces", "masses")


class System:
    def __init__(self, natoms, nreplicas, precision, device):
        shapes = {
            "box": (nreplicas, 3, 3),
            "pos": (nreplicas, natoms, 3),
            "vel": (nreplicas, natoms, 3),
            "forces": (nreplicas, natoms, 3),
            "masses": (natoms, 1),
        }
        for name in _FIELDS:
            setattr(self, name, torch.zeros(shapes[name], dtype=precision, device=device))

    @property
    def natoms(self):
        return self.pos.shape[1]

    @property
    def nreplicas(self):
        return self.pos.shape[0]

    def to_(self, device):
        for name in _FIELDS:
            setattr(self, name, getattr(self, name).to(device))

    def precision_(self, precision):
        for name in _FIELDS:
            setattr(self, name, getattr(self, name).type(precision))

    def _like(self, value, target):
        if isinstance(value, np.ndarray):
            return torch.tensor(value, dtype=target.dtype, device=target.device)
        return value.clone().detach().type(target.dtype).to(target.device)

    def set_positions(self, pos):
        if pos.shape[1] != 3:
            raise RuntimeError(
                f"Positions shape must be (natoms, 3, 1) or (natoms, 3, nreplicas) but were given {pos.shape} instead"
            )
        p = self._like(pos, self.pos)
        if p.ndim == 2:
            p = p[:, :, None]
        p = p.permute(2, 0, 1)  # (frames, N, 3)
        if self.nreplicas > 1 and p.shape[0] != self.nreplicas:
            p = p[0:1].expand(self.nreplicas, -1, -1)
        self.pos[:] = p

    def set_velocities(self, vel):
        if tuple(vel.shape) != (self.nreplicas, self.natoms, 3):
            raise RuntimeError("Velocities shape must be (nreplicas, natoms, 3)")
        self.vel[:] = self._like(vel, self.vel)

    def set_box(self, box):
        box = np.asarray(box.detach().cpu() if torch.is_tensor(box) else box)
        if box.ndim == 1:
            if len(box) != 3:
                raise RuntimeError("Box must have at least 3 elements")
            box = box[:, None]
        if box.shape[0] != 3:
            raise RuntimeError("Box shape must be (3, 1) or (3, nreplicas)")
        rows = box.T  # (frames, 3)
        if self.nreplicas > 1 and rows.shape[0] != self.nreplicas:
            rows = np.repeat(rows[0:1], self.nreplicas, axis=0)
        diag = torch.tensor(np.ascontiguousarray(rows), dtype=self.box.dtype, device=self.box.device)
        for r in range(rows.shape[0]):
            self.box[r, [0, 1, 2], [0, 1, 2]] = diag[r]

    def set_forces(self, forces):
        if tuple(forces.shape) != (self.nreplicas, self.natoms, 3):
            raise RuntimeError("Forces shape must be (nreplicas, natoms, 3)")
        self.forces[:] = self._like(forces, self.forces)

    def set_masses(self, masses):
        if tuple(masses.shape) != (self.natoms,):
            raise RuntimeError("Masses shape must be (natoms,)")
        self.masses[:, 0] = self._like(masses, self.masses)
