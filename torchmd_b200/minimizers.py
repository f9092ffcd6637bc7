"""``minimize_bfgs`` -- the reference's L-BFGS-B minimiser (torchmd/minimizers.py:8-51) over
``Forces.compute``.  scipy drives the optimisation on the host, as in the reference; what
changes is the traffic per evaluation: the trial coordinates go into one persistent device
tensor through a pinned staging buffer and the gradient comes back the same way, instead of a
fresh tensor, a ``type_as`` copy and a pageable ``.cpu()`` per call (minimizers.py:19-23).
Same arguments, same printed table, same result: ``system.pos`` is replaced by the minimum.
"""
import numpy as np
import torch


def minimize_bfgs(system, forces, fmax=0.5, steps=1000):
    from scipy.optimize import minimize

    if steps == 0:
        return
    if system.pos.shape[0] != 1:
        raise RuntimeError("System minimization currently doesn't support replicas.")  # minimizers.py:14-17

    pos = system.pos
    cuda = pos.device.type == "cuda"
    trial = torch.empty_like(pos)  # the tensor every evaluation runs on
    h_in = torch.empty(pos.shape, dtype=pos.dtype, pin_memory=cuda)
    h_out = torch.empty(pos.shape, dtype=pos.dtype, pin_memory=cuda)

    def evalfunc(coords, info):
        h_in.copy_(torch.from_numpy(coords.reshape(1, -1, 3)))  # fp64 -> run precision, like type_as
        trial.copy_(h_in, non_blocking=True)
        epot = forces.compute(trial, system.box, system.forces)[0]
        h_out.copy_(system.forces.detach(), non_blocking=True)
        if cuda:
            torch.cuda.current_stream(pos.device).synchronize()
        grad = -h_out.numpy().astype(np.float64)[0]
        if info["Nfeval"] % 1 == 0:  # the reference's progress table
            print("{0:4d}   {1: 3.6f}   {2: 3.6f}".format(info["Nfeval"], epot, np.max(np.linalg.norm(grad, axis=1))))
        info["Nfeval"] += 1
        return epot, grad.reshape(-1)

    print("{0:4s} {1:9s}       {2:9s}".format("Iter", " Epot", " fmax"))
    x0 = pos.detach().cpu().numpy()[0].astype(np.float64).flatten()
    res = minimize(evalfunc, x0, method="L-BFGS-B", jac=True, options={"gtol": fmax, "maxiter": steps},  # ("disp": False in the reference: the default, and newer scipy rejects the key)
                   args=({"Nfeval": 0},))
    system.pos = torch.tensor(res.x.reshape(1, -1, 3), dtype=pos.dtype, device=pos.device, requires_grad=pos.requires_grad)


def minimize_pytorch_bfgs(system, calculator, steps=10, max_iter=20, tolerance_change=1e-9):
    """torchmd/minimizers.py:54-95: ``torch.optim.LBFGS`` over the flattened positions, the gradient from
    ``Epot.sum().backward()`` through ``Forces.compute(..., toNumpy=False)`` (here: the backward of the kernel pass,
    ``forces._EnergyFunction``).  ``steps`` outer calls of ``opt.step`` with up to ``max_iter`` closure evaluations each;
    prints the reference's table, writes the result into ``system.pos`` and returns the energies of every evaluation,
    concatenated along the last axis as the reference does."""
    if steps == 0:
        return
    flat = system.pos.detach().reshape(-1, 3).clone().requires_grad_(True)
    opt = torch.optim.LBFGS([flat], max_iter=max_iter, tolerance_change=tolerance_change)
    history, count = [], [0]

    def closure():
        opt.zero_grad()
        epot = calculator.compute(flat.view(system.nreplicas, system.natoms, 3), system.box, system.forces, toNumpy=False)
        if not torch.is_tensor(epot):  # (forces.py:339 already stacks the replicas; the reference's torch.stack of that tensor raises)
            epot = torch.stack(list(epot))
        history.append(epot.detach().cpu().numpy())
        total = epot.sum()
        total.backward()
        fmax = np.max(np.linalg.norm(flat.grad.detach().cpu().numpy(), axis=1))
        print("{0:4d}   {1: 3.6f}   {2: 3.6f}".format(count[0], float(total.detach()), fmax))
        count[0] += 1
        return total

    print("{0:4s} {1:9s}       {2:9s}".format("Iter", " Epot", " fmax"))
    for _ in range(steps):
        opt.step(closure)
    system.pos[:] = flat.detach().view(1, -1, 3)
    return np.concatenate(history, axis=history[0].ndim - 1) if history[0].ndim else np.asarray(history)


_TAU = 0.618033988749895  # (sqrt(5) - 1) / 2


class _Probe:
    """Evaluations along one search line x0 + a*d for ``minimize_cg`` (minimizers.py:98-105: every evaluation is a
    full ``Forces.compute``; the forces of the latest one are what the caller continues with)."""

    def __init__(self, forces, system, x0, d):
        self.forces, self.system, self.x0, self.d = forces, system, x0, d
        self.pos = self.frc = None

    def energy(self, a, keep=False):
        x = self.x0 + a * self.d
        u = self.forces.compute(x, self.system.box, self.system.forces)[0]
        if keep:
            self.pos, self.frc = x, self.system.forces.detach()[0]
        return u


def _bracket_and_golden_section_search(forces, system, initpos, search_dir, u):
    """minimizers.py:108-273: bracket a minimum of the potential along ``search_dir`` (largest atomic displacement
    1 A, interval shrunk or shifted by golden ratios until it is unimodal), then golden-section search down to an
    interval of 0.01 A.  Returns (positions, forces, potential) of the LAST point evaluated with forces, as the
    reference does (not necessarily the lowest one seen)."""
    reach, tol = 1.0, 1.0e-2
    line = _Probe(forces, system, initpos, search_dir)
    smax = torch.sqrt(torch.max(torch.sum(search_dir**2, dim=1)))
    lo, u_lo = 0.0, u
    hi = reach / smax
    width = hi - lo
    a1, a2 = lo + (1 - _TAU) * width, lo + _TAU * width
    u_hi = line.energy(hi)
    u1 = line.energy(a1)
    u2 = line.energy(a2, keep=True)
    u = u2
    bracketed = (width * smax) <= tol
    while not bracketed:
        if u1 >= u_lo:  # minimum left of a1: keep [lo, a1], both inner points are new
            hi, u_hi = a1, u1
            width = hi - lo
            a1, a2 = lo + (1 - _TAU) * width, lo + _TAU * width
            u1 = line.energy(a1)
            u2 = line.energy(a2, keep=True)
            bracketed = width * smax <= tol
            u = u2
        elif u2 >= u_lo:  # keep [lo, a2]: the old a1 becomes a2
            hi, u_hi = a2, u2
            a2, u2 = a1, u1
            width = hi - lo
            a1 = lo + (1 - _TAU) * width
            u1 = line.energy(a1, keep=True)
            bracketed = width * smax <= tol
            u = u1
        elif u_hi < u1 and u_hi < u2:  # still going down at the far end: slide the interval to [a2, a2 + width]
            lo, u_lo = a2, u2
            a1, u1 = hi, u_hi
            hi = lo + width
            a2 = lo + _TAU * width
            u_hi = line.energy(hi)
            u2 = line.energy(a2, keep=True)
        else:
            bracketed = True
    while width * smax > tol:
        if u1 > u2:
            lo, u_lo = a1, u1
            width = hi - lo
            a1, u1 = a2, u2
            a2 = lo + _TAU * width
            u2 = line.energy(a2, keep=True)
            u = u2
        else:
            hi, u_hi = a2, u2
            width = hi - lo
            a2, u2 = a1, u1
            a1 = lo + (1 - _TAU) * width
            u1 = line.energy(a1, keep=True)
            u = u1
    return line.pos, line.frc, u


def minimize_cg(system, forces, steps=1000, start_step: int = 0, threshold=None, update_system=False):
    """minimizers.py:276-309: Fletcher-Reeves conjugate gradient with the line search above; logs step, energy and the
    largest force component, stops when that drops under ``threshold`` and returns the index of the last step.  The
    reference keeps the minimised coordinates in a local and leaves ``system.pos`` as it was; ``update_system=True``
    (not in the reference) stores them."""
    import logging

    log = logging.getLogger(__name__)
    pos = system.pos.detach()
    u = forces.compute(pos, system.box, system.forces)[0]
    frc = system.forces.detach()[0]
    search_dir = frc.clone()
    fdf = torch.sum(frc**2)
    last = steps - 1
    for step in range(start_step, steps):
        pos, frc, u = _bracket_and_golden_section_search(forces, system, pos.clone(), search_dir, u)
        fdf_prev, fdf = fdf, torch.sum(frc**2)
        search_dir = frc + (fdf / fdf_prev) * search_dir
        energy = forces.compute(pos, system.box, system.forces)[0]
        maxforce = torch.max(torch.abs(system.forces.detach()[0]))
        log.info(f"{step:12d} {energy:14.4f} {maxforce:16.4f}")
        if threshold is not None and maxforce < threshold:
            last = step
            break
    if update_system:
        system.pos[:] = pos.reshape(system.pos.shape)
    return last
