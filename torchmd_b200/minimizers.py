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
