"""Multi-rank check of the peer-to-peer exchange: the same 3000-atom water box is advanced
with the NCCL all-gather exchange and with the fused push exchange; trajectories must agree
bit for bit on every rank (both equal the single-GPU trajectory by construction).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 scripts/p2p_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(dev=None, waters=1000, cutoff=9.0, switch=7.5, skin=None, steps=(1, 2, 61), graphs=(False, True)):
    """(The keyword arguments exist for tests/test_mirrors_on_interpreter.py, which runs this function with one
    rank on the host interpreter build; under torchrun on a GPU box the defaults apply.)"""
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ.pop("NCCL_DEBUG", None)
    if dev is None:
        torch.cuda.set_device(local)
        dev = f"cuda:{local}"
        dist.init_process_group("nccl", device_id=torch.device(dev))
    else:
        dist.init_process_group("gloo")
    from torchmd_b200 import Forces, System, maxwell_boltzmann, testsystems
    from torchmd_b200.domain import DecomposedIntegrator

    def make(exchange, use_graph):
        sysd = testsystems.water_box(waters, seed=3)
        par = testsystems.water_parameters(sysd, device=dev)
        n = len(sysd["coords"])
        system = System(n, 1, torch.float32, dev)
        system.set_positions(sysd["coords"])
        system.set_box(sysd["box"])
        torch.manual_seed(5)
        system.set_velocities(maxwell_boltzmann(par.masses, 300.0, 1))
        forces = Forces(par, terms=["lj", "electrostatics", "bonds", "angles"], cutoff=cutoff, rfa=True, switch_dist=switch, skin=skin)
        torch.manual_seed(9)
        integ = DecomposedIntegrator(system, forces, 1.0, dev, gamma=0.1, T=300.0, use_graph=use_graph, exchange=exchange)
        forces.compute(system.pos, system.box, system.forces)
        return system, forces, integ

    ok = True
    for use_graph in graphs:
        sa, fa, ia = make("allgather", use_graph)
        sb, fb, ib = make("p2p", use_graph)
        if ib.exchange != "p2p":
            if rank == 0:
                print("P2P_CHECK FAIL: the peer-to-peer exchange could not be set up (see stderr)", flush=True)
            dist.destroy_process_group()
            sys.exit(2)
        for niter in steps:
            ea = ia.step(niter=niter)
            eb = ib.step(niter=niter)
            same = bool(torch.equal(sa.pos, sb.pos) and torch.equal(sa.vel, sb.vel))
            t = torch.tensor([int(same)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if rank == 0:
                print(f"graph={use_graph} niter={niter}: trajectories identical on all ranks: {bool(t.item())}; "
                      f"Epot {ea[1][0]:.6f} / {eb[1][0]:.6f}", flush=True)
            ok &= bool(t.item())
        fb.stats()
    dist.barrier()
    if rank == 0:
        print("P2P_CHECK", "PASS" if ok else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
