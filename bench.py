"""bench.py -- MD steps/s on the synthetic 100k-atom TIP3P water box (BASELINE.json).

    python bench.py --gpus 1 --steps K --warmup W          # our arm
    python bench.py --impl reference --steps K --warmup W  # reference CPU arm (oracle port)

One "step" is one MD step (velocity Verlet + Langevin, LJ + reaction-field
electrostatics + bonds + angles, cutoff 9 / switch 7.5, dt 1 fs) of 33,333 flexible
TIP3P waters (99,999 atoms, fp32).  Rank 0 prints ONE JSON line:

  value      steps/s with the state resident in HBM, K steps enqueued back to back
             through Integrator.step (CUDA events on the launching stream, max over ranks)
  e2e        steps/s through the C-ABI host entry tmd_md_steps_host: pinned HOST
             positions+velocities copied in, one step, positions+velocities+energies
             copied back, every step
  roofline   non-bonded pair kernel: algorithmic bytes (32 N + 4 P_rc, SURVEY.md
             section 8d) / its mean duration (CUDA events around each launch inside the
             timed region) against the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the oracle port of the reference's all-pairs PyTorch-CPU path timed on
             this box's host cores on a bounded sample (smaller box, O(N^2) scaled)
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_WATERS = 33333
TERMS = ["lj", "electrostatics", "bonds", "angles"]
CFG = dict(cutoff=9.0, rfa=True, switch_dist=7.5)
TIMESTEP_FS = 1.0
GAMMA_PS = 0.1
TEMPERATURE = 300.0
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md


# ------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's CPU path
# ------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, budget_s=150.0):
    """Time the reference algorithm (all-pairs + mask, torch CPU ops, all host threads)
    on a bounded sample: the same generator/settings at a box size whose (warmup+steps)
    steps fit the budget; scaled to the 99,999-atom metric with the O(N^2) law the
    all-pairs evaluation follows (100k atoms are infeasible for
    the reference -- 80 GB pair table)."""
    import torch

    from oracle import refmd
    from torchmd_b200 import testsystems

    ncores = os.cpu_count() or 1
    # ATen's elementwise kernels stop scaling (and degrade) on very wide hosts: probe a few thread
    # counts on a 3,000-atom box and give the reference the fastest one
    best = (float("inf"), ncores)
    try:
        probe = testsystems.water_box(1000, seed=0)
        ppar = testsystems.water_parameters(probe, precision=torch.float32)
        pof = refmd.OracleForces(ppar, TERMS, **CFG)
        ppos = torch.tensor(probe["coords"])[None].clone()
        pbox = torch.zeros(1, 3, 3)
        for k in range(3):
            pbox[0, k, k] = float(probe["box"][k])
        pF = torch.zeros_like(ppos)
        for nt in sorted({min(ncores, c) for c in (8, 16, 32, 64, ncores)}):
            torch.set_num_threads(nt)
            pof.compute(ppos, pbox, pF)
            t0 = time.perf_counter()
            pof.compute(ppos, pbox, pF)
            dt = time.perf_counter() - t0
            if dt < best[0]:
                best = (dt, nt)
    except Exception:
        pass
    nthreads = best[1]
    torch.set_num_threads(nthreads)
    est = {3333: 1.6, 1000: 0.22, 333: 0.05}  # s/step measured on 8 cores
    scale = 8.0 / max(1, min(ncores, 32))
    nw = 333
    for cand in (3333, 1000, 333):
        if (steps + warmup) * est[cand] * max(scale, 0.25) <= budget_s:
            nw = cand
            break
    sysd = testsystems.water_box(nw, seed=0)
    par = testsystems.water_parameters(sysd, precision=torch.float32)
    n = len(sysd["coords"])
    t0 = time.perf_counter()
    of = refmd.OracleForces(par, TERMS, **CFG)
    t_init = time.perf_counter() - t0
    pos = torch.tensor(sysd["coords"])[None].clone()
    box = torch.zeros(1, 3, 3)
    for k in range(3):
        box[0, k, k] = float(sysd["box"][k])
    torch.manual_seed(1)
    vel = refmd.maxwell_boltzmann(par.masses, TEMPERATURE, 1)
    F = torch.zeros_like(pos)
    fn = lambda p, b, f: [sum(e.values()) for e in of.compute(p, b, f)]  # noqa: E731
    fn(pos, box, F)
    integ = refmd.OracleIntegrator(pos, vel, box, F, par.masses, fn, TIMESTEP_FS, gamma_ps=GAMMA_PS, T=TEMPERATURE)
    integ.step(warmup)
    t0 = time.perf_counter()
    integ.step(steps)
    dt = time.perf_counter() - t0
    measured = steps / dt
    target_n = 3 * N_WATERS
    value = measured * (n / target_n) ** 2
    return {
        "value": value,
        "unit": "steps/s",
        "cores": nthreads,
        "kind": "port",
        "sample": (
            f"oracle/refmd.py (torch-CPU restatement of the reference; {nthreads} torch threads, the fastest of a probe "
            f"over 8..{ncores} on this {ncores}-core host) on a {n}-atom water box, "
            f"{steps} steps after {warmup} warm-up: measured {measured:.4g} steps/s ({dt / steps:.3f} s/step, "
            f"pair-table init {t_init:.1f} s); value = measured x ({n}/{target_n})^2 (all-pairs O(N^2)); "
            f"99,999 atoms are infeasible for the reference (O(N^2) memory)"
        ),
        "measured_steps_per_s": measured,
        "measured_natoms": n,
        "ms_per_step_measured": 1e3 * dt / steps,
    }


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    base = cpu_reference_run(steps, max(0, args.warmup))
    line = {
        "impl": "reference",
        "metric": "MD steps/sec (100k-atom water, fp32)",
        "value": base["value"],
        "unit": "steps/s",
        "n_gpus": args.gpus,
        "steps": steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 / base["value"],
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(ngpus):
    return {
        "workload": "synthetic TIP3P water box, 33333 waters = 99999 atoms, L=99.93 A, LJ(switch 7.5)+RF electrostatics cutoff 9 A, "
        "flexible bonds+angles, Langevin 300 K gamma 0.1/ps, dt 1 fs, 1 replica",
        "natoms": 3 * N_WATERS,
        "pair_kernel": {"1": "fixed-point separations (TMD_B200_FX=1)", "2": "fixed-point separations + packed fp32x2 arithmetic (TMD_B200_FX=2)"}.get(
            os.environ.get("TMD_B200_FX", "")[:1], "float separations (default)"),
        "parallelism": "single GPU" if ngpus == 1 else (f"spatial slabs over {ngpus} GPUs, " + ("positions pushed to all ranks over NVLink peer memory by the integration kernel"
                                                   if os.environ.get("TMD_B200_EXCHANGE", "").lower() == "p2p" else "position all-gather")),
        "l2": "no flush between steps: consecutive MD steps are data-dependent; the neighbour list streamed by the "
        "pair kernel (>150 MB) exceeds the 126 MB L2",
    }


# ------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index=0):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=self.tmp, stderr=subprocess.DEVNULL,
            )
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.tmp.read().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.tmp.name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------
DEVICE_OVERRIDE = None  # tests/test_mirrors_on_interpreter.py dry-runs gpu_arm on the host interpreter build with "cpu"
MIN_DECOMPOSED_WARMUP = 5000  # N > 1: steps before the timed region (>= 0.4 s under load for the clock sampler)


def gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from torchmd_b200 import Forces, Integrator, System, _lib, maxwell_boltzmann, testsystems

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.pop("NCCL_DEBUG", None)  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = DEVICE_OVERRIDE or f"cuda:{local}"
    if DEVICE_OVERRIDE is None:
        torch.cuda.set_device(local)
    if world > 1:
        from torchmd_b200 import domain  # spatial decomposition driver

        try:
            return domain.bench_decomposed(args, world, rank, local, workload_config(world))
        finally:
            dist.destroy_process_group()

    sysd = testsystems.water_box(N_WATERS, seed=0)
    n = len(sysd["coords"])
    par = testsystems.water_parameters(sysd, device=dev)
    system = System(n, 1, torch.float32, dev)
    system.set_positions(sysd["coords"])
    system.set_box(sysd["box"])
    torch.manual_seed(1)
    system.set_velocities(maxwell_boltzmann(par.masses, TEMPERATURE, 1))
    forces = Forces(par, terms=TERMS, **CFG)
    forces.compute(system.pos, system.box, system.forces)

    # relax the lattice start into a liquid: strong coupling, then the production thermostat
    eq = Integrator(system, forces, TIMESTEP_FS, dev, gamma=10.0, T=TEMPERATURE)
    for _ in range(args.equil // 100):
        eq.step(niter=100)
    integ = Integrator(system, forces, TIMESTEP_FS, dev, gamma=GAMMA_PS, T=TEMPERATURE)
    sampler = ClockSampler(local)  # runs through the warm-up too (same load), so short runs still get samples
    t_w = time.perf_counter()
    done = 0
    while done < max(3, args.warmup) or time.perf_counter() - t_w < 0.6:  # >= 0.6 s under load for the sampler
        ekin, pot, T = integ.step(niter=50)
        done += 50

    L = _lib.lib()
    stream = torch.cuda.current_stream().cuda_stream
    st0 = forces.stats()
    torch.cuda.synchronize()
    _lib.check(L.tmd_profile_begin(forces._ctx, args.steps))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    ekin, pot, T = integ.step(niter=args.steps)
    ev1.record()
    torch.cuda.synchronize()
    import ctypes as C

    pair_ms, pair_n = C.c_double(), C.c_int()
    _lib.check(L.tmd_profile_end(forces._ctx, C.byref(pair_ms), C.byref(pair_n), stream))
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    st1 = forces.stats()
    ms_per_step = ms_total / args.steps
    value = 1e3 / ms_per_step

    # exact number of in-cutoff pairs of the final configuration (reference predicate)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    dummy = torch.zeros(2, dtype=torch.int32, device=dev)
    scratch = torch.empty_like(system.pos)
    forces.compute(system.pos, system.box, scratch)
    _lib.check(L.tmd_export_pairs(forces._ctx, system.pos.data_ptr(), 0, dummy.data_ptr(), 0, count.data_ptr(), stream))
    p_rc = int(count.item())

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", FALLBACK_HBM_GBS))
    pair_avg_ms = pair_ms.value / max(1, pair_n.value)
    alg_bytes = 32.0 * n + 4.0 * p_rc
    achieved = alg_bytes / (pair_avg_ms * 1e-3) / 1e9 if pair_avg_ms > 0 else 0.0
    roofline = {
        "kernel": "k_pair<false,true> (non-bonded pair kernel)",
        "bound": "hbm",
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak,
        "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
        "algorithmic_bytes": alg_bytes,
        "pairs_in_cutoff": p_rc,
        "avg_kernel_ms": pair_avg_ms,
        "launches_sampled": pair_n.value,
        "share_of_step": pair_avg_ms / ms_per_step,
        "traffic": None,
    }
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "pair_kernel_traffic.json")))
        roofline["traffic"] = prof.get("dram_bytes_per_launch")
        roofline["traffic_source"] = prof.get("source")
    except Exception:
        pass

    # end to end: host-resident state through the C-ABI host entry, one step per call
    e2e_steps = min(args.steps, args.e2e_steps)
    hpos = torch.empty(system.pos.shape, dtype=torch.float32, pin_memory=True)
    hvel = torch.empty(system.vel.shape, dtype=torch.float32, pin_memory=True)
    hpos.copy_(system.pos)
    hvel.copy_(system.vel)
    hene = np.zeros((1, _lib.NUM_ENERGIES), dtype=np.float64)
    hke = np.zeros(1, dtype=np.float64)
    gamma_int = GAMMA_PS / (1000.0 / 48.88821)

    def host_step(k):
        _lib.check(
            L.tmd_md_steps_host(
                forces._ctx, 1, hpos.data_ptr(), hvel.data_ptr(), system.forces.data_ptr(), integ.masses.data_ptr(),
                system.pos.data_ptr(), system.vel.data_ptr(), integ.dt, gamma_int, integ.vcoeff.data_ptr(),
                integ.seed, 10_000_000 + k, hene.ctypes.data, hke.ctypes.data, stream,
            )
        )

    for k in range(5):
        host_step(k)
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        host_step(5 + k)
    t_e2e = time.perf_counter() - t0
    bytes_each_way = 2 * system.pos.numel() * 4
    e2e = {
        "value": e2e_steps / t_e2e,
        "unit": "steps/s",
        "h2d_bytes_per_step": bytes_each_way,
        "d2h_bytes_per_step": bytes_each_way + hene.nbytes + hke.nbytes,
        "steps": e2e_steps,
        "api": "tmd_md_steps_host (C ABI, pinned host positions+velocities in and out every step)",
    }

    base = cpu_reference_run(3, 1, budget_s=40.0) if not args.no_cpu_baseline else None

    line = {
        "metric": "MD steps/sec (100k-atom water, fp32)",
        "value": value,
        "unit": "steps/s",
        "n_gpus": 1,
        "steps": args.steps,
        "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(1),
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(st1["kernel_launches"] - st0["kernel_launches"]),
        "roofline": roofline,
        "cpu_baseline": base,
        "state": {
            "temperature_K": float(T[0]),
            "epot": float(pot[0]),
            "rebuilds_in_timed_region": int(st1["rebuilds"] - st0["rebuilds"]),
            "max_neighbours": int(st1["max_neighbours"]),
            "row_capacity": int(st1["row_capacity"]),
            "skin_A": forces.skin,
            "equilibration_steps": args.equil,
        },
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--equil", type=int, default=600, help="relaxation steps before warm-up (lattice start)")
    ap.add_argument("--e2e-steps", type=int, default=300)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    main()
