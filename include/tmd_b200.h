/*
 * tmd_b200.h -- C ABI of the B200-native MD inner loop that sits behind
 * torchmd's Forces.compute() / Integrator.step().
 *
 * The reference (torchmd/torchmd) has no FFI: its hot path is stock torch ops
 * issued from Python.  Each entry point below therefore names the reference
 * Python it replaces (paths relative to the reference checkout), and
 * INTEGRATION.md shows the ctypes binding a maintainer would add there.
 *
 * Conventions
 *   - every function returns 0 on success, a negative TMD_ERR_* otherwise;
 *     tmd_last_error() gives the message (thread-local).  Nothing aborts.
 *   - "dev" pointers are CUDA device pointers owned by the caller (PyTorch
 *     tensors: fp32, contiguous).  "host" pointers are ordinary host memory,
 *     read during the call only (topology is copied into the context).
 *   - per-step calls only ENQUEUE work on the given CUDA stream: no
 *     allocation, no synchronisation, CUDA-graph capturable.  Calls marked
 *     [sync] synchronise the stream and may (re)allocate context scratch.
 *   - a context belongs to one device and is not thread-safe.
 *   - arithmetic is fp32 ("precision: single" in torchmd), energies are
 *     accumulated and returned in fp64.
 */
#ifndef TMD_B200_H
#define TMD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tmd_ctx tmd_ctx;
typedef void* tmd_stream; /* cudaStream_t */

enum {
  TMD_OK = 0,
  TMD_ERR_ARG = -1,      /* bad argument */
  TMD_ERR_CUDA = -2,     /* CUDA runtime error (message has the string) */
  TMD_ERR_STATE = -3,    /* call order / missing setup */
  TMD_ERR_OVERFLOW = -4, /* neighbour rows overflowed their capacity */
  TMD_ERR_UNSUPPORTED = -5
};

/* Energy slots, in the order of torchmd Forces.terms (forces.py:23-25). */
enum {
  TMD_E_BONDS = 0,
  TMD_E_ANGLES = 1,
  TMD_E_DIHEDRALS = 2,
  TMD_E_IMPROPERS = 3,
  TMD_E_14 = 4, /* stays 0: 1-4 energies are booked under lj/electrostatics, forces.py:185-236 */
  TMD_E_ELECTROSTATICS = 5,
  TMD_E_LJ = 6,
  TMD_E_REPULSION = 7,
  TMD_E_REPULSIONCG = 8,
  TMD_NUM_ENERGIES = 9
};

/* Term bit mask (1 << energy slot). */
#define TMD_TERM(slot) (1u << (slot))

const char* tmd_last_error(void);
int tmd_version(void);

/* ---- lifetime ------------------------------------------------------------ */

/* One context per (device, system).  Replaces nothing in the reference; it is
 * the state Forces.__init__ (forces.py:27-74) keeps in Python attributes. */
int tmd_create(tmd_ctx** out, int device, int natoms, int nreplicas);
int tmd_destroy(tmd_ctx* ctx);

/* ---- topology / parameters: host pointers, copied; once per run ----------- */

/* Per-atom charge and atom-type id, and the LJ A/B type tables
 * (Parameters.charges, .mapped_atom_types, .A, .B; parameters.py:449-457).
 * A and B may be NULL when no LJ/repulsion term is used. */
int tmd_set_atoms(tmd_ctx* ctx, const float* charges_host, const int32_t* types_host,
                  int ntypes, const float* A_host, const float* B_host);

/* Symmetric exclusion adjacency in CSR form (row_ptr has natoms+1 entries).
 * Replaces the N x N bool matrix of Forces._make_indeces (forces.py:348-357). */
int tmd_set_exclusions(tmd_ctx* ctx, const int64_t* row_ptr_host, const int32_t* cols_host);

/* Non-bonded set-up: which pair terms, cutoff (< 0: none), switch distance
 * (< 0: none), reaction field, solvent dielectric, Coulomb constant
 * (forces.py:375-378) and Verlet skin.  Arguments of Forces.__init__. */
int tmd_set_nonbonded(tmd_ctx* ctx, uint32_t term_mask, double cutoff, double switch_dist,
                      int rfa, double solvent_dielectric, double coulomb_constant, double skin);

/* Bonded terms, one parameter row per term instance (the reference's
 * params[map[:,1]] gather, forces.py:122-258, done once here).
 *   bonds    idx (n,2)  prm (n,2) = k, r0
 *   angles   idx (n,3)  prm (n,2) = k, theta0 [rad]
 *   torsions idx (n,4)  term_ptr (n+1) into terms (nterms,3) = k, phi0 [rad], periodicity
 *            which = 0 proper dihedrals, 1 impropers; amber_form = all(per>0) (forces.py:566)
 *   pairs14  idx (n,2)  prm (n,4) = A, B, scnb, scee */
int tmd_set_bonds(tmd_ctx* ctx, int n, const int32_t* idx_host, const float* prm_host);
int tmd_set_angles(tmd_ctx* ctx, int n, const int32_t* idx_host, const float* prm_host);
int tmd_set_torsions(tmd_ctx* ctx, int which, int n, const int32_t* idx_host,
                     const int32_t* term_ptr_host, const float* terms_host, int amber_form);
int tmd_set_pairs14(tmd_ctx* ctx, int n, const int32_t* idx_host, const float* prm_host);

/* Box diagonal per replica, host (nreplicas,3).  All zeros = no periodic
 * wrapping (forces.py:361).  Sizes the cell grid; call again if the box changes. [sync] */
int tmd_set_box(tmd_ctx* ctx, const float* box_diag_host);

/* ---- per-step work: device pointers, enqueue only -------------------------- */

/* Forces.compute(pos, box, forces) explicit-force path (forces.py:83-346):
 * overwrites forces_dev (R,N,3) and, if energies_dev != NULL, writes
 * (R, TMD_NUM_ENERGIES) doubles.  Internally: displacement check, (gated)
 * cell-list + neighbour-list rebuild, non-bonded pair kernel, bonded kernels. */
int tmd_forces(tmd_ctx* ctx, const float* pos_dev, float* forces_dev, double* energies_dev,
               tmd_stream stream);

/* _first_VV (integrator.py:61-64): pos += v dt + 0.5 (F/m) dt^2 ; v += 0.5 dt F/m */
int tmd_vv_first(tmd_ctx* ctx, float* pos_dev, float* vel_dev, const float* forces_dev,
                 const float* masses_dev, double dt, tmd_stream stream);

/* langevin (integrator.py:72-74) followed by _second_VV (integrator.py:67-69).
 * gamma < 0 or vcoeff_dev == NULL: no thermostat.  noise_dev (R,N,3) N(0,1)
 * draws, or NULL to draw in-kernel from Philox4x32-10(seed, step_index + number of
 * tmd_vv_first calls on this context so far -- the count lives on the device so a
 * captured CUDA graph of one step can be replayed).
 * ke_dev != NULL: also write the kinetic energy per replica (R doubles)
 * (kinetic_energy, integrator.py:8-30). */
int tmd_vv_second(tmd_ctx* ctx, float* vel_dev, const float* forces_dev, const float* masses_dev,
                  double dt, double gamma, const float* vcoeff_dev, const float* noise_dev,
                  uint64_t seed, uint64_t step_index, double* ke_dev, tmd_stream stream);

/* kinetic_energy (integrator.py:8-30) on its own. */
int tmd_kinetic_energy(tmd_ctx* ctx, const float* vel_dev, const float* masses_dev, double* ke_dev,
                       tmd_stream stream);

/* niter iterations of Integrator.step's loop body (integrator.py:115-120) with
 * no host round trip; energies/ke are those of the LAST iteration, which is all
 * Integrator.step returns (integrator.py:122-125).  noise_dev: (niter,R,N,3) or NULL. */
int tmd_md_steps(tmd_ctx* ctx, int niter, float* pos_dev, float* vel_dev, float* forces_dev,
                 const float* masses_dev, double dt, double gamma, const float* vcoeff_dev,
                 const float* noise_dev, uint64_t seed, uint64_t first_step_index,
                 double* energies_dev, double* ke_dev, tmd_stream stream);

/* Same as tmd_md_steps but with HOST state: copies pos/vel (R,N,3) in, runs,
 * copies pos/vel/energies/ke out and synchronises.  The end-to-end entry a
 * host-resident caller (e.g. minimizers.py:19-32 style) would use. [sync] */
int tmd_md_steps_host(tmd_ctx* ctx, int niter, float* pos_host, float* vel_host,
                      float* forces_dev, const float* masses_dev, float* pos_dev, float* vel_dev,
                      double dt, double gamma, const float* vcoeff_dev, uint64_t seed,
                      uint64_t first_step_index, double* energies_host, double* ke_host,
                      tmd_stream stream);

/* Which force the switched LJ term returns.  0 (default): the reference's explicit formula
 * s*dE/dr + E*s'/r with its extra 1/r (forces.py:410-412), what Forces.compute returns with
 * explicit_forces=True.  1: the exact derivative d(E*s)/dr = s*dE/dr + E*s', what the
 * reference obtains by autograd with explicit_forces=False (forces.py:328-336).  Every other
 * term's explicit force already is its exact gradient.  Takes effect at the next call. */
int tmd_set_force_convention(tmd_ctx* ctx, int exact_gradient);

/* ---- decomposed runs (one context per rank, every rank holds all positions) --------- */

/* Restrict the FORCE and INTEGRATION work of this context to the atoms
 * [first_atom, first_atom+count) (original indices): tmd_forces fills forces only for
 * them (their complete force: every partner is visible locally), tmd_vv_first /
 * tmd_vv_second / tmd_kinetic_energy update only them, energies are this rank's share
 * (sum over ranks = total).  The caller exchanges positions between tmd_vv_first and
 * tmd_forces (one all-gather per step).  Default: all atoms.  No reference counterpart
 * (the reference is single-device). */
int tmd_set_owned_atoms(tmd_ctx* ctx, int first_atom, int count);

/* ---- decomposed runs: integrate + position exchange over NVLink peer memory ------------
 *
 * Instead of a collective between tmd_vv_first and tmd_forces, the integration kernel of
 * every rank stores the new positions of its owned atoms straight into the position
 * buffers of ALL ranks (peer-to-peer stores through NVLink / NVSwitch), then raises a flag
 * in every rank's flag array; a one-warp kernel on each rank waits until all flags of the
 * step have arrived.  One launch + one tiny wait per step, no library collective on the
 * path.  Positions live in two buffers per rank (read buffer / write buffer alternate each
 * step) allocated by the context and shared between the processes of one node through CUDA
 * IPC.  Single replica.  No reference counterpart.
 *
 * Set-up, once: every rank calls tmd_dd_create (after tmd_set_owned_atoms), all-gathers the
 * 64-byte handles with whatever transport it has (torch.distributed), then tmd_dd_connect.
 * Per step, parity p = 0,1,0,...:  tmd_dd_vv_first_push(p)  tmd_dd_wait  tmd_dd_forces(1-p)
 * tmd_vv_second.  All four only enqueue; every per-step counter is device resident, so a
 * captured CUDA graph of a step of given parity can be replayed. */
#define TMD_MAX_PEERS 16
#define TMD_IPC_HANDLE_BYTES 64

/* Allocates the two position buffers and the flag array of this rank; handle_out receives
 * TMD_IPC_HANDLE_BYTES bytes to be sent to the other ranks. [sync] */
int tmd_dd_create(tmd_ctx* ctx, int rank, int world, unsigned char* handle_out_host);
/* handles_host: world * TMD_IPC_HANDLE_BYTES bytes, rank order (own entry ignored).  Opens the
 * peers' buffers in this process. [sync] */
int tmd_dd_connect(tmd_ctx* ctx, const unsigned char* handles_host);
/* Copy caller positions (1,N,3) into / out of position buffer `which` (0 or 1). */
int tmd_dd_load(tmd_ctx* ctx, int which, const float* pos_dev, tmd_stream stream);
int tmd_dd_store(tmd_ctx* ctx, int which, float* pos_dev, tmd_stream stream);
/* _first_VV (integrator.py:61-64) on the owned atoms: reads positions from buffer which_in,
 * writes the new ones into buffer 1-which_in of EVERY rank, then signals all ranks. */
int tmd_dd_vv_first_push(tmd_ctx* ctx, int which_in, float* vel_dev, const float* forces_dev,
                         const float* masses_dev, double dt, tmd_stream stream);
/* Wait (on the device) until every rank's stores of this step have landed here. */
int tmd_dd_wait(tmd_ctx* ctx, tmd_stream stream);
/* tmd_forces on position buffer `which`. */
int tmd_dd_forces(tmd_ctx* ctx, int which, float* forces_dev, double* energies_dev, tmd_stream stream);

/* ---- Wrapper.wrap (wrapper.py:8-30): molecules back into the box -----------------------
 *
 * Groups are the connected components of the bond graph (calculate_molecule_groups,
 * wrapper.py:33-55) as a CSR over atom indices: group g holds
 * group_atoms[group_ptr[g] .. group_ptr[g+1]); an atom without bonds is a group of one atom
 * (the reference's "nongrouped" branch is the same arithmetic).  tmd_wrapper_wrap moves every
 * group by  -floor(com / box) * box  per dimension, com being the plain mean of the group's
 * coordinates, in place on pos_dev (R,N,3), for every replica; box_dev is the (R,3,3) box
 * tensor (diagonal used).  If every box length is zero nothing happens (wrapper.py:14-15).
 * The optional re-centring on a wrap-index group (wrapper.py:17-21) rebinds a local name in
 * the reference and never reaches the caller's tensor; it is not part of this entry point.
 * Independent of tmd_ctx; only enqueues. */
typedef struct tmd_wrapper tmd_wrapper;
int tmd_wrapper_create(tmd_wrapper** out, int device, int natoms, int ngroups,
                       const int32_t* group_ptr_host, const int32_t* group_atoms_host);
int tmd_wrapper_wrap(tmd_wrapper* w, float* pos_dev, const float* box_dev, int nreplicas,
                     tmd_stream stream);
int tmd_wrapper_destroy(tmd_wrapper* w);

/* ---- inspection ------------------------------------------------------------ */

/* The reference's neighbour list for one replica: every non-excluded pair
 * (i<j, original atom indices) with dist <= cutoff under the reference's own
 * fp32 predicate (forces.py:76-81,264-269), unordered.  pairs_dev holds
 * capacity*2 int32; *count_dev receives the number found (may exceed capacity). */
int tmd_export_pairs(tmd_ctx* ctx, const float* pos_dev, int replica, int32_t* pairs_dev,
                     int64_t capacity, int64_t* count_dev, tmd_stream stream);

/* Which pair kernel the last tmd_forces / tmd_md_steps launched: 0 k_pair (float separations, the
 * default), 1 k_pair_fx (fixed-point separations), 2 k_pair_fx2 (fixed point + packed fp32x2
 * arithmetic), 3 k_pair2_open (no box, packed arithmetic).  For tests and bench labels. */
int tmd_pair_kernel(tmd_ctx* ctx);

typedef struct {
  int64_t rebuilds;        /* neighbour-list rebuilds so far (all replicas) */
  int64_t force_calls;     /* tmd_forces invocations */
  int32_t max_neighbours;  /* longest neighbour row seen */
  int32_t row_capacity;    /* entries reserved per atom */
  int32_t overflow;        /* 1 if a row ever overflowed (results invalid) */
  int32_t ncells[3];       /* cell grid of replica 0 */
  int64_t kernel_launches; /* kernels launched by this context so far */
} tmd_stats;

/* Reads counters back. [sync]  If a row overflowed, grows the capacity,
 * forces a rebuild on the next call and returns TMD_ERR_OVERFLOW once. */
int tmd_get_stats(tmd_ctx* ctx, tmd_stats* out, tmd_stream stream);

/* Device-side timing of the non-bonded pair kernel: between begin and end every
 * tmd_forces brackets its pair-kernel launch with CUDA events on the launching
 * stream (at most max_samples of them).  tmd_profile_end synchronises and returns
 * the summed and the number of sampled launches. [sync] */
int tmd_profile_begin(tmd_ctx* ctx, int max_samples);
int tmd_profile_end(tmd_ctx* ctx, double* total_ms, int* nsamples, tmd_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* TMD_B200_H */
