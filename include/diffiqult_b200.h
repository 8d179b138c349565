/* diffiqult_b200 — C ABI of the B200-native RHF hot path.
 *
 * Drop-in boundary: these entry points are what a binding of the reference's hot path
 * would call instead of the Python functions cited beside each one.  Plain pointers and sizes only; all
 * pointers are HOST memory unless stated, FP64, row-major, int32 indices.  Every function returns a status
 * (0 = ok) and never exits the process (the reference's Boys failure paths call exit(): Tools.py:91,125,151).
 * Negative statuses: -1 bad argument, -2 CUDA error, -3 no CUDA device (there is no CPU path), -4 size limit of the
 * batched path, -5 a Boys-function loop did not converge (non-finite input; where the reference exits), -9 other.
 * No global mutable state besides the launch counter: callable from several host threads, one per GPU.
 *
 * Conventions shared with the reference:
 *   ltype[i]      0 = s, 1/2/3 = p_x/p_y/p_z   (the reference's l tuples (0,0,0),(1,0,0),(0,1,0),(0,0,1))
 *   contr[i]      primitives of basis function i (Molecule.py:65 list_contr); primitives are stored
 *                 function after function in alpha[] / coef[]
 *   packed ERI    Eri_vec[eri_index(i,j,k,l)]  (Tools.py:17-51), length M(M+1)/2, M = N(N+1)/2
 *   ne            number of doubly-occupied orbitals = electrons // 2 (Molecule.py:184)
 */
#ifndef DIFFIQULT_B200_H
#define DIFFIQULT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dq_system {
  int32_t natoms;
  const int32_t* charges;  /* [natoms]    System_mol.charges */
  const double* atoms;     /* [natoms*3]  System_mol.atom (bohr), kept fixed: Gaussians float */
  int32_t nbasis;
  const int32_t* contr;    /* [nbasis]    System_mol.list_contr */
  const int32_t* ltype;    /* [nbasis] */
  int32_t nprim;
  const double* alpha;     /* [nprim]     exponents, or ln(exponent) when log_alpha != 0 (Task.py:135-136) */
  const double* coef;      /* [nprim]     contraction coefficients (raw or normalised, see each function) */
  const double* xyz;       /* [nbasis*3]  Gaussian centres */
  int32_t ne;
  int32_t log_alpha;
} dq_system;

/* in-place sum over ranks of `count` doubles at DEVICE pointer `dev`.  The collective must be enqueued ON `stream` (the
 * cudaStream_t the library runs on, = dq_options.stream; NULL = the default stream): the buffer is written by kernels
 * queued on that stream before the call and read by kernels queued on it afterwards.  Supplied by the host language
 * (Python: torch.distributed.all_reduce over NCCL, diffiqult_b200/distributed.py).  If a rank fails before reaching a
 * collective the other ranks block in it: abort the job on any non-zero status. */
typedef void (*dq_allreduce_fn)(void* dev, int64_t count, void* stream, void* user);

typedef struct dq_options {
  int32_t max_scf;         /* Energy.py:164   (Tasks defaults: 30 'Energy'/'Grad', 100 'Opt') */
  double e_tol;            /* Energy.py:117-118, 1e-8 when <= 0 */
  int32_t device;          /* CUDA device ordinal */
  void* stream;            /* cudaStream_t to run on (NULL: the default stream) */
  int32_t world_size;      /* ERI tiles are dealt round-robin to `world_size` ranks ... */
  int32_t rank;            /* ... this process computes the tiles of `rank` */
  dq_allreduce_fn allreduce;   /* required when world_size > 1: sums partial J|K and gradient buffers */
  void* allreduce_user;
  int32_t jk_mode;         /* Fock build (Energy.py:29-38): 0 auto, 1 stored ERI tiles (computed once, 8 B per contracted
                            * quartet kept in HBM), 2 integral-direct (every build re-evaluates the tiles and contracts each one
                            * with D inside the same kernel: nothing is stored) */
  int64_t direct_batch_bytes;  /* unused since the fused integral-direct kernel (kept for ABI stability; was: transient buffer size) */
  double screen_threshold; /* 0 (default): every quartet is evaluated, as in the reference, which screens nothing.
                            * > 0 (opt-in): a tile is skipped when 2 pi^(5/2) max|K_ab| max|K_cd| over its two block
                            * pairs is below the threshold (an upper bound of its primitive prefactors) */
  const double* guess_C;   /* readguess (Energy.py:148-157): [nbasis*nbasis] AO x MO coefficients; the first Fock matrix is
                            * built from D = C_occ C_occ^T instead of the core Hamiltonian.  NULL: core guess */
  int32_t accumulate;      /* how J, K, the pair adjoints and the gradient are summed over warps / CTAs: 0 or 1 (default) =
                            * fixed point - every accumulator is a pair of 64-bit integer limbs (2^-24 and 2^-67 units), so the
                            * sums do not depend on the order of the atomics and two calls on the same inputs return identical
                            * bits (the reference's BFGS line search amplifies last-bit noise into different
                            * iteration counts); 2 = double-precision floating-point atomics (results vary in the last bits
                            * from run to run).  The environment variable DQ_ACCUM=fixed|float overrides. */
} dq_options;

typedef struct dq_stats {
  /* CUDA-event times of the last call: host planning + uploads, S/T/V, the ERI tile kernels, the SCF loop, its reverse
   * sweep, the ERI-gradient tile kernels alone, the rest of the gradient assembly, and the whole call */
  double ms_setup, ms_one_electron, ms_eri, ms_scf, ms_reverse, ms_eri_grad, ms_grad_other, ms_total;
  int64_t launches;        /* kernels launched by the last call */
  int64_t n_tiles;         /* ERI tiles evaluated by this rank */
  int64_t n_quartets;      /* contracted quartets (valid lanes) evaluated by this rank */
  int64_t n_prim_quartets; /* primitive quartets evaluated by this rank (per ERI pass) */
  int32_t scf_iterations;
  int32_t fock_builds;     /* J/K contractions (forward + reverse sweep) */
  int32_t grad_terms;      /* rank of the two-particle weight used by the ERI-gradient pass */
  int32_t jk_direct;       /* 1 if the Fock builds were integral-direct, 0 if they used the stored tiles */
  int64_t n_tiles_skipped; /* tiles skipped by the opt-in screening in the (last) ERI pass */
  double ms_jk;            /* sum of the J/K contraction kernels' times over all Fock builds (CUDA events, read once at the end) */
  int64_t n_grad_quartets_zero_weight;  /* contracted quartets of this rank whose two-particle weight in the ERI-gradient pass is
                            * EXACTLY 0.0 (both density-like factors vanish identically, e.g. between uncoupled far-apart
                            * fragments): they add exactly nothing and their primitive loops are skipped */
  int64_t n_prim_pairs;    /* primitive pairs over all function-block pairs (both orders of a diagonal block counted) */
  int64_t n_prim_pairs_underflowed;  /* ... of which the Gaussian-product exponential underflowed to EXACTLY 0: primitive
                            * quartets containing one are exact zeros (value and every derivative) and are left out of the
                            * kernels' loops - bit-identical results; DQ_NO_ZERO_SKIP=1 evaluates them anyway */
  int32_t grad_kernel_sparse;   /* 1: the ERI-gradient pass used the lane-independent kernel for sparse weights (chosen when
                                 * fewer than 25 % of the elements of the final density matrix are nonzero), 0: the dense one */
  int32_t eri_kernel_compact;   /* 1: the ERI pass used the kernel that compacts the surviving primitive pairs per tile (chosen
                                 * when more than 25 % of the primitive-pair prefactors underflowed to exactly 0), 0: the plain one */
} dq_stats;

/* Work decomposition of the ERI passes for one rank, computed on the host only (usable without a GPU):
 * out[6] = {function blocks, block pairs, tiles of this rank, contracted quartets of this rank, primitive
 * quartets of this rank, tiles of all ranks}.  Replaces the loop nest of Integrals.py:190-216 (erivector). */
int dq_plan(const dq_system* sys, int32_t world_size, int32_t rank, int64_t* out);

/* Integrals.py:477-501 normalization(alpha, c, l, list_contr): sys->coef raw -> out[nprim] */
int dq_normalization(const dq_system* sys, const dq_options* opt, double* out);

/* Integrals.py:10-20, 51-62, 135-145 overlapmatrix / nuclearmatrix / kineticmatrix.
 * sys->coef must already be normalised (as the reference's callers pass it). S,T,V: [nbasis*nbasis]. */
int dq_one_electron(const dq_system* sys, const dq_options* opt, double* S, double* T, double* V);

/* Integrals.py:190-216 erivector: sys->coef normalised; eri: packed, reference order. */
int dq_eri_packed(const dq_system* sys, const dq_options* opt, double* eri);

/* Energy.py:29-38 fockmatrix(Hcore, Eri, D): integrals are recomputed from sys (coef normalised);
 * D: [nbasis*nbasis] symmetric; outputs F = Hcore + G[D] and (optional, may be NULL) Hcore. */
int dq_fock(const dq_system* sys, const dq_options* opt, const double* D, double* F, double* Hcore);

/* Energy.py:75-324 rhfenergy (eigen=True, readguess=None): sys->coef RAW.
 * returns 0 converged, 1 = SCF not converged (the reference's 99999 sentinel; *E still holds the last
 * E_elec+E_nuc), negative = error.  esteps: [max_scf] per-iteration E_elec; C: [nbasis*nbasis] AO x MO;
 * eps: [nbasis] orbital energies.  Any of esteps, C, eps may be NULL. */
int dq_rhf(const dq_system* sys, const dq_options* opt, double* E, int32_t* niter, double* esteps, double* C,
           double* eps);

/* Task.py:25-51 function_grads_algopy / grad_evaluator_algopy: gradient of the SAME truncated SCF energy,
 * w.r.t. sys->alpha as passed (ln alpha if log_alpha), the raw coefficients, and the Gaussian centres
 * (row-major [nbasis][3]).  One call returns all three groups; any output may be NULL. */
int dq_rhf_grad(const dq_system* sys, const dq_options* opt, double* E, int32_t* niter, double* g_alpha,
                double* g_coef, double* g_xyz);

/* dq_rhf_grad plus the gradient w.r.t. the NUCLEAR coordinates, g_atoms[natoms*3] (may be NULL): the reference's
 * `type(xyz_atom) != np.ndarray` branch (Energy.py:125-130, UTPM on argument slot 5): the Gaussians do not follow the
 * nuclei, so only V (through the SCF response) and E_nuc depend on them.  Unreachable from the reference's Tasks. */
int dq_rhf_grad_ex(const dq_system* sys, const dq_options* opt, double* E, int32_t* niter, double* g_alpha,
                   double* g_coef, double* g_xyz, double* g_atoms);

/* Batch of `nmol` molecules of ONE structure (shape: natoms, charges, nbasis, contr, ltype, nprim, ne, log_alpha; its
 * atoms/alpha/coef/xyz pointers are ignored): the reference evaluates such a batch as nmol separate
 * Tasks.runtask('Energy'|'Grad') calls (Task.py:242-278, 223-240).  Per-molecule inputs: atoms [nmol][natoms*3],
 * alpha [nmol][nprim] (ln alpha if shape->log_alpha), coef [nmol][nprim] raw, xyz [nmol][nbasis*3].  Outputs: E [nmol],
 * status [nmol] (0 converged / 1 = 99999 sentinel), niter [nmol] and, when want_grad, the three gradient groups
 * [nmol][nprim], [nmol][nprim], [nmol][nbasis*3] (any may be NULL).  Every molecule follows its own truncated SCF exactly
 * as dq_rhf_grad does.  nbasis <= 40, at most 40 primitive pairs per function pair (status -4 otherwise). */
int dq_rhf_grad_batch(const dq_system* shape, int32_t nmol, const double* atoms, const double* alpha, const double* coef,
                      const double* xyz, const dq_options* opt, int32_t want_grad, double* E, int32_t* status, int32_t* niter,
                      double* g_alpha, double* g_coef, double* g_xyz);

/* Eigen.py:69-74 eigensolver on plain doubles: batch of symmetric n x n matrices -> ascending eigenvalues
 * w[batch*n], eigenvectors in the columns of V[batch*n*n]. */
int dq_eigh(int32_t n, int32_t batch, const double* A, double* w, double* V, int32_t device);

/* Tools.py:78-83 incompletegammaf(t, m) for m = 0..4 and d/dt through the same truncated loops.
 * F, dF: [n*5]. */
int dq_boys(int32_t n, const double* t, double* F, double* dF, int32_t device);

/* measured FP64 FMA throughput of the device (TFLOP/s): the FP64 roofline denominator */
int dq_fp64_peak(int32_t device, double* tflops);

void dq_get_stats(dq_stats* out);       /* stats of the last call made by this thread */
const char* dq_last_error(void);        /* message of the last non-zero status of this thread */
int dq_device_count(void);
/* device buffers are pooled between calls (same-size reuse); this returns them to the driver */
void dq_release_cache(void);

#ifdef __cplusplus
}
#endif
#endif
