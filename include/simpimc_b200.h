/* simpimc_b200 -- C ABI of the B200-native action-evaluation path.
 *
 * Drop-in boundary for simpimc's `Action` operator API (src/actions/action_class.h:7-74)
 * restricted to the pair actions, the Ewald k-space sums, the move action-deltas and the
 * E / g(r) / S(k) estimator reductions.  Plain C: opaque handles, pointers and sizes, no
 * exceptions and no torch / C++ types cross this boundary.  Every entry point returns
 * PIMC_OK (0) or a negative pimc_status; pimc_last_error() gives the message.
 *
 * Where the reference holds ONE walker per process (src/framework/framework_class.h:44-53)
 * a context here holds `n_clones` independent walkers of identical shape on one GPU; every
 * per-walker scalar of the reference becomes an array of n_clones doubles.  With
 * n_clones = 1 the calls are one-to-one with the reference's.
 *
 * Host layouts are the reference's: positions R[clone][particle][bead][dim]
 * (species_class.h:66-70, bead_(p,b)).  Device layouts are private to the library.
 *
 * OLD/NEW semantics (path_class.h:95, bead_class.h:98): the context keeps the committed
 * path (the reference's r_c / rho_k_c) plus, per species and clone, at most one pending
 * proposal written by pimc_propose (the reference's r of the beads a move touched).
 * mode = PIMC_OLD reads the committed path, PIMC_NEW reads it overlaid with the proposal.
 * pimc_commit is Move::Accept / Move::Reject (bisect_class.h:24-36,127-139).
 *
 * Thread-compatibility: calls on one context must not overlap; different contexts are
 * independent.  All work of a context is issued on one CUDA stream.
 */
#ifndef SIMPIMC_B200_H_
#define SIMPIMC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    PIMC_OK = 0,
    PIMC_ERR_INVALID = -1,     /* bad argument (reference: assert / abort, io_xml.h:118) */
    PIMC_ERR_UNSUPPORTED = -2, /* valid in the reference, outside this path (e.g. max_level > 0) */
    PIMC_ERR_CUDA = -3,        /* CUDA runtime failure; no CPU fallback exists */
    PIMC_ERR_TABLE = -4        /* malformed pair-action table (reference: exit(1), david...:241) */
} pimc_status;

enum { PIMC_OLD = 0, PIMC_NEW = 1 }; /* ModeType, bead_class.h:7-8 */

typedef struct pimc_ctx pimc_ctx;       /* Path + Species[] + KSpace for n_clones walkers */
typedef struct pimc_action pimc_action; /* one PairAction subclass instance */

/* <System> and <Species> attributes (path_class.h:25-69, species_class.h:48-60). */
typedef struct {
    int32_t n_d;           /* spatial dimensions; this build evaluates n_d = 3 */
    int32_t pbc;           /* periodic box (path_class.h:29) */
    double L;              /* box side, ignored when pbc = 0 */
    double beta;           /* inverse temperature; tau = beta / n_bead */
    int32_t n_bead;        /* M: time slices of the whole path */
    int32_t n_species;
    const int32_t *n_part; /* [n_species] */
    const double *lambda;  /* [n_species] hbar^2/2m */
    int32_t n_clones;      /* independent walkers batched in this context */
    int32_t device;        /* CUDA device ordinal */
    int32_t slice_lo;      /* time-slice shard [slice_lo, slice_hi) owned by this context; */
    int32_t slice_hi;      /* 0, n_bead for an unsharded path */
} pimc_config;

/* ---- pair-action tables: the datasets the reference constructors read ------------------ */
typedef struct {
    int32_t n;
    const double *r; /* grid, ascending */
    const double *f; /* values on the grid */
} pimc_table_1d;

typedef struct {
    int32_t n_x, n_y;
    const double *x, *y;
    const double *f; /* row-major f[ix*n_y + iy] (ilkka_pair_action_class.h:272-280) */
} pimc_table_2d;

/* <obj>/diag/{r_long, <obj>_long_r, <obj>_long_r_0, k, <obj>_long_k, <obj>_long_k_0} */
typedef struct {
    pimc_table_1d f_r;
    double f_r_0;
    int32_t n_k;
    const double *k;   /* |k| shell list, first entry 0 (Ewald.py:295-318) */
    const double *f_k; /* value per shell */
    double f_k_0;
} pimc_long_range;

typedef struct { /* ilkka_pair_action_class.h:266-418 */
    pimc_table_2d u_xy, du_xy;
    pimc_table_1d v_r;
    pimc_long_range u_long, du_long, v_long; /* read only when use_long_range */
} pimc_ilkka_tables;

typedef struct { /* bare_pair_action_class.h:38-95 */
    pimc_table_1d v_r;
    pimc_long_range v_long;
    int32_t is_coulomb; /* analytic 1/2r + 1/2r' instead of the v spline (bare...:106-108) */
} pimc_bare_tables;

enum { PIMC_GRID_GENERAL = 0, PIMC_GRID_LOG = 1, PIMC_GRID_LINEAR = 2 };
typedef struct { /* david_pair_action_class.h:194-336 */
    int32_t grid_type;
    double r_start, r_end;
    int32_t n_grid;
    const double *grid_points; /* used when grid_type = GENERAL */
    int32_t n_order;           /* n_val = 1 + sum_{i=1..n_order}(1+i) */
    int32_t n_tau;             /* = max_level + 1; this build takes 1 */
    const double *taus;        /* [n_tau] */
    const double *u_kj;        /* file order [n_grid][n_val][n_tau] */
    const double *du_kj_dbeta; /* same shape */
    const double *potential;   /* [n_grid] */
    int32_t n_k;               /* long_range/{n_k,k_points,u_k}, squarer/v_image */
    const double *k_points;
    const double *u_k;
    double v_image;
} pimc_david_tables;

/* ---- context ------------------------------------------------------------------------ */
const char *pimc_last_error(void);
int pimc_version(void);
int pimc_ctx_create(const pimc_config *cfg, pimc_ctx **out);
int pimc_ctx_destroy(pimc_ctx *ctx);
int pimc_ctx_sync(pimc_ctx *ctx);           /* wait for the context's stream */
void *pimc_ctx_stream(pimc_ctx *ctx);       /* the cudaStream_t all kernels run on */

/* KSpace::Setup (k_space_class.h:33-80): grow-only, returns the vector count. */
int pimc_kspace_setup(pimc_ctx *ctx, double k_cut, int32_t *n_k);
/* k_index[n_k][n_d] (signed lattice indices, reference order), k_mag[n_k]. */
int pimc_kspace_get(pimc_ctx *ctx, int32_t *k_index, double *k_mag);

/* Species::InitPaths + InitRhoK (species_class.h:222-403): set r and r_c of clones
 * [clone_lo, clone_hi) of one species from host memory and rebuild rho_k.
 * R[clone][particle][bead][dim], bead running over the context's shard
 * [slice_lo, slice_hi] -- INCLUDING the halo slice slice_hi (mod n_bead) when sharded;
 * exactly n_bead beads when unsharded. */
int pimc_positions_upload(pimc_ctx *ctx, int32_t species, int32_t clone_lo, int32_t clone_hi, const double *R);
int pimc_positions_download(pimc_ctx *ctx, int32_t species, int32_t mode, int32_t clone_lo, int32_t clone_hi, double *R);
/* Same, from / to DEVICE memory in the library's own layout R[clone][particle][dim][slice], each
 * slice row padded to a multiple of 4 doubles. */
int pimc_positions_set_device(pimc_ctx *ctx, int32_t species, const double *d_R);
double *pimc_positions_device_ptr(pimc_ctx *ctx, int32_t species);
int pimc_rhok_rebuild(pimc_ctx *ctx, int32_t species); /* Species::InitRhoK */
/* Slice-shard halo (no reference counterpart; the reference never splits a path): pack the
 * FIRST owned slice of every clone and particle into d_buf[clone][particle][dim] (device
 * memory) for the rank that owns the preceding shard, and store a received buffer into the
 * halo slot that follows the last owned slice.  pimc_halo_exchange (below) does pack, the
 * NCCL ring and unpack in one call; the two halves are exported for callers with their own transport. */
int pimc_halo_pack(pimc_ctx *ctx, int32_t species, double *d_buf);
int pimc_halo_unpack(pimc_ctx *ctx, int32_t species, const double *d_buf);
/* Ring rotation of a slice-sharded path by `shift` slices (1..slices of the shard): pack this
 * shard's first `shift` owned slices ([n_clones][N][3][shift] doubles, device memory), send them to
 * the PREVIOUS rank, and apply what the NEXT rank sent: owned slices slide left, the received ones
 * are appended.  Global slice labels rotate, every shard keeps its range; imaginary time is a ring,
 * so actions and estimators are unchanged.  Afterwards refresh the halos (pimc_halo_*) and rho_k
 * (pimc_rhok_rebuild).  This is what lets shard-boundary slices be moved by pimc_bisect_sweep, whose
 * windows on a sharded context stay inside the shard's stored slices. */
int pimc_rotate_pack(pimc_ctx *ctx, int32_t species, int32_t shift, double *d_buf);
int pimc_rotate_apply(pimc_ctx *ctx, int32_t species, int32_t shift, const double *d_buf);
/* rho_k of one clone, out[bead][n_k][2] (re, im) (Species::GetRhoK, species_class.h:428). */
int pimc_rhok_download(pimc_ctx *ctx, int32_t species, int32_t mode, int32_t clone, double *out);

/* ---- actions (ActionFactory, actions.h:13-35; PairAction ctor, pair_action_class.h:204-238) */
int pimc_action_create_ilkka(pimc_ctx *ctx, int32_t species_a, int32_t species_b, const pimc_ilkka_tables *t,
                             int32_t max_level, int32_t use_long_range, double k_cut, pimc_action **out);
int pimc_action_create_bare(pimc_ctx *ctx, int32_t species_a, int32_t species_b, const pimc_bare_tables *t,
                            int32_t max_level, int32_t use_long_range, double k_cut, pimc_action **out);
int pimc_action_create_david(pimc_ctx *ctx, int32_t species_a, int32_t species_b, const pimc_david_tables *t,
                             int32_t max_level, int32_t use_long_range, pimc_action **out);
/* Kinetic (src/actions/single_action/kinetic_class.h): the free-particle action of one species with
 * n_images periodic images of the density matrix (FreeSpline, src/actions/free_spline_class.h:25-84:
 * the image sum tabulated on 10 000 points over [-L/2, L/2], natural cubic spline per dimension;
 * n_images = 0 is the closed form -|r|^2 / 4 lambda tau).  The handle answers pimc_action_dbeta
 * (DActionDBeta, kinetic_class.h:35-45), pimc_action_get at any level 0..5 (GetAction, :105-122: only
 * the listed particles of its species), pimc_action_total, and pimc_action_potential (0, Action's
 * default); the device-resident sweeps use its n_images for the kinetic part of the acceptance. */
int pimc_action_create_kinetic(pimc_ctx *ctx, int32_t species, int32_t n_images, pimc_action **out);
int pimc_action_destroy(pimc_action *act);

/* Action::DActionDBeta (pair_action_class.h:241-264) and Action::Potential (:369-395),
 * one double per clone, long-range k-sum and constants included.  `out` is host memory.
 * For a sharded context the value is this shard's partial sum (constants on the rank
 * that owns slice 0); the caller all-reduces. */
int pimc_action_dbeta(pimc_action *act, double *out);
int pimc_action_potential(pimc_action *act, double *out);
/* Same, result left in device memory (for NCCL all-reduce without a host round trip). */
int pimc_action_dbeta_device(pimc_action *act, double *d_out);
int pimc_action_potential_device(pimc_action *act, double *d_out);

/* Action::GetAction(b0, b1, particles, level) (pair_action_class.h:267-302).
 *   b0[n_clones]                first slice of each clone's window; b1 = b0 + n_window
 *   moved_species[n_moved]      species of each moved particle (same for all clones); up to sixteen
 *                               particles of each of the action's species (permutation cycles)
 *   moved_particle[n_clones][n_moved]
 * Returns 0 for level > max_level or constant actions, as the reference does.  In NEW mode
 * with use_long_range the first call after pimc_commit / pimc_action_accept refreshes the
 * proposal's rho_k (Species::UpdateRhoK, species_class.h:406-425). */
int pimc_action_get(pimc_action *act, int32_t mode, const int32_t *b0, int32_t n_window, int32_t n_moved,
                    const int32_t *moved_species, const int32_t *moved_particle, int32_t level, double *out);
/* Action::GetActionGradient / GetActionLaplacian (action_class.h:46,49; pair_action_class.h:305-366):
 * derivative of the action of the window [b0, b0 + n_window) with respect to the species-a bead of
 * every pair that touches a listed particle (forward and backward link of each slice; Ilkka:
 * analytic gradient + k-space force, Bare / David and every Laplacian: the reference's central
 * differences with eps = 1e-4).  Arguments as pimc_action_get; grad is [n_clones][3], lap
 * [n_clones] (host memory).  No proposal may be pending (the estimators run between moves). */
int pimc_action_gradient(pimc_action *act, const int32_t *b0, int32_t n_window, int32_t n_moved, const int32_t *moved_species,
                         const int32_t *moved_particle, int32_t level, double *grad);
int pimc_action_laplacian(pimc_action *act, const int32_t *b0, int32_t n_window, int32_t n_moved, const int32_t *moved_species,
                          const int32_t *moved_particle, int32_t level, double *lap);
/* Whole-path action of every clone: GetAction(0, n_bead, all particles, 0) in OLD mode. */
int pimc_action_total(pimc_action *act, double *out);
int pimc_action_total_device(pimc_action *act, double *d_out);
int pimc_action_accept(pimc_action *act); /* PairAction::Accept, pair_action_class.h:406-411 */
int pimc_action_reject(pimc_action *act); /* PairAction::Reject, :414-419 */

/* Per-pair kernels on caller-supplied distances (tests; CalcU / CalcdUdBeta / CalcV).
 * which: 0 = U, 1 = dU/dbeta, 2 = V.  Arrays are host memory of length n. */
int pimc_action_calc_pair(pimc_action *act, int32_t which, int32_t n, const double *r, const double *r_p,
                          const double *s, int32_t level, double *out);

/* Same through the fast Ilkka evaluation the whole-path kernel uses (shared-memory table layout,
 * uniform interval tables; csrc/pair_fast.cuh).  which: 0 = U, 1 = dU/dbeta. */
int pimc_action_calc_pair_fast(pimc_action *act, int32_t which, int32_t n, const double *r, const double *r_p,
                               const double *s, double *out);
/* Device square root of the fast path on caller data (tests: <= 1 ulp from sqrt). */
int pimc_debug_fast_sqrt(pimc_ctx *ctx, int32_t n, const double *x, double *out);
/* HOST-ONLY (no device work; callable without a GPU): builds the interval table of the fast kernels for
 * grid[n] -- kind 0: uniform buckets, kind 1: IEEE-754 bit-pattern buckets (spline_build.h) -- and runs the
 * device's lookup arithmetic (key, table entry, one compare against the next knot) for x[m] >= 0 on the host.
 * out[i] = einspline's interval of x[i] (nubspline general_grid_reverse_map semantics: 0 below the grid,
 * n - 1 at or above its end); n_keys = table size.  PIMC_ERR_UNSUPPORTED if the grid admits no such table. */
int pimc_debug_interval_table(int32_t kind, int32_t n, const double *grid, int32_t m, const double *x, int32_t *out,
                              int32_t *n_keys);
/* HOST-ONLY: the natural cubic spline through (grid[n], values[n]) evaluated at x[m] (clamped to the grid, as SetLimits
 * does) in the two layouts of the fast 1-D tables (csrc/pair_fast.cuh: FastPP1) -- out_interval: one cubic per interval about
 * its knot (einspline's interval); out_bucket: the bucket-centred records with the device's lookup arithmetic.  The two are
 * the same piecewise polynomial: they agree to rounding everywhere, the knots and their floating-point neighbours included. */
int pimc_debug_bucket_spline(int32_t n, const double *grid, const double *values, int32_t m, const double *x, double *out_interval,
                             double *out_bucket, int32_t *n_keys);
/* enable != 0: evaluate with the general kernels even where the fast path applies (tests). */
int pimc_ctx_force_general(pimc_ctx *ctx, int32_t enable);

/* ---- moves' contract ------------------------------------------------------------------ */
/* NEW-mode Bead::SetR for one particle per clone (bisect_class.h:89-94,
 * displace_particle_class.h:40-50): beads b_first[c] .. b_first[c]+n_beads-1 (mod n_bead)
 * of particle[c] take newR[c][i][dim].  Calling it again for a species that already has a pending
 * proposal ADDS a particle (same n_beads, a different particle in every clone; at most sixteen per
 * species): the particles of a permutation cycle. */
int pimc_propose(pimc_ctx *ctx, int32_t species, const int32_t *particle, const int32_t *b_first, int32_t n_beads,
                 const double *newR);
/* Committed positions of beads b_first[c] .. b_first[c]+n_beads-1 (mod n_bead) of
 * particle[c], out[c][i][dim] (what Bisect::Attempt reads through GetBead / GetNextBead,
 * bisect_class.h:51-58). */
int pimc_beads_download(pimc_ctx *ctx, int32_t species, const int32_t *particle, const int32_t *b_first, int32_t n_beads,
                        double *out);
/* Move::Accept (accept[c] != 0) or Move::Reject for the pending proposals of all species:
 * StoreR/StoreRhoK or RestoreR/RestoreRhoK, then every action's flag is re-armed. */
int pimc_commit(pimc_ctx *ctx, const int32_t *accept);

/* n_attempts x Bisect::DoEvent (move_class.h:61-77 -> bisect_class.h:39-139) on every clone,
 * entirely on the device: particle and window choice, Levy construction, the free-particle
 * action (with_kinetic = 0 leaves it out; the species' Kinetic handle, if one exists, supplies its
 * n_images, otherwise n_images = 0; sampling probabilities: pimc_move_set_images), every pair
 * action of the context that involves `species` in OLD and NEW mode, the Metropolis test per
 * level, Accept / Reject.  Random numbers are Philox4x32-10 (key = seed, counter = attempt0 + i,
 * clone, slot): the stream is reproducible and documented in csrc/mc.cuh, but it is not
 * std::mt19937 -- sampled runs match the reference statistically.  n_accept[n_clones] (host,
 * may be NULL) is ADDED to. */
int pimc_bisect_sweep(pimc_ctx *ctx, int32_t species, int32_t n_level, int32_t n_attempts, uint64_t seed, uint64_t attempt0,
                      int32_t with_kinetic, int64_t *n_accept);

/* The same move on EVERY disjoint window of a walker at once -- for one large path (few walkers, e.g. a slice
 * shard of BASELINE config C5), where one window per launch leaves the GPU idle.  A round tiles the path (the
 * shard's stored slices) with W = floor(slices / 2^n_level) windows starting at offset + w 2^n_level, the offset
 * drawn once per walker and round; every window picks its own particle, builds its own Levy bridge and takes its
 * own Metropolis decision.  The windows share only their fixed end-point slices: the pair action at level 0
 * couples a moved bead with the other particles' beads of the same and the neighbouring slice
 * (pair_action_class.h:282-288) and rho_k is slice-local, so the W updates are independent and their product
 * satisfies detailed balance like W sequential Bisect::DoEvent calls on those windows.  Philox counter:
 * (attempt0 + round, walker * W + window, slot).  n_accept[n_clones] is ADDED to; *n_windows returns W. */
int pimc_bisect_sweep_windows(pimc_ctx *ctx, int32_t species, int32_t n_level, int32_t n_rounds, uint64_t seed, uint64_t attempt0,
                              int32_t with_kinetic, int64_t *n_accept, int32_t *n_windows);

/* Bisect's own n_images attribute (bisect_class.h:173: the FreeSplines of the Levy sampling
 * probabilities, tau 2^level / 2) for later pimc_bisect_sweep calls on this species; default 0. */
int pimc_move_set_images(pimc_ctx *ctx, int32_t species, int32_t n_images);

/* n_attempts x DisplaceParticle::DoEvent (displace_particle_class.h:13-89) on every clone, on the
 * device: one particle's whole path shifted by a vector of length step_size (direction: normalised
 * uniform point of the cube, scaffold/rng/rng.h:38-56), every pair action that involves `species`
 * over all slices in OLD and NEW mode, Metropolis, Accept / Reject.  Same Philox convention as
 * pimc_bisect_sweep (slots: 0 particle, 1-2 direction, 3 Metropolis uniform). */
int pimc_displace_sweep(pimc_ctx *ctx, int32_t species, double step_size, int32_t n_attempts, uint64_t seed, uint64_t attempt0,
                        int64_t *n_accept);

/* Permutation table of the permuting bisection moves for the window [b0, b0 + n_bisect_beads] of
 * every clone: t[c][i][j] = exp(e) if e > log(epsilon) else 0 with
 *   relative == 0: e = -|Dr(r_i(b0), r_j(b1))|^2 / (4 lambda tau n)     PermBisectIterative::UpdatePermTable
 *                                                                        (perm_bisect_iterative_class.h:10-30)
 *   relative != 0: e = (-|Dr_ij|^2 + |Dr_ii|^2) / (4 lambda tau n)      PermBisectTable (perm_bisect_table_class.h:36-50)
 * on the unpermuted path.  Cycle selection and the relabelling of an accepted permutation stay with
 * the caller.  t is host memory, [n_clones][N][N]. */
int pimc_perm_table(pimc_ctx *ctx, int32_t species, const int32_t *b0, int32_t n_bisect_beads, double epsilon, int32_t relative, double *t);

/* Permuting bisection on the device: n_attempts x PermBisectIterative::DoEvent
 * (src/events/moves/single_species_move/bisect/perm_bisect/perm_bisect_iterative_class.h:113-222 on
 * perm_bisect_class.h:32-82) on every clone.  Per attempt: first bead and first particle, the cycle grown from rows of
 * the permutation table (SelectCycleIterative, :33-111; cycles longer than 8 particles are counted as not attempted),
 * PermuteBeads + the Levy bridge of every member level by level, the kinetic action along the links, every pair
 * action of the species in OLD and NEW mode over the listed particles, Metropolis per level starting from
 * -log(cycle weight), and on acceptance AssignParticleLabels (perm_bisect_class.h:47-56).  The path is kept by particle
 * LABEL plus the permutation at the beta seam (pimc_permutation_get): a chain keeps its label inside the path and
 * continues as next[label] across the seam; pair actions pair labels at equal slices, as the reference's GetBead(p, b)
 * does.  A cycle of one particle is the plain bisection with the links followed.  Philox slots are
 * documented in csrc/perm.cuh.  n_accept[n_clones], perm_attempt / perm_accept[n_clones][8] (by cycle length - 1;
 * the reference's perm_attempt / perm_accept vectors, perm_bisect_class.h:21-22) are host memory, ADDED to, may be NULL. */
int pimc_perm_bisect_sweep(pimc_ctx *ctx, int32_t species, int32_t n_level, int32_t n_attempts, uint64_t seed, uint64_t attempt0,
                           int32_t with_kinetic, double epsilon, int64_t *n_accept, int64_t *perm_attempt, int64_t *perm_accept);
/* The permutation at the beta seam, next[n_clones][N] (host memory): next[c][p] = label of the bead that follows
 * (p, n_bead - 1) of clone c -- the reference's Bead::next link of the last slice (species_class.h:302-321 writes the
 * same array as the `permutation` dataset of a path dump).  Identity until a permuting move or _set changes it.
 * While a species is permuted, the entry points that read one particle's path by label across the seam
 * (pimc_bisect_sweep, pimc_displace_sweep, a Kinetic handle's pimc_action_get) fail with PIMC_ERR_UNSUPPORTED;
 * whole-path evaluations follow the links (Kinetic) or do not depend on them (pair actions, g(r), S(k)). */
int pimc_permutation_get(pimc_ctx *ctx, int32_t species, int32_t *next);
int pimc_permutation_set(pimc_ctx *ctx, int32_t species, const int32_t *next);
/* Tests: the cycle of the LAST attempt of pimc_perm_bisect_sweep per clone -- window start, cycle length (0: the
 * selection stopped, -1: longer than 8), members[n_clones][8] (-1 padded), selection steps, accept flag.  Any may be NULL. */
int pimc_perm_last_cycle(pimc_ctx *ctx, int32_t *b0, int32_t *n_perm, int32_t *particles, int32_t *n_steps, int32_t *accept);

/* ---- slice sharding over several GPUs ---------------------------------------------------- */
/* One large path split by imaginary-time slice (pimc_config.slice_lo / slice_hi), one context and one
 * process per GPU.  No reference counterpart: the reference's MPI ranks are independent walkers
 * (framework_class.h:44-53).  What makes the split exact: the pair action at level 0 couples slice b
 * only with b + 1 (pair_action_class.h:282-288), rho_k(b) and the k sums are slice-local
 * (species_class.h:391-395, ilkka_pair_action_class.h:114-116).  The collectives run on NCCL (bound at
 * run time: libnccl.so.2 of the process or the system), issued on the context's stream. */
typedef struct pimc_comm pimc_comm;
/* Rank 0 creates the 128-byte NCCL unique id and hands it to the other ranks by whatever channel the
 * host program has (MPI_Bcast in the reference's world, a torch.distributed broadcast in bench.py). */
int pimc_comm_unique_id(void *id128);
/* Collective over the `world` ranks; world = 1 yields a communicator whose collectives are no-ops. */
int pimc_comm_init(pimc_ctx *ctx, const void *id128, int32_t rank, int32_t world, pimc_comm **out);
int pimc_comm_destroy(pimc_comm *comm);
int64_t pimc_comm_bytes_sent(pimc_comm *comm); /* payload bytes this rank has sent so far */
/* After this rank's positions changed: its first owned slice of `species` goes to the previous rank,
 * the next rank's first slice arrives in the halo slot (pack -> ncclSend/ncclRecv ring -> unpack). */
int pimc_halo_exchange(pimc_ctx *ctx, pimc_comm *comm, int32_t species);
/* In-place SUM over the ranks of n doubles in device memory (shard partial sums, estimator rows). */
int pimc_allreduce_sum(pimc_ctx *ctx, pimc_comm *comm, double *d_buf, int64_t n);
/* pimc_rotate_pack / ring / pimc_rotate_apply for every species, then halos and rho_k refreshed. */
int pimc_rotate(pimc_ctx *ctx, pimc_comm *comm, int32_t shift);
/* which = 0: whole-path action, 1: DActionDBeta, 2: Potential of n_actions actions of this context
 * into d_out[action][clone] (device memory), the shard partial sums combined by ONE all-reduce. */
int pimc_sharded_evaluate(pimc_ctx *ctx, pimc_comm *comm, int32_t which, pimc_action *const *actions, int32_t n_actions,
                          double *d_out);

/* ---- CUDA graphs -------------------------------------------------------------------------- */
/* Capture the device work of the calls made between begin and end on this context (kernels, copies
 * between device buffers, the NCCL collectives above) into a graph that replays with one launch.
 * Only entry points that neither synchronise nor take host output may be captured (the *_device
 * variants, pimc_rhok_rebuild, pimc_halo_exchange, pimc_allreduce_sum, pimc_sharded_evaluate); run the
 * sequence once before capturing it so that every scratch buffer exists. */
typedef struct pimc_graph pimc_graph;
int pimc_capture_begin(pimc_ctx *ctx);
int pimc_capture_end(pimc_ctx *ctx, pimc_graph **out);
int pimc_graph_launch(pimc_graph *graph);
int64_t pimc_graph_nodes(pimc_graph *graph);
int pimc_graph_destroy(pimc_graph *graph);

/* ---- estimators ------------------------------------------------------------------------ */
/* PairCorrelation::Accumulate (pair_correlation_class.h:15-28): y[c][i] += cofactor[c] for
 * every pair and slice, bin i = (uint32)nearbyint((|dr|-r_min)*d_ir - 0.5), i < n_r kept.
 * y is host memory [n_clones][n_r] and is ADDED to.  cofactor may be NULL (= 1). */
int pimc_est_gofr(pimc_ctx *ctx, int32_t species_a, int32_t species_b, double r_min, double r_max, int32_t n_r,
                  const double *cofactor, double *y);
/* Integer bin counts (bit-exact part): counts[c][i], host memory, overwritten. */
int pimc_est_gofr_counts(pimc_ctx *ctx, int32_t species_a, int32_t species_b, double r_min, double r_max, int32_t n_r,
                         uint64_t *counts);
/* StructureFactor::Accumulate (structure_factor_class.h:15-32): sk[c][k] += cofactor[c] *
 * sum_b Re(rho_a rho_b^*) for |k| < k_cut.  Host memory [n_clones][n_k], ADDED to. */
int pimc_est_sofk(pimc_ctx *ctx, int32_t species_a, int32_t species_b, double k_cut, const double *cofactor, double *sk);

/* ---- measurement helpers ------------------------------------------------------------- */
/* Number of kernels this library launched since the context was created. */
int64_t pimc_ctx_launch_count(pimc_ctx *ctx);
/* Per-kernel device time, measured with CUDA events on the context's stream around every
 * launch while timing is enabled (enable = 0/1 also clears the totals). */
enum {
    PIMC_KERNEL_PAIR_FULL = 1,   /* K1 */
    PIMC_KERNEL_RHOK_BUILD = 2,  /* K2 */
    PIMC_KERNEL_KSUM = 3,        /* K3 */
    PIMC_KERNEL_PAIR_WINDOW = 4, /* K4 */
    PIMC_KERNEL_GOFR = 5,        /* K5 */
    PIMC_KERNEL_SOFK = 6         /* K6 */
};
int pimc_ctx_set_timing(pimc_ctx *ctx, int32_t enable);
int pimc_ctx_kernel_time(pimc_ctx *ctx, int32_t kernel_id, double *total_ms, int64_t *n_launches);
/* FP64 FMA micro-benchmark on the context's device: returns achieved TFLOP/s (2 flop/FMA). */
int pimc_fp64_peak(pimc_ctx *ctx, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* SIMPIMC_B200_H_ */
