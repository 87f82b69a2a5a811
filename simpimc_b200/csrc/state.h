// Host-side state of libsimpimc_b200.so shared by its CUDA translation units (not part of the public ABI):
// capi.cu   context, k-space, table packing, actions, estimators, graphs, test hooks
// moves.cu  the device-resident moves (bisection sweeps, permuting bisection, DisplaceParticle) and the permutation table
// Every kernel header declares its kernels `static`, so each translation unit carries its own copies of the ones it launches.
#ifndef SIMPIMC_B200_STATE_H_
#define SIMPIMC_B200_STATE_H_

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/simpimc_b200.h"
#include "internal.h"
#include "kernels.cuh"
#include "pair_fast.cuh"
#include "kinetic.cuh"

namespace pimc_host {

using namespace pimc;


/// Records the pimc_last_error() text of the calling thread and returns `code` (defined in capi.cu).
int Fail(int code, const std::string &msg);

#define PIMC_CUDA(expr)                                                                                   \
    do {                                                                                                  \
        cudaError_t e__ = (expr);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            return Fail(PIMC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));              \
    } while (0)

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t Alloc(size_t count) {
        Free();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void **)&p, count * sizeof(T));
    }
    void Free() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { Free(); }
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

/// The FreeSplines of one (species, n_images): tau_s = tau 2^s / 2, s = 0..6, plus the tau derivative
/// table of tau itself (Kinetic::SetupSpline, kinetic_class.h:16-24; Bisect::SetupSpline, bisect_class.h:158-163).
struct FreeSet {
    DevBuf<double> pp[kMaxFreeSplines], pp_dtau;
    FreeSplineSet view;
    FreeSplineTab dtau;
};

struct SpeciesState {
    int N = 0;
    double lambda = 0.;
    std::map<int, std::unique_ptr<FreeSet>> free_sets;  // by n_images, built on first use
    int move_images = 0;                                 // Bisect's n_images attribute (bisect_class.h:173)
    pimc_action *kinetic = nullptr;                      // the species' Kinetic action, if one was created
    DevBuf<double> R;        // committed positions [C][N][3][Ms]
    // slice-major mirror [C][Mstore][N][3] for the bisection windows of the single-launch sweep (a 9-slice
    // window is 9 contiguous 6 KB blocks here, 768 72-byte fragments of 128-byte lines in R); built on
    // demand, kept in step by the sweep's own commits, invalidated by every other writer of R
    DevBuf<double> R2;
    bool R2_valid = false;
    DevBuf<double2> rho;     // committed rho_k [C][Mloc][n_k]
    // pending proposals: n_slots particles of the species per clone, n_prop beads each
    DevBuf<double> P;        // [slot][C][n_prop][3]
    DevBuf<int32_t> P_particle, P_first;  // [slot][C]
    int n_prop = 0, n_slots = 0;
    std::vector<int32_t> h_particle;      // host copy of P_particle (duplicate check)
    // rho_k increment of the proposal over the window it was computed for
    DevBuf<double2> drho;
    DevBuf<int32_t> drho_b0;
    int drho_window = 0;
    bool drho_valid = false;
    bool need_update_rho_k = true;  // Species::need_update_rho_k_ (species_class.h:25)
    // permutation at the beta seam: next[c][p] = label of the bead that follows (p, n_bead - 1);
    // allocated (as the identity) by the first permuting move or pimc_permutation_set, absent = unpermuted
    DevBuf<int32_t> perm_next;
    bool perm_tracked = false;
};

}  // namespace pimc_host

using namespace pimc_host;

struct pimc_ctx {
    int n_d = 3, pbc = 1, M = 0, C = 0, device = 0;
    double L = 0, iL = 0, vol = 1, beta = 0, tau = 0;
    int slice_lo = 0, slice_hi = 0, Mloc = 0, Mstore = 0, Ms = 0, sharded = 0;
    cudaStream_t stream = nullptr;
    int n_sm = 148;
    size_t smem_optin = 0;
    DevBuf<double2> rho_part;  // partial rho_k of a particle-split build
    std::vector<std::unique_ptr<SpeciesState>> species;
    // KSpace (k_space_class.h:6-15)
    double k_cutoff = 0.;
    int max_index = 0;
    std::vector<int32_t> k_index;  // [n_k][3] signed lattice indices
    std::vector<double> k_mag;
    DevBuf<int32_t> d_kidx;        // [n_k][3] offset by max_index
    DevBuf<double> d_kmag;
    // the same set grouped into (i_x, i_y) columns for rhok_build_cols_kernel
    DevBuf<int32_t> d_col_info, d_kmap;
    int n_cols = 0, cols_tm = 0;
    // scratch
    DevBuf<double> partial, out_dev, lr_dev, stage;
    DevBuf<int32_t> i32_a, i32_b, i32_c, i32_d;
    DevBuf<unsigned long long> counts;
    DevBuf<double> est;
    // device-resident moves (mc.cuh): per-clone scalars of the attempt in flight
    DevBuf<double> mc_f64;    // partial, logu0, pair_old, pair_new, lr_old, lr_new: 6 x [C]
    DevBuf<int32_t> mc_i32;   // alive, b0, accept: 3 x [C]
    DevBuf<long long> mc_naccept;
    // permuting bisection (perm.cuh): cycle of the attempt in flight, link sums per action and mode, label-rotation scratch
    DevBuf<int32_t> perm_i32;      // n_perm [C], n_steps [C], particles [C][kPermMaxLen]
    DevBuf<double> perm_f64;       // weight [C], lr_old [C], lr_new [C], pair_parts [n_actions][2][C][nb]
    DevBuf<double> perm_stage;     // [C][kPermMaxLen][3][M]
    DevBuf<long long> perm_counts; // attempted, accepted: 2 x [C][kPermMaxLen]
    std::vector<pimc_action *> actions;
    // side streams of pimc_internal_evaluate_many: the whole-path evaluations of several actions forked off the
    // context's stream and joined back into it (events), so that one action's CTAs fill the SMs another has left
    std::vector<cudaStream_t> side_streams;
    std::vector<cudaEvent_t> side_events;
    cudaEvent_t fork_event = nullptr;
    int64_t launches = 0;
    bool force_general = false;  // tests: evaluate with the general kernels even where the fast path applies
    // optional per-kernel device timing (CUDA events on the context's stream)
    bool timing = false;
    struct KTimer {
        std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
        double total_ms = 0.;
        int64_t n = 0;
    } timers[8];

    int n_k() const { return (int)k_mag.size(); }
    PathView View() const {
        PathView v;
        v.C = C;
        v.M = M;
        v.Mloc = Mloc;
        v.Mstore = Mstore;
        v.Ms = Ms;
        v.slice_lo = slice_lo;
        v.sharded = sharded;
        v.vdiv = 1;
        v.box.L = L;
        v.box.iL = iL;
        return v;
    }
    SpeciesView SView(int s, bool with_proposal) const {
        const SpeciesState &st = *species[s];
        SpeciesView v;
        v.R = st.R.p;
        v.N = st.N;
        v.P = st.P.p;
        v.P_particle = st.P_particle.p;
        v.P_first = st.P_first.p;
        v.n_prop = with_proposal ? st.n_prop : 0;
        v.n_slots = with_proposal ? st.n_slots : 0;
        return v;
    }
    KSpaceView KView() const {
        KSpaceView k;
        k.n_k = n_k();
        k.max_index = max_index;
        k.kidx = d_kidx.p;
        k.kbox = 2. * M_PI / L;
        return k;
    }
};

struct pimc_action {
    pimc_ctx *ctx = nullptr;
    int atype = ATYPE_ILKKA;
    int sa = 0, sb = 0;
    int max_level = 0;
    bool use_long_range = false;
    bool is_constant = false;
    int n_images = 0;  // Kinetic: periodic images of the free-particle density matrix (action_class.h:30)
    // per evaluated quantity (U, dU, V): the stageable blob (grids, LUTs, 1-D pp coefficients;
    // for David the B-spline multi-spline, read from global memory), the 2-D cell polynomials
    DevBuf<double> blob[3];
    DevBuf<double> cells[3];
    bool stageable[3] = {true, true, true};
    PairTable table[3];
    // long range: weight per k vector and scaled constants
    DevBuf<double> wk[3];
    // host copies of the |k|-shell tables the weights were matched from: KSpace::Setup is grow-only
    // (k_space_class.h:34-41), so a later, larger k_cut rebuilds the vector list and every action's
    // weights have to be re-matched against it (RematchWeights)
    std::vector<double> shell_k[3], shell_f[3];
    double k0[3] = {0, 0, 0}, r0[3] = {0, 0, 0};
    double ulong_scale = 1.;  // Bare CalcULong: level_tau
    // the action's own partial-sum scratch for evaluations that run beside other actions' (pimc_internal_evaluate_many)
    DevBuf<double> own_partial, own_lr;
    // Ilkka U / dU fast path (pair_fast.cuh): every table in one shared-memory block
    DevBuf<unsigned char> fast_tab[2];
    FastTable fast[2];
    bool fast_ok[2] = {false, false};
    // fast Potential() kernel (Ilkka / Bare): v(r) and v_long(r) in the shared-memory layout
    DevBuf<unsigned char> fastv_tab;
    FastVTable fastv;
    bool fastv_ok = false;
    // David U / dU fast path: endpoint spline and off-diagonal multi-spline in shared memory
    DevBuf<unsigned char> fastd_tab[2];
    FastDavidTable fastd[2];
    bool fastd_ok[2] = {false, false};
};

struct pimc_graph {
    pimc_ctx *ctx = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int64_t n_nodes = 0;
};

namespace pimc_host {

/// Brackets one kernel launch with CUDA events when timing is enabled.
struct ScopedKernelTimer {
    pimc_ctx *ctx;
    int id;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ScopedKernelTimer(pimc_ctx *c, int kernel_id);
    ~ScopedKernelTimer();
};

// helpers defined in capi.cu
int EnsureI32(pimc_ctx *ctx, DevBuf<int32_t> &buf, const int32_t *host, size_t n);  ///< host int32 array -> device scratch
int GridFor(const pimc_ctx *ctx, int items);
int GetFreeSet(pimc_ctx *ctx, int s, int n_images, FreeSet **out);                  ///< FreeSplines of (species, n_images), built on first use
int ToHost(pimc_ctx *ctx, const double *d_src, double *host, size_t n);             ///< device -> host on the context's stream, synchronised
// defined in moves.cu
/// PIMC_OK unless the species' path is permuted at the beta seam (entry points that read one particle's path by label).
int RequireUnpermuted(pimc_ctx *ctx, int s, const char *what);

}  // namespace pimc_host

#endif  // SIMPIMC_B200_STATE_H_
