// Free-particle density matrix with periodic images on the device: FreeSpline
// (src/actions/free_spline_class.h:25-84) and the Kinetic action built on it
// (src/actions/single_action/kinetic_class.h:35-45 DActionDBeta, :105-122 GetAction).
//
// FreeSpline tabulates, on a uniform grid of 10 000 points over [-L/2, L/2],
//   image_action(r)  = -log1p( sum_{i=1..n_images} exp((r^2 - (r + iL)^2)/4 lambda tau) + exp((r^2 - (r - iL)^2)/4 lambda tau) )
// (and its tau derivative) and interpolates them with a natural cubic spline per dimension:
//   log rho_free(r)      = -sum_d image_action(r_d) - |r|^2 / (4 lambda tau)
//   dlog rho_free / dtau = -sum_d d_image_action_d_tau(r_d) - |r|^2 / (4 lambda tau^2)
// The tables are converted to one cubic per interval on the host (spline_build.h:
// BuildFreeSpline -- the same natural interpolant, pp form) and live in global memory (320 KB
// each, L2-resident).  Where the thermal wavelength is small against the box the image sum
// underflows to exactly zero over the whole centre of the grid; the run of all-zero intervals
// is recorded and lookups that land in it touch no memory.  n_images = 0 has no table at all:
// the closed form -|r|^2 / (4 lambda tau) the reference's spline of zeros reduces to.
#ifndef SIMPIMC_B200_KINETIC_CUH_
#define SIMPIMC_B200_KINETIC_CUH_

#include "kernels.cuh"

namespace pimc {

struct FreeSplineTab {
    const double *pp;       // [n_int][4]: cubic in (x - g_i), g_i = start + i dr; nullptr: identically zero (n_images = 0)
    int n_int;              // intervals (grid points - 1)
    int zero_lo, zero_hi;   // intervals [zero_lo, zero_hi) are identically zero
    double start, inv_dr, dr;
    double i4lt;            // 1 / (4 lambda tau_s)        (dtau table: 1 / (4 lambda tau_s^2))
};

constexpr int kMaxFreeSplines = 7;  // tau_s = tau / 2, tau, 2 tau, ... 32 tau: sampling of level l = s, Kinetic of level l = s - 1

struct FreeSplineSet {
    FreeSplineTab s[kMaxFreeSplines];
    int n_images;  // 0: every table is nullptr, closed forms
};

/// eval_UBspline_1d_d of the image table at x in [-L/2, L/2] (free_spline_class.h:79).
__device__ __forceinline__ double FreeImageAction(const FreeSplineTab &T, double x) {
    int i = (int)floor((x - T.start) * T.inv_dr);
    i = min(max(i, 0), T.n_int - 1);
    if (i >= T.zero_lo && i < T.zero_hi) return 0.;
    const double t = x - fma((double)i, T.dr, T.start);
    const double2 *p = reinterpret_cast<const double2 *>(T.pp) + 2 * (size_t)i;
    const double2 c01 = __ldg(p), c23 = __ldg(p + 1);
    return fma(fma(fma(c23.y, t, c23.x), t, c01.y), t, c01.x);
}

/// FreeSpline::GetLogRhoFree (free_spline_class.h:75-83); also GetDLogRhoFreeDTau (:91-99) when T
/// is the tau-derivative table (its i4lt is 1 / (4 lambda tau^2)).
__device__ __forceinline__ double FreeLogRho(const FreeSplineTab &T, const double r[3]) {
    double tot = 0.;
    if (T.pp) {
#pragma unroll
        for (int d = 0; d < 3; ++d) tot -= FreeImageAction(T, r[d]);
    }
    return tot - ((r[0] * r[0] + r[1] * r[1]) + r[2] * r[2]) * T.i4lt;
}

// ------------------------------------------------------------------ Kinetic::DActionDBeta
struct KineticFullArgs {
    PathView pv;
    SpeciesView sv;
    FreeSplineTab dtau;   // d_image_action_d_tau of tau (rho_free_splines[0], kinetic_class.h:20)
    int n_chunks;         // chunks of 32 links
    double *partial;      // [C][n_chunks]
    const int32_t *next = nullptr;  // [C][N] permutation at the beta seam (nullptr: identity): the bead after (p, M - 1) is (next[p], 0)
};

/// sum over particles and links (b, b + 1) of GetDLogRhoFreeDTau(Dr(bead, next)) (kinetic_class.h:38-43);
/// the constant N M n_d / (2 tau) is added by the caller.  CTA = (clone, 32-link chunk), lanes = links.
static __global__ void __launch_bounds__(256) kinetic_dbeta_kernel(const KineticFullArgs a) {
    __shared__ double red[256 / 32];
    const PathView &pv = a.pv;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int item = blockIdx.x; item < pv.C * a.n_chunks; item += gridDim.x) {
        const int c = item / a.n_chunks, chunk = item - c * a.n_chunks;
        const int b = chunk * 32 + lane;
        double acc = 0.;
        if (b < pv.Mloc) {
            const int bn = NextSlice(pv, b);
            const bool seam = a.next && !pv.sharded && b + 1 == pv.M;
            for (int p = warp; p < a.sv.N; p += 256 / 32) {
                const int pn = seam ? a.next[(size_t)c * a.sv.N + p] : p;
                double dr[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const double x = a.sv.R[PosIndex(pv, a.sv.N, c, p, d, b)] - a.sv.R[PosIndex(pv, a.sv.N, c, pn, d, bn)];
                    dr[d] = pv.box.L > 0. ? x - rint(x * pv.box.iL) * pv.box.L : x;   // Path::Dr (path_class.h:108-112)
                }
                acc += FreeLogRho(a.dtau, dr);
            }
        }
        const double tot = BlockSum<256>(acc, red);
        if (threadIdx.x == 0) a.partial[item] = tot;
        __syncthreads();
    }
}

/// out[c] = constant + sign * sum of the clone's chunk partial sums, in chunk order.
static __global__ void kinetic_finalize_kernel(const double *__restrict__ partial, int C, int n_chunks, double sign, double constant,
                                        double *__restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double tot = 0.;
    for (int i = 0; i < n_chunks; ++i) tot += partial[(size_t)c * n_chunks + i];
    out[c] = constant + sign * tot;
}

// --------------------------------------------------------------------- Kinetic::GetAction
struct KineticWindowArgs {
    PathView pv;
    SpeciesView sv;           // with the proposal overlay in NEW mode
    FreeSplineTab tab;        // rho_free_splines[skip - 1]: tau_s = tau * 2^level
    int n_listed;             // listed particles of this species per clone (<= kMaxPropSlots)
    const int32_t *part;      // [n_listed][C]
    const int32_t *b0;        // [C]
    int n_links;              // links of stride `skip` in the window: (b1 - b0) / skip
    int skip;
    int mode;
    double *out;              // [C]
};

/// -sum over listed particles and links (a, a + skip), a = b0, b0 + skip, ... < b1, of
/// GetLogRhoFree(Dr(bead_a, bead_{a + skip})) (kinetic_class.h:105-122).  One warp per clone.
static __global__ void __launch_bounds__(128) kinetic_window_kernel(const KineticWindowArgs a) {
    const PathView &pv = a.pv;
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= pv.C) return;
    double acc = 0.;
    const int n_items = a.n_listed * a.n_links;
    for (int t = lane; t < n_items; t += 32) {
        const int i = t / a.n_links, j = t - i * a.n_links;
        const int p = a.part[(size_t)i * pv.C + c];
        const int bg = a.b0[c] + j * a.skip;
        double x0[3], x1[3], dr[3];
        LoadPos(pv, a.sv, c, p, bg, a.mode, x0);
        LoadPos(pv, a.sv, c, p, bg + a.skip, a.mode, x1);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double x = x0[d] - x1[d];
            dr[d] = pv.box.L > 0. ? x - rint(x * pv.box.iL) * pv.box.L : x;
        }
        acc -= FreeLogRho(a.tab, dr);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) a.out[c] = acc;
}

}  // namespace pimc

#endif  // SIMPIMC_B200_KINETIC_CUH_
