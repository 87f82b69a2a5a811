// Spatial derivatives of a pair action over a window of slices:
//   PairAction::GetActionGradient   src/actions/pair_action/pair_action_class.h:305-337
//   PairAction::GetActionLaplacian  pair_action_class.h:340-366
// called by the contact-density and virial estimators (src/events/observables/
// contact_density_class.h:120-121, energy_class.h:88).  For every pair that touches a moved
// particle and every slice b_i of the window the reference adds the derivative, with respect to
// the bead (species a, first particle of the pair, b_i), of CalcU on the forward link
// (b_i, b_i + 1) and on the backward link (b_i, b_i - 1):
//   * gradient, Ilkka: analytic -- chain rule through x = q + s/2, y = q - s/2 on the 2-D spline's
//     value-and-gradient evaluation, minus the long-range r-space slope
//     (ilkka_pair_action_class.h:172-222);
//   * gradient, Bare / David: central differences of CalcU, eps = 1e-4 (pair_action_class.h:135-155);
//   * Laplacian, all types: second central differences, eps = 1e-4 (pair_action_class.h:168-203);
//   * Ilkka long-range gradient: sum_k u_k k (Re rho_bead Im rho_b - Im rho_bead Re rho_b), added
//     once per PAIR for the pair's species-a particle (pair_action_class.h:330-334 calling
//     ilkka_pair_action_class.h:231-249) -- linear in rho_bead, so the kernel first accumulates
//     A_k = sum_pairs rho_bead(first of the pair) and contracts once.
#ifndef SIMPIMC_B200_GRAD_CUH_
#define SIMPIMC_B200_GRAD_CUH_

#include "kernels.cuh"

namespace pimc {

constexpr double kGradEps = 1.e-4;  // pair_action_class.h:141,175

/// Value and slope of a pp-form 1-D spline.
__device__ __forceinline__ void PP1EvalVG(const double *__restrict__ blob, const PP1Desc &d, double x, double &val, double &grad) {
    const double *g = blob + d.off_g;
    const int i = LutInterval(blob, d.lut, g, d.n, x);
    const double t = x - g[i];
    const double2 *c = reinterpret_cast<const double2 *>(blob + d.off_pp) + 2 * i;
    const double2 c01 = c[0], c23 = c[1];
    val = fma(fma(fma(c23.y, t, c23.x), t, c01.y), t, c01.x);
    grad = fma(fma(3. * c23.y, t, 2. * c23.x), t, c01.y);
}

/// Value and gradient of a pp-form 2-D spline (eval_NUBspline_2d_d_vg).
__device__ __forceinline__ void PP2EvalVG(const double *__restrict__ blob, const PP2Desc &d, double x, double y, double &val, double &gx,
                                          double &gy) {
    const double *gxs = blob + d.off_gx, *gys = blob + d.off_gy;
    const int ix = LutInterval(blob, d.lutx, gxs, d.nx, x);
    const int iy = LutInterval(blob, d.luty, gys, d.ny, y);
    const double tx = x - gxs[ix], ty = y - gys[iy];
    const double2 *c = reinterpret_cast<const double2 *>(d.cells) + ((size_t)ix * d.ny + iy) * 8;
    double row[4], drow[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const double2 c01 = __ldg(c + 2 * m), c23 = __ldg(c + 2 * m + 1);
        row[m] = fma(fma(fma(c23.y, ty, c23.x), ty, c01.y), ty, c01.x);
        drow[m] = fma(fma(3. * c23.y, ty, 2. * c23.x), ty, c01.y);
    }
    val = fma(fma(fma(row[3], tx, row[2]), tx, row[1]), tx, row[0]);
    gx = fma(fma(3. * row[3], tx, 2. * row[2]), tx, row[1]);
    gy = fma(fma(fma(drow[3], tx, drow[2]), tx, drow[1]), tx, drow[0]);
}

/// Vector-keeping Path::DrDrpDrrp (path_class.h:153-166).
__device__ __forceinline__ void DrDrpDrrpVec(const double a0[3], const double b0[3], const double a1[3], const double b1[3], const Box &bx,
                                             double r[3], double rrp[3], double &r_mag, double &rp_mag, double &rrp_mag) {
    double rp[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        r[d] = b0[d] - a0[d];
        rp[d] = b1[d] - a1[d];
        r[d] -= rint(r[d] * bx.iL) * bx.L;
        rp[d] += rint((r[d] - rp[d]) * bx.iL) * bx.L;
        rrp[d] = r[d] - rp[d];
        rrp[d] -= rint(rrp[d] * bx.iL) * bx.L;
    }
    r_mag = Mag3(r[0], r[1], r[2]);
    rp_mag = Mag3(rp[0], rp[1], rp[2]);
    rrp_mag = Mag3(rrp[0], rrp[1], rrp[2]);
}

/// CalcGradientU of one link: a0/a1 = the pair's species-a particle at b_i and at the link's
/// other slice, q0/q1 = its partner.  Adds to g[3].
template <int ATYPE>
__device__ __forceinline__ void LinkGradient(const double *__restrict__ blob, const PairTable &T, const double a0[3], const double q0[3],
                                             const double a1[3], const double q1[3], const Box &bx, double g[3]) {
    if (ATYPE == ATYPE_ILKKA) {
        double r[3], rrp[3], r_mag, rp_mag, s_mag;
        DrDrpDrrpVec(a0, q0, a1, q1, bx, r, rrp, r_mag, rp_mag, s_mag);
        const double q = 0.5 * (r_mag + rp_mag);
        double u, gx, gy;
        PP2EvalVG(blob, T.xy, q + 0.5 * s_mag, q - 0.5 * s_mag, u, gx, gy);
        double rh[3], sh[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            rh[d] = r_mag == 0. ? 0. : r[d] / r_mag;
            sh[d] = s_mag == 0. ? 0. : rrp[d] / s_mag;
        }
        double slope = 0.;
        if (T.use_lr) {
            SetLimits(T.lr.r_min, T.lr.r_max, r_mag, rp_mag);
            double v;
            PP1EvalVG(blob, T.lr, r_mag, v, slope);
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double t = -0.5 * (gx * (rh[d] + sh[d]) + gy * (rh[d] - sh[d]));
            t -= 0.5 * slope * rh[d];
            g[d] += t;
        }
    } else {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double ap[3] = {a0[0], a0[1], a0[2]};
            double r, rp, s;
            ap[d] = a0[d] + kGradEps;
            DrDrpDrrp(ap, q0, a1, q1, bx, r, rp, s);
            const double f1 = PairEval<ATYPE, WHICH_U>(blob, T, r, rp, s);
            ap[d] = a0[d] - kGradEps;
            DrDrpDrrp(ap, q0, a1, q1, bx, r, rp, s);
            const double f2 = PairEval<ATYPE, WHICH_U>(blob, T, r, rp, s);
            g[d] += (f1 - f2) / (2. * kGradEps);
        }
    }
}

/// CalcLaplacianU of one link.
template <int ATYPE>
__device__ __forceinline__ double LinkLaplacian(const double *__restrict__ blob, const PairTable &T, const double a0[3], const double q0[3],
                                                const double a1[3], const double q1[3], const Box &bx) {
    double r, rp, s, tot = 0.;
    DrDrpDrrp(a0, q0, a1, q1, bx, r, rp, s);
    const double f0 = PairEval<ATYPE, WHICH_U>(blob, T, r, rp, s);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double ap[3] = {a0[0], a0[1], a0[2]};
        ap[d] = a0[d] + kGradEps;
        DrDrpDrrp(ap, q0, a1, q1, bx, r, rp, s);
        const double fp1 = PairEval<ATYPE, WHICH_U>(blob, T, r, rp, s);
        ap[d] = a0[d] - kGradEps;
        DrDrpDrrp(ap, q0, a1, q1, bx, r, rp, s);
        const double fm1 = PairEval<ATYPE, WHICH_U>(blob, T, r, rp, s);
        tot += (fp1 + fm1 - 2 * f0) / (kGradEps * kGradEps);
    }
    return tot;
}

struct PairGradArgs {
    PathView pv;
    SpeciesView A, B;
    int same;
    int moved_a, moved_b;
    const int32_t *part_a, *part_b, *b0;  // [C]
    int n_links;
    int what;                             // 0 = gradient (3 values per item), 1 = Laplacian (1 value)
    PairTable T;
    const double *blob;
    double *partial;                      // [C][n_links][what ? 1 : 3]
};

/// One CTA per (clone, window slice): the pairs of pair_window_kernel, derivative with respect
/// to the species-a bead of each pair on its forward and backward links.
template <int ATYPE>
static __global__ void __launch_bounds__(128) pair_grad_kernel(const PairGradArgs a) {
    __shared__ double red[128 / 32];
    const PathView &pv = a.pv;
    const int nv = a.what ? 1 : 3;
    for (int item = blockIdx.x; item < pv.C * a.n_links; item += gridDim.x) {
        const int c = item / a.n_links, j = item - c * a.n_links;
        int bg = a.b0[c] + j;
        if (bg >= pv.M) bg -= pv.M;
        const int bn = bg + 1 == pv.M ? 0 : bg + 1, bp = bg == 0 ? pv.M - 1 : bg - 1;
        double acc[3] = {0., 0., 0.};
        // every (first = a particle, second = b particle) pair the reference lists, visited through its a particle
        const int ma = a.moved_a ? a.part_a[c] : -1, mb = (!a.same && a.moved_b) ? a.part_b[c] : -1;
        const int n_t = max(ma >= 0 ? a.B.N : 0, mb >= 0 ? a.A.N : 0);
        for (int t = threadIdx.x; t < n_t; t += blockDim.x) {
            // moved a with every partner t of species b; or (b moved) first = t with the moved b particle
            for (int pass = 0; pass < 2; ++pass) {
                int first, second;
                if (pass == 0) {
                    if (ma < 0 || t >= a.B.N) continue;
                    first = ma;
                    second = t;
                    if (a.same && second == ma) continue;
                    if (mb >= 0 && second == mb) continue;  // (moved a, moved b) is visited in pass 1
                } else {
                    if (mb < 0 || t >= a.A.N) continue;
                    first = t;
                    second = mb;
                }
                double f0[3], fn[3], fp[3], s0[3], sn[3], sp[3];
                LoadPos(pv, a.A, c, first, bg, 0, f0);
                LoadPos(pv, a.A, c, first, bn, 0, fn);
                LoadPos(pv, a.A, c, first, bp, 0, fp);
                LoadPos(pv, a.B, c, second, bg, 0, s0);
                LoadPos(pv, a.B, c, second, bn, 0, sn);
                LoadPos(pv, a.B, c, second, bp, 0, sp);
                if (a.what == 0) {
                    LinkGradient<ATYPE>(a.blob, a.T, f0, s0, fn, sn, pv.box, acc);
                    LinkGradient<ATYPE>(a.blob, a.T, f0, s0, fp, sp, pv.box, acc);
                } else {
                    acc[0] += LinkLaplacian<ATYPE>(a.blob, a.T, f0, s0, fn, sn, pv.box) + LinkLaplacian<ATYPE>(a.blob, a.T, f0, s0, fp, sp, pv.box);
                }
            }
        }
        for (int v = 0; v < nv; ++v) {
            const double tot = BlockSum<128>(acc[v], red);
            if (threadIdx.x == 0) a.partial[(size_t)item * nv + v] = tot;
            __syncthreads();
        }
    }
}

struct GradLongArgs {
    PathView pv;
    SpeciesView A;
    KSpaceView ks;
    const double2 *rho_b;      // [C][Mloc][n_k] of species b
    const double *wk;          // [n_k] u_long_k per k vector
    const int32_t *part_a;     // [C] moved a particle (or null)
    const int32_t *b0;
    int n_links;
    int mult_moved;            // pairs whose first particle is the moved a particle
    int mult_others;           // 1 if every other a particle is the first of one pair (a b particle moved)
    double factor;             // 2 if the species differ
    double *out;               // [C][3]
};

/// CalcGradientULong summed over the pairs: one CTA per clone, threads over k.
static __global__ void __launch_bounds__(256) grad_long_kernel(const GradLongArgs a) {
    extern __shared__ __align__(16) double2 gtab[];  // [32 particles][3 axes][2 m + 1]
    __shared__ double red[256 / 32];
    const PathView &pv = a.pv;
    const int tl = 2 * a.ks.max_index + 1;
    const int c = blockIdx.x;
    const int ma = a.part_a ? a.part_a[c] : -1;
    double g[3] = {0., 0., 0.};
    for (int j = 0; j < a.n_links; ++j) {
        int bg = a.b0[c] + j;
        if (bg >= pv.M) bg -= pv.M;
        const int p_lo = a.mult_others ? 0 : ma, p_hi = a.mult_others ? a.A.N : ma + 1;
        for (int p0 = p_lo; p0 < p_hi; p0 += 32) {
            const int np = min(32, p_hi - p0);
            __syncthreads();
            for (int t = threadIdx.x; t < np * 3; t += blockDim.x) {
                const int pp = t / 3, d = t - pp * 3;
                double r[3];
                LoadPos(pv, a.A, c, p0 + pp, bg, 0, r);
                PhaseTable(r[d], a.ks.kbox, a.ks.max_index, gtab + (size_t)t * tl);
            }
            __syncthreads();
            for (int k = threadIdx.x; k < a.ks.n_k; k += blockDim.x) {
                const int i0 = a.ks.kidx[3 * k], i1 = a.ks.kidx[3 * k + 1], i2 = a.ks.kidx[3 * k + 2];
                double ar = 0., ai = 0.;
                for (int pp = 0; pp < np; ++pp) {
                    const double m = (p0 + pp == ma) ? (double)a.mult_moved : (double)a.mult_others;
                    const double2 *tb = gtab + (size_t)pp * 3 * tl;
                    const double2 f = CMul(CMul(tb[i0], tb[tl + i1]), tb[2 * tl + i2]);
                    ar += m * f.x;
                    ai += m * f.y;
                }
                const double2 rb = a.rho_b[((size_t)c * pv.Mloc + (bg - pv.slice_lo)) * a.ks.n_k + k];
                const double f = a.wk[k] * (ar * rb.y - ai * rb.x);
                g[0] += (double)(i0 - a.ks.max_index) * a.ks.kbox * f;
                g[1] += (double)(i1 - a.ks.max_index) * a.ks.kbox * f;
                g[2] += (double)(i2 - a.ks.max_index) * a.ks.kbox * f;
            }
        }
    }
    for (int d = 0; d < 3; ++d) {
        const double tot = BlockSum<256>(g[d], red);
        if (threadIdx.x == 0) a.out[(size_t)c * 3 + d] = a.factor * tot;
        __syncthreads();
    }
}

/// out[c][v] = sum over the window items in order (+ the long-range gradient).
static __global__ void grad_finalize_kernel(const double *__restrict__ partial, int C, int n_links, int nv, const double *__restrict__ lr, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * nv) return;
    const int c = i / nv, v = i - c * nv;
    double tot = 0.;
    for (int j = 0; j < n_links; ++j) tot += partial[((size_t)c * n_links + j) * nv + v];
    if (lr) tot += lr[(size_t)c * nv + v];
    out[i] = tot;
}

}  // namespace pimc

#endif  // SIMPIMC_B200_GRAD_CUH_
