// Device-side arithmetic of the pair-action path: minimum-image distance triples and
// non-uniform cubic B-spline evaluation on tables staged in shared (or left in global) memory.
//
// Reference semantics being reproduced (all /root/reference/src):
//   Path::DrDrpDrrp          data_structures/path_class.h:137-150
//   Path::Dr / PutInBox      data_structures/path_class.h:108-134
//   PairAction::SetLimits    actions/pair_action/pair_action_class.h:32-42
//   Ilkka CalcU/CalcdUdBeta/CalcV   actions/pair_action/ilkka_pair_action_class.h:77-101,125-149,34-54
//   Bare  CalcV/CalcU               actions/pair_action/bare_pair_action_class.h:99-123,146-150
//   David CalcU/CalcdUdBeta/CalcV   actions/pair_action/david_pair_action_class.h:63-102,126-168,26-39
// The spline evaluation follows the published einspline algorithm the reference calls
// (eval_NUBspline_{1d,2d}_d, eval_multi_NUBspline_1d_d): interval search on the knot grid,
// Cox-de Boor recursion for the four live basis functions, 4 / 16-tap contraction.
#ifndef SIMPIMC_B200_DEVICE_MATH_CUH_
#define SIMPIMC_B200_DEVICE_MATH_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

namespace pimc {

enum { ATYPE_ILKKA = 0, ATYPE_BARE = 1, ATYPE_DAVID = 2, ATYPE_KINETIC = 3 };
enum { WHICH_U = 0, WHICH_DU = 1, WHICH_V = 2 };
enum { GRIDCODE_GENERAL = 0, GRIDCODE_LOG = 1 };

struct Box {
    double L, iL;
};

/// Offsets (in doubles) of one B-spline-form 1-D / multi spline inside a table blob.
struct Sp1Desc {
    int n;            // grid points
    int off_t, off_w, off_c;
    int n_splines;    // 1 for a plain spline; stride of the multi-spline coefficient rows
    int code;         // GRIDCODE_*
    double r_min, r_max;  // grid start / end (SetLimits)
    double ainv, startinv;  // log grid reverse map
};
/// Interval-search accelerator: key = (bits(x) >> shift) - key0 clamped to [0, n_keys);
/// lut[key] (uint16, `off` counted in uint16 units from the blob start) is a lower bound of
/// the interval index, finished by a forward scan on the grid.
struct LutDesc {
    int off, n_keys, shift;
    long long key0;
};
/// pp-form 1-D spline inside the (stageable) blob: grid g[n], coefficients pp[n][4].
struct PP1Desc {
    int n, off_g, off_pp;
    LutDesc lut;
    double r_min, r_max;
};
/// pp-form 2-D spline: grids and LUTs inside the blob, the cell polynomials
/// cells[ix][iy][4][4] in global memory (read through the read-only path).
struct PP2Desc {
    int nx, ny, off_gx, off_gy;
    LutDesc lutx, luty;
    const double *cells;
};
/// Interval table of the David grid.  kind 0: uniform buckets (ULookup's key).  kind 1: buckets of
/// the IEEE-754 bit pattern, key = (high word of x >> shift) - key0 -- the exponent and the top
/// 20 - shift mantissa bits, i.e. buckets of constant RELATIVE width, which is what a
/// logarithmic grid (the David squarer's usual choice) needs: a uniform table fine enough for
/// its first interval would not fit.  Either way lut[key] is the interval of the bucket's lower
/// edge and at most one knot lies inside a bucket.
struct FastLut {
    int kind;
    int off_lut, key_max;
    int shift, key0;
    double inv_h;
};
/// DavidPairAction tables in the shared-memory layout: the endpoint spline e(r) (value 1 of the
/// multi-spline for U; value 0 + value 1 for dU/dbeta, david...:137-145) as split (c0,c1)/(c2,c3)
/// arrays, and one record per interval holding the pp coefficients of the n_q = n_val - 1
/// off-diagonal values u_kj(q) (32 bytes each, record stride an odd number of 16-byte slots).
struct FastDavidTable {
    FastLut lut;
    int off_gpair;        // double2 [n]: (g[i], g[i+1]); g[n] = +inf
    int off_e01, off_e23; // double2 [n] each
    int off_q, q_stride;  // bytes
    int n_order;
    double r_min, r_max;
    int n_bytes;
};

/// Everything a pair kernel needs to evaluate one of U / dU/dbeta / V of one action.
struct PairTable {
    int use_lr;
    int is_coulomb;
    int n_order;
    double u_scale;      // Bare CalcU: (1 >> level) * tau
    PP2Desc xy;          // Ilkka u_xy / du_xy
    PP1Desc a;           // Ilkka / Bare v_r
    PP1Desc lr;          // long-range r-space spline subtracted from the short-range part
    Sp1Desc dav;         // David multi-spline (B-spline form) ...
    const double *dav_blob;  // ... in global memory
    // David U / dU: the pp-form block of the fast whole-path kernel, read through the read-only
    // path by every other kernel (windows, displacements, per-pair hooks) unless the context
    // forces the general B-spline evaluation
    int dav_use_fast;
    const unsigned char *dav_fast_tab;
    FastDavidTable dav_fast;
};

// nearbyint() in the default rounding mode is round-half-even == rint()
__device__ __forceinline__ double MinImage(double d, const Box &bx) { return d - rint(d * bx.iL) * bx.L; }

// scaffold mag() = arma::norm(v,2) for 3 elements accumulates (x0^2 + x2^2) + x1^2
__device__ __forceinline__ double Mag3(double x, double y, double z) { return sqrt((x * x + z * z) + y * y); }

// Same without FMA contraction and with correctly rounded sqrt: g(r) bin indices depend on it.
__device__ __forceinline__ double Mag3Exact(double x, double y, double z) {
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(z, z)), __dmul_rn(y, y)));
}

/// Path::DrDrpDrrp: r between the two particles at slice b0 (minimum image), r' at slice b1
/// moved to the SAME image as r, and |r - r'| minimum-imaged.
__device__ __forceinline__ void DrDrpDrrp(const double a0[3], const double b0[3], const double a1[3], const double b1[3],
                                          const Box &bx, double &r_mag, double &rp_mag, double &rrp_mag) {
    double r[3], rp[3], rrp[3];
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

__device__ __forceinline__ void SetLimits(double r_min, double r_max, double &r, double &r_p) {
    r = r > r_max ? r_max : (r < r_min ? r_min : r);
    r_p = r_p > r_max ? r_max : (r_p < r_min ? r_min : r_p);
}

/// Interval index of x on the grid g[0..n) (einspline general_grid_reverse_map): last i with
/// g[i] <= x, 0 below the grid, n-1 at or above its end.
__device__ __forceinline__ int GridInterval(const double *__restrict__ g, int n, double x) {
    if (x <= g[0]) return 0;
    if (x >= g[n - 1]) return n - 1;
    int lo = 0, hi = n - 1;
    while (hi - lo >= 2) {
        int mid = (hi + lo) >> 1;
        if (g[mid] > x)
            hi = mid;
        else
            lo = mid;
    }
    return lo;
}

/// Cox-de Boor recursion: the four cubic B-spline basis values alive on interval i at x.
/// t = knots (grid at t+2), w[3i+j] = 1/(t[i+j+1]-t[i]).
__device__ __forceinline__ void BasisOnInterval(const double *__restrict__ t, const double *__restrict__ w, int i, double x,
                                                double b[4]) {
    const int i2 = i + 2;
    const double tm2 = t[i2 - 2], tm1 = t[i2 - 1], t0 = t[i2], t1 = t[i2 + 1], t2 = t[i2 + 2], t3 = t[i2 + 3];
    const double w00 = w[3 * i + 2];
    const double w11 = w[3 * (i + 1) + 1], w12 = w[3 * (i + 1) + 2];
    const double w20 = w[3 * (i + 2) + 0], w21 = w[3 * (i + 2) + 1], w22 = w[3 * (i + 2) + 2];
    const double l0 = (t1 - x) * w20;
    const double l1 = (x - t0) * w20;
    const double q0 = (t1 - x) * w11 * l0;
    const double q1 = ((x - tm1) * w11 * l0 + (t2 - x) * w21 * l1);
    const double q2 = (x - t0) * w21 * l1;
    b[0] = (t1 - x) * w00 * q0;
    b[1] = ((x - tm2) * w00 * q0 + (t2 - x) * w12 * q1);
    b[2] = ((x - tm1) * w12 * q1 + (t3 - x) * w22 * q2);
    b[3] = (x - t0) * w22 * q2;
}

__device__ __forceinline__ int Sp1Interval(const double *__restrict__ blob, const Sp1Desc &s, double x) {
    if (s.code == GRIDCODE_LOG) {
        int idx = (int)floor(s.ainv * log(x * s.startinv));
        idx = idx < 0 ? 0 : idx;
        return idx > s.n - 1 ? s.n - 1 : idx;
    }
    return GridInterval(blob + s.off_t + 2, s.n, x);
}

/// One value of a multi-spline (coefficient rows of n_splines doubles).
__device__ __forceinline__ double MultiTap(const double *__restrict__ blob, const Sp1Desc &s, int i, const double b[4], int v) {
    const double *c = blob + s.off_c + (size_t)i * s.n_splines + v;
    const int st = s.n_splines;
    return c[0] * b[0] + c[st] * b[1] + c[2 * st] * b[2] + c[3 * st] * b[3];
}

/// Interval of x on grid g[0..n): LUT lower bound + forward scan.  Equals GridInterval.
__device__ __forceinline__ int LutInterval(const double *__restrict__ blob, const LutDesc &L, const double *__restrict__ g, int n,
                                           double x) {
    long long key = (__double_as_longlong(x) >> L.shift) - L.key0;
    key = key < 0 ? 0 : (key > (long long)(L.n_keys - 1) ? (long long)(L.n_keys - 1) : key);
    int i = reinterpret_cast<const unsigned short *>(blob)[L.off + (int)key];
    while (i + 1 < n && g[i + 1] <= x) ++i;
    return i;
}

__device__ __forceinline__ double PP1Eval(const double *__restrict__ blob, const PP1Desc &d, double x) {
    const double *g = blob + d.off_g;
    const int i = LutInterval(blob, d.lut, g, d.n, x);
    const double t = x - g[i];
    const double2 *c = reinterpret_cast<const double2 *>(blob + d.off_pp) + 2 * i;
    const double2 c01 = c[0], c23 = c[1];
    return fma(fma(fma(c23.y, t, c23.x), t, c01.y), t, c01.x);
}

__device__ __forceinline__ double PP2Eval(const double *__restrict__ blob, const PP2Desc &d, double x, double y) {
    const double *gx = blob + d.off_gx, *gy = blob + d.off_gy;
    const int ix = LutInterval(blob, d.lutx, gx, d.nx, x);
    const int iy = LutInterval(blob, d.luty, gy, d.ny, y);
    const double tx = x - gx[ix], ty = y - gy[iy];
    const double2 *c = reinterpret_cast<const double2 *>(d.cells) + ((size_t)ix * d.ny + iy) * 8;
    double row[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const double2 c01 = __ldg(c + 2 * m), c23 = __ldg(c + 2 * m + 1);
        row[m] = fma(fma(fma(c23.y, ty, c23.x), ty, c01.y), ty, c01.x);
    }
    return fma(fma(fma(row[3], tx, row[2]), tx, row[1]), tx, row[0]);
}

/// Interval lookup on the David fast tables held in GLOBAL memory (see DLookup in pair_fast.cuh
/// for the shared-memory twin); the table kind is a run-time value here.
__device__ __forceinline__ void DavidLookupGlobal(const unsigned char *__restrict__ tb, const FastLut &L, int off_gpair, double x, int &i,
                                                  double &t) {
    int key;
    if (L.kind == 0)
        key = min(__double2loint(fma(x, L.inv_h, 6755399441055744.0)), L.key_max);  // 2^52 + 2^51: round to nearest integer
    else
        key = min(max((__double2hiint(x) >> L.shift) - L.key0, 0), L.key_max);
    const int i0 = __ldg(reinterpret_cast<const unsigned short *>(tb + L.off_lut) + key);
    const double2 g = __ldg(reinterpret_cast<const double2 *>(tb + off_gpair) + i0);
    const bool up = x >= g.y;
    i = up ? i0 + 1 : i0;
    t = x - (up ? g.y : g.x);
}

__device__ __forceinline__ double DavidEndpointGlobal(const unsigned char *__restrict__ tb, const FastDavidTable &T, double x) {
    x = x > T.r_max ? T.r_max : (x < T.r_min ? T.r_min : x);
    int i;
    double t;
    DavidLookupGlobal(tb, T.lut, T.off_gpair, x, i, t);
    const double2 c01 = __ldg(reinterpret_cast<const double2 *>(tb + T.off_e01) + i);
    const double2 c23 = __ldg(reinterpret_cast<const double2 *>(tb + T.off_e23) + i);
    return fma(fma(fma(c23.y, t, c23.x), t, c01.y), t, c01.x);
}

/// DavidPairAction::CalcU / CalcdUdBeta (david_pair_action_class.h:63-102,126-168) in pp form from
/// the fast kernel's table block; n_order (1..3) and the table kind are run-time values.
__device__ __forceinline__ double DavidFastGlobal(const unsigned char *__restrict__ tb, const FastDavidTable &T, double r, double r_p,
                                                  double s) {
    const double q = 0.5 * (r + r_p), z = r - r_p;  // before the clamp (david...:65-70)
    double u = 0.5 * DavidEndpointGlobal(tb, T, r);
    u = fma(0.5, DavidEndpointGlobal(tb, T, r_p), u);
    if (s > 0.0 && q < T.r_max) {
        int i;
        double t;
        DavidLookupGlobal(tb, T.lut, T.off_gpair, q, i, t);
        const double2 *rec = reinterpret_cast<const double2 *>(tb + T.off_q + (size_t)i * T.q_stride);
        const double s_2 = s * s, z_2 = z * z;
        double sp[4] = {1., s_2, s_2 * s_2, s_2 * s_2 * s_2}, zp[4] = {1., z_2, z_2 * z_2, z_2 * z_2 * z_2};
        int v = 0;
        double od = 0.;
#pragma unroll
        for (int k = 1; k <= 3; ++k) {
            if (k <= T.n_order) {
#pragma unroll
                for (int j = 0; j <= k; ++j, ++v) {
                    const double2 c01 = __ldg(rec + 2 * v), c23 = __ldg(rec + 2 * v + 1);
                    const double cof = fma(fma(fma(c23.y, t, c23.x), t, c01.y), t, c01.x);
                    od = fma(cof, zp[j] * sp[k - j], od);
                }
            }
        }
        u += od;
    }
    return u;
}

/// CalcV of the three action families.
template <int ATYPE>
__device__ __forceinline__ double PairV(const double *__restrict__ blob, const PairTable &T, double r, double r_p) {
    if (ATYPE == ATYPE_DAVID) {
        const double *db = T.dav_blob;
        SetLimits(T.dav.r_min, T.dav.r_max, r, r_p);
        double b[4];
        int i = Sp1Interval(db, T.dav, r);
        BasisOnInterval(db + T.dav.off_t, db + T.dav.off_w, i, r, b);
        double v0 = MultiTap(db, T.dav, i, b, 0);
        i = Sp1Interval(db, T.dav, r_p);
        BasisOnInterval(db + T.dav.off_t, db + T.dav.off_w, i, r_p, b);
        double v1 = MultiTap(db, T.dav, i, b, 0);
        return 0.5 * (v0 + v1);
    }
    SetLimits(T.a.r_min, T.a.r_max, r, r_p);
    double v = 0.;
    if (ATYPE == ATYPE_BARE && T.is_coulomb) {
        v += (0.5 / r) + (0.5 / r_p);
    } else {
        v += 0.5 * PP1Eval(blob, T.a, r);
        v += 0.5 * PP1Eval(blob, T.a, r_p);
    }
    if (T.use_lr) {
        SetLimits(T.lr.r_min, T.lr.r_max, r, r_p);
        v -= 0.5 * PP1Eval(blob, T.lr, r);
        v -= 0.5 * PP1Eval(blob, T.lr, r_p);
    }
    return v;
}

/// CalcU (WHICH_U) / CalcdUdBeta (WHICH_DU) / CalcV (WHICH_V) for one (r, r', s) triple.
template <int ATYPE, int WHICH>
__device__ __forceinline__ double PairEval(const double *__restrict__ blob, const PairTable &T, double r, double r_p, double s) {
    if (WHICH == WHICH_V) return PairV<ATYPE>(blob, T, r, r_p);
    if (ATYPE == ATYPE_BARE) {
        const double v = PairV<ATYPE_BARE>(blob, T, r, r_p);
        return WHICH == WHICH_U ? T.u_scale * v : v;
    }
    if (ATYPE == ATYPE_ILKKA) {
        const double q = 0.5 * (r + r_p);
        const double x = q + 0.5 * s;
        const double y = q - 0.5 * s;
        double u = PP2Eval(blob, T.xy, x, y);
        if (T.use_lr) {
            SetLimits(T.lr.r_min, T.lr.r_max, r, r_p);
            u -= 0.5 * PP1Eval(blob, T.lr, r);
            u -= 0.5 * PP1Eval(blob, T.lr, r_p);
        }
        return u;
    }
    // David: endpoint term plus the off-diagonal polynomial in z^2 and s^2
    if (T.dav_use_fast) return DavidFastGlobal(T.dav_fast_tab, T.dav_fast, r, r_p, s);
    const double *db = T.dav_blob;
    const double q = 0.5 * (r + r_p);
    const double z = r - r_p;
    const double r_max = T.dav.r_max;
    SetLimits(T.dav.r_min, T.dav.r_max, r, r_p);
    double b[4];
    int i = Sp1Interval(db, T.dav, r);
    BasisOnInterval(db + T.dav.off_t, db + T.dav.off_w, i, r, b);
    double e1 = MultiTap(db, T.dav, i, b, 1);
    double e0 = (WHICH == WHICH_DU) ? MultiTap(db, T.dav, i, b, 0) : 0.;
    i = Sp1Interval(db, T.dav, r_p);
    BasisOnInterval(db + T.dav.off_t, db + T.dav.off_w, i, r_p, b);
    double f1 = MultiTap(db, T.dav, i, b, 1);
    double f0 = (WHICH == WHICH_DU) ? MultiTap(db, T.dav, i, b, 0) : 0.;
    double u = 0.5 * (e1 + f1);
    if (WHICH == WHICH_DU) u += 0.5 * (e0 + f0);
    if (s > 0.0 && q < r_max) {
        i = Sp1Interval(db, T.dav, q);
        BasisOnInterval(db + T.dav.off_t, db + T.dav.off_w, i, q, b);
        const double z_2 = z * z, s_2 = s * s, i_s_2 = 1. / s_2;
        double s_2_k = s_2;
        for (int k = 1; k <= T.n_order; k++) {
            double z_2_j = 1, current_s = s_2_k;
            for (int j = 0; j <= k; j++) {
                const double cof = MultiTap(db, T.dav, i, b, k * (k + 1) / 2 + (j + 1));
                u += cof * z_2_j * current_s;
                z_2_j *= z_2;
                current_s *= i_s_2;
            }
            s_2_k *= s_2;
        }
    }
    return u;
}

/// Deterministic block-wide sum: warp shuffles, then warp leaders in fixed order.
template <int NT>
__device__ __forceinline__ double BlockSum(double v, double *scratch /* NT/32 doubles */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double tot = 0.;
    if (threadIdx.x == 0) {
#pragma unroll 1
        for (int i = 0; i < NT / 32; ++i) tot += scratch[i];
    }
    return tot;  // valid on thread 0
}

}  // namespace pimc

#endif  // SIMPIMC_B200_DEVICE_MATH_CUH_
