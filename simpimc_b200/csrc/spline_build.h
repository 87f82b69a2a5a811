// Host-side construction of the spline tables the device kernels evaluate.
//
// The reference builds its tables with einspline's create_NUBspline_{1d,2d}_d /
// create_multi_NUBspline_1d_d using NATURAL boundary conditions
// (src/actions/pair_action/ilkka_pair_action_class.h:263,278-280,294-297;
//  bare_pair_action_class.h:37,47-50; david_pair_action_class.h:263,277-282).
// A table is packed into one contiguous "blob" of doubles so a kernel can stage it into
// shared memory with a single bulk copy:
//
//   1-D spline blob:  [ knots t (n+5) | inverse knot spans w (3(n+2)) | coefficients (n+3) ]
//   2-D spline blob:  [ tx | wx | ty | wy | coefficients (nx+3) x (ny+2) row-major, +4 guard ]
//   multi blob:       [ t | w | coefficients (n+3) x n_splines ]
//
// Knot vector: two phantom knots below the grid and three above, spaced like the first /
// last grid interval; w[3i+j] = 1/(t[i+j+1]-t[i]).  Coefficients solve the interpolation
// conditions plus zero second derivative at both grid ends; one extra (zero) coefficient
// keeps the zero-weight fourth tap at x == grid end in bounds.
#ifndef SIMPIMC_B200_SPLINE_BUILD_H_
#define SIMPIMC_B200_SPLINE_BUILD_H_

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace pimc {

struct KnotBasis {
    int n = 0;               // grid points
    std::vector<double> t;   // n+5
    std::vector<double> w;   // 3(n+2)

    void Build(const double *grid, int n_grid) {
        if (n_grid < 4) throw std::invalid_argument("spline grid needs at least 4 points");
        for (int i = 1; i < n_grid; ++i)
            if (!(grid[i] > grid[i - 1])) throw std::invalid_argument("spline grid is not strictly ascending");
        n = n_grid;
        t.assign(n + 5, 0.0);
        const double h0 = grid[1] - grid[0], h1 = grid[n - 1] - grid[n - 2];
        t[0] = grid[0] - 2.0 * h0;
        t[1] = grid[0] - 1.0 * h0;
        for (int i = 0; i < n; ++i) t[i + 2] = grid[i];
        t[n + 2] = grid[n - 1] + 1.0 * h1;
        t[n + 3] = grid[n - 1] + 2.0 * h1;
        t[n + 4] = grid[n - 1] + 3.0 * h1;
        w.assign(3 * (n + 2), 0.0);
        for (int i = 0; i < n + 2; ++i)
            for (int j = 0; j < 3; ++j) w[3 * i + j] = 1.0 / (t[i + j + 1] - t[i]);
    }

    // The three cubic basis values that are non-zero AT grid point i (functions i, i+1, i+2)
    // and their second derivatives, from the Cox-de Boor recursion on interval [t[i+2], t[i+3]].
    void AtKnot(int i, double val[3], double d2[3]) const {
        const int i2 = i + 2;
        const double x = t[i2];
        const double lin0 = (t[i2 + 1] - x) * w[3 * (i + 2)];  // == 1 at the left end of the interval
        const double q0 = (t[i2 + 1] - x) * w[3 * (i + 1) + 1] * lin0;
        const double q1 = (x - t[i2 - 1]) * w[3 * (i + 1) + 1] * lin0;
        val[0] = (t[i2 + 1] - x) * w[3 * i + 2] * q0;
        val[1] = (x - t[i2 - 2]) * w[3 * i + 2] * q0 + (t[i2 + 2] - x) * w[3 * (i + 1) + 2] * q1;
        val[2] = (x - t[i2 - 1]) * w[3 * (i + 1) + 2] * q1;
        d2[0] = 6.0 * w[3 * i + 2] * w[3 * (i + 1) + 1] * lin0;
        d2[1] = -6.0 * w[3 * (i + 1) + 1] * (w[3 * i + 2] + w[3 * (i + 1) + 2]) * lin0;
        d2[2] = 6.0 * w[3 * (i + 1) + 2] * w[3 * (i + 1) + 1] * lin0;
    }
};

// Natural interpolating coefficients: unknowns c[0..n+1].
//   row 0      : sum_j d2_0[j] c[j]         = 0
//   row i+1    : sum_j val_i[j] c[i+j]      = data[i],  i = 0..n-1
//   row n+1    : sum_j d2_{n-1}[j] c[n-1+j] = 0
// A pentadiagonal band (two sub- and two super-diagonals) holds every row; the collocation
// rows are diagonally dominant, so plain Gaussian elimination inside the band is stable.
inline void SolveNatural(const KnotBasis &kb, const double *data, std::size_t dstride, double *c, std::size_t cstride) {
    const int n = kb.n, m = n + 2;
    std::vector<double> A((std::size_t)m * 5, 0.0), rhs(m, 0.0);
    auto at = [&](int r, int col) -> double & { return A[(std::size_t)r * 5 + (col - r + 2)]; };
    double v[3], d2[3];
    kb.AtKnot(0, v, d2);
    for (int j = 0; j < 3; ++j) at(0, j) = d2[j];
    for (int i = 0; i < n; ++i) {
        kb.AtKnot(i, v, d2);
        for (int j = 0; j < 3; ++j) at(i + 1, i + j) = v[j];
        rhs[i + 1] = data[(std::size_t)i * dstride];
    }
    kb.AtKnot(n - 1, v, d2);
    for (int j = 0; j < 3; ++j) at(m - 1, n - 1 + j) = d2[j];
    for (int col = 0; col < m; ++col) {
        const double piv = at(col, col);
        if (piv == 0.0) throw std::runtime_error("singular spline system");
        for (int r = col + 1; r <= col + 2 && r < m; ++r) {
            const double f = at(r, col) / piv;
            if (f == 0.0) continue;
            for (int cc = col; cc <= col + 2 && cc < m; ++cc) at(r, cc) -= f * at(col, cc);
            rhs[r] -= f * rhs[col];
        }
    }
    for (int r = m - 1; r >= 0; --r) {
        double s = rhs[r];
        for (int cc = r + 1; cc <= r + 2 && cc < m; ++cc) s -= at(r, cc) * c[(std::size_t)cc * cstride];
        c[(std::size_t)r * cstride] = s / at(r, r);
    }
}

inline std::size_t Blob1DSize(int n) { return (std::size_t)(n + 5) + 3 * (n + 2) + (n + 3); }

/// [t | w | c] for a natural 1-D spline through (grid, data).
inline std::vector<double> BuildBlob1D(const double *grid, const double *data, int n) {
    KnotBasis kb;
    kb.Build(grid, n);
    std::vector<double> blob(Blob1DSize(n), 0.0);
    std::copy(kb.t.begin(), kb.t.end(), blob.begin());
    std::copy(kb.w.begin(), kb.w.end(), blob.begin() + (n + 5));
    SolveNatural(kb, data, 1, blob.data() + (n + 5) + 3 * (n + 2), 1);
    return blob;
}

inline std::size_t Blob2DSize(int nx, int ny) {
    return (std::size_t)(nx + 5) + 3 * (nx + 2) + (ny + 5) + 3 * (ny + 2) + (std::size_t)(nx + 3) * (ny + 2) + 4;
}

/// [tx | wx | ty | wy | C] for the tensor-product natural spline through data[ix*ny+iy].
inline std::vector<double> BuildBlob2D(const double *gx, int nx, const double *gy, int ny, const double *data) {
    KnotBasis bx, by;
    bx.Build(gx, nx);
    by.Build(gy, ny);
    std::vector<double> blob(Blob2DSize(nx, ny), 0.0);
    double *p = blob.data();
    std::copy(bx.t.begin(), bx.t.end(), p);
    p += nx + 5;
    std::copy(bx.w.begin(), bx.w.end(), p);
    p += 3 * (nx + 2);
    std::copy(by.t.begin(), by.t.end(), p);
    p += ny + 5;
    std::copy(by.w.begin(), by.w.end(), p);
    p += 3 * (ny + 2);
    const int sy = ny + 2;
    // along x for every data column, then along y for every coefficient row
    for (int iy = 0; iy < ny; ++iy) SolveNatural(bx, data + iy, ny, p + iy, sy);
    std::vector<double> row(ny);
    for (int ix = 0; ix < nx + 2; ++ix) {
        for (int iy = 0; iy < ny; ++iy) row[iy] = p[(std::size_t)ix * sy + iy];
        SolveNatural(by, row.data(), 1, p + (std::size_t)ix * sy, 1);
    }
    return blob;
}

inline std::size_t BlobMultiSize(int n, int n_splines) {
    return (std::size_t)(n + 5) + 3 * (n + 2) + (std::size_t)(n + 3) * n_splines;
}

/// [t | w | C[knot][spline]]; values[s] points at n grid values of spline s.
inline std::vector<double> BuildBlobMulti(const double *grid, int n, const std::vector<std::vector<double>> &values) {
    KnotBasis kb;
    kb.Build(grid, n);
    const int ns = (int)values.size();
    std::vector<double> blob(BlobMultiSize(n, ns), 0.0);
    std::copy(kb.t.begin(), kb.t.end(), blob.begin());
    std::copy(kb.w.begin(), kb.w.end(), blob.begin() + (n + 5));
    double *c = blob.data() + (n + 5) + 3 * (n + 2);
    for (int s = 0; s < ns; ++s) SolveNatural(kb, values[s].data(), 1, c + s, ns);
    return blob;
}


// ---------------------------------------------------------------------------------------------
// Piecewise-polynomial (pp) form of the same splines.
//
// On grid interval i the spline is ONE cubic in tau = x - g[i]; expanding the four live
// B-spline basis functions of that interval into monomials of tau (in long double, from the
// same knots t and inverse spans w the B-spline form uses) and contracting with the
// coefficients gives it directly.  The device then needs one interval lookup, 4 (1-D) or 16
// (2-D cell) coefficients and a Horner evaluation instead of the Cox-de Boor recursion: the
// function evaluated is identical up to rounding (~1e-16 relative to the local magnitude).
// Interval n-1 (the phantom interval right of the grid end, which einspline selects for
// x >= grid end) is included, so x == end and x > end reproduce the B-spline form as well.

struct CubicLD {
    long double c[4] = {0, 0, 0, 0};
    CubicLD TimesLinear(long double a, long double b) const {  // this * (a + b tau), degree stays <= 3
        CubicLD r;
        for (int k = 0; k < 4; ++k) {
            r.c[k] += c[k] * a;
            if (k + 1 < 4) r.c[k + 1] += c[k] * b;
        }
        return r;
    }
    CubicLD Plus(const CubicLD &o) const {
        CubicLD r;
        for (int k = 0; k < 4; ++k) r.c[k] = c[k] + o.c[k];
        return r;
    }
    CubicLD Scaled(long double f) const {
        CubicLD r;
        for (int k = 0; k < 4; ++k) r.c[k] = c[k] * f;
        return r;
    }
};

/// The four cubic basis polynomials alive on interval i, in tau = x - t[i+2].
inline void BasisPolynomials(const KnotBasis &kb, int i, CubicLD b[4]) {
    const int i2 = i + 2;
    const std::vector<double> &t = kb.t, &w = kb.w;
    const long double t0 = t[i2];
    const long double dm2 = t0 - (long double)t[i2 - 2], dm1 = t0 - (long double)t[i2 - 1];
    const long double d1 = (long double)t[i2 + 1] - t0, d2 = (long double)t[i2 + 2] - t0, d3 = (long double)t[i2 + 3] - t0;
    const long double w00 = w[3 * i + 2], w11 = w[3 * (i + 1) + 1], w12 = w[3 * (i + 1) + 2];
    const long double w20 = w[3 * (i + 2) + 0], w21 = w[3 * (i + 2) + 1], w22 = w[3 * (i + 2) + 2];
    CubicLD one;
    one.c[0] = 1;
    // (t1 - x) = d1 - tau, (x - t0) = tau, (x - tm1) = tau + dm1, (t2 - x) = d2 - tau, ...
    const CubicLD l0 = one.TimesLinear(d1, -1).Scaled(w20);
    const CubicLD l1 = one.TimesLinear(0, 1).Scaled(w20);
    const CubicLD q0 = l0.TimesLinear(d1, -1).Scaled(w11);
    const CubicLD q1 = l0.TimesLinear(dm1, 1).Scaled(w11).Plus(l1.TimesLinear(d2, -1).Scaled(w21));
    const CubicLD q2 = l1.TimesLinear(0, 1).Scaled(w21);
    b[0] = q0.TimesLinear(d1, -1).Scaled(w00);
    b[1] = q0.TimesLinear(dm2, 1).Scaled(w00).Plus(q1.TimesLinear(d2, -1).Scaled(w12));
    b[2] = q1.TimesLinear(dm1, 1).Scaled(w12).Plus(q2.TimesLinear(d3, -1).Scaled(w22));
    b[3] = q2.TimesLinear(0, 1).Scaled(w22);
}

/// pp[i][4] for i = 0..n-1 from the n+3 B-spline coefficients c.
inline std::vector<double> PPFrom1D(const KnotBasis &kb, const double *c, std::size_t cstride = 1) {
    const int n = kb.n;
    std::vector<double> pp((std::size_t)n * 4, 0.0);
    for (int i = 0; i < n; ++i) {
        CubicLD b[4];
        BasisPolynomials(kb, i, b);
        for (int m = 0; m < 4; ++m) {
            long double acc = 0;
            for (int k = 0; k < 4; ++k) acc += (long double)c[(std::size_t)(i + k) * cstride] * b[k].c[m];
            pp[(std::size_t)i * 4 + m] = (double)acc;
        }
    }
    return pp;
}

/// cells[ix][iy][m][n] (power m of tau_x, power n of tau_y) from the 2-D coefficient array
/// C[(nx+3)][(ny+2)] (row stride sy = ny + 2, plus guard).
inline std::vector<double> PPFrom2D(const KnotBasis &bx, const KnotBasis &by, const double *C) {
    const int nx = bx.n, ny = by.n, sy = ny + 2;
    std::vector<double> cells((std::size_t)nx * ny * 16, 0.0);
    std::vector<CubicLD> polyy((std::size_t)ny * 4);
    for (int iy = 0; iy < ny; ++iy) BasisPolynomials(by, iy, &polyy[(std::size_t)iy * 4]);
    for (int ix = 0; ix < nx; ++ix) {
        CubicLD a[4];
        BasisPolynomials(bx, ix, a);
        for (int iy = 0; iy < ny; ++iy) {
            const CubicLD *b = &polyy[(std::size_t)iy * 4];
            long double P[4][4] = {{0}};
            for (int k = 0; k < 4; ++k)
                for (int l = 0; l < 4; ++l) {
                    // the tap past the last coefficient column has zero weight in the B-spline
                    // form (x == grid end); it must not pick up the next row's first entry
                    const long double cc = (iy + l < sy) ? (long double)C[(std::size_t)(ix + k) * sy + iy + l] : 0.0L;
                    if (cc == 0) continue;
                    for (int m = 0; m < 4; ++m)
                        for (int nn = 0; nn < 4; ++nn) P[m][nn] += cc * a[k].c[m] * b[l].c[nn];
                }
            double *dst = &cells[((std::size_t)ix * ny + iy) * 16];
            for (int m = 0; m < 4; ++m)
                for (int nn = 0; nn < 4; ++nn) dst[m * 4 + nn] = (double)P[m][nn];
        }
    }
    return cells;
}

/// Interval-search accelerator: buckets of the IEEE-754 bit pattern of x (exponent plus the
/// top `mant_bits` mantissa bits), lut[key] = a lower bound of the interval index for every x
/// in the bucket; the device finishes with a forward scan on the grid.
// ------------------------------------------------------------------------------ FreeSpline
/// FreeSpline (src/actions/free_spline_class.h:25-67) in pp form: the image-sum tables on the
/// uniform 10 000-point grid over [-L/2, L/2] (L = 1000 for an open box, :33-35) and their natural
/// cubic interpolants -- einspline's create_UBspline_1d_d with NATURAL ends is the unique C2
/// piecewise cubic through the points with zero second derivative at both ends, which is what
/// KnotBasis + SolveNatural build on the same points.  [zero_lo, zero_hi) is the run of intervals
/// whose polynomials are zero (|coefficient| < 1e-290: the image sum underflows there).
struct FreeSplineHost {
    int n = 0;                         // grid points
    double start = 0., dr = 0.;
    double i4lt = 0., i4ltt = 0.;      // 1 / (4 lambda tau), 1 / (4 lambda tau^2)
    std::vector<double> pp_action;     // [n][4]
    std::vector<double> pp_dtau;       // [n][4] (use_tau_derivative only)
    int zero_lo = 0, zero_hi = 0, dtau_zero_lo = 0, dtau_zero_hi = 0;
};

inline void ZeroRun(const std::vector<double> &pp, int n_int, int &lo, int &hi) {
    // the longest run of all-zero intervals that contains the grid centre (none: lo = hi = 0)
    auto is_zero = [&](int i) {
        for (int k = 0; k < 4; ++k)
            if (std::fabs(pp[(std::size_t)i * 4 + k]) >= 1e-290) return false;
        return true;
    };
    const int mid = n_int / 2;
    lo = hi = 0;
    if (!is_zero(mid)) return;
    lo = mid;
    hi = mid + 1;
    while (lo > 0 && is_zero(lo - 1)) --lo;
    while (hi < n_int && is_zero(hi)) ++hi;
}

inline FreeSplineHost BuildFreeSpline(double L, unsigned n_images, double lambda, double tau, bool use_tau_derivative) {
    FreeSplineHost f;
    f.i4lt = 1. / (4. * lambda * tau);
    f.i4ltt = 1. / (4. * lambda * tau * tau);
    double t_l = L;
    if (L == 0.) t_l = 1000.;
    const int num = 10000;
    f.n = num;
    f.start = -t_l / 2.;
    const double end = t_l / 2.;
    f.dr = (end - f.start) / (num - 1);
    std::vector<double> grid(num), action(num, 0.), dtau(num, 0.);
    for (int i = 0; i < num; ++i) {
        const double r = f.start + i * f.dr;
        grid[i] = r;
        const double r2_i4lt = r * r * f.i4lt;
        for (unsigned image = 1; image <= n_images; ++image) {
            const double r_p = r + image * t_l, r_m = r - image * t_l;
            const double d_p = r2_i4lt - r_p * r_p * f.i4lt, d_m = r2_i4lt - r_m * r_m * f.i4lt;
            const double e_p = std::exp(d_p), e_m = std::exp(d_m);
            action[i] += e_p + e_m;
            if (use_tau_derivative) dtau[i] += (d_p * e_p + d_m * e_m) / tau;
        }
        if (use_tau_derivative) dtau[i] = dtau[i] / (1. + action[i]);
        action[i] = -std::log1p(action[i]);
    }
    KnotBasis kb;
    kb.Build(grid.data(), num);
    std::vector<double> coefs(num + 3, 0.0);
    SolveNatural(kb, action.data(), 1, coefs.data(), 1);
    f.pp_action = PPFrom1D(kb, coefs.data());
    ZeroRun(f.pp_action, num - 1, f.zero_lo, f.zero_hi);
    if (use_tau_derivative) {
        std::fill(coefs.begin(), coefs.end(), 0.0);
        SolveNatural(kb, dtau.data(), 1, coefs.data(), 1);
        f.pp_dtau = PPFrom1D(kb, coefs.data());
        ZeroRun(f.pp_dtau, num - 1, f.dtau_zero_lo, f.dtau_zero_hi);
    }
    return f;
}

struct BitLut {
    int shift = 0;        // key = (bits(x) >> shift) - key0, clamped to [0, n_keys)
    long long key0 = 0;
    std::vector<uint16_t> lut;
};

inline long long DoubleBits(double x) {
    long long b;
    static_assert(sizeof(b) == sizeof(x), "size");
    std::memcpy(&b, &x, sizeof(b));
    return b;
}

inline BitLut BuildBitLut(const double *g, int n) {
    if (n > 65535) throw std::invalid_argument("grid too long for the 16-bit interval LUT");
    BitLut best;
    // everything below x_lo falls into bucket 0, which must map to interval 0
    const double x_lo = g[0] > 0 ? g[0] : g[1] * (1.0 / 64.0);
    const double x_hi = g[n - 1];
    if (!(x_lo > 0) || !(x_hi > x_lo)) throw std::invalid_argument("interval LUT needs a positive ascending grid");
    for (int mant_bits = 3; mant_bits <= 14; ++mant_bits) {
        BitLut L;
        L.shift = 52 - mant_bits;
        L.key0 = DoubleBits(x_lo) >> L.shift;
        const long long n_keys = (DoubleBits(x_hi) >> L.shift) - L.key0 + 1;
        if (n_keys > 16384 && !best.lut.empty()) break;
        if (n_keys > 65536) break;
        L.lut.assign((std::size_t)n_keys, 0);
        int max_per_bucket = 0, i = 0;
        for (long long key = 1; key < n_keys; ++key) {
            const long long bits = (key + L.key0) << L.shift;
            double edge;
            std::memcpy(&edge, &bits, sizeof(edge));
            int before = i;
            while (i + 1 < n && g[i + 1] <= edge) ++i;
            L.lut[(std::size_t)key] = (uint16_t)i;
            max_per_bucket = std::max(max_per_bucket, i - before);
        }
        best = L;
        if (max_per_bucket <= 1) break;
    }
    if (best.lut.empty()) throw std::runtime_error("could not build the interval LUT");
    return best;
}


/// Uniform-bucket interval table for the fast kernels: bucket k holds every x whose x/h rounds
/// to k (the device computes the key with one fused multiply-add against 2^52+2^51).  lut[k] is
/// the einspline interval of the bucket's lower edge; h is chosen below the smallest knot
/// spacing so that at most one knot lies inside a bucket and a single compare against the
/// next knot finishes the search.  Returns false when the table would need more than
/// `max_keys` buckets (the caller then keeps the general kernel).
struct ULut {
    double inv_h = 0.;
    std::vector<uint16_t> lut;
    double h = 0.;
};

/// Packed form of a uniform interval table (pair_fast.cuh: ULookup16): entry k = lut[k] << 16 | position of the knot that
/// lies inside bucket k = [(k - 1/2) h, (k + 1/2) h), in 1/32768 of h from the lower edge, rounded up and at most 0x7FFF
/// (0: the knot sits at or just below the lower edge -- BuildULut's slack; 0xFFFF: no knot -- the 15-bit position of an
/// x never reaches it).
inline std::vector<uint32_t> PackULut(const ULut &L, const double *g, int n) {
    std::vector<uint32_t> out(L.lut.size());
    for (std::size_t k = 0; k < L.lut.size(); ++k) {
        const int i = L.lut[k];
        uint32_t pos = 0xFFFFu;
        const long double lo = ((long double)k - 0.5L) * (long double)L.h, hi = lo + (long double)L.h;
        if (i + 1 < n && (long double)g[i + 1] < hi) {
            const long double p = std::ceil(((long double)g[i + 1] - lo) / (long double)L.h * 32768.0L);
            pos = p <= 0 ? 0u : (p >= 32767.0L ? 0x7FFFu : (uint32_t)p);
        }
        out[k] = ((uint32_t)i << 16) | pos;
    }
    return out;
}

inline bool BuildULut(const double *g, int n, int max_keys, ULut &out) {
    if (n < 2 || n > 65535 || g[0] < 0.) return false;
    double min_sp = g[1] - g[0];
    for (int i = 2; i < n; ++i) min_sp = std::min(min_sp, g[i] - g[i - 1]);
    if (!(min_sp > 0.)) return false;
    const double h = 0.96 * min_sp;
    const double inv_h = 1.0 / h;
    const double n_keys_d = std::ceil(g[n - 1] * inv_h) + 3.0;
    if (n_keys_d > (double)max_keys) return false;
    const int n_keys = (int)n_keys_d;
    out.inv_h = inv_h;
    out.h = h;
    out.lut.assign((std::size_t)n_keys, 0);
    // interval of x (einspline general-grid reverse map): last i with g[i] <= x, 0 below the
    // grid, n-1 at or above its end
    auto interval = [&](double x) {
        if (x <= g[0]) return 0;
        if (x >= g[n - 1]) return n - 1;
        return (int)(std::upper_bound(g, g + n, x) - g) - 1;
    };
    const double slack = 1e-6;  // covers the rounding of x * inv_h at the bucket edges
    for (int k = 0; k < n_keys; ++k) {
        const double lo = ((double)k - 0.5 - slack) * h, hi = ((double)k + 0.5 + slack) * h;
        const int i_lo = interval(lo), i_hi = interval(hi);
        if (i_hi - i_lo > 1) return false;  // cannot happen while h < min spacing
        out.lut[(std::size_t)k] = (uint16_t)i_lo;
    }
    // the last bucket must lie wholly at or above the grid end: keys are clamped to it
    if (out.lut.back() != n - 1) return false;
    return true;
}

/// Bucket-centred form of a 1-D natural spline (pair_fast.cuh: FastLR2): the uniform buckets of BuildULut
/// ([(k - 1/2) h, (k + 1/2) h), h = 0.96 x the smallest knot spacing: at most one knot per bucket) carry the cubic piece
/// that holds their lower edge, re-expanded in long double about the bucket centre k h (records 0 .. n_keys: one more
/// than buckets, the piece above a knot is the NEXT bucket's record), and the position of their knot in 1/65536 of h
/// from the lower edge (0xFFFF: none, or so close to the upper edge that the next bucket takes it over).
struct LR2Host {
    double h = 0., inv_h16 = 0.;
    std::vector<uint16_t> knot;      // [n_keys]
    std::vector<double> c01, c23;    // [n_keys + 1][2]
};

inline bool BuildLR2(const double *g, int n, const KnotBasis &kb, const double *c, int max_keys, LR2Host &out) {
    if (n < 2 || g[0] < 0.) return false;
    double min_sp = g[1] - g[0];
    for (int i = 2; i < n; ++i) min_sp = std::min(min_sp, g[i] - g[i - 1]);
    if (!(min_sp > 0.)) return false;
    const double h = 0.96 * min_sp;
    const double n_keys_d = std::ceil(g[n - 1] / h) + 3.0;
    if (n_keys_d > (double)std::min(max_keys, 32767)) return false;   // the device packs bucket and position into 31 bits
    const int n_keys = (int)n_keys_d;
    out.h = h;
    out.inv_h16 = 65536.0 / h;
    out.knot.assign((std::size_t)n_keys, 0xFFFF);
    out.c01.assign(2 * ((std::size_t)n_keys + 1), 0.0);
    out.c23.assign(2 * ((std::size_t)n_keys + 1), 0.0);
    auto interval = [&](long double x) {   // einspline's reverse map: last i with g[i] <= x, 0 below the grid, n - 1 at or above its end
        if (x <= g[0]) return 0;
        if (x >= g[n - 1]) return n - 1;
        return (int)(std::upper_bound(g, g + n, (double)x) - g) - 1;
    };
    for (int k = 0; k <= n_keys; ++k) {
        const long double centre = (long double)k * (long double)h;
        const long double lo = centre - 0.5L * h, hi = centre + 0.5L * h;
        const int i = interval(lo);
        // piece i about g[i] (PPFrom1D's sums, kept in long double), then about the bucket centre
        CubicLD b[4];
        BasisPolynomials(kb, i, b);
        long double a[4];
        for (int m = 0; m < 4; ++m) {
            long double acc = 0;
            for (int t = 0; t < 4; ++t) acc += (long double)c[i + t] * b[t].c[m];
            a[m] = acc;
        }
        const long double d = centre - (long double)g[i];
        out.c01[2 * (std::size_t)k] = (double)(a[0] + d * (a[1] + d * (a[2] + d * a[3])));
        out.c01[2 * (std::size_t)k + 1] = (double)(a[1] + d * (2.0L * a[2] + 3.0L * d * a[3]));
        out.c23[2 * (std::size_t)k] = (double)(a[2] + 3.0L * d * a[3]);
        out.c23[2 * (std::size_t)k + 1] = (double)a[3];
        if (k < n_keys && i + 1 < n && (long double)g[i + 1] > lo && (long double)g[i + 1] < hi) {
            if (i + 2 < n && (long double)g[i + 2] < hi) return false;   // cannot happen while h < min spacing
            const long double pos = std::ceil(((long double)g[i + 1] - lo) / h * 65536.0L);
            out.knot[(std::size_t)k] = (uint16_t)std::min<long double>(pos, 65535.0L);
        }
    }
    return true;
}

/// Bit-pattern interval table for the fast David kernel: bucket k holds every x >= 0 whose
/// (high 32 bits >> shift) equals k + key0, i.e. one exponent value and the top (20 - shift)
/// mantissa bits -- buckets of constant relative width 2^-(20-shift), the natural partner of a
/// logarithmic grid.  lut[k] = interval of the bucket's lower edge; the smallest number of
/// mantissa bits that leaves at most one knot per bucket is chosen.  Keys are clamped on the
/// device to [0, n_keys-1]: bucket 0 holds the grid start (everything below it maps to interval
/// 0), the last bucket lies wholly above the grid end (interval n-1).
struct BLut {
    int shift = 0, key0 = 0;
    std::vector<uint16_t> lut;
};

inline bool BuildBLut(const double *g, int n, int max_keys, BLut &out) {
    if (n < 2 || n > 65535 || !(g[0] > 0.)) return false;
    for (int i = 1; i < n; ++i)
        if (!(g[i] > g[i - 1])) return false;
    auto hi_word = [](double x) {
        uint64_t b;
        std::memcpy(&b, &x, sizeof(b));
        return (long long)(b >> 32);
    };
    auto from_hi = [](long long hi) {
        const uint64_t b = (uint64_t)hi << 32;
        double x;
        std::memcpy(&x, &b, sizeof(x));
        return x;
    };
    auto interval = [&](double x) {
        if (x <= g[0]) return 0;
        if (x >= g[n - 1]) return n - 1;
        return (int)(std::upper_bound(g, g + n, x) - g) - 1;
    };
    for (int mant = 0; mant <= 14; ++mant) {
        const int shift = 20 - mant;
        const long long k_first = hi_word(g[0]) >> shift, k_last = (hi_word(g[n - 1]) >> shift) + 1;
        const long long n_keys = k_last - k_first + 1;
        if (n_keys > (long long)max_keys) return false;
        std::vector<uint16_t> lut((std::size_t)n_keys);
        bool ok = true;
        for (long long k = 0; k < n_keys && ok; ++k) {
            const double lo = from_hi((k + k_first) << shift);
            const double hi = std::nextafter(from_hi((k + k_first + 1) << shift), 0.0);  // largest x of the bucket
            const int i_lo = interval(lo), i_hi = interval(hi);
            if (i_hi - i_lo > 1) ok = false;
            lut[(std::size_t)k] = (uint16_t)i_lo;
        }
        if (!ok) continue;
        if (lut.back() != n - 1 || lut.front() != 0) return false;
        out.shift = shift;
        out.key0 = (int)k_first;
        out.lut.swap(lut);
        return true;
    }
    return false;
}

}  // namespace pimc

#endif  // SIMPIMC_B200_SPLINE_BUILD_H_
