// TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked into the product library.
//
// Restatement of the non-uniform cubic B-spline arithmetic that simpimc's pair actions
// call through the einspline library.  einspline is NOT vendored in /root/reference and
// the reference pins no version of it: CMake/FindEinspline.cmake:17-24 downloads
// github.com/etano/meinspline `master`.  The reference has no test or golden vector at
// this boundary, so PARITY IS UNPINNED against the reference here; what is pinned is the
// mathematics: the interpolant (natural cubic spline through the table, tensor product in
// 2-D) is unique, and tests/test_oracle_cpu.py (test_spline_definition_matches_scipy_...) checks this restatement against scipy's
// independent natural cubic spline (1-D and tensor-product 2-D) to 1e-12.
//
// What follows restates the published einspline 0.9.2 algorithm (nugrid.c, nubasis.c,
// nubspline_create.c, nubspline_eval_std_d.h, bspline_create.c) as called from
//   src/actions/pair_action/ilkka_pair_action_class.h:41-49,85,94-96,133,142-144,278-280,294-297
//   src/actions/pair_action/bare_pair_action_class.h:47-50,64-67,109-118
//   src/actions/pair_action/david_pair_action_class.h:34-35,75-76,82,220-230,277-282
//   src/actions/free_spline_class.h:65-67,80,96
#ifndef ORACLE_SPLINE_ORACLE_H_
#define ORACLE_SPLINE_ORACLE_H_

#include <cmath>
#include <cstdlib>
#include <vector>

namespace orc {

enum GridCode { GRID_GENERAL = 0, GRID_LOG = 1 };

/// Non-uniform grid (einspline NUgrid).
struct Grid {
    GridCode code = GRID_GENERAL;
    double start = 0, end = 0;
    int num_points = 0;
    std::vector<double> points;
    double a = 0, ainv = 0, startinv = 0;  // log grid only

    /// einspline general_grid_reverse_map / log_grid_reverse_map: interval index of x.
    int ReverseMap(double x) const {
        if (code == GRID_LOG) {
            int index = (int)std::floor(ainv * std::log(x * startinv));
            return index < 0 ? 0 : index;
        }
        const int N = num_points;
        if (x <= points[0]) return 0;
        if (x >= points[N - 1]) return N - 1;
        int hi = N - 1, lo = 0;
        while (hi - lo >= 2) {
            int mid = (hi + lo) >> 1;
            if (points[mid] > x)
                hi = mid;
            else
                lo = mid;
        }
        return lo;
    }
};

inline Grid MakeGeneralGrid(const double *pts, int n) {
    Grid g;
    g.code = GRID_GENERAL;
    g.points.assign(pts, pts + n);
    g.num_points = n;
    g.start = pts[0];
    g.end = pts[n - 1];
    return g;
}

inline Grid MakeLogGrid(double start, double end, int n) {
    Grid g;
    g.code = GRID_LOG;
    g.start = start;
    g.end = end;
    g.num_points = n;
    g.a = 1.0 / (double)(n - 1) * std::log(end / start);
    g.ainv = 1.0 / g.a;
    g.startinv = 1.0 / start;
    g.points.resize(n);
    for (int i = 0; i < n; i++) g.points[i] = start * std::exp(g.a * (double)i);
    return g;
}

inline Grid MakeLinearGrid(double start, double end, int n) {
    std::vector<double> p(n);
    for (int i = 0; i < n; i++) p[i] = start + (end - start) * (double)i / (double)(n - 1);
    return MakeGeneralGrid(p.data(), n);
}

/// Non-periodic cubic B-spline basis on a non-uniform grid (einspline NUBasis).
/// Knots xv[0..N+4]: two phantom knots below, the grid, three above.
struct Basis {
    Grid grid;
    std::vector<double> xv;     // N+5
    std::vector<double> dxinv;  // 3*(N+2): dxinv[3i+j] = 1/(xv[i+j+1]-xv[i])

    void Init(const Grid &g) {
        grid = g;
        const int N = g.num_points;
        const double *p = g.points.data();
        xv.assign(N + 5, 0.0);
        for (int i = 0; i < N; i++) xv[i + 2] = p[i];
        xv[0] = p[0] - 2.0 * (p[1] - p[0]);
        xv[1] = p[0] - 1.0 * (p[1] - p[0]);
        xv[N + 2] = p[N - 1] + 1.0 * (p[N - 1] - p[N - 2]);
        xv[N + 3] = p[N - 1] + 2.0 * (p[N - 1] - p[N - 2]);
        xv[N + 4] = p[N - 1] + 3.0 * (p[N - 1] - p[N - 2]);
        dxinv.assign(3 * (N + 2), 0.0);
        for (int i = 0; i < N + 2; i++)
            for (int j = 0; j < 3; j++) dxinv[3 * i + j] = 1.0 / (xv[i + j + 1] - xv[i]);
    }

    /// Cox-de Boor recursion for the four cubic functions alive on interval i at x.
    void FuncsOnInterval(int i, double x, double b[4], double b1[2], double b2[3]) const {
        const int i2 = i + 2;
        const double *t = xv.data();
        const double *w = dxinv.data();
        b1[0] = (t[i2 + 1] - x) * w[3 * (i + 2) + 0];
        b1[1] = (x - t[i2]) * w[3 * (i + 2) + 0];
        b2[0] = (t[i2 + 1] - x) * w[3 * (i + 1) + 1] * b1[0];
        b2[1] = ((x - t[i2 - 1]) * w[3 * (i + 1) + 1] * b1[0] + (t[i2 + 2] - x) * w[3 * (i + 2) + 1] * b1[1]);
        b2[2] = (x - t[i2]) * w[3 * (i + 2) + 1] * b1[1];
        b[0] = (t[i2 + 1] - x) * w[3 * (i) + 2] * b2[0];
        b[1] = ((x - t[i2 - 2]) * w[3 * (i) + 2] * b2[0] + (t[i2 + 2] - x) * w[3 * (i + 1) + 2] * b2[1]);
        b[2] = ((x - t[i2 - 1]) * w[3 * (i + 1) + 2] * b2[1] + (t[i2 + 3] - x) * w[3 * (i + 2) + 2] * b2[2]);
        b[3] = (x - t[i2]) * w[3 * (i + 2) + 2] * b2[2];
    }

    /// get_NUBasis_funcs_d: returns the interval index and the 4 basis values.
    int Funcs(double x, double b[4]) const {
        double b1[2], b2[3];
        int i = grid.ReverseMap(x);
        FuncsOnInterval(i, x, b, b1, b2);
        return i;
    }

    /// Basis values and first derivatives (eval_*_vg).
    int FuncsD(double x, double b[4], double db[4]) const {
        double b1[2], b2[3];
        int i = grid.ReverseMap(x);
        FuncsOnInterval(i, x, b, b1, b2);
        const double *w = dxinv.data();
        db[0] = -3.0 * (w[3 * (i) + 2] * b2[0]);
        db[1] = 3.0 * (w[3 * (i) + 2] * b2[0] - w[3 * (i + 1) + 2] * b2[1]);
        db[2] = 3.0 * (w[3 * (i + 1) + 2] * b2[1] - w[3 * (i + 2) + 2] * b2[2]);
        db[3] = 3.0 * (w[3 * (i + 2) + 2] * b2[2]);
        return i;
    }

    /// Values at grid point i (get_NUBasis_funcs_di); the 4th value is exactly 0.
    void FuncsAtPoint(int i, double b[4]) const {
        double b1[2], b2[3];
        FuncsOnInterval(i, grid.points[i], b, b1, b2);
    }

    /// Second derivatives at grid point i (get_NUBasis_d2funcs_di).
    void D2FuncsAtPoint(int i, double d2b[4]) const {
        double b[4], b1[2], b2[3];
        FuncsOnInterval(i, grid.points[i], b, b1, b2);
        const double *w = dxinv.data();
        d2b[0] = 6.0 * (+w[3 * (i + 0) + 2] * w[3 * (i + 1) + 1] * b1[0]);
        d2b[1] = 6.0 * (-w[3 * (i + 1) + 1] * (w[3 * (i + 0) + 2] + w[3 * (i + 1) + 2]) * b1[0] +
                        w[3 * (i + 1) + 2] * w[3 * (i + 2) + 1] * b1[1]);
        d2b[2] = 6.0 * (+w[3 * (i + 1) + 2] * w[3 * (i + 1) + 1] * b1[0] -
                        w[3 * (i + 2) + 1] * (w[3 * (i + 1) + 2] + w[3 * (i + 2) + 2]) * b1[1]);
        d2b[3] = 6.0 * (+w[3 * (i + 2) + 2] * w[3 * (i + 2) + 1] * b1[1]);
    }
};

/// Solve for the M+2 coefficients of the NATURAL (zero second derivative at both ends)
/// interpolating spline through data[0..M) (strided).  Tridiagonal system with the two
/// boundary rows folded in, eliminated top-down and back-substituted, as
/// solve_NUB_deriv_interp_1d_d does.
inline void SolveNatural1D(const Basis &basis, const double *data, int dstride, double *p, int pstride) {
    const int M = basis.grid.num_points;
    const int N = M + 2;
    std::vector<double> bands(4 * N, 0.0);
    double left[4], right[4];
    basis.D2FuncsAtPoint(0, left);
    basis.D2FuncsAtPoint(M - 1, right);
    left[3] = 0.0;   // NATURAL: rhs (second derivative) = 0
    right[3] = 0.0;
    for (int i = 0; i < 4; i++) {
        bands[i] = left[i];
        bands[4 * (N - 1) + i] = right[i];
    }
    for (int i = 0; i < M; i++) {
        basis.FuncsAtPoint(i, &bands[4 * (i + 1)]);
        bands[4 * (i + 1) + 3] = data[dstride * i];
    }
    // first row
    bands[1] /= bands[0];
    bands[2] /= bands[0];
    bands[3] /= bands[0];
    bands[0] = 1.0;
    bands[4 * 1 + 1] -= bands[4 * 1 + 0] * bands[1];
    bands[4 * 1 + 2] -= bands[4 * 1 + 0] * bands[2];
    bands[4 * 1 + 3] -= bands[4 * 1 + 0] * bands[3];
    bands[0] = 0.0;
    bands[4 * 1 + 2] /= bands[4 * 1 + 1];
    bands[4 * 1 + 3] /= bands[4 * 1 + 1];
    bands[4 * 1 + 1] = 1.0;
    // rows 2 .. M
    for (int row = 2; row < N - 1; row++) {
        bands[4 * row + 1] -= bands[4 * row + 0] * bands[4 * (row - 1) + 2];
        bands[4 * row + 3] -= bands[4 * row + 0] * bands[4 * (row - 1) + 3];
        bands[4 * row + 2] /= bands[4 * row + 1];
        bands[4 * row + 3] /= bands[4 * row + 1];
        bands[4 * row + 0] = 0.0;
        bands[4 * row + 1] = 1.0;
    }
    // last row
    bands[4 * (M + 1) + 1] -= bands[4 * (M + 1) + 0] * bands[4 * (M - 1) + 2];
    bands[4 * (M + 1) + 3] -= bands[4 * (M + 1) + 0] * bands[4 * (M - 1) + 3];
    bands[4 * (M + 1) + 2] -= bands[4 * (M + 1) + 1] * bands[4 * (M) + 2];
    bands[4 * (M + 1) + 3] -= bands[4 * (M + 1) + 1] * bands[4 * (M) + 3];
    bands[4 * (M + 1) + 3] /= bands[4 * (M + 1) + 2];
    bands[4 * (M + 1) + 2] = 1.0;
    p[pstride * (M + 1)] = bands[4 * (M + 1) + 3];
    for (int row = M; row > 0; row--)
        p[pstride * row] = bands[4 * row + 3] - bands[4 * row + 2] * p[pstride * (row + 1)];
    p[0] = bands[3] - bands[1] * p[pstride * 1] - bands[2] * p[pstride * 2];
}

/// NUBspline_1d_d with NATURAL/NATURAL boundary conditions.
struct Spline1D {
    Basis basis;
    std::vector<double> coefs;  // M+2 (+1 guard so the zero-weight 4th tap at x==end is defined)

    void Create(const Grid &g, const double *data) {
        basis.Init(g);
        coefs.assign(g.num_points + 3, 0.0);
        SolveNatural1D(basis, data, 1, coefs.data(), 1);
    }
    double Eval(double x) const {
        double b[4];
        int i = basis.Funcs(x, b);
        const double *c = coefs.data();
        return (c[i + 0] * b[0] + c[i + 1] * b[1] + c[i + 2] * b[2] + c[i + 3] * b[3]);
    }
    void EvalVG(double x, double *val, double *grad) const {
        double b[4], db[4];
        int i = basis.FuncsD(x, b, db);
        const double *c = coefs.data();
        *val = (c[i + 0] * b[0] + c[i + 1] * b[1] + c[i + 2] * b[2] + c[i + 3] * b[3]);
        *grad = (c[i + 0] * db[0] + c[i + 1] * db[1] + c[i + 2] * db[2] + c[i + 3] * db[3]);
    }
};

/// NUBspline_2d_d, NATURAL in both directions; data row-major [ix*My+iy].
struct Spline2D {
    Basis xb, yb;
    int Nx = 0, Ny = 0;         // coefficient counts Mx+2, My+2
    std::vector<double> coefs;  // (Nx+1)*Ny + guard row so zero-weight taps stay in bounds

    void Create(const Grid &gx, const Grid &gy, const double *data) {
        xb.Init(gx);
        yb.Init(gy);
        const int Mx = gx.num_points, My = gy.num_points;
        Nx = Mx + 2;
        Ny = My + 2;
        coefs.assign((size_t)(Nx + 1) * Ny + 4, 0.0);
        for (int iy = 0; iy < My; iy++) SolveNatural1D(xb, data + iy, My, coefs.data() + iy, Ny);
        std::vector<double> tmp(My);
        for (int ix = 0; ix < Nx; ix++) {
            // in-place solve along y: copy the row first (source and destination overlap)
            for (int iy = 0; iy < My; iy++) tmp[iy] = coefs[(size_t)ix * Ny + iy];
            SolveNatural1D(yb, tmp.data(), 1, coefs.data() + (size_t)ix * Ny, 1);
        }
    }
    double Eval(double x, double y) const {
        double a[4], b[4];
        int ix = xb.Funcs(x, a);
        int iy = yb.Funcs(y, b);
        const double *c = coefs.data();
        const int xs = Ny;
#define ORC_C(i, j) c[(size_t)(ix + (i)) * xs + iy + (j)]
        double v = (a[0] * (ORC_C(0, 0) * b[0] + ORC_C(0, 1) * b[1] + ORC_C(0, 2) * b[2] + ORC_C(0, 3) * b[3]) +
                    a[1] * (ORC_C(1, 0) * b[0] + ORC_C(1, 1) * b[1] + ORC_C(1, 2) * b[2] + ORC_C(1, 3) * b[3]) +
                    a[2] * (ORC_C(2, 0) * b[0] + ORC_C(2, 1) * b[1] + ORC_C(2, 2) * b[2] + ORC_C(2, 3) * b[3]) +
                    a[3] * (ORC_C(3, 0) * b[0] + ORC_C(3, 1) * b[1] + ORC_C(3, 2) * b[2] + ORC_C(3, 3) * b[3]));
#undef ORC_C
        return v;
    }
    void EvalVG(double x, double y, double *val, double *grad) const {
        double a[4], da[4], b[4], db[4];
        int ix = xb.FuncsD(x, a, da);
        int iy = yb.FuncsD(y, b, db);
        const double *c = coefs.data();
        const int xs = Ny;
        double v = 0, gx = 0, gy = 0;
        for (int m = 0; m < 4; m++)
            for (int n = 0; n < 4; n++) {
                double cc = c[(size_t)(ix + m) * xs + iy + n];
                v += a[m] * b[n] * cc;
                gx += da[m] * b[n] * cc;
                gy += a[m] * db[n] * cc;
            }
        *val = v;
        grad[0] = gx;
        grad[1] = gy;
    }
};

/// multi_NUBspline_1d_d: several splines on one grid, coefficient layout [knot][spline].
struct MultiSpline1D {
    Basis basis;
    int num_splines = 0;
    std::vector<double> coefs;  // (M+3)*num_splines

    void Create(const Grid &g, int n_splines) {
        basis.Init(g);
        num_splines = n_splines;
        coefs.assign((size_t)(g.num_points + 3) * n_splines, 0.0);
    }
    void Set(int which, const double *data) { SolveNatural1D(basis, data, 1, coefs.data() + which, num_splines); }
    void Eval(double x, double *vals) const {
        double b[4];
        int i = basis.Funcs(x, b);
        const int xs = num_splines;
        const double *c0 = coefs.data() + (size_t)(i + 0) * xs;
        const double *c1 = coefs.data() + (size_t)(i + 1) * xs;
        const double *c2 = coefs.data() + (size_t)(i + 2) * xs;
        const double *c3 = coefs.data() + (size_t)(i + 3) * xs;
        for (int n = 0; n < num_splines; n++) vals[n] = c0[n] * b[0] + c1[n] * b[1] + c2[n] * b[2] + c3[n] * b[3];
    }
};

/// UBspline_1d_d (uniform grid), NATURAL/NATURAL; used only by FreeSpline (kinetic action).
struct USpline1D {
    double start = 0, end = 0, delta = 0, delta_inv = 0;
    int num = 0;
    std::vector<double> coefs;  // num+2 (+1 guard)

    void Create(double t_start, double t_end, int t_num, const double *data) {
        start = t_start;
        end = t_end;
        num = t_num;
        delta = (end - start) / (double)(num - 1);
        delta_inv = 1.0 / delta;
        const int M = num, N = M + 2;
        coefs.assign(N + 1, 0.0);
        // rows: natural second-derivative rows (1,-2,1)/delta^2, interior rows (1/6,2/3,1/6)
        std::vector<double> bands(4 * N, 0.0);
        bands[0] = 1.0 * delta_inv * delta_inv;
        bands[1] = -2.0 * delta_inv * delta_inv;
        bands[2] = 1.0 * delta_inv * delta_inv;
        bands[3] = 0.0;
        bands[4 * (N - 1) + 0] = 1.0 * delta_inv * delta_inv;
        bands[4 * (N - 1) + 1] = -2.0 * delta_inv * delta_inv;
        bands[4 * (N - 1) + 2] = 1.0 * delta_inv * delta_inv;
        bands[4 * (N - 1) + 3] = 0.0;
        for (int i = 0; i < M; i++) {
            bands[4 * (i + 1) + 0] = 1.0 / 6.0;
            bands[4 * (i + 1) + 1] = 2.0 / 3.0;
            bands[4 * (i + 1) + 2] = 1.0 / 6.0;
            bands[4 * (i + 1) + 3] = data[i];
        }
        double *p = coefs.data();
        bands[1] /= bands[0];
        bands[2] /= bands[0];
        bands[3] /= bands[0];
        bands[0] = 1.0;
        bands[4 * 1 + 1] -= bands[4 * 1 + 0] * bands[1];
        bands[4 * 1 + 2] -= bands[4 * 1 + 0] * bands[2];
        bands[4 * 1 + 3] -= bands[4 * 1 + 0] * bands[3];
        bands[0] = 0.0;
        bands[4 * 1 + 2] /= bands[4 * 1 + 1];
        bands[4 * 1 + 3] /= bands[4 * 1 + 1];
        bands[4 * 1 + 1] = 1.0;
        for (int row = 2; row < N - 1; row++) {
            bands[4 * row + 1] -= bands[4 * row + 0] * bands[4 * (row - 1) + 2];
            bands[4 * row + 3] -= bands[4 * row + 0] * bands[4 * (row - 1) + 3];
            bands[4 * row + 2] /= bands[4 * row + 1];
            bands[4 * row + 3] /= bands[4 * row + 1];
            bands[4 * row + 0] = 0.0;
            bands[4 * row + 1] = 1.0;
        }
        bands[4 * (M + 1) + 1] -= bands[4 * (M + 1) + 0] * bands[4 * (M - 1) + 2];
        bands[4 * (M + 1) + 3] -= bands[4 * (M + 1) + 0] * bands[4 * (M - 1) + 3];
        bands[4 * (M + 1) + 2] -= bands[4 * (M + 1) + 1] * bands[4 * (M) + 2];
        bands[4 * (M + 1) + 3] -= bands[4 * (M + 1) + 1] * bands[4 * (M) + 3];
        bands[4 * (M + 1) + 3] /= bands[4 * (M + 1) + 2];
        bands[4 * (M + 1) + 2] = 1.0;
        p[M + 1] = bands[4 * (M + 1) + 3];
        for (int row = M; row > 0; row--) p[row] = bands[4 * row + 3] - bands[4 * row + 2] * p[row + 1];
        p[0] = bands[3] - bands[1] * p[1] - bands[2] * p[2];
    }
    double Eval(double x) const {
        double u = (x - start) * delta_inv;
        double ipart;
        double t = std::modf(u, &ipart);
        int i = (int)ipart;
        if (i < 0) { i = 0; }
        if (i > num - 1) { i = num - 1; }
        const double tp0 = t * t * t, tp1 = t * t, tp2 = t;
        const double *c = coefs.data();
        return (c[i + 0] * (-1.0 / 6.0 * tp0 + 3.0 / 6.0 * tp1 - 3.0 / 6.0 * tp2 + 1.0 / 6.0) +
                c[i + 1] * (3.0 / 6.0 * tp0 - 6.0 / 6.0 * tp1 + 0.0 / 6.0 * tp2 + 4.0 / 6.0) +
                c[i + 2] * (-3.0 / 6.0 * tp0 + 3.0 / 6.0 * tp1 + 3.0 / 6.0 * tp2 + 1.0 / 6.0) +
                c[i + 3] * (1.0 / 6.0 * tp0 + 0.0 / 6.0 * tp1 + 0.0 / 6.0 * tp2 + 0.0 / 6.0));
    }
};

}  // namespace orc

#endif  // ORACLE_SPLINE_ORACLE_H_
