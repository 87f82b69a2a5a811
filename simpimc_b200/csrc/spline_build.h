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

#include <cmath>
#include <cstddef>
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

}  // namespace pimc

#endif  // SIMPIMC_B200_SPLINE_BUILD_H_
