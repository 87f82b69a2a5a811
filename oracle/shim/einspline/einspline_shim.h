// TEST INFRASTRUCTURE ONLY -- C-style einspline API surface used by /root/reference/src,
// implemented on top of the oracle's spline restatement (oracle/spline_oracle.h).
// einspline ("meinspline" fork, unpinned) is not vendored in the reference: see the header
// of spline_oracle.h.  create_linear_grid / create_loglin_grid exist only in the fork;
// LINEAR is provided as an equally spaced general grid, LOGLIN aborts.
#ifndef ORACLE_SHIM_EINSPLINE_SHIM_H_
#define ORACLE_SHIM_EINSPLINE_SHIM_H_

#include <iostream>
#include "../../spline_oracle.h"

typedef enum { PERIODIC, DERIV1, DERIV2, FLAT, NATURAL, ANTIPERIODIC } bc_code;
typedef struct {
    bc_code lCode, rCode;
    double lVal, rVal;
} BCtype_d;

typedef struct {
    double start, end;
    int num;
    double delta, delta_inv;
} Ugrid;

struct NUgrid : public orc::Grid {};

typedef orc::Spline1D NUBspline_1d_d;
typedef orc::Spline2D NUBspline_2d_d;
typedef orc::MultiSpline1D multi_NUBspline_1d_d;
typedef orc::USpline1D UBspline_1d_d;

inline void einspline_shim_require_natural(const BCtype_d &bc) {
    if (bc.lCode != NATURAL || bc.rCode != NATURAL) {
        std::cerr << "oracle einspline shim: only NATURAL boundary conditions are provided" << std::endl;
        std::abort();
    }
}

inline NUgrid *create_general_grid(double *points, int num_points) {
    NUgrid *g = new NUgrid;
    static_cast<orc::Grid &>(*g) = orc::MakeGeneralGrid(points, num_points);
    return g;
}
inline NUgrid *create_log_grid(double start, double end, int num_points) {
    NUgrid *g = new NUgrid;
    static_cast<orc::Grid &>(*g) = orc::MakeLogGrid(start, end, num_points);
    return g;
}
inline NUgrid *create_linear_grid(double start, double end, int num_points) {
    NUgrid *g = new NUgrid;
    static_cast<orc::Grid &>(*g) = orc::MakeLinearGrid(start, end, num_points);
    return g;
}
inline NUgrid *create_loglin_grid(double, double, double, int) {
    std::cerr << "oracle einspline shim: create_loglin_grid is fork-only and not provided" << std::endl;
    std::abort();
}

inline NUBspline_1d_d *create_NUBspline_1d_d(NUgrid *x_grid, BCtype_d xBC, double *data) {
    einspline_shim_require_natural(xBC);
    NUBspline_1d_d *s = new NUBspline_1d_d;
    s->Create(*x_grid, data);
    return s;
}
inline NUBspline_2d_d *create_NUBspline_2d_d(NUgrid *x_grid, NUgrid *y_grid, BCtype_d xBC, BCtype_d yBC, double *data) {
    einspline_shim_require_natural(xBC);
    einspline_shim_require_natural(yBC);
    NUBspline_2d_d *s = new NUBspline_2d_d;
    s->Create(*x_grid, *y_grid, data);
    return s;
}
inline multi_NUBspline_1d_d *create_multi_NUBspline_1d_d(NUgrid *x_grid, BCtype_d xBC, int num_splines) {
    einspline_shim_require_natural(xBC);
    multi_NUBspline_1d_d *s = new multi_NUBspline_1d_d;
    s->Create(*x_grid, num_splines);
    return s;
}
inline void set_multi_NUBspline_1d_d(multi_NUBspline_1d_d *spline, int which, double *data) { spline->Set(which, data); }
inline UBspline_1d_d *create_UBspline_1d_d(Ugrid x_grid, BCtype_d xBC, double *data) {
    einspline_shim_require_natural(xBC);
    UBspline_1d_d *s = new UBspline_1d_d;
    s->Create(x_grid.start, x_grid.end, x_grid.num, data);
    return s;
}

inline void eval_NUBspline_1d_d(NUBspline_1d_d *s, double x, double *val) { *val = s->Eval(x); }
inline void eval_NUBspline_1d_d_vg(NUBspline_1d_d *s, double x, double *val, double *grad) { s->EvalVG(x, val, grad); }
inline void eval_NUBspline_2d_d(NUBspline_2d_d *s, double x, double y, double *val) { *val = s->Eval(x, y); }
inline void eval_NUBspline_2d_d_vg(NUBspline_2d_d *s, double x, double y, double *val, double *grad) { s->EvalVG(x, y, val, grad); }
inline void eval_multi_NUBspline_1d_d(multi_NUBspline_1d_d *s, double x, double *vals) { s->Eval(x, vals); }
inline void eval_UBspline_1d_d(UBspline_1d_d *s, double x, double *val) { *val = s->Eval(x); }
inline void eval_UBspline_1d_d_vg(UBspline_1d_d *s, double x, double *val, double *grad) {
    // central difference is enough: only the (out-of-scope) gradient estimators call this
    const double h = 1e-6 * (s->end - s->start);
    *val = s->Eval(x);
    *grad = (s->Eval(x + h) - s->Eval(x - h)) / (2 * h);
}

#endif  // ORACLE_SHIM_EINSPLINE_SHIM_H_
