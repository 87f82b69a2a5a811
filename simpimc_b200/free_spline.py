"""Host mirror of FreeSpline (src/actions/free_spline_class.h:25-84): the free-particle density
matrix of one dimension with `n_images` periodic images, tabulated on the reference's uniform
grid of 10 000 points over [-L/2, L/2] (L = 1000 for an open box) and interpolated by the
natural cubic spline through those points -- what create_UBspline_1d_d with NATURAL ends
builds; here scipy's CubicSpline(bc_type="natural"), an independent implementation of the same
interpolant.  The device tables (csrc/spline_build.h: BuildFreeSpline, csrc/kinetic.cuh) are
tested against this mirror and against the reference's own Kinetic action.

    log rho_free(r)      = -sum_d image_action(r_d) - |r|^2 / (4 lambda tau)
    dlog rho_free / dtau = -sum_d d_image_action_d_tau(r_d) - |r|^2 / (4 lambda tau^2)
"""
import numpy as np


class FreeSpline:
    N_GRID = 10000

    def __init__(self, L, n_images, lam, tau, use_tau_derivative=False):
        self.i_4_lambda_tau = 1.0 / (4.0 * lam * tau)
        self.i_4_lambda_tau_tau = 1.0 / (4.0 * lam * tau * tau)
        self.n_images = int(n_images)
        self.image_action = None
        self.d_image_action_d_tau = None
        if self.n_images == 0:      # the reference splines a table of exact zeros
            return
        from scipy.interpolate import CubicSpline
        t_l = 1000.0 if L == 0.0 else L
        start, end = -t_l / 2.0, t_l / 2.0
        dr = (end - start) / (self.N_GRID - 1)
        r = start + np.arange(self.N_GRID) * dr
        r2 = r * r * self.i_4_lambda_tau
        action = np.zeros(self.N_GRID)
        dtau = np.zeros(self.N_GRID)
        with np.errstate(under="ignore"):
            for image in range(1, self.n_images + 1):
                d_p = r2 - (r + image * t_l) ** 2 * self.i_4_lambda_tau
                d_m = r2 - (r - image * t_l) ** 2 * self.i_4_lambda_tau
                e_p, e_m = np.exp(d_p), np.exp(d_m)
                action += e_p + e_m
                if use_tau_derivative:
                    dtau += (d_p * e_p + d_m * e_m) / tau
        if use_tau_derivative:
            self.d_image_action_d_tau = CubicSpline(r, dtau / (1.0 + action), bc_type="natural")
        self.image_action = CubicSpline(r, -np.log1p(action), bc_type="natural")

    def GetLogRhoFree(self, r):
        """r[..., n_d] -> log rho_free (free_spline_class.h:75-83)."""
        r = np.asarray(r, dtype=np.float64)
        tot = 0.0 if self.image_action is None else -np.sum(self.image_action(r), axis=-1)
        return tot - np.sum(r * r, axis=-1) * self.i_4_lambda_tau

    def GetDLogRhoFreeDTau(self, r):
        """free_spline_class.h:91-99."""
        r = np.asarray(r, dtype=np.float64)
        tot = 0.0 if self.d_image_action_d_tau is None else -np.sum(self.d_image_action_d_tau(r), axis=-1)
        return tot - np.sum(r * r, axis=-1) * self.i_4_lambda_tau_tau
