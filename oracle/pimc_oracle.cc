// TEST INFRASTRUCTURE ONLY -- CPU oracle (builds oracle/liboracle.so).
//
// Plain C++ restatement, on flat arrays, of the reference's action-evaluation path for ONE
// walker.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load it, and only as the checker / reported baseline; the product
// library (simpimc_b200/csrc) never links, imports or calls anything in this directory.
//
// Pinning: oracle/make_golden.py runs the reference's own classes (oracle/_ref, built from
// /root/reference/src in place) and this restatement on the same seeded inputs and commits
// the reference's outputs under tests/golden/; tests/test_oracle_*.py hold this file to
// them.  The spline arithmetic (oracle/spline_oracle.h) is the one part the reference
// delegates to an un-vendored library (einspline) -- "parity unpinned" there, see that
// header.
//
// Each function cites the reference file:line (relative to /root/reference) it follows.
// Compile WITHOUT -ffast-math and with -ffp-contract=off: bin indices and k ordering are
// compared bit-exactly.

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../include/simpimc_b200.h"
#include "spline_oracle.h"

namespace {

typedef std::complex<double> cplx;

// scaffold mag() = arma::norm(v,2): two accumulators over even / odd elements
// (include/scaffold/matrix/armadillo.h:32-33; Armadillo op_norm::vec_norm_2_direct_std).
inline double Mag(const double *v, int n) {
    double acc1 = 0., acc2 = 0.;
    int i, j;
    for (i = 0, j = 1; j < n; i += 2, j += 2) {
        acc1 += v[i] * v[i];
        acc2 += v[j] * v[j];
    }
    if (i < n) acc1 += v[i] * v[i];
    return std::sqrt(acc1 + acc2);
}
// scaffold dot() = arma::dot (op_dot::direct_dot_arma), same pairing.
inline double Dot(const double *a, const double *b, int n) {
    double v1 = 0., v2 = 0.;
    int i, j;
    for (i = 0, j = 1; j < n; i += 2, j += 2) {
        v1 += a[i] * b[i];
        v2 += a[j] * b[j];
    }
    if (i < n) v1 += a[i] * b[i];
    return v1 + v2;
}
// include/scaffold/algorithm/algorithm.h:22-25
inline double CMag2(const cplx &z1, const cplx &z2) { return z1.real() * z2.real() + z1.imag() * z2.imag(); }
// include/scaffold/algorithm/algorithm.h:46-50
inline bool FEqual(double a, double b, double tol) { return std::fabs(a - b) < tol; }

// ------------------------------------------------------------------------------- KSpace
struct KSpace {  // src/data_structures/k_space_class.h:6-15
    double cutoff = 0., L = 0.;
    int n_d = 3;
    std::vector<double> mags;
    std::vector<double> vecs;   // [k][d]
    std::vector<int> indices;   // [k][d], offset by max_index like the reference
    std::vector<double> box;
    std::vector<int> max_index;

    // k_space_class.h:17-30
    bool Include(const double *k, double k_cut) const {
        double k2 = Dot(k, k, n_d);
        if (k2 < k_cut * k_cut && k2 != 0.) {
            if (k[0] > 0.) return true;
            if (n_d >= 2 && (k[0] == 0. && k[1] > 0.)) return true;
            if (n_d >= 3 && (k[0] == 0. && k[1] == 0. && k[2] > 0.)) return true;
            return false;
        }
        return false;
    }

    // k_space_class.h:33-80.  GenCombPermK(t_indices, is, n_d, false, true)
    // (algorithm.h:101-121) over the multiset holding every index n_d times walks the
    // sorted n_d-multisets in lexicographic order and, for each, its distinct permutations
    // in lexicographic order; that order is reproduced here directly.
    void Setup(double k_cut) {
        if (k_cut <= cutoff) return;
        vecs.clear();
        indices.clear();
        mags.clear();
        cutoff = k_cut;
        box.assign(n_d, 2. * M_PI / L);
        max_index.assign(n_d, 0);
        for (int d = 0; d < n_d; d++) max_index[d] = (uint32_t)std::ceil(1.1 * k_cut / box[d]);
        const int m = max_index[0];
        std::vector<int> comb(n_d, -m);
        while (true) {
            std::vector<int> perm(comb);  // already sorted ascending
            do {
                double k[3];
                for (int d = 0; d < n_d; d++) k[d] = perm[d] * box[d];
                if (Include(k, k_cut)) {
                    for (int d = 0; d < n_d; d++) {
                        vecs.push_back(k[d]);
                        indices.push_back(max_index[d] + perm[d]);
                    }
                    mags.push_back(std::sqrt(Dot(k, k, n_d)));
                }
            } while (std::next_permutation(perm.begin(), perm.end()));
            // next non-decreasing tuple
            int pos = n_d - 1;
            while (pos >= 0 && comb[pos] == m) pos--;
            if (pos < 0) break;
            int v = comb[pos] + 1;
            for (int d = pos; d < n_d; d++) comb[d] = v;
        }
    }
    size_t n_k() const { return mags.size(); }
};

// ------------------------------------------------------------------------------ Species
struct Species {  // src/data_structures/species_class.h:8-25 + bead_class.h:11-21
    int n_part = 0, n_bead = 0, n_d = 3;
    double lambda = 0.;
    std::vector<double> r, r_c;       // [p][b][d]   Bead::r_, Bead::r_c_
    std::vector<cplx> rho_k, rho_k_c; // [b][k]      Species::rho_k_, rho_k_c_
    bool need_update_rho_k = true;

    int bead_loop(int b) const { return b >= n_bead ? b - n_bead : b; }  // species_class.h:101
    double *R(int mode, int p, int b) { return (mode ? r : r_c).data() + ((size_t)p * n_bead + bead_loop(b)) * n_d; }
};

struct World;

// ------------------------------------------------------------------------------- Actions
struct LongRange {
    bool present = false;
    orc::Spline1D r_spline;
    double r_min = 0, r_max = 0, r_0 = 0, k_0 = 0;
    std::vector<double> k_tab, fk_tab;  // as read from the table
    std::vector<double> f_k;            // per k vector (ilkka...:308-315)
};

struct Action {
    enum Type { ILKKA, BARE, DAVID } type;
    World *w;
    int sa, sb;
    int max_level = 0;
    bool use_long_range = false;
    bool is_constant = false, is_first_time = true;
    double dUdB_constant = 0., potential_constant = 0.;
    double k_cut = 0.;
    // Ilkka
    orc::Spline2D u_xy, du_xy;
    orc::Spline1D v_r;
    double r_v_min = 0, r_v_max = 0;
    LongRange u_long, du_long, v_long;
    // Bare
    bool is_coulomb = false;
    // David
    int n_order = 0, n_val = 0;
    orc::Grid grid;
    orc::MultiSpline1D u_kj, du_kj;
    std::vector<double> d_u_long_k, d_du_long_k, d_v_long_k;
    double d_u_long_k_0 = 0, d_u_long_r_0 = 0, d_du_long_k_0 = 0, d_du_long_r_0 = 0, d_v_long_k_0 = 0, d_v_long_r_0 = 0;
};

struct World {  // src/data_structures/path_class.h:8-23
    int n_d = 3, n_bead = 0;
    bool pbc = true;
    double L = 0, iL = 0, vol = 1, beta = 0, tau = 0;
    int mode = 1;
    KSpace ks;
    std::vector<Species> species;
    std::vector<std::unique_ptr<Action>> actions;

    // path_class.h:131-134
    void PutInBox(double *r) const {
        for (int d = 0; d < n_d; ++d) r[d] -= std::nearbyint(r[d] * iL) * L;
    }
    // path_class.h:108-112
    void Dr(const double *r0, const double *r1, double *dr) const {
        for (int d = 0; d < n_d; ++d) dr[d] = r0[d] - r1[d];
        PutInBox(dr);
    }
    // path_class.h:137-150
    void DrDrpDrrp(int b0, int b1, int s0, int s1, int p0, int p1, double &r_mag, double &r_p_mag, double &r_r_p_mag) {
        double r[3], r_p[3], r_r_p[3];
        const double *a0 = species[s0].R(mode, p0, b0), *a1 = species[s1].R(mode, p1, b0);
        const double *c0 = species[s0].R(mode, p0, b1), *c1 = species[s1].R(mode, p1, b1);
        for (int d = 0; d < n_d; ++d) {
            r[d] = a1[d] - a0[d];
            r_p[d] = c1[d] - c0[d];
        }
        for (int d = 0; d < n_d; ++d) {
            r[d] -= std::nearbyint(r[d] * iL) * L;
            r_p[d] += std::nearbyint((r[d] - r_p[d]) * iL) * L;
        }
        for (int d = 0; d < n_d; ++d) r_r_p[d] = r[d] - r_p[d];
        for (int d = 0; d < n_d; ++d) r_r_p[d] -= std::nearbyint(r_r_p[d] * iL) * L;
        r_mag = Mag(r, n_d);
        r_p_mag = Mag(r_p, n_d);
        r_r_p_mag = Mag(r_r_p, n_d);
    }

    // KSpace::CalcC + Bead::CalcRhoK (k_space_class.h:83-94, bead_class.h:125-133)
    void BeadRhoK(const double *r, cplx *out) const {
        if (ks.n_k() == 0) return;  // no k vectors (open boundary, or cutoff at its default): nothing to fill
        std::vector<std::vector<cplx>> c_k(n_d);
        for (int d = 0; d < n_d; d++) {
            const int mx = ks.max_index[d];
            c_k[d].assign(2 * mx + 1, cplx(0., 0.));
            double phi = r[d] * ks.box[d];
            cplx tmp(std::cos(phi), std::sin(phi));
            c_k[d][mx] = 1.;
            for (int k_i = 1; k_i <= mx; k_i++) {
                c_k[d][mx + k_i] = tmp * c_k[d][mx + k_i - 1];
                c_k[d][mx - k_i] = std::conj(c_k[d][mx + k_i]);
            }
        }
        const size_t n_k = ks.n_k();
        for (size_t k_i = 0; k_i < n_k; ++k_i) {
            cplx factor = 1.;
            for (int d = 0; d < n_d; d++) factor *= c_k[d][ks.indices[k_i * n_d + d]];
            out[k_i] = factor;
        }
    }

    // Species::InitRhoK (species_class.h:380-403)
    void InitRhoK(int sp) {
        Species &s = species[sp];
        const size_t n_k = ks.n_k();
        s.rho_k.assign((size_t)n_bead * n_k, cplx(0., 0.));
        std::vector<cplx> tmp(n_k);
        for (int b = 0; b < n_bead; b++)
            for (int p = 0; p < s.n_part; p++) {
                BeadRhoK(s.R(1, p, b), tmp.data());
                for (size_t k = 0; k < n_k; ++k) s.rho_k[b * n_k + k] += tmp[k];
            }
        s.rho_k_c = s.rho_k;
    }

    // Species::UpdateRhoK (species_class.h:406-425).  The reference keeps every bead's own
    // rho_k and its copy; the copy always equals CalcRhoK of r_c, so it is recomputed here.
    void UpdateRhoK(int sp, int b0, int b1, const std::vector<int> &particles, int level) {
        Species &s = species[sp];
        const size_t n_k = ks.n_k();
        const int skip = 1 << level;
        for (int b = b0; b < b1; b += skip) {
            int bl = s.bead_loop(b);
            for (size_t k = 0; k < n_k; ++k) s.rho_k[bl * n_k + k] = s.rho_k_c[bl * n_k + k];
        }
        std::vector<cplx> nw(n_k), od(n_k);
        for (int p : particles)
            for (int b = b0; b < b1; b += skip) {
                int bl = s.bead_loop(b);
                BeadRhoK(s.R(1, p, b), nw.data());
                BeadRhoK(s.R(0, p, b), od.data());
                for (size_t k = 0; k < n_k; ++k) s.rho_k[bl * n_k + k] += nw[k] - od[k];
            }
        s.need_update_rho_k = false;
    }
    std::vector<cplx> &RhoK(int sp) { return mode ? species[sp].rho_k : species[sp].rho_k_c; }
};

// pair_action_class.h:32-42
inline void SetLimits(double r_min, double r_max, double &r, double &r_p) {
    if (r > r_max)
        r = r_max;
    else if (r < r_min)
        r = r_min;
    if (r_p > r_max)
        r_p = r_max;
    else if (r_p < r_min)
        r_p = r_min;
}

// ---- per-pair kernels -------------------------------------------------------------------
double CalcV(Action &a, double r, double r_p, int level) {
    World &w = *a.w;
    if (a.type == Action::DAVID) {  // david_pair_action_class.h:26-39
        SetLimits(a.grid.start, a.grid.end, r, r_p);
        std::vector<double> rv(a.n_val + 1), rpv(a.n_val + 1);
        a.u_kj.Eval(r, rv.data());
        a.u_kj.Eval(r_p, rpv.data());
        return 0.5 * (rv[0] + rpv[0]);
    }
    // ilkka_pair_action_class.h:34-54 and bare_pair_action_class.h:99-123
    SetLimits(a.r_v_min, a.r_v_max, r, r_p);
    double v = 0.;
    if (a.type == Action::BARE && a.is_coulomb) {
        v += (0.5 / r) + (0.5 / r_p);
    } else {
        v += 0.5 * a.v_r.Eval(r);
        v += 0.5 * a.v_r.Eval(r_p);
    }
    if (a.use_long_range) {
        SetLimits(a.v_long.r_min, a.v_long.r_max, r, r_p);
        v -= 0.5 * a.v_long.r_spline.Eval(r);
        v -= 0.5 * a.v_long.r_spline.Eval(r_p);
    }
    (void)w;
    return v;
}

double CalcU(Action &a, double r, double r_p, double s, int level) {
    World &w = *a.w;
    if (a.type == Action::BARE) {  // bare_pair_action_class.h:146-150
        uint32_t skip = 1 >> level;  // sic: 1 at level 0, 0 above (App. A-2)
        double level_tau = skip * w.tau;
        return level_tau * CalcV(a, r, r_p, level);
    }
    if (a.type == Action::ILKKA) {  // ilkka_pair_action_class.h:77-101
        double q = 0.5 * (r + r_p);
        double x = q + 0.5 * s;
        double y = q - 0.5 * s;
        double u = 0.;
        u = a.u_xy.Eval(x, y);
        if (a.use_long_range) {
            SetLimits(a.u_long.r_min, a.u_long.r_max, r, r_p);
            u -= 0.5 * a.u_long.r_spline.Eval(r);
            u -= 0.5 * a.u_long.r_spline.Eval(r_p);
        }
        return u;
    }
    // david_pair_action_class.h:63-102
    double q = 0.5 * (r + r_p);
    double z = r - r_p;
    double r_max = a.grid.end;
    SetLimits(a.grid.start, a.grid.end, r, r_p);
    std::vector<double> rv(a.n_val + 1), rpv(a.n_val + 1);
    a.u_kj.Eval(r, rv.data());
    a.u_kj.Eval(r_p, rpv.data());
    double u = 0.5 * (rv[1] + rpv[1]);
    if (s > 0.0 && q < r_max) {
        std::vector<double> qv(a.n_val + 1);
        a.u_kj.Eval(q, qv.data());
        double z_2 = z * z, s_2 = s * s, i_s_2 = 1. / s_2, s_2_k = s_2;
        for (int k = 1; k <= a.n_order; k++) {
            double z_2_j = 1, current_s = s_2_k;
            for (int j = 0; j <= k; j++) {
                double u_cof = qv[k * (k + 1) / 2 + (j + 1)];
                u += u_cof * z_2_j * current_s;
                z_2_j *= z_2;
                current_s *= i_s_2;
            }
            s_2_k *= s_2;
        }
    }
    return u;
}

double CalcdUdBeta(Action &a, double r, double r_p, double s, int level) {
    if (a.type == Action::BARE) return CalcV(a, r, r_p, level);  // bare...:175-177
    if (a.type == Action::ILKKA) {                               // ilkka...:125-149
        double q = 0.5 * (r + r_p);
        double x = q + 0.5 * s;
        double y = q - 0.5 * s;
        double du = a.du_xy.Eval(x, y);
        if (a.use_long_range) {
            SetLimits(a.du_long.r_min, a.du_long.r_max, r, r_p);
            du -= 0.5 * a.du_long.r_spline.Eval(r);
            du -= 0.5 * a.du_long.r_spline.Eval(r_p);
        }
        return du;
    }
    // david_pair_action_class.h:126-168
    double q = 0.5 * (r + r_p);
    double z = r - r_p;
    double r_max = a.grid.end;
    SetLimits(a.grid.start, a.grid.end, r, r_p);
    std::vector<double> rv(a.n_val + 1), rpv(a.n_val + 1);
    a.du_kj.Eval(r, rv.data());
    a.du_kj.Eval(r_p, rpv.data());
    double v = 0.5 * (rv[0] + rpv[0]);
    double du = 0.5 * (rv[1] + rpv[1]);
    du += v;
    if (s > 0.0 && q < r_max) {
        std::vector<double> qv(a.n_val + 1);
        a.du_kj.Eval(q, qv.data());
        double z_2 = z * z, s_2 = s * s, i_s_2 = 1. / s_2, s_2_k = s_2;
        for (int k = 1; k <= a.n_order; k++) {
            double z_2_j = 1, current_s = s_2_k;
            for (int j = 0; j <= k; j++) {
                double du_cof = qv[k * (k + 1) / 2 + j + 1];
                du += du_cof * z_2_j * current_s;
                z_2_j *= z_2;
                current_s *= i_s_2;
            }
            s_2_k *= s_2;
        }
    }
    return du;
}

// ---- long range k sums ------------------------------------------------------------------
const std::vector<double> &KWeights(Action &a, int which) {  // 0 u, 1 du, 2 v
    if (a.type == Action::DAVID) return which == 0 ? a.d_u_long_k : (which == 1 ? a.d_du_long_k : a.d_v_long_k);
    if (a.type == Action::BARE) return a.v_long.f_k;
    return which == 0 ? a.u_long.f_k : (which == 1 ? a.du_long.f_k : a.v_long.f_k);
}

// ilkka...:104-122, david...:105-123, bare...:153-172
double CalcULong(Action &a, int b0, int b1, int level) {
    World &w = *a.w;
    const std::vector<cplx> &ra = w.RhoK(a.sa), &rb = w.RhoK(a.sb);
    const std::vector<double> &wk = KWeights(a, 0);
    const int skip = 1 << level;
    const size_t n_k = w.ks.n_k();
    double tot = 0.;
    for (size_t k = 0; k < n_k; k++)
        for (int b = b0; b < b1; b += skip) {
            int bl = w.species[a.sa].bead_loop(b);
            tot += wk[k] * CMag2(ra[bl * n_k + k], rb[bl * n_k + k]);
        }
    if (a.sa != a.sb) tot *= 2.;
    if (a.type == Action::BARE) {
        double level_tau = skip * w.tau;
        return level_tau * tot;
    }
    return tot;
}

// ilkka...:152-169 / :57-74, david...:171-188 / :42-60, bare...:126-143,180-182
double CalcFullLong(Action &a, int which) {
    World &w = *a.w;
    const std::vector<cplx> &ra = w.RhoK(a.sa), &rb = w.RhoK(a.sb);
    const std::vector<double> &wk = KWeights(a, a.type == Action::BARE ? 2 : which);
    const size_t n_k = w.ks.n_k();
    double tot = 0.;
    for (size_t k = 0; k < n_k; k++)
        for (int b = 0; b < w.n_bead; b++) tot += wk[k] * CMag2(ra[b * n_k + k], rb[b * n_k + k]);
    if (a.sa != a.sb) tot *= 2.;
    if (a.type == Action::DAVID)
        return which == 1 ? tot + a.d_du_long_k_0 + a.d_du_long_r_0 : tot + a.d_v_long_k_0 + a.d_v_long_r_0;
    if (a.type == Action::BARE) return tot + a.v_long.k_0 + a.v_long.r_0;
    return which == 1 ? tot + a.du_long.k_0 + a.du_long.r_0 : tot + a.v_long.k_0 + a.v_long.r_0;
}

// ---- pair lists -------------------------------------------------------------------------
// pair_action_class.h:45-60
void AllPairs(Action &a, std::vector<std::pair<int, int>> &pairs) {
    World &w = *a.w;
    const int na = w.species[a.sa].n_part, nb = w.species[a.sb].n_part;
    if (a.sa == a.sb) {
        for (int p = 0; p < na - 1; ++p)
            for (int q = p + 1; q < nb; ++q) pairs.push_back(std::make_pair(p, q));
    } else {
        for (int p = 0; p < na; ++p)
            for (int q = 0; q < nb; ++q) pairs.push_back(std::make_pair(p, q));
    }
}

// pair_action_class.h:63-114
void MovedPairs(Action &a, int n, const int *sp, const int *pi, std::vector<int> &pa, std::vector<int> &pb,
                std::vector<std::pair<int, int>> &pairs) {
    World &w = *a.w;
    for (int i = 0; i < n; ++i) {
        if (sp[i] == a.sa)
            pa.push_back(pi[i]);
        else if (sp[i] == a.sb)
            pb.push_back(pi[i]);
    }
    const int n_a = pa.size(), n_b = pb.size();
    if (n_a == 0 && n_b == 0) return;
    std::vector<int> other_a, other_b;
    for (int p = 0; p < w.species[a.sa].n_part; ++p)
        if (std::find(pa.begin(), pa.end(), p) == pa.end()) other_a.push_back(p);
    for (int p = 0; p < w.species[a.sb].n_part; ++p)
        if (std::find(pb.begin(), pb.end(), p) == pb.end()) other_b.push_back(p);
    if (a.sa == a.sb) {
        for (int p : pa)
            for (int q : other_a) pairs.push_back(std::make_pair(p, q));
        for (int p = 0; p < n_a - 1; ++p)
            for (int q = p + 1; q < n_a; ++q) pairs.push_back(std::make_pair(pa[p], pa[q]));
    } else {
        for (int p : pa)
            for (int q : other_b) pairs.push_back(std::make_pair(p, q));
        for (int p : pb)
            for (int q : other_a) pairs.push_back(std::make_pair(q, p));
        for (int p : pa)
            for (int q : pb) pairs.push_back(std::make_pair(p, q));
    }
}

// ---- the Action API ---------------------------------------------------------------------
// pair_action_class.h:241-264
double DActionDBeta(Action &a) {
    World &w = *a.w;
    if (a.is_constant && !a.is_first_time) return a.dUdB_constant;
    double tot = 0.;
    std::vector<std::pair<int, int>> pairs;
    AllPairs(a, pairs);
    for (size_t i = 0; i < pairs.size(); i++)
        for (int b = 0; b < w.n_bead; ++b) {
            double r, rp, s;
            w.DrDrpDrrp(b, b + 1, a.sa, a.sb, pairs[i].first, pairs[i].second, r, rp, s);
            tot += CalcdUdBeta(a, r, rp, s, 0);
        }
    if (a.use_long_range) tot += CalcFullLong(a, 1);
    if (a.is_first_time) {
        a.is_first_time = false;
        a.dUdB_constant = tot;
    }
    return tot;
}

// pair_action_class.h:369-395 -- two independent minimum-image distances (App. A-6)
double Potential(Action &a) {
    World &w = *a.w;
    if (a.is_constant && !a.is_first_time) return a.potential_constant;
    double tot = 0.;
    std::vector<std::pair<int, int>> pairs;
    AllPairs(a, pairs);
    for (size_t i = 0; i < pairs.size(); i++)
        for (int b = 0; b < w.n_bead; ++b) {
            int bj = b + 1;
            double dr[3];
            w.Dr(w.species[a.sa].R(w.mode, pairs[i].first, b), w.species[a.sb].R(w.mode, pairs[i].second, b), dr);
            double r_mag = Mag(dr, w.n_d);
            w.Dr(w.species[a.sa].R(w.mode, pairs[i].first, bj), w.species[a.sb].R(w.mode, pairs[i].second, bj), dr);
            double r_p_mag = Mag(dr, w.n_d);
            tot += CalcV(a, r_mag, r_p_mag, 0);
        }
    if (a.use_long_range) tot += CalcFullLong(a, 2);
    if (a.is_first_time) {
        a.is_first_time = false;
        a.potential_constant = tot;
    }
    return tot;
}

// pair_action_class.h:267-302
double GetAction(Action &a, int b0, int b1, int n, const int *sp, const int *pi, int level) {
    World &w = *a.w;
    if (level > a.max_level || a.is_constant) return 0.;
    std::vector<int> pa, pb;
    std::vector<std::pair<int, int>> pairs;
    MovedPairs(a, n, sp, pi, pa, pb, pairs);
    if (pairs.size() == 0) return 0.;
    const int skip = 1 << level;
    double tot = 0.;
    for (int b = b0; b < b1; b += skip) {
        int bj = b + skip;
        for (auto &pp : pairs) {
            double r, rp, s;
            w.DrDrpDrrp(b, bj, a.sa, a.sb, pp.first, pp.second, r, rp, s);
            tot += CalcU(a, r, rp, s, level);
        }
    }
    if (a.use_long_range) {
        if (w.species[a.sa].need_update_rho_k && w.mode == 1) w.UpdateRhoK(a.sa, b0, b1, pa, level);
        if (w.species[a.sb].need_update_rho_k && w.mode == 1) w.UpdateRhoK(a.sb, b0, b1, pb, level);
        tot += CalcULong(a, b0, b1, level);
    }
    return tot;
}

// ---- spatial derivatives of the action (contact-density and virial estimators) -------------
// Vector-keeping Path::DrDrpDrrp (path_class.h:153-166).
void DrDrpDrrpVec(World &w, int b0, int b1, int s0, int s1, int p0, int p1, double &r_mag, double &r_p_mag, double &r_r_p_mag,
                  double *r, double *r_p, double *r_r_p) {
    const double *a0 = w.species[s0].R(w.mode, p0, b0), *a1 = w.species[s1].R(w.mode, p1, b0);
    const double *c0 = w.species[s0].R(w.mode, p0, b1), *c1 = w.species[s1].R(w.mode, p1, b1);
    for (int d = 0; d < w.n_d; ++d) {
        r[d] = a1[d] - a0[d];
        r_p[d] = c1[d] - c0[d];
    }
    for (int d = 0; d < w.n_d; ++d) {
        r[d] -= std::nearbyint(r[d] * w.iL) * w.L;
        r_p[d] += std::nearbyint((r[d] - r_p[d]) * w.iL) * w.L;
    }
    for (int d = 0; d < w.n_d; ++d) r_r_p[d] = r[d] - r_p[d];
    for (int d = 0; d < w.n_d; ++d) r_r_p[d] -= std::nearbyint(r_r_p[d] * w.iL) * w.L;
    r_mag = Mag(r, w.n_d);
    r_p_mag = Mag(r_p, w.n_d);
    r_r_p_mag = Mag(r_r_p, w.n_d);
}

// PairAction::CalcGradientU (pair_action_class.h:135-155): central differences, eps = 1e-4,
// of CalcU with respect to the bead (species a, p_i, b_i); Ilkka overrides it analytically
// (ilkka_pair_action_class.h:172-222).
void CalcGradientU(Action &a, int b_i, int b_j, int p_i, int p_j, int level, double *tot) {
    World &w = *a.w;
    if (a.type == Action::ILKKA) {
        double r[3], r_p[3], r_r_p[3], r_mag, r_p_mag, r_r_p_mag;
        DrDrpDrrpVec(w, b_i, b_j, a.sa, a.sb, p_i, p_j, r_mag, r_p_mag, r_r_p_mag, r, r_p, r_r_p);
        double q = 0.5 * (r_mag + r_p_mag);
        double x = q + 0.5 * r_r_p_mag;
        double y = q - 0.5 * r_r_p_mag;
        double u = 0., g[2];
        a.u_xy.EvalVG(x, y, &u, g);
        double rh[3], sh[3];
        for (int d = 0; d < w.n_d; ++d) {
            rh[d] = r[d] / r_mag;
            sh[d] = r_r_p[d] / r_r_p_mag;
        }
        if (r_mag == 0.)
            for (int d = 0; d < w.n_d; ++d) rh[d] = 0.;
        if (r_r_p_mag == 0.)
            for (int d = 0; d < w.n_d; ++d) sh[d] = 0.;
        for (int d = 0; d < w.n_d; ++d) tot[d] = -0.5 * (g[0] * (rh[d] + sh[d]) + g[1] * (rh[d] - sh[d]));
        if (a.use_long_range) {
            SetLimits(a.u_long.r_min, a.u_long.r_max, r_mag, r_p_mag);
            double tmp_u, tmp_du_dr;
            a.u_long.r_spline.EvalVG(r_mag, &tmp_u, &tmp_du_dr);
            for (int d = 0; d < w.n_d; ++d) tot[d] -= 0.5 * tmp_du_dr * rh[d];
        }
        return;
    }
    double *b = w.species[a.sa].R(w.mode, p_i, b_i);
    double r0[3] = {b[0], b[1], b[2]};
    const double eps = 1.e-4;
    double r_mag, r_p_mag, r_r_p_mag;
    for (int d = 0; d < w.n_d; ++d) {
        b[d] = r0[d] + eps;
        w.DrDrpDrrp(b_i, b_j, a.sa, a.sb, p_i, p_j, r_mag, r_p_mag, r_r_p_mag);
        double f1 = CalcU(a, r_mag, r_p_mag, r_r_p_mag, level);
        b[d] = r0[d] - eps;
        w.DrDrpDrrp(b_i, b_j, a.sa, a.sb, p_i, p_j, r_mag, r_p_mag, r_r_p_mag);
        double f2 = CalcU(a, r_mag, r_p_mag, r_r_p_mag, level);
        tot[d] = (f1 - f2) / (2. * eps);
        b[d] = r0[d];
    }
}

// CalcGradientULong(b_0, b_1, p_i, level): zero in the base class (pair_action_class.h:163-165),
// the k-space force on bead (species a, p_i) for Ilkka (ilkka_pair_action_class.h:231-249).
void CalcGradientULong(Action &a, int b_0, int b_1, int p_i, int level, double *tot) {
    World &w = *a.w;
    for (int d = 0; d < w.n_d; ++d) tot[d] = 0.;
    if (a.type != Action::ILKKA) return;
    const std::vector<cplx> &rb = w.RhoK(a.sb);
    const size_t n_k = w.ks.n_k();
    const int skip = 1 << level;
    std::vector<cplx> ra(n_k);
    for (int b_i = b_0; b_i < b_1; b_i += skip) {
        w.BeadRhoK(w.species[a.sa].R(w.mode, p_i, b_i), ra.data());  // the bead's own rho_k (bead_class.h:125-133)
        const int bl = w.species[a.sb].bead_loop(b_i);
        for (size_t k = 0; k < n_k; k++) {
            const double f = ra[k].real() * rb[bl * n_k + k].imag() - ra[k].imag() * rb[bl * n_k + k].real();
            for (int d = 0; d < w.n_d; ++d) tot[d] += (a.u_long.f_k[k] * w.ks.vecs[k * w.n_d + d]) * f;
        }
    }
    if (a.sa != a.sb)
        for (int d = 0; d < w.n_d; ++d) tot[d] *= 2.;
}

// PairAction::CalcLaplacianU (pair_action_class.h:168-203): second central differences of CalcU.
double CalcLaplacianU(Action &a, int b_i, int b_j, int p_i, int p_j, int level) {
    World &w = *a.w;
    double *b = w.species[a.sa].R(w.mode, p_i, b_i);
    double r0[3] = {b[0], b[1], b[2]};
    const double eps = 1.e-4;
    double r_mag, r_p_mag, r_r_p_mag, tot = 0.;
    w.DrDrpDrrp(b_i, b_j, a.sa, a.sb, p_i, p_j, r_mag, r_p_mag, r_r_p_mag);
    double f0 = CalcU(a, r_mag, r_p_mag, r_r_p_mag, level);
    for (int d = 0; d < w.n_d; ++d) {
        b[d] = r0[d] + eps;
        w.DrDrpDrrp(b_i, b_j, a.sa, a.sb, p_i, p_j, r_mag, r_p_mag, r_r_p_mag);
        double fp1 = CalcU(a, r_mag, r_p_mag, r_r_p_mag, level);
        b[d] = r0[d] - eps;
        w.DrDrpDrrp(b_i, b_j, a.sa, a.sb, p_i, p_j, r_mag, r_p_mag, r_r_p_mag);
        double fm1 = CalcU(a, r_mag, r_p_mag, r_r_p_mag, level);
        tot += (fp1 + fm1 - 2 * f0) / (eps * eps);
        b[d] = r0[d];
    }
    return tot;
}

// PairAction::GetActionGradient (pair_action_class.h:305-337): the forward link (b_i, b_i + skip)
// and the backward link (b_i, b_i - skip + n_bead) of every pair; the long-range term once per
// PAIR (sic), always for the pair's species-a particle.
void GetActionGradient(Action &a, int b0, int b1, int n, const int *sp, const int *pi, int level, double *tot) {
    World &w = *a.w;
    for (int d = 0; d < w.n_d; ++d) tot[d] = 0.;
    if (level > a.max_level || a.is_constant) return;
    std::vector<int> pa, pb;
    std::vector<std::pair<int, int>> pairs;
    MovedPairs(a, n, sp, pi, pa, pb, pairs);
    if (pairs.size() == 0) return;
    const int skip = 1 << level;
    double g[3];
    for (int b_i = b0; b_i < b1; b_i += skip) {
        int b_j = b_i + skip;
        int b_k = b_i - skip + w.species[a.sa].n_bead;
        for (auto &pp : pairs) {
            CalcGradientU(a, b_i, b_j, pp.first, pp.second, level, g);
            for (int d = 0; d < w.n_d; ++d) tot[d] += g[d];
            CalcGradientU(a, b_i, b_k, pp.first, pp.second, level, g);
            for (int d = 0; d < w.n_d; ++d) tot[d] += g[d];
        }
    }
    if (a.use_long_range)
        for (auto &pp : pairs) {
            CalcGradientULong(a, b0, b1, pp.first, level, g);
            for (int d = 0; d < w.n_d; ++d) tot[d] += g[d];
        }
}

// PairAction::GetActionLaplacian (pair_action_class.h:340-366); no long-range part ("FIXME" there).
double GetActionLaplacian(Action &a, int b0, int b1, int n, const int *sp, const int *pi, int level) {
    World &w = *a.w;
    if (level > a.max_level || a.is_constant) return 0.;
    std::vector<int> pa, pb;
    std::vector<std::pair<int, int>> pairs;
    MovedPairs(a, n, sp, pi, pa, pb, pairs);
    if (pairs.size() == 0) return 0.;
    const int skip = 1 << level;
    double tot = 0.;
    for (int b_i = b0; b_i < b1; b_i += skip) {
        int b_j = b_i + skip;
        int b_k = b_i - skip + w.species[a.sa].n_bead;
        for (auto &pp : pairs)
            tot += CalcLaplacianU(a, b_i, b_j, pp.first, pp.second, level) + CalcLaplacianU(a, b_i, b_k, pp.first, pp.second, level);
    }
    return tot;
}

// ---- construction -----------------------------------------------------------------------
void LoadLongRange(World &w, LongRange &lr, const pimc_long_range &t) {
    // ilkka_pair_action_class.h:283-316 (same block three times; bare...:53-86)
    lr.present = true;
    orc::Grid g = orc::MakeGeneralGrid(t.f_r.r, t.f_r.n);
    lr.r_min = g.start;
    lr.r_max = g.end;
    lr.r_spline.Create(g, t.f_r.f);
    lr.r_0 = t.f_r_0;
    lr.k_0 = t.f_k_0;
    lr.k_tab.assign(t.k, t.k + t.n_k);
    lr.fk_tab.assign(t.f_k, t.f_k + t.n_k);
    lr.f_k.assign(w.ks.n_k(), 0.);
    for (size_t k_i = 0; k_i < w.ks.n_k(); ++k_i)
        for (int k_t = 0; k_t < t.n_k; ++k_t)
            if (FEqual(w.ks.mags[k_i], t.k[k_t], 1.e-8)) lr.f_k[k_i] = t.f_k[k_t];
}

// pair_action_class.h:204-238
Action *NewAction(World &w, Action::Type type, int sa, int sb, int max_level, int use_lr, double k_cut) {
    Action *a = new Action;
    a->type = type;
    a->w = &w;
    a->sa = sa;
    a->sb = sb;
    a->max_level = max_level;
    a->use_long_range = use_lr != 0;
    if (a->use_long_range) {
        a->k_cut = k_cut;
        w.ks.Setup(k_cut);
        w.InitRhoK(sa);
        w.InitRhoK(sb);
    }
    a->is_constant = ((sa == sb) && (w.species[sa].n_part == 1 || w.species[sa].lambda == 0.));
    a->is_first_time = true;
    return a;
}

// ilkka...:421-431, bare...:89-95, david...:347-358
void ScaleConstants(World &w, Action &a, double &k0, double &r0) {
    const Species &A = w.species[a.sa], &B = w.species[a.sb];
    if (a.sa == a.sb) {
        k0 *= 0.5 * A.n_part * B.n_part * w.n_bead;
        r0 *= -0.5 * A.n_part * w.n_bead;
    } else {
        k0 *= A.n_part * B.n_part * w.n_bead;
        r0 *= 0.;
    }
}

}  // namespace

extern "C" {

void *orc_create(const pimc_config *cfg) {
    World *w = new World;
    w->n_d = cfg->n_d;
    w->n_bead = cfg->n_bead;
    w->beta = cfg->beta;
    w->pbc = cfg->pbc != 0;
    if (w->pbc) {  // path_class.h:33-42
        w->L = cfg->L;
        w->iL = 1. / w->L;
        w->vol = std::pow(w->L, w->n_d);
    } else {
        w->L = 0.;
        w->iL = 0.;
        w->vol = 1.;
    }
    w->tau = w->beta / (1. * w->n_bead);
    w->ks.n_d = w->n_d;
    w->ks.L = w->L;
    w->ks.cutoff = 0.;
    w->species.resize(cfg->n_species);
    for (int s = 0; s < cfg->n_species; ++s) {
        Species &sp = w->species[s];
        sp.n_part = cfg->n_part[s];
        sp.n_bead = cfg->n_bead;
        sp.n_d = cfg->n_d;
        sp.lambda = cfg->lambda[s];
        sp.r.assign((size_t)sp.n_part * sp.n_bead * sp.n_d, 0.);
        sp.r_c = sp.r;
    }
    return w;
}
void orc_destroy(void *h) { delete (World *)h; }

int orc_kspace_setup(void *h, double k_cut) {
    World *w = (World *)h;
    w->ks.Setup(k_cut);
    for (size_t s = 0; s < w->species.size(); ++s) w->InitRhoK(s);
    return (int)w->ks.n_k();
}
void orc_kspace_get(void *h, int32_t *idx, double *mags) {
    World *w = (World *)h;
    for (size_t k = 0; k < w->ks.n_k(); ++k) {
        for (int d = 0; d < w->n_d; ++d) idx[k * w->n_d + d] = w->ks.indices[k * w->n_d + d] - w->ks.max_index[d];
        mags[k] = w->ks.mags[k];
    }
}

void orc_set_positions(void *h, int sp, const double *R) {
    World *w = (World *)h;
    Species &s = w->species[sp];
    std::memcpy(s.r.data(), R, s.r.size() * sizeof(double));
    s.r_c = s.r;
    w->InitRhoK(sp);
    s.need_update_rho_k = true;
}
void orc_get_positions(void *h, int sp, int mode, double *R) {
    World *w = (World *)h;
    Species &s = w->species[sp];
    std::memcpy(R, (mode ? s.r : s.r_c).data(), s.r.size() * sizeof(double));
}
void orc_rhok(void *h, int sp, int mode, double *out) {
    World *w = (World *)h;
    const std::vector<cplx> &rho = mode ? w->species[sp].rho_k : w->species[sp].rho_k_c;
    for (size_t i = 0; i < rho.size(); ++i) {
        out[2 * i] = rho[i].real();
        out[2 * i + 1] = rho[i].imag();
    }
}
void orc_set_mode(void *h, int mode) { ((World *)h)->mode = mode; }

int orc_action_create_ilkka(void *h, int sa, int sb, const pimc_ilkka_tables *t, int max_level, int use_lr, double k_cut) {
    World &w = *(World *)h;
    Action *a = NewAction(w, Action::ILKKA, sa, sb, max_level, use_lr, k_cut);
    // ilkka_pair_action_class.h:266-280, 318-332
    a->u_xy.Create(orc::MakeGeneralGrid(t->u_xy.x, t->u_xy.n_x), orc::MakeGeneralGrid(t->u_xy.y, t->u_xy.n_y), t->u_xy.f);
    a->du_xy.Create(orc::MakeGeneralGrid(t->du_xy.x, t->du_xy.n_x), orc::MakeGeneralGrid(t->du_xy.y, t->du_xy.n_y), t->du_xy.f);
    // :371-382
    orc::Grid gv = orc::MakeGeneralGrid(t->v_r.r, t->v_r.n);
    a->r_v_min = gv.start;
    a->r_v_max = gv.end;
    a->v_r.Create(gv, t->v_r.f);
    if (a->use_long_range) {
        LoadLongRange(w, a->u_long, t->u_long);
        LoadLongRange(w, a->du_long, t->du_long);
        LoadLongRange(w, a->v_long, t->v_long);
        ScaleConstants(w, *a, a->du_long.k_0, a->du_long.r_0);
        ScaleConstants(w, *a, a->v_long.k_0, a->v_long.r_0);
    }
    w.actions.emplace_back(a);
    return (int)w.actions.size() - 1;
}

int orc_action_create_bare(void *h, int sa, int sb, const pimc_bare_tables *t, int max_level, int use_lr, double k_cut) {
    World &w = *(World *)h;
    Action *a = NewAction(w, Action::BARE, sa, sb, max_level, use_lr, k_cut);
    a->is_coulomb = t->is_coulomb != 0;
    orc::Grid gv = orc::MakeGeneralGrid(t->v_r.r, t->v_r.n);
    a->r_v_min = gv.start;
    a->r_v_max = gv.end;
    a->v_r.Create(gv, t->v_r.f);
    if (a->use_long_range) {
        LoadLongRange(w, a->v_long, t->v_long);
        ScaleConstants(w, *a, a->v_long.k_0, a->v_long.r_0);
    }
    w.actions.emplace_back(a);
    return (int)w.actions.size() - 1;
}

int orc_action_create_david(void *h, int sa, int sb, const pimc_david_tables *t, int max_level, int use_lr) {
    World &w = *(World *)h;
    Action *a = NewAction(w, Action::DAVID, sa, sb, max_level, 0, 0.);
    // david_pair_action_class.h:209-230
    if (t->grid_type == PIMC_GRID_LOG)
        a->grid = orc::MakeLogGrid(t->r_start, t->r_end, t->n_grid);
    else if (t->grid_type == PIMC_GRID_LINEAR)
        a->grid = orc::MakeLinearGrid(t->r_start, t->r_end, t->n_grid);
    else
        a->grid = orc::MakeGeneralGrid(t->grid_points, t->n_grid);
    const int n_grid = t->n_grid, n_tau = t->n_tau;
    // :232-245
    bool tau_found = false;
    for (int i = 0; i < n_tau; ++i)
        if (std::fabs(t->taus[i] - w.tau) < 1.0e-6) tau_found = true;
    if (!tau_found || n_tau != 1) {
        delete a;
        return -1;
    }
    a->n_order = t->n_order;
    a->n_val = 1;  // :252-254
    for (int i = 1; i <= a->n_order; ++i) a->n_val += 1 + i;
    const int n_val = a->n_val;
    // :256-283 -- cube(n_val, n_grid, n_tau) filled in file order; value 0 of every knot is
    // the potential, and the LAST grid point of every value is left at 0 (loop to n_grid-1)
    for (int which = 0; which < 2; ++which) {
        const double *data = which == 0 ? t->u_kj : t->du_kj_dbeta;
        orc::MultiSpline1D &ms = which == 0 ? a->u_kj : a->du_kj;
        ms.Create(a->grid, n_val + 1);
        std::vector<double> tmp(n_grid);
        for (int v = 0; v < n_val + 1; ++v) {
            for (int g = 0; g < n_grid; ++g) {
                if (g == n_grid - 1)
                    tmp[g] = 0.;
                else if (v == 0)
                    tmp[g] = t->potential[g];
                else
                    tmp[g] = data[(size_t)(v - 1) + (size_t)n_val * g];  // tau index 0
            }
            ms.Set(v, tmp.data());
        }
    }
    if (use_lr) {  // :311-358
        a->use_long_range = true;
        // base-class part of the constructor (pair_action_class.h:216-222) with the System k_cut
        w.InitRhoK(sa);
        w.InitRhoK(sb);
        const size_t n_k = w.ks.n_k();
        std::vector<double> v_long_k(t->n_k);
        for (int i = 0; i < t->n_k; ++i) v_long_k[i] = t->u_k[i] / w.vol;
        a->d_u_long_k.assign(n_k, 0.);
        a->d_du_long_k.assign(n_k, 0.);
        double v_long_k_0 = 0.;
        for (int kv = 0; kv < t->n_k; ++kv) {
            if (FEqual(0., t->k_points[kv], 1.e-8)) v_long_k_0 = v_long_k[kv];
            for (size_t k_i = 0; k_i < n_k; ++k_i)
                if (FEqual(w.ks.mags[k_i], t->k_points[kv], 1.e-8)) {
                    a->d_u_long_k[k_i] = v_long_k[kv] * w.tau;
                    a->d_du_long_k[k_i] = v_long_k[kv];
                }
        }
        // CalcVLong (david...:42-60) indexes v_long_k -- an array over table SHELLS -- with the
        // k-VECTOR index, reading past its end whenever there are more vectors than shells
        // (undefined behaviour under ARMA_NO_DEBUG).  The oracle defines the evident intent:
        // the shell value matched by |k|, as for du_long_k.
        a->d_v_long_k.assign(n_k, 0.);
        for (int kv = 0; kv < t->n_k; ++kv)
            for (size_t k_i = 0; k_i < n_k; ++k_i)
                if (FEqual(w.ks.mags[k_i], t->k_points[kv], 1.e-8)) a->d_v_long_k[k_i] = v_long_k[kv];
        a->d_v_long_r_0 = t->v_image;
        a->d_u_long_r_0 = a->d_v_long_r_0 * w.tau;
        a->d_du_long_r_0 = a->d_v_long_r_0;
        a->d_v_long_k_0 = v_long_k_0;
        a->d_u_long_k_0 = v_long_k_0 * w.tau;
        a->d_du_long_k_0 = v_long_k_0;
        ScaleConstants(w, *a, a->d_du_long_k_0, a->d_du_long_r_0);
        ScaleConstants(w, *a, a->d_v_long_k_0, a->d_v_long_r_0);
    }
    w.actions.emplace_back(a);
    return (int)w.actions.size() - 1;
}

double orc_dbeta(void *h, int a) {
    World *w = (World *)h;
    w->mode = 1;
    return DActionDBeta(*w->actions[a]);
}
double orc_potential(void *h, int a) {
    World *w = (World *)h;
    w->mode = 1;
    return Potential(*w->actions[a]);
}
double orc_get_action(void *h, int a, int mode, int b0, int b1, int n, const int *sp, const int *pi, int level) {
    World *w = (World *)h;
    w->mode = mode;
    return GetAction(*w->actions[a], b0, b1, n, sp, pi, level);
}
void orc_action_gradient(void *h, int a, int mode, int b0, int b1, int n, const int *sp, const int *pi, int level, double *out) {
    World *w = (World *)h;
    w->mode = mode;
    GetActionGradient(*w->actions[a], b0, b1, n, sp, pi, level, out);
}
double orc_action_laplacian(void *h, int a, int mode, int b0, int b1, int n, const int *sp, const int *pi, int level) {
    World *w = (World *)h;
    w->mode = mode;
    return GetActionLaplacian(*w->actions[a], b0, b1, n, sp, pi, level);
}
// PermBisectIterative::UpdatePermTable (perm_bisect_iterative_class.h:10-30); relative != 0: the
// table of PermBisectTable (perm_bisect_table_class.h:36-50).  Unpermuted path.
void orc_perm_table(void *h, int sp, int bead0, int n_bisect_beads, double epsilon, int relative, double *t) {
    World *w = (World *)h;
    Species &s = w->species[sp];
    const double i_4_lambda_tau = 1. / (4. * s.lambda * w->tau);           // bisect_class.h:147
    const double i_4_lambda_tau_n_bisect_beads = i_4_lambda_tau / n_bisect_beads;
    const double log_epsilon = std::log(epsilon);
    int b1 = bead0 + n_bisect_beads;
    while (b1 >= s.n_bead) b1 -= s.n_bead;
    for (int i = 0; i < s.n_part; i++) {
        double dr_ii[3];
        w->Dr(s.R(w->mode, i, bead0), s.R(w->mode, i, b1), dr_ii);
        for (int j = 0; j < s.n_part; j++) {
            double dr_ij[3];
            w->Dr(s.R(w->mode, i, bead0), s.R(w->mode, j, b1), dr_ij);
            double exponent = relative ? (-Dot(dr_ij, dr_ij, w->n_d) + Dot(dr_ii, dr_ii, w->n_d)) * i_4_lambda_tau_n_bisect_beads
                                       : (-Dot(dr_ij, dr_ij, w->n_d)) * i_4_lambda_tau_n_bisect_beads;
            t[(size_t)i * s.n_part + j] = exponent > log_epsilon ? std::exp(exponent) : 0.;
        }
    }
}
void orc_calc_pair(void *h, int a, int which, int n, const double *r, const double *rp, const double *s, int level, double *out) {
    World *w = (World *)h;
    Action &act = *w->actions[a];
    for (int i = 0; i < n; ++i)
        out[i] = which == 0 ? CalcU(act, r[i], rp[i], s[i], level)
                            : (which == 1 ? CalcdUdBeta(act, r[i], rp[i], s[i], level) : CalcV(act, r[i], rp[i], level));
}
double orc_calc_long(void *h, int a, int which, int b0, int b1, int level) {
    World *w = (World *)h;
    Action &act = *w->actions[a];
    if (which == 0) return CalcULong(act, b0, b1, level);
    return CalcFullLong(act, which);
}
void orc_dr_drp_drrp(void *h, int b0, int b1, int sa, int sb, int p0, int p1, double *out3) {
    World *w = (World *)h;
    w->DrDrpDrrp(b0, b1, sa, sb, p0, p1, out3[0], out3[1], out3[2]);
}

// NEW-mode Bead::SetR (bisect_class.h:93)
void orc_propose(void *h, int sp, int p, int b_first, int n, const double *newR) {
    World *w = (World *)h;
    Species &s = w->species[sp];
    for (int i = 0; i < n; ++i) std::memcpy(s.R(1, p, s.bead_loop(b_first + i) ), newR + (size_t)i * w->n_d, w->n_d * sizeof(double));
}
// Bisect::Accept / Reject (bisect_class.h:24-36,127-139)
void orc_finish_move(void *h, int sp, int p, int b0, int b1, int accept) {
    World *w = (World *)h;
    Species &s = w->species[sp];
    const size_t n_k = w->ks.n_k();
    for (int b = b0; b <= b1; ++b) {
        double *rn = s.R(1, p, b), *ro = s.R(0, p, b);
        if (accept)
            std::memcpy(ro, rn, w->n_d * sizeof(double));
        else
            std::memcpy(rn, ro, w->n_d * sizeof(double));
    }
    if (!s.rho_k.empty())
        for (int b = b0; b < b1; ++b) {
            int bl = s.bead_loop(b);
            for (size_t k = 0; k < n_k; ++k) {
                if (accept)
                    s.rho_k_c[bl * n_k + k] = s.rho_k[bl * n_k + k];
                else
                    s.rho_k[bl * n_k + k] = s.rho_k_c[bl * n_k + k];
            }
        }
    // PairAction::Accept / Reject (pair_action_class.h:406-419)
    for (auto &a : w->actions)
        if (a->use_long_range) {
            w->species[a->sa].need_update_rho_k = true;
            w->species[a->sb].need_update_rho_k = true;
        }
}

// LinearGrid::ReverseMap (observable_class.h:56-62) with CreateGrid (:30-40)
void orc_gofr_bins(double r_min, double r_max, int n_r, int n, const double *r, uint32_t *bins) {
    double dr = (r_max - r_min) / (n_r - 1.);
    double d_ir = 1. / dr;
    for (int i = 0; i < n; ++i) bins[i] = (uint32_t)std::nearbyint((r[i] - r_min) * d_ir - 0.5);
}

// PairCorrelation::Accumulate (pair_correlation_class.h:15-28)
void orc_gofr(void *h, int sa, int sb, double r_min, double r_max, int n_r, double cofactor, double *y, uint64_t *counts) {
    World *w = (World *)h;
    w->mode = 1;
    double dr = (r_max - r_min) / (n_r - 1.);
    double d_ir = 1. / dr;
    Action tmp;
    tmp.w = w;
    tmp.sa = sa;
    tmp.sb = sb;
    std::vector<std::pair<int, int>> pairs;
    AllPairs(tmp, pairs);
    for (int b = 0; b < w->n_bead; ++b)
        for (auto &p : pairs) {
            double d[3];
            w->Dr(w->species[sa].R(1, p.first, b), w->species[sb].R(1, p.second, b), d);
            uint32_t i = (uint32_t)std::nearbyint((Mag(d, w->n_d) - r_min) * d_ir - 0.5);
            if (i < (uint32_t)n_r) {
                y[i] = y[i] + 1. * cofactor;
                if (counts) counts[i] += 1;
            }
        }
}

// StructureFactor::Accumulate (structure_factor_class.h:15-32)
void orc_sofk(void *h, int sa, int sb, double k_cut, double cofactor, double *sk) {
    World *w = (World *)h;
    w->mode = 1;
    const std::vector<cplx> &ra = w->RhoK(sa), &rb = w->RhoK(sb);
    const size_t n_k = w->ks.n_k();
    for (size_t k = 0; k < n_k; k++)
        if (w->ks.mags[k] < k_cut)
            for (int b = 0; b < w->n_bead; ++b) sk[k] += cofactor * CMag2(ra[b * n_k + k], rb[b * n_k + k]);
}

// Stand-alone spline hooks for tests/test_oracle_spline.py
void orc_spline1d_eval(int n, const double *grid, const double *data, int m, const double *x, double *out) {
    orc::Spline1D s;
    s.Create(orc::MakeGeneralGrid(grid, n), data);
    for (int i = 0; i < m; ++i) out[i] = s.Eval(x[i]);
}
void orc_spline1d_coefs(int n, const double *grid, const double *data, double *coefs) {
    orc::Spline1D s;
    s.Create(orc::MakeGeneralGrid(grid, n), data);
    for (int i = 0; i < n + 2; ++i) coefs[i] = s.coefs[i];
}
void orc_spline2d_eval(int nx, int ny, const double *gx, const double *gy, const double *data, int m, const double *x,
                       const double *y, double *out) {
    orc::Spline2D s;
    s.Create(orc::MakeGeneralGrid(gx, nx), orc::MakeGeneralGrid(gy, ny), data);
    for (int i = 0; i < m; ++i) out[i] = s.Eval(x[i], y[i]);
}
void orc_spline2d_coefs(int nx, int ny, const double *gx, const double *gy, const double *data, double *coefs) {
    orc::Spline2D s;
    s.Create(orc::MakeGeneralGrid(gx, nx), orc::MakeGeneralGrid(gy, ny), data);
    for (int i = 0; i < (nx + 2) * (ny + 2); ++i) coefs[i] = s.coefs[i];
}
int orc_grid_reverse_map(int n, const double *grid, int m, const double *x, int32_t *out) {
    orc::Grid g = orc::MakeGeneralGrid(grid, n);
    for (int i = 0; i < m; ++i) out[i] = g.ReverseMap(x[i]);
    return 0;
}

}  // extern "C"
