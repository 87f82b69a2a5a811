// Permuting bisection on the device: PermBisectIterative::Attempt / Accept / Reject
// (src/events/moves/single_species_move/bisect/perm_bisect/perm_bisect_iterative_class.h:10-222 on
// perm_bisect_class.h:32-82) for every walker of a context at once, in the dense representation of
// positions by particle LABEL plus the permutation at the beta seam,
// next[c][p] = label of the bead that follows (p, n_bead - 1) of walker c.  A chain keeps its label inside the
// path and continues as next[label] across the seam.
//
//   perm_select_kernel         first bead, first particle and SelectCycleIterative (:33-111): rows of
//                              UpdatePermTable's t_ij (:10-30) computed on the fly with the chains followed across
//                              the seam, "continue?" / "which next particle?" from two uniforms per step, weight
//   perm_sample_kernel         PermuteBeads (the last link of member i leads to the old end point of member i + 1),
//                              Levy construction of every member level by level, Kinetic along the links, the
//                              Metropolis tests of the levels above 0 starting from -log(weight); leaves the
//                              proposal as label windows in the species' proposal slots (the pair actions pair
//                              LABELS at equal slices: the reference's GetBead(p, b), App. A-4)
//   (pair_window_kernel, rhok_delta_kernel, ksum_kernel from kernels.cuh: every pair action in OLD and NEW mode)
//   perm_decide_commit_kernel  level-0 Metropolis test, Move::Accept of the label windows and rho_k
//   perm_apply_kernel          PermBisect::AssignParticleLabels (perm_bisect_class.h:47-56) after an accepted cycle:
//                              from the slice after the window's last moved bead to the end of the path the
//                              members' labels rotate, and so does the seam permutation
//
// Philox slots (counter = attempt, clone, slot), shared with the host mirror simpimc_b200/perm_moves.py:
//   slot 0                first bead (words 0-1), first particle of the cycle (words 2-3)
//   slot 1 + k            step k of the cycle selection: continue? (words 0-1), next particle (words 2-3)
//   slot 16 + 128 i + s   Levy displacement slots s (the single-particle move's layout, s >= 1) of cycle member i
//   slot 1040 + level     Metropolis uniform of the level
// Integer results (cycle members, labels, permutation) are bit-exact against the reference run on the same numbers
// (tests/test_stream_ref_cpu.py pins the mirror to the reference, tests/test_gpu_perm.py the kernels to the mirror).
#ifndef SIMPIMC_B200_PERM_CUH_
#define SIMPIMC_B200_PERM_CUH_

#include "mc.cuh"

namespace pimc {

constexpr int kPermMaxLen = 8;  // longest cycle built on the device (the reference has no bound; longer ones are counted and skipped)
constexpr uint32_t kPermSlotCycle0 = 1u, kPermSlotLevy0 = 16u, kPermSlotLevyStride = 128u;
constexpr uint32_t kPermSlotMetro0 = kPermSlotLevy0 + kPermMaxLen * kPermSlotLevyStride;
static_assert(kMaxPropSlots >= 2 * kPermMaxLen, "members plus their end points' labels must fit the proposal slots");

struct PermSelectArgs {
    PathView pv;
    const double *R;           // committed positions of the species
    int N;
    const int32_t *next;       // [C][N] seam permutation
    int n_bisect_beads;
    double i_4_lambda_tau_n;   // (1 / (4 lambda tau)) / n_bisect_beads
    double log_epsilon;
    uint32_t seed_lo, seed_hi, attempt_lo, attempt_hi;
    int32_t *b0;               // [C] first slice of the window (out)
    int32_t *n_perm;           // [C] cycle length; 0: the selection stopped (no bisection); -1: longer than kPermMaxLen
    int32_t *particles;        // [C][kPermMaxLen]
    double *weight;            // [C]
    int32_t *n_steps;          // [C] selection steps taken (two uniforms each)
};

/// One warp per walker; the row of the table lives in shared memory ([2][N] doubles).
static __global__ void __launch_bounds__(32) perm_select_kernel(const PermSelectArgs a) {
    extern __shared__ double row[];  // t(p, .) then t_c(p, .): [2][N]
    const PathView &pv = a.pv;
    const int c = blockIdx.x, lane = threadIdx.x, N = a.N;
    double *t_row = row, *t_c = row + N;
    uint32_t rnd[4];
    Philox4x32(a.attempt_lo, a.attempt_hi, (uint32_t)c, 0u, a.seed_lo, a.seed_hi, rnd);
    int bead0 = (int)(UniformFromBits(rnd[0], rnd[1]) * pv.M);
    bead0 = bead0 < pv.M ? bead0 : pv.M - 1;
    const int be_raw = bead0 + a.n_bisect_beads;
    const bool wrapped = be_raw >= pv.M;
    const int bs = bead0, be = wrapped ? be_raw - pv.M : be_raw;
    int p0 = (int)(UniformFromBits(rnd[2], rnd[3]) * N);
    p0 = p0 < N ? p0 : N - 1;
    int p = p0, n = 0, steps = 0, ps[kPermMaxLen];
    double weight = 1.;
    int result = 0;
    while (true) {
        if (n == kPermMaxLen) {
            result = -1;
            break;
        }
        ps[n++] = p;
        // row p: Dr(r_p(b0), r_j(b1)), the end bead of chain j taken across the seam when the window rolls over
        double rp[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) rp[d] = a.R[PosIndex(pv, N, c, p, d, bs)];
        for (int j = lane; j < N; j += 32) {
            const int lj = wrapped ? a.next[(size_t)c * N + j] : j;
            double x[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) x[d] = MinImage(rp[d] - a.R[PosIndex(pv, N, c, lj, d, be)], pv.box);
            const double e = (-__dadd_rn(__dadd_rn(__dmul_rn(x[0], x[0]), __dmul_rn(x[2], x[2])), __dmul_rn(x[1], x[1]))) * a.i_4_lambda_tau_n;
            const double t = e > a.log_epsilon ? exp(e) : 0.;
            t_row[j] = t;
            t_c[j] = t;
        }
        __syncwarp();
        uint32_t ru[4];
        Philox4x32(a.attempt_lo, a.attempt_hi, (uint32_t)c, kPermSlotCycle0 + (uint32_t)steps, a.seed_lo, a.seed_hi, ru);
        ++steps;
        int nxt = p, stop = 0;
        double t_next = 0., t_self = 0.;
        if (lane == 0) {  // the reference's sequential sums and scan, in index order
            for (int i = 0; i < n; ++i) t_c[ps[i]] = 0.;
            t_c[p0] = t_row[p0];
            double Q_p = 0., Q_p_c = 0.;
            for (int i = 0; i < N; ++i) {
                Q_p = __dadd_rn(Q_p, t_row[i]);
                Q_p_c = __dadd_rn(Q_p_c, t_c[i]);
            }
            if (__ddiv_rn(Q_p_c, Q_p) < UniformFromBits(ru[0], ru[1])) {
                stop = 1;
            } else {
                const double x = UniformFromBits(ru[2], ru[3]);
                double t_Q = 0.;
                for (int i = 0; i < N; ++i) {
                    t_Q = __dadd_rn(t_Q, __ddiv_rn(t_c[i], Q_p_c));
                    if (t_Q > x) {
                        nxt = i;
                        break;
                    }
                }
                t_next = t_row[nxt];
                t_self = t_row[p];
            }
        }
        stop = __shfl_sync(0xffffffffu, stop, 0);
        if (stop) {
            result = 0;
            n = 0;
            break;
        }
        nxt = __shfl_sync(0xffffffffu, nxt, 0);
        t_next = __shfl_sync(0xffffffffu, t_next, 0);
        t_self = __shfl_sync(0xffffffffu, t_self, 0);
        weight = __dmul_rn(weight, __ddiv_rn(t_next, t_self));
        __syncwarp();
        p = nxt;
        if (p == p0) {
            result = n;
            break;
        }
    }
    if (lane == 0) {
        a.b0[c] = bead0;
        a.n_perm[c] = result;
        a.weight[c] = result > 0 ? weight : 0.;
        a.n_steps[c] = steps;
        for (int i = 0; i < kPermMaxLen; ++i) a.particles[(size_t)c * kPermMaxLen + i] = (result > 0 && i < result) ? ps[i] : -1;
    }
}

struct PermSampleArgs {
    PathView pv;
    const double *R;
    int N;
    const int32_t *next;
    double lambda, tau;
    int n_level;
    int with_kinetic;
    FreeSplineSet fs_move;  // the move's own n_images (sampling probabilities), tau_s = tau 2^level / 2
    FreeSplineSet fs_kin;   // the Kinetic action's, tau_s = tau 2^level
    uint32_t seed_lo, seed_hi, attempt_lo, attempt_hi;
    const int32_t *b0;         // [C]
    const int32_t *n_perm;     // [C]
    const int32_t *particles;  // [C][kPermMaxLen]
    const double *weight;      // [C]
    // outputs: the proposal as label windows, kMaxPropSlots slots (unused ones carry label -1)
    double *P;                 // [slot][C][nb - 1][3]
    int32_t *P_particle;       // [slot][C]
    int32_t *P_first;          // [slot][C]
    double *partial;           // [C] log_sample_ratio - kinetic change at level 0 + previous level's change
    double *logu0;             // [C]
    int32_t *alive;            // [C] a cycle closed and passed the levels above 0
    long long *perm_attempt;   // [C][kPermMaxLen], += 1 at (cycle length - 1) when a cycle closed (perm_bisect_iterative_class.h:131)
};

/// One warp per walker: the lanes load the members' chains and draw the Levy displacements, lane 0 walks the levels.
static __global__ void __launch_bounds__(32) perm_sample_kernel(const PermSampleArgs a) {
    __shared__ double s_old[kPermMaxLen][kMaxBisectBeads + 1][3], s_new[kPermMaxLen][kMaxBisectBeads + 1][3];
    __shared__ double s_del[kPermMaxLen][kMaxBisectBeads][3], s_d2[kPermMaxLen][kMaxBisectBeads], s_logu[8];
    __shared__ int s_ps[kPermMaxLen], s_nx[kPermMaxLen], s_lab[kMaxPropSlots], s_nlab;
    const PathView &pv = a.pv;
    const int c = blockIdx.x, lane = threadIdx.x, N = a.N, M = pv.M;
    const int n = a.n_perm[c], nb = 1 << a.n_level, n_prop = nb - 1, bead0 = a.b0[c];
    int first = bead0 + 1;
    first = first >= M ? first - M : first;
    if (n <= 0) {  // the selection stopped (or overflowed): no bisection, Reject
        if (lane < kMaxPropSlots) {
            a.P_particle[(size_t)lane * pv.C + c] = -1;
            a.P_first[(size_t)lane * pv.C + c] = first;
        }
        if (lane == 0) {
            a.alive[c] = 0;
            a.partial[c] = 0.;
            a.logu0[c] = 0.;
        }
        return;
    }
    if (lane < n) {
        const int p = a.particles[(size_t)c * kPermMaxLen + lane];
        s_ps[lane] = p;
        s_nx[lane] = a.next[(size_t)c * N + p];
    }
    __syncwarp();
    // OLD chains follow the committed links
    const int per = (nb + 1) * 3;
    for (int t = lane; t < n * per; t += 32) {
        const int i = t / per, r = t - i * per, k = r / 3, d = r - k * 3;
        const int bg = bead0 + k;
        const int l = bg < M ? s_ps[i] : s_nx[i], b = bg < M ? bg : bg - M;
        s_old[i][k][d] = a.R[PosIndex(pv, N, c, l, d, b)];
    }
    __syncwarp();
    // NEW chains end on the next member's old end point (PermuteBeads, perm_bisect_class.h:58-82)
    for (int t = lane; t < n * per; t += 32) {
        const int i = t / per, r = t - i * per, k = r / 3, d = r - k * 3;
        s_new[i][k][d] = (k == nb) ? s_old[i + 1 == n ? 0 : i + 1][nb][d] : s_old[i][k][d];
    }
    for (int t = lane; t < n * n_prop; t += 32) {  // Levy displacement sigma * normal of member i, midpoint ib
        const int i = t / n_prop, ib = 1 + t - i * n_prop;
        const int level = __ffs(ib) - 1, skip = 1 << level;
        const int idx = (ib - skip) >> (level + 1);
        const uint32_t slot = kPermSlotLevy0 + kPermSlotLevyStride * (uint32_t)i + SweepSlotStart(level, a.n_level, nb) + 2u * (uint32_t)idx;
        uint32_t r0[4], r1[4];
        Philox4x32(a.attempt_lo, a.attempt_hi, (uint32_t)c, slot, a.seed_lo, a.seed_hi, r0);
        Philox4x32(a.attempt_lo, a.attempt_hi, (uint32_t)c, slot + 1, a.seed_lo, a.seed_hi, r1);
        const double ua = UniformFromBits(r0[0], r0[1]), ub = UniformFromBits(r0[2], r0[3]);
        const double uc = UniformFromBits(r1[0], r1[1]), ud = UniformFromBits(r1[2], r1[3]);
        const double ra = sqrt(-2. * log(ua)), rc = sqrt(-2. * log(uc));
        double sb, cb, sd, cd;
        sincospi(2. * ub, &sb, &cb);
        sincospi(2. * ud, &sd, &cd);
        (void)sd;
        const double nrm[3] = {ra * cb, ra * sb, rc * cd};
        const double sigma = sqrt(a.lambda * (a.tau * skip));
        double d2 = 0., delv[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double del = PutInBox1(sigma * nrm[d], pv.box);
            s_del[i][ib][d] = del;
            delv[d] = del;
            d2 += del * del;
        }
        s_d2[i][ib] = a.fs_move.n_images ? -FreeLogRho(a.fs_move.s[level], delv) : d2;
    }
    if (lane < a.n_level) {  // Metropolis uniform of level = lane
        uint32_t ru[4];
        Philox4x32(a.attempt_lo, a.attempt_hi, (uint32_t)c, kPermSlotMetro0 + (uint32_t)lane, a.seed_lo, a.seed_hi, ru);
        s_logu[lane] = log(UniformFromBits(ru[0], ru[1]));
    }
    __syncwarp();
    if (lane == 0) {
        bool alive = true;
        double prev_change = -log(a.weight[c]), partial = 0.;  // perm_bisect_iterative_class.h:141
        for (int level = a.n_level - 1; level >= 0; --level) {
            const int skip = 1 << level;
            const double level_tau = a.tau * skip;
            const double i4lt_sample = 1. / (4. * a.lambda * (0.5 * level_tau));
            const double i4lt_kin = 1. / (4. * a.lambda * level_tau);
            double old_lp = 0., new_lp = 0.;
            for (int i = 0; i < n; ++i) {
                double(*oldb)[3] = s_old[i];
                double(*newb)[3] = s_new[i];
                for (int ia = 0; ia < nb; ia += 2 * skip) {
                    const int ib = ia + skip, ic = ia + 2 * skip;
                    double d2_old = 0., delo[3];
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const double rbar_old = oldb[ia][d] + 0.5 * PutInBox1(oldb[ic][d] - oldb[ia][d], pv.box);
                        const double del_old = PutInBox1(oldb[ib][d] - rbar_old, pv.box);
                        delo[d] = del_old;
                        d2_old += del_old * del_old;
                        const double rbar_new = newb[ia][d] + 0.5 * PutInBox1(newb[ic][d] - newb[ia][d], pv.box);
                        newb[ib][d] = rbar_new + s_del[i][ib][d];
                    }
                    if (a.fs_move.n_images) {
                        old_lp += FreeLogRho(a.fs_move.s[level], delo);
                        new_lp -= s_d2[i][ib];
                    } else {
                        old_lp -= d2_old * i4lt_sample;
                        new_lp -= s_d2[i][ib] * i4lt_sample;
                    }
                }
            }
            double old_kin = 0., new_kin = 0.;
            if (a.with_kinetic) {  // Kinetic::GetAction follows the links (GetNextBead): the members' chains
                for (int i = 0; i < n; ++i) {
                    double(*oldb)[3] = s_old[i];
                    double(*newb)[3] = s_new[i];
                    for (int ia = 0; ia < nb; ia += skip) {
                        double d2o = 0., d2n = 0., ov[3], nv[3];
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            const double o = PutInBox1(oldb[ia][d] - oldb[ia + skip][d], pv.box);
                            const double nn = PutInBox1(newb[ia][d] - newb[ia + skip][d], pv.box);
                            ov[d] = o;
                            nv[d] = nn;
                            d2o += o * o;
                            d2n += nn * nn;
                        }
                        if (a.fs_kin.n_images) {
                            old_kin -= FreeLogRho(a.fs_kin.s[level + 1], ov);
                            new_kin -= FreeLogRho(a.fs_kin.s[level + 1], nv);
                        } else {
                            old_kin += d2o * i4lt_kin;
                            new_kin += d2n * i4lt_kin;
                        }
                    }
                }
            }
            const double lsr = -new_lp + old_lp;
            const double change = new_kin - old_kin;
            if (level > 0) {
                if (lsr - change + prev_change < s_logu[level]) alive = false;
                prev_change = change;
            } else {
                partial = lsr - change + prev_change;
            }
        }
        // the listed particles: the members and, when the window rolls over the seam, the labels their chains end on
        // (bead_f(i)->GetP() of the OLD links), sorted
        int nlab = 0;
        const bool roll_over = bead0 + nb >= M;
        for (int pass = 0; pass < (roll_over ? 2 : 1); ++pass)
            for (int i = 0; i < n; ++i) {
                const int l = pass ? s_nx[i] : s_ps[i];
                int pos = 0;
                while (pos < nlab && s_lab[pos] < l) ++pos;
                if (pos < nlab && s_lab[pos] == l) continue;
                for (int q = nlab; q > pos; --q) s_lab[q] = s_lab[q - 1];
                s_lab[pos] = l;
                ++nlab;
            }
        s_nlab = nlab;
        a.alive[c] = alive ? 1 : 0;
        a.partial[c] = partial;
        a.logu0[c] = s_logu[0];
        a.perm_attempt[(size_t)c * kPermMaxLen + (n - 1)] += 1;
    }
    __syncwarp();
    const int nlab = s_nlab;
    if (lane < kMaxPropSlots) {
        a.P_particle[(size_t)lane * pv.C + c] = lane < nlab ? s_lab[lane] : -1;
        a.P_first[(size_t)lane * pv.C + c] = first;
    }
    // label windows: the committed beads of the label overlaid with the moved beads of the chain that runs through it
    for (int t = lane; t < nlab * n_prop * 3; t += 32) {
        const int sl = t / (n_prop * 3), r = t - sl * n_prop * 3, j = r / 3, d = r - j * 3;
        const int l = s_lab[sl], k = j + 1, bg = bead0 + k, b = bg < M ? bg : bg - M;
        double x = a.R[PosIndex(pv, N, c, l, d, b)];
        for (int i = 0; i < n; ++i)
            if ((bg < M ? s_ps[i] : s_nx[i]) == l) x = s_new[i][k][d];
        a.P[(((size_t)sl * pv.C + c) * n_prop + j) * 3 + d] = x;
    }
}

struct PermPairArgs {
    PathView pv;
    SpeciesView A_old, A_new, B_old, B_new;  // the action's two species without / with the label windows as their proposal
    int same;                   // species_a == species_b
    int moved_is_a;             // the moved species is species a of the action (always, when same)
    const int32_t *part;        // [kMaxPropSlots][C] listed labels of the moved species, packed, -1 = unused slot
    const int32_t *b0;          // [C]
    const int32_t *alive;       // [C]
    int n_links;
    FastTable FT;
    const unsigned char *fast_tables;
    double *partial_old, *partial_new;   // [C][n_links]
};

/// PairAction::GetAction over the window for the listed labels (GenerateParticlePairs, pair_action_class.h:63-114) in
/// OLD and NEW mode at once, through the fast Ilkka evaluation with its tables read from global memory: one CTA per
/// (walker, link), threads over the partners.  Same pair set and order as pair_window_kernel (kernels.cuh), which stays
/// the path of every other action type.
static __global__ void __launch_bounds__(128) perm_pair_fast_kernel(const PermPairArgs a) {
    __shared__ double red[128 / 32];
    const PathView &pv = a.pv;
    const GlobalTab tb(a.fast_tables);
    for (int item = blockIdx.x; item < pv.C * a.n_links; item += gridDim.x) {
        const int c = item / a.n_links, j = item - c * a.n_links;
        if (!a.alive[c]) continue;  // uniform over the CTA
        const int bg = a.b0[c] + j;
        int la[kMaxPropSlots], n_l = 0;
        for (int i = 0; i < kMaxPropSlots; ++i) {
            const int l = a.part[(size_t)i * pv.C + c];
            if (l < 0) break;
            la[n_l++] = l;
        }
        double acc_o = 0., acc_n = 0.;
        const SpeciesView &M_old = a.moved_is_a ? a.A_old : a.B_old, &M_new = a.moved_is_a ? a.A_new : a.B_new;
        const SpeciesView &Q_old = a.moved_is_a ? a.B_old : a.A_old, &Q_new = a.moved_is_a ? a.B_new : a.A_new;
        // an unlisted partner has no proposal: its two beads come straight from the committed array, once for both modes
        const int s0 = WrapSlice(pv, bg) - pv.slice_lo, s1 = WrapSlice(pv, bg + 1) - pv.slice_lo;
        for (int i = 0; i < n_l; ++i) {
            const int m = la[i];
            double m0o[3], m1o[3], m0n[3], m1n[3];
            LoadPos(pv, M_old, c, m, bg, 0, m0o);
            LoadPos(pv, M_old, c, m, bg + 1, 0, m1o);
            LoadPos(pv, M_new, c, m, bg, 1, m0n);
            LoadPos(pv, M_new, c, m, bg + 1, 1, m1n);
            for (int q = threadIdx.x; q < Q_old.N; q += blockDim.x) {
                bool q_listed = false;
                if (a.same) {
                    if (q == m) continue;
                    bool earlier = false;  // (listed_i, listed_j) once: only from the lower list index
                    for (int i2 = 0; i2 < n_l; ++i2) {
                        earlier = earlier || (i2 < i && la[i2] == q);
                        q_listed = q_listed || la[i2] == q;
                    }
                    if (earlier) continue;
                }
                double q0[3], q1[3], r, rp, s;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    q0[d] = Q_old.R[PosIndex(pv, Q_old.N, c, q, d, s0)];
                    q1[d] = Q_old.R[PosIndex(pv, Q_old.N, c, q, d, s1)];
                }
                DrDrpDrrpFast(m0o, q0, m1o, q1, pv.box, r, rp, s);
                acc_o += FastIlkkaEval(tb, a.FT, r, rp, s);
                if (q_listed) {  // the partner is a listed label itself: its NEW beads come from its label window
                    LoadPos(pv, Q_new, c, q, bg, 1, q0);
                    LoadPos(pv, Q_new, c, q, bg + 1, 1, q1);
                }
                DrDrpDrrpFast(m0n, q0, m1n, q1, pv.box, r, rp, s);
                acc_n += FastIlkkaEval(tb, a.FT, r, rp, s);
            }
        }
        const double to = BlockSum<128>(acc_o, red);
        const double tn = BlockSum<128>(acc_n, red);
        if (threadIdx.x == 0) {
            a.partial_old[item] = to;
            a.partial_new[item] = tn;
        }
        __syncthreads();
    }
}

struct PermDecideArgs {
    PathView pv;
    int N, n_k, nb;
    const int32_t *alive;
    const double *partial, *logu0;      // [C]
    const double *pair_parts;           // [action][mode (OLD, NEW)][C][nb] link sums of pair_window_kernel
    int n_pair_actions;
    const double *lr_old, *lr_new;      // [C] (nullptr: no long-range action)
    const double *P;
    const int32_t *P_particle, *b0;
    const double2 *drho;                // [C][nb][n_k] (nullptr: no long-range action)
    double *R;
    double2 *rho;
    const int32_t *n_perm;
    int32_t *accept;                    // [C]
    long long *n_accept;                // [C]
    long long *perm_accept;             // [C][kPermMaxLen]
};

/// Level-0 Metropolis test (perm_bisect_iterative_class.h:196-208) + Move::Accept of the label windows: one CTA per walker.
static __global__ void __launch_bounds__(256) perm_decide_commit_kernel(const PermDecideArgs a) {
    __shared__ int s_acc;
    const PathView &pv = a.pv;
    const int c = blockIdx.x, C = pv.C, nb = a.nb, n_prop = nb - 1;
    if (threadIdx.x == 0) {
        int acc = 0;
        if (a.alive[c]) {
            double old_action = 0., new_action = 0.;
            for (int t = 0; t < a.n_pair_actions; ++t) {
                const double *po = a.pair_parts + ((size_t)(2 * t) * C + c) * nb, *pn = a.pair_parts + ((size_t)(2 * t + 1) * C + c) * nb;
                double so = 0., sn = 0.;
                for (int j = 0; j < nb; ++j) {
                    so += po[j];
                    sn += pn[j];
                }
                old_action += so;
                new_action += sn;
            }
            if (a.lr_old) {
                old_action += a.lr_old[c];
                new_action += a.lr_new[c];
            }
            acc = (a.partial[c] - (new_action - old_action)) < a.logu0[c] ? 0 : 1;
        }
        a.accept[c] = acc;
        a.n_accept[c] += acc;
        if (acc) a.perm_accept[(size_t)c * kPermMaxLen + (a.n_perm[c] - 1)] += 1;
        s_acc = acc;
    }
    __syncthreads();
    if (!s_acc) return;
    const int bead0 = a.b0[c];
    for (int sl = 0; sl < kMaxPropSlots; ++sl) {
        const int l = a.P_particle[(size_t)sl * C + c];
        if (l < 0) break;  // the labels are packed into the leading slots
        for (int t = threadIdx.x; t < n_prop * 3; t += blockDim.x) {
            const int j = t / 3, d = t - j * 3;
            int bg = bead0 + 1 + j;
            bg = bg >= pv.M ? bg - pv.M : bg;
            a.R[PosIndex(pv, a.N, c, l, d, bg)] = a.P[(((size_t)sl * C + c) * n_prop + j) * 3 + d];
        }
    }
    if (a.drho) {
        for (int t = threadIdx.x; t < nb * a.n_k; t += blockDim.x) {
            const int j = t / a.n_k, k = t - j * a.n_k;
            int bg = bead0 + j;
            bg = bg >= pv.M ? bg - pv.M : bg;
            double2 *dst = a.rho + ((size_t)c * pv.Mloc + bg) * a.n_k + k;
            const double2 d = a.drho[((size_t)c * nb + j) * a.n_k + k];
            dst->x += d.x;
            dst->y += d.y;
        }
    }
}

struct PermApplyArgs {
    PathView pv;
    double *R;
    int N;
    int32_t *next;              // [C][N], updated
    const int32_t *b0;          // [C]
    int n_bisect_beads;
    const int32_t *accept;      // [C]
    const int32_t *n_perm;      // [C]: cycles of fewer than two particles are left alone
    const int32_t *particles;   // [C][kPermMaxLen] labels of the members at slice b0
    double *scratch;            // [C][kPermMaxLen * 3 * M]
};

/// One CTA per walker: rows (member, dim) x slices (e, M) copied to the walker's scratch, written back rotated by one member.
static __global__ void __launch_bounds__(256) perm_apply_kernel(const PermApplyArgs a) {
    __shared__ int labels[kPermMaxLen], old_next[kPermMaxLen];
    const PathView &pv = a.pv;
    const int c = blockIdx.x, N = a.N, M = pv.M;
    const int n = a.n_perm[c];
    if (!a.accept[c] || n < 2) return;
    double *stage = a.scratch + (size_t)c * kPermMaxLen * 3 * M;
    const int last = a.b0[c] + a.n_bisect_beads - 1;   // the window's last moved slice, before the wrap
    const bool wrapped = last >= M;
    const int e = wrapped ? last - M : last;
    if (threadIdx.x < n) {
        const int p = a.particles[(size_t)c * kPermMaxLen + threadIdx.x];
        const int l = wrapped ? a.next[(size_t)c * N + p] : p;
        labels[threadIdx.x] = l;
        old_next[threadIdx.x] = a.next[(size_t)c * N + l];
    }
    __syncthreads();
    const int n_sl = M - 1 - e;   // slices e + 1 .. M - 1
    for (int t = threadIdx.x; t < n * 3 * n_sl; t += blockDim.x) {
        const int i = t / (3 * n_sl), r = t - i * 3 * n_sl, d = r / n_sl, s = r - d * n_sl;
        stage[t] = a.R[PosIndex(pv, N, c, labels[i], d, e + 1 + s)];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n * 3 * n_sl; t += blockDim.x) {
        const int i = t / (3 * n_sl), r = t - i * 3 * n_sl, d = r / n_sl, s = r - d * n_sl;
        const int src = (i + 1 == n ? 0 : i + 1);
        a.R[PosIndex(pv, N, c, labels[i], d, e + 1 + s)] = stage[(size_t)src * 3 * n_sl + r];
    }
    if (threadIdx.x < n) a.next[(size_t)c * N + labels[threadIdx.x]] = old_next[threadIdx.x + 1 == n ? 0 : threadIdx.x + 1];
}

/// next[c][p] = p: the unpermuted path.
static __global__ void perm_identity_kernel(int32_t *__restrict__ next, int C, int N) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < C * N) next[t] = t % N;
}

}  // namespace pimc

#endif  // SIMPIMC_B200_PERM_CUH_
