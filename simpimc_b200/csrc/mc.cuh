// Device-resident bisection moves: Bisect::Attempt / Accept / Reject
// (src/events/moves/single_species_move/bisect/bisect_class.h:24-139) for every clone of a
// context at once, without a host round trip per attempt.
//
//   bisect_sample_kernel      pick particle and window, Levy construction level by level, the
//                             free-particle (Kinetic, n_images = 0) action and sampling
//                             probabilities, Metropolis tests of the levels above 0 (pair
//                             actions return 0 there, pair_action_class.h:269)
//   pair_window_both_kernel   sum over the window's links and all partner particles of the
//                             pair action in OLD and in NEW mode (PairAction::GetAction,
//                             pair_action_class.h:267-302, as called at bisect_class.h:103-108)
//   (rhok_delta_kernel + ksum_kernel from kernels.cuh give the long-range part)
//   bisect_decide_kernel      level-0 Metropolis test (bisect_class.h:110-115) -> accept flags
//   (commit_positions_kernel / commit_rhok_kernel apply Move::Accept)
//
// Random numbers: Philox4x32-10 keyed by the caller's seed, counter = (attempt, clone, slot);
// the reference draws from std::mt19937, so sampled runs agree with it statistically, not
// stream for stream.  simpimc_b200/moves.py holds the host mirror of the same stream.
#ifndef SIMPIMC_B200_MC_CUH_
#define SIMPIMC_B200_MC_CUH_

#include "kernels.cuh"
#include "pair_fast.cuh"
#include "kinetic.cuh"

namespace pimc {

__host__ __device__ __forceinline__ void PhiloxRound(uint32_t c[4], uint32_t k[2]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c[1] ^ k[0], n2 = hi0 ^ c[3] ^ k[1];
    c[0] = n0;
    c[1] = lo1;
    c[2] = n2;
    c[3] = lo0;
}

/// Philox4x32-10 (Salmon et al., SC'11): out = f_key(counter).
__host__ __device__ __forceinline__ void Philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                                    uint32_t out[4]) {
    uint32_t c[4] = {c0, c1, c2, c3}, k[2] = {k0, k1};
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        PhiloxRound(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    out[0] = c[0];
    out[1] = c[1];
    out[2] = c[2];
    out[3] = c[3];
}

/// Uniform double in (0, 1] from two 32-bit words (53 bits; never 0, so log(u) is finite).
__host__ __device__ __forceinline__ double UniformFromBits(uint32_t a, uint32_t b) {
    const uint64_t x = ((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6);
    return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
}

constexpr int kMaxBisectBeads = 32;  // n_level <= 5

struct BisectArgs {
    PathView pv;
    const double *R;   // committed positions of the species
    int N;
    double lambda, tau;
    int n_level;
    int with_kinetic;
    FreeSplineSet fs_move;  // Bisect's rho_free_splines (bisect_class.h:158-163): s[level], tau_s = tau 2^level / 2
    FreeSplineSet fs_kin;   // Kinetic's (kinetic_class.h:16-24): s[level + 1], tau_s = tau 2^level
    int b0_lo, b0_count;  // window starts are uniform in [b0_lo, b0_lo + b0_count): the whole path, or a shard's interior windows
    uint32_t seed_lo, seed_hi;
    uint32_t attempt_lo, attempt_hi;
    // outputs (proposal of the species + per-clone scalars)
    double *P;              // [C][nb-1][3]
    int32_t *P_particle;    // [C]
    int32_t *P_first;       // [C]
    int32_t *b0;            // [C]
    double *partial;        // [C] log_sample_ratio - kinetic change at level 0 + previous level's change
    double *logu0;          // [C]
    int32_t *alive;         // [C] passed the levels above 0
    double *pair_old, *pair_new, *lr_old, *lr_new;  // [C] accumulators, zeroed here
};

__device__ __forceinline__ double PutInBox1(double d, const Box &bx) { return d - rint(d * bx.iL) * bx.L; }

/// First Philox slot of a level: slot 0 picks particle and window, then every level from the
/// top takes two slots per midpoint and one for its Metropolis uniform.
__device__ __forceinline__ uint32_t SweepSlotStart(int level, int n_level, int nb) {
    uint32_t s = 1;
    for (int l = n_level - 1; l > level; --l) s += 2u * (uint32_t)(nb >> (l + 1)) + 1u;
    return s;
}

constexpr int kSampleWarps = 4;  // clones per CTA

/// One WARP per clone: the Philox draws (Levy displacements of all midpoints, Metropolis
/// uniforms) and the window's bead loads run in parallel over the lanes, lane 0 then walks the
/// levels (bisect_class.h:69-98) on shared memory, and the lanes write the proposal.
static __global__ void __launch_bounds__(kSampleWarps * 32) bisect_sample_kernel(const BisectArgs a) {
    __shared__ double s_old[kSampleWarps][kMaxBisectBeads + 1][3], s_new[kSampleWarps][kMaxBisectBeads + 1][3];
    __shared__ double s_del[kSampleWarps][kMaxBisectBeads][3], s_d2[kSampleWarps][kMaxBisectBeads], s_logu[kSampleWarps][8];
    __shared__ int s_alive[kSampleWarps];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * kSampleWarps + w;
    const PathView &pv = a.pv;
    if (c >= pv.C) return;  // whole warp
    const int nb = 1 << a.n_level;
    uint32_t rnd[4];
    Philox4x32(a.attempt_lo, a.attempt_hi, (uint32_t)c, 0u, a.seed_lo, a.seed_hi, rnd);
    // rng.UnifRand(n) - 1: uniform integer in [0, n)
    int p_i = (int)(UniformFromBits(rnd[0], rnd[1]) * a.N);
    p_i = p_i < a.N ? p_i : a.N - 1;
    int bead0 = (int)(UniformFromBits(rnd[2], rnd[3]) * a.b0_count);
    bead0 = a.b0_lo + (bead0 < a.b0_count ? bead0 : a.b0_count - 1);
    if (pv.vdiv > 1) {
        // disjoint windows of one walker: window w starts at b0_lo + offset + w * nb, the offset drawn once per walker
        // and attempt from the stream of its window 0 (b0_count = number of admissible offsets)
        const int c0 = (c / pv.vdiv) * pv.vdiv, wi = c - c0;
        uint32_t r0[4];
        Philox4x32(a.attempt_lo, a.attempt_hi, (uint32_t)c0, 0u, a.seed_lo, a.seed_hi, r0);
        int off = (int)(UniformFromBits(r0[2], r0[3]) * a.b0_count);
        off = off < a.b0_count ? off : a.b0_count - 1;
        bead0 = a.b0_lo + off + wi * nb;
        bead0 = WrapSlice(pv, bead0);
    }
    double(*oldb)[3] = s_old[w];
    double(*newb)[3] = s_new[w];
    for (int t = lane; t < (nb + 1) * 3; t += 32) {
        const int j = t / 3, d = t - j * 3;
        int bg = bead0 + j;
        bg = WrapSlice(pv, bg);
        const double x = a.R[PosIndex(pv, a.N, RealClone(pv, c), p_i, d, bg - pv.slice_lo)];
        oldb[j][d] = x;
        newb[j][d] = x;
    }
    if (lane >= 1 && lane < nb) {  // Levy displacement sigma * normal of midpoint ib = lane
        const int ib = lane;
        const int level = __ffs(ib) - 1, skip = 1 << level;
        const int idx = (ib - skip) >> (level + 1);
        const uint32_t slot = SweepSlotStart(level, a.n_level, nb) + 2u * (uint32_t)idx;
        uint32_t r0[4], r1[4];
        Philox4x32(a.attempt_lo, a.attempt_hi, (uint32_t)c, slot, a.seed_lo, a.seed_hi, r0);
        Philox4x32(a.attempt_lo, a.attempt_hi, (uint32_t)c, slot + 1, a.seed_lo, a.seed_hi, r1);
        // Box-Muller: three of the four normals
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
            s_del[w][ib][d] = del;
            delv[d] = del;
            d2 += del * del;
        }
        // with images: -log rho_free of the new displacement (bisect_class.h:94); without: |delta|^2
        s_d2[w][ib] = a.fs_move.n_images ? -FreeLogRho(a.fs_move.s[level], delv) : d2;
    }
    if (lane < a.n_level) {  // Metropolis uniform of level = lane
        const uint32_t slot = SweepSlotStart(lane, a.n_level, nb) + 2u * (uint32_t)(nb >> (lane + 1));
        uint32_t ru[4];
        Philox4x32(a.attempt_lo, a.attempt_hi, (uint32_t)c, slot, a.seed_lo, a.seed_hi, ru);
        s_logu[w][lane] = log(UniformFromBits(ru[0], ru[1]));
    }
    __syncwarp();
    if (lane == 0) {
        bool alive = true;
        double prev_change = 0., partial = 0.;
        for (int level = a.n_level - 1; level >= 0; --level) {
            const int skip = 1 << level;
            const double level_tau = a.tau * skip;
            // FreeSpline(L, n_images = 0, lambda, 0.5 * tau * 2^level): log rho = -|r|^2 / (4 lambda (level_tau / 2))
            const double i4lt_sample = 1. / (4. * a.lambda * (0.5 * level_tau));
            const double i4lt_kin = 1. / (4. * a.lambda * level_tau);
            double old_lp = 0., new_lp = 0.;
            for (int ia = 0; ia < nb; ia += 2 * skip) {
                const int ib = ia + skip, ic = ia + 2 * skip;
                double d2_old = 0., delo[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    // RBar(bead_c, bead_a) = r_a + 0.5 * Dr(r_c, r_a)   (path_class.h:124)
                    const double rbar_old = oldb[ia][d] + 0.5 * PutInBox1(oldb[ic][d] - oldb[ia][d], pv.box);
                    const double del_old = PutInBox1(oldb[ib][d] - rbar_old, pv.box);
                    delo[d] = del_old;
                    d2_old += del_old * del_old;
                    const double rbar_new = newb[ia][d] + 0.5 * PutInBox1(newb[ic][d] - newb[ia][d], pv.box);
                    newb[ib][d] = rbar_new + s_del[w][ib][d];
                }
                if (a.fs_move.n_images) {  // FreeSpline with images (free_spline_class.h:75-83)
                    old_lp += FreeLogRho(a.fs_move.s[level], delo);
                    new_lp -= s_d2[w][ib];
                } else {
                    old_lp -= d2_old * i4lt_sample;
                    new_lp -= s_d2[w][ib] * i4lt_sample;
                }
            }
            double old_kin = 0., new_kin = 0.;
            if (a.with_kinetic) {  // Kinetic::GetAction (kinetic_class.h:105-122)
                for (int ia = 0; ia < nb; ia += skip) {
                    double d2o = 0., d2n = 0., ov[3], nv[3];
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const double o = PutInBox1(oldb[ia][d] - oldb[ia + skip][d], pv.box);
                        const double n = PutInBox1(newb[ia][d] - newb[ia + skip][d], pv.box);
                        ov[d] = o;
                        nv[d] = n;
                        d2o += o * o;
                        d2n += n * n;
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
            const double lsr = -new_lp + old_lp;
            const double change = new_kin - old_kin;
            if (level > 0) {
                if (lsr - change + prev_change < s_logu[w][level]) alive = false;
                prev_change = change;
            } else {
                partial = lsr - change + prev_change;
            }
        }
        s_alive[w] = alive ? 1 : 0;
        a.P_particle[c] = p_i;
        int first = bead0 + 1;
        first = WrapSlice(pv, first);
        a.P_first[c] = first;
        a.b0[c] = bead0;
        a.partial[c] = partial;
        a.logu0[c] = s_logu[w][0];
        a.alive[c] = alive ? 1 : 0;
        a.pair_old[c] = 0.;
        a.pair_new[c] = 0.;
        a.lr_old[c] = 0.;
        a.lr_new[c] = 0.;
    }
    __syncwarp();
    const bool alive = s_alive[w] != 0;
    const int n_prop = nb - 1;
    for (int t = lane; t < n_prop * 3; t += 32) {
        const int j = t / 3, d = t - j * 3;
        a.P[((size_t)c * n_prop + j) * 3 + d] = alive ? newb[j + 1][d] : oldb[j + 1][d];
    }
}

struct WindowBothArgs {
    PathView pv;
    const double *R_moved;    // committed positions of the moved species
    int N_moved;
    const double *R_partner;  // committed positions of the partner species
    int N_partner;
    int same;                 // partner species == moved species (skip the moved particle itself)
    const double *P;          // proposal [C][n_links-1][3]
    const int32_t *P_particle;
    const int32_t *b0;
    const int32_t *alive;
    int n_links;
    // tables: fast Ilkka block (global memory) or the general blob
    FastTable FT;
    const unsigned char *fast_tables;
    PairTable T;
    const double *blob;
    double *out_old, *out_new;  // [C], added to
};

constexpr int kWindowThreads = 256;

/// One CTA per clone: thread t walks partner particles q = t, t + 256, ...; for each it evaluates
/// the window's links in OLD and NEW positions of the moved particle.
template <int ATYPE, bool FAST>
static __global__ void __launch_bounds__(kWindowThreads) pair_window_both_kernel(const WindowBothArgs a) {
    __shared__ double pold[kMaxBisectBeads + 1][3], pnew[kMaxBisectBeads + 1][3];
    __shared__ double red[2][kWindowThreads / 32];
    const PathView &pv = a.pv;
    const int c = blockIdx.x;
    if (!a.alive[c]) return;  // rejected above level 0: the decision does not need the sums
    const int p = a.P_particle[c], bead0 = a.b0[c], nl = a.n_links;
    for (int t = threadIdx.x; t < (nl + 1) * 3; t += blockDim.x) {
        const int j = t / 3, d = t - j * 3;
        int bg = bead0 + j;
        bg = WrapSlice(pv, bg);
        const double x = a.R_moved[PosIndex(pv, a.N_moved, RealClone(pv, c), p, d, bg - pv.slice_lo)];
        pold[j][d] = x;
        pnew[j][d] = (j >= 1 && j < nl) ? a.P[((size_t)c * (nl - 1) + (j - 1)) * 3 + d] : x;
    }
    __syncthreads();
    double acc_old = 0., acc_new = 0.;
    for (int q = threadIdx.x; q < a.N_partner; q += blockDim.x) {
        if (a.same && q == p) continue;
        double q0[3], q1[3];
        int bg = bead0;
#pragma unroll
        for (int d = 0; d < 3; ++d) q0[d] = a.R_partner[PosIndex(pv, a.N_partner, RealClone(pv, c), q, d, bg - pv.slice_lo)];
        for (int j = 0; j < nl; ++j) {
            int bn = bg + 1;
            bn = WrapSlice(pv, bn);
#pragma unroll
            for (int d = 0; d < 3; ++d) q1[d] = a.R_partner[PosIndex(pv, a.N_partner, RealClone(pv, c), q, d, bn - pv.slice_lo)];
            double r, rp, s;
            if (FAST) {
                DrDrpDrrpFast(pold[j], q0, pold[j + 1], q1, pv.box, r, rp, s);
                acc_old += FastIlkkaEval(GlobalTab(a.fast_tables), a.FT, r, rp, s);
                DrDrpDrrpFast(pnew[j], q0, pnew[j + 1], q1, pv.box, r, rp, s);
                acc_new += FastIlkkaEval(GlobalTab(a.fast_tables), a.FT, r, rp, s);
            } else {
                DrDrpDrrp(pold[j], q0, pold[j + 1], q1, pv.box, r, rp, s);
                acc_old += PairEval<ATYPE, WHICH_U>(a.blob, a.T, r, rp, s);
                DrDrpDrrp(pnew[j], q0, pnew[j + 1], q1, pv.box, r, rp, s);
                acc_new += PairEval<ATYPE, WHICH_U>(a.blob, a.T, r, rp, s);
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) q0[d] = q1[d];
            bg = bn;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_old += __shfl_down_sync(0xffffffffu, acc_old, o);
        acc_new += __shfl_down_sync(0xffffffffu, acc_new, o);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        red[0][wid] = acc_old;
        red[1][wid] = acc_new;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double to = 0., tn = 0.;
        for (int i = 0; i < kWindowThreads / 32; ++i) {
            to += red[0][i];
            tn += red[1][i];
        }
        a.out_old[c] += to;
        a.out_new[c] += tn;
    }
}

/// Level-0 Metropolis test: log_accept = log_sample_ratio - (new_action - old_action) + previous change.
static __global__ void bisect_decide_kernel(int C, const int32_t *__restrict__ alive, const double *__restrict__ partial,
                                     const double *__restrict__ logu0, const double *__restrict__ pair_old,
                                     const double *__restrict__ pair_new, const double *__restrict__ lr_old,
                                     const double *__restrict__ lr_new, int32_t *__restrict__ accept, int64_t *__restrict__ n_accept) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    int acc = 0;
    if (alive[c]) {
        const double old_action = pair_old[c] + lr_old[c], new_action = pair_new[c] + lr_new[c];
        const double log_accept = partial[c] - (new_action - old_action);
        acc = log_accept < logu0[c] ? 0 : 1;
    }
    accept[c] = acc;
    n_accept[c] += acc;
}

// ---------------------------------------------------------------- staged window kernel
constexpr int kWinFastThreads = 1024;

constexpr int kWinTeams = 8;                                  // independent sub-CTA teams (named barriers)
constexpr int kWinTeamThreads = kWinFastThreads / kWinTeams;
constexpr int kWinTeamWarps = kWinTeamThreads / 32;

/// Same sums as pair_window_both_kernel<ILKKA, fast> with the fast tables staged in shared
/// memory once per CTA: persistent CTAs (one per SM) split into kWinTeams teams of 4 warps, each
/// walking its own clones (a whole CTA per clone leaves one iteration per warp between barriers
/// at N = 128).  Inside a clone, n_links consecutive lanes own the links of one (moved particle,
/// partner) pair and groups of 32 / n_links partners per warp keep every warp converged: OLD and
/// NEW come from one set of partner loads and the long-range spline is evaluated twice per link
/// pair instead of four times (FastIlkkaEvalWindow).
static __global__ void __launch_bounds__(kWinFastThreads, 1) pair_window_fast_kernel(const WindowBothArgs a) {
    extern __shared__ __align__(16) unsigned char wsm[];
    __shared__ double pold[kWinTeams][kMaxBisectBeads + 1][3], pnew[kWinTeams][kMaxBisectBeads + 1][3];
    __shared__ double red[kWinTeams][2][kWinTeamWarps];
    __shared__ unsigned long long stage_bar;
    StageBlockTma(wsm, a.fast_tables, a.FT.n_bytes, &stage_bar);  // the only CTA-wide synchronisation
    const SharedTab wtab(wsm);
    const PathView &pv = a.pv;
    const int nl = a.n_links;
    const int team = threadIdx.x / kWinTeamThreads, tid = threadIdx.x - team * kWinTeamThreads;
    const int lane = tid & 31, warp = tid >> 5;
    const int j = lane & (nl - 1), sub = lane / nl, per_warp = 32 / nl;
    const int n_groups = (a.N_partner + per_warp - 1) / per_warp;
    auto team_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(kWinTeamThreads) : "memory"); };
    for (int c = (int)blockIdx.x + team * (int)gridDim.x; c < pv.C; c += kWinTeams * (int)gridDim.x) {
        team_sync();  // previous clone's pold, pnew, red consumed
        if (!a.alive[c]) continue;
        const int p = a.P_particle[c], bead0 = a.b0[c];
        for (int t = tid; t < (nl + 1) * 3; t += kWinTeamThreads) {
            const int jj = t / 3, d = t - jj * 3;
            int bg = bead0 + jj;
            bg = WrapSlice(pv, bg);
            const double x = a.R_moved[PosIndex(pv, a.N_moved, RealClone(pv, c), p, d, bg - pv.slice_lo)];
            pold[team][jj][d] = x;
            pnew[team][jj][d] = (jj >= 1 && jj < nl) ? a.P[((size_t)c * (nl - 1) + (jj - 1)) * 3 + d] : x;
        }
        team_sync();
        int b0s = bead0 + j, b1s = bead0 + j + 1;
        b0s = WrapSlice(pv, b0s);
        b1s = WrapSlice(pv, b1s);
        double acc_old = 0., acc_new = 0.;
        const uint32_t po_addr = (uint32_t)__cvta_generic_to_shared(&pold[team][j][0]);
        const uint32_t pn_addr = (uint32_t)__cvta_generic_to_shared(&pnew[team][j][0]);
        for (int g = warp; g < n_groups; g += kWinTeamWarps) {
            const int q = g * per_warp + sub;
            const bool on = q < a.N_partner && !(a.same && q == p);
            const int qc = q < a.N_partner ? q : a.N_partner - 1;
            double q0[3], q1[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                q0[d] = a.R_partner[PosIndex(pv, a.N_partner, RealClone(pv, c), qc, d, b0s - pv.slice_lo)];
                q1[d] = a.R_partner[PosIndex(pv, a.N_partner, RealClone(pv, c), qc, d, b1s - pv.slice_lo)];
            }
            double ro, rpo, so, rn, rpn, sn, m0[3], m1[3];
            LdsBeadPair(po_addr, m0, m1);
            DrDrpDrrpFast(m0, q0, m1, q1, pv.box, ro, rpo, so);
            LdsBeadPair(pn_addr, m0, m1);
            DrDrpDrrpFast(m0, q0, m1, q1, pv.box, rn, rpn, sn);
            double uo, un;
            FastIlkkaEvalWindow(wtab, a.FT, j, nl, lane, ro, rpo, so, rn, rpn, sn, uo, un);
            acc_old += on ? uo : 0.;
            acc_new += on ? un : 0.;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc_old += __shfl_down_sync(0xffffffffu, acc_old, o);
            acc_new += __shfl_down_sync(0xffffffffu, acc_new, o);
        }
        if (lane == 0) {
            red[team][0][warp] = acc_old;
            red[team][1][warp] = acc_new;
        }
        team_sync();
        if (tid == 0) {
            double to = 0., tn = 0.;
            for (int i = 0; i < kWinTeamWarps; ++i) {
                to += red[team][0][i];
                tn += red[team][1][i];
            }
            a.out_old[c] += to;
            a.out_new[c] += tn;
        }
    }
}

// ------------------------------------------------------------ long-range part of a window
constexpr int kMaxLrActions = 8;  // long-range actions that involve one species (inputs/C/c.xml: 4)
struct LrWindowArgs {
    PathView pv;
    SpeciesView sv;          // moved species with the pending proposal
    KSpaceView ks;
    const int32_t *b0;
    int n_window;
    const double2 *rho_self;  // committed rho_k of the moved species
    double2 *drho;            // [C][n_window][n_k] out
    int n_actions;
    const double2 *rho_other[kMaxLrActions];  // partner species' rho_k (nullptr: same species)
    const double *wk[kMaxLrActions];
    double factor[kMaxLrActions];             // scale * (2 if the species differ)
    double *lr_old, *lr_new;                  // [C], overwritten
    // fuse_decide != 0: the kernel also takes the level-0 Metropolis decision and commits
    // (bisect_decide_commit_kernel's work, one launch less per attempt)
    int fuse_decide;
    const int32_t *alive;    // nullptr: every clone is alive (DisplaceParticle has no earlier levels)
    const double *partial, *logu0, *pair_old, *pair_new;  // [C]; partial == nullptr: 0
    // pair sums left per 32-link chunk by displace_pair_kernel, [C][n_chunks][2] (OLD, NEW), summed here in chunk order
    const double *chunk_partial;
    int n_chunks;
    int N;
    double *R;               // committed positions of the moved species
    double2 *rho_commit;     // == rho_self, writable
    int32_t *accept;
    long long *n_accept;
    // recompute_commit != 0 (with fuse_decide): drho is NOT stored; an accepted proposal rebuilds the phase tables and adds
    // the same deltas (same arithmetic, same bits) to rho_k in the commit -- for a window that is the whole path
    // (DisplaceParticle) this takes the rho_k traffic of an attempt from five passes over the array (read rho, write drho;
    // read drho, read rho, write rho) to three
    int recompute_commit = 0;
};

/// Species::UpdateRhoK for the proposal (species_class.h:406-425) fused with CalcULong over the
/// window in OLD and NEW mode (ilkka_pair_action_class.h:104-122) for every long-range action
/// that involves the moved species.  One CTA per clone, thread per k vector.
constexpr int kLrChunk = 16;  // window slices whose phase tables are built together

static __global__ void __launch_bounds__(256) lr_window_kernel(const LrWindowArgs a) {
    extern __shared__ __align__(16) double2 ptab[];  // [kLrChunk][2 modes][3 axes][2m+1]
    __shared__ double red[2][256 / 32];
    __shared__ int s_accept;
    const PathView &pv = a.pv;
    const int c = blockIdx.x, tid = threadIdx.x;
    const int tl = 2 * a.ks.max_index + 1, n_k = a.ks.n_k;
    const int p = a.sv.P_particle[c];
    double acc_old = 0., acc_new = 0.;
    for (int j0 = 0; j0 < a.n_window; j0 += kLrChunk) {
        const int nj = min(kLrChunk, a.n_window - j0);
        __syncthreads();
        for (int t = tid; t < nj * 6; t += blockDim.x) {  // one thread per (slice, mode, axis)
            const int jj = t / 6, md = t - jj * 6;
            const int mode = md / 3, d = md - mode * 3;
            int bg = a.b0[c] + j0 + jj;
            bg = WrapSlice(pv, bg);
            double r[3];
            LoadPos(pv, a.sv, c, p, bg, mode, r);
            PhaseTable(r[d], a.ks.kbox, a.ks.max_index, ptab + (size_t)t * tl);
        }
        __syncthreads();
        for (int t = tid; t < nj * n_k; t += blockDim.x) {
            const int jj = t / n_k, k = t - jj * n_k;
            const int j = j0 + jj;
            int bg = a.b0[c] + j;
            bg = WrapSlice(pv, bg);
            const int i0 = a.ks.kidx[3 * k], i1 = a.ks.kidx[3 * k + 1], i2 = a.ks.kidx[3 * k + 2];
            const double2 *to = ptab + (size_t)jj * 6 * tl, *tn = to + 3 * tl;
            const double2 fo = CMul(CMul(to[i0], to[tl + i1]), to[2 * tl + i2]);
            const double2 fn = CMul(CMul(tn[i0], tn[tl + i1]), tn[2 * tl + i2]);
            const double2 d = make_double2(fn.x - fo.x, fn.y - fo.y);
            if (!a.recompute_commit) a.drho[((size_t)c * a.n_window + j) * n_k + k] = d;
            const size_t ri = ((size_t)RealClone(pv, c) * pv.Mloc + (bg - pv.slice_lo)) * n_k + k;
            const double2 rs = a.rho_self[ri];
            const double2 rn = make_double2(rs.x + d.x, rs.y + d.y);
            for (int t2 = 0; t2 < a.n_actions; ++t2) {
                const double w = a.wk[t2][k] * a.factor[t2];
                if (a.rho_other[t2]) {
                    const double2 ro = a.rho_other[t2][ri];
                    acc_old += w * (rs.x * ro.x + rs.y * ro.y);
                    acc_new += w * (rn.x * ro.x + rn.y * ro.y);
                } else {
                    acc_old += w * (rs.x * rs.x + rs.y * rs.y);
                    acc_new += w * (rn.x * rn.x + rn.y * rn.y);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_old += __shfl_down_sync(0xffffffffu, acc_old, o);
        acc_new += __shfl_down_sync(0xffffffffu, acc_new, o);
    }
    if ((tid & 31) == 0) {
        red[0][tid >> 5] = acc_old;
        red[1][tid >> 5] = acc_new;
    }
    __syncthreads();
    if (tid == 0) {
        double to = 0., tn = 0.;
        for (int i = 0; i < 256 / 32; ++i) {
            to += red[0][i];
            tn += red[1][i];
        }
        a.lr_old[c] = to;
        a.lr_new[c] = tn;
        if (a.fuse_decide) {  // bisect_class.h:110-115 / displace_particle_class.h:65-71, same arithmetic as the *_decide_commit kernels
            int acc = 0;
            if (!a.alive || a.alive[c]) {
                double po, pn;
                if (a.chunk_partial) {
                    po = pn = 0.;
                    for (int i = 0; i < a.n_chunks; ++i) {
                        po += a.chunk_partial[((size_t)c * a.n_chunks + i) * 2];
                        pn += a.chunk_partial[((size_t)c * a.n_chunks + i) * 2 + 1];
                    }
                } else {
                    po = a.pair_old[c];
                    pn = a.pair_new[c];
                }
                const double old_action = po + to, new_action = pn + tn;
                acc = a.partial ? ((a.partial[c] - (new_action - old_action)) < a.logu0[c] ? 0 : 1)
                                : ((old_action - new_action) < a.logu0[c] ? 0 : 1);
            }
            a.accept[c] = acc;
            a.n_accept[c] += acc;
            s_accept = acc;
        }
    }
    if (!a.fuse_decide) return;
    __syncthreads();
    if (!s_accept) return;
    // Move::Accept: the proposal's beads (n_prop beads from P_first: the window's interior for a bisection, the whole
    // path for a displacement) and rho_k += delta on the window (every read of rho_self is behind the barrier)
    const int n_prop = a.sv.n_prop, bead0 = a.b0[c], first = a.sv.P_first[c];
    if (a.recompute_commit) {
        // rho_k first, while the committed positions are still the OLD ones: the same tables, the same deltas as above
        for (int j0 = 0; j0 < a.n_window; j0 += kLrChunk) {
            const int nj = min(kLrChunk, a.n_window - j0);
            __syncthreads();
            for (int t = tid; t < nj * 6; t += blockDim.x) {
                const int jj = t / 6, md = t - jj * 6;
                const int mode = md / 3, d = md - mode * 3;
                int bg = bead0 + j0 + jj;
                bg = WrapSlice(pv, bg);
                double r[3];
                LoadPos(pv, a.sv, c, p, bg, mode, r);
                PhaseTable(r[d], a.ks.kbox, a.ks.max_index, ptab + (size_t)t * tl);
            }
            __syncthreads();
            for (int t = tid; t < nj * n_k; t += blockDim.x) {
                const int jj = t / n_k, k = t - jj * n_k;
                int bg = bead0 + j0 + jj;
                bg = WrapSlice(pv, bg);
                const int i0 = a.ks.kidx[3 * k], i1 = a.ks.kidx[3 * k + 1], i2 = a.ks.kidx[3 * k + 2];
                const double2 *to = ptab + (size_t)jj * 6 * tl, *tn = to + 3 * tl;
                const double2 fo = CMul(CMul(to[i0], to[tl + i1]), to[2 * tl + i2]);
                const double2 fn = CMul(CMul(tn[i0], tn[tl + i1]), tn[2 * tl + i2]);
                double2 *dst = a.rho_commit + ((size_t)RealClone(pv, c) * pv.Mloc + (bg - pv.slice_lo)) * n_k + k;
                dst->x += fn.x - fo.x;
                dst->y += fn.y - fo.y;
            }
        }
        __syncthreads();  // every OLD position has been read
    }
    for (int t = tid; t < n_prop * 3; t += blockDim.x) {
        const int j = t / 3, d = t - j * 3;
        int bg = first + j;
        bg = WrapSlice(pv, bg);
        a.R[PosIndex(pv, a.N, RealClone(pv, c), p, d, bg - pv.slice_lo)] = a.sv.P[((size_t)c * n_prop + j) * 3 + d];
    }
    if (a.recompute_commit) return;
    for (int t = tid; t < a.n_window * n_k; t += blockDim.x) {
        const int j = t / n_k, k = t - j * n_k;
        int bg = bead0 + j;
        bg = WrapSlice(pv, bg);
        double2 *dst = a.rho_commit + ((size_t)RealClone(pv, c) * pv.Mloc + (bg - pv.slice_lo)) * n_k + k;
        const double2 d = a.drho[((size_t)c * a.n_window + j) * n_k + k];
        dst->x += d.x;
        dst->y += d.y;
    }
}

/// bisect_decide_kernel + Move::Accept in one launch: one CTA per clone.
static __global__ void __launch_bounds__(256) bisect_decide_commit_kernel(PathView pv, int N, int n_k, int n_window, const int32_t *__restrict__ alive,
                                                                   const double *__restrict__ partial, const double *__restrict__ logu0,
                                                                   const double *__restrict__ pair_old, const double *__restrict__ pair_new,
                                                                   const double *__restrict__ lr_old, const double *__restrict__ lr_new,
                                                                   const double *__restrict__ P, const int32_t *__restrict__ P_particle,
                                                                   const int32_t *__restrict__ b0, const double2 *__restrict__ drho,
                                                                   double *__restrict__ R, double2 *__restrict__ rho,
                                                                   int32_t *__restrict__ accept, long long *__restrict__ n_accept) {
    const int c = blockIdx.x;
    int acc = 0;
    if (alive[c]) {
        const double old_action = pair_old[c] + lr_old[c], new_action = pair_new[c] + lr_new[c];
        acc = (partial[c] - (new_action - old_action)) < logu0[c] ? 0 : 1;
    }
    if (threadIdx.x == 0) {
        accept[c] = acc;
        n_accept[c] += acc;
    }
    if (!acc) return;
    const int p = P_particle[c], bead0 = b0[c], n_prop = n_window - 1;
    for (int t = threadIdx.x; t < n_prop * 3; t += blockDim.x) {
        const int j = t / 3, d = t - j * 3;
        int bg = bead0 + 1 + j;
        bg = WrapSlice(pv, bg);
        R[PosIndex(pv, N, RealClone(pv, c), p, d, bg - pv.slice_lo)] = P[((size_t)c * n_prop + j) * 3 + d];
    }
    if (drho) {
        for (int t = threadIdx.x; t < n_window * n_k; t += blockDim.x) {
            const int j = t / n_k, k = t - j * n_k;
            int bg = bead0 + j;
            bg = WrapSlice(pv, bg);
            double2 *dst = rho + ((size_t)RealClone(pv, c) * pv.Mloc + (bg - pv.slice_lo)) * n_k + k;
            const double2 d = drho[((size_t)c * n_window + j) * n_k + k];
            dst->x += d.x;
            dst->y += d.y;
        }
    }
}

}  // namespace pimc

#endif  // SIMPIMC_B200_MC_CUH_
