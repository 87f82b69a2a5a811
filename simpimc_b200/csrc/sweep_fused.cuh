// One launch = n_attempts x Bisect::DoEvent on every clone
// (src/events/moves/move_class.h:61-77 -> single_species_move/bisect/bisect_class.h:39-139)
// for the common case of ONE same-species Ilkka pair action (with or without its long-range
// part) acting on the moved species -- the uniform electron gas of BASELINE.json.
//
// Clones are independent walkers, so a sweep needs no grid-wide synchronisation: persistent
// CTAs (one per SM, 1024 threads) stage the fast Ilkka tables in shared memory ONCE and then
// split into kSweepTeams independent TEAMS (8 teams of 4 warps, one clone each, measured best:
// 2 teams 2863, 4 teams 3037, 8 teams 3194 clone-sweeps/s at C3).  A team owns up to kTeamClones
// clones at a time and runs their attempts back to back, synchronising only with itself (named
// barrier) between the phases of an attempt -- the teams drift apart, so the serial phases of
// one (Levy construction, Metropolis test: one thread per clone) are hidden behind the pair
// work of the others while all of them share one copy of the tables:
//
//   A   Philox draws for every (clone, bead) in parallel: particle and window, the Levy
//       displacements sigma * normal of every midpoint, the Metropolis uniforms; the window's
//       committed beads are loaded
//   A'  one thread per clone: Levy construction level by level (bisect_class.h:69-98), kinetic
//       action in closed form, Metropolis tests of the levels above 0
//   B   PairAction::GetAction in OLD and NEW mode (pair_action_class.h:267-302): lanes = (partner
//       particle, link); the partner's beads are read once for both modes; per-axis phase
//       tables of the window's old and new beads are built on the side
//   C   Species::UpdateRhoK for the proposal (species_class.h:406-425) fused with CalcULong over
//       the window in OLD and NEW mode (ilkka_pair_action_class.h:104-122)
//   D   level-0 Metropolis test (bisect_class.h:110-115)
//   E   Move::Accept: positions and rho_k += (new - old) (bisect_class.h:24-36)
//
// The multi-kernel path in mc.cuh (any action mix, any species pair) draws the same Philox
// stream and remains the general implementation; both are tested against the same host mirror.
#ifndef SIMPIMC_B200_SWEEP_FUSED_CUH_
#define SIMPIMC_B200_SWEEP_FUSED_CUH_

#include "mc.cuh"

namespace pimc {

// A/B (tools/gpu_ab_sweep.sh): 512-thread CTAs (four teams, 123 registers, no spills) 2530 clone-sweeps/s against 3300 with
// 1024 threads x 64 registers: like K1, this kernel wants the warps more than the registers
#ifndef PIMC_SWEEP_THREADS
#define PIMC_SWEEP_THREADS 1024
#endif
constexpr int kSweepThreads = PIMC_SWEEP_THREADS;
#ifndef PIMC_SWEEP_TEAMS
#define PIMC_SWEEP_TEAMS 8
#endif
constexpr int kSweepTeams = PIMC_SWEEP_TEAMS;                // independent sub-CTA teams
constexpr int kTeamThreads = kSweepThreads / kSweepTeams;
constexpr int kTeamWarps = kTeamThreads / 32;
constexpr int kTeamClones = (kSweepThreads / 128) / kSweepTeams;  // clones a team advances together (one per 4 warps)
constexpr int kSweepClones = kSweepTeams * kTeamClones;      // clones in flight per CTA
constexpr int kSweepGroup = kTeamThreads / kTeamClones;      // threads that own one clone in phases C-E
constexpr int kSweepGroupWarps = kSweepGroup / 32;
// Where the windows' partner beads are gathered from (A/B, tools/gpu_ab_sweep2.sh): 0 = the whole-path layout R
// ([particle][dim][slice]: 72-byte fragments of 128-byte lines), 1 = slice-major mirror [slice][particle][dim],
// 2 = slice-major mirror [slice][dim][particle]
// Measured at C3 (1024 clones, 256 attempts, ms per attempt): 0 -> 0.0750, 1 -> 0.0815, 2 -> 0.0860.  The mirrors cut the
// DRAM traffic but every load instruction then touches 8 lines (one per slice) instead of 4 (one per partner):
// the gather is bound by L1 request processing and latency, not by DRAM bytes, so the whole-path layout stays.
#ifndef PIMC_SWEEP_MIRROR
#define PIMC_SWEEP_MIRROR 0
#endif
constexpr int kSweepMaxBeads = 16;                           // 2^n_level <= 16
constexpr int kSweepMaxLevel = 4;

struct SweepFusedArgs {
    PathView pv;
    double *R;   // committed positions [C][N][3][Ms] (the layout of the whole-path kernels): written on acceptance
    double *R2;  // slice-major mirror [C][Mstore][N][3]: what the windows are gathered from; written on acceptance too
    int N;
    double lambda, tau;
    int n_level;
    int with_kinetic;
    FreeSplineSet fs_move, fs_kin;  // as in BisectArgs (mc.cuh)
    int b0_lo, b0_count;  // window starts: uniform in [b0_lo, b0_lo + b0_count)
    uint32_t seed_lo, seed_hi;
    unsigned long long attempt0;
    int n_attempts;
    FastTable FT;
    const unsigned char *fast_tables;
    int use_lr;
    KSpaceView ks;
    double2 *rho;      // committed rho_k of the species (read and written)
    double2 *rho_new;  // scratch [C][2^n_level][n_k]: rho_k + delta of the window, written in phase C, committed in phase E
    const double *wk;  // [n_k]
    double lr_factor;
    long long *n_accept;  // [C], added to
};

/// Shared state of the clones a CTA is advancing.
struct SweepShared {
    double pold[kSweepClones][kSweepMaxBeads + 1][3];
    double pnew[kSweepClones][kSweepMaxBeads + 1][3];
    double del_new[kSweepClones][kSweepMaxBeads][3];  // sigma * normal, folded into the box
    double d2_new[kSweepClones][kSweepMaxBeads];
    double logu[kSweepClones][kSweepMaxLevel];
    double partial[kSweepClones];
    double wsum[kSweepClones][2][kTeamWarps];
    double lrsum[kSweepClones][2][kSweepGroupWarps];
    int particle[kSweepClones];
    int bead0[kSweepClones];
    int alive[kSweepClones];
    int accept[kSweepClones];
};

/// Barrier among the kTeamThreads threads of one team (barrier 0 is __syncthreads).
__device__ __forceinline__ void TeamSync(int team) { asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(kTeamThreads) : "memory"); }

__device__ __forceinline__ void PrefetchL2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

/// Offset of (particle q, dim d) inside one slice block of the mirror.
__device__ __forceinline__ unsigned MirrorRow(int N, int q, int d) {
#if PIMC_SWEEP_MIRROR == 2
    return (unsigned)(d * N + q);
#else
    (void)N;
    return (unsigned)(q * 3 + d);
#endif
}

/// R[c][row = particle * 3 + dim][slice (row length Ms)] -> R2[c][slice][row]: 32 x 32 tiles through shared
/// memory, both sides coalesced.  grid (slice tiles, row tiles, clones), block (32, 8).
static __global__ void __launch_bounds__(256) slice_major_kernel(const double *__restrict__ R, int n_rows, int Mstore, int Ms, double *__restrict__ R2) {
    __shared__ double tile[32][33];
    const int c = blockIdx.z, s0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const double *src = R + (size_t)c * n_rows * Ms;
    double *dst = R2 + (size_t)c * Mstore * n_rows;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int row = r0 + i, sl = s0 + threadIdx.x;
        if (row < n_rows && sl < Mstore) tile[i][threadIdx.x] = src[(size_t)row * Ms + sl];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int sl = s0 + i, row = r0 + threadIdx.x;
        if (row < n_rows && sl < Mstore) dst[(size_t)sl * n_rows + MirrorRow(n_rows / 3, row / 3, row % 3)] = tile[threadIdx.x][i];
    }
}

/// IMAGES: the FreeSpline branches (periodic images of the free-particle density matrix) are compiled in; the
/// n_images = 0 instantiation keeps the closed forms only (0.0750 vs 0.0790 ms per attempt at C3).
template <bool IMAGES>
static __global__ void __launch_bounds__(kSweepThreads, 1) bisect_sweep_fused_kernel(const SweepFusedArgs a) {
    extern __shared__ __align__(16) unsigned char ssm[];
    __shared__ SweepShared sh;
    const int lane = threadIdx.x & 31;
    const int team = threadIdx.x / kTeamThreads, tid = threadIdx.x - team * kTeamThreads;  // tid: inside the team
    const int warp = tid >> 5;                                                              // warp inside the team
    const PathView &pv = a.pv;
    const int nb = 1 << a.n_level;
    const int tl = 2 * a.ks.max_index + 1, n_k = a.use_lr ? a.ks.n_k : 0;
    // dynamic shared memory: [fast tables][phase tables: clone slot, window slice, mode, axis, 2m+1]
    __shared__ unsigned long long stage_bar;
    StageBlockTma(ssm, a.fast_tables, a.FT.n_bytes, &stage_bar);  // the only CTA-wide synchronisation: from here on the teams run on their own
    const SharedTab tb(ssm);
    const int s0 = team * kTeamClones;  // first shared-memory clone slot of this team
    double2 *ptab = reinterpret_cast<double2 *>(ssm + a.FT.n_bytes) + (size_t)s0 * nb * 6 * tl;
    const int grp = tid / kSweepGroup, tg = tid - grp * kSweepGroup;
    // clones of a CTA: blockIdx.x + i * gridDim.x, i < per_cta (every CTA holds the same count +-1);
    // team t takes i = t, t + kSweepTeams, ... , kTeamClones of them at a time
    const int per_cta = (pv.C + gridDim.x - 1) / gridDim.x;
    const int per_team = (per_cta - team + kSweepTeams - 1) / kSweepTeams;
#define SWEEP_CLONE(lc) ((int)blockIdx.x + (team + (batch + (lc)) * kSweepTeams) * (int)gridDim.x)
    for (int batch = 0; batch < per_team; batch += kTeamClones) {
        int nlc = 0;
        for (int lc = 0; lc < kTeamClones; ++lc)
            if (batch + lc < per_team && SWEEP_CLONE(lc) < pv.C) nlc = lc + 1;
        if (nlc == 0) break;
        const int c_grp = SWEEP_CLONE(grp);  // clone of this thread's group (phases C-E)
        const int sg = s0 + grp;             // its shared-memory slot
        long long my_accepts = 0;
        for (int it = 0; it < a.n_attempts; ++it) {
            const unsigned long long attempt = a.attempt0 + (unsigned long long)it;
            const uint32_t at_lo = (uint32_t)attempt, at_hi = (uint32_t)(attempt >> 32);
            // ---------------------------------------------------------------- phase A
            if (tid < nlc * (nb + 1) * 3) {  // committed beads of the window
                const int lc = tid / ((nb + 1) * 3), rem = tid - lc * (nb + 1) * 3;
                const int j = rem / 3, d = rem - j * 3;
                const int c = SWEEP_CLONE(lc);
                uint32_t rnd[4];
                Philox4x32(at_lo, at_hi, (uint32_t)c, 0u, a.seed_lo, a.seed_hi, rnd);
                int p_i = (int)(UniformFromBits(rnd[0], rnd[1]) * a.N);
                p_i = p_i < a.N ? p_i : a.N - 1;
                int b0 = (int)(UniformFromBits(rnd[2], rnd[3]) * a.b0_count);
                b0 = a.b0_lo + (b0 < a.b0_count ? b0 : a.b0_count - 1);
                int bg = b0 + j;
                bg = WrapSlice(pv, bg);
#if PIMC_SWEEP_MIRROR
                const double x = a.R2[((size_t)c * pv.Mstore + (bg - pv.slice_lo)) * a.N * 3 + MirrorRow(a.N, p_i, d)];
#else
                const double x = a.R[PosIndex(pv, a.N, c, p_i, d, bg - pv.slice_lo)];
#endif
                sh.pold[s0 + lc][j][d] = x;
                sh.pnew[s0 + lc][j][d] = x;
                if (rem == 0) {
                    sh.particle[s0 + lc] = p_i;
                    sh.bead0[s0 + lc] = b0;
                }
            } else if (tid >= kTeamThreads / 2 && tid < kTeamThreads / 2 + nlc * (nb - 1)) {  // Levy displacements
                const int t = tid - kTeamThreads / 2;
                const int lc = t / (nb - 1), ib = t - lc * (nb - 1) + 1;
                const int c = SWEEP_CLONE(lc);
                const int level = __ffs(ib) - 1, skip = 1 << level;
                const int idx = (ib - skip) >> (level + 1);
                const uint32_t slot = SweepSlotStart(level, a.n_level, nb) + 2u * (uint32_t)idx;
                uint32_t r0[4], r1[4];
                Philox4x32(at_lo, at_hi, (uint32_t)c, slot, a.seed_lo, a.seed_hi, r0);
                Philox4x32(at_lo, at_hi, (uint32_t)c, slot + 1, a.seed_lo, a.seed_hi, r1);
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
                    sh.del_new[s0 + lc][ib][d] = del;
                    delv[d] = del;
                    d2 += del * del;
                }
                sh.d2_new[s0 + lc][ib] = (IMAGES && a.fs_move.n_images) ? -FreeLogRho(a.fs_move.s[level], delv) : d2;
            } else if (tid >= 3 * kTeamThreads / 4 && tid < 3 * kTeamThreads / 4 + nlc * a.n_level) {  // Metropolis uniforms
                const int t = tid - 3 * kTeamThreads / 4;
                const int lc = t / a.n_level, level = t - lc * a.n_level;
                const int c = SWEEP_CLONE(lc);
                const uint32_t slot = SweepSlotStart(level, a.n_level, nb) + 2u * (uint32_t)(nb >> (level + 1));
                uint32_t ru[4];
                Philox4x32(at_lo, at_hi, (uint32_t)c, slot, a.seed_lo, a.seed_hi, ru);
                sh.logu[s0 + lc][level] = log(UniformFromBits(ru[0], ru[1]));
            }
            TeamSync(team);
            // ---------------------------------------------------------------- phase A'
            if (tg == 0 && grp < nlc) {
                double(*oldb)[3] = sh.pold[sg];
                double(*newb)[3] = sh.pnew[sg];
                bool alive = true;
                double prev_change = 0., partial = 0.;
                for (int level = a.n_level - 1; level >= 0; --level) {
                    const int skip = 1 << level;
                    const double level_tau = a.tau * skip;
                    const double i4lt_sample = 1. / (4. * a.lambda * (0.5 * level_tau));
                    const double i4lt_kin = 1. / (4. * a.lambda * level_tau);
                    double old_lp = 0., new_lp = 0.;
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
                            newb[ib][d] = rbar_new + sh.del_new[sg][ib][d];
                        }
                        if (IMAGES && a.fs_move.n_images) {
                            old_lp += FreeLogRho(a.fs_move.s[level], delo);
                            new_lp -= sh.d2_new[sg][ib];
                        } else {
                            old_lp -= d2_old * i4lt_sample;
                            new_lp -= sh.d2_new[sg][ib] * i4lt_sample;
                        }
                    }
                    double old_kin = 0., new_kin = 0.;
                    if (a.with_kinetic) {
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
                            if (IMAGES && a.fs_kin.n_images) {
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
                        if (lsr - change + prev_change < sh.logu[sg][level]) alive = false;
                        prev_change = change;
                    } else {
                        partial = lsr - change + prev_change;
                    }
                }
                sh.partial[sg] = partial;
                sh.alive[sg] = alive ? 1 : 0;
            } else if (grp < nlc) {
                // the other threads of the group pull what phases B and C will read towards L2 while
                // the Levy construction runs: the window's slices of every particle row (first and
                // last byte: a 72-byte window touches one or two 128-byte lines) and of rho_k
                const int bead0 = sh.bead0[sg];
                int b_last = bead0 + nb;
                b_last = WrapSlice(pv, b_last);
#if PIMC_SWEEP_MIRROR
                (void)b_last;
                // the window's nb + 1 slices of the slice-major mirror: contiguous blocks of N * 24 bytes
                const int slice_bytes = a.N * 24, lines = (slice_bytes + 127) / 128;
                for (int t = tg - 1; t < (nb + 1) * lines; t += kSweepGroup - 1) {
                    const int j = t / lines, l = t - j * lines;
                    int bg = bead0 + j;
                    bg = WrapSlice(pv, bg);
                    const char *p = reinterpret_cast<const char *>(a.R2 + ((size_t)c_grp * pv.Mstore + (bg - pv.slice_lo)) * a.N * 3);
                    PrefetchL2(p + min(l * 128, slice_bytes - 8));
                }
#else
                const double *Rc = a.R + PosIndex(pv, a.N, c_grp, 0, 0, 0);
                const int n_rows = a.N * 3;
                for (int row = tg - 1; row < n_rows; row += kSweepGroup - 1) {
                    const double *first = Rc + (size_t)row * pv.Ms + bead0 - pv.slice_lo;
                    const double *last = Rc + (size_t)row * pv.Ms + b_last - pv.slice_lo;
                    PrefetchL2(first);
                    if (((uintptr_t)first >> 7) != ((uintptr_t)last >> 7)) PrefetchL2(last);  // second 128-byte line, or the wrapped end
                }
#endif
                if (n_k > 0) {
                    const int lines = (n_k * (int)sizeof(double2) + 127) / 128;
                    for (int t = tg - 1; t < nb * lines; t += kSweepGroup - 1) {
                        const int j = t / lines, l = t - j * lines;
                        int bg = bead0 + j;
                        bg = WrapSlice(pv, bg);
                        const char *p = reinterpret_cast<const char *>(a.rho + ((size_t)c_grp * pv.Mloc + (bg - pv.slice_lo)) * n_k);
                        PrefetchL2(p + min(l * 128, n_k * (int)sizeof(double2) - 1));
                    }
                }
            }
            TeamSync(team);
            // ---------------------------------------------------------------- phase B
            if (n_k > 0 && tid < nlc * nb * 6) {  // phase tables of the window's old and new beads
                const int lc = tid / (nb * 6), rem = tid - lc * nb * 6;
                const int j = rem / 6, md = rem - j * 6;
                const int mode = md / 3, d = md - mode * 3;
                if (sh.alive[s0 + lc])
                    PhaseTable(mode ? sh.pnew[s0 + lc][j][d] : sh.pold[s0 + lc][j][d], a.ks.kbox, a.ks.max_index, ptab + (size_t)tid * tl);
            }
            {
                const int per_warp = 32 / nb;  // partner particles per warp item
                const int n_groups = (a.N + per_warp - 1) / per_warp;
                const int j = lane & (nb - 1), sub = lane / nb;
                for (int lc = 0; lc < nlc; ++lc) {
                    if (!sh.alive[s0 + lc]) continue;  // rejected above level 0
                    const int c = SWEEP_CLONE(lc);
                    const int p = sh.particle[s0 + lc];
                    int b0s = sh.bead0[s0 + lc] + j, b1s = b0s + 1;
                    b0s = WrapSlice(pv, b0s);
                    b1s = WrapSlice(pv, b1s);
                    // moved-particle beads are re-read from shared memory for every evaluation (asm
                    // volatile: not hoisted) -- holding OLD and NEW copies in registers spills at 64
                    const uint32_t po_addr = (uint32_t)__cvta_generic_to_shared(&sh.pold[s0 + lc][j][0]);
                    const uint32_t pn_addr = (uint32_t)__cvta_generic_to_shared(&sh.pnew[s0 + lc][j][0]);
#if PIMC_SWEEP_MIRROR
                    // 32-bit offsets inside the clone's block of the slice-major mirror
                    const double *Rc = a.R2 + (size_t)c * pv.Mstore * a.N * 3;
                    const unsigned o0 = (unsigned)(b0s - pv.slice_lo) * (unsigned)a.N * 3u, o1 = (unsigned)(b1s - pv.slice_lo) * (unsigned)a.N * 3u;
#else
                    const double *Rc = a.R + PosIndex(pv, a.N, c, 0, 0, 0);
                    const unsigned row_stride = 3u * (unsigned)pv.Ms, ms = (unsigned)pv.Ms;
                    const unsigned o0 = (unsigned)(b0s - pv.slice_lo), o1 = (unsigned)(b1s - pv.slice_lo);
#endif
                    double acc_old = 0., acc_new = 0.;
                    for (int g = warp; g < n_groups; g += kTeamWarps) {
                        const int q = g * per_warp + sub;
                        const bool on = q < a.N && q != p;
                        double q0[3], q1[3];
#if PIMC_SWEEP_MIRROR
                        const int qq = q < a.N ? q : a.N - 1;
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            q0[d] = Rc[o0 + MirrorRow(a.N, qq, d)];
                            q1[d] = Rc[o1 + MirrorRow(a.N, qq, d)];
                        }
#else
                        const unsigned row = (unsigned)(q < a.N ? q : a.N - 1) * row_stride;
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            q0[d] = Rc[row + d * ms + o0];
                            q1[d] = Rc[row + d * ms + o1];
                        }
#endif
                        // both sets of distances first: the partner's beads die before the table work
                        double ro, rpo, so, rn, rpn, sn, m0[3], m1[3];
                        LdsBeadPair(po_addr, m0, m1);
                        DrDrpDrrpFast(m0, q0, m1, q1, pv.box, ro, rpo, so);
                        LdsBeadPair(pn_addr, m0, m1);
                        DrDrpDrrpFast(m0, q0, m1, q1, pv.box, rn, rpn, sn);
                        double uo, un;
                        FastIlkkaEvalWindow(tb, a.FT, j, nb, lane, ro, rpo, so, rn, rpn, sn, uo, un);
                        acc_old += on ? uo : 0.;
                        acc_new += on ? un : 0.;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        acc_old += __shfl_down_sync(0xffffffffu, acc_old, o);
                        acc_new += __shfl_down_sync(0xffffffffu, acc_new, o);
                    }
                    if (lane == 0) {
                        sh.wsum[s0 + lc][0][warp] = acc_old;
                        sh.wsum[s0 + lc][1][warp] = acc_new;
                    }
                }
            }
            TeamSync(team);
            // ---------------------------------------------------------------- phase C
            const bool grp_on = grp < nlc && sh.alive[grp < nlc ? sg : s0];
            if (n_k > 0) {
                double acc_old = 0., acc_new = 0.;
                if (grp_on) {
                    // thread = k vector (its indices and weight are read once), inner loop = the
                    // window's slices, unrolled so that the rho_k loads of several slices are in flight
                    const int bead0 = sh.bead0[sg];
                    const double2 *pt = ptab + (size_t)grp * nb * 6 * tl;
                    const double2 *rho_c = a.rho + (size_t)c_grp * pv.Mloc * n_k;
                    double2 *rn_c = a.rho_new + (size_t)c_grp * nb * n_k;
                    for (int k = tg; k < n_k; k += kSweepGroup) {
                        const int i0 = a.ks.kidx[3 * k], i1 = tl + a.ks.kidx[3 * k + 1], i2 = 2 * tl + a.ks.kidx[3 * k + 2];
                        const double w = a.wk[k] * a.lr_factor;
#pragma unroll 4
                        for (int j = 0; j < nb; ++j) {
                            int bg = bead0 + j;
                            bg = WrapSlice(pv, bg);
                            const double2 rs = rho_c[(size_t)(bg - pv.slice_lo) * n_k + k];
                            const double2 *to = pt + (size_t)j * 6 * tl, *tn = to + 3 * tl;
                            const double2 fo = CMul(CMul(to[i0], to[i1]), to[i2]);
                            const double2 fn = CMul(CMul(tn[i0], tn[i1]), tn[i2]);
                            const double2 rn = make_double2(rs.x + (fn.x - fo.x), rs.y + (fn.y - fo.y));
                            rn_c[(size_t)j * n_k + k] = rn;  // phase E copies it (same thread: program order)
                            acc_old += w * (rs.x * rs.x + rs.y * rs.y);
                            acc_new += w * (rn.x * rn.x + rn.y * rn.y);
                        }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    acc_old += __shfl_down_sync(0xffffffffu, acc_old, o);
                    acc_new += __shfl_down_sync(0xffffffffu, acc_new, o);
                }
                if (lane == 0) {
                    sh.lrsum[sg][0][warp - grp * kSweepGroupWarps] = acc_old;
                    sh.lrsum[sg][1][warp - grp * kSweepGroupWarps] = acc_new;
                }
                TeamSync(team);
            }
            // ---------------------------------------------------------------- phase D
            if (tg == 0 && grp < nlc) {
                int acc = 0;
                if (sh.alive[sg]) {
                    double po = 0., pn = 0., lo = 0., ln = 0.;
                    for (int w = 0; w < kTeamWarps; ++w) {
                        po += sh.wsum[sg][0][w];
                        pn += sh.wsum[sg][1][w];
                    }
                    if (n_k > 0)
                        for (int w = 0; w < kSweepGroupWarps; ++w) {
                            lo += sh.lrsum[sg][0][w];
                            ln += sh.lrsum[sg][1][w];
                        }
                    const double old_action = po + lo, new_action = pn + ln;
                    acc = (sh.partial[sg] - (new_action - old_action)) < sh.logu[sg][0] ? 0 : 1;
                }
                sh.accept[sg] = acc;
                my_accepts += acc;
            }
            TeamSync(team);
            // ---------------------------------------------------------------- phase E
            if (grp < nlc && sh.accept[sg]) {
                const int p = sh.particle[sg], bead0 = sh.bead0[sg];
                for (int t = tg; t < (nb - 1) * 3; t += kSweepGroup) {
                    const int j = t / 3 + 1, d = t - (j - 1) * 3;
                    int bg = bead0 + j;
                    bg = WrapSlice(pv, bg);
                    a.R[PosIndex(pv, a.N, c_grp, p, d, bg - pv.slice_lo)] = sh.pnew[sg][j][d];
#if PIMC_SWEEP_MIRROR
                    a.R2[((size_t)c_grp * pv.Mstore + (bg - pv.slice_lo)) * a.N * 3 + MirrorRow(a.N, p, d)] = sh.pnew[sg][j][d];
#endif
                }
                if (n_k > 0) {
                    // slice 0 of the window keeps its bead; the others take rho_k + delta as phase C left it
                    double2 *rho_c = a.rho + (size_t)c_grp * pv.Mloc * n_k;
                    const double2 *rn_c = a.rho_new + (size_t)c_grp * nb * n_k;
                    for (int k = tg; k < n_k; k += kSweepGroup) {
#pragma unroll 4
                        for (int j = 1; j < nb; ++j) {
                            int bg = bead0 + j;
                            bg = WrapSlice(pv, bg);
                            rho_c[(size_t)(bg - pv.slice_lo) * n_k + k] = rn_c[(size_t)j * n_k + k];
                        }
                    }
                }
            }
            TeamSync(team);
        }
        if (tg == 0 && grp < nlc) a.n_accept[c_grp] += my_accepts;
    }
#undef SWEEP_CLONE
}

}  // namespace pimc

#endif  // SIMPIMC_B200_SWEEP_FUSED_CUH_
