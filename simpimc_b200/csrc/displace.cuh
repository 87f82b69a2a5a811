// Device-resident DisplaceParticle::DoEvent
// (src/events/moves/single_species_move/displace_particle_class.h:13-89) on every clone at once:
// pick a particle, shift its WHOLE path rigidly by a vector of fixed length step_size whose
// direction is a normalised uniform point of the cube (scaffold/rng/rng.h:38-56 -- not isotropic,
// App. A-16), evaluate every pair action touching the species over all slices in OLD and NEW mode
// (GetAction(0, n_bead, particles, 0), displace_particle_class.h:54-63), Metropolis, commit.
// A rigid shift leaves the free-particle (Kinetic, n_images = 0) action unchanged.
//
//   displace_sample_kernel   Philox draws, proposal P = committed beads + dr (the proposal overlay
//                            that rhok / long-range kernels already understand)
//   displace_pair_kernel     persistent CTAs; item = (clone, 32 links); lanes = links, warps walk
//                            the partner particles; OLD and NEW from the same partner loads
//   lr_window_kernel         (mc.cuh) with the window = the whole path: drho and the k-sums
//   displace_decide_commit_kernel
//
// Philox slots of an attempt (counter = attempt, clone, slot): 0 particle, 1-2 direction, 3
// Metropolis uniform.  simpimc_b200/moves.py (displace_attempt_philox) mirrors the stream.
#ifndef SIMPIMC_B200_DISPLACE_CUH_
#define SIMPIMC_B200_DISPLACE_CUH_

#include "mc.cuh"

namespace pimc {

struct DisplaceSampleArgs {
    PathView pv;
    const double *R;
    int N;
    double step;
    uint32_t seed_lo, seed_hi, attempt_lo, attempt_hi;
    double *P;             // [C][M][3]
    int32_t *P_particle;   // [C]
    int32_t *P_first;      // [C] = 0
    int32_t *b0;           // [C] = 0
    double *dr;            // [C][3]
    double *logu;          // [C]
};

/// One CTA per clone.
static __global__ void __launch_bounds__(128) displace_sample_kernel(const DisplaceSampleArgs a) {
    __shared__ double sdr[3];
    __shared__ int sp;
    const PathView &pv = a.pv;
    const int c = blockIdx.x;
    if (threadIdx.x == 0) {
        uint32_t r0[4], r1[4], r2[4], r3[4];
        Philox4x32(a.attempt_lo, a.attempt_hi, (uint32_t)c, 0u, a.seed_lo, a.seed_hi, r0);
        Philox4x32(a.attempt_lo, a.attempt_hi, (uint32_t)c, 1u, a.seed_lo, a.seed_hi, r1);
        Philox4x32(a.attempt_lo, a.attempt_hi, (uint32_t)c, 2u, a.seed_lo, a.seed_hi, r2);
        Philox4x32(a.attempt_lo, a.attempt_hi, (uint32_t)c, 3u, a.seed_lo, a.seed_hi, r3);
        int p_i = (int)(UniformFromBits(r0[0], r0[1]) * a.N);
        p_i = p_i < a.N ? p_i : a.N - 1;
        // UnifRand(-1, 1) = (b - a) * u + a per component, r /= mag(r), r *= l
        double v[3] = {2. * UniformFromBits(r1[0], r1[1]) + -1., 2. * UniformFromBits(r1[2], r1[3]) + -1.,
                       2. * UniformFromBits(r2[0], r2[1]) + -1.};
        const double m = Mag3(v[0], v[1], v[2]);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            v[d] = (v[d] / m) * a.step;
            sdr[d] = v[d];
            a.dr[(size_t)c * 3 + d] = v[d];
        }
        sp = p_i;
        a.P_particle[c] = p_i;
        a.P_first[c] = 0;
        a.b0[c] = 0;
        a.logu[c] = log(UniformFromBits(r3[0], r3[1]));
    }
    __syncthreads();
    const int p = sp;
    for (int t = threadIdx.x; t < pv.M * 3; t += blockDim.x) {
        const int d = t / pv.M, b = t - d * pv.M;  // slice fastest: coalesced reads
        a.P[((size_t)c * pv.M + b) * 3 + d] = a.R[PosIndex(pv, a.N, c, p, d, b)] + sdr[d];
    }
}

// A/B on one box (tools/gpu_ab_disp.sh, ms per attempt over 1024 clones): 1024 threads x 64 registers (112 B of
// spills) 1.468, 768 x 80 1.396, 512 x 125 (no spills) 1.371 -- unlike K1, this kernel holds OLD and NEW state at once
#ifndef PIMC_DISP_THREADS
#define PIMC_DISP_THREADS 512
#endif
constexpr int kDispThreads = PIMC_DISP_THREADS;
constexpr int kDispWarps = kDispThreads / 32;
// Partner rows staged per warp through shared memory with cp.async, kDispStages partners ahead (0: each warp loads its
// partner rows straight into registers when it needs them -- the round-1 kernel, long-scoreboard bound at 16 warps per SM).
// The ring lives in the shared memory K1 keeps for its partner tile, which this kernel does not use.
#ifndef PIMC_DISP_STAGES
#define PIMC_DISP_STAGES 3
#endif
constexpr int kDispStages = PIMC_DISP_STAGES;
constexpr int kDispRingRow = 34;                                  // 33 slices of a chunk (+1 pad)
constexpr int kDispRingDoubles = 3 * kDispRingRow;                // one partner: 3 dims
constexpr size_t kDispRingBytes = (size_t)kDispWarps * (kDispStages > 0 ? kDispStages : 1) * kDispRingDoubles * sizeof(double);

__device__ __forceinline__ void CpAsync8(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void CpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void CpAsyncWait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct DisplacePairArgs {
    PathView pv;
    const double *R_moved, *R_partner;
    int N_moved, N_partner, same;
    const int32_t *particle;   // [C]
    const double *dr;          // [C][3]
    int n_chunks;
    // fast Ilkka evaluation (tables staged in shared memory) or the general evaluation
    FastTable FT;
    const unsigned char *fast_tables;
    PairTable T;
    const double *blob;
    int accumulate;            // add to the partial sums (second and later actions)
    double *partial;           // [C][n_chunks][2]: OLD, NEW
};

/// ATYPE < 0: fast Ilkka path.
template <int ATYPE>
static __global__ void __launch_bounds__(kDispThreads, 1) displace_pair_kernel(const DisplacePairArgs a) {
    extern __shared__ __align__(16) unsigned char dsm[];
    __shared__ double red[2][kDispWarps];
    __shared__ double sdr[3];                // the item's shift vector
    __shared__ double ring[kDispWarps][32];  // lane 31's parked r': slots 0..15 OLD, 16..31 NEW (FastIlkkaEvalWarpBoth)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const PathView &pv = a.pv;
    if (ATYPE < 0) {
        __shared__ unsigned long long stage_bar;
        StageBlockTma(dsm, a.fast_tables, a.FT.n_bytes, &stage_bar);
    }
    const SharedTab tb(dsm);
    // per-warp ring of staged partner rows behind the table block (fast path only)
    double *ring_q = reinterpret_cast<double *>(dsm + ((ATYPE < 0 ? a.FT.n_bytes : 0) + 15) / 16 * 16) + (size_t)warp * (kDispStages > 0 ? kDispStages : 1) * kDispRingDoubles;
    const int n_items = pv.C * a.n_chunks;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int c = item / a.n_chunks, chunk = item - c * a.n_chunks;
        const int b = chunk * 32 + lane;
        const bool lane_on = b < pv.M;
        const int bn = b + 1 >= pv.M ? b + 1 - pv.M : b + 1;
        const int p = a.particle[c];
        double p0[3], p1[3];
        __syncthreads();  // the previous item's shift has been read
        if (tid < 3) sdr[tid] = a.dr[(size_t)c * 3 + tid];
        __syncthreads();
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            p0[d] = lane_on ? a.R_moved[PosIndex(pv, a.N_moved, c, p, d, b)] : 0.;
            p1[d] = lane_on ? a.R_moved[PosIndex(pv, a.N_moved, c, p, d, bn)] : 0.;
        }
        const uint32_t sdr_addr = (uint32_t)__cvta_generic_to_shared(sdr);
        double acc_old = 0., acc_new = 0.;
        int n_parked = 0;                                     // warp-uniform
        const bool lane31_counts = chunk * 32 + 31 < pv.M;    // lane 31 holds a real link
        // -u_long/2 of the parked r' (n_parked OLD in lanes 0.., n_parked NEW in lanes 16..), one per lane
        auto flush_ring = [&]() {
            __syncwarp();
            double v = 0.;
            if ((lane & 15) < n_parked) v = -0.5 * FastLrEval(tb, a.FT, Clamp(ring[warp][lane], a.FT.lr.r_min, a.FT.lr.r_max));
            __syncwarp();
            if (lane < 16)
                acc_old += v;
            else
                acc_new += v;
            n_parked = 0;
        };
        // element i of a staged row = slice chunk * 32 + i; the element after the chunk's last link is that link's next
        // slice (the beta-periodic wrap included), fetched by lane 0
        const int n_valid = min(32, pv.M - chunk * 32);
        const int b_extra = chunk * 32 + n_valid >= pv.M ? chunk * 32 + n_valid - pv.M : chunk * 32 + n_valid;
        auto stage_partner = [&](int q, int slot) {
            double *dst = ring_q + (size_t)slot * kDispRingDoubles;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const double *src = a.R_partner + PosIndex(pv, a.N_partner, c, q, d, 0);
                if (lane_on)
                    CpAsync8(dst + d * kDispRingRow + lane, src + b);
                else
                    dst[d * kDispRingRow + lane + 1] = 0.;   // unused elements stay finite (they feed lookups of masked lanes)
                if (lane == 0) CpAsync8(dst + d * kDispRingRow + n_valid, src + b_extra);
            }
        };
        int q_stage = warp;   // next partner to stage
        if (ATYPE < 0 && kDispStages > 0) {
            __syncwarp();
#pragma unroll
            for (int st = 0; st < (kDispStages > 0 ? kDispStages - 1 : 0); ++st) {
                if (q_stage < a.N_partner) stage_partner(q_stage, st);
                CpAsyncCommit();
                q_stage += kDispWarps;
            }
        }
        int slot = 0;
        for (int q = warp; q < a.N_partner; q += kDispWarps) {
            double q0[3], q1[3];
            if (ATYPE < 0 && kDispStages > 0) {
                // keep kDispStages - 1 partners in flight, then wait for the oldest
                const int fill = slot == 0 ? kDispStages - 1 : slot - 1;
                if (q_stage < a.N_partner) stage_partner(q_stage, fill);
                CpAsyncCommit();
                q_stage += kDispWarps;
                CpAsyncWait<(kDispStages > 0 ? kDispStages - 1 : 0)>();
                __syncwarp();
                const double *src = ring_q + (size_t)slot * kDispRingDoubles + lane;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    q0[d] = src[d * kDispRingRow];
                    q1[d] = src[d * kDispRingRow + 1];
                }
                __syncwarp();   // the slot is read before a later iteration refills it
                slot = slot + 1 == kDispStages ? 0 : slot + 1;
                if (a.same && q == p) continue;
            } else {
                if (a.same && q == p) continue;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    q0[d] = lane_on ? a.R_partner[PosIndex(pv, a.N_partner, c, q, d, b)] : 0.;
                    q1[d] = lane_on ? a.R_partner[PosIndex(pv, a.N_partner, c, q, d, bn)] : 0.;
                }
            }
            // the shifted beads p + dr (rounded as the stored proposal is) are formed per step from a shift
            // re-read from shared memory (asm volatile: not hoisted): 12 registers fewer live across the loop
            double n0[3], n1[3];
            {
                double s0, s1, s2;
                asm volatile("ld.shared.f64 %0, [%3];\n\tld.shared.f64 %1, [%3+8];\n\tld.shared.f64 %2, [%3+16];"
                             : "=d"(s0), "=d"(s1), "=d"(s2)
                             : "r"(sdr_addr));
                n0[0] = p0[0] + s0;
                n0[1] = p0[1] + s1;
                n0[2] = p0[2] + s2;
                n1[0] = p1[0] + s0;
                n1[1] = p1[1] + s1;
                n1[2] = p1[2] + s2;
            }
            double r, rp, s, uo, un;
            if (ATYPE < 0) {
                double rn, rpn, sn;
                DrDrpDrrpFast(p0, q0, p1, q1, pv.box, r, rp, s);
                DrDrpDrrpFast(n0, q0, n1, q1, pv.box, rn, rpn, sn);
                // u_long(r') of a link is the next lane's u_long(r) (same bead pair one slice later): one
                // long-range lookup per lane and mode instead of two (A/B on one box: 1.644 -> 1.519 ms per attempt;
                // evaluating OLD and NEW one after the other through FastIlkkaEvalWarp: 1.617)
                FastIlkkaEvalWarpBoth(tb, a.FT, r, rp, s, rn, rpn, sn, lane, &ring[warp][n_parked], &ring[warp][16 + n_parked], uo, un);
                if (a.FT.use_lr && lane31_counts && ++n_parked == 16) flush_ring();
            } else {
                DrDrpDrrp(p0, q0, p1, q1, pv.box, r, rp, s);
                uo = PairEval<(ATYPE < 0 ? 0 : ATYPE), WHICH_U>(a.blob, a.T, r, rp, s);
                DrDrpDrrp(n0, q0, n1, q1, pv.box, r, rp, s);
                un = PairEval<(ATYPE < 0 ? 0 : ATYPE), WHICH_U>(a.blob, a.T, r, rp, s);
            }
            acc_old += lane_on ? uo : 0.;
            acc_new += lane_on ? un : 0.;
        }
        if (ATYPE < 0 && n_parked > 0) flush_ring();
        if (ATYPE < 0 && kDispStages > 0) CpAsyncWait<0>();   // nothing of this item is in flight into the ring any more
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc_old += __shfl_down_sync(0xffffffffu, acc_old, o);
            acc_new += __shfl_down_sync(0xffffffffu, acc_new, o);
        }
        __syncthreads();  // red of the previous item has been read
        if (lane == 0) {
            red[0][warp] = acc_old;
            red[1][warp] = acc_new;
        }
        __syncthreads();
        if (tid < 2) {
            double tot = 0.;
            for (int w = 0; w < kDispWarps; ++w) tot += red[tid][w];
            double *dst = a.partial + (size_t)item * 2 + tid;
            *dst = a.accumulate ? *dst + tot : tot;
        }
    }
}

/// One CTA per clone: Metropolis test (displace_particle_class.h:65-71) and Move::Accept.
static __global__ void __launch_bounds__(256) displace_decide_commit_kernel(PathView pv, int N, int n_k, int n_chunks, const double *__restrict__ partial,
                                                                     int use_lr, const double *__restrict__ lr_old, const double *__restrict__ lr_new,
                                                                     const double *__restrict__ logu, const double *__restrict__ P,
                                                                     const int32_t *__restrict__ P_particle, const double2 *__restrict__ drho,
                                                                     double *__restrict__ R, double2 *__restrict__ rho, int32_t *__restrict__ accept,
                                                                     long long *__restrict__ n_accept) {
    const int c = blockIdx.x;
    double old_action = 0., new_action = 0.;
    for (int i = 0; i < n_chunks; ++i) {  // every thread the same fixed-order sum
        old_action += partial[((size_t)c * n_chunks + i) * 2];
        new_action += partial[((size_t)c * n_chunks + i) * 2 + 1];
    }
    if (use_lr) {
        old_action += lr_old[c];
        new_action += lr_new[c];
    }
    const int acc = (old_action - new_action) < logu[c] ? 0 : 1;
    if (threadIdx.x == 0) {
        accept[c] = acc;
        n_accept[c] += acc;
    }
    if (!acc) return;
    const int p = P_particle[c];
    for (int t = threadIdx.x; t < pv.M * 3; t += blockDim.x) {
        const int d = t / pv.M, b = t - d * pv.M;
        R[PosIndex(pv, N, c, p, d, b)] = P[((size_t)c * pv.M + b) * 3 + d];
    }
    if (drho) {
        const size_t n = (size_t)pv.M * n_k;
        double2 *dst = rho + (size_t)c * n;
        const double2 *src = drho + (size_t)c * n;
        for (size_t t = threadIdx.x; t < n; t += blockDim.x) {
            double2 v = dst[t];
            v.x += src[t].x;
            v.y += src[t].y;
            dst[t] = v;
        }
    }
}

}  // namespace pimc

#endif  // SIMPIMC_B200_DISPLACE_CUH_
