// sm_100a kernels of the action-evaluation path.  See DESIGN.md for the layout and the
// roofline of each kernel.
//
//   K1 pair_full_kernel     PairAction::DActionDBeta / Potential / whole-path GetAction
//                           (pair_action_class.h:241-264, 369-395, 267-302)
//   K2 rhok_build_kernel    Species::InitRhoK + Bead::CalcRhoK + KSpace::CalcC
//                           (species_class.h:380-403, bead_class.h:125-133, k_space_class.h:83-94)
//      rhok_delta_kernel    Species::UpdateRhoK (species_class.h:406-425)
//   K3 ksum_kernel          Calc{ULong,dUdBetaLong,VLong} (ilkka_pair_action_class.h:104-122,152-169,57-74)
//   K4 pair_window_kernel   PairAction::GetAction over a move's slice window (pair_action_class.h:267-302)
//   K5 gofr_kernel          PairCorrelation::Accumulate (pair_correlation_class.h:15-28)
//   K6 sofk_kernel          StructureFactor::Accumulate (structure_factor_class.h:15-32)
//
// Device layouts (private to the library):
//   positions  R[clone][slice][dim][particle (padded to a multiple of 4)]       double
//   rho_k      rho[clone][slice][k]                                              double2 (re, im)
//   proposal   P[clone][bead of the proposal][dim]                               double
//   drho       D[clone][window slice][k]                                         double2
#ifndef SIMPIMC_B200_KERNELS_CUH_
#define SIMPIMC_B200_KERNELS_CUH_

#include "device_math.cuh"

namespace pimc {

constexpr int kPairThreads = 512;
constexpr int kPairCtasPerSm = 2;  // 2 x 512 threads x 64 registers fill the register file

/// One species' committed positions plus its pending proposal.
struct SpeciesView {
    const double *R;     // committed positions
    int N, Npad;
    // proposal overlay (NEW mode)
    const double *P;          // [C][n_prop][3]
    const int32_t *P_particle;  // [C]
    const int32_t *P_first;     // [C] first bead (global slice index)
    int n_prop;                 // 0 = no pending proposal
};

struct PathView {
    int C;        // clones
    int M;        // slices of the whole path
    int Mloc;     // slices owned by this context
    int Mstore;   // slices stored (Mloc, +1 halo when sharded)
    int slice_lo; // first owned slice
    int sharded;
    Box box;
};

/// Local storage index of the slice that follows local slice b.
__device__ __forceinline__ int NextSlice(const PathView &pv, int b) {
    return pv.sharded ? b + 1 : (b + 1 == pv.M ? 0 : b + 1);
}

__device__ __forceinline__ size_t PosIndex(const PathView &pv, int Npad, int c, int b, int d, int p) {
    return (((size_t)c * pv.Mstore + b) * 3 + d) * Npad + p;
}

/// Position of (species view, clone c, particle p, GLOBAL slice bg) in OLD or NEW mode.
__device__ __forceinline__ void LoadPos(const PathView &pv, const SpeciesView &sv, int c, int p, int bg, int mode, double out[3]) {
    int bl = bg >= pv.M ? bg - pv.M : bg;
    if (mode && sv.n_prop > 0 && sv.P_particle[c] == p) {
        int off = bl - sv.P_first[c];
        if (off < 0) off += pv.M;
        if (off < sv.n_prop) {
            const double *q = sv.P + ((size_t)c * sv.n_prop + off) * 3;
            out[0] = q[0];
            out[1] = q[1];
            out[2] = q[2];
            return;
        }
    }
    const int b = bl - pv.slice_lo;
#pragma unroll
    for (int d = 0; d < 3; ++d) out[d] = sv.R[PosIndex(pv, sv.Npad, c, b, d, p)];
}

// ------------------------------------------------------------------------------------ K1
struct PairFullArgs {
    PathView pv;
    SpeciesView A, B;
    int same;               // species_a == species_b
    PairTable T;
    const double *blob;     // table blob in global memory
    int blob_doubles;
    int stage;              // 1: copy the blob into shared memory first
    int independent_images; // Potential(): r and r' minimum-imaged independently (App. A-6)
    double *partial;        // [C][Mloc]
};

/// Persistent CTAs: stage the table once, then walk (clone, slice) items.  Per item the two
/// slices' positions are staged in shared memory (SoA) and the threads stride over pairs.
template <int ATYPE, int WHICH>
__global__ void __launch_bounds__(kPairThreads, kPairCtasPerSm) pair_full_kernel(const PairFullArgs a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ double red[kPairThreads / 32];
    const int tid = threadIdx.x;
    double *pos = smem;  // [2 slices][2 species][3][Npad]
    const int NpA = a.A.Npad, NpB = a.B.Npad;
    const int pos_doubles = 2 * 3 * (NpA + (a.same ? 0 : NpB));
    const double *tab = a.blob;
    if (a.stage) {
        double *stab = smem + pos_doubles;
        for (int i = tid; i < a.blob_doubles; i += kPairThreads) stab[i] = a.blob[i];
        tab = stab;
    }
    __syncthreads();
    const PathView &pv = a.pv;
    const int n_items = pv.C * pv.Mloc;
    const int Na = a.A.N, Nb = a.B.N;
    // same species: particle p pairs with (p+d) mod N for d = 1..N/2 (the last offset only for
    // the first half when N is even) -- every unordered pair exactly once
    const int half = Na / 2;
    const int n_work = a.same ? ((Na & 1) ? Na * half : Na * (half - 1) + half) : Na * Nb;
    double *xa0 = pos, *xa1 = pos + 3 * NpA;
    double *xb0 = a.same ? xa0 : pos + 6 * NpA, *xb1 = a.same ? xa1 : pos + 6 * NpA + 3 * NpB;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int c = item / pv.Mloc, b = item - c * pv.Mloc;
        const int b1 = NextSlice(pv, b);
        __syncthreads();
        {
            const double *s0 = a.A.R + PosIndex(pv, NpA, c, b, 0, 0);
            const double *s1 = a.A.R + PosIndex(pv, NpA, c, b1, 0, 0);
            for (int i = tid; i < 3 * NpA; i += kPairThreads) {
                xa0[i] = s0[i];
                xa1[i] = s1[i];
            }
            if (!a.same) {
                const double *u0 = a.B.R + PosIndex(pv, NpB, c, b, 0, 0);
                const double *u1 = a.B.R + PosIndex(pv, NpB, c, b1, 0, 0);
                for (int i = tid; i < 3 * NpB; i += kPairThreads) {
                    xb0[i] = u0[i];
                    xb1[i] = u1[i];
                }
            }
        }
        __syncthreads();
        double acc = 0.;
        for (int w = tid; w < n_work; w += kPairThreads) {
            int p, q;
            if (a.same) {
                const int d = w / Na;
                p = w - d * Na;
                q = p + d + 1;
                if (q >= Na) q -= Na;
            } else {
                p = w / Nb;
                q = w - p * Nb;
            }
            const double pa0[3] = {xa0[p], xa0[NpA + p], xa0[2 * NpA + p]};
            const double pa1[3] = {xa1[p], xa1[NpA + p], xa1[2 * NpA + p]};
            const double pb0[3] = {xb0[q], xb0[NpB + q], xb0[2 * NpB + q]};
            const double pb1[3] = {xb1[q], xb1[NpB + q], xb1[2 * NpB + q]};
            double r, rp, s;
            if (a.independent_images) {
                r = Mag3(MinImage(pa0[0] - pb0[0], pv.box), MinImage(pa0[1] - pb0[1], pv.box), MinImage(pa0[2] - pb0[2], pv.box));
                rp = Mag3(MinImage(pa1[0] - pb1[0], pv.box), MinImage(pa1[1] - pb1[1], pv.box), MinImage(pa1[2] - pb1[2], pv.box));
                s = 0.;
            } else {
                DrDrpDrrp(pa0, pb0, pa1, pb1, pv.box, r, rp, s);
            }
            acc += PairEval<ATYPE, WHICH>(tab, a.T, r, rp, s);
        }
        const double tot = BlockSum<kPairThreads>(acc, red);
        if (tid == 0) a.partial[item] = tot;
    }
}

/// out[c] = sum_b partial[c][b] in slice order, then the long-range term and constants in the
/// reference's association: tot + ((ksum [*2]) + k_0 + r_0).
__global__ void finalize_kernel(const double *__restrict__ partial, int C, int Mloc, const double *__restrict__ lr_sum,
                                int add_lr, double lr_const_k, double lr_const_r, int add_const, double *__restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double tot = 0.;
    for (int b = 0; b < Mloc; ++b) tot += partial[(size_t)c * Mloc + b];
    if (add_lr) {
        double l = lr_sum[c];
        if (add_const) l = l + lr_const_k + lr_const_r;
        tot += l;
    }
    out[c] = tot;
}

// ------------------------------------------------------------------------------------ K2
struct KSpaceView {
    int n_k;
    int max_index;          // same for every dimension (cubic box)
    const int32_t *kidx;    // [n_k][3] table positions max_index + lattice index
    double kbox;            // 2 pi / L
};

/// Per-axis phase table of one position: c[m] = 1, c[m+j] = e^{i phi} c[m+j-1], c[m-j] = conj.
__device__ __forceinline__ void PhaseTable(double x, double kbox, int m, double2 *tab /* 2m+1 */) {
    double sn, cs;
    sincos(x * kbox, &sn, &cs);
    double2 cur = make_double2(1., 0.);
    tab[m] = cur;
    for (int j = 1; j <= m; ++j) {
        // complex product (cs + i sn) * cur, written as the four-multiply form
        const double re = cs * cur.x - sn * cur.y;
        const double im = cs * cur.y + sn * cur.x;
        cur = make_double2(re, im);
        tab[m + j] = cur;
        tab[m - j] = make_double2(re, -im);
    }
}

__device__ __forceinline__ double2 CMul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

/// rho_k(c, b) = sum_p prod_d table_d[kidx(k, d)], particles added in index order.
/// One CTA per (clone, slice); the phase tables of `chunk` particles at a time live in shared
/// memory and thread k accumulates its k vector over them.
__global__ void __launch_bounds__(256) rhok_build_kernel(PathView pv, SpeciesView sv, KSpaceView ks, int chunk, double2 *__restrict__ rho) {
    extern __shared__ __align__(16) double2 ctab[];  // [chunk][3][2m+1]
    const int tl = 2 * ks.max_index + 1;
    const int tid = threadIdx.x;
    for (int item = blockIdx.x; item < pv.C * pv.Mloc; item += gridDim.x) {
        const int c = item / pv.Mloc, b = item - c * pv.Mloc;
        for (int k0 = 0; k0 < ks.n_k; k0 += blockDim.x) {
            const int k = k0 + tid;
            int i0 = 0, i1 = 0, i2 = 0;
            if (k < ks.n_k) {
                i0 = ks.kidx[3 * k];
                i1 = ks.kidx[3 * k + 1];
                i2 = ks.kidx[3 * k + 2];
            }
            double2 acc = make_double2(0., 0.);
            for (int p0 = 0; p0 < sv.N; p0 += chunk) {
                const int np = min(chunk, sv.N - p0);
                __syncthreads();
                for (int t = tid; t < np * 3; t += blockDim.x) {
                    const int pp = t / 3, d = t - pp * 3;
                    PhaseTable(sv.R[PosIndex(pv, sv.Npad, c, b, d, p0 + pp)], ks.kbox, ks.max_index, ctab + (size_t)(pp * 3 + d) * tl);
                }
                __syncthreads();
                if (k < ks.n_k) {
                    for (int pp = 0; pp < np; ++pp) {
                        const double2 *tb = ctab + (size_t)pp * 3 * tl;
                        const double2 f = CMul(CMul(tb[i0], tb[tl + i1]), tb[2 * tl + i2]);
                        acc.x += f.x;
                        acc.y += f.y;
                    }
                }
            }
            if (k < ks.n_k) rho[((size_t)c * pv.Mloc + b) * ks.n_k + k] = acc;
        }
    }
}

/// drho(c, j, k) = rho_bead(new position) - rho_bead(old position) of the proposal's particle
/// at window slice j (global slice b0[c] + j), zero where the proposal does not cover it.
__global__ void __launch_bounds__(256) rhok_delta_kernel(PathView pv, SpeciesView sv, KSpaceView ks, const int32_t *__restrict__ b0,
                                                        int n_window, double2 *__restrict__ drho) {
    extern __shared__ __align__(16) double2 ctab[];  // [2][3][2m+1]
    const int tl = 2 * ks.max_index + 1;
    const int tid = threadIdx.x;
    for (int item = blockIdx.x; item < pv.C * n_window; item += gridDim.x) {
        const int c = item / n_window, j = item - c * n_window;
        const int bg = b0[c] + j;
        const int p = sv.P_particle[c];
        __syncthreads();
        if (tid < 6) {
            const int mode = tid / 3, d = tid - mode * 3;
            double r[3];
            LoadPos(pv, sv, c, p, bg, mode, r);
            PhaseTable(r[d], ks.kbox, ks.max_index, ctab + (size_t)(mode * 3 + d) * tl);
        }
        __syncthreads();
        for (int k = tid; k < ks.n_k; k += blockDim.x) {
            const int i0 = ks.kidx[3 * k], i1 = ks.kidx[3 * k + 1], i2 = ks.kidx[3 * k + 2];
            const double2 fo = CMul(CMul(ctab[i0], ctab[tl + i1]), ctab[2 * tl + i2]);
            const double2 *tn = ctab + 3 * tl;
            const double2 fn = CMul(CMul(tn[i0], tn[tl + i1]), tn[2 * tl + i2]);
            drho[((size_t)c * n_window + j) * ks.n_k + k] = make_double2(fn.x - fo.x, fn.y - fo.y);
        }
    }
}

// ------------------------------------------------------------------------------------ K3
struct KSumArgs {
    PathView pv;
    int n_k;
    const double2 *rho_a, *rho_b;   // committed rho_k of the two species
    const double2 *drho_a, *drho_b; // NEW-mode increments (or nullptr)
    const double *wk;               // [n_k] table weight per k vector
    const int32_t *b0;              // window start per clone (nullptr: whole shard)
    int n_window;
    int twice;                      // species differ
    double scale;                   // Bare CalcULong: level_tau
    double *out;                    // [C]
};

/// out[c] = scale * (twice?2:1) * sum_k sum_b w_k Re(rho_a rho_b^*).
__global__ void __launch_bounds__(256) ksum_kernel(const KSumArgs a) {
    __shared__ double red[256 / 32];
    const int c = blockIdx.x;
    const PathView &pv = a.pv;
    const int nb = a.b0 ? a.n_window : pv.Mloc;
    double acc = 0.;
    for (int k = threadIdx.x; k < a.n_k; k += blockDim.x) {
        const double w = a.wk[k];
        for (int j = 0; j < nb; ++j) {
            int b = j;
            if (a.b0) {
                b = a.b0[c] + j;
                if (b >= pv.M) b -= pv.M;
                b -= pv.slice_lo;
            }
            double2 ra = a.rho_a[((size_t)c * pv.Mloc + b) * a.n_k + k];
            double2 rb = a.rho_b[((size_t)c * pv.Mloc + b) * a.n_k + k];
            if (a.drho_a) {
                const double2 d = a.drho_a[((size_t)c * a.n_window + j) * a.n_k + k];
                ra.x += d.x;
                ra.y += d.y;
            }
            if (a.drho_b) {
                const double2 d = a.drho_b[((size_t)c * a.n_window + j) * a.n_k + k];
                rb.x += d.x;
                rb.y += d.y;
            }
            acc += w * (ra.x * rb.x + ra.y * rb.y);
        }
    }
    double tot = BlockSum<256>(acc, red);
    if (threadIdx.x == 0) {
        if (a.twice) tot *= 2.;
        a.out[c] = a.scale * tot;
    }
}

// ------------------------------------------------------------------------------------ K4
struct PairWindowArgs {
    PathView pv;
    SpeciesView A, B;
    int same;
    int moved_a, moved_b;       // 1 if the moved list holds a particle of species a / b
    const int32_t *part_a;      // [C] moved particle of species a (if moved_a)
    const int32_t *part_b;      // [C]
    const int32_t *b0;          // [C] window start (global slice)
    int n_links;                // window length (level 0: one link per slice)
    int mode;
    PairTable T;
    const double *blob;
    double *partial;            // [C][n_links]
};

/// One CTA per (clone, link): all pairs that touch a moved particle, OLD or NEW positions.
template <int ATYPE>
__global__ void __launch_bounds__(128) pair_window_kernel(const PairWindowArgs a) {
    __shared__ double red[128 / 32];
    const PathView &pv = a.pv;
    for (int item = blockIdx.x; item < pv.C * a.n_links; item += gridDim.x) {
        const int c = item / a.n_links, j = item - c * a.n_links;
        const int bg = a.b0[c] + j;
        double acc = 0.;
        // pairs (moved a, every b partner)
        if (a.moved_a) {
            const int m = a.part_a[c];
            double m0[3], m1[3];
            LoadPos(pv, a.A, c, m, bg, a.mode, m0);
            LoadPos(pv, a.A, c, m, bg + 1, a.mode, m1);
            for (int q = threadIdx.x; q < a.B.N; q += blockDim.x) {
                if (a.same && q == m) continue;
                double q0[3], q1[3];
                LoadPos(pv, a.B, c, q, bg, a.mode, q0);
                LoadPos(pv, a.B, c, q, bg + 1, a.mode, q1);
                double r, rp, s;
                // the reference stores same-species pairs as (moved, other) and differences
                // are taken as second minus first; the three magnitudes do not depend on order
                DrDrpDrrp(m0, q0, m1, q1, pv.box, r, rp, s);
                acc += PairEval<ATYPE, WHICH_U>(a.blob, a.T, r, rp, s);
            }
        }
        // pairs (every a partner, moved b), skipping the moved a particle counted above
        if (!a.same && a.moved_b) {
            const int m = a.part_b[c];
            const int skip = a.moved_a ? a.part_a[c] : -1;
            double m0[3], m1[3];
            LoadPos(pv, a.B, c, m, bg, a.mode, m0);
            LoadPos(pv, a.B, c, m, bg + 1, a.mode, m1);
            for (int p = threadIdx.x; p < a.A.N; p += blockDim.x) {
                if (p == skip) continue;
                double p0[3], p1[3];
                LoadPos(pv, a.A, c, p, bg, a.mode, p0);
                LoadPos(pv, a.A, c, p, bg + 1, a.mode, p1);
                double r, rp, s;
                DrDrpDrrp(p0, m0, p1, m1, pv.box, r, rp, s);
                acc += PairEval<ATYPE, WHICH_U>(a.blob, a.T, r, rp, s);
            }
        }
        const double tot = BlockSum<128>(acc, red);
        if (threadIdx.x == 0) a.partial[item] = tot;
        __syncthreads();
    }
}

/// Per-pair test hook: out[i] = Calc{U,dUdBeta,V}(r[i], rp[i], s[i]).
template <int ATYPE, int WHICH>
__global__ void calc_pair_kernel(const double *__restrict__ blob, PairTable T, int n, const double *__restrict__ r,
                                 const double *__restrict__ rp, const double *__restrict__ s, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = PairEval<ATYPE, WHICH>(blob, T, r[i], rp[i], s[i]);
}

// ------------------------------------------------------------------------------------ K5
struct GofrArgs {
    PathView pv;
    SpeciesView A, B;
    int same;
    double r_min, d_ir;
    int n_r;
    unsigned long long *counts;  // [C][n_r]
};

/// Histogram of minimum-image distances; bin = (uint32)rint((|dr| - r_min) * d_ir - 0.5), with
/// the distance and the bin argument computed without FMA contraction (bit-exact bins).
__global__ void __launch_bounds__(256) gofr_kernel(const GofrArgs a) {
    extern __shared__ unsigned int hist[];  // [n_r]
    const PathView &pv = a.pv;
    const int Na = a.A.N, Nb = a.B.N;
    const int half = Na / 2;
    const int n_work = a.same ? ((Na & 1) ? Na * half : Na * (half - 1) + half) : Na * Nb;
    for (int item = blockIdx.x; item < pv.C * pv.Mloc; item += gridDim.x) {
        const int c = item / pv.Mloc, b = item - c * pv.Mloc;
        __syncthreads();
        for (int i = threadIdx.x; i < a.n_r; i += blockDim.x) hist[i] = 0u;
        __syncthreads();
        for (int w = threadIdx.x; w < n_work; w += blockDim.x) {
            int p, q;
            if (a.same) {
                const int d = w / Na;
                p = w - d * Na;
                q = p + d + 1;
                if (q >= Na) q -= Na;
            } else {
                p = w / Nb;
                q = w - p * Nb;
            }
            double dr[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const double x = __dsub_rn(a.A.R[PosIndex(pv, a.A.Npad, c, b, d, p)], a.B.R[PosIndex(pv, a.B.Npad, c, b, d, q)]);
                dr[d] = __dsub_rn(x, __dmul_rn(rint(__dmul_rn(x, pv.box.iL)), pv.box.L));
            }
            const double dist = Mag3Exact(dr[0], dr[1], dr[2]);
            const double arg = __dsub_rn(__dmul_rn(__dsub_rn(dist, a.r_min), a.d_ir), 0.5);
            // (uint32_t) of a negative double is undefined in C++; on x86-64 the reference's
            // cvttsd2si path wraps to a huge value that fails i < n_r, i.e. the sample is dropped
            const double ri = rint(arg);
            if (ri >= 0. && ri < (double)a.n_r) atomicAdd(&hist[(unsigned int)ri], 1u);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < a.n_r; i += blockDim.x)
            if (hist[i]) atomicAdd(&a.counts[(size_t)c * a.n_r + i], (unsigned long long)hist[i]);
    }
}

// ------------------------------------------------------------------------------------ K6
/// sk[c][k] += cofactor[c] * CMag2(rho_a, rho_b) for b = 0..Mloc-1 in order (no FMA contraction:
/// the accumulation then matches the reference bit for bit given equal rho_k).
__global__ void sofk_kernel(PathView pv, int n_k, const double2 *__restrict__ rho_a, const double2 *__restrict__ rho_b,
                            const double *__restrict__ kmag, double k_cut, const double *__restrict__ cofactor,
                            double *__restrict__ sk) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (k >= n_k || !(kmag[k] < k_cut)) return;
    const double cf = cofactor ? cofactor[c] : 1.0;
    double acc = sk[(size_t)c * n_k + k];
    for (int b = 0; b < pv.Mloc; ++b) {
        const double2 ra = rho_a[((size_t)c * pv.Mloc + b) * n_k + k];
        const double2 rb = rho_b[((size_t)c * pv.Mloc + b) * n_k + k];
        const double m2 = __dadd_rn(__dmul_rn(ra.x, rb.x), __dmul_rn(ra.y, rb.y));
        acc = __dadd_rn(acc, __dmul_rn(cf, m2));
    }
    sk[(size_t)c * n_k + k] = acc;
}

// ------------------------------------------------------------------------- data movement
/// host order R[clone][particle][bead][dim] -> device order R[clone][slice][dim][particle].
__global__ void positions_in_kernel(const double *__restrict__ src, int n_clones, int N, int Npad, int Mstore, double *__restrict__ dst) {
    const size_t total = (size_t)n_clones * Mstore * 3 * Npad;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int p = i % Npad;
        size_t r = i / Npad;
        const int d = r % 3;
        r /= 3;
        const int b = r % Mstore;
        const size_t c = r / Mstore;
        dst[i] = p < N ? src[((c * N + p) * Mstore + b) * 3 + d] : 0.0;
    }
}
__global__ void positions_out_kernel(const double *__restrict__ src, int n_clones, int N, int Npad, int Mstore, double *__restrict__ dst) {
    const size_t total = (size_t)n_clones * N * Mstore * 3;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int d = i % 3;
        size_t r = i / 3;
        const int b = r % Mstore;
        r /= Mstore;
        const int p = r % N;
        const size_t c = r / N;
        dst[i] = src[((c * Mstore + b) * 3 + d) * Npad + p];
    }
}

/// out[c][j][d] = committed position of particle[c] at slice b_first[c] + j (mod M).
__global__ void gather_beads_kernel(PathView pv, const double *__restrict__ R, int Npad, const int32_t *__restrict__ particle,
                                    const int32_t *__restrict__ b_first, int n_beads, double *__restrict__ out) {
    const int c = blockIdx.x;
    const int p = particle[c];
    for (int t = threadIdx.x; t < n_beads * 3; t += blockDim.x) {
        const int j = t / 3, d = t - j * 3;
        int bg = b_first[c] + j;
        while (bg >= pv.M) bg -= pv.M;
        out[((size_t)c * n_beads + j) * 3 + d] = R[PosIndex(pv, Npad, c, bg - pv.slice_lo, d, p)];
    }
}

/// Move::Accept for the clones whose accept flag is set: committed positions take the
/// proposal, committed rho_k takes rho_k + drho on the window slices.
__global__ void commit_positions_kernel(PathView pv, int Npad, const double *__restrict__ P, const int32_t *__restrict__ P_particle,
                                        const int32_t *__restrict__ P_first, int n_prop, const int32_t *__restrict__ accept,
                                        double *__restrict__ R) {
    const int c = blockIdx.x;
    if (!accept[c]) return;
    const int p = P_particle[c];
    for (int t = threadIdx.x; t < n_prop * 3; t += blockDim.x) {
        const int j = t / 3, d = t - j * 3;
        int bg = P_first[c] + j;
        if (bg >= pv.M) bg -= pv.M;
        const int b = bg - pv.slice_lo;
        if (b < 0 || b >= pv.Mstore) continue;
        R[PosIndex(pv, Npad, c, b, d, p)] = P[((size_t)c * n_prop + j) * 3 + d];
    }
}
__global__ void commit_rhok_kernel(PathView pv, int n_k, const double2 *__restrict__ drho, const int32_t *__restrict__ b0, int n_window,
                                   const int32_t *__restrict__ accept, double2 *__restrict__ rho) {
    const int c = blockIdx.y;
    if (!accept[c]) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_window * n_k) return;
    const int j = t / n_k, k = t - j * n_k;
    int bg = b0[c] + j;
    if (bg >= pv.M) bg -= pv.M;
    const int b = bg - pv.slice_lo;
    double2 *dst = rho + ((size_t)c * pv.Mloc + b) * n_k + k;
    const double2 d = drho[((size_t)c * n_window + j) * n_k + k];
    dst->x += d.x;
    dst->y += d.y;
}

/// Dependent-FMA chains: FP64 pipe throughput (2 flop per FMA).
__global__ void fp64_peak_kernel(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1., a2 = a0 + 2., a3 = a0 + 3., a4 = a0 + 4., a5 = a0 + 5., a6 = a0 + 6., a7 = a0 + 7.;
    const double m = 1.0000001, k = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, k);
        a1 = fma(a1, m, k);
        a2 = fma(a2, m, k);
        a3 = fma(a3, m, k);
        a4 = fma(a4, m, k);
        a5 = fma(a5, m, k);
        a6 = fma(a6, m, k);
        a7 = fma(a7, m, k);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace pimc

#endif  // SIMPIMC_B200_KERNELS_CUH_
