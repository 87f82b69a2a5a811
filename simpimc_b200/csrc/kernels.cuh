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
//   positions  R[clone][particle][dim][slice (row padded to a multiple of 4)]   double
//              -- the imaginary-time index is fastest: a warp's 32 lanes walk 32 consecutive
//              slices of one particle pair, loads coalesce, and because paths are continuous
//              in imaginary time neighbouring lanes land in the same spline cells
//   rho_k      rho[clone][slice][k]                                              double2 (re, im)
//   proposal   P[clone][bead of the proposal][dim]                               double
//   drho       D[clone][window slice][k]                                         double2
#ifndef SIMPIMC_B200_KERNELS_CUH_
#define SIMPIMC_B200_KERNELS_CUH_

#include "device_math.cuh"

namespace pimc {

/// One species' committed positions plus its pending proposal.
struct SpeciesView {
    const double *R;     // committed positions
    int N;
    // proposal overlay (NEW mode): up to kMaxPropSlots particles of the species per clone (one for
    // Bisect / DisplaceParticle, the particles of a cycle for permutation moves), all with n_prop beads
    const double *P;          // [slot][C][n_prop][3]
    const int32_t *P_particle;  // [slot][C]
    const int32_t *P_first;     // [slot][C] first bead (global slice index)
    int n_prop;                 // 0 = no pending proposal
    int n_slots;                // proposals pending (0 with n_prop > 0 is read as 1)
};

constexpr int kMaxPropSlots = 16;  // members of a permutation cycle (<= 8) plus, when the window rolls over the seam, their end points' labels

struct PathView {
    int C;        // clones
    int M;        // slices of the whole path
    int Mloc;     // slices owned by this context
    int Mstore;   // slices stored (Mloc, +1 halo when sharded)
    int Ms;       // row length of the position arrays (Mstore rounded up to a multiple of 4)
    int slice_lo; // first owned slice
    int sharded;
    // Disjoint bisection windows of one walker moved in the same launch (pimc_bisect_sweep_windows): the move kernels
    // then run over C = n_walkers * vdiv "virtual clones" -- index c = walker * vdiv + window for everything an
    // attempt owns (proposal, window start, partial sums), walker = c / vdiv for positions and rho_k.  1 elsewhere.
    int vdiv;
    Box box;
};

/// Walker whose positions / rho_k virtual clone c works on.
__device__ __forceinline__ int RealClone(const PathView &pv, int c) { return pv.vdiv > 1 ? c / pv.vdiv : c; }

/// Global slice index after the beta-periodic wrap (bg < 2 M).  A slice shard's move windows never
/// leave its stored slices [slice_lo, slice_hi]; the halo -- global slice slice_hi, which equals M
/// on the last rank -- is stored at local index Mloc and must NOT wrap, so sharded views do not.
__device__ __forceinline__ int WrapSlice(const PathView &pv, int bg) { return (!pv.sharded && bg >= pv.M) ? bg - pv.M : bg; }

/// Local storage index of the slice that follows local slice b.
__device__ __forceinline__ int NextSlice(const PathView &pv, int b) {
    return pv.sharded ? b + 1 : (b + 1 == pv.M ? 0 : b + 1);
}

__device__ __forceinline__ size_t PosIndex(const PathView &pv, int N, int c, int p, int d, int b) {
    return (((size_t)c * N + p) * 3 + d) * pv.Ms + b;
}

/// Position of (species view, clone c, particle p, GLOBAL slice bg) in OLD or NEW mode.
__device__ __forceinline__ void LoadPos(const PathView &pv, const SpeciesView &sv, int c, int p, int bg, int mode, double out[3]) {
    const int bl = WrapSlice(pv, bg);
    if (mode && sv.n_prop > 0) {
        const int ns = sv.n_slots > 0 ? sv.n_slots : 1;
        for (int sl = 0; sl < ns; ++sl) {
            if (sv.P_particle[(size_t)sl * pv.C + c] != p) continue;
            int off = bl - sv.P_first[(size_t)sl * pv.C + c];
            if (off < 0) off += pv.M;
            if (off < sv.n_prop) {
                const double *q = sv.P + (((size_t)sl * pv.C + c) * sv.n_prop + off) * 3;
                out[0] = q[0];
                out[1] = q[1];
                out[2] = q[2];
                return;
            }
        }
    }
    const int b = bl - pv.slice_lo, cr = RealClone(pv, c);
#pragma unroll
    for (int d = 0; d < 3; ++d) out[d] = sv.R[PosIndex(pv, sv.N, cr, p, d, b)];
}

// ------------------------------------------------------------------------------------ K1
constexpr int kPairThreads = 1024;  // one persistent CTA per SM, 32 warps
constexpr int kPairWarps = kPairThreads / 32;
constexpr int kChunk = 32;          // links per work item = lanes of a warp
constexpr int kQTile = 64;          // partner particles staged per shared-memory tile
constexpr int kRow = 34;            // 33 slices of a chunk (+1 pad) per staged row

struct PairFullArgs {
    PathView pv;
    SpeciesView A, B;
    int same;               // species_a == species_b
    PairTable T;
    const double *blob;     // stageable table blob in global memory
    int blob_doubles;
    int stage;              // 1: copy the blob into shared memory first
    int independent_images; // Potential(): r and r' minimum-imaged independently (App. A-6)
    int n_chunks;           // ceil(Mloc / 32)
    int n_pgroups;          // ceil(Na / 32)
    double *partial;        // [C][n_chunks][n_pgroups]
};

/// Storage index of local slice s (s == Mloc is the slice after the shard: wraps to 0 on an
/// unsharded path, is the halo otherwise); -1 beyond.
__device__ __forceinline__ int StoreIndex(const PathView &pv, int s) {
    if (s < pv.Mloc) return s;
    if (s == pv.Mloc) return pv.sharded ? s : 0;
    return -1;
}

/// Persistent CTAs, one per SM.  A work item is (clone, 32-link chunk, group of 32 particles
/// of species a): warp w owns particle p of the group, its 32 lanes own the chunk's links, and
/// the warp walks over the partner particles q, staged tile by tile in shared memory.  The
/// tables' 1-D parts, grids and LUTs are staged once per CTA.
template <int ATYPE, int WHICH>
static __global__ void __launch_bounds__(kPairThreads, 1) pair_full_kernel(const PairFullArgs a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ double red[kPairWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *qpos = smem;  // [kQTile][3][kRow]
    const double *tab = a.blob;
    if (a.stage) {
        double *stab = smem + kQTile * 3 * kRow;
        for (int i = tid; i < a.blob_doubles; i += kPairThreads) stab[i] = a.blob[i];
        tab = stab;
    }
    __syncthreads();
    const PathView &pv = a.pv;
    const int Na = a.A.N, Nb = a.B.N;
    const int half = Na / 2;
    const bool even = (Na & 1) == 0;
    const int per_clone = a.n_chunks * a.n_pgroups;
    const int n_items = pv.C * per_clone;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int c = item / per_clone;
        const int rem = item - c * per_clone;
        const int chunk = rem / a.n_pgroups, pg = rem - chunk * a.n_pgroups;
        const int s0 = chunk * kChunk;            // first local slice of the chunk
        const int p = pg * kPairWarps + warp;
        const bool warp_on = p < Na;
        const bool lane_on = s0 + lane < pv.Mloc;
        double p0[3] = {0., 0., 0.}, p1[3] = {0., 0., 0.};
        if (warp_on && lane_on) {
            const int i0 = s0 + lane, i1 = StoreIndex(pv, s0 + lane + 1);
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                p0[d] = a.A.R[PosIndex(pv, Na, c, p, d, i0)];
                p1[d] = a.A.R[PosIndex(pv, Na, c, p, d, i1)];
            }
        }
        double acc = 0.;
        for (int q0 = 0; q0 < Nb; q0 += kQTile) {
            const int nq = min(kQTile, Nb - q0);
            __syncthreads();  // the previous tile has been consumed
            for (int row = warp; row < nq * 3; row += kPairWarps) {
                const int qq = row / 3, d = row - qq * 3;
                const double *src = a.B.R + PosIndex(pv, Nb, c, q0 + qq, d, 0);
                const int i0 = StoreIndex(pv, s0 + lane);
                qpos[row * kRow + lane] = i0 >= 0 ? src[i0] : 0.;
                if (lane == 0) {
                    const int i1 = StoreIndex(pv, s0 + kChunk);
                    qpos[row * kRow + kChunk] = i1 >= 0 ? src[i1] : 0.;
                }
            }
            __syncthreads();
            if (!warp_on) continue;
            for (int qq = 0; qq < nq; ++qq) {
                const int q = q0 + qq;
                if (a.same) {
                    // every unordered pair once: q = p + dd (mod N) for dd = 1..N/2, the last
                    // offset only from the lower half when N is even (warp-uniform test)
                    int dd = q - p;
                    if (dd < 0) dd += Na;
                    if (dd == 0 || dd > half || (even && dd == half && p >= half)) continue;
                }
                if (!lane_on) continue;
                const double *row = qpos + (qq * 3) * kRow + lane;
                const double b0[3] = {row[0], row[kRow], row[2 * kRow]};
                const double b1[3] = {row[1], row[kRow + 1], row[2 * kRow + 1]};
                double r, rp, s;
                if (a.independent_images) {
                    r = Mag3(MinImage(p0[0] - b0[0], pv.box), MinImage(p0[1] - b0[1], pv.box), MinImage(p0[2] - b0[2], pv.box));
                    rp = Mag3(MinImage(p1[0] - b1[0], pv.box), MinImage(p1[1] - b1[1], pv.box), MinImage(p1[2] - b1[2], pv.box));
                    s = 0.;
                } else {
                    DrDrpDrrp(p0, b0, p1, b1, pv.box, r, rp, s);
                }
                acc += PairEval<ATYPE, WHICH>(tab, a.T, r, rp, s);
            }
        }
        const double tot = BlockSum<kPairThreads>(acc, red);
        if (tid == 0) a.partial[item] = tot;
    }
}

/// out[c] = sum_b partial[c][b] in slice order, then the long-range term and constants in the
/// reference's association: tot + ((ksum [*2]) + k_0 + r_0).
static __global__ void finalize_kernel(const double *__restrict__ partial, int C, int Mloc, const double *__restrict__ lr_sum,
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
static __global__ void __launch_bounds__(256) rhok_build_kernel(PathView pv, SpeciesView sv, KSpaceView ks, int chunk, double2 *__restrict__ rho) {
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
                    PhaseTable(sv.R[PosIndex(pv, sv.N, c, p0 + pp, d, b)], ks.kbox, ks.max_index, ctab + (size_t)(pp * 3 + d) * tl);
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

// K2 v2: the k set of KSpace::Setup is a half ball, so it factors into COLUMNS (i_x, i_y) that
// each hold the lattice indices i_z = -z_max .. z_max (the column (0, 0) only i_z > 0).  With
// A_p = c_x[i_x] c_y[i_y] of particle p and z_p = c_z[|i_z|],
//     rho(i_x, i_y, +|i_z|) = sum_p A_p z_p       = (S1 - S2) + i (S3 + S4)
//     rho(i_x, i_y, -|i_z|) = sum_p A_p conj(z_p) = (S1 + S2) + i (S4 - S3)
// with S1 = sum Re A Re z, S2 = sum Im A Im z, S3 = sum Re A Im z, S4 = sum Im A Re z: four FMAs
// per particle serve BOTH k vectors, and c_z is the same address for every lane (one broadcast
// LDS.128).  Lane = column, warp = slice; per particle and warp 22 FP64 instructions and 6 LDS
// instead of 182 x (10 FP64 + 3 LDS) / 32 = 57 + 17 of the thread-per-k form above.  The sums run
// over particles in index order but associate differently from species_class.h:391-395
// (products are added component-wise): |difference| <= a few N ulp, inside the 1e-12 N bound the
// parity tests hold rho_k to.
struct KColsView {
    int n_k, n_cols, n_groups;     // n_groups = ceil(n_cols / 32) warps share one slice
    const int32_t *col_info;       // [n_cols] |i_x| | |i_y| << 8 | (i_x < 0) << 16 | (i_y < 0) << 17
    const int32_t *kmap;           // [n_cols][2 TM + 1] index of the k vector (i_x, i_y, i_z), -1 = absent
    double kbox;
    int stage_out;                 // n_k complex numbers fit in one slice's table region
    int p_split, p_chunk;          // particles split p_split ways (p_chunk each, a multiple of 32): a launch with
                                   // few (clone, slice) items -- one slice-sharded path -- still fills the GPU;
                                   // part ps writes its partial sums to rho + ps * C * Mloc * n_k
};

constexpr int kColsWarps = 8;
__host__ __device__ constexpr int ColsEntries(int TM) { return (3 * TM + 2) | 1; }  // odd: 16-byte records of 8 lanes hit disjoint banks

__device__ __forceinline__ double FlipSign(double v, int flip) { return __hiloint2double(__double2hiint(v) ^ flip, __double2loint(v)); }

template <int TM>
static __global__ void __launch_bounds__(kColsWarps * 32) rhok_build_cols_kernel(PathView pv, SpeciesView sv, KColsView kc, double2 *__restrict__ rho) {
    constexpr int E = ColsEntries(TM);
    extern __shared__ __align__(16) double2 ctab[];  // [slice of the CTA][32 particles][E]
    const int G = kc.n_groups, S = kColsWarps / G;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = warp / G, g = warp - s * G;
    const int groups_per_clone = (pv.Mloc + S - 1) / S;
    const int ps = blockIdx.x % kc.p_split, blk = blockIdx.x / kc.p_split;
    const int c = blk / groups_per_clone;
    const int b = (blk - c * groups_per_clone) * S + s;
    const int p_begin = ps * kc.p_chunk, p_end = min(sv.N, p_begin + kc.p_chunk);
    const bool valid = s < S && b < pv.Mloc;
    const int col = g * 32 + lane;
    const bool col_ok = col < kc.n_cols;
    const int info = col_ok ? kc.col_info[col] : 0;
    const int xo = info & 0xff, yo = (TM + 1) + ((info >> 8) & 0xff);
    const int fx = (info >> 16) & 1 ? (int)0x80000000 : 0, fy = (info >> 17) & 1 ? (int)0x80000000 : 0;
    double2 *T = ctab + (size_t)(s < S ? s : 0) * 32 * E;
    double a0r = 0., a0i = 0.;
    double S1[TM], S2[TM], S3[TM], S4[TM];
#pragma unroll
    for (int j = 0; j < TM; ++j) S1[j] = S2[j] = S3[j] = S4[j] = 0.;
    for (int p0 = p_begin; p0 < p_end; p0 += 32) {
        if (G > 1) __syncthreads(); else __syncwarp();
        if (g == 0 && valid && p0 + lane < p_end) {
            // KSpace::CalcC (k_space_class.h:83-94): c[j] = e^{i phi} c[j-1], entries 0..TM (z: 1..TM)
            double2 *t = T + lane * E;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                double sn, cs;
                sincos(sv.R[PosIndex(pv, sv.N, c, p0 + lane, d, b)] * kc.kbox, &sn, &cs);
                double2 cur = make_double2(1., 0.);
                double2 *td = d < 2 ? t + d * (TM + 1) : t + 2 * TM + 1;
                if (d < 2) td[0] = cur;
#pragma unroll
                for (int j = 1; j <= TM; ++j) {
                    cur = make_double2(cs * cur.x - sn * cur.y, cs * cur.y + sn * cur.x);
                    td[j] = cur;
                }
            }
        }
        if (G > 1) __syncthreads(); else __syncwarp();
        if (valid) {
            const int np = min(32, p_end - p0);
#pragma unroll 4
            for (int pp = 0; pp < np; ++pp) {
                const double2 *t = T + pp * E;
                const double2 X = t[xo], Y = t[yo];
                const double xi = FlipSign(X.y, fx), yi = FlipSign(Y.y, fy);
                const double ar = X.x * Y.x - xi * yi, ai = X.x * yi + xi * Y.x;
                a0r += ar;
                a0i += ai;
#pragma unroll
                for (int j = 0; j < TM; ++j) {
                    const double2 Z = t[2 * TM + 2 + j];
                    S1[j] = fma(ar, Z.x, S1[j]);
                    S2[j] = fma(ai, Z.y, S2[j]);
                    S3[j] = fma(ar, Z.y, S3[j]);
                    S4[j] = fma(ai, Z.x, S4[j]);
                }
            }
        }
    }
    // emit: through the slice's table region when it is large enough (coalesced 16-byte stores)
    double2 *out = rho + (((size_t)ps * pv.C + c) * pv.Mloc + (valid ? b : 0)) * kc.n_k;
    double2 *dst = kc.stage_out ? T : out;
    if (kc.stage_out) { if (G > 1) __syncthreads(); else __syncwarp(); }
    if (valid && col_ok) {
        const int32_t *km = kc.kmap + (size_t)col * (2 * TM + 1);
        int k = km[TM];
        if (k >= 0) dst[k] = make_double2(a0r, a0i);
#pragma unroll
        for (int j = 0; j < TM; ++j) {
            k = km[TM + 1 + j];
            if (k >= 0) dst[k] = make_double2(S1[j] - S2[j], S3[j] + S4[j]);
            k = km[TM - 1 - j];
            if (k >= 0) dst[k] = make_double2(S1[j] + S2[j], S4[j] - S3[j]);
        }
    }
    if (kc.stage_out) {
        if (G > 1) __syncthreads(); else __syncwarp();
        if (valid)
            for (int k = g * 32 + lane; k < kc.n_k; k += 32 * G) out[k] = T[k];
    }
}

/// rho[i] = sum over the particle parts of a split build, in part order.
static __global__ void rhok_reduce_kernel(const double2 *__restrict__ part, int n_parts, size_t n, double2 *__restrict__ rho) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double2 acc = part[i];
    for (int ps = 1; ps < n_parts; ++ps) {
        const double2 v = part[(size_t)ps * n + i];
        acc.x += v.x;
        acc.y += v.y;
    }
    rho[i] = acc;
}

/// drho(c, j, k) = sum over the species' pending proposals of rho_bead(new position) -
/// rho_bead(old position) at window slice j (global slice b0[c] + j); a proposal that does not
/// cover the slice contributes zero.
static __global__ void __launch_bounds__(256) rhok_delta_kernel(PathView pv, SpeciesView sv, KSpaceView ks, const int32_t *__restrict__ b0,
                                                        int n_window, double2 *__restrict__ drho) {
    extern __shared__ __align__(16) double2 ctab[];  // [2][3][2m+1]
    const int tl = 2 * ks.max_index + 1;
    const int tid = threadIdx.x;
    const int ns = sv.n_slots > 0 ? sv.n_slots : 1;
    for (int item = blockIdx.x; item < pv.C * n_window; item += gridDim.x) {
        const int c = item / n_window, j = item - c * n_window;
        const int bg = b0[c] + j;
        bool wrote = false;
        for (int sl = 0; sl < ns; ++sl) {
            const int p = sv.P_particle[(size_t)sl * pv.C + c];
            if (p < 0) continue;  // unused slot (uniform over the CTA)
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
                double2 *dst = drho + ((size_t)c * n_window + j) * ks.n_k + k;
                double2 v = make_double2(fn.x - fo.x, fn.y - fo.y);
                if (wrote) {  // same thread wrote this element for the previous slot
                    v.x += dst->x;
                    v.y += dst->y;
                }
                *dst = v;
            }
            wrote = true;
        }
        if (!wrote)
            for (int k = tid; k < ks.n_k; k += blockDim.x) drho[((size_t)c * n_window + j) * ks.n_k + k] = make_double2(0., 0.);
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
    int accumulate;                 // add to out[c] instead of overwriting it
    double *out;                    // [C]
};

/// out[c] = scale * (twice?2:1) * sum_k sum_b w_k Re(rho_a rho_b^*).
static __global__ void __launch_bounds__(256) ksum_kernel(const KSumArgs a) {
    __shared__ double red[256 / 32];
    const int c = blockIdx.x;
    const PathView &pv = a.pv;
    const int nb = a.b0 ? a.n_window : pv.Mloc;
    double acc = 0.;
    if (!a.b0 && !a.drho_a && !a.drho_b) {
        // whole shard, committed rho_k: every (k, slice) element is read once -- keep four slices'
        // loads in flight per thread (one CTA per clone cannot hide the HBM latency otherwise) and
        // read a same-species rho_k once; the accumulation order (k outer, slices in order) stays
        const bool same = a.rho_a == a.rho_b;
        for (int k = threadIdx.x; k < a.n_k; k += blockDim.x) {
            const double w = a.wk[k];
            const double2 *pa = a.rho_a + (size_t)c * pv.Mloc * a.n_k + k;
            const double2 *pb = a.rho_b + (size_t)c * pv.Mloc * a.n_k + k;
            int j = 0;
            for (; j + 4 <= nb; j += 4) {
                double2 ra[4], rb[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) ra[u] = pa[(size_t)(j + u) * a.n_k];
#pragma unroll
                for (int u = 0; u < 4; ++u) rb[u] = same ? ra[u] : pb[(size_t)(j + u) * a.n_k];
#pragma unroll
                for (int u = 0; u < 4; ++u) acc += w * (ra[u].x * rb[u].x + ra[u].y * rb[u].y);
            }
            for (; j < nb; ++j) {
                const double2 ra = pa[(size_t)j * a.n_k];
                const double2 rb = same ? ra : pb[(size_t)j * a.n_k];
                acc += w * (ra.x * rb.x + ra.y * rb.y);
            }
        }
    } else
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
        if (a.accumulate)
            a.out[c] += a.scale * tot;
        else
            a.out[c] = a.scale * tot;
    }
}

// ------------------------------------------------------------------------------------ K4
struct PairWindowArgs {
    PathView pv;
    SpeciesView A, B;
    int same;
    int n_a, n_b;               // listed ("moved") particles of species a / b per clone, <= kMaxPropSlots each
    const int32_t *part_a;      // [n_a][C]
    const int32_t *part_b;      // [n_b][C]
    const int32_t *b0;          // [C] window start (global slice)
    int n_links;                // window length (level 0: one link per slice)
    int mode;
    PairTable T;
    const double *blob;
    double *partial;            // [C][n_links]
    const int32_t *alive = nullptr;  // [C] (optional): clones whose flag is 0 are skipped, their partial sums left untouched
};

/// One CTA per (clone, link): all pairs that touch a listed particle (PairAction::
/// GenerateParticlePairs, pair_action_class.h:63-114), OLD or NEW positions.  Same species:
/// (listed, other) and (listed_i, listed_j), i < j.  Different species: (listed a, every b),
/// then (every unlisted a, listed b).
template <int ATYPE>
static __global__ void __launch_bounds__(128) pair_window_kernel(const PairWindowArgs a) {
    __shared__ double red[128 / 32];
    const PathView &pv = a.pv;
    for (int item = blockIdx.x; item < pv.C * a.n_links; item += gridDim.x) {
        const int c = item / a.n_links, j = item - c * a.n_links;
        if (a.alive && !a.alive[c]) continue;  // uniform over the CTA
        const int bg = a.b0[c] + j;
        // a negative entry is an unused slot (the device-resident permutation move lists a different number of labels per clone)
        int la[kMaxPropSlots], lb[kMaxPropSlots];
        for (int i = 0; i < a.n_a; ++i) la[i] = a.part_a[(size_t)i * pv.C + c];
        for (int i = 0; i < a.n_b; ++i) lb[i] = a.part_b[(size_t)i * pv.C + c];
        double acc = 0.;
        // pairs (listed a, partner of species b)
        for (int i = 0; i < a.n_a; ++i) {
            const int m = la[i];
            if (m < 0) continue;
            double m0[3], m1[3];
            LoadPos(pv, a.A, c, m, bg, a.mode, m0);
            LoadPos(pv, a.A, c, m, bg + 1, a.mode, m1);
            for (int q = threadIdx.x; q < a.B.N; q += blockDim.x) {
                if (a.same) {
                    if (q == m) continue;
                    bool earlier = false;  // (listed_i, listed_j) once: only from the lower list index
                    for (int i2 = 0; i2 < i; ++i2) earlier = earlier || la[i2] == q;
                    if (earlier) continue;
                }
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
        // pairs (unlisted a partner, listed b)
        if (!a.same) {
            for (int i = 0; i < a.n_b; ++i) {
                const int m = lb[i];
                if (m < 0) continue;
                double m0[3], m1[3];
                LoadPos(pv, a.B, c, m, bg, a.mode, m0);
                LoadPos(pv, a.B, c, m, bg + 1, a.mode, m1);
                for (int p = threadIdx.x; p < a.A.N; p += blockDim.x) {
                    bool listed = false;
                    for (int i2 = 0; i2 < a.n_a; ++i2) listed = listed || la[i2] == p;
                    if (listed) continue;
                    double p0[3], p1[3];
                    LoadPos(pv, a.A, c, p, bg, a.mode, p0);
                    LoadPos(pv, a.A, c, p, bg + 1, a.mode, p1);
                    double r, rp, s;
                    DrDrpDrrp(p0, m0, p1, m1, pv.box, r, rp, s);
                    acc += PairEval<ATYPE, WHICH_U>(a.blob, a.T, r, rp, s);
                }
            }
        }
        const double tot = BlockSum<128>(acc, red);
        if (threadIdx.x == 0) a.partial[item] = tot;
        __syncthreads();
    }
}

/// Per-pair test hook: out[i] = Calc{U,dUdBeta,V}(r[i], rp[i], s[i]).
template <int ATYPE, int WHICH>
static __global__ void calc_pair_kernel(const double *__restrict__ blob, PairTable T, int n, const double *__restrict__ r,
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
static __global__ void __launch_bounds__(256) gofr_kernel(const GofrArgs a) {
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
                const double x = __dsub_rn(a.A.R[PosIndex(pv, a.A.N, c, p, d, b)], a.B.R[PosIndex(pv, a.B.N, c, q, d, b)]);
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

/// K5 v2.  Work item = (clone, chunk of 32 slices, tile I of species a, tile J of species b):
/// the tiles' positions are staged in shared memory with coalesced loads (32 consecutive slices
/// per row), warp w owns slice w of the chunk, its lanes own 32 particles p of tile I and walk
/// the particles q of tile J together, so q's position is one broadcast LDS per axis for 32
/// pairs and p's position stays in registers.  Rows are 33 doubles apart (odd lane stride:
/// conflict-free).  Each warp counts into its own shared histogram when n_r allows (lanes hold
/// different pairs: few same-bin collisions), else into one per CTA.  Same arithmetic as
/// gofr_kernel: bins are bit-exact and the counts are integers, so the order does not matter.
constexpr int kGofrThreads = 1024;
constexpr int kGofrRow = 33;

struct GofrTiledArgs {
    GofrArgs g;
    int T;            // particles per tile
    int n_ti, n_tj;   // tiles of species a / b
    int n_chunks;     // ceil(Mloc / 32)
    int warp_hist;    // 1: one histogram per warp
};

static __global__ void __launch_bounds__(kGofrThreads, 1) gofr_tiled_kernel(const GofrTiledArgs t) {
    extern __shared__ __align__(16) unsigned char gsm[];
    const GofrArgs &a = t.g;
    const PathView &pv = a.pv;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *ti = reinterpret_cast<double *>(gsm);                // [T][3][33]
    double *tj = ti + (size_t)t.T * 3 * kGofrRow;                // [T][3][33]
    unsigned int *hist = reinterpret_cast<unsigned int *>(tj + (size_t)t.T * 3 * kGofrRow);
    unsigned int *my_hist = hist + (t.warp_hist ? warp * a.n_r : 0);
    const int n_hist = (t.warp_hist ? kGofrThreads / 32 : 1) * a.n_r;
    const int Na = a.A.N, Nb = a.B.N;
    // tile pairs: same species -> I <= J
    const int n_tp = a.same ? t.n_ti * (t.n_ti + 1) / 2 : t.n_ti * t.n_tj;
    const int per_clone = t.n_chunks * n_tp;
    const int n_items = pv.C * per_clone;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int c = item / per_clone;
        int rem = item - c * per_clone;
        const int chunk = rem / n_tp;
        rem -= chunk * n_tp;
        int I, J;
        if (a.same) {
            I = 0;
            while (rem >= t.n_ti - I) {
                rem -= t.n_ti - I;
                ++I;
            }
            J = I + rem;
        } else {
            I = rem / t.n_tj;
            J = rem - I * t.n_tj;
        }
        const int s0 = chunk * 32;
        const int p_lo = I * t.T, q_lo = J * t.T;
        const int np = min(t.T, Na - p_lo), nq = min(t.T, Nb - q_lo);
        const bool diag = a.same && I == J;
        __syncthreads();  // previous item fully consumed
        for (int i = tid; i < n_hist; i += kGofrThreads) hist[i] = 0u;
        for (int row = warp; row < np * 3; row += kGofrThreads / 32) {
            const int pp = row / 3, d = row - pp * 3;
            ti[row * kGofrRow + lane] = s0 + lane < pv.Mloc ? a.A.R[PosIndex(pv, Na, c, p_lo + pp, d, s0 + lane)] : 0.;
        }
        if (!diag)
            for (int row = warp; row < nq * 3; row += kGofrThreads / 32) {
                const int qq = row / 3, d = row - qq * 3;
                tj[row * kGofrRow + lane] = s0 + lane < pv.Mloc ? a.B.R[PosIndex(pv, Nb, c, q_lo + qq, d, s0 + lane)] : 0.;
            }
        __syncthreads();
        const double *qt = diag ? ti : tj;
        if (s0 + warp < pv.Mloc) {
            for (int pb = 0; pb < np; pb += 32) {
                const int pp = pb + lane;
                const bool p_on = pp < np;
                const double *prow = ti + (size_t)(p_on ? pp : 0) * 3 * kGofrRow + warp;
                const double px = prow[0], py = prow[kGofrRow], pz = prow[2 * kGofrRow];
                // same tile: pairs p < q only; the warp starts at the first q any of its lanes needs
                for (int qq = diag ? pb + 1 : 0; qq < nq; ++qq) {
                    const double *qrow = qt + (size_t)qq * 3 * kGofrRow + warp;
                    double dr[3];
                    const double pxyz[3] = {px, py, pz};
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const double x = __dsub_rn(pxyz[d], qrow[d * kGofrRow]);
                        dr[d] = __dsub_rn(x, __dmul_rn(rint(__dmul_rn(x, pv.box.iL)), pv.box.L));
                    }
                    const double dist = Mag3Exact(dr[0], dr[1], dr[2]);
                    const double arg = __dsub_rn(__dmul_rn(__dsub_rn(dist, a.r_min), a.d_ir), 0.5);
                    const double ri = rint(arg);
                    if (p_on && (!diag || pp < qq) && ri >= 0. && ri < (double)a.n_r) atomicAdd(&my_hist[(unsigned int)ri], 1u);
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < a.n_r; i += kGofrThreads) {
            unsigned long long tot = 0;
            if (t.warp_hist) {
                for (int w = 0; w < kGofrThreads / 32; ++w) tot += hist[w * a.n_r + i];
            } else {
                tot = hist[i];
            }
            if (tot) atomicAdd(&a.counts[(size_t)c * a.n_r + i], tot);
        }
    }
}

// ------------------------------------------------------------------------------------ K6
/// sk[c][k] += cofactor[c] * CMag2(rho_a, rho_b) for b = 0..Mloc-1 in order (no FMA contraction:
/// the accumulation then matches the reference bit for bit given equal rho_k).
static __global__ void sofk_kernel(PathView pv, int n_k, const double2 *__restrict__ rho_a, const double2 *__restrict__ rho_b,
                            const double *__restrict__ kmag, double k_cut, const double *__restrict__ cofactor,
                            double *__restrict__ sk) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (k >= n_k || !(kmag[k] < k_cut)) return;
    const double cf = cofactor ? cofactor[c] : 1.0;
    double acc = sk[(size_t)c * n_k + k];
    const bool same = rho_a == rho_b;
    const double2 *pa = rho_a + (size_t)c * pv.Mloc * n_k + k, *pb = rho_b + (size_t)c * pv.Mloc * n_k + k;
    int b = 0;
    for (; b + 4 <= pv.Mloc; b += 4) {  // four slices' loads in flight; the accumulation stays in slice order
        double2 ra[4], rb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) ra[u] = pa[(size_t)(b + u) * n_k];
#pragma unroll
        for (int u = 0; u < 4; ++u) rb[u] = same ? ra[u] : pb[(size_t)(b + u) * n_k];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double m2 = __dadd_rn(__dmul_rn(ra[u].x, rb[u].x), __dmul_rn(ra[u].y, rb[u].y));
            acc = __dadd_rn(acc, __dmul_rn(cf, m2));
        }
    }
    for (; b < pv.Mloc; ++b) {
        const double2 ra = pa[(size_t)b * n_k];
        const double2 rb = same ? ra : pb[(size_t)b * n_k];
        const double m2 = __dadd_rn(__dmul_rn(ra.x, rb.x), __dmul_rn(ra.y, rb.y));
        acc = __dadd_rn(acc, __dmul_rn(cf, m2));
    }
    sk[(size_t)c * n_k + k] = acc;
}

// --------------------------------------------------------------------- permutation table
/// PermBisectIterative::UpdatePermTable (perm_bisect_iterative_class.h:10-30) / the table of
/// PermBisectTable (perm_bisect_table_class.h:36-50): t(i, j) = exp(e) if e > log_epsilon else 0,
/// e = -|Dr(r_i(b0), r_j(b1))|^2 / (4 lambda tau n) [+ |Dr(r_i(b0), r_i(b1))|^2 / (4 lambda tau n) for the
/// table variant], b1 = b0 + n_bisect_beads.  One CTA per (clone, i), threads over j; the path
/// is taken unpermuted (the bead n slices ahead of particle j is particle j's).
static __global__ void __launch_bounds__(128) perm_table_kernel(PathView pv, const double *__restrict__ R, int N, const int32_t *__restrict__ b0,
                                                        int n_bisect_beads, double i_4_lambda_tau_n, double log_epsilon, int relative,
                                                        double *__restrict__ t) {
    const int c = blockIdx.x / N, i = blockIdx.x - c * N;
    const int bs = WrapSlice(pv, b0[c]) - pv.slice_lo;
    int be = b0[c] + n_bisect_beads;
    while (be >= pv.M) be -= pv.M;
    be -= pv.slice_lo;
    double ri[3], d_ii = 0.;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        ri[d] = R[PosIndex(pv, N, c, i, d, bs)];
        const double x = MinImage(ri[d] - R[PosIndex(pv, N, c, i, d, be)], pv.box);
        d_ii += x * x;
    }
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        double d_ij = 0.;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double x = MinImage(ri[d] - R[PosIndex(pv, N, c, j, d, be)], pv.box);
            d_ij += x * x;
        }
        const double e = relative ? (-d_ij + d_ii) * i_4_lambda_tau_n : (-d_ij) * i_4_lambda_tau_n;
        t[((size_t)c * N + i) * N + j] = e > log_epsilon ? exp(e) : 0.;
    }
}

// ------------------------------------------------------------------------- data movement
/// host order R[clone][particle][bead][dim] -> device order R[clone][particle][dim][slice].
static __global__ void positions_in_kernel(const double *__restrict__ src, int n_clones, int N, int Mstore, int Ms, double *__restrict__ dst) {
    const size_t total = (size_t)n_clones * N * 3 * Ms;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int b = i % Ms;
        size_t r = i / Ms;
        const int d = r % 3;
        const size_t cp = r / 3;  // clone * N + particle
        dst[i] = b < Mstore ? src[(cp * Mstore + b) * 3 + d] : 0.0;
    }
}
static __global__ void positions_out_kernel(const double *__restrict__ src, int n_clones, int N, int Mstore, int Ms, double *__restrict__ dst) {
    const size_t total = (size_t)n_clones * N * Mstore * 3;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int d = i % 3;
        size_t r = i / 3;
        const int b = r % Mstore;
        const size_t cp = r / Mstore;
        dst[i] = src[(cp * 3 + d) * Ms + b];
    }
}

/// out[c][j][d] = committed position of particle[c] at slice b_first[c] + j (mod M).
static __global__ void gather_beads_kernel(PathView pv, const double *__restrict__ R, int N, const int32_t *__restrict__ particle,
                                    const int32_t *__restrict__ b_first, int n_beads, double *__restrict__ out) {
    const int c = blockIdx.x;
    const int p = particle[c];
    for (int t = threadIdx.x; t < n_beads * 3; t += blockDim.x) {
        const int j = t / 3, d = t - j * 3;
        int bg = b_first[c] + j;
        while (bg >= pv.M) bg -= pv.M;
        out[((size_t)c * n_beads + j) * 3 + d] = R[PosIndex(pv, N, c, p, d, bg - pv.slice_lo)];
    }
}

/// Move::Accept for the clones whose accept flag is set: committed positions take the
/// proposal, committed rho_k takes rho_k + drho on the window slices.
static __global__ void commit_positions_kernel(PathView pv, int N, const double *__restrict__ P, const int32_t *__restrict__ P_particle,
                                        const int32_t *__restrict__ P_first, int n_prop, int n_slots, const int32_t *__restrict__ accept,
                                        double *__restrict__ R) {
    const int c = blockIdx.x;
    if (!accept[c]) return;
    for (int sl = 0; sl < n_slots; ++sl) {
        const int p = P_particle[(size_t)sl * pv.C + c];
        for (int t = threadIdx.x; t < n_prop * 3; t += blockDim.x) {
            const int j = t / 3, d = t - j * 3;
            int bg = P_first[(size_t)sl * pv.C + c] + j;
            if (bg >= pv.M) bg -= pv.M;
            const int b = bg - pv.slice_lo;
            if (b < 0 || b >= pv.Mstore) continue;
            R[PosIndex(pv, N, c, p, d, b)] = P[(((size_t)sl * pv.C + c) * n_prop + j) * 3 + d];
        }
    }
}
static __global__ void commit_rhok_kernel(PathView pv, int n_k, const double2 *__restrict__ drho, const int32_t *__restrict__ b0, int n_window,
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

/// Slice-shard halo: buf[clone][particle][dim] <-> one stored slice of the position array.
static __global__ void halo_pack_kernel(const double *__restrict__ R, size_t n_rows, int Ms, int slot, double *__restrict__ buf) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n_rows) buf[i] = R[i * Ms + slot];
}
static __global__ void halo_unpack_kernel(double *__restrict__ R, size_t n_rows, int Ms, int slot, const double *__restrict__ buf) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n_rows) R[i * Ms + slot] = buf[i];
}

/// Ring rotation of a slice-sharded path by `shift` slices: every shard hands its first `shift`
/// owned slices to the previous rank and appends the ones it receives, i.e. the global slice labels
/// rotate while every shard keeps its range -- imaginary time is a ring, so actions and estimators
/// are unchanged, and the slices that were shard boundaries (never moved by shard-interior windows)
/// become interior.  buf is [row][shift], row = (clone, particle, dim).
static __global__ void rotate_pack_kernel(const double *__restrict__ R, size_t n_rows, int Ms, int shift, double *__restrict__ buf) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n_rows * shift) return;
    const size_t row = i / shift;
    const int j = (int)(i - row * shift);
    buf[i] = R[row * Ms + j];
}
/// One thread per row: slide the owned slices left by `shift` (ascending order: in place) and put the
/// received ones at the end.  The halo slot is left stale (the caller exchanges halos next).
static __global__ void rotate_apply_kernel(double *__restrict__ R, size_t n_rows, int Ms, int Mloc, int shift, const double *__restrict__ buf) {
    const size_t row = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    double *r = R + row * Ms;
    for (int i = 0; i + shift < Mloc; ++i) r[i] = r[i + shift];
    for (int j = 0; j < shift; ++j) r[Mloc - shift + j] = buf[row * shift + j];
}

/// Dependent-FMA chains: FP64 pipe throughput (2 flop per FMA).
static __global__ void fp64_peak_kernel(double *out, int iters) {
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
