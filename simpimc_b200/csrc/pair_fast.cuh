// Fast path of the Ilkka pair action (U and dU/dbeta): every table the evaluation touches
// lives in shared memory and every access is sized for the LSU, which -- not the FP64 pipe --
// bounds the first versions of this kernel (profiles/r1_k1_v3_*.txt: 145 LSU wavefronts per
// warp-evaluation, 80 % of the pipe).  Measured on B200 (tools/microbench.cu,
// profiles/r1_microbench.txt): an LDS.128 costs 2 wavefronts when <= 16 distinct
// bank-disjoint addresses are read, 4 for 32; an L1-hit LDG.128 costs 4 even fully broadcast
// and one per distinct line beyond; records 32 B apart conflict 2-way.  Hence:
//
//   * interval search = ONE fused multiply-add: key = low word of fma(x, 1/h, 2^52+2^51)
//     (round-to-nearest integer of x/h), a uint16 table gives the interval of the bucket's
//     lower edge, and because h is smaller than the smallest knot spacing at most one knot
//     lies inside a bucket: one compare against g[i+1] finishes the search.  The result is the
//     einspline interval index exactly (integer work: bit-exact, tested against the oracle);
//   * knots are stored as pairs (g[i], g[i+1]) so that one LDS.128 feeds the compare and
//     tau = x - g[i];
//   * 1-D pp coefficients are split into two 16-byte-stride arrays (c0,c1) and (c2,c3);
//   * the bicubic cell polynomials of the block of cells the box can reach are staged with a
//     144-byte record stride (neighbouring cells fall into disjoint banks); cells outside the
//     block are read from global memory;
//   * sqrt = MUFU.RSQ64H seed + one coupled Newton step + one Markstein correction (7 FP64
//     instructions, no slow-path branch).
//
// Reference semantics: Path::DrDrpDrrp (src/data_structures/path_class.h:137-150),
// IlkkaPairAction::CalcU / CalcdUdBeta (src/actions/pair_action/ilkka_pair_action_class.h:77-101,
// 125-149), PairAction::SetLimits (pair_action_class.h:32-42).
#ifndef SIMPIMC_B200_PAIR_FAST_CUH_
#define SIMPIMC_B200_PAIR_FAST_CUH_

#include "device_math.cuh"

// -DPIMC_EXPERIMENT=7 builds K1 without the neighbour-lane reuse of the long-range spline
// (A/B timing, tools/gpu_ab.sh)
#ifndef PIMC_EXPERIMENT
#define PIMC_EXPERIMENT 0
#endif
// partner steps per trip of K1's inner loop (A/B: 2 with 512 threads x 128 registers, tools/gpu_ab.sh)
#ifndef PIMC_K1_UNROLL
#define PIMC_K1_UNROLL 1
#endif
// 1: x / y interval of the 2-D table from the packed interval table (ULookup16: one load, then knot and cell together).
// Measured (tools/gpu_ab_lr2.sh, same box): K1 44.77 -> 44.64 ms, displace 1.310 -> 1.326 ms per attempt -- inside the
// run-to-run noise, so it stays OFF and the x / y intervals remain einspline's bit for bit (ULookup).
#ifndef PIMC_XY16
#define PIMC_XY16 0
#endif
// 1: 1-D splines with a uniform interval table in bucket-centred form (FastPP1); 0: interval form everywhere (A/B)
#ifndef PIMC_LR2
#define PIMC_LR2 1
#endif

namespace pimc {

constexpr double kRoundMagic = 6755399441055744.0;  // 2^52 + 2^51: low word of (x + magic) = rint(x)
constexpr int kCellRecord = 144;                    // bytes per staged bicubic cell (128 + 16 pad)

/// Uniform-bucket interval table; offsets are bytes from the start of the staged table block.
struct ULutDesc {
    int off_lut;    // uint16 [n_keys]
    int key_max;    // n_keys - 1
    double inv_h;   // 1 / bucket width
    int shift, key0;  // bit-pattern buckets (KIND 1, see FastLut): key = (high word of x >> shift) - key0
    // packed form (PIMC_XY16, the x / y grids of the 2-D table): entry k = interval of bucket k's lower edge << 16 | position
    // of the bucket's knot in 1/32768 of the bucket width from its lower edge (0xFFFF: none), so the interval is known
    // after ONE load and the knot g[i] (for tau) is fetched beside the cell's coefficients instead of before them
    int off_lut32;  // uint32 [n_keys]
    double inv_h16; // 32768 / bucket width
    double x_cap;   // key_max * bucket width: keys of larger x are those of x_cap (the last bucket lies above the grid end)
};
/// A 1-D pp-form spline.  Two layouts:
///   interval form (bit-pattern interval tables, and every table without PIMC_LR2): lut -> interval i, knot pairs
///   (g[i], g[i+1]) for the finishing compare and tau = x - g[i], coefficients of piece i in off_c01 / off_c23;
///   BUCKET-CENTRED form (uniform interval tables with PIMC_LR2): the table's uniform buckets
///   [(k - 1/2) h, (k + 1/2) h) -- h below the smallest knot spacing, so at most one knot per bucket -- carry the
///   polynomial themselves.  Record k (off_c01 / off_c23, n_keys + 1 of them) is the cubic piece that holds the bucket's
///   lower edge, re-expanded (long double) about the bucket centre k h; knot[k] (off_knot) is the position of the
///   bucket's knot in 1/65536 of h from its lower edge (0xFFFF: none).  x at or above the knot belongs to the piece that
///   holds the NEXT bucket's lower edge: record k + 1, evaluated at the negative offset x - (k + 1) h.  One small load
///   decides, then the two coefficient loads go out together: the knot-pair load of the interval form (one of three
///   16-byte gathers at ~8.7 wavefronts each when 32 lanes sit in 32 unrelated intervals) and one level of the
///   dependent chain are gone (K1: 46.50 -> 44.82 ms, same-box A/B).  The compare is made on a 16-bit fraction of the
///   bucket: within h / 65536 of a knot either neighbouring piece may be taken, and there the two cubics differ by
///   (jump of the third derivative) (h / 65536)^3 / 6 -- a natural spline is C2 -- far below one ulp of the value
///   (host emulation against the interval form, knots and their floating-point neighbours included: 1e-16).
struct FastPP1 {
    ULutDesc lut;
    int off_gpair;  // double2 [n]: (g[i], g[i+1]); g[n] = +inf
    int off_c01;    // double2 [n]: (c0, c1) of interval i          | bucket-centred: [n_keys + 1]
    int off_c23;    // double2 [n]: (c2, c3)                         | bucket-centred: [n_keys + 1]
    double r_min, r_max;
    int off_knot;   // bucket-centred: uint16 [n_keys]
    double inv_h16; // bucket-centred: 65536 / h
    double h;
};
struct FastPP2 {
    ULutDesc lutx, luty;
    int off_gxpair, off_gypair;
    int off_cells;      // staged block: record (ix, iy) at off_cells + ix * row_stride + iy * kCellRecord
    int n_stage;        // cells with ix < n_stage and iy < n_stage are staged
    int row_stride;     // bytes
    int ny;             // cells per row of the global array
    const double *cells_global;  // [nx][ny][16]
};
struct FastTable {
    FastPP2 xy;
    FastPP1 lr;
    int use_lr;
    int n_bytes;  // size of the staged block
};

/// Where the staged table block lives: shared memory addressed by explicit 32-bit window
/// addresses (LDS with register + immediate operands, no generic-address arithmetic), or
/// global memory for the kernels that evaluate too little to stage 160 KB.
struct SharedTab {
    uint32_t base;
    __device__ __forceinline__ explicit SharedTab(const void *p) : base((uint32_t)__cvta_generic_to_shared(p)) {}
    __device__ __forceinline__ double2 LdV2(int off) const {
        double2 v;
        asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(base + (uint32_t)off));
        return v;
    }
    __device__ __forceinline__ int LdU16(int off) const {
        uint32_t v;
        asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(base + (uint32_t)off));
        return (int)v;
    }
    __device__ __forceinline__ uint32_t LdU32(int off) const {
        uint32_t v;
        asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + (uint32_t)off));
        return v;
    }
    __device__ __forceinline__ double LdF64(int off) const {
        double v;
        asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + (uint32_t)off));
        return v;
    }
};
struct GlobalTab {
    const unsigned char *base;
    __device__ __forceinline__ explicit GlobalTab(const void *p) : base(reinterpret_cast<const unsigned char *>(p)) {}
    __device__ __forceinline__ double2 LdV2(int off) const { return __ldg(reinterpret_cast<const double2 *>(base + off)); }
    __device__ __forceinline__ int LdU16(int off) const { return __ldg(reinterpret_cast<const unsigned short *>(base + off)); }
    __device__ __forceinline__ uint32_t LdU32(int off) const { return __ldg(reinterpret_cast<const uint32_t *>(base + off)); }
    __device__ __forceinline__ double LdF64(int off) const { return __ldg(reinterpret_cast<const double *>(base + off)); }
};

/// Stages `n_bytes` (a multiple of 16; source and destination 16-byte aligned) of table data from
/// global into shared memory with the TMA engine: one thread arms an mbarrier with the byte count
/// and issues 1-D bulk copies (cp.async.bulk, SASS UBLKCP), every thread waits on the barrier's
/// phase 0.  No registers or LSU instructions are spent on the copy and the whole block is in
/// flight at once (the thread-strided int4 loop it replaces made ~7 dependent L2 round trips).
/// Ends with the block synchronised and the data visible to ordinary shared-memory loads.
__device__ __forceinline__ void StageBlockTma(void *smem_dst, const void *gsrc, int n_bytes, unsigned long long *mbar) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)n_bytes) : "memory");
        const char *src = reinterpret_cast<const char *>(gsrc);
        constexpr int kChunk = 32768;
        for (int off = 0; off < n_bytes; off += kChunk) {
            const uint32_t sz = (uint32_t)min(kChunk, n_bytes - off);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + (uint32_t)off),
                         "l"(src + off), "r"(sz), "r"(bar)
                         : "memory");
        }
    }
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "STAGE_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra STAGE_DONE;\n"
        "bra STAGE_WAIT;\n"
        "STAGE_DONE:\n"
        "}\n" ::"r"(bar)
        : "memory");
}

/// Interval of x >= 0 and tau = x - g[interval].  `clamp_key`: x may lie beyond the grid end
/// (keys above the table are clamped to its last bucket, whose interval is the last one).
template <class Tab, int KIND = 0>
__device__ __forceinline__ void ULookup(const Tab &tb, const ULutDesc &L, int off_gpair, double x, bool clamp_key, int &i, double &t) {
    int key;
    if (KIND == 0) {
        key = __double2loint(fma(x, L.inv_h, kRoundMagic));
        if (clamp_key) key = min(key, L.key_max);
    } else {
        key = min(max((__double2hiint(x) >> L.shift) - L.key0, 0), L.key_max);
    }
    const int i0 = tb.LdU16(L.off_lut + 2 * key);
    const double2 g = tb.LdV2(off_gpair + 16 * i0);
    const bool up = x >= g.y;
    i = up ? i0 + 1 : i0;
    t = x - (up ? g.y : g.x);
}

/// Interval of x >= 0 (any x: beyond the grid end it is the last one) and tau = x - g[interval] from the packed interval
/// table.  Within 1/32768 of a bucket width of a knot either neighbouring interval may come out; the piecewise polynomial is
/// C2 there, so the two pieces differ by (jump of the third derivative) (h / 32768)^3 / 6 -- far below one ulp of the value.
template <class Tab>
__device__ __forceinline__ void ULookup16(const Tab &tb, const ULutDesc &L, int off_gpair, double x, int &i, double &t) {
    const double xc = x > L.x_cap ? L.x_cap : x;
    // low word of fma(x, 32768 / h, 2^52 + 2^51) = rint(32768 x / h): bucket (nearest) above bit 15, position in it below
    const int k32 = __double2loint(fma(xc, L.inv_h16, kRoundMagic)) + 0x4000;
    const uint32_t pk = tb.LdU32(L.off_lut32 + 4 * (k32 >> 15));
    i = (int)(pk >> 16) + (((uint32_t)k32 & 0x7FFFu) >= (pk & 0xFFFFu) ? 1 : 0);
    t = x - tb.LdF64(off_gpair + 16 * i);
}

/// x must already lie inside [r_min, r_max] (SetLimits).
template <class Tab, int KIND = 0>
__device__ __forceinline__ double FastPP1Eval(const Tab &tb, const FastPP1 &d, double x) {
#if PIMC_LR2
    if (KIND == 0) {
        // low word of fma(x, 65536 / h, 2^52 + 2^51) = rint(65536 x / h): bucket (nearest) in the high half, position in it below
        const int k32 = __double2loint(fma(x, d.inv_h16, kRoundMagic)) + 0x8000;
        const int key = k32 >> 16;
        const int knot = tb.LdU16(d.off_knot + 2 * key);
        const int rec = key + ((k32 & 0xFFFF) >= knot ? 1 : 0);
        const double t = fma(-(double)rec, d.h, x);
        const double2 c01 = tb.LdV2(d.off_c01 + 16 * rec);
        const double2 c23 = tb.LdV2(d.off_c23 + 16 * rec);
        return fma(fma(fma(c23.y, t, c23.x), t, c01.y), t, c01.x);
    }
#endif
    int i;
    double t;
    ULookup<Tab, KIND>(tb, d.lut, d.off_gpair, x, false, i, t);
    const double2 c01 = tb.LdV2(d.off_c01 + 16 * i);
    const double2 c23 = tb.LdV2(d.off_c23 + 16 * i);
    return fma(fma(fma(c23.y, t, c23.x), t, c01.y), t, c01.x);
}

/// u_long(x) of a fast Ilkka table; x must already lie inside [r_min, r_max] (SetLimits).
template <class Tab>
__device__ __forceinline__ double FastLrEval(const Tab &tb, const FastTable &T, double x) {
    return FastPP1Eval<Tab, 0>(tb, T.lr, x);
}

template <class Tab>
__device__ __forceinline__ double FastPP2Eval(const Tab &tb, const FastPP2 &d, double x, double y) {
    int ix, iy;
    double tx, ty;
#if PIMC_XY16
    ULookup16(tb, d.lutx, d.off_gxpair, x, ix, tx);
    ULookup16(tb, d.luty, d.off_gypair, y, iy, ty);
#else
    ULookup(tb, d.lutx, d.off_gxpair, x, true, ix, tx);
    ULookup(tb, d.luty, d.off_gypair, y, true, iy, ty);
#endif
    double2 c[8];
    if (ix < d.n_stage && iy < d.n_stage) {
        const int off = d.off_cells + ix * d.row_stride + iy * kCellRecord;
#pragma unroll
        for (int k = 0; k < 8; ++k) c[k] = tb.LdV2(off + 16 * k);
    } else {
        const double2 *p = reinterpret_cast<const double2 *>(d.cells_global) + ((size_t)ix * d.ny + iy) * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) c[k] = __ldg(p + k);
    }
    double row[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) row[m] = fma(fma(fma(c[2 * m + 1].y, ty, c[2 * m + 1].x), ty, c[2 * m].y), ty, c[2 * m].x);
    return fma(fma(fma(row[3], tx, row[2]), tx, row[1]), tx, row[0]);
}

/// sqrt(x) for x >= 0 to within 1 ulp (0 for x = 0): reciprocal-square-root seed (relative
/// error 2^-26), one coupled Newton step for sqrt and 1/(2 sqrt), one Markstein correction.
__device__ __forceinline__ double FastSqrt(double x) {
    const int hi = max(__double2hiint(x), 0x00100000);  // keeps the seed finite at x = 0 (s = 0 for beads at rest)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(__hiloint2double(hi, __double2loint(x))));
    double g = x * y;
    double h = 0.5 * y;
    const double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    const double d = fma(-g, g, x);
    return fma(d, h, g);
}

/// The three magnitudes of Path::DrDrpDrrp.  The reference's last step, PutInBox(r - r'), is
/// skipped: r' has just been moved to the image nearest to r, so every component of r - r'
/// already lies in [-L/2, L/2] and the rounding it would apply is zero (it could differ only
/// when a component sits within one rounding error of exactly L/2, where the reference's own
/// result depends on its compiler's contraction choices).
__device__ __forceinline__ void DrDrpDrrpFast(const double a0[3], const double b0[3], const double a1[3], const double b1[3],
                                              const Box &bx, double &r_mag, double &rp_mag, double &rrp_mag) {
    double r[3], rp[3], s[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        r[d] = b0[d] - a0[d];
        rp[d] = b1[d] - a1[d];
        r[d] = fma(-rint(r[d] * bx.iL), bx.L, r[d]);
        rp[d] = fma(rint((r[d] - rp[d]) * bx.iL), bx.L, rp[d]);
        s[d] = r[d] - rp[d];
    }
    r_mag = FastSqrt(fma(r[1], r[1], fma(r[2], r[2], r[0] * r[0])));
    rp_mag = FastSqrt(fma(rp[1], rp[1], fma(rp[2], rp[2], rp[0] * rp[0])));
    rrp_mag = FastSqrt(fma(s[1], s[1], fma(s[2], s[2], s[0] * s[0])));
}

__device__ __forceinline__ double Clamp(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }

/// IlkkaPairAction::CalcU / CalcdUdBeta (the table decides which).
template <class Tab>
__device__ __forceinline__ double FastIlkkaEval(const Tab &tb, const FastTable &T, double r, double r_p, double s) {
    const double q = 0.5 * (r + r_p);
    const double x = fma(0.5, s, q);
    const double y = fma(-0.5, s, q);  // >= 0 up to rounding (r + r' >= s); a tiny negative still rounds to key 0
    double u = FastPP2Eval(tb, T.xy, x, y);
    if (T.use_lr) {
        // SetLimits (pair_action_class.h:32-42); distances inside the grid range -- all of them
        // in a periodic box whose table reaches sqrt(3) L / 2 -- skip the selects
        const bool outside = r < T.lr.r_min || r > T.lr.r_max || r_p < T.lr.r_min || r_p > T.lr.r_max;
        if (__any_sync(__activemask(), outside)) {  // a real (warp-uniform, rarely taken) branch, not eight selects
            r = Clamp(r, T.lr.r_min, T.lr.r_max);
            r_p = Clamp(r_p, T.lr.r_min, T.lr.r_max);
        }
        u = fma(-0.5, FastLrEval(tb, T, r), u);
        u = fma(-0.5, FastLrEval(tb, T, r_p), u);
    }
    return u;
}

/// SetLimits for one value, as a warp-uniform rarely taken branch.
__device__ __forceinline__ double ClampRare(double x, const FastPP1 &d) {
    if (__any_sync(__activemask(), x < d.r_min || x > d.r_max)) x = Clamp(x, d.r_min, d.r_max);
    return x;
}

/// FastIlkkaEval for a converged warp whose 32 lanes hold consecutive links (b, b+1) of ONE
/// particle pair.  u_long is evaluated once per lane, for r; u_long(r') is the next lane's
/// u_long(r) -- the same bead pair one slice later -- whenever r' == r(next lane) bit for bit
/// (equal inputs, equal outputs: exact).  That fails only where the image shift between the
/// two slices differs (those lanes evaluate u_long(r') themselves: a rare branch with one or
/// two active lanes) and on lane 31, which has no next lane: it parks its r' in `ring` (one
/// slot per step) and the caller evaluates 32 parked values together (LrRingFlush) -- the
/// term -u_long(r')/2 is simply added to whichever lane's partial sum.
template <class Tab>
__device__ __forceinline__ double FastIlkkaEvalWarp(const Tab &tb, const FastTable &T, double r, double r_p, double s, int lane,
                                                    double *ring_slot) {
    const double q = 0.5 * (r + r_p);
    const double x = fma(0.5, s, q);
    const double y = fma(-0.5, s, q);
    double u = FastPP2Eval(tb, T.xy, x, y);
    if (T.use_lr) {
        const double lr_r = FastLrEval(tb, T, ClampRare(r, T.lr));
        const double r_next = __shfl_down_sync(0xffffffffu, r, 1);
        double lr_p = __shfl_down_sync(0xffffffffu, lr_r, 1);
        if (lane == 31) {
            *ring_slot = r_p;
            lr_p = 0.;
        } else if (r_p != r_next) {
            lr_p = FastLrEval(tb, T, Clamp(r_p, T.lr.r_min, T.lr.r_max));
        }
        u = fma(-0.5, lr_r, u);
        u = fma(-0.5, lr_p, u);
    }
    return u;
}

/// FastIlkkaEvalWarp for an OLD and a NEW set of distances at once (whole-path displacement: both
/// sets run over the same 32 consecutive links).  Lane 31 parks its two r' in *ring_old / *ring_new.
template <class Tab>
__device__ __forceinline__ void FastIlkkaEvalWarpBoth(const Tab &tb, const FastTable &T, double ro, double rpo, double so, double rn, double rpn,
                                                      double sn, int lane, double *ring_old, double *ring_new, double &uo, double &un) {
    const double qo = 0.5 * (ro + rpo), qn = 0.5 * (rn + rpn);
    uo = FastPP2Eval(tb, T.xy, fma(0.5, so, qo), fma(-0.5, so, qo));
    un = FastPP2Eval(tb, T.xy, fma(0.5, sn, qn), fma(-0.5, sn, qn));
    if (T.use_lr) {
        const unsigned full = 0xffffffffu;
        const double lro = FastLrEval(tb, T, ClampRare(ro, T.lr));
        const double lrn = FastLrEval(tb, T, ClampRare(rn, T.lr));
        const double ro_next = __shfl_down_sync(full, ro, 1), rn_next = __shfl_down_sync(full, rn, 1);
        double lpo = __shfl_down_sync(full, lro, 1), lpn = __shfl_down_sync(full, lrn, 1);
        if (lane == 31) {
            *ring_old = rpo;
            *ring_new = rpn;
            lpo = 0.;
            lpn = 0.;
        } else {
            if (rpo != ro_next) lpo = FastLrEval(tb, T, Clamp(rpo, T.lr.r_min, T.lr.r_max));
            if (rpn != rn_next) lpn = FastLrEval(tb, T, Clamp(rpn, T.lr.r_min, T.lr.r_max));
        }
        uo = fma(-0.5, lro, uo);
        uo = fma(-0.5, lpo, uo);
        un = fma(-0.5, lrn, un);
        un = fma(-0.5, lpn, un);
    }
}

/// -u_long/2 of the first `count` parked values of a warp's ring, one per lane.
template <class Tab>
__device__ __forceinline__ double LrRingFlush(const Tab &tb, const FastTable &T, const double *ring, int lane, int count) {
    __syncwarp();
    double v = 0.;
    if (lane < count) v = -0.5 * FastLrEval(tb, T, Clamp(ring[lane], T.lr.r_min, T.lr.r_max));
    __syncwarp();
    return v;
}

// ------------------------------------------------- window evaluation (bisection moves)
/// IlkkaPairAction::CalcU of one link of a bisection window in OLD and in NEW mode, for a
/// converged warp whose lanes are (partner, link j) with the nb links of a partner in consecutive
/// lanes.  The 2-D part is evaluated per lane and mode.  The long-range spline is needed at the
/// nb + 1 bead distances of the OLD window and at the nb - 1 moved ones of the NEW window (beads 0
/// and nb are common) = 2 nb values for 2 nb evaluations, so every lane evaluates it exactly
/// twice instead of four times: lrA = u_long(r_old(j)); lrB = u_long(r_new(j)) -- except that
/// lane j = 0, whose r is the same in both modes, spends its second evaluation on the window's
/// last bead nb.  u_long(r') of a link is the next lane's u_long(r) whenever r' equals that r
/// bit for bit (same inputs, same image: exact); otherwise -- the image shift changed between
/// the two slices -- the lane evaluates it itself (rare branch).
template <class Tab>
__device__ __forceinline__ void FastIlkkaEvalWindow(const Tab &tb, const FastTable &T, int j, int nb, int lane, double ro, double rpo,
                                                    double so, double rn, double rpn, double sn, double &uo, double &un) {
    const double qo = 0.5 * (ro + rpo), qn = 0.5 * (rn + rpn);
    uo = FastPP2Eval(tb, T.xy, fma(0.5, so, qo), fma(-0.5, so, qo));
    un = FastPP2Eval(tb, T.xy, fma(0.5, sn, qn), fma(-0.5, sn, qn));
    if (T.use_lr) {
        const unsigned full = 0xffffffffu;
        const double r_end = __shfl_sync(full, rpo, lane | (nb - 1));  // bead nb as the group's last lane sees it
        const double lrA = FastLrEval(tb, T, ClampRare(ro, T.lr));
        const double lrB = FastLrEval(tb, T, ClampRare(j == 0 ? r_end : rn, T.lr));
        const double lr_end = __shfl_sync(full, lrB, lane & ~(nb - 1));
        double nxt_o = __shfl_down_sync(full, lrA, 1), rnx_o = __shfl_down_sync(full, ro, 1);
        double nxt_n = __shfl_down_sync(full, lrB, 1), rnx_n = __shfl_down_sync(full, rn, 1);
        if (j == nb - 1) {
            nxt_o = lr_end;
            nxt_n = lr_end;
            rnx_o = r_end;
            rnx_n = r_end;
        }
        if (rpo != rnx_o) nxt_o = FastLrEval(tb, T, Clamp(rpo, T.lr.r_min, T.lr.r_max));
        if (rpn != rnx_n) nxt_n = FastLrEval(tb, T, Clamp(rpn, T.lr.r_min, T.lr.r_max));
        uo = fma(-0.5, lrA, uo);
        uo = fma(-0.5, nxt_o, uo);
        un = fma(-0.5, j == 0 ? lrA : lrB, un);
        un = fma(-0.5, nxt_n, un);
    }
}

/// Beads j and j + 1 of one clone's window ([j][3] doubles, consecutive) from shared memory.
__device__ __forceinline__ void LdsBeadPair(uint32_t addr, double b0[3], double b1[3]) {
    asm volatile("ld.shared.f64 %0, [%6];\n\tld.shared.f64 %1, [%6+8];\n\tld.shared.f64 %2, [%6+16];\n\t"
                 "ld.shared.f64 %3, [%6+24];\n\tld.shared.f64 %4, [%6+32];\n\tld.shared.f64 %5, [%6+40];"
                 : "=d"(b0[0]), "=d"(b0[1]), "=d"(b0[2]), "=d"(b1[0]), "=d"(b1[1]), "=d"(b1[2])
                 : "r"(addr));
}

// ------------------------------------------------------------------------------ K1 (fast)
#ifndef PIMC_FAST_THREADS
#define PIMC_FAST_THREADS 1024
#endif
constexpr int kFastThreads = PIMC_FAST_THREADS;
constexpr int kK1Unroll = PIMC_K1_UNROLL;
constexpr int kFastWarps = kFastThreads / 32;
constexpr int kFastQ = 32;                         // partner offsets handled per staged window
constexpr int kFastRows = kFastQ + kFastWarps - 1; // window rows (same species: sliding window)
constexpr int kFastRow = 34;                       // 33 slices of a chunk (+1 pad) per row

struct PairFastArgs {
    PathView pv;
    SpeciesView A, B;
    int same;
    FastTable T;
    const unsigned char *tables;  // the block to stage (T.n_bytes)
    int n_chunks, n_pgroups;
    int n_tsplit, t_windows;      // the partner windows of an item group are split n_tsplit ways, t_windows windows each
    double *partial;              // [C][n_chunks][n_pgroups][n_tsplit]
};

/// Persistent CTAs, one per SM.  Work item = (clone, 32-slice chunk, group of 32 particles of
/// species a, range of partner windows -- the whole partner loop unless the launch has too few
/// items to fill the GPU, e.g. one slice-sharded path): warp w owns particle p = 32 g + w, its
/// lanes own the chunk's 32 links.  Same
/// species: the warp pairs p with q = p + dd (mod N) for dd = 1..N/2 (the last offset only
/// from the lower half when N is even), so every unordered pair is visited once and every
/// warp does the same amount of work per staged window; window t holds the Q + 31 particles
/// q = 32 g + t Q + 1 ... and warp w reads row w + i at step i.  Different species: the window
/// holds Q particles of species b and every warp walks all of them.
static __global__ void __launch_bounds__(kFastThreads, 1) pair_full_fast_kernel(const PairFastArgs a) {
    extern __shared__ __align__(16) unsigned char fsm[];
    __shared__ double red[kFastWarps];
    __shared__ double ring[kFastWarps][32];  // lane 31's parked r' per warp and step (FastIlkkaEvalWarp)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *qpos = reinterpret_cast<double *>(fsm);               // [kFastRows][3][kFastRow]
    unsigned char *tb_ptr = fsm + sizeof(double) * kFastRows * 3 * kFastRow;
    const SharedTab tb(tb_ptr);
    __shared__ unsigned long long stage_bar;
    StageBlockTma(tb_ptr, a.tables, a.T.n_bytes, &stage_bar);
    const PathView &pv = a.pv;
    const int Na = a.A.N, Nb = a.B.N;
    const int half = Na / 2;
    const int n_dd = a.same ? half : Nb;  // partner steps per particle
    const int per_clone = a.n_chunks * a.n_pgroups * a.n_tsplit;
    const int n_items = pv.C * per_clone;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int c = item / per_clone;
        int rem = item - c * per_clone;
        const int ts = rem % a.n_tsplit;
        rem /= a.n_tsplit;
        const int chunk = rem / a.n_pgroups, pg = rem - chunk * a.n_pgroups;
        const int t_begin = ts * a.t_windows * kFastQ, t_end = min(n_dd, t_begin + a.t_windows * kFastQ);
        const int s0 = chunk * kChunk;
        const int p_lo = pg * kFastWarps;
        const int p = p_lo + warp;
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
        // steps this warp really takes: the offset N/2 of an even N belongs to the lower half
        int my_dd = n_dd;
        if (a.same && (Na & 1) == 0 && p >= half) my_dd = half - 1;
        if (!warp_on) my_dd = 0;
        double acc = 0.;
        const bool lane31_counts = s0 + 31 < pv.Mloc;
        for (int t0 = t_begin; t0 < t_end; t0 += kFastQ) {
            const int n_rows = a.same ? min(kFastQ, n_dd - t0) + kFastWarps - 1 : min(kFastQ, Nb - t0);
            const int q_first = a.same ? p_lo + t0 + 1 : t0;
            __syncthreads();  // the previous window has been consumed
            for (int row = warp; row < n_rows * 3; row += kFastWarps) {
                const int qq = row / 3, d = row - qq * 3;
                int q = q_first + qq;
                if (a.same) q %= Na;
                const double *src = a.B.R + PosIndex(pv, Nb, c, q, d, 0);
                const int i0 = StoreIndex(pv, s0 + lane);
                qpos[row * kFastRow + lane] = i0 >= 0 ? src[i0] : 0.;
                if (lane == 0) {
                    const int i1 = StoreIndex(pv, s0 + kChunk);
                    qpos[row * kFastRow + kChunk] = i1 >= 0 ? src[i1] : 0.;
                }
            }
            __syncthreads();
            const int n_step = min(kFastQ, my_dd - t0);
            const double *rowp = qpos + (a.same ? warp * 3 * kFastRow : 0) + lane;
#pragma unroll kK1Unroll
            for (int i = 0; i < n_step; ++i, rowp += 3 * kFastRow) {
                const double b0[3] = {rowp[0], rowp[kFastRow], rowp[2 * kFastRow]};
                const double b1[3] = {rowp[1], rowp[kFastRow + 1], rowp[2 * kFastRow + 1]};
                double r, rp, s;
                DrDrpDrrpFast(p0, b0, p1, b1, pv.box, r, rp, s);
#if PIMC_EXPERIMENT == 7
                const double u = FastIlkkaEval(tb, a.T, r, rp, s);
#else
                const double u = FastIlkkaEvalWarp(tb, a.T, r, rp, s, lane, &ring[warp][i]);
#endif
                acc += lane_on ? u : 0.;
            }
#if PIMC_EXPERIMENT != 7
            // a window holds at most kFastQ = 32 steps: one parked value per step
            if (a.T.use_lr && lane31_counts && n_step > 0) acc += LrRingFlush(tb, a.T, ring[warp], lane, n_step);
#endif
        }
        const double tot = BlockSum<kFastThreads>(acc, red);
        if (tid == 0) a.partial[item] = tot;
    }
}

// ------------------------------------------------------------------------- K1 (potential)
/// Tables of the fast Potential() kernel: v(r) and the long-range r-space part, both pp-form
/// with uniform interval tables, staged in shared memory.
struct FastVTable {
    FastPP1 v, lr;
    int use_lr, is_coulomb;
    int v_kind;  // interval table of v: 0 uniform buckets, 1 bit-pattern buckets (David's logarithmic grid)
    int n_bytes;
};

struct PotFastArgs {
    PathView pv;
    SpeciesView A, B;
    int same;
    FastVTable T;
    const unsigned char *tables;
    int n_chunks, n_pgroups;   // chunks of 32 STORED slices (the halo slice of a shard included)
    int n_tsplit, t_windows;
    double *partial;           // [C][n_chunks][n_pgroups][n_tsplit]
};

/// PairAction::Potential (pair_action_class.h:369-395) for Ilkka / Bare / David CalcV
/// (ilkka_pair_action_class.h:34-54, bare_pair_action_class.h:99-123, david_pair_action_class.h:26-39:
/// value 0 of the multi-spline, no r-space long-range part).  Potential() measures r at
/// slice b and r' at slice b + 1 with INDEPENDENT minimum images (App. A-6), so r' of link
/// (b, b+1) is exactly r of link (b+1, b+2): with g(r) = v(clamp_v(r))/2 - v_long(clamp_l(clamp_v(r)))/2
/// the sum over links of g(r) + g(r') is the sum over slices of w_s g(r_s), w_s = number of the
/// shard's links that touch slice s (2; 1 for a shard's first slice and for its halo slice).
/// One distance and two 1-D lookups per pair and slice instead of two and four.  Same work
/// decomposition as pair_full_fast_kernel (warp = particle of species a, lanes = slices,
/// partner rows staged per window of 32 offsets).
template <int KIND>
static __global__ void __launch_bounds__(kFastThreads, 1) potential_fast_kernel(const PotFastArgs a) {
    extern __shared__ __align__(16) unsigned char fsm[];
    __shared__ double red[kFastWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *qpos = reinterpret_cast<double *>(fsm);               // [kFastRows][3][kFastRow]
    unsigned char *tb_ptr = fsm + sizeof(double) * kFastRows * 3 * kFastRow;
    const SharedTab tb(tb_ptr);
    __shared__ unsigned long long stage_bar;
    StageBlockTma(tb_ptr, a.tables, a.T.n_bytes, &stage_bar);
    const PathView &pv = a.pv;
    const int Na = a.A.N, Nb = a.B.N;
    const int half = Na / 2;
    const int n_dd = a.same ? half : Nb;
    const int per_clone = a.n_chunks * a.n_pgroups * a.n_tsplit;
    const int n_items = pv.C * per_clone;
    const int n_store = pv.Mloc + (pv.sharded ? 1 : 0);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int c = item / per_clone;
        int rem = item - c * per_clone;
        const int ts = rem % a.n_tsplit;
        rem /= a.n_tsplit;
        const int chunk = rem / a.n_pgroups, pg = rem - chunk * a.n_pgroups;
        const int t_begin = ts * a.t_windows * kFastQ, t_end = min(n_dd, t_begin + a.t_windows * kFastQ);
        const int s = chunk * kChunk + lane;  // stored slice of this lane
        const int p_lo = pg * kFastWarps;
        const int p = p_lo + warp;
        const bool warp_on = p < Na;
        const bool lane_on = s < n_store;
        const double weight = !lane_on ? 0. : ((pv.sharded && (s == 0 || s == pv.Mloc)) ? 1. : 2.);
        double p0[3] = {0., 0., 0.};
        if (warp_on && lane_on) {
#pragma unroll
            for (int d = 0; d < 3; ++d) p0[d] = a.A.R[PosIndex(pv, Na, c, p, d, s)];
        }
        int my_dd = n_dd;
        if (a.same && (Na & 1) == 0 && p >= half) my_dd = half - 1;
        if (!warp_on) my_dd = 0;
        double acc = 0.;
        for (int t0 = t_begin; t0 < t_end; t0 += kFastQ) {
            const int n_rows = a.same ? min(kFastQ, n_dd - t0) + kFastWarps - 1 : min(kFastQ, Nb - t0);
            const int q_first = a.same ? p_lo + t0 + 1 : t0;
            __syncthreads();
            for (int row = warp; row < n_rows * 3; row += kFastWarps) {
                const int qq = row / 3, d = row - qq * 3;
                int q = q_first + qq;
                if (a.same) q %= Na;
                qpos[row * kFastRow + lane] = lane_on ? a.B.R[PosIndex(pv, Nb, c, q, d, s)] : 0.;
            }
            __syncthreads();
            const int n_step = min(kFastQ, my_dd - t0);
            const double *rowp = qpos + (a.same ? warp * 3 * kFastRow : 0) + lane;
            for (int i = 0; i < n_step; ++i, rowp += 3 * kFastRow) {
                double dr[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const double x = p0[d] - rowp[d * kFastRow];   // Path::Dr (path_class.h:108-112)
                    dr[d] = fma(-rint(x * pv.box.iL), pv.box.L, x);
                }
                double r = FastSqrt(fma(dr[1], dr[1], fma(dr[2], dr[2], dr[0] * dr[0])));
                r = ClampRare(r, a.T.v);
                double g = a.T.is_coulomb ? 0.5 / r : 0.5 * FastPP1Eval<SharedTab, KIND>(tb, a.T.v, r);
                if (a.T.use_lr) g = fma(-0.5, FastPP1Eval(tb, a.T.lr, ClampRare(r, a.T.lr)), g);
                acc = fma(weight, g, acc);
            }
        }
        const double tot = BlockSum<kFastThreads>(acc, red);
        if (tid == 0) a.partial[item] = tot;
    }
}

// ------------------------------------------------------------------------- K1 (Bare U, dU)
struct BareFastArgs {
    PathView pv;
    SpeciesView A, B;
    int same;
    FastVTable T;
    const unsigned char *tables;
    double scale;              // tau for CalcU at level 0 (bare_pair_action_class.h:146-150), 1 for CalcdUdBeta
    int n_chunks, n_pgroups, n_tsplit, t_windows;
    double *partial;           // [C][n_chunks][n_pgroups][n_tsplit]
};

/// One distance's half of BarePairAction::CalcV (bare_pair_action_class.h:99-123):
/// v(clamp_v r) / 2 - v_long(clamp_l clamp_v r) / 2.
template <class Tab>
__device__ __forceinline__ double BareHalfV(const Tab &tb, const FastVTable &T, double r) {
    r = ClampRare(r, T.v);
    double g = T.is_coulomb ? 0.5 / r : 0.5 * FastPP1Eval(tb, T.v, r);
    if (T.use_lr) g = fma(-0.5, FastPP1Eval(tb, T.lr, ClampRare(r, T.lr)), g);
    return g;
}

/// BarePairAction::CalcU / CalcdUdBeta over all pairs and links (pair_action_class.h:241-264,267-302
/// with bare_pair_action_class.h:146-150,175-177): both are CalcV(r, r') of Path::DrDrpDrrp's
/// distances -- r' in the image of r, unlike Potential().  Same decomposition as
/// pair_full_fast_kernel; the half g(r') of a link is the next lane's g(r) whenever r' equals that
/// r bit for bit (same image shift, the usual case), lane 31 parks its r' in the ring.
static __global__ void __launch_bounds__(kFastThreads, 1) bare_full_fast_kernel(const BareFastArgs a) {
    extern __shared__ __align__(16) unsigned char fsm[];
    __shared__ double red[kFastWarps];
    __shared__ double ring[kFastWarps][32];
    __shared__ unsigned long long stage_bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *qpos = reinterpret_cast<double *>(fsm);               // [kFastRows][3][kFastRow]
    unsigned char *tb_ptr = fsm + sizeof(double) * kFastRows * 3 * kFastRow;
    const SharedTab tb(tb_ptr);
    StageBlockTma(tb_ptr, a.tables, a.T.n_bytes, &stage_bar);
    const PathView &pv = a.pv;
    const int Na = a.A.N, Nb = a.B.N;
    const int half = Na / 2;
    const int n_dd = a.same ? half : Nb;
    const int per_clone = a.n_chunks * a.n_pgroups * a.n_tsplit;
    const int n_items = pv.C * per_clone;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int c = item / per_clone;
        int rem = item - c * per_clone;
        const int ts = rem % a.n_tsplit;
        rem /= a.n_tsplit;
        const int chunk = rem / a.n_pgroups, pg = rem - chunk * a.n_pgroups;
        const int t_begin = ts * a.t_windows * kFastQ, t_end = min(n_dd, t_begin + a.t_windows * kFastQ);
        const int s0 = chunk * kChunk;
        const int p_lo = pg * kFastWarps;
        const int p = p_lo + warp;
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
        int my_dd = n_dd;
        if (a.same && (Na & 1) == 0 && p >= half) my_dd = half - 1;
        if (!warp_on) my_dd = 0;
        double acc = 0.;
        const bool lane31_counts = s0 + 31 < pv.Mloc;
        for (int t0 = t_begin; t0 < t_end; t0 += kFastQ) {
            const int n_rows = a.same ? min(kFastQ, n_dd - t0) + kFastWarps - 1 : min(kFastQ, Nb - t0);
            const int q_first = a.same ? p_lo + t0 + 1 : t0;
            __syncthreads();
            for (int row = warp; row < n_rows * 3; row += kFastWarps) {
                const int qq = row / 3, d = row - qq * 3;
                int q = q_first + qq;
                if (a.same) q %= Na;
                const double *src = a.B.R + PosIndex(pv, Nb, c, q, d, 0);
                const int i0 = StoreIndex(pv, s0 + lane);
                qpos[row * kFastRow + lane] = i0 >= 0 ? src[i0] : 0.;
                if (lane == 0) {
                    const int i1 = StoreIndex(pv, s0 + kChunk);
                    qpos[row * kFastRow + kChunk] = i1 >= 0 ? src[i1] : 0.;
                }
            }
            __syncthreads();
            const int n_step = min(kFastQ, my_dd - t0);
            const double *rowp = qpos + (a.same ? warp * 3 * kFastRow : 0) + lane;
            for (int i = 0; i < n_step; ++i, rowp += 3 * kFastRow) {
                // r and r' of Path::DrDrpDrrp (the third magnitude is not used by CalcV)
                double r2 = 0., rp2 = 0.;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    double r = rowp[d * kFastRow] - p0[d];
                    double rp = rowp[d * kFastRow + 1] - p1[d];
                    r = fma(-rint(r * pv.box.iL), pv.box.L, r);
                    rp = fma(rint((r - rp) * pv.box.iL), pv.box.L, rp);
                    r2 = d == 0 ? r * r : fma(r, r, r2);
                    rp2 = d == 0 ? rp * rp : fma(rp, rp, rp2);
                }
                const double r_mag = FastSqrt(r2), rp_mag = FastSqrt(rp2);
                const double g_r = BareHalfV(tb, a.T, r_mag);
                const double r_next = __shfl_down_sync(0xffffffffu, r_mag, 1);
                double g_p = __shfl_down_sync(0xffffffffu, g_r, 1);
                if (lane == 31) {
                    ring[warp][i] = rp_mag;
                    g_p = 0.;
                } else if (rp_mag != r_next) {
                    g_p = BareHalfV(tb, a.T, rp_mag);
                }
                acc += lane_on ? g_r + g_p : 0.;
            }
            if (lane31_counts && n_step > 0) {  // the parked r' of lane 31, one per step
                __syncwarp();
                if (lane < n_step) acc += BareHalfV(tb, a.T, ring[warp][lane]);
                __syncwarp();
            }
        }
        const double tot = BlockSum<kFastThreads>(acc, red);
        if (tid == 0) a.partial[item] = a.scale * tot;
    }
}

// ------------------------------------------------------------------------- K1 (David U, dU)
// FastLut / FastDavidTable (the David tables' shared-memory layout) are declared in device_math.cuh.
template <int KIND, class Tab>
__device__ __forceinline__ void DLookup(const Tab &tb, const FastLut &L, int off_gpair, double x, int &i, double &t) {
    int key;
    if (KIND == 0) {
        key = min(__double2loint(fma(x, L.inv_h, kRoundMagic)), L.key_max);
    } else {
        key = min(max((__double2hiint(x) >> L.shift) - L.key0, 0), L.key_max);
    }
    const int i0 = tb.LdU16(L.off_lut + 2 * key);
    const double2 g = tb.LdV2(off_gpair + 16 * i0);
    const bool up = x >= g.y;
    i = up ? i0 + 1 : i0;
    t = x - (up ? g.y : g.x);
}

/// e(x) for x already inside [r_min, r_max].
template <int KIND, class Tab>
__device__ __forceinline__ double DavidEndpoint(const Tab &tb, const FastDavidTable &T, double x) {
    int i;
    double t;
    DLookup<KIND>(tb, T.lut, T.off_gpair, x, i, t);
    const double2 c01 = tb.LdV2(T.off_e01 + 16 * i);
    const double2 c23 = tb.LdV2(T.off_e23 + 16 * i);
    return fma(fma(fma(c23.y, t, c23.x), t, c01.y), t, c01.x);
}

/// The off-diagonal sum of DavidPairAction::CalcU / CalcdUdBeta (david...:80-99, 146-165):
/// sum_{k=1..n_order} sum_{j=0..k} u_kj(q) z^(2j) s^(2(k-j)), terms in the reference's order.  The
/// reference walks the powers of s down by multiplying with 1/s^2; here s^(2m) and z^(2j) are
/// built upward by multiplication (no division), which differs by rounding only.
template <int NORD, int KIND, class Tab>
__device__ __forceinline__ double DavidOffDiagonal(const Tab &tb, const FastDavidTable &T, double q, double z, double s) {
    int i;
    double t;
    DLookup<KIND>(tb, T.lut, T.off_gpair, q, i, t);
    const int base = T.off_q + i * T.q_stride;
    double sp[NORD + 1], zp[NORD + 1];
    sp[0] = 1.;
    zp[0] = 1.;
    const double s_2 = s * s, z_2 = z * z;
#pragma unroll
    for (int m = 1; m <= NORD; ++m) {
        sp[m] = sp[m - 1] * s_2;
        zp[m] = zp[m - 1] * z_2;
    }
    double u = 0.;
    int v = 0;
#pragma unroll
    for (int k = 1; k <= NORD; ++k) {
#pragma unroll
        for (int j = 0; j <= k; ++j, ++v) {
            const double2 c01 = tb.LdV2(base + 32 * v);
            const double2 c23 = tb.LdV2(base + 32 * v + 16);
            const double cof = fma(fma(fma(c23.y, t, c23.x), t, c01.y), t, c01.x);
            u = fma(cof, zp[j] * sp[k - j], u);
        }
    }
    return u;
}

/// DavidPairAction::CalcU / CalcdUdBeta for one (r, r', s) triple (test hook and ring flush).
template <int NORD, int KIND, class Tab>
__device__ __forceinline__ double FastDavidEval(const Tab &tb, const FastDavidTable &T, double r, double r_p, double s) {
    const double q = 0.5 * (r + r_p), z = r - r_p;  // before the clamp (david...:65-70)
    double u = 0.5 * DavidEndpoint<KIND>(tb, T, Clamp(r, T.r_min, T.r_max));
    u = fma(0.5, DavidEndpoint<KIND>(tb, T, Clamp(r_p, T.r_min, T.r_max)), u);
    if (s > 0.0 && q < T.r_max) u += DavidOffDiagonal<NORD, KIND>(tb, T, q, z, s);
    return u;
}

struct DavidFastArgs {
    PathView pv;
    SpeciesView A, B;
    int same;
    FastDavidTable T;
    const unsigned char *tables;
    int n_chunks, n_pgroups, n_tsplit, t_windows;
    double *partial;  // [C][n_chunks][n_pgroups][n_tsplit]
};

/// DavidPairAction::CalcU / CalcdUdBeta over all pairs and links (pair_action_class.h:241-264,
/// 267-302 with david_pair_action_class.h:63-102,126-168).  Decomposition of
/// pair_full_fast_kernel; the endpoint half e(r')/2 of a link is the next lane's e(r)/2 whenever r'
/// equals that r bit for bit (same image shift between the two slices, the usual case), lane 31
/// parks its r' in the ring and the warp evaluates the parked values together.
template <int NORD, int KIND>
static __global__ void __launch_bounds__(kFastThreads, 1) david_full_fast_kernel(const DavidFastArgs a) {
    extern __shared__ __align__(16) unsigned char fsm[];
    __shared__ double red[kFastWarps];
    __shared__ double ring[kFastWarps][32];
    __shared__ unsigned long long stage_bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *qpos = reinterpret_cast<double *>(fsm);               // [kFastRows][3][kFastRow]
    unsigned char *tb_ptr = fsm + sizeof(double) * kFastRows * 3 * kFastRow;
    const SharedTab tb(tb_ptr);
    StageBlockTma(tb_ptr, a.tables, a.T.n_bytes, &stage_bar);
    const PathView &pv = a.pv;
    const int Na = a.A.N, Nb = a.B.N;
    const int half = Na / 2;
    const int n_dd = a.same ? half : Nb;
    const int per_clone = a.n_chunks * a.n_pgroups * a.n_tsplit;
    const int n_items = pv.C * per_clone;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int c = item / per_clone;
        int rem = item - c * per_clone;
        const int ts = rem % a.n_tsplit;
        rem /= a.n_tsplit;
        const int chunk = rem / a.n_pgroups, pg = rem - chunk * a.n_pgroups;
        const int t_begin = ts * a.t_windows * kFastQ, t_end = min(n_dd, t_begin + a.t_windows * kFastQ);
        const int s0 = chunk * kChunk;
        const int p_lo = pg * kFastWarps;
        const int p = p_lo + warp;
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
        int my_dd = n_dd;
        if (a.same && (Na & 1) == 0 && p >= half) my_dd = half - 1;
        if (!warp_on) my_dd = 0;
        double acc = 0.;
        const bool lane31_counts = s0 + 31 < pv.Mloc;
        for (int t0 = t_begin; t0 < t_end; t0 += kFastQ) {
            const int n_rows = a.same ? min(kFastQ, n_dd - t0) + kFastWarps - 1 : min(kFastQ, Nb - t0);
            const int q_first = a.same ? p_lo + t0 + 1 : t0;
            __syncthreads();
            for (int row = warp; row < n_rows * 3; row += kFastWarps) {
                const int qq = row / 3, d = row - qq * 3;
                int q = q_first + qq;
                if (a.same) q %= Na;
                const double *src = a.B.R + PosIndex(pv, Nb, c, q, d, 0);
                const int i0 = StoreIndex(pv, s0 + lane);
                qpos[row * kFastRow + lane] = i0 >= 0 ? src[i0] : 0.;
                if (lane == 0) {
                    const int i1 = StoreIndex(pv, s0 + kChunk);
                    qpos[row * kFastRow + kChunk] = i1 >= 0 ? src[i1] : 0.;
                }
            }
            __syncthreads();
            const int n_step = min(kFastQ, my_dd - t0);
            const double *rowp = qpos + (a.same ? warp * 3 * kFastRow : 0) + lane;
            for (int i = 0; i < n_step; ++i, rowp += 3 * kFastRow) {
                const double b0[3] = {rowp[0], rowp[kFastRow], rowp[2 * kFastRow]};
                const double b1[3] = {rowp[1], rowp[kFastRow + 1], rowp[2 * kFastRow + 1]};
                double r, rp, s;
                DrDrpDrrpFast(p0, b0, p1, b1, pv.box, r, rp, s);
                const double q = 0.5 * (r + rp), z = r - rp;  // before the clamp (david...:65-70)
                double rc = r;
                if (__any_sync(0xffffffffu, r < a.T.r_min || r > a.T.r_max)) rc = Clamp(r, a.T.r_min, a.T.r_max);
                const double e_r = DavidEndpoint<KIND>(tb, a.T, rc);
                const double r_next = __shfl_down_sync(0xffffffffu, r, 1);
                double e_p = __shfl_down_sync(0xffffffffu, e_r, 1);
                if (lane == 31) {
                    ring[warp][i] = rp;
                    e_p = 0.;
                } else if (rp != r_next) {
                    e_p = DavidEndpoint<KIND>(tb, a.T, Clamp(rp, a.T.r_min, a.T.r_max));
                }
                double u = 0.5 * (e_r + e_p);
                if (s > 0.0 && q < a.T.r_max) u += DavidOffDiagonal<NORD, KIND>(tb, a.T, q, z, s);
                acc += lane_on ? u : 0.;
            }
            if (lane31_counts && n_step > 0) {  // the parked r' of lane 31, one per step
                __syncwarp();
                if (lane < n_step) acc = fma(0.5, DavidEndpoint<KIND>(tb, a.T, Clamp(ring[warp][lane], a.T.r_min, a.T.r_max)), acc);
                __syncwarp();
            }
        }
        const double tot = BlockSum<kFastThreads>(acc, red);
        if (tid == 0) a.partial[item] = tot;
    }
}

/// Test hooks: the fast evaluation on caller-supplied triples (tables read from global memory
/// through the same code) and the square root on its own.
static __global__ void calc_pair_fast_kernel(const unsigned char *__restrict__ tables, FastTable T, int n, const double *__restrict__ r,
                                      const double *__restrict__ rp, const double *__restrict__ s, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = FastIlkkaEval(GlobalTab(tables), T, r[i], rp[i], s[i]);
}
template <int NORD, int KIND>
static __global__ void calc_david_fast_kernel(const unsigned char *__restrict__ tables, FastDavidTable T, int n, const double *__restrict__ r,
                                       const double *__restrict__ rp, const double *__restrict__ s, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = FastDavidEval<NORD, KIND>(GlobalTab(tables), T, r[i], rp[i], s[i]);
}
static __global__ void fast_sqrt_kernel(int n, const double *__restrict__ x, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = FastSqrt(x[i]);
}

}  // namespace pimc

#endif  // SIMPIMC_B200_PAIR_FAST_CUH_
