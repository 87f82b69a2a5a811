// C ABI of the B200-native action-evaluation path: implementation of include/simpimc_b200.h.
//
// Host side of the boundary: owns the device buffers of a context (n_clones walkers), builds
// the spline tables and k-space lists the way the reference's constructors do, and launches
// the kernels in kernels.cuh on the context's stream.  There is no CPU fallback: every
// compute entry point fails with PIMC_ERR_CUDA when no device is usable.
#include "state.h"
#include "mc.cuh"
#include "grad.cuh"
#include "spline_build.h"

using namespace pimc;

namespace {
}  // namespace

namespace pimc_host {

thread_local std::string g_last_error;

int Fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

// -------------------------------------------------------------------------------- helpers

ScopedKernelTimer::ScopedKernelTimer(pimc_ctx *c, int kernel_id) : ctx(c), id(kernel_id) {
    if (!ctx->timing) return;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
        e0 = e1 = nullptr;
        return;
    }
    cudaEventRecord(e0, ctx->stream);
}
ScopedKernelTimer::~ScopedKernelTimer() {
    if (!e0) return;
    cudaEventRecord(e1, ctx->stream);
    ctx->timers[id].pending.push_back(std::make_pair(e0, e1));
}

int EnsureI32(pimc_ctx *ctx, DevBuf<int32_t> &buf, const int32_t *host, size_t n) {
    if (buf.n < n) PIMC_CUDA(buf.Alloc(n));
    PIMC_CUDA(cudaMemcpyAsync(buf.p, host, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    return PIMC_OK;
}

int GridFor(const pimc_ctx *ctx, int items) { return std::max(1, std::min(items, ctx->n_sm * 8)); }

// KSpace::Setup (k_space_class.h:33-80).  The reference walks GenCombPermK's list
// (algorithm.h:101-121): sorted index multisets in lexicographic order, each followed by its
// distinct permutations in lexicographic order.  Equivalent total order: sort all lattice
// triples by (sorted triple, triple).
void BuildKSpace(pimc_ctx *ctx, double k_cut) {
    ctx->k_index.clear();
    ctx->k_mag.clear();
    ctx->k_cutoff = k_cut;
    const double kb = 2. * M_PI / ctx->L;
    const int m = (int)(uint32_t)std::ceil(1.1 * k_cut / kb);
    ctx->max_index = m;
    struct Cand {
        int s[3], v[3];
    };
    std::vector<Cand> cands;
    for (int i = -m; i <= m; ++i)
        for (int j = -m; j <= m; ++j)
            for (int k = -m; k <= m; ++k) {
                Cand c;
                c.v[0] = i;
                c.v[1] = j;
                c.v[2] = k;
                c.s[0] = i;
                c.s[1] = j;
                c.s[2] = k;
                std::sort(c.s, c.s + 3);
                cands.push_back(c);
            }
    std::sort(cands.begin(), cands.end(), [](const Cand &a, const Cand &b) {
        for (int d = 0; d < 3; ++d)
            if (a.s[d] != b.s[d]) return a.s[d] < b.s[d];
        for (int d = 0; d < 3; ++d)
            if (a.v[d] != b.v[d]) return a.v[d] < b.v[d];
        return false;
    });
    for (const Cand &c : cands) {
        const double k0 = c.v[0] * kb, k1 = c.v[1] * kb, k2 = c.v[2] * kb;
        // dot(k,k) in the reference's pairing (x0^2 + x2^2) + x1^2; no contraction on the host
        const double kk = (k0 * k0 + k2 * k2) + k1 * k1;
        if (!(kk < k_cut * k_cut && kk != 0.)) continue;
        const bool keep = (k0 > 0.) || (k0 == 0. && k1 > 0.) || (k0 == 0. && k1 == 0. && k2 > 0.);
        if (!keep) continue;
        ctx->k_index.push_back(c.v[0]);
        ctx->k_index.push_back(c.v[1]);
        ctx->k_index.push_back(c.v[2]);
        ctx->k_mag.push_back(std::sqrt(kk));
    }
}

int UploadKSpace(pimc_ctx *ctx) {
    const int n_k = ctx->n_k();
    std::vector<int32_t> off(ctx->k_index);
    for (auto &v : off) v += ctx->max_index;
    PIMC_CUDA(ctx->d_kidx.Alloc(std::max(1, n_k * 3)));
    PIMC_CUDA(ctx->d_kmag.Alloc(std::max(1, n_k)));
    if (n_k) {
        PIMC_CUDA(cudaMemcpyAsync(ctx->d_kidx.p, off.data(), off.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        PIMC_CUDA(cudaMemcpyAsync(ctx->d_kmag.p, ctx->k_mag.data(), n_k * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    // columns (i_x, i_y) in order of first appearance; kmap[col][i_z + TM] = k
    int tm = 0;
    for (int32_t v : ctx->k_index) tm = std::max(tm, std::abs(v));
    std::vector<int32_t> col_info, col_key;
    for (int k = 0; k < n_k; ++k) {
        const int32_t key = (ctx->k_index[3 * k] + 128) * 256 + ctx->k_index[3 * k + 1] + 128;
        if (std::find(col_key.begin(), col_key.end(), key) == col_key.end()) col_key.push_back(key);
    }
    ctx->n_cols = (int)col_key.size();
    ctx->cols_tm = tm;
    std::vector<int32_t> kmap((size_t)std::max(1, ctx->n_cols) * (2 * tm + 1), -1);
    for (int k = 0; k < n_k; ++k) {
        const int32_t ix = ctx->k_index[3 * k], iy = ctx->k_index[3 * k + 1], iz = ctx->k_index[3 * k + 2];
        const int col = (int)(std::find(col_key.begin(), col_key.end(), (ix + 128) * 256 + iy + 128) - col_key.begin());
        kmap[(size_t)col * (2 * tm + 1) + iz + tm] = k;
    }
    for (int32_t key : col_key) {
        const int ix = key / 256 - 128, iy = key % 256 - 128;
        col_info.push_back(std::abs(ix) | (std::abs(iy) << 8) | ((ix < 0) << 16) | ((iy < 0) << 17));
    }
    PIMC_CUDA(ctx->d_col_info.Alloc(std::max(1, ctx->n_cols)));
    PIMC_CUDA(ctx->d_kmap.Alloc(kmap.size()));
    if (n_k) {
        PIMC_CUDA(cudaMemcpyAsync(ctx->d_col_info.p, col_info.data(), col_info.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        PIMC_CUDA(cudaMemcpyAsync(ctx->d_kmap.p, kmap.data(), kmap.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    PIMC_CUDA(cudaStreamSynchronize(ctx->stream));
    return PIMC_OK;
}

template <int TM>
int LaunchRhokCols(pimc_ctx *ctx, int s, KColsView kc, double2 *rho) {
    const int S = kColsWarps / kc.n_groups;
    const size_t smem = (size_t)S * 32 * ColsEntries(TM) * sizeof(double2);
    PIMC_CUDA(cudaFuncSetAttribute(rhok_build_cols_kernel<TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int base_grid = ctx->C * ((ctx->Mloc + S - 1) / S);
    // few (clone, slice) items and many particles: split the particle loop (partial sums, then
    // a fixed-order reduction) until there are about two CTAs per SM
    const int N = ctx->species[s]->N;
    const int n_chunks32 = (N + 31) / 32;
    int p_split = std::max(1, std::min(n_chunks32 / 4, (2 * ctx->n_sm + base_grid - 1) / base_grid));
    kc.p_chunk = 32 * ((n_chunks32 + p_split - 1) / p_split);
    kc.p_split = p_split = (N + kc.p_chunk - 1) / kc.p_chunk;
    const size_t n = (size_t)ctx->C * ctx->Mloc * kc.n_k;
    double2 *dst = rho;
    if (p_split > 1) {
        if (ctx->rho_part.n < n * p_split) PIMC_CUDA(ctx->rho_part.Alloc(n * p_split));
        dst = ctx->rho_part.p;
    }
    {
        ScopedKernelTimer t(ctx, PIMC_KERNEL_RHOK_BUILD);
        rhok_build_cols_kernel<TM><<<base_grid * p_split, kColsWarps * 32, smem, ctx->stream>>>(ctx->View(), ctx->SView(s, false), kc, dst);
        if (p_split > 1) rhok_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(dst, p_split, n, rho);
    }
    ctx->launches += p_split > 1 ? 2 : 1;
    PIMC_CUDA(cudaGetLastError());
    return PIMC_OK;
}

int RebuildRhoK(pimc_ctx *ctx, int s) {
    SpeciesState &st = *ctx->species[s];
    const int n_k = ctx->n_k();
    st.drho_valid = false;
    st.need_update_rho_k = true;
    if (n_k == 0) {
        st.rho.Free();
        return PIMC_OK;
    }
    const size_t need = (size_t)ctx->C * ctx->Mloc * n_k;
    if (st.rho.n != need) PIMC_CUDA(st.rho.Alloc(need));
    // column-factorised build (every KSpace::Setup set with |index| <= 6 and <= 256 columns)
    if (!ctx->force_general && ctx->cols_tm >= 1 && ctx->cols_tm <= 6 && ctx->n_cols <= 32 * kColsWarps) {
        KColsView kc;
        kc.n_k = n_k;
        kc.n_cols = ctx->n_cols;
        kc.n_groups = (ctx->n_cols + 31) / 32;
        kc.col_info = ctx->d_col_info.p;
        kc.kmap = ctx->d_kmap.p;
        kc.kbox = 2. * M_PI / ctx->L;
        kc.stage_out = n_k <= 32 * ColsEntries(ctx->cols_tm) ? 1 : 0;
        switch (ctx->cols_tm) {
            case 1: return LaunchRhokCols<1>(ctx, s, kc, st.rho.p);
            case 2: return LaunchRhokCols<2>(ctx, s, kc, st.rho.p);
            case 3: return LaunchRhokCols<3>(ctx, s, kc, st.rho.p);
            case 4: return LaunchRhokCols<4>(ctx, s, kc, st.rho.p);
            case 5: return LaunchRhokCols<5>(ctx, s, kc, st.rho.p);
            default: return LaunchRhokCols<6>(ctx, s, kc, st.rho.p);
        }
    }
    const int tl = 2 * ctx->max_index + 1;
    int chunk = std::max(1, std::min(32, (int)(40000 / (3 * tl * sizeof(double2)))));
    const size_t smem = (size_t)chunk * 3 * tl * sizeof(double2);
    const int items = ctx->C * ctx->Mloc;
    {
        ScopedKernelTimer t(ctx, PIMC_KERNEL_RHOK_BUILD);
        rhok_build_kernel<<<GridFor(ctx, items), 256, smem, ctx->stream>>>(ctx->View(), ctx->SView(s, false), ctx->KView(), chunk, st.rho.p);
    }
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return PIMC_OK;
}

// --------------------------------------------------------------------------- table packing
void PadEven(std::vector<double> &blob) {
    if (blob.size() & 1) blob.push_back(0.0);
}

void AppendLut(std::vector<double> &blob, LutDesc &d, const double *grid, int n) {
    PadEven(blob);
    BitLut L = BuildBitLut(grid, n);
    d.off = (int)(blob.size() * 4);
    d.n_keys = (int)L.lut.size();
    d.shift = L.shift;
    d.key0 = L.key0;
    const size_t n_dbl = (L.lut.size() + 3) / 4;
    const size_t base = blob.size();
    blob.resize(base + n_dbl, 0.0);
    std::memcpy(reinterpret_cast<char *>(blob.data() + base), L.lut.data(), L.lut.size() * sizeof(uint16_t));
}

/// Appends [grid | pp | lut] of the natural spline through (grid, data) and describes it.
void AppendPP1(std::vector<double> &blob, PP1Desc &d, const double *grid, const double *data, int n) {
    KnotBasis kb;
    kb.Build(grid, n);
    std::vector<double> coefs(n + 3, 0.0);
    SolveNatural(kb, data, 1, coefs.data(), 1);
    std::vector<double> pp = PPFrom1D(kb, coefs.data());
    PadEven(blob);
    d.n = n;
    d.off_g = (int)blob.size();
    blob.insert(blob.end(), grid, grid + n);
    PadEven(blob);
    d.off_pp = (int)blob.size();
    blob.insert(blob.end(), pp.begin(), pp.end());
    AppendLut(blob, d.lut, grid, n);
    d.r_min = grid[0];
    d.r_max = grid[n - 1];
}

void Fill1D(Sp1Desc &d, int n, int base, const double *grid) {
    d.n = n;
    d.off_t = base;
    d.off_w = base + (n + 5);
    d.off_c = base + (n + 5) + 3 * (n + 2);
    d.n_splines = 1;
    d.code = GRIDCODE_GENERAL;
    d.r_min = grid[0];
    d.r_max = grid[n - 1];
    d.ainv = 0.;
    d.startinv = 0.;
}

void CheckTable1D(const pimc_table_1d &t, const char *what) {
    if (t.n < 4 || !t.r || !t.f) throw std::invalid_argument(std::string(what) + ": missing or too short");
}

// ilkka_pair_action_class.h:308-315: weight of each k vector = table value of the shell whose
// |k| matches within 1e-8 (last match wins), else 0.
std::vector<double> MatchShells(const pimc_ctx *ctx, const std::vector<double> &k, const std::vector<double> &f_k) {
    std::vector<double> w(ctx->n_k(), 0.);
    for (int k_i = 0; k_i < ctx->n_k(); ++k_i)
        for (size_t k_t = 0; k_t < k.size(); ++k_t)
            if (std::fabs(ctx->k_mag[k_i] - k[k_t]) < 1.e-8) w[k_i] = f_k[k_t];
    return w;
}

// ilkka...:421-431, bare...:89-95, david...:347-358
void ScaleConstants(const pimc_ctx *ctx, const pimc_action *a, double &k0, double &r0) {
    const int Na = ctx->species[a->sa]->N, Nb = ctx->species[a->sb]->N;
    if (a->sa == a->sb) {
        k0 *= 0.5 * Na * Nb * ctx->M;
        r0 *= -0.5 * Na * ctx->M;
    } else {
        k0 *= Na * Nb * ctx->M;
        r0 *= 0.;
    }
}

int UploadBlob(pimc_ctx *ctx, DevBuf<double> &dst, const std::vector<double> &blob) {
    PIMC_CUDA(dst.Alloc(blob.size()));
    PIMC_CUDA(cudaMemcpy(dst.p, blob.data(), blob.size() * sizeof(double), cudaMemcpyHostToDevice));
    return PIMC_OK;
}
int UploadVec(DevBuf<double> &dst, const std::vector<double> &v) {
    PIMC_CUDA(dst.Alloc(std::max<size_t>(1, v.size())));
    if (!v.empty()) PIMC_CUDA(cudaMemcpy(dst.p, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice));
    return PIMC_OK;
}

// PairAction constructor (pair_action_class.h:204-238)
int NewAction(pimc_ctx *ctx, int atype, int sa, int sb, int max_level, int use_lr, double k_cut, pimc_action **out) {
    if (!ctx || !out) return Fail(PIMC_ERR_INVALID, "null context or output");
    if (sa < 0 || sb < 0 || sa >= (int)ctx->species.size() || sb >= (int)ctx->species.size())
        return Fail(PIMC_ERR_INVALID, "species index out of range");
    if (max_level != 0)
        return Fail(PIMC_ERR_UNSUPPORTED,
                    "max_level > 0: the reference's long-range update assumes level 0 (pair_action_class.h:293) and every "
                    "shipped input uses max_level=0");
    if (use_lr && !ctx->pbc) return Fail(PIMC_ERR_INVALID, "use_long_range needs a periodic box");
    pimc_action *a = new pimc_action;
    a->ctx = ctx;
    a->atype = atype;
    a->sa = sa;
    a->sb = sb;
    a->max_level = max_level;
    a->use_long_range = use_lr != 0;
    if (a->use_long_range && k_cut > ctx->k_cutoff) {
        int32_t n_k;
        int rc = pimc_kspace_setup(ctx, k_cut, &n_k);
        if (rc != PIMC_OK) {
            delete a;
            return rc;
        }
    }
    a->is_constant = (sa == sb) && (ctx->species[sa]->N == 1 || ctx->species[sa]->lambda == 0.);
    *out = a;
    return PIMC_OK;
}

int LoadLongRange(pimc_ctx *ctx, pimc_action *a, int which, const pimc_long_range &lr, std::vector<double> &blob, PP1Desc &desc) {
    CheckTable1D(lr.f_r, "long-range r table");
    if (lr.n_k < 1 || !lr.k || !lr.f_k) throw std::invalid_argument("long-range k table missing");
    AppendPP1(blob, desc, lr.f_r.r, lr.f_r.f, lr.f_r.n);
    a->shell_k[which].assign(lr.k, lr.k + lr.n_k);
    a->shell_f[which].assign(lr.f_k, lr.f_k + lr.n_k);
    int rc = UploadVec(a->wk[which], MatchShells(ctx, a->shell_k[which], a->shell_f[which]));
    if (rc != PIMC_OK) return rc;
    a->k0[which] = lr.f_k_0;
    a->r0[which] = lr.f_r_0;
    return PIMC_OK;
}

/// Re-matches every long-range action's k-vector weights after KSpace::Setup rebuilt the list.
int RematchWeights(pimc_ctx *ctx) {
    for (pimc_action *a : ctx->actions) {
        if (!a->use_long_range) continue;
        for (int which = 0; which < 3; ++which) {
            if (a->shell_k[which].empty()) continue;
            int rc = UploadVec(a->wk[which], MatchShells(ctx, a->shell_k[which], a->shell_f[which]));
            if (rc != PIMC_OK) return rc;
        }
    }
    return PIMC_OK;
}

/// The FreeSplines of (species, n_images), built and uploaded on first use.
int GetFreeSet(pimc_ctx *ctx, int s, int n_images, FreeSet **out) {
    SpeciesState &st = *ctx->species[s];
    auto it = st.free_sets.find(n_images);
    if (it != st.free_sets.end()) {
        *out = it->second.get();
        return PIMC_OK;
    }
    if (n_images < 0) return Fail(PIMC_ERR_INVALID, "negative n_images");
    if (!(st.lambda > 0.)) return Fail(PIMC_ERR_INVALID, "free-particle splines of a species with lambda = 0");
    std::unique_ptr<FreeSet> fs(new FreeSet);
    std::memset(&fs->view, 0, sizeof(fs->view));
    std::memset(&fs->dtau, 0, sizeof(fs->dtau));
    fs->view.n_images = n_images;
    try {
        for (int k = 0; k < kMaxFreeSplines; ++k) {
            const double tau_s = 0.5 * ctx->tau * (double)(1 << k);
            FreeSplineTab &T = fs->view.s[k];
            T.i4lt = 1. / (4. * st.lambda * tau_s);
            if (n_images == 0) {  // the reference's spline of zeros: closed form, no table
                if (k == 1) {
                    fs->dtau = T;
                    fs->dtau.i4lt = 1. / (4. * st.lambda * tau_s * tau_s);
                }
                continue;
            }
            const FreeSplineHost h = BuildFreeSpline(ctx->pbc ? ctx->L : 0., (unsigned)n_images, st.lambda, tau_s, k == 1);
            int rc = UploadVec(fs->pp[k], h.pp_action);
            if (rc != PIMC_OK) return rc;
            T.pp = fs->pp[k].p;
            T.n_int = h.n - 1;
            T.zero_lo = h.zero_lo;
            T.zero_hi = h.zero_hi;
            T.start = h.start;
            T.dr = h.dr;
            T.inv_dr = 1. / h.dr;
            T.i4lt = h.i4lt;
            if (k == 1) {
                if ((rc = UploadVec(fs->pp_dtau, h.pp_dtau)) != PIMC_OK) return rc;
                fs->dtau = T;
                fs->dtau.pp = fs->pp_dtau.p;
                fs->dtau.zero_lo = h.dtau_zero_lo;
                fs->dtau.zero_hi = h.dtau_zero_hi;
                fs->dtau.i4lt = h.i4ltt;
            }
        }
    } catch (const std::exception &e) {
        return Fail(PIMC_ERR_TABLE, e.what());
    }
    *out = fs.get();
    st.free_sets[n_images] = std::move(fs);
    return PIMC_OK;
}

// ------------------------------------------------------------------- fast Ilkka tables
struct ByteBlob {
    std::vector<unsigned char> b;
    int Reserve(size_t bytes) {  // 16-byte aligned region, returns its offset
        const size_t off = (b.size() + 15) & ~(size_t)15;
        b.resize(off + bytes, 0);
        return (int)off;
    }
    template <class T>
    T *At(int off) { return reinterpret_cast<T *>(b.data() + off); }
};

int AppendULut(ByteBlob &blob, ULutDesc &d, const ULut &L) {
    d.off_lut = blob.Reserve(L.lut.size() * sizeof(uint16_t));
    std::memcpy(blob.At<uint16_t>(d.off_lut), L.lut.data(), L.lut.size() * sizeof(uint16_t));
    d.key_max = (int)L.lut.size() - 1;
    d.inv_h = L.inv_h;
    return d.off_lut;
}

#if PIMC_XY16
/// The packed interval table of ULookup16 (pair_fast.cuh) for a grid whose uniform table is L.
void AppendULut32(ByteBlob &blob, ULutDesc &d, const ULut &L, const double *g, int n) {
    const std::vector<uint32_t> pk = PackULut(L, g, n);
    d.off_lut32 = blob.Reserve(pk.size() * sizeof(uint32_t));
    std::memcpy(blob.At<uint32_t>(d.off_lut32), pk.data(), pk.size() * sizeof(uint32_t));
    d.inv_h16 = 32768.0 / L.h;
    d.x_cap = (double)((int)L.lut.size() - 1) * L.h;
}
#endif

int AppendKnotPairs(ByteBlob &blob, const double *g, int n) {
    const int off = blob.Reserve((size_t)n * 16);
    double *p = blob.At<double>(off);
    for (int i = 0; i < n; ++i) {
        p[2 * i] = g[i];
        p[2 * i + 1] = (i + 1 < n) ? g[i + 1] : HUGE_VAL;
    }
    return off;
}

/// The arrays of one 1-D pp-form spline in the fast layout (pair_fast.cuh: FastPP1).  A uniform interval table
/// (`ulut`) gives the bucket-centred form with PIMC_LR2 (records per bucket + knot positions), the interval form
/// otherwise; a bit-pattern table (`blut`) always the interval form.
bool AppendPP1Arrays(ByteBlob &blob, FastPP1 &d, const pimc_table_1d &f, const KnotBasis &kb, const double *coefs, const ULut *ulut,
                     const BLut *blut, int max_keys) {
#if PIMC_LR2
    if (ulut) {
        LR2Host l2;
        if (!BuildLR2(f.r, f.n, kb, coefs, max_keys, l2)) return false;
        d.h = l2.h;
        d.inv_h16 = l2.inv_h16;
        d.off_c01 = blob.Reserve(l2.c01.size() * sizeof(double));
        std::memcpy(blob.At<double>(d.off_c01), l2.c01.data(), l2.c01.size() * sizeof(double));
        d.off_c23 = blob.Reserve(l2.c23.size() * sizeof(double));
        std::memcpy(blob.At<double>(d.off_c23), l2.c23.data(), l2.c23.size() * sizeof(double));
        d.off_knot = blob.Reserve(l2.knot.size() * sizeof(uint16_t));
        std::memcpy(blob.At<uint16_t>(d.off_knot), l2.knot.data(), l2.knot.size() * sizeof(uint16_t));
        return true;
    }
#else
    (void)max_keys;
#endif
    const std::vector<double> pp = PPFrom1D(kb, coefs);
    d.off_gpair = AppendKnotPairs(blob, f.r, f.n);
    d.off_c01 = blob.Reserve((size_t)f.n * 16);
    d.off_c23 = blob.Reserve((size_t)f.n * 16);
    for (int i = 0; i < f.n; ++i) {
        blob.At<double>(d.off_c01)[2 * i] = pp[4 * (size_t)i];
        blob.At<double>(d.off_c01)[2 * i + 1] = pp[4 * (size_t)i + 1];
        blob.At<double>(d.off_c23)[2 * i] = pp[4 * (size_t)i + 2];
        blob.At<double>(d.off_c23)[2 * i + 1] = pp[4 * (size_t)i + 3];
    }
    if (ulut) {
        AppendULut(blob, d.lut, *ulut);
    } else {
        d.lut.off_lut = blob.Reserve(blut->lut.size() * sizeof(uint16_t));
        std::memcpy(blob.At<uint16_t>(d.lut.off_lut), blut->lut.data(), blut->lut.size() * sizeof(uint16_t));
        d.lut.key_max = (int)blut->lut.size() - 1;
        d.lut.inv_h = 0.;
        d.lut.shift = blut->shift;
        d.lut.key0 = blut->key0;
    }
    return true;
}

/// Packs the shared-memory block of the fast pair kernel for one of u_xy / du_xy (+ its
/// long-range r-space spline).  `cells` = PPFrom2D output [nx][ny][16].  Leaves ok = false when
/// a grid does not admit the uniform interval table or nothing fits.
int BuildFastIlkka(pimc_ctx *ctx, pimc_action *a, int which, const pimc_table_2d &g, const std::vector<double> &cells,
                   const pimc_long_range *lr) {
    a->fast_ok[which] = false;
    FastTable &T = a->fast[which];
    std::memset(&T, 0, sizeof(T));
    ULut lx, ly, ll;
    const int kMaxKeys = 16384;
    if (!BuildULut(g.x, g.n_x, kMaxKeys, lx) || !BuildULut(g.y, g.n_y, kMaxKeys, ly)) return PIMC_OK;
    ByteBlob blob;
    T.xy.off_gxpair = AppendKnotPairs(blob, g.x, g.n_x);
    T.xy.off_gypair = AppendKnotPairs(blob, g.y, g.n_y);
    AppendULut(blob, T.xy.lutx, lx);
    AppendULut(blob, T.xy.luty, ly);
#if PIMC_XY16
    if (lx.lut.size() > 65000 || ly.lut.size() > 65000) return PIMC_OK;  // bucket and position share 31 bits
    AppendULut32(blob, T.xy.lutx, lx, g.x, g.n_x);
    AppendULut32(blob, T.xy.luty, ly, g.y, g.n_y);
#endif
    T.xy.ny = g.n_y;
    T.xy.cells_global = a->cells[which].p;
    T.use_lr = a->use_long_range ? 1 : 0;
    if (a->use_long_range) {
        const pimc_table_1d &f = lr->f_r;
        if (!BuildULut(f.r, f.n, kMaxKeys, ll)) return PIMC_OK;
        KnotBasis kb;
        kb.Build(f.r, f.n);
        std::vector<double> coefs(f.n + 3, 0.0);
        SolveNatural(kb, f.f, 1, coefs.data(), 1);
        if (!AppendPP1Arrays(blob, T.lr, f, kb, coefs.data(), &ll, nullptr, kMaxKeys)) return PIMC_OK;
        T.lr.r_min = f.r[0];
        T.lr.r_max = f.r[f.n - 1];
    }
    // block of cells the box can reach: x = q + s/2 <= 1.5 * (sqrt(3) L / 2)
    int n_need = std::min(g.n_x, g.n_y);
    if (ctx->pbc) {
        const double x_max = 1.5 * 0.5 * std::sqrt(3.0) * ctx->L;
        const int ix = (int)(std::upper_bound(g.x, g.x + g.n_x, x_max) - g.x);
        const int iy = (int)(std::upper_bound(g.y, g.y + g.n_y, x_max) - g.y);
        n_need = std::min(n_need, std::max(ix, iy) + 1);
    }
    const size_t fixed = sizeof(double) * kFastRows * 3 * kFastRow + ((blob.b.size() + 15) & ~(size_t)15) + 2048 +
                         sizeof(double) * kFastWarps * 33;  // static shared memory of the kernel (ring, red)
    int n_stage = 0, row_stride = 0;
    for (int n = n_need; n >= 1; --n) {
        const int pad_slots = (4 - (n % 8) + 8) % 8;  // row stride = 4 (mod 8) sixteen-byte slots: rows fall into disjoint banks
        const int rs = n * kCellRecord + pad_slots * 16;
        if (fixed + (size_t)n * rs <= ctx->smem_optin) {
            n_stage = n;
            row_stride = rs;
            break;
        }
    }
    T.xy.n_stage = n_stage;
    T.xy.row_stride = row_stride;
    if (n_stage > 0) {
        T.xy.off_cells = blob.Reserve((size_t)n_stage * row_stride);
        for (int ix = 0; ix < n_stage; ++ix)
            for (int iy = 0; iy < n_stage; ++iy)
                std::memcpy(blob.b.data() + T.xy.off_cells + (size_t)ix * row_stride + (size_t)iy * kCellRecord,
                            &cells[((size_t)ix * g.n_y + iy) * 16], 128);
    }
    blob.Reserve(0);
    blob.b.resize((blob.b.size() + 15) & ~(size_t)15, 0);
    T.n_bytes = (int)blob.b.size();
    PIMC_CUDA(a->fast_tab[which].Alloc(blob.b.size()));
    PIMC_CUDA(cudaMemcpy(a->fast_tab[which].p, blob.b.data(), blob.b.size(), cudaMemcpyHostToDevice));
    a->fast_ok[which] = true;
    return PIMC_OK;
}

/// One pp-form 1-D spline in the fast shared-memory layout; false if the grid admits no uniform interval table.
bool AppendFastPP1(ByteBlob &blob, FastPP1 &d, const pimc_table_1d &f, int max_keys, int *kind = nullptr) {
    ULut lut;
    BLut blut;
    const bool uniform = BuildULut(f.r, f.n, max_keys, lut);
    // callers that pass `kind` can evaluate through the bit-pattern table as well (FastLut)
    if (!uniform && !(kind && BuildBLut(f.r, f.n, max_keys, blut))) return false;
    if (kind) *kind = uniform ? 0 : 1;
    KnotBasis kb;
    kb.Build(f.r, f.n);
    std::vector<double> coefs(f.n + 3, 0.0);
    SolveNatural(kb, f.f, 1, coefs.data(), 1);
    if (!AppendPP1Arrays(blob, d, f, kb, coefs.data(), uniform ? &lut : nullptr, uniform ? nullptr : &blut, max_keys)) return false;
    d.r_min = f.r[0];
    d.r_max = f.r[f.n - 1];
    return true;
}

/// Tables of potential_fast_kernel for an Ilkka, Bare or David action; fastv_ok stays false when a
/// grid admits no interval table (uniform; for v also bit-pattern) or the block does not fit in shared memory.
int BuildFastV(pimc_ctx *ctx, pimc_action *a, const pimc_table_1d &v_r, int is_coulomb, const pimc_long_range *lr) {
    a->fastv_ok = false;
    FastVTable &T = a->fastv;
    std::memset(&T, 0, sizeof(T));
    ByteBlob blob;
    const int kMaxKeys = 16384;
    T.is_coulomb = is_coulomb ? 1 : 0;
    const bool use_lr = a->use_long_range && lr != nullptr;  // David has no r-space long-range part (lr == nullptr)
    T.use_lr = use_lr ? 1 : 0;
    if (is_coulomb) {  // analytic 1/r: only the clamp limits of the (unused) table matter
        T.v.r_min = v_r.r[0];
        T.v.r_max = v_r.r[v_r.n - 1];
    } else if (!AppendFastPP1(blob, T.v, v_r, kMaxKeys, &T.v_kind)) {
        return PIMC_OK;
    }
    if (use_lr && !AppendFastPP1(blob, T.lr, lr->f_r, kMaxKeys)) return PIMC_OK;
    blob.b.resize((blob.b.size() + 15) & ~(size_t)15, 0);
    if (blob.b.empty()) blob.b.resize(16, 0);
    const size_t need = sizeof(double) * kFastRows * 3 * kFastRow + blob.b.size() + 10240;  // + static shared memory (ring, red) of the kernels that use it
    if (need > ctx->smem_optin) return PIMC_OK;
    T.n_bytes = (int)blob.b.size();
    PIMC_CUDA(a->fastv_tab.Alloc(blob.b.size()));
    PIMC_CUDA(cudaMemcpy(a->fastv_tab.p, blob.b.data(), blob.b.size(), cudaMemcpyHostToDevice));
    a->fastv_ok = true;
    return PIMC_OK;
}

/// Packs the shared-memory block of david_full_fast_kernel for U (which = WHICH_U) or dU/dbeta.
/// values[v][g] = the multi-spline's data as the reference fills it (david...:262-283).  Leaves
/// fastd_ok false when the grid admits neither interval table, n_order is outside 1..3 or the
/// block does not fit in shared memory (the general kernel then evaluates the action).
int BuildFastDavid(pimc_ctx *ctx, pimc_action *a, int which, const double *grid, int n, const std::vector<std::vector<double>> &values,
                   int n_order, double r_min, double r_max) {
    a->fastd_ok[which] = false;
    FastDavidTable &T = a->fastd[which];
    std::memset(&T, 0, sizeof(T));
    if (n_order < 1 || n_order > 3) return PIMC_OK;
    const int n_q = (int)values.size() - 2;
    if (n_q != n_order * (n_order + 3) / 2) return PIMC_OK;
    const int kMaxKeys = 16384;
    ByteBlob blob;
    ULut ul;
    BLut bl;
    if (BuildULut(grid, n, kMaxKeys, ul)) {
        T.lut.kind = 0;
        T.lut.off_lut = blob.Reserve(ul.lut.size() * sizeof(uint16_t));
        std::memcpy(blob.At<uint16_t>(T.lut.off_lut), ul.lut.data(), ul.lut.size() * sizeof(uint16_t));
        T.lut.key_max = (int)ul.lut.size() - 1;
        T.lut.inv_h = ul.inv_h;
    } else if (BuildBLut(grid, n, kMaxKeys, bl)) {
        T.lut.kind = 1;
        T.lut.off_lut = blob.Reserve(bl.lut.size() * sizeof(uint16_t));
        std::memcpy(blob.At<uint16_t>(T.lut.off_lut), bl.lut.data(), bl.lut.size() * sizeof(uint16_t));
        T.lut.key_max = (int)bl.lut.size() - 1;
        T.lut.shift = bl.shift;
        T.lut.key0 = bl.key0;
    } else {
        return PIMC_OK;
    }
    KnotBasis kb;
    kb.Build(grid, n);
    std::vector<double> coefs(n + 3, 0.0);
    auto pp_of = [&](const std::vector<double> &data) {
        std::fill(coefs.begin(), coefs.end(), 0.0);
        SolveNatural(kb, data.data(), 1, coefs.data(), 1);
        return PPFrom1D(kb, coefs.data());
    };
    T.off_gpair = AppendKnotPairs(blob, grid, n);
    // endpoint: value 1; dU/dbeta adds the potential (value 0), david...:137-145
    std::vector<double> pe = pp_of(values[1]);
    if (which == WHICH_DU) {
        const std::vector<double> p0 = pp_of(values[0]);
        for (size_t i = 0; i < pe.size(); ++i) pe[i] += p0[i];
    }
    T.off_e01 = blob.Reserve((size_t)n * 16);
    T.off_e23 = blob.Reserve((size_t)n * 16);
    for (int i = 0; i < n; ++i) {
        blob.At<double>(T.off_e01)[2 * i] = pe[4 * (size_t)i];
        blob.At<double>(T.off_e01)[2 * i + 1] = pe[4 * (size_t)i + 1];
        blob.At<double>(T.off_e23)[2 * i] = pe[4 * (size_t)i + 2];
        blob.At<double>(T.off_e23)[2 * i + 1] = pe[4 * (size_t)i + 3];
    }
    T.q_stride = 32 * n_q + 16;  // 2 n_q + 1 sixteen-byte slots: odd, neighbouring intervals fall into disjoint banks
    T.off_q = blob.Reserve((size_t)n * T.q_stride);
    for (int v = 0; v < n_q; ++v) {
        const std::vector<double> pq = pp_of(values[v + 2]);
        for (int i = 0; i < n; ++i)
            std::memcpy(blob.b.data() + T.off_q + (size_t)i * T.q_stride + 32 * (size_t)v, &pq[4 * (size_t)i], 32);
    }
    blob.b.resize((blob.b.size() + 15) & ~(size_t)15, 0);
    const size_t need = sizeof(double) * kFastRows * 3 * kFastRow + blob.b.size() + 10240;  // + the kernel's static shared memory
    if (need > ctx->smem_optin) return PIMC_OK;
    T.n_order = n_order;
    T.r_min = r_min;
    T.r_max = r_max;
    T.n_bytes = (int)blob.b.size();
    PIMC_CUDA(a->fastd_tab[which].Alloc(blob.b.size()));
    PIMC_CUDA(cudaMemcpy(a->fastd_tab[which].p, blob.b.data(), blob.b.size(), cudaMemcpyHostToDevice));
    a->fastd_ok[which] = true;
    return PIMC_OK;
}

// ------------------------------------------------------------------------------ launchers
template <int ATYPE, int WHICH>
int LaunchPairFullT(pimc_ctx *ctx, const PairFullArgs &args, size_t smem, int grid) {
    if (smem > 48 * 1024)
        PIMC_CUDA(cudaFuncSetAttribute(pair_full_kernel<ATYPE, WHICH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        ScopedKernelTimer t(ctx, PIMC_KERNEL_PAIR_FULL);
        pair_full_kernel<ATYPE, WHICH><<<grid, kPairThreads, smem, ctx->stream>>>(args);
    }
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return PIMC_OK;
}

int LaunchPairFast(pimc_action *a, int which, int *n_per_clone) {
    pimc_ctx *ctx = a->ctx;
    PairFastArgs args;
    args.pv = ctx->View();
    args.A = ctx->SView(a->sa, false);
    args.B = ctx->SView(a->sb, false);
    args.same = a->sa == a->sb;
    args.T = a->fast[which];
    args.tables = a->fast_tab[which].p;
    args.n_chunks = (ctx->Mloc + kChunk - 1) / kChunk;
    args.n_pgroups = (args.A.N + kFastWarps - 1) / kFastWarps;
    // too few items to fill the SMs (one large slice-sharded path): split the partner loop of an
    // item by windows of kFastQ partners until there are about six items per SM
    const int n_dd = args.same ? args.A.N / 2 : args.B.N;
    const int n_windows = std::max(1, (n_dd + kFastQ - 1) / kFastQ);
    const size_t base_items = (size_t)ctx->C * args.n_chunks * args.n_pgroups;
    int want = (int)std::min<size_t>((size_t)n_windows, ((size_t)6 * ctx->n_sm + base_items - 1) / std::max<size_t>(base_items, 1));
    want = std::max(want, 1);
    args.t_windows = std::max(1, n_windows / want);
    args.n_tsplit = (n_windows + args.t_windows - 1) / args.t_windows;
    *n_per_clone = args.n_chunks * args.n_pgroups * args.n_tsplit;
    const size_t items = (size_t)ctx->C * *n_per_clone;
    if (ctx->partial.n < items) PIMC_CUDA(ctx->partial.Alloc(items));
    args.partial = ctx->partial.p;
    const size_t smem = sizeof(double) * kFastRows * 3 * kFastRow + (size_t)args.T.n_bytes;
    PIMC_CUDA(cudaFuncSetAttribute(pair_full_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<size_t>(items, (size_t)ctx->n_sm);
    {
        ScopedKernelTimer t(ctx, PIMC_KERNEL_PAIR_FULL);
        pair_full_fast_kernel<<<grid, kFastThreads, smem, ctx->stream>>>(args);
    }
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return PIMC_OK;
}

/// Items and partner-window split shared by the fast whole-path kernels (see LaunchPairFast).
void FastItems(const pimc_ctx *ctx, int Na, int Nb, bool same, int n_chunks, int &n_pgroups, int &n_tsplit, int &t_windows) {
    n_pgroups = (Na + kFastWarps - 1) / kFastWarps;
    const int n_dd = same ? Na / 2 : Nb;
    const int n_windows = std::max(1, (n_dd + kFastQ - 1) / kFastQ);
    const size_t base_items = (size_t)ctx->C * n_chunks * n_pgroups;
    int want = (int)std::min<size_t>((size_t)n_windows, ((size_t)6 * ctx->n_sm + base_items - 1) / std::max<size_t>(base_items, 1));
    want = std::max(want, 1);
    t_windows = std::max(1, n_windows / want);
    n_tsplit = (n_windows + t_windows - 1) / t_windows;
}

int LaunchPotentialFast(pimc_action *a, int *n_per_clone) {
    pimc_ctx *ctx = a->ctx;
    PotFastArgs args;
    args.pv = ctx->View();
    args.A = ctx->SView(a->sa, false);
    args.B = ctx->SView(a->sb, false);
    args.same = a->sa == a->sb;
    args.T = a->fastv;
    args.tables = a->fastv_tab.p;
    args.n_chunks = (ctx->Mloc + (ctx->sharded ? 1 : 0) + kChunk - 1) / kChunk;
    FastItems(ctx, args.A.N, args.B.N, args.same != 0, args.n_chunks, args.n_pgroups, args.n_tsplit, args.t_windows);
    *n_per_clone = args.n_chunks * args.n_pgroups * args.n_tsplit;
    const size_t items = (size_t)ctx->C * *n_per_clone;
    if (ctx->partial.n < items) PIMC_CUDA(ctx->partial.Alloc(items));
    args.partial = ctx->partial.p;
    const size_t smem = sizeof(double) * kFastRows * 3 * kFastRow + (size_t)args.T.n_bytes;
    auto *kernel = args.T.v_kind == 1 ? potential_fast_kernel<1> : potential_fast_kernel<0>;
    PIMC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<size_t>(items, (size_t)ctx->n_sm);
    {
        ScopedKernelTimer t(ctx, PIMC_KERNEL_PAIR_FULL);
        kernel<<<grid, kFastThreads, smem, ctx->stream>>>(args);
    }
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return PIMC_OK;
}

/// Bare U / dU/dbeta over the whole path through the fast V tables (CalcV of DrDrpDrrp's r, r').
int LaunchBareFast(pimc_action *a, int which, int *n_per_clone) {
    pimc_ctx *ctx = a->ctx;
    BareFastArgs args;
    args.pv = ctx->View();
    args.A = ctx->SView(a->sa, false);
    args.B = ctx->SView(a->sb, false);
    args.same = a->sa == a->sb;
    args.T = a->fastv;
    args.tables = a->fastv_tab.p;
    args.scale = which == WHICH_U ? a->table[WHICH_U].u_scale : 1.;
    args.n_chunks = (ctx->Mloc + kChunk - 1) / kChunk;
    FastItems(ctx, args.A.N, args.B.N, args.same != 0, args.n_chunks, args.n_pgroups, args.n_tsplit, args.t_windows);
    *n_per_clone = args.n_chunks * args.n_pgroups * args.n_tsplit;
    const size_t items = (size_t)ctx->C * *n_per_clone;
    if (ctx->partial.n < items) PIMC_CUDA(ctx->partial.Alloc(items));
    args.partial = ctx->partial.p;
    const size_t smem = sizeof(double) * kFastRows * 3 * kFastRow + (size_t)args.T.n_bytes;
    PIMC_CUDA(cudaFuncSetAttribute(bare_full_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<size_t>(items, (size_t)ctx->n_sm);
    {
        ScopedKernelTimer t(ctx, PIMC_KERNEL_PAIR_FULL);
        bare_full_fast_kernel<<<grid, kFastThreads, smem, ctx->stream>>>(args);
    }
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return PIMC_OK;
}

template <int NORD, int KIND>
int LaunchDavidFastT(pimc_ctx *ctx, const DavidFastArgs &args, size_t smem, int grid) {
    PIMC_CUDA(cudaFuncSetAttribute(david_full_fast_kernel<NORD, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        ScopedKernelTimer t(ctx, PIMC_KERNEL_PAIR_FULL);
        david_full_fast_kernel<NORD, KIND><<<grid, kFastThreads, smem, ctx->stream>>>(args);
    }
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return PIMC_OK;
}

/// David U / dU/dbeta over the whole path with every table in shared memory.
int LaunchDavidFast(pimc_action *a, int which, int *n_per_clone) {
    pimc_ctx *ctx = a->ctx;
    DavidFastArgs args;
    args.pv = ctx->View();
    args.A = ctx->SView(a->sa, false);
    args.B = ctx->SView(a->sb, false);
    args.same = a->sa == a->sb;
    args.T = a->fastd[which];
    args.tables = a->fastd_tab[which].p;
    args.n_chunks = (ctx->Mloc + kChunk - 1) / kChunk;
    FastItems(ctx, args.A.N, args.B.N, args.same != 0, args.n_chunks, args.n_pgroups, args.n_tsplit, args.t_windows);
    *n_per_clone = args.n_chunks * args.n_pgroups * args.n_tsplit;
    const size_t items = (size_t)ctx->C * *n_per_clone;
    if (ctx->partial.n < items) PIMC_CUDA(ctx->partial.Alloc(items));
    args.partial = ctx->partial.p;
    const size_t smem = sizeof(double) * kFastRows * 3 * kFastRow + (size_t)args.T.n_bytes;
    const int grid = (int)std::min<size_t>(items, (size_t)ctx->n_sm);
    const int key = args.T.n_order * 2 + args.T.lut.kind;
    switch (key) {
        case 2: return LaunchDavidFastT<1, 0>(ctx, args, smem, grid);
        case 3: return LaunchDavidFastT<1, 1>(ctx, args, smem, grid);
        case 4: return LaunchDavidFastT<2, 0>(ctx, args, smem, grid);
        case 5: return LaunchDavidFastT<2, 1>(ctx, args, smem, grid);
        case 6: return LaunchDavidFastT<3, 0>(ctx, args, smem, grid);
        default: return LaunchDavidFastT<3, 1>(ctx, args, smem, grid);
    }
}

int LaunchPairFull(pimc_action *a, int which, bool independent_images, int *n_per_clone) {
    pimc_ctx *ctx = a->ctx;
    if (which == WHICH_V && independent_images && a->fastv_ok && !ctx->force_general) return LaunchPotentialFast(a, n_per_clone);
    if (a->atype == ATYPE_BARE && which != WHICH_V && !independent_images && a->fastv_ok && a->fastv.v_kind == 0 && !ctx->force_general)
        return LaunchBareFast(a, which, n_per_clone);
    if (a->atype == ATYPE_ILKKA && which != WHICH_V && !independent_images && a->fast_ok[which] && !ctx->force_general)
        return LaunchPairFast(a, which, n_per_clone);
    if (a->atype == ATYPE_DAVID && which != WHICH_V && !independent_images && a->fastd_ok[which] && !ctx->force_general)
        return LaunchDavidFast(a, which, n_per_clone);
    PairFullArgs args;
    args.pv = ctx->View();
    args.A = ctx->SView(a->sa, false);
    args.B = ctx->SView(a->sb, false);
    args.same = a->sa == a->sb;
    args.T = a->table[which];
    args.blob = a->blob[which].p;
    args.blob_doubles = (int)a->blob[which].n;
    args.independent_images = independent_images ? 1 : 0;
    args.n_chunks = (ctx->Mloc + kChunk - 1) / kChunk;
    args.n_pgroups = (args.A.N + kPairWarps - 1) / kPairWarps;
    *n_per_clone = args.n_chunks * args.n_pgroups;
    const size_t items = (size_t)ctx->C * *n_per_clone;
    if (ctx->partial.n < items) PIMC_CUDA(ctx->partial.Alloc(items));
    args.partial = ctx->partial.p;
    const size_t pos_bytes = sizeof(double) * kQTile * 3 * kRow;
    const size_t blob_bytes = a->stageable[which] ? sizeof(double) * a->blob[which].n : 0;
    args.stage = (a->stageable[which] && pos_bytes + blob_bytes + 1024 <= ctx->smem_optin) ? 1 : 0;
    const size_t smem = pos_bytes + (args.stage ? blob_bytes : 0);
    // persistent CTAs: one per SM (1024 threads), never more than the items
    const int grid = (int)std::min<size_t>(items, (size_t)ctx->n_sm);
#define PIMC_DISPATCH(AT)                                                                    \
    switch (which) {                                                                         \
        case WHICH_U: return LaunchPairFullT<AT, WHICH_U>(ctx, args, smem, grid);            \
        case WHICH_DU: return LaunchPairFullT<AT, WHICH_DU>(ctx, args, smem, grid);          \
        default: return LaunchPairFullT<AT, WHICH_V>(ctx, args, smem, grid);                 \
    }
    switch (a->atype) {
        case ATYPE_ILKKA: PIMC_DISPATCH(ATYPE_ILKKA)
        case ATYPE_BARE: PIMC_DISPATCH(ATYPE_BARE)
        default: PIMC_DISPATCH(ATYPE_DAVID)
    }
#undef PIMC_DISPATCH
}

/// Long-range k sum over the whole shard (b0 == nullptr) or over a window, into ctx->lr_dev.
int LaunchKSum(pimc_action *a, int which, const int32_t *d_b0, int n_window, bool use_delta, double scale) {
    pimc_ctx *ctx = a->ctx;
    if ((int)ctx->lr_dev.n < ctx->C) PIMC_CUDA(ctx->lr_dev.Alloc(ctx->C));
    KSumArgs k;
    k.pv = ctx->View();
    k.n_k = ctx->n_k();
    k.rho_a = ctx->species[a->sa]->rho.p;
    k.rho_b = ctx->species[a->sb]->rho.p;
    k.drho_a = (use_delta && ctx->species[a->sa]->drho_valid) ? ctx->species[a->sa]->drho.p : nullptr;
    k.drho_b = (use_delta && ctx->species[a->sb]->drho_valid) ? ctx->species[a->sb]->drho.p : nullptr;
    k.wk = a->wk[which].p;
    k.b0 = d_b0;
    k.n_window = n_window;
    k.twice = a->sa != a->sb;
    k.scale = scale;
    k.accumulate = 0;
    k.out = ctx->lr_dev.p;
    {
        ScopedKernelTimer t(ctx, PIMC_KERNEL_KSUM);
        ksum_kernel<<<ctx->C, 256, 0, ctx->stream>>>(k);
    }
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return PIMC_OK;
}

int Finalize(pimc_ctx *ctx, int n_per_clone, bool add_lr, double k0, double r0, bool add_const, double *d_out) {
    finalize_kernel<<<(ctx->C + 127) / 128, 128, 0, ctx->stream>>>(ctx->partial.p, ctx->C, n_per_clone, ctx->lr_dev.p, add_lr ? 1 : 0, k0,
                                                                   r0, add_const ? 1 : 0, d_out);
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return PIMC_OK;
}

/// DActionDBeta (which = DU), Potential (V) or the whole-path action (U) into device memory.
int KineticFull(pimc_action *a, int which, double *d_out);

int FullEvaluation(pimc_action *a, int which, double *d_out) {
    pimc_ctx *ctx = a->ctx;
    if (a->atype == ATYPE_KINETIC) return KineticFull(a, which, d_out);
    if (a->is_constant) {
        // App. A-1: a constant action has no pairs (one particle) or never changes; the
        // reference caches its first value.  Only the no-pair case is evaluated here.
        if (ctx->species[a->sa]->N != 1)
            return Fail(PIMC_ERR_UNSUPPORTED, "constant action with lambda = 0 and more than one particle");
    }
    int n_per_clone = 0;
    int rc = LaunchPairFull(a, which, which == WHICH_V, &n_per_clone);
    if (rc != PIMC_OK) return rc;
    if (a->use_long_range) {
        if (ctx->n_k() > 0) {
            const double scale = (which == WHICH_U) ? a->ulong_scale : 1.0;
            rc = LaunchKSum(a, which, nullptr, 0, false, scale);
            if (rc != PIMC_OK) return rc;
        } else {  // no k vector inside the cutoff: empty k sum, the constants remain
            if ((int)ctx->lr_dev.n < ctx->C) PIMC_CUDA(ctx->lr_dev.Alloc(ctx->C));
            PIMC_CUDA(cudaMemsetAsync(ctx->lr_dev.p, 0, ctx->C * sizeof(double), ctx->stream));
        }
    }
    // constants belong to the rank that owns slice 0 when the path is sharded
    const bool add_const = (which != WHICH_U) && (!ctx->sharded || ctx->slice_lo == 0);
    const double k0 = a->k0[which], r0 = a->r0[which];
    return Finalize(ctx, n_per_clone, a->use_long_range, k0, r0, add_const, d_out);
}

/// Kinetic::GetAction (kinetic_class.h:105-122) for `n_listed` particles per clone (device array part[n_listed][C])
/// over the window [b0, b0 + n_window) at `level`, OLD or NEW mode, into d_out[C].
int KineticWindow(pimc_action *a, int mode, const int32_t *d_b0, int n_window, int n_listed, const int32_t *d_part, int level, double *d_out) {
    pimc_ctx *ctx = a->ctx;
    if (level + 1 >= kMaxFreeSplines) return Fail(PIMC_ERR_UNSUPPORTED, "Kinetic action above level 5");
    const int skip = 1 << level;
    if (n_window % skip) return Fail(PIMC_ERR_INVALID, "window is not a multiple of 2^level");
    FreeSet *fs = nullptr;
    int rc = GetFreeSet(ctx, a->sa, a->n_images, &fs);
    if (rc != PIMC_OK) return rc;
    KineticWindowArgs w;
    w.pv = ctx->View();
    w.sv = ctx->SView(a->sa, mode == PIMC_NEW);
    w.tab = fs->view.s[level + 1];
    w.n_listed = n_listed;
    w.part = d_part;
    w.b0 = d_b0;
    w.n_links = n_window / skip;
    w.skip = skip;
    w.mode = mode;
    w.out = d_out;
    kinetic_window_kernel<<<(ctx->C + 3) / 4, 128, 0, ctx->stream>>>(w);
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return PIMC_OK;
}

/// Kinetic::DActionDBeta (kinetic_class.h:35-45); Potential() of a Kinetic action is Action's default 0
/// (action_class.h:43); the whole-path action is GetAction(0, n_bead, every particle, 0).
int KineticFull(pimc_action *a, int which, double *d_out) {
    pimc_ctx *ctx = a->ctx;
    SpeciesState &st = *ctx->species[a->sa];
    if (which == WHICH_V) {
        PIMC_CUDA(cudaMemsetAsync(d_out, 0, ctx->C * sizeof(double), ctx->stream));
        return PIMC_OK;
    }
    FreeSet *fs = nullptr;
    int rc = GetFreeSet(ctx, a->sa, a->n_images, &fs);
    if (rc != PIMC_OK) return rc;
    KineticFullArgs k;
    k.pv = ctx->View();
    k.sv = ctx->SView(a->sa, false);
    k.n_chunks = (ctx->Mloc + 31) / 32;
    k.next = st.perm_tracked ? st.perm_next.p : nullptr;  // the link over the beta seam follows the permutation (GetNextBead)
    const size_t items = (size_t)ctx->C * k.n_chunks;
    if (ctx->partial.n < items) PIMC_CUDA(ctx->partial.Alloc(items));
    k.partial = ctx->partial.p;
    double constant = 0.;
    if (which == WHICH_DU) {
        k.dtau = fs->dtau;
        // d_action_d_beta_const = N M n_d / (2 tau) (kinetic_class.h:31), on the rank that owns slice 0
        constant = (!ctx->sharded || ctx->slice_lo == 0) ? st.N * (double)ctx->M * ctx->n_d / (2. * ctx->tau) : 0.;
    } else {  // -log rho_free of every link: the same sum with the action table and the opposite sign
        k.dtau = fs->view.s[1];
    }
    kinetic_dbeta_kernel<<<(int)std::min<size_t>(items, (size_t)ctx->n_sm * 8), 256, 0, ctx->stream>>>(k);
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    if ((int)ctx->lr_dev.n < ctx->C) PIMC_CUDA(ctx->lr_dev.Alloc(ctx->C));
    PIMC_CUDA(cudaMemsetAsync(ctx->lr_dev.p, 0, ctx->C * sizeof(double), ctx->stream));
    kinetic_finalize_kernel<<<(ctx->C + 127) / 128, 128, 0, ctx->stream>>>(ctx->partial.p, ctx->C, k.n_chunks, which == WHICH_DU ? 1. : -1., constant, d_out);
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return PIMC_OK;
}

int ToHost(pimc_ctx *ctx, const double *d_src, double *host, size_t n) {
    PIMC_CUDA(cudaMemcpyAsync(host, d_src, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PIMC_CUDA(cudaStreamSynchronize(ctx->stream));
    return PIMC_OK;
}

int EnsureOut(pimc_ctx *ctx) {
    if ((int)ctx->out_dev.n < ctx->C) PIMC_CUDA(ctx->out_dev.Alloc(ctx->C));
    return PIMC_OK;
}

}  // namespace pimc_host

// =========================================================================== C ABI
extern "C" {

const char *pimc_last_error(void) { return g_last_error.c_str(); }
int pimc_version(void) { return 100; }

int pimc_ctx_create(const pimc_config *cfg, pimc_ctx **out) {
    if (!cfg || !out) return Fail(PIMC_ERR_INVALID, "null config or output");
    if (cfg->n_d != 3) return Fail(PIMC_ERR_UNSUPPORTED, "only n_d = 3 is evaluated on the device");
    if (cfg->n_bead < 1 || cfg->n_species < 1 || cfg->n_clones < 1) return Fail(PIMC_ERR_INVALID, "n_bead, n_species, n_clones must be >= 1");
    if (cfg->pbc && !(cfg->L > 0.)) return Fail(PIMC_ERR_INVALID, "periodic box needs L > 0");
    int lo = cfg->slice_lo, hi = cfg->slice_hi;
    if (lo == 0 && hi == 0) hi = cfg->n_bead;
    if (lo < 0 || hi > cfg->n_bead || lo >= hi) return Fail(PIMC_ERR_INVALID, "bad slice shard");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return Fail(PIMC_ERR_CUDA, "no CUDA device visible; simpimc_b200 has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= n_dev) return Fail(PIMC_ERR_INVALID, "device ordinal out of range");
    PIMC_CUDA(cudaSetDevice(cfg->device));
    std::unique_ptr<pimc_ctx> ctx(new pimc_ctx);
    ctx->n_d = cfg->n_d;
    ctx->pbc = cfg->pbc != 0;
    ctx->M = cfg->n_bead;
    ctx->C = cfg->n_clones;
    ctx->device = cfg->device;
    ctx->beta = cfg->beta;
    ctx->tau = cfg->beta / (1. * cfg->n_bead);
    if (ctx->pbc) {  // path_class.h:33-42
        ctx->L = cfg->L;
        ctx->iL = 1. / cfg->L;
        ctx->vol = std::pow(cfg->L, cfg->n_d);
    } else {
        ctx->L = 0.;
        ctx->iL = 0.;
        ctx->vol = 1.;
    }
    ctx->slice_lo = lo;
    ctx->slice_hi = hi;
    ctx->Mloc = hi - lo;
    ctx->sharded = (ctx->Mloc != ctx->M);
    ctx->Mstore = ctx->Mloc + (ctx->sharded ? 1 : 0);
    ctx->Ms = (ctx->Mstore + 3) & ~3;
    PIMC_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    cudaDeviceProp prop;
    PIMC_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    ctx->n_sm = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    for (int s = 0; s < cfg->n_species; ++s) {
        std::unique_ptr<SpeciesState> st(new SpeciesState);
        st->N = cfg->n_part[s];
        if (st->N < 1) return Fail(PIMC_ERR_INVALID, "species without particles");
        st->lambda = cfg->lambda[s];
        const size_t n = (size_t)ctx->C * st->N * 3 * ctx->Ms;
        PIMC_CUDA(st->R.Alloc(n));
        PIMC_CUDA(cudaMemsetAsync(st->R.p, 0, n * sizeof(double), ctx->stream));
        PIMC_CUDA(st->P_particle.Alloc((size_t)kMaxPropSlots * ctx->C));
        PIMC_CUDA(st->P_first.Alloc((size_t)kMaxPropSlots * ctx->C));
        PIMC_CUDA(st->drho_b0.Alloc(ctx->C));
        ctx->species.push_back(std::move(st));
    }
    PIMC_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = ctx.release();
    return PIMC_OK;
}

int pimc_ctx_destroy(pimc_ctx *ctx) {
    if (!ctx) return PIMC_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (pimc_action *a : ctx->actions) delete a;
    for (cudaStream_t st : ctx->side_streams) {
        cudaStreamSynchronize(st);
        cudaStreamDestroy(st);
    }
    for (cudaEvent_t ev : ctx->side_events) cudaEventDestroy(ev);
    if (ctx->fork_event) cudaEventDestroy(ctx->fork_event);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return PIMC_OK;
}

int pimc_ctx_sync(pimc_ctx *ctx) {
    if (!ctx) return Fail(PIMC_ERR_INVALID, "null context");
    PIMC_CUDA(cudaStreamSynchronize(ctx->stream));
    return PIMC_OK;
}
void *pimc_ctx_stream(pimc_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int64_t pimc_ctx_launch_count(pimc_ctx *ctx) { return ctx ? ctx->launches : 0; }

int pimc_kspace_setup(pimc_ctx *ctx, double k_cut, int32_t *n_k) {
    if (!ctx) return Fail(PIMC_ERR_INVALID, "null context");
    if (!ctx->pbc) return Fail(PIMC_ERR_INVALID, "k space needs a periodic box");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    if (k_cut > ctx->k_cutoff) {  // grow-only (k_space_class.h:34-41)
        BuildKSpace(ctx, k_cut);
        int rc = UploadKSpace(ctx);
        if (rc != PIMC_OK) return rc;
        // actions created for the smaller set keep their shell tables: match them to the new vectors
        // (the reference's actions would keep stale per-vector arrays here; its inputs never grow the
        // set after an action exists, so the well-defined behaviour is the one implemented)
        if ((rc = RematchWeights(ctx)) != PIMC_OK) return rc;
        for (size_t s = 0; s < ctx->species.size(); ++s) {
            rc = RebuildRhoK(ctx, (int)s);
            if (rc != PIMC_OK) return rc;
        }
    }
    if (n_k) *n_k = ctx->n_k();
    return PIMC_OK;
}

int pimc_kspace_get(pimc_ctx *ctx, int32_t *k_index, double *k_mag) {
    if (!ctx) return Fail(PIMC_ERR_INVALID, "null context");
    if (k_index) std::memcpy(k_index, ctx->k_index.data(), ctx->k_index.size() * sizeof(int32_t));
    if (k_mag) std::memcpy(k_mag, ctx->k_mag.data(), ctx->k_mag.size() * sizeof(double));
    return PIMC_OK;
}

int pimc_positions_upload(pimc_ctx *ctx, int32_t s, int32_t clone_lo, int32_t clone_hi, const double *R) {
    if (!ctx || !R) return Fail(PIMC_ERR_INVALID, "null argument");
    if (s < 0 || s >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    if (clone_lo < 0 || clone_hi > ctx->C || clone_lo >= clone_hi) return Fail(PIMC_ERR_INVALID, "bad clone range");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    SpeciesState &st = *ctx->species[s];
    const int nc = clone_hi - clone_lo;
    const size_t n_host = (size_t)nc * st.N * ctx->Mstore * 3;
    if (ctx->stage.n < n_host) PIMC_CUDA(ctx->stage.Alloc(n_host));
    PIMC_CUDA(cudaMemcpyAsync(ctx->stage.p, R, n_host * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    double *dst = st.R.p + (size_t)clone_lo * st.N * 3 * ctx->Ms;
    positions_in_kernel<<<ctx->n_sm * 8, 256, 0, ctx->stream>>>(ctx->stage.p, nc, st.N, ctx->Mstore, ctx->Ms, dst);
    st.R2_valid = false;
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    st.n_prop = 0;
    st.n_slots = 0;
    if (clone_lo == 0 && clone_hi == ctx->C) return RebuildRhoK(ctx, s);
    st.need_update_rho_k = true;
    st.drho_valid = false;
    return PIMC_OK;  // partial upload: caller finishes with pimc_rhok_rebuild
}

int pimc_positions_download(pimc_ctx *ctx, int32_t s, int32_t mode, int32_t clone_lo, int32_t clone_hi, double *R) {
    if (!ctx || !R) return Fail(PIMC_ERR_INVALID, "null argument");
    if (s < 0 || s >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    if (clone_lo < 0 || clone_hi > ctx->C || clone_lo >= clone_hi) return Fail(PIMC_ERR_INVALID, "bad clone range");
    if (mode != PIMC_OLD) return Fail(PIMC_ERR_UNSUPPORTED, "download returns the committed path; proposals live with the caller");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    SpeciesState &st = *ctx->species[s];
    const int nc = clone_hi - clone_lo;
    const size_t n_host = (size_t)nc * st.N * ctx->Mstore * 3;
    if (ctx->stage.n < n_host) PIMC_CUDA(ctx->stage.Alloc(n_host));
    const double *src = st.R.p + (size_t)clone_lo * st.N * 3 * ctx->Ms;
    positions_out_kernel<<<ctx->n_sm * 8, 256, 0, ctx->stream>>>(src, nc, st.N, ctx->Mstore, ctx->Ms, ctx->stage.p);
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return ToHost(ctx, ctx->stage.p, R, n_host);
}

int pimc_positions_set_device(pimc_ctx *ctx, int32_t s, const double *d_R) {
    if (!ctx || !d_R) return Fail(PIMC_ERR_INVALID, "null argument");
    if (s < 0 || s >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    SpeciesState &st = *ctx->species[s];
    PIMC_CUDA(cudaMemcpyAsync(st.R.p, d_R, st.R.n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    st.R2_valid = false;
    st.n_prop = 0;
    st.n_slots = 0;
    return RebuildRhoK(ctx, s);
}

double *pimc_positions_device_ptr(pimc_ctx *ctx, int32_t s) {
    if (!ctx || s < 0 || s >= (int)ctx->species.size()) return nullptr;
    ctx->species[s]->R2_valid = false;  // the caller may write through the pointer
    return ctx->species[s]->R.p;
}

int pimc_halo_pack(pimc_ctx *ctx, int32_t s, double *d_buf) {
    if (!ctx || !d_buf) return Fail(PIMC_ERR_INVALID, "null argument");
    if (s < 0 || s >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    SpeciesState &st = *ctx->species[s];
    const size_t n_rows = (size_t)ctx->C * st.N * 3;
    halo_pack_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, ctx->stream>>>(st.R.p, n_rows, ctx->Ms, 0, d_buf);
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return PIMC_OK;
}

int pimc_halo_unpack(pimc_ctx *ctx, int32_t s, const double *d_buf) {
    if (!ctx || !d_buf) return Fail(PIMC_ERR_INVALID, "null argument");
    if (s < 0 || s >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    if (!ctx->sharded) return Fail(PIMC_ERR_INVALID, "an unsharded context has no halo slice");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    SpeciesState &st = *ctx->species[s];
    const size_t n_rows = (size_t)ctx->C * st.N * 3;
    halo_unpack_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, ctx->stream>>>(st.R.p, n_rows, ctx->Ms, ctx->Mloc, d_buf);
    st.R2_valid = false;
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return PIMC_OK;
}

int pimc_rotate_pack(pimc_ctx *ctx, int32_t s, int32_t shift, double *d_buf) {
    if (!ctx || !d_buf) return Fail(PIMC_ERR_INVALID, "null argument");
    if (s < 0 || s >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    if (!ctx->sharded) return Fail(PIMC_ERR_INVALID, "rotation applies to slice-sharded contexts");
    if (shift < 1 || shift > ctx->Mloc) return Fail(PIMC_ERR_INVALID, "shift must be in 1..slices of the shard");
    for (auto &sp : ctx->species)
        if (sp->n_prop > 0) return Fail(PIMC_ERR_INVALID, "a proposal is pending: call pimc_commit first");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    SpeciesState &st = *ctx->species[s];
    const size_t n = (size_t)ctx->C * st.N * 3 * shift;
    rotate_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(st.R.p, (size_t)ctx->C * st.N * 3, ctx->Ms, shift, d_buf);
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return PIMC_OK;
}

int pimc_rotate_apply(pimc_ctx *ctx, int32_t s, int32_t shift, const double *d_buf) {
    if (!ctx || !d_buf) return Fail(PIMC_ERR_INVALID, "null argument");
    if (s < 0 || s >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    if (!ctx->sharded) return Fail(PIMC_ERR_INVALID, "rotation applies to slice-sharded contexts");
    if (shift < 1 || shift > ctx->Mloc) return Fail(PIMC_ERR_INVALID, "shift must be in 1..slices of the shard");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    SpeciesState &st = *ctx->species[s];
    const size_t n_rows = (size_t)ctx->C * st.N * 3;
    rotate_apply_kernel<<<(unsigned)((n_rows + 127) / 128), 128, 0, ctx->stream>>>(st.R.p, n_rows, ctx->Ms, ctx->Mloc, shift, d_buf);
    st.R2_valid = false;
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    // rho_k of the shifted slices: the caller rebuilds (pimc_rhok_rebuild) after the halo exchange
    st.drho_valid = false;
    st.need_update_rho_k = true;
    return PIMC_OK;
}

int pimc_rhok_rebuild(pimc_ctx *ctx, int32_t s) {
    if (!ctx) return Fail(PIMC_ERR_INVALID, "null context");
    if (s < 0 || s >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    return RebuildRhoK(ctx, s);
}

int pimc_rhok_download(pimc_ctx *ctx, int32_t s, int32_t mode, int32_t clone, double *out) {
    if (!ctx || !out) return Fail(PIMC_ERR_INVALID, "null argument");
    if (s < 0 || s >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    if (clone < 0 || clone >= ctx->C) return Fail(PIMC_ERR_INVALID, "clone out of range");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    SpeciesState &st = *ctx->species[s];
    const int n_k = ctx->n_k();
    if (n_k == 0) return PIMC_OK;
    const size_t n = (size_t)ctx->Mloc * n_k;
    std::vector<double> host(2 * n);
    PIMC_CUDA(cudaMemcpyAsync(host.data(), st.rho.p + (size_t)clone * n, n * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    PIMC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (mode == PIMC_NEW && st.drho_valid) {
        // rho_k (NEW) = committed + increment on the window the increment was computed for
        std::vector<double> d((size_t)2 * st.drho_window * n_k);
        int32_t b0;
        PIMC_CUDA(cudaMemcpy(d.data(), st.drho.p + (size_t)clone * st.drho_window * n_k, d.size() * sizeof(double), cudaMemcpyDeviceToHost));
        PIMC_CUDA(cudaMemcpy(&b0, st.drho_b0.p + clone, sizeof(int32_t), cudaMemcpyDeviceToHost));
        for (int j = 0; j < st.drho_window; ++j) {
            int b = (b0 + j) % ctx->M - ctx->slice_lo;
            for (int k = 0; k < 2 * n_k; ++k) host[(size_t)b * 2 * n_k + k] += d[(size_t)j * 2 * n_k + k];
        }
    }
    std::memcpy(out, host.data(), host.size() * sizeof(double));
    return PIMC_OK;
}

// -------------------------------------------------------------------------------- actions
int pimc_action_create_ilkka(pimc_ctx *ctx, int32_t sa, int32_t sb, const pimc_ilkka_tables *t, int32_t max_level,
                             int32_t use_lr, double k_cut, pimc_action **out) {
    if (!t) return Fail(PIMC_ERR_INVALID, "null tables");
    pimc_action *a = nullptr;
    int rc = NewAction(ctx, ATYPE_ILKKA, sa, sb, max_level, use_lr, k_cut, &a);
    if (rc != PIMC_OK) return rc;
    std::unique_ptr<pimc_action> guard(a);
    PIMC_CUDA(cudaSetDevice(ctx->device));
    try {
        const pimc_table_2d *xy[2] = {&t->u_xy, &t->du_xy};
        const pimc_long_range *lrs[3] = {&t->u_long, &t->du_long, &t->v_long};
        for (int which = 0; which < 2; ++which) {
            const pimc_table_2d &g = *xy[which];
            if (g.n_x < 4 || g.n_y < 4 || !g.x || !g.y || !g.f) throw std::invalid_argument("off-diagonal table missing");
            // tensor-product natural spline (create_NUBspline_2d_d), then one bicubic per cell
            std::vector<double> bs = BuildBlob2D(g.x, g.n_x, g.y, g.n_y, g.f);
            KnotBasis bx, by;
            bx.Build(g.x, g.n_x);
            by.Build(g.y, g.n_y);
            const size_t off_c = (size_t)(g.n_x + 5) + 3 * (g.n_x + 2) + (g.n_y + 5) + 3 * (g.n_y + 2);
            std::vector<double> cells = PPFrom2D(bx, by, bs.data() + off_c);
            rc = UploadBlob(ctx, a->cells[which], cells);
            if (rc != PIMC_OK) return rc;
            PairTable &T = a->table[which];
            std::memset(&T, 0, sizeof(T));
            std::vector<double> blob;
            T.xy.nx = g.n_x;
            T.xy.ny = g.n_y;
            T.xy.off_gx = (int)blob.size();
            blob.insert(blob.end(), g.x, g.x + g.n_x);
            T.xy.off_gy = (int)blob.size();
            blob.insert(blob.end(), g.y, g.y + g.n_y);
            AppendLut(blob, T.xy.lutx, g.x, g.n_x);
            AppendLut(blob, T.xy.luty, g.y, g.n_y);
            T.xy.cells = a->cells[which].p;
            T.use_lr = a->use_long_range;
            if (a->use_long_range) {
                rc = LoadLongRange(ctx, a, which, *lrs[which], blob, T.lr);
                if (rc != PIMC_OK) return rc;
            }
            rc = UploadBlob(ctx, a->blob[which], blob);
            if (rc != PIMC_OK) return rc;
            rc = BuildFastIlkka(ctx, a, which, g, cells, a->use_long_range ? lrs[which] : nullptr);
            if (rc != PIMC_OK) return rc;
        }
        {
            CheckTable1D(t->v_r, "v_r table");
            PairTable &T = a->table[WHICH_V];
            std::memset(&T, 0, sizeof(T));
            std::vector<double> blob;
            AppendPP1(blob, T.a, t->v_r.r, t->v_r.f, t->v_r.n);
            T.use_lr = a->use_long_range;
            if (a->use_long_range) {
                rc = LoadLongRange(ctx, a, WHICH_V, t->v_long, blob, T.lr);
                if (rc != PIMC_OK) return rc;
            }
            rc = UploadBlob(ctx, a->blob[WHICH_V], blob);
            if (rc != PIMC_OK) return rc;
            rc = BuildFastV(ctx, a, t->v_r, 0, a->use_long_range ? &t->v_long : nullptr);
            if (rc != PIMC_OK) return rc;
        }
    } catch (const std::exception &e) {
        return Fail(PIMC_ERR_TABLE, e.what());
    }
    if (a->use_long_range) {
        // only du and v constants are scaled and used (ilkka...:421-431); u has none in CalcULong
        ScaleConstants(ctx, a, a->k0[WHICH_DU], a->r0[WHICH_DU]);
        ScaleConstants(ctx, a, a->k0[WHICH_V], a->r0[WHICH_V]);
    }
    ctx->actions.push_back(guard.release());
    *out = a;
    return PIMC_OK;
}

int pimc_action_create_bare(pimc_ctx *ctx, int32_t sa, int32_t sb, const pimc_bare_tables *t, int32_t max_level,
                            int32_t use_lr, double k_cut, pimc_action **out) {
    if (!t) return Fail(PIMC_ERR_INVALID, "null tables");
    pimc_action *a = nullptr;
    int rc = NewAction(ctx, ATYPE_BARE, sa, sb, max_level, use_lr, k_cut, &a);
    if (rc != PIMC_OK) return rc;
    std::unique_ptr<pimc_action> guard(a);
    PIMC_CUDA(cudaSetDevice(ctx->device));
    try {
        CheckTable1D(t->v_r, "v_r table");
        // U, dU/dbeta and V all evaluate CalcV (bare...:146-150,175-182); one blob each keeps
        // the kernels' addressing uniform
        for (int which = 0; which < 3; ++which) {
            PairTable &T = a->table[which];
            std::memset(&T, 0, sizeof(T));
            std::vector<double> blob;
            AppendPP1(blob, T.a, t->v_r.r, t->v_r.f, t->v_r.n);
            T.use_lr = a->use_long_range;
            T.is_coulomb = t->is_coulomb != 0;
            T.u_scale = ctx->tau;  // level 0: (1 >> 0) * tau
            if (a->use_long_range) {
                rc = LoadLongRange(ctx, a, which, t->v_long, blob, T.lr);
                if (rc != PIMC_OK) return rc;
            }
            rc = UploadBlob(ctx, a->blob[which], blob);
            if (rc != PIMC_OK) return rc;
        }
        rc = BuildFastV(ctx, a, t->v_r, t->is_coulomb != 0, a->use_long_range ? &t->v_long : nullptr);
        if (rc != PIMC_OK) return rc;
    } catch (const std::exception &e) {
        return Fail(PIMC_ERR_TABLE, e.what());
    }
    if (a->use_long_range) {
        ScaleConstants(ctx, a, a->k0[WHICH_DU], a->r0[WHICH_DU]);
        ScaleConstants(ctx, a, a->k0[WHICH_V], a->r0[WHICH_V]);
        a->ulong_scale = ctx->tau;  // bare...:170-171 at level 0
    }
    ctx->actions.push_back(guard.release());
    *out = a;
    return PIMC_OK;
}

int pimc_action_create_david(pimc_ctx *ctx, int32_t sa, int32_t sb, const pimc_david_tables *t, int32_t max_level,
                             int32_t use_lr, pimc_action **out) {
    if (!t) return Fail(PIMC_ERR_INVALID, "null tables");
    pimc_action *a = nullptr;
    int rc = NewAction(ctx, ATYPE_DAVID, sa, sb, max_level, use_lr, ctx ? ctx->k_cutoff : 0., &a);
    if (rc != PIMC_OK) return rc;
    std::unique_ptr<pimc_action> guard(a);
    PIMC_CUDA(cudaSetDevice(ctx->device));
    if (t->n_tau != 1) return Fail(PIMC_ERR_UNSUPPORTED, "DavidPairAction tables with more than one tau (max_level > 0)");
    // david...:232-245: the path's tau has to be in the table
    if (!t->taus || std::fabs(t->taus[0] - ctx->tau) >= 1.0e-6) return Fail(PIMC_ERR_TABLE, "tau of the path not found in the table");
    if (t->n_grid < 4 || !t->u_kj || !t->du_kj_dbeta || !t->potential) return Fail(PIMC_ERR_TABLE, "DavidPairAction table missing");
    try {
        const int n = t->n_grid;
        std::vector<double> grid(n);
        double ainv = 0., startinv = 0.;
        int code = GRIDCODE_GENERAL;
        if (t->grid_type == PIMC_GRID_LOG) {  // einspline create_log_grid
            const double aa = 1.0 / (double)(n - 1) * std::log(t->r_end / t->r_start);
            for (int i = 0; i < n; ++i) grid[i] = t->r_start * std::exp(aa * (double)i);
            ainv = 1.0 / aa;
            startinv = 1.0 / t->r_start;
            code = GRIDCODE_LOG;
        } else if (t->grid_type == PIMC_GRID_LINEAR) {
            for (int i = 0; i < n; ++i) grid[i] = t->r_start + (t->r_end - t->r_start) * (double)i / (double)(n - 1);
        } else {
            if (!t->grid_points) return Fail(PIMC_ERR_TABLE, "general grid without grid_points");
            std::copy(t->grid_points, t->grid_points + n, grid.begin());
        }
        int n_val = 1;  // david...:252-254
        for (int i = 1; i <= t->n_order; ++i) n_val += 1 + i;
        for (int which = 0; which < 3; ++which) {
            const double *data = (which == WHICH_DU) ? t->du_kj_dbeta : t->u_kj;
            // david...:262-283: value 0 of every knot is the potential; the last grid point of
            // every value stays 0 (the fill loop stops at n_grid-1)
            std::vector<std::vector<double>> values(n_val + 1, std::vector<double>(n, 0.));
            for (int v = 0; v < n_val + 1; ++v)
                for (int g = 0; g < n - 1; ++g) values[v][g] = (v == 0) ? t->potential[g] : data[(size_t)(v - 1) + (size_t)n_val * g];
            std::vector<double> blob = BuildBlobMulti(grid.data(), n, values);
            PairTable &T = a->table[which];
            std::memset(&T, 0, sizeof(T));
            Fill1D(T.dav, n, 0, grid.data());
            T.dav.n_splines = n_val + 1;
            T.dav.code = code;
            T.dav.ainv = ainv;
            T.dav.startinv = startinv;
            T.dav.r_min = (code == GRIDCODE_LOG) ? t->r_start : grid[0];
            T.dav.r_max = (code == GRIDCODE_LOG) ? t->r_end : grid[n - 1];
            T.n_order = t->n_order;
            T.use_lr = 0;  // David subtracts nothing in r space
            rc = UploadBlob(ctx, a->blob[which], blob);
            if (rc != PIMC_OK) return rc;
            T.dav_blob = a->blob[which].p;
            a->stageable[which] = false;
            if (which != WHICH_V) {
                rc = BuildFastDavid(ctx, a, which, grid.data(), n, values, t->n_order, T.dav.r_min, T.dav.r_max);
                if (rc != PIMC_OK) return rc;
                if (a->fastd_ok[which]) {  // every other kernel reads the same block from global memory
                    T.dav_fast = a->fastd[which];
                    T.dav_fast_tab = a->fastd_tab[which].p;
                    T.dav_use_fast = ctx->force_general ? 0 : 1;
                }
            } else {
                // CalcV = value 0 of the multi-spline at r and r' (david...:26-39): the fast Potential() kernel's v table
                pimc_table_1d v0;
                v0.n = n;
                v0.r = grid.data();
                v0.f = values[0].data();
                rc = BuildFastV(ctx, a, v0, 0, nullptr);
                if (rc != PIMC_OK) return rc;
                a->fastv.v.r_min = T.dav.r_min;  // the LOG grid clamps to its nominal start / end (GetLimits)
                a->fastv.v.r_max = T.dav.r_max;
            }
        }
    } catch (const std::exception &e) {
        return Fail(PIMC_ERR_TABLE, e.what());
    }
    if (a->use_long_range) {  // david...:311-358
        if (t->n_k < 1 || !t->k_points || !t->u_k) return Fail(PIMC_ERR_TABLE, "DavidPairAction long_range table missing");
        double v_long_k_0 = 0.;
        for (int which = 0; which < 3; ++which) {
            a->shell_k[which].assign(t->k_points, t->k_points + t->n_k);
            a->shell_f[which].resize(t->n_k);
        }
        for (int kv = 0; kv < t->n_k; ++kv) {
            const double vk = t->u_k[kv] / ctx->vol;
            if (std::fabs(0. - t->k_points[kv]) < 1.e-8) v_long_k_0 = vk;
            a->shell_f[WHICH_U][kv] = vk * ctx->tau;
            a->shell_f[WHICH_DU][kv] = vk;
            a->shell_f[WHICH_V][kv] = vk;  // the reference indexes shells by vector index here (UB); see DESIGN.md
        }
        for (int which = 0; which < 3; ++which)
            if ((rc = UploadVec(a->wk[which], MatchShells(ctx, a->shell_k[which], a->shell_f[which]))) != PIMC_OK) return rc;
        a->r0[WHICH_DU] = t->v_image;
        a->r0[WHICH_V] = t->v_image;
        a->k0[WHICH_DU] = v_long_k_0;
        a->k0[WHICH_V] = v_long_k_0;
        ScaleConstants(ctx, a, a->k0[WHICH_DU], a->r0[WHICH_DU]);
        ScaleConstants(ctx, a, a->k0[WHICH_V], a->r0[WHICH_V]);
    }
    ctx->actions.push_back(guard.release());
    *out = a;
    return PIMC_OK;
}

int pimc_action_create_kinetic(pimc_ctx *ctx, int32_t species, int32_t n_images, pimc_action **out) {
    if (!ctx || !out) return Fail(PIMC_ERR_INVALID, "null context or output");
    if (species < 0 || species >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    if (n_images < 0) return Fail(PIMC_ERR_INVALID, "negative n_images");
    SpeciesState &st = *ctx->species[species];
    if (st.kinetic) return Fail(PIMC_ERR_INVALID, "the species already has a Kinetic action");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    std::unique_ptr<pimc_action> a(new pimc_action);
    a->ctx = ctx;
    a->atype = ATYPE_KINETIC;
    a->sa = a->sb = species;
    a->n_images = n_images;
    FreeSet *fs = nullptr;  // Kinetic::SetupSpline runs in the constructor (kinetic_class.h:32)
    int rc = GetFreeSet(ctx, species, n_images, &fs);
    if (rc != PIMC_OK) return rc;
    st.kinetic = a.get();
    ctx->actions.push_back(a.get());
    *out = a.release();
    return PIMC_OK;
}

int pimc_move_set_images(pimc_ctx *ctx, int32_t species, int32_t n_images) {
    if (!ctx) return Fail(PIMC_ERR_INVALID, "null context");
    if (species < 0 || species >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    if (n_images < 0) return Fail(PIMC_ERR_INVALID, "negative n_images");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    FreeSet *fs = nullptr;
    int rc = GetFreeSet(ctx, species, n_images, &fs);
    if (rc != PIMC_OK) return rc;
    ctx->species[species]->move_images = n_images;
    return PIMC_OK;
}

int pimc_action_destroy(pimc_action *act) {
    if (!act) return PIMC_OK;
    pimc_ctx *ctx = act->ctx;
    if (act->atype == ATYPE_KINETIC && ctx->species[act->sa]->kinetic == act) ctx->species[act->sa]->kinetic = nullptr;
    auto it = std::find(ctx->actions.begin(), ctx->actions.end(), act);
    if (it != ctx->actions.end()) ctx->actions.erase(it);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    delete act;
    return PIMC_OK;
}

int pimc_action_dbeta_device(pimc_action *act, double *d_out) {
    if (!act || !d_out) return Fail(PIMC_ERR_INVALID, "null argument");
    PIMC_CUDA(cudaSetDevice(act->ctx->device));
    return FullEvaluation(act, WHICH_DU, d_out);
}
int pimc_action_potential_device(pimc_action *act, double *d_out) {
    if (!act || !d_out) return Fail(PIMC_ERR_INVALID, "null argument");
    PIMC_CUDA(cudaSetDevice(act->ctx->device));
    return FullEvaluation(act, WHICH_V, d_out);
}
int pimc_action_total_device(pimc_action *act, double *d_out) {
    if (!act || !d_out) return Fail(PIMC_ERR_INVALID, "null argument");
    PIMC_CUDA(cudaSetDevice(act->ctx->device));
    return FullEvaluation(act, WHICH_U, d_out);
}
static int FullToHost(pimc_action *act, int which, double *out) {
    if (!act || !out) return Fail(PIMC_ERR_INVALID, "null argument");
    pimc_ctx *ctx = act->ctx;
    PIMC_CUDA(cudaSetDevice(ctx->device));
    int rc = EnsureOut(ctx);
    if (rc != PIMC_OK) return rc;
    rc = FullEvaluation(act, which, ctx->out_dev.p);
    if (rc != PIMC_OK) return rc;
    return ToHost(ctx, ctx->out_dev.p, out, ctx->C);
}
int pimc_action_dbeta(pimc_action *act, double *out) { return FullToHost(act, WHICH_DU, out); }
int pimc_action_potential(pimc_action *act, double *out) { return FullToHost(act, WHICH_V, out); }
int pimc_action_total(pimc_action *act, double *out) { return FullToHost(act, WHICH_U, out); }

int pimc_action_get(pimc_action *act, int32_t mode, const int32_t *b0, int32_t n_window, int32_t n_moved,
                    const int32_t *moved_species, const int32_t *moved_particle, int32_t level, double *out) {
    if (!act || !out || !b0 || (n_moved > 0 && (!moved_species || !moved_particle))) return Fail(PIMC_ERR_INVALID, "null argument");
    pimc_ctx *ctx = act->ctx;
    PIMC_CUDA(cudaSetDevice(ctx->device));
    if (act->atype == ATYPE_KINETIC) {  // Kinetic::GetAction: every level, only the listed particles of its species
        if (ctx->sharded) return Fail(PIMC_ERR_UNSUPPORTED, "move windows on a slice-sharded context");
        if (n_window < 1 || n_window > ctx->M) return Fail(PIMC_ERR_INVALID, "window must cover 1..n_bead slices");
        if (level < 0) return Fail(PIMC_ERR_INVALID, "negative level");
        std::vector<int> idx;
        for (int i = 0; i < n_moved; ++i)
            if (moved_species[i] == act->sa) idx.push_back(i);
        if (idx.empty()) {
            for (int c = 0; c < ctx->C; ++c) out[c] = 0.;
            return PIMC_OK;
        }
        if ((int)idx.size() > kMaxPropSlots) return Fail(PIMC_ERR_UNSUPPORTED, "more than 16 listed particles of one species");
        {  // the window kernel reads a particle's beads by label: exact on an unpermuted path only
            const int rc_p = RequireUnpermuted(ctx, act->sa, "Kinetic::GetAction over a window");
            if (rc_p != PIMC_OK) return rc_p;
        }
        if (n_window == ctx->M) {  // bead_a->GetNextBead(b1 - b0) is bead_a itself: the reference's loop body never runs (kinetic_class.h:112-113)
            for (int c = 0; c < ctx->C; ++c) out[c] = 0.;
            return PIMC_OK;
        }
        std::vector<int32_t> part(idx.size() * (size_t)ctx->C);
        for (int c = 0; c < ctx->C; ++c) {
            if (b0[c] < 0 || b0[c] >= ctx->M) return Fail(PIMC_ERR_INVALID, "window start out of range");
            for (size_t i = 0; i < idx.size(); ++i) {
                const int32_t p = moved_particle[(size_t)c * n_moved + idx[i]];
                if (p < 0 || p >= ctx->species[act->sa]->N) return Fail(PIMC_ERR_INVALID, "moved particle out of range");
                part[i * ctx->C + c] = p;
            }
        }
        int rc;
        if ((rc = EnsureI32(ctx, ctx->i32_a, part.data(), part.size())) != PIMC_OK) return rc;
        if ((rc = EnsureI32(ctx, ctx->i32_c, b0, ctx->C)) != PIMC_OK) return rc;
        if ((rc = EnsureOut(ctx)) != PIMC_OK) return rc;
        if ((rc = KineticWindow(act, mode, ctx->i32_c.p, n_window, (int)idx.size(), ctx->i32_a.p, level, ctx->out_dev.p)) != PIMC_OK) return rc;
        return ToHost(ctx, ctx->out_dev.p, out, ctx->C);
    }
    // pair_action_class.h:269
    if (level > act->max_level || act->is_constant) {
        for (int c = 0; c < ctx->C; ++c) out[c] = 0.;
        return PIMC_OK;
    }
    if (ctx->sharded) return Fail(PIMC_ERR_UNSUPPORTED, "move windows on a slice-sharded context");
    if (n_window < 1 || n_window > ctx->M) return Fail(PIMC_ERR_INVALID, "window must cover 1..n_bead slices");
    // the listed particles of the action's two species (GenerateParticlePairs, pair_action_class.h:63-76):
    // one for Bisect / DisplaceParticle, the particles of a cycle for permutation moves
    std::vector<int> ia, ib;
    for (int i = 0; i < n_moved; ++i) {
        if (moved_species[i] == act->sa)
            ia.push_back(i);
        else if (moved_species[i] == act->sb)
            ib.push_back(i);
    }
    if (act->sa == act->sb) ib.clear();
    if ((int)ia.size() > kMaxPropSlots || (int)ib.size() > kMaxPropSlots)
        return Fail(PIMC_ERR_UNSUPPORTED, "more than 16 listed particles of one species");
    if (ia.empty() && ib.empty()) {  // pair_action_class.h:77-78
        for (int c = 0; c < ctx->C; ++c) out[c] = 0.;
        return PIMC_OK;
    }
    const int na = (int)ia.size(), nb_l = (int)ib.size();
    std::vector<int32_t> pa((size_t)std::max(na, 1) * ctx->C, 0), pb((size_t)std::max(nb_l, 1) * ctx->C, 0);
    for (int c = 0; c < ctx->C; ++c) {
        if (b0[c] < 0 || b0[c] >= ctx->M) return Fail(PIMC_ERR_INVALID, "window start out of range");
        for (int i = 0; i < na; ++i) {
            const int32_t p = moved_particle[(size_t)c * n_moved + ia[i]];
            if (p < 0 || p >= ctx->species[act->sa]->N) return Fail(PIMC_ERR_INVALID, "moved particle out of range");
            for (int i2 = 0; i2 < i; ++i2)
                if (pa[(size_t)i2 * ctx->C + c] == p) return Fail(PIMC_ERR_INVALID, "particle listed twice");
            pa[(size_t)i * ctx->C + c] = p;
        }
        for (int i = 0; i < nb_l; ++i) {
            const int32_t p = moved_particle[(size_t)c * n_moved + ib[i]];
            if (p < 0 || p >= ctx->species[act->sb]->N) return Fail(PIMC_ERR_INVALID, "moved particle out of range");
            for (int i2 = 0; i2 < i; ++i2)
                if (pb[(size_t)i2 * ctx->C + c] == p) return Fail(PIMC_ERR_INVALID, "particle listed twice");
            pb[(size_t)i * ctx->C + c] = p;
        }
    }
    int rc;
    if ((rc = EnsureI32(ctx, ctx->i32_a, pa.data(), pa.size())) != PIMC_OK) return rc;
    if ((rc = EnsureI32(ctx, ctx->i32_b, pb.data(), pb.size())) != PIMC_OK) return rc;
    if ((rc = EnsureI32(ctx, ctx->i32_c, b0, ctx->C)) != PIMC_OK) return rc;
    PairWindowArgs w;
    w.pv = ctx->View();
    w.A = ctx->SView(act->sa, mode == PIMC_NEW);
    w.B = ctx->SView(act->sb, mode == PIMC_NEW);
    w.same = act->sa == act->sb;
    w.n_a = na;
    w.n_b = nb_l;
    w.part_a = ctx->i32_a.p;
    w.part_b = ctx->i32_b.p;
    w.b0 = ctx->i32_c.p;
    w.n_links = n_window;
    w.mode = mode;
    w.T = act->table[WHICH_U];
    w.blob = act->blob[WHICH_U].p;
    const size_t items = (size_t)ctx->C * n_window;
    if (ctx->partial.n < items) PIMC_CUDA(ctx->partial.Alloc(items));
    w.partial = ctx->partial.p;
    const int grid = (int)std::min<size_t>(items, (size_t)ctx->n_sm * 16);
    {
        ScopedKernelTimer t(ctx, PIMC_KERNEL_PAIR_WINDOW);
        switch (act->atype) {
            case ATYPE_ILKKA: pair_window_kernel<ATYPE_ILKKA><<<grid, 128, 0, ctx->stream>>>(w); break;
            case ATYPE_BARE: pair_window_kernel<ATYPE_BARE><<<grid, 128, 0, ctx->stream>>>(w); break;
            default: pair_window_kernel<ATYPE_DAVID><<<grid, 128, 0, ctx->stream>>>(w); break;
        }
    }
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    const bool lr = act->use_long_range && ctx->n_k() > 0;
    if (lr) {
        // pair_action_class.h:293-299: refresh the proposal's rho_k once per move per species
        if (mode == PIMC_NEW) {
            const int sp[2] = {act->sa, act->sb};
            const int idx[2] = {ia.empty() ? -1 : 0, ib.empty() ? -1 : 0};
            for (int t = 0; t < (act->sa == act->sb ? 1 : 2); ++t) {
                SpeciesState &st = *ctx->species[sp[t]];
                if (!st.need_update_rho_k) continue;
                st.drho_valid = false;
                if (idx[t] >= 0 && st.n_prop > 0) {
                    const size_t need = (size_t)ctx->C * n_window * ctx->n_k();
                    if (st.drho.n < need) PIMC_CUDA(st.drho.Alloc(need));
                    PIMC_CUDA(cudaMemcpyAsync(st.drho_b0.p, ctx->i32_c.p, ctx->C * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
                    const int tl = 2 * ctx->max_index + 1;
                    rhok_delta_kernel<<<GridFor(ctx, ctx->C * n_window), 256, (size_t)6 * tl * sizeof(double2), ctx->stream>>>(
                        ctx->View(), ctx->SView(sp[t], true), ctx->KView(), st.drho_b0.p, n_window, st.drho.p);
                    ctx->launches++;
                    PIMC_CUDA(cudaGetLastError());
                    st.drho_window = n_window;
                    st.drho_valid = true;
                }
                st.need_update_rho_k = false;
            }
            for (int t = 0; t < 2; ++t) {
                SpeciesState &st = *ctx->species[sp[t]];
                if (st.drho_valid && st.drho_window != n_window)
                    return Fail(PIMC_ERR_INVALID, "NEW-mode GetAction window differs from the window rho_k was refreshed for");
            }
        }
        rc = LaunchKSum(act, WHICH_U, ctx->i32_c.p, n_window, mode == PIMC_NEW, act->ulong_scale);
        if (rc != PIMC_OK) return rc;
    }
    if ((rc = EnsureOut(ctx)) != PIMC_OK) return rc;
    if (act->use_long_range && !lr) {
        if ((int)ctx->lr_dev.n < ctx->C) PIMC_CUDA(ctx->lr_dev.Alloc(ctx->C));
        PIMC_CUDA(cudaMemsetAsync(ctx->lr_dev.p, 0, ctx->C * sizeof(double), ctx->stream));
    }
    rc = Finalize(ctx, n_window, act->use_long_range, 0., 0., false, ctx->out_dev.p);
    if (rc != PIMC_OK) return rc;
    return ToHost(ctx, ctx->out_dev.p, out, ctx->C);
}

// --------------------------------------------------------- spatial derivatives of the action
/// what = 0: GetActionGradient -> out[C][3]; what = 1: GetActionLaplacian -> out[C].
static int ActionDerivative(pimc_action *act, const int32_t *b0, int32_t n_window, int32_t n_moved, const int32_t *moved_species,
                            const int32_t *moved_particle, int32_t level, int what, double *out) {
    if (!act || !out || !b0 || (n_moved > 0 && (!moved_species || !moved_particle))) return Fail(PIMC_ERR_INVALID, "null argument");
    pimc_ctx *ctx = act->ctx;
    PIMC_CUDA(cudaSetDevice(ctx->device));
    if (act->atype == ATYPE_KINETIC) return Fail(PIMC_ERR_UNSUPPORTED, "gradient / Laplacian of the Kinetic action (closed forms, kinetic_class.h:125-165) are not on the device path");
    const int nv = what ? 1 : 3;
    auto zero = [&]() {
        for (size_t i = 0; i < (size_t)ctx->C * nv; ++i) out[i] = 0.;
        return PIMC_OK;
    };
    if (level > act->max_level || act->is_constant) return zero();  // pair_action_class.h:308,342
    if (ctx->sharded) return Fail(PIMC_ERR_UNSUPPORTED, "action derivatives on a slice-sharded context");
    if (n_window < 1 || n_window > ctx->M) return Fail(PIMC_ERR_INVALID, "window must cover 1..n_bead slices");
    if (ctx->species[act->sa]->n_prop > 0 || ctx->species[act->sb]->n_prop > 0)
        return Fail(PIMC_ERR_UNSUPPORTED, "action derivatives while a proposal is pending (estimators run between moves)");
    int ia = -1, ib = -1;
    for (int i = 0; i < n_moved; ++i) {
        if (moved_species[i] == act->sa) {
            if (ia >= 0) return Fail(PIMC_ERR_UNSUPPORTED, "two listed particles of one species");
            ia = i;
        } else if (moved_species[i] == act->sb) {
            if (ib >= 0) return Fail(PIMC_ERR_UNSUPPORTED, "two listed particles of one species");
            ib = i;
        }
    }
    if (act->sa == act->sb) ib = -1;
    if (ia < 0 && ib < 0) return zero();
    const int Na = ctx->species[act->sa]->N, Nb = ctx->species[act->sb]->N;
    std::vector<int32_t> pa(ctx->C, 0), pb(ctx->C, 0);
    for (int c = 0; c < ctx->C; ++c) {
        if (ia >= 0) pa[c] = moved_particle[(size_t)c * n_moved + ia];
        if (ib >= 0) pb[c] = moved_particle[(size_t)c * n_moved + ib];
        if (b0[c] < 0 || b0[c] >= ctx->M) return Fail(PIMC_ERR_INVALID, "window start out of range");
        if (ia >= 0 && (pa[c] < 0 || pa[c] >= Na)) return Fail(PIMC_ERR_INVALID, "listed particle out of range");
        if (ib >= 0 && (pb[c] < 0 || pb[c] >= Nb)) return Fail(PIMC_ERR_INVALID, "listed particle out of range");
    }
    int rc;
    if ((rc = EnsureI32(ctx, ctx->i32_a, pa.data(), ctx->C)) != PIMC_OK) return rc;
    if ((rc = EnsureI32(ctx, ctx->i32_b, pb.data(), ctx->C)) != PIMC_OK) return rc;
    if ((rc = EnsureI32(ctx, ctx->i32_c, b0, ctx->C)) != PIMC_OK) return rc;
    PairGradArgs w;
    w.pv = ctx->View();
    w.A = ctx->SView(act->sa, false);
    w.B = ctx->SView(act->sb, false);
    w.same = act->sa == act->sb;
    w.moved_a = ia >= 0;
    w.moved_b = ib >= 0;
    w.part_a = ctx->i32_a.p;
    w.part_b = ctx->i32_b.p;
    w.b0 = ctx->i32_c.p;
    w.n_links = n_window;
    w.what = what;
    w.T = act->table[WHICH_U];
    w.blob = act->blob[WHICH_U].p;
    const size_t items = (size_t)ctx->C * n_window;
    if (ctx->partial.n < items * nv) PIMC_CUDA(ctx->partial.Alloc(items * nv));
    w.partial = ctx->partial.p;
    const int grid = (int)std::min<size_t>(items, (size_t)ctx->n_sm * 16);
    switch (act->atype) {
        case ATYPE_ILKKA: pair_grad_kernel<ATYPE_ILKKA><<<grid, 128, 0, ctx->stream>>>(w); break;
        case ATYPE_BARE: pair_grad_kernel<ATYPE_BARE><<<grid, 128, 0, ctx->stream>>>(w); break;
        default: pair_grad_kernel<ATYPE_DAVID><<<grid, 128, 0, ctx->stream>>>(w); break;
    }
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    // long-range part of the gradient: Ilkka only (the base class returns zero, pair_action_class.h:163-165)
    const bool lr = what == 0 && act->atype == ATYPE_ILKKA && act->use_long_range && ctx->n_k() > 0;
    if ((size_t)ctx->lr_dev.n < (size_t)ctx->C * 3) PIMC_CUDA(ctx->lr_dev.Alloc((size_t)ctx->C * 3));
    if (lr) {
        GradLongArgs g;
        g.pv = w.pv;
        g.A = w.A;
        g.ks = ctx->KView();
        g.rho_b = ctx->species[act->sb]->rho.p;
        g.wk = act->wk[WHICH_U].p;
        g.part_a = ia >= 0 ? ctx->i32_a.p : nullptr;
        g.b0 = ctx->i32_c.p;
        g.n_links = n_window;
        g.mult_moved = w.same ? Na - 1 : Nb;
        g.mult_others = ib >= 0 ? 1 : 0;
        g.factor = w.same ? 1. : 2.;
        g.out = ctx->lr_dev.p;
        const size_t smem = (size_t)32 * 3 * (2 * ctx->max_index + 1) * sizeof(double2);
        if (smem > 48 * 1024) PIMC_CUDA(cudaFuncSetAttribute(grad_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        grad_long_kernel<<<ctx->C, 256, smem, ctx->stream>>>(g);
        ctx->launches++;
        PIMC_CUDA(cudaGetLastError());
    }
    if ((size_t)ctx->out_dev.n < (size_t)ctx->C * 3) PIMC_CUDA(ctx->out_dev.Alloc((size_t)ctx->C * 3));
    grad_finalize_kernel<<<(ctx->C * nv + 127) / 128, 128, 0, ctx->stream>>>(ctx->partial.p, ctx->C, n_window, nv, lr ? ctx->lr_dev.p : nullptr,
                                                                              ctx->out_dev.p);
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return ToHost(ctx, ctx->out_dev.p, out, (size_t)ctx->C * nv);
}

int pimc_action_gradient(pimc_action *act, const int32_t *b0, int32_t n_window, int32_t n_moved, const int32_t *moved_species,
                         const int32_t *moved_particle, int32_t level, double *grad) {
    return ActionDerivative(act, b0, n_window, n_moved, moved_species, moved_particle, level, 0, grad);
}

int pimc_action_laplacian(pimc_action *act, const int32_t *b0, int32_t n_window, int32_t n_moved, const int32_t *moved_species,
                          const int32_t *moved_particle, int32_t level, double *lap) {
    return ActionDerivative(act, b0, n_window, n_moved, moved_species, moved_particle, level, 1, lap);
}

static int ArmFlags(pimc_action *act) {
    if (!act) return Fail(PIMC_ERR_INVALID, "null action");
    if (act->use_long_range) {
        act->ctx->species[act->sa]->need_update_rho_k = true;
        act->ctx->species[act->sb]->need_update_rho_k = true;
    }
    return PIMC_OK;
}
int pimc_action_accept(pimc_action *act) { return ArmFlags(act); }
int pimc_action_reject(pimc_action *act) { return ArmFlags(act); }

int pimc_action_calc_pair(pimc_action *act, int32_t which, int32_t n, const double *r, const double *r_p, const double *s,
                          int32_t level, double *out) {
    if (!act || !r || !r_p || !s || !out || n < 0 || which < 0 || which > 2) return Fail(PIMC_ERR_INVALID, "bad argument");
    if (level != 0) return Fail(PIMC_ERR_UNSUPPORTED, "level > 0");
    if (act->atype == ATYPE_KINETIC) return Fail(PIMC_ERR_INVALID, "a Kinetic action has no pair kernels");
    if (n == 0) return PIMC_OK;
    pimc_ctx *ctx = act->ctx;
    PIMC_CUDA(cudaSetDevice(ctx->device));
    DevBuf<double> buf;
    PIMC_CUDA(buf.Alloc((size_t)4 * n));
    PIMC_CUDA(cudaMemcpyAsync(buf.p, r, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PIMC_CUDA(cudaMemcpyAsync(buf.p + n, r_p, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PIMC_CUDA(cudaMemcpyAsync(buf.p + 2 * (size_t)n, s, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const int grid = (n + 127) / 128;
    const double *blob = act->blob[which].p;
    const PairTable &T = act->table[which];
#define PIMC_CP(AT, WH) calc_pair_kernel<AT, WH><<<grid, 128, 0, ctx->stream>>>(blob, T, n, buf.p, buf.p + n, buf.p + 2 * (size_t)n, buf.p + 3 * (size_t)n)
#define PIMC_CPW(AT)                      \
    if (which == WHICH_U)                 \
        PIMC_CP(AT, WHICH_U);             \
    else if (which == WHICH_DU)           \
        PIMC_CP(AT, WHICH_DU);            \
    else                                  \
        PIMC_CP(AT, WHICH_V);
    if (act->atype == ATYPE_ILKKA) {
        PIMC_CPW(ATYPE_ILKKA)
    } else if (act->atype == ATYPE_BARE) {
        PIMC_CPW(ATYPE_BARE)
    } else {
        PIMC_CPW(ATYPE_DAVID)
    }
#undef PIMC_CPW
#undef PIMC_CP
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return ToHost(ctx, buf.p + 3 * (size_t)n, out, n);
}

// ---------------------------------------------------------------------------------- moves
int pimc_propose(pimc_ctx *ctx, int32_t s, const int32_t *particle, const int32_t *b_first, int32_t n_beads, const double *newR) {
    if (!ctx || !particle || !b_first || !newR) return Fail(PIMC_ERR_INVALID, "null argument");
    if (s < 0 || s >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    if (n_beads < 1 || n_beads > ctx->M) return Fail(PIMC_ERR_INVALID, "proposal must cover 1..n_bead beads");
    if (ctx->sharded) return Fail(PIMC_ERR_UNSUPPORTED, "proposals on a slice-sharded context");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    SpeciesState &st = *ctx->species[s];
    for (int c = 0; c < ctx->C; ++c) {
        if (particle[c] < 0 || particle[c] >= st.N) return Fail(PIMC_ERR_INVALID, "proposal particle out of range");
        if (b_first[c] < 0 || b_first[c] >= ctx->M) return Fail(PIMC_ERR_INVALID, "proposal bead out of range");
    }
    // a second proposal on a species that already has one pending ADDS a particle (the cycle of a
    // permutation move): same bead count, another particle in every clone
    int slot = 0;
    if (st.n_prop > 0 && st.n_slots > 0) {
        if (n_beads != st.n_prop) return Fail(PIMC_ERR_INVALID, "proposals of one species must cover the same number of beads");
        if (st.n_slots == kMaxPropSlots) return Fail(PIMC_ERR_UNSUPPORTED, "more than 16 proposed particles of one species");
        for (int sl = 0; sl < st.n_slots; ++sl)
            for (int c = 0; c < ctx->C; ++c)
                if (st.h_particle[(size_t)sl * ctx->C + c] == particle[c]) return Fail(PIMC_ERR_INVALID, "particle proposed twice");
        slot = st.n_slots;
    }
    const size_t n = (size_t)ctx->C * n_beads * 3;
    if (st.P.n < n * kMaxPropSlots) {
        if (slot > 0) return Fail(PIMC_ERR_INVALID, "internal: proposal buffer smaller than its slots");
        PIMC_CUDA(st.P.Alloc(n * kMaxPropSlots));
    }
    PIMC_CUDA(cudaMemcpyAsync(st.P.p + slot * n, newR, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PIMC_CUDA(cudaMemcpyAsync(st.P_particle.p + (size_t)slot * ctx->C, particle, ctx->C * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    PIMC_CUDA(cudaMemcpyAsync(st.P_first.p + (size_t)slot * ctx->C, b_first, ctx->C * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    PIMC_CUDA(cudaStreamSynchronize(ctx->stream));  // host buffers may be reused by the caller
    st.h_particle.resize((size_t)kMaxPropSlots * ctx->C);
    std::copy(particle, particle + ctx->C, st.h_particle.begin() + (size_t)slot * ctx->C);
    st.n_prop = n_beads;
    st.n_slots = slot + 1;
    st.drho_valid = false;
    st.need_update_rho_k = true;  // the refreshed rho_k (if any) did not include this particle
    return PIMC_OK;
}

int pimc_beads_download(pimc_ctx *ctx, int32_t s, const int32_t *particle, const int32_t *b_first, int32_t n_beads, double *out) {
    if (!ctx || !particle || !b_first || !out) return Fail(PIMC_ERR_INVALID, "null argument");
    if (s < 0 || s >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    if (n_beads < 1 || n_beads > 2 * ctx->M) return Fail(PIMC_ERR_INVALID, "bad bead count");
    if (ctx->sharded) return Fail(PIMC_ERR_UNSUPPORTED, "bead windows on a slice-sharded context");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    SpeciesState &st = *ctx->species[s];
    for (int c = 0; c < ctx->C; ++c) {
        if (particle[c] < 0 || particle[c] >= st.N) return Fail(PIMC_ERR_INVALID, "particle out of range");
        if (b_first[c] < 0 || b_first[c] >= ctx->M) return Fail(PIMC_ERR_INVALID, "bead out of range");
    }
    int rc;
    if ((rc = EnsureI32(ctx, ctx->i32_a, particle, ctx->C)) != PIMC_OK) return rc;
    if ((rc = EnsureI32(ctx, ctx->i32_b, b_first, ctx->C)) != PIMC_OK) return rc;
    const size_t n = (size_t)ctx->C * n_beads * 3;
    if (ctx->stage.n < n) PIMC_CUDA(ctx->stage.Alloc(n));
    gather_beads_kernel<<<ctx->C, 64, 0, ctx->stream>>>(ctx->View(), st.R.p, st.N, ctx->i32_a.p, ctx->i32_b.p, n_beads, ctx->stage.p);
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return ToHost(ctx, ctx->stage.p, out, n);
}

int pimc_commit(pimc_ctx *ctx, const int32_t *accept) {
    if (!ctx || !accept) return Fail(PIMC_ERR_INVALID, "null argument");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = EnsureI32(ctx, ctx->i32_d, accept, ctx->C)) != PIMC_OK) return rc;
    for (auto &sp : ctx->species) {
        SpeciesState &st = *sp;
        if (st.n_prop > 0) {
            commit_positions_kernel<<<ctx->C, 64, 0, ctx->stream>>>(ctx->View(), st.N, st.P.p, st.P_particle.p, st.P_first.p, st.n_prop,
                                                                  std::max(1, st.n_slots), ctx->i32_d.p, st.R.p);
            st.R2_valid = false;
            ctx->launches++;
            PIMC_CUDA(cudaGetLastError());
            if (st.drho_valid) {
                const int n_k = ctx->n_k();
                dim3 grid((st.drho_window * n_k + 255) / 256, ctx->C);
                commit_rhok_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->View(), n_k, st.drho.p, st.drho_b0.p, st.drho_window, ctx->i32_d.p,
                                                                  st.rho.p);
                ctx->launches++;
                PIMC_CUDA(cudaGetLastError());
            }
            // If no long-range action evaluated this proposal in NEW mode, rho_k is NOT refreshed:
            // the reference only updates it inside PairAction::GetAction
            // (pair_action_class.h:293-299), so StoreRhoK commits the old values there too.
        }
        st.n_prop = 0;
        st.n_slots = 0;
        st.drho_valid = false;
        st.need_update_rho_k = true;  // every action's Accept/Reject re-arms the flag
    }
    PIMC_CUDA(cudaStreamSynchronize(ctx->stream));
    return PIMC_OK;
}


// ----------------------------------------------------------------------------- estimators
int pimc_est_gofr_counts(pimc_ctx *ctx, int32_t sa, int32_t sb, double r_min, double r_max, int32_t n_r, uint64_t *counts) {
    if (!ctx || !counts) return Fail(PIMC_ERR_INVALID, "null argument");
    if (sa < 0 || sb < 0 || sa >= (int)ctx->species.size() || sb >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    if (n_r < 2 || n_r > 8192) return Fail(PIMC_ERR_INVALID, "n_r must be in 2..8192");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->C * n_r;
    if (ctx->counts.n < n) PIMC_CUDA(ctx->counts.Alloc(n));
    PIMC_CUDA(cudaMemsetAsync(ctx->counts.p, 0, n * sizeof(unsigned long long), ctx->stream));
    GofrArgs g;
    g.pv = ctx->View();
    g.A = ctx->SView(sa, false);
    g.B = ctx->SView(sb, false);
    g.same = sa == sb;
    // LinearGrid::CreateGrid (observable_class.h:30-40)
    const double dr = (r_max - r_min) / (n_r - 1.);
    g.r_min = r_min;
    g.d_ir = 1. / dr;
    g.n_r = n_r;
    g.counts = ctx->counts.p;
    // tiled kernel (K5 v2): the largest tile whose two copies leave room for the histograms
    const int Na = g.A.N, Nb = g.B.N;
    GofrTiledArgs t;
    t.g = g;
    t.T = 0;
    t.warp_hist = 0;
    for (int pass = 0; pass < 2 && t.T == 0; ++pass)  // per-warp histograms if they fit beside some tile size, else one per CTA
        for (int T : {128, 64, 32}) {
            const size_t need = (size_t)2 * T * 3 * kGofrRow * sizeof(double) + (size_t)(pass == 0 ? kGofrThreads / 32 : 1) * n_r * sizeof(unsigned int) + 1024;
            if (need > ctx->smem_optin) continue;
            t.T = T;
            t.warp_hist = pass == 0 ? 1 : 0;
            break;
        }
    if (t.T > 0 && !ctx->force_general) {
        t.T = std::min(t.T, 32 * ((std::max(Na, Nb) + 31) / 32));
        t.n_ti = (Na + t.T - 1) / t.T;
        t.n_tj = (Nb + t.T - 1) / t.T;
        t.n_chunks = (ctx->Mloc + 31) / 32;
        const size_t smem = (size_t)2 * t.T * 3 * kGofrRow * sizeof(double) + (size_t)(t.warp_hist ? kGofrThreads / 32 : 1) * n_r * sizeof(unsigned int);
        const size_t n_items = (size_t)ctx->C * t.n_chunks * (g.same ? (size_t)t.n_ti * (t.n_ti + 1) / 2 : (size_t)t.n_ti * t.n_tj);
        PIMC_CUDA(cudaFuncSetAttribute(gofr_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ScopedKernelTimer tm(ctx, PIMC_KERNEL_GOFR);
        gofr_tiled_kernel<<<(int)std::min<size_t>(n_items, (size_t)ctx->n_sm), kGofrThreads, smem, ctx->stream>>>(t);
    } else {
        const int items = ctx->C * ctx->Mloc;
        ScopedKernelTimer tm(ctx, PIMC_KERNEL_GOFR);
        gofr_kernel<<<GridFor(ctx, items), 256, n_r * sizeof(unsigned int), ctx->stream>>>(g);
    }
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    PIMC_CUDA(cudaMemcpyAsync(counts, ctx->counts.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    PIMC_CUDA(cudaStreamSynchronize(ctx->stream));
    return PIMC_OK;
}

int pimc_est_gofr(pimc_ctx *ctx, int32_t sa, int32_t sb, double r_min, double r_max, int32_t n_r, const double *cofactor, double *y) {
    if (!ctx || !y) return Fail(PIMC_ERR_INVALID, "null argument");
    std::vector<uint64_t> counts((size_t)ctx->C * std::max(n_r, 0));
    int rc = pimc_est_gofr_counts(ctx, sa, sb, r_min, r_max, n_r, counts.data());
    if (rc != PIMC_OK) return rc;
    // gr.y(i) = gr.y(i) + 1.*cofactor, once per counted pair (pair_correlation_class.h:23)
    for (int c = 0; c < ctx->C; ++c) {
        const double cf = cofactor ? cofactor[c] : 1.0;
        for (int i = 0; i < n_r; ++i) y[(size_t)c * n_r + i] += cf * (double)counts[(size_t)c * n_r + i];
    }
    return PIMC_OK;
}

int pimc_est_sofk(pimc_ctx *ctx, int32_t sa, int32_t sb, double k_cut, const double *cofactor, double *sk) {
    if (!ctx || !sk) return Fail(PIMC_ERR_INVALID, "null argument");
    if (sa < 0 || sb < 0 || sa >= (int)ctx->species.size() || sb >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    if (k_cut > ctx->k_cutoff) {  // structure_factor_class.h:49-51
        int32_t n_k;
        int rc = pimc_kspace_setup(ctx, k_cut, &n_k);
        if (rc != PIMC_OK) return rc;
    }
    const int n_k = ctx->n_k();
    if (n_k == 0) return PIMC_OK;
    const size_t n = (size_t)ctx->C * n_k;
    if (ctx->est.n < n + ctx->C) PIMC_CUDA(ctx->est.Alloc(n + ctx->C));
    PIMC_CUDA(cudaMemcpyAsync(ctx->est.p, sk, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    double *d_cf = nullptr;
    if (cofactor) {
        d_cf = ctx->est.p + n;
        PIMC_CUDA(cudaMemcpyAsync(d_cf, cofactor, ctx->C * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    dim3 grid((n_k + 127) / 128, ctx->C);
    {
        ScopedKernelTimer t(ctx, PIMC_KERNEL_SOFK);
        sofk_kernel<<<grid, 128, 0, ctx->stream>>>(ctx->View(), n_k, ctx->species[sa]->rho.p, ctx->species[sb]->rho.p, ctx->d_kmag.p, k_cut,
                                                   d_cf, ctx->est.p);
    }
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return ToHost(ctx, ctx->est.p, sk, n);
}

int pimc_action_calc_pair_fast(pimc_action *act, int32_t which, int32_t n, const double *r, const double *r_p, const double *s,
                               double *out) {
    if (!act || !r || !r_p || !s || !out || n < 0 || which < 0 || which > 1) return Fail(PIMC_ERR_INVALID, "bad argument");
    const bool david = act->atype == ATYPE_DAVID && act->fastd_ok[which];
    if (!david && (act->atype != ATYPE_ILKKA || !act->fast_ok[which])) return Fail(PIMC_ERR_UNSUPPORTED, "action has no fast-path tables");
    if (n == 0) return PIMC_OK;
    pimc_ctx *ctx = act->ctx;
    PIMC_CUDA(cudaSetDevice(ctx->device));
    DevBuf<double> buf;
    PIMC_CUDA(buf.Alloc((size_t)4 * n));
    PIMC_CUDA(cudaMemcpyAsync(buf.p, r, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PIMC_CUDA(cudaMemcpyAsync(buf.p + n, r_p, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PIMC_CUDA(cudaMemcpyAsync(buf.p + 2 * (size_t)n, s, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (david) {
        const FastDavidTable &T = act->fastd[which];
        const unsigned char *tab = act->fastd_tab[which].p;
        double *pr = buf.p, *pp = buf.p + n, *ps = buf.p + 2 * (size_t)n, *po = buf.p + 3 * (size_t)n;
        const int grid = (n + 127) / 128;
#define PIMC_DAVID_HOOK(NORD, KIND) calc_david_fast_kernel<NORD, KIND><<<grid, 128, 0, ctx->stream>>>(tab, T, n, pr, pp, ps, po)
        switch (T.n_order * 2 + T.lut.kind) {
            case 2: PIMC_DAVID_HOOK(1, 0); break;
            case 3: PIMC_DAVID_HOOK(1, 1); break;
            case 4: PIMC_DAVID_HOOK(2, 0); break;
            case 5: PIMC_DAVID_HOOK(2, 1); break;
            case 6: PIMC_DAVID_HOOK(3, 0); break;
            default: PIMC_DAVID_HOOK(3, 1); break;
        }
#undef PIMC_DAVID_HOOK
    } else {
        calc_pair_fast_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(act->fast_tab[which].p, act->fast[which], n, buf.p, buf.p + n,
                                                                        buf.p + 2 * (size_t)n, buf.p + 3 * (size_t)n);
    }
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return ToHost(ctx, buf.p + 3 * (size_t)n, out, n);
}

int pimc_debug_fast_sqrt(pimc_ctx *ctx, int32_t n, const double *x, double *out) {
    if (!ctx || !x || !out || n < 0) return Fail(PIMC_ERR_INVALID, "bad argument");
    if (n == 0) return PIMC_OK;
    PIMC_CUDA(cudaSetDevice(ctx->device));
    DevBuf<double> buf;
    PIMC_CUDA(buf.Alloc((size_t)2 * n));
    PIMC_CUDA(cudaMemcpyAsync(buf.p, x, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    fast_sqrt_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, buf.p, buf.p + n);
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return ToHost(ctx, buf.p + n, out, n);
}

int pimc_debug_interval_table(int32_t kind, int32_t n, const double *grid, int32_t m, const double *x, int32_t *out, int32_t *n_keys) {
    if (!grid || !x || !out || n < 2 || m < 0 || kind < 0 || kind > 1) return Fail(PIMC_ERR_INVALID, "bad argument");
    const int kMaxKeys = 16384;
    std::vector<uint16_t> lut;
    ULut ul;
    BLut bl;
    if (kind == 0) {
        if (!BuildULut(grid, n, kMaxKeys, ul)) return Fail(PIMC_ERR_UNSUPPORTED, "grid admits no uniform interval table");
        lut = ul.lut;
    } else {
        if (!BuildBLut(grid, n, kMaxKeys, bl)) return Fail(PIMC_ERR_UNSUPPORTED, "grid admits no bit-pattern interval table");
        lut = bl.lut;
    }
    const int key_max = (int)lut.size() - 1;
    if (n_keys) *n_keys = (int32_t)lut.size();
    for (int i = 0; i < m; ++i) {
        int key;
        if (kind == 0) {  // ULookup: low word of fma(x, 1/h, 2^52 + 2^51)
            const double kd = std::fma(x[i], ul.inv_h, kRoundMagic);
            uint64_t bits;
            std::memcpy(&bits, &kd, sizeof(bits));
            key = std::min((int)(int32_t)(uint32_t)bits, key_max);
        } else {          // DLookup<1>: (high word >> shift) - key0
            uint64_t bits;
            std::memcpy(&bits, &x[i], sizeof(bits));
            key = std::min(std::max(((int)(int32_t)(bits >> 32) >> bl.shift) - bl.key0, 0), key_max);
        }
        const int i0 = lut[(size_t)key];
        const double g_next = (i0 + 1 < n) ? grid[i0 + 1] : HUGE_VAL;
        out[i] = x[i] >= g_next ? i0 + 1 : i0;
    }
    return PIMC_OK;
}

int pimc_debug_bucket_spline(int32_t n, const double *grid, const double *values, int32_t m, const double *x, double *out_interval,
                             double *out_bucket, int32_t *n_keys) {
    if (!grid || !values || !x || !out_interval || !out_bucket || n < 2 || m < 0) return Fail(PIMC_ERR_INVALID, "bad argument");
    try {
        KnotBasis kb;
        kb.Build(grid, n);
        std::vector<double> coefs((size_t)n + 3, 0.0);
        SolveNatural(kb, values, 1, coefs.data(), 1);
        const std::vector<double> pp = PPFrom1D(kb, coefs.data());
        LR2Host l2;
        if (!BuildLR2(grid, n, kb, coefs.data(), 16384, l2)) return Fail(PIMC_ERR_UNSUPPORTED, "grid admits no uniform interval table");
        if (n_keys) *n_keys = (int32_t)l2.knot.size();
        for (int j = 0; j < m; ++j) {
            const double xx = std::min(std::max(x[j], grid[0]), grid[n - 1]);  // SetLimits: the device evaluates inside the grid only
            // interval form: einspline's interval, tau = x - g[i]
            const int i = xx >= grid[n - 1] ? n - 1 : (int)(std::upper_bound(grid, grid + n, xx) - grid) - 1;
            const double t = xx - grid[i];
            out_interval[j] = std::fma(std::fma(std::fma(pp[4 * (size_t)i + 3], t, pp[4 * (size_t)i + 2]), t, pp[4 * (size_t)i + 1]), t, pp[4 * (size_t)i]);
            // bucket-centred form: the device's arithmetic (FastPP1Eval, pair_fast.cuh)
            const double kd = std::fma(xx, l2.inv_h16, kRoundMagic);
            uint64_t bits;
            std::memcpy(&bits, &kd, sizeof(bits));
            const int k32 = (int)(int32_t)(uint32_t)bits + 0x8000;
            const int key = k32 >> 16;
            if (key < 0 || key >= (int)l2.knot.size()) return Fail(PIMC_ERR_INVALID, "internal: bucket outside the table");
            const int rec = key + ((k32 & 0xFFFF) >= (int)l2.knot[(size_t)key] ? 1 : 0);
            const double tb = std::fma(-(double)rec, l2.h, xx);
            out_bucket[j] = std::fma(std::fma(std::fma(l2.c23[2 * (size_t)rec + 1], tb, l2.c23[2 * (size_t)rec]), tb, l2.c01[2 * (size_t)rec + 1]), tb,
                                     l2.c01[2 * (size_t)rec]);
        }
    } catch (const std::exception &e) {
        return Fail(PIMC_ERR_TABLE, e.what());
    }
    return PIMC_OK;
}

int pimc_ctx_force_general(pimc_ctx *ctx, int32_t enable) {
    if (!ctx) return Fail(PIMC_ERR_INVALID, "null context");
    ctx->force_general = enable != 0;
    for (pimc_action *a : ctx->actions)
        if (a->atype == ATYPE_DAVID)
            for (int which = 0; which < 2; ++which) a->table[which].dav_use_fast = (!ctx->force_general && a->fastd_ok[which]) ? 1 : 0;
    return PIMC_OK;
}

int pimc_ctx_set_timing(pimc_ctx *ctx, int32_t enable) {
    if (!ctx) return Fail(PIMC_ERR_INVALID, "null context");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    PIMC_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto &t : ctx->timers) {
        for (auto &p : t.pending) {
            cudaEventDestroy(p.first);
            cudaEventDestroy(p.second);
        }
        t.pending.clear();
        t.total_ms = 0.;
        t.n = 0;
    }
    ctx->timing = enable != 0;
    return PIMC_OK;
}

int pimc_ctx_kernel_time(pimc_ctx *ctx, int32_t kernel_id, double *total_ms, int64_t *n_launches) {
    if (!ctx || kernel_id < 0 || kernel_id >= 8) return Fail(PIMC_ERR_INVALID, "bad kernel id");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    PIMC_CUDA(cudaStreamSynchronize(ctx->stream));
    auto &t = ctx->timers[kernel_id];
    for (auto &p : t.pending) {
        float ms = 0.f;
        PIMC_CUDA(cudaEventElapsedTime(&ms, p.first, p.second));
        t.total_ms += ms;
        t.n += 1;
        cudaEventDestroy(p.first);
        cudaEventDestroy(p.second);
    }
    t.pending.clear();
    if (total_ms) *total_ms = t.total_ms;
    if (n_launches) *n_launches = t.n;
    return PIMC_OK;
}

// ------------------------------------------------------------------------ CUDA graphs
// Stream capture of a sequence of calls on the context's stream (a launch-latency-bound step such as
// the slice-sharded evaluation: ~13 short kernels + one NCCL all-reduce).  Everything the sequence
// allocates must exist already: run it once un-captured first.
int pimc_capture_begin(pimc_ctx *ctx) {
    if (!ctx) return Fail(PIMC_ERR_INVALID, "null context");
    if (ctx->timing) return Fail(PIMC_ERR_INVALID, "per-kernel timing is on: events cannot be recorded inside a capture");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    PIMC_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    return PIMC_OK;
}

int pimc_capture_end(pimc_ctx *ctx, pimc_graph **out) {
    if (!ctx || !out) return Fail(PIMC_ERR_INVALID, "null argument");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    cudaGraph_t g = nullptr;
    PIMC_CUDA(cudaStreamEndCapture(ctx->stream, &g));
    if (!g) return Fail(PIMC_ERR_CUDA, "stream capture was invalidated (a captured call synchronised or allocated)");
    pimc_graph *pg = new pimc_graph;
    pg->ctx = ctx;
    pg->graph = g;
    cudaError_t e = cudaGraphInstantiate(&pg->exec, g, 0);
    if (e != cudaSuccess) {
        cudaGraphDestroy(g);
        delete pg;
        return Fail(PIMC_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    }
    size_t n_nodes = 0;
    cudaGraphGetNodes(g, nullptr, &n_nodes);
    pg->n_nodes = (int64_t)n_nodes;
    *out = pg;
    return PIMC_OK;
}

int pimc_graph_launch(pimc_graph *g) {
    if (!g) return Fail(PIMC_ERR_INVALID, "null graph");
    PIMC_CUDA(cudaSetDevice(g->ctx->device));
    PIMC_CUDA(cudaGraphLaunch(g->exec, g->ctx->stream));
    g->ctx->launches += g->n_nodes;
    return PIMC_OK;
}

int64_t pimc_graph_nodes(pimc_graph *g) { return g ? g->n_nodes : 0; }

int pimc_graph_destroy(pimc_graph *g) {
    if (!g) return PIMC_OK;
    cudaSetDevice(g->ctx->device);
    cudaStreamSynchronize(g->ctx->stream);
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
    return PIMC_OK;
}

// ------------------------------------------------------------ internal.h (other translation units)
int pimc_internal_evaluate_many(pimc_ctx *ctx, int which, pimc_action *const *actions, int n, double *d_out) {
    if (!ctx || !actions || !d_out || n < 1) return Fail(PIMC_ERR_INVALID, "bad argument");
    if (which != WHICH_U && which != WHICH_DU && which != WHICH_V) return Fail(PIMC_ERR_INVALID, "which must be 0 (action), 1 (dU/dbeta) or 2 (potential)");
    for (int i = 0; i < n; ++i)
        if (!actions[i] || actions[i]->ctx != ctx) return Fail(PIMC_ERR_INVALID, "action of another context");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    if (n == 1 || ctx->timing) {  // per-kernel event pairs would time a kernel's wait for SMs as well: one after the other
        for (int i = 0; i < n; ++i) {
            const int rc = FullEvaluation(actions[i], which, d_out + (size_t)i * ctx->C);
            if (rc != PIMC_OK) return rc;
        }
        return PIMC_OK;
    }
    while ((int)ctx->side_streams.size() < n - 1) {
        cudaStream_t st = nullptr;
        cudaEvent_t ev = nullptr;
        PIMC_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        ctx->side_streams.push_back(st);
        PIMC_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->side_events.push_back(ev);
    }
    if (!ctx->fork_event) PIMC_CUDA(cudaEventCreateWithFlags(&ctx->fork_event, cudaEventDisableTiming));
    cudaStream_t main_stream = ctx->stream;
    PIMC_CUDA(cudaEventRecord(ctx->fork_event, main_stream));
    // an action evaluated beside others works in its own scratch and on its own stream: the launch helpers read both
    // from the context, so they are swapped in for the duration of the call and restored on every path out
    struct Swap {
        pimc_ctx *ctx;
        pimc_action *a;
        cudaStream_t main_stream;
        void Exchange() {
            std::swap(ctx->partial.p, a->own_partial.p);
            std::swap(ctx->partial.n, a->own_partial.n);
            std::swap(ctx->lr_dev.p, a->own_lr.p);
            std::swap(ctx->lr_dev.n, a->own_lr.n);
        }
        Swap(pimc_ctx *c, pimc_action *act, cudaStream_t s) : ctx(c), a(act), main_stream(c->stream) {
            Exchange();
            ctx->stream = s;
        }
        ~Swap() {
            ctx->stream = main_stream;
            Exchange();
        }
    };
    int rc = PIMC_OK;
    int forked = 0;
    for (int i = 0; i < n && rc == PIMC_OK; ++i) {
        cudaStream_t st = i == 0 ? main_stream : ctx->side_streams[(size_t)i - 1];
        if (i > 0) {
            PIMC_CUDA(cudaStreamWaitEvent(st, ctx->fork_event, 0));
            forked = i;
        }
        {
            Swap sw(ctx, actions[i], st);
            rc = FullEvaluation(actions[i], which, d_out + (size_t)i * ctx->C);
        }
        if (i > 0) {  // joined even after a failure: a stream forked inside a capture must come back
            const cudaError_t e = cudaEventRecord(ctx->side_events[(size_t)i - 1], st);
            if (e != cudaSuccess && rc == PIMC_OK) rc = Fail(PIMC_ERR_CUDA, std::string("cudaEventRecord: ") + cudaGetErrorString(e));
        }
    }
    for (int i = 1; i <= forked; ++i) {
        const cudaError_t e = cudaStreamWaitEvent(main_stream, ctx->side_events[(size_t)i - 1], 0);
        if (e != cudaSuccess && rc == PIMC_OK) rc = Fail(PIMC_ERR_CUDA, std::string("cudaStreamWaitEvent: ") + cudaGetErrorString(e));
    }
    return rc;
}

int pimc_internal_fail(int code, const char *msg) { return Fail(code, msg ? msg : ""); }
int pimc_internal_device(const pimc_ctx *ctx) { return ctx->device; }
int pimc_internal_n_clones(const pimc_ctx *ctx) { return ctx->C; }
int pimc_internal_n_species(const pimc_ctx *ctx) { return (int)ctx->species.size(); }
int pimc_internal_n_part(const pimc_ctx *ctx, int s) { return (s < 0 || s >= (int)ctx->species.size()) ? -1 : ctx->species[s]->N; }

int pimc_fp64_peak(pimc_ctx *ctx, double *tflops) {
    if (!ctx || !tflops) return Fail(PIMC_ERR_INVALID, "null argument");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    const int blocks = ctx->n_sm * 8, threads = 256, iters = 1 << 16;
    DevBuf<double> buf;
    PIMC_CUDA(buf.Alloc((size_t)blocks * threads));
    cudaEvent_t e0, e1;
    PIMC_CUDA(cudaEventCreate(&e0));
    PIMC_CUDA(cudaEventCreate(&e1));
    double best = 0.;
    for (int rep = 0; rep < 4; ++rep) {
        PIMC_CUDA(cudaEventRecord(e0, ctx->stream));
        fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(buf.p, iters);
        PIMC_CUDA(cudaEventRecord(e1, ctx->stream));
        PIMC_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        PIMC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return PIMC_OK;
}

}  // extern "C"
