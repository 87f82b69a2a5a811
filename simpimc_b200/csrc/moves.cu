// Device-resident moves behind the C ABI (include/simpimc_b200.h): pimc_bisect_sweep / _windows (Bisect::DoEvent),
// pimc_perm_bisect_sweep (PermBisectIterative::DoEvent) with the permutation at the beta seam, pimc_displace_sweep
// (DisplaceParticle::DoEvent) and pimc_perm_table.  The context and everything else live in capi.cu; state.h is shared.
#include "state.h"
#include "mc.cuh"
#include "sweep_fused.cuh"
#include "displace.cuh"
#include "perm.cuh"

using namespace pimc;

extern "C" {

// ------------------------------------------------------------------- device-resident moves
/// windows = 0: one window per walker and attempt (Bisect::DoEvent); != 0: every walker's path is tiled with disjoint
/// windows that are all attempted in the same launches (pimc_bisect_sweep_windows), *n_windows_out = windows per walker.
static int BisectSweepImpl(pimc_ctx *ctx, int32_t s, int32_t n_level, int32_t n_attempts, uint64_t seed, uint64_t attempt0,
                           int32_t with_kinetic, int64_t *n_accept, int windows, int32_t *n_windows_out) {
    if (!ctx) return Fail(PIMC_ERR_INVALID, "null context");
    if (s < 0 || s >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    if (n_level < 1 || (1 << n_level) > kMaxBisectBeads || (1 << n_level) > ctx->M)
        return Fail(PIMC_ERR_INVALID, "n_level must satisfy 2 <= 2^n_level <= min(32, n_bead)");
    if (n_attempts < 0) return Fail(PIMC_ERR_INVALID, "negative attempt count");
    // a slice shard moves the windows that lie inside its stored slices [slice_lo, slice_hi]: its first
    // slice and its halo stay fixed until the caller rotates the ring (pimc_rotate_*) and refreshes the halos
    if (ctx->sharded && (1 << n_level) > ctx->Mloc) return Fail(PIMC_ERR_INVALID, "window longer than the slice shard");
    int b0_lo = ctx->sharded ? ctx->slice_lo : 0, b0_count = ctx->sharded ? ctx->Mloc - (1 << n_level) + 1 : ctx->M;
    // disjoint windows: W = floor(slices / 2^n_level) windows per walker at b0_lo + offset + w 2^n_level; the offset is
    // one of the shifts that keep every window inside the path (unsharded: any of 2^n_level when the windows tile the
    // ring exactly, else the slack; a shard: the slack of its stored slices -- its first slice and halo stay fixed)
    int W = 1;
    if (windows) {
        const int span = ctx->sharded ? ctx->Mloc : ctx->M, nbw = 1 << n_level;
        W = span / nbw;
        if (W < 1) return Fail(PIMC_ERR_INVALID, "window longer than the path / shard");
        b0_count = (!ctx->sharded && W * nbw == span) ? nbw : span - W * nbw + 1;
        if (n_windows_out) *n_windows_out = W;
    }
    SpeciesState &st = *ctx->species[s];
    if (!(st.lambda > 0.)) return Fail(PIMC_ERR_INVALID, "bisection of a species with lambda = 0");
    for (auto &sp : ctx->species)
        if (sp->n_prop > 0) return Fail(PIMC_ERR_INVALID, "a proposal is pending: call pimc_commit first");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    {
        const int rc_p = RequireUnpermuted(ctx, s, "pimc_bisect_sweep");
        if (rc_p != PIMC_OK) return rc_p;
    }
    const int Cw = ctx->C;                       // walkers
    const int C = Cw * W;                        // virtual clones: (walker, window); == walkers in the classic mode
    const int nb = 1 << n_level, n_prop = nb - 1, n_k = ctx->n_k();
    if (windows) {
        if (st.P_particle.n < (size_t)kMaxPropSlots * C) PIMC_CUDA(st.P_particle.Alloc((size_t)kMaxPropSlots * C));
        if (st.P_first.n < (size_t)kMaxPropSlots * C) PIMC_CUDA(st.P_first.Alloc((size_t)kMaxPropSlots * C));
    }
    if (ctx->mc_f64.n < (size_t)6 * C) PIMC_CUDA(ctx->mc_f64.Alloc((size_t)6 * C));
    if (ctx->mc_i32.n < (size_t)3 * C) PIMC_CUDA(ctx->mc_i32.Alloc((size_t)3 * C));
    if (ctx->mc_naccept.n < (size_t)C) PIMC_CUDA(ctx->mc_naccept.Alloc(C));
    PIMC_CUDA(cudaMemsetAsync(ctx->mc_naccept.p, 0, C * sizeof(long long), ctx->stream));
    if (st.P.n < (size_t)C * n_prop * 3) PIMC_CUDA(st.P.Alloc((size_t)C * n_prop * 3));
    double *partial = ctx->mc_f64.p, *logu0 = partial + C, *pair_old = logu0 + C, *pair_new = pair_old + C, *lr_old = pair_new + C,
           *lr_new = lr_old + C;
    int32_t *alive = ctx->mc_i32.p, *b0 = alive + C, *accept = b0 + C;
    // move_class.h:27-31: the actions that involve this species
    std::vector<pimc_action *> acts;
    bool any_lr = false;
    for (pimc_action *a : ctx->actions) {
        if (a->sa != s && a->sb != s) continue;
        if (a->is_constant || a->atype == ATYPE_KINETIC) continue;  // the kinetic action is evaluated with the Levy construction
        acts.push_back(a);
        any_lr = any_lr || (a->use_long_range && n_k > 0);
    }
    if (any_lr) {
        const size_t need = (size_t)C * nb * n_k;
        if (st.drho.n < need) PIMC_CUDA(st.drho.Alloc(need));
    }
    // free-particle splines with periodic images: the move's own (Bisect n_images) for the sampling
    // probabilities, the Kinetic action's for the action (closed forms when n_images = 0)
    FreeSet *fs_move = nullptr, *fs_kin = nullptr;
    int rc_fs;
    if ((rc_fs = GetFreeSet(ctx, s, st.move_images, &fs_move)) != PIMC_OK) return rc_fs;
    if ((rc_fs = GetFreeSet(ctx, s, st.kinetic ? st.kinetic->n_images : 0, &fs_kin)) != PIMC_OK) return rc_fs;
    if ((fs_move->view.n_images || fs_kin->view.n_images) && n_level + 1 > kMaxFreeSplines)
        return Fail(PIMC_ERR_UNSUPPORTED, "n_level too large for the tabulated free-particle splines");
    PathView pv = ctx->View();
    if (windows) {
        pv.C = C;
        pv.vdiv = W;
    }
    // one same-species fast Ilkka action on the moved species: the whole sweep is one launch
    // (sweep_fused.cuh); everything else takes the kernel-per-phase path below
    bool fused = !windows && acts.size() == 1 && acts[0]->sa == s && acts[0]->sb == s && acts[0]->atype == ATYPE_ILKKA && acts[0]->fast_ok[WHICH_U] &&
                 !ctx->force_general && nb <= kSweepMaxBeads && n_level <= kSweepMaxLevel;
    size_t fused_smem = 0;
    if (fused) {
        const int tl = 2 * ctx->max_index + 1;
        fused_smem = (size_t)acts[0]->fast[WHICH_U].n_bytes + (any_lr ? (size_t)kSweepClones * nb * 6 * tl * sizeof(double2) : 0);
        if (fused_smem + sizeof(SweepShared) + 1024 > (size_t)ctx->smem_optin) fused = false;
    }
    if (fused && n_attempts > 0) {
        pimc_action *a = acts[0];
        SweepFusedArgs f;
        if (PIMC_SWEEP_MIRROR && !st.R2_valid) {
            const size_t n2 = (size_t)C * ctx->Mstore * st.N * 3;
            if (st.R2.n != n2) PIMC_CUDA(st.R2.Alloc(n2));
            dim3 tgrid((ctx->Mstore + 31) / 32, (st.N * 3 + 31) / 32, C);
            slice_major_kernel<<<tgrid, dim3(32, 8), 0, ctx->stream>>>(st.R.p, st.N * 3, ctx->Mstore, ctx->Ms, st.R2.p);
            ctx->launches++;
            st.R2_valid = true;
        }
        f.pv = pv;
        f.R = st.R.p;
        f.R2 = st.R2.p;
        f.N = st.N;
        f.lambda = st.lambda;
        f.tau = ctx->tau;
        f.n_level = n_level;
        f.with_kinetic = with_kinetic ? 1 : 0;
        f.fs_move = fs_move->view;
        f.fs_kin = fs_kin->view;
        f.b0_lo = b0_lo;
        f.b0_count = b0_count;
        f.seed_lo = (uint32_t)seed;
        f.seed_hi = (uint32_t)(seed >> 32);
        f.attempt0 = attempt0;
        f.n_attempts = n_attempts;
        f.FT = a->fast[WHICH_U];
        f.fast_tables = a->fast_tab[WHICH_U].p;
        f.use_lr = any_lr ? 1 : 0;
        f.ks = ctx->KView();
        f.rho = any_lr ? st.rho.p : nullptr;
        f.rho_new = any_lr ? st.drho.p : nullptr;  // the proposal's rho_k buffer doubles as the scratch
        f.wk = any_lr ? a->wk[WHICH_U].p : nullptr;
        f.lr_factor = a->ulong_scale;
        f.n_accept = ctx->mc_naccept.p;
        const bool images = f.fs_move.n_images > 0 || f.fs_kin.n_images > 0;
        auto *kernel = images ? bisect_sweep_fused_kernel<true> : bisect_sweep_fused_kernel<false>;
        PIMC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem));
        {
            ScopedKernelTimer t(ctx, PIMC_KERNEL_PAIR_WINDOW);
            kernel<<<std::min(C, ctx->n_sm), kSweepThreads, fused_smem, ctx->stream>>>(f);
        }
        ctx->launches++;
    }
    if (!fused && n_attempts > 0) st.R2_valid = false;  // the kernel-per-phase path commits into R only
    for (int it = 0; it < (fused ? 0 : n_attempts); ++it) {
        const uint64_t attempt = attempt0 + (uint64_t)it;
        BisectArgs ba;
        ba.pv = pv;
        ba.R = st.R.p;
        ba.N = st.N;
        ba.lambda = st.lambda;
        ba.tau = ctx->tau;
        ba.n_level = n_level;
        ba.with_kinetic = with_kinetic ? 1 : 0;
        ba.fs_move = fs_move->view;
        ba.fs_kin = fs_kin->view;
        ba.b0_lo = b0_lo;
        ba.b0_count = b0_count;
        ba.seed_lo = (uint32_t)seed;
        ba.seed_hi = (uint32_t)(seed >> 32);
        ba.attempt_lo = (uint32_t)attempt;
        ba.attempt_hi = (uint32_t)(attempt >> 32);
        ba.P = st.P.p;
        ba.P_particle = st.P_particle.p;
        ba.P_first = st.P_first.p;
        ba.b0 = b0;
        ba.partial = partial;
        ba.logu0 = logu0;
        ba.alive = alive;
        ba.pair_old = pair_old;
        ba.pair_new = pair_new;
        ba.lr_old = lr_old;
        ba.lr_new = lr_new;
        bisect_sample_kernel<<<(C + kSampleWarps - 1) / kSampleWarps, kSampleWarps * 32, 0, ctx->stream>>>(ba);
        ctx->launches++;
        for (pimc_action *a : acts) {
            const int partner = (a->sa == s) ? a->sb : a->sa;
            WindowBothArgs w;
            w.pv = pv;
            w.R_moved = st.R.p;
            w.N_moved = st.N;
            w.R_partner = ctx->species[partner]->R.p;
            w.N_partner = ctx->species[partner]->N;
            w.same = partner == s;
            w.P = st.P.p;
            w.P_particle = st.P_particle.p;
            w.b0 = b0;
            w.alive = alive;
            w.n_links = nb;
            w.FT = a->fast[WHICH_U];
            w.fast_tables = a->fast_tab[WHICH_U].p;
            w.T = a->table[WHICH_U];
            w.blob = a->blob[WHICH_U].p;
            w.out_old = pair_old;
            w.out_new = pair_new;
            {
                ScopedKernelTimer t(ctx, PIMC_KERNEL_PAIR_WINDOW);
                if (a->atype == ATYPE_ILKKA && a->fast_ok[WHICH_U] && !ctx->force_general) {
                    const size_t smem = (size_t)w.FT.n_bytes;
                    PIMC_CUDA(cudaFuncSetAttribute(pair_window_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    pair_window_fast_kernel<<<std::min(C, ctx->n_sm), kWinFastThreads, smem, ctx->stream>>>(w);
                } else if (a->atype == ATYPE_ILKKA)
                    pair_window_both_kernel<ATYPE_ILKKA, false><<<C, kWindowThreads, 0, ctx->stream>>>(w);
                else if (a->atype == ATYPE_BARE)
                    pair_window_both_kernel<ATYPE_BARE, false><<<C, kWindowThreads, 0, ctx->stream>>>(w);
                else
                    pair_window_both_kernel<ATYPE_DAVID, false><<<C, kWindowThreads, 0, ctx->stream>>>(w);
            }
            ctx->launches++;
        }
        if (any_lr) {
            LrWindowArgs l;
            l.pv = pv;
            l.sv = ctx->SView(s, false);
            l.sv.n_prop = n_prop;  // the proposal the sample kernel has just written
            l.sv.n_slots = 1;
            l.ks = ctx->KView();
            l.b0 = b0;
            l.n_window = nb;
            l.rho_self = st.rho.p;
            l.drho = st.drho.p;
            l.lr_old = lr_old;
            l.lr_new = lr_new;
            l.fuse_decide = 1;  // decision and commit ride on this launch
            l.chunk_partial = nullptr;
            l.n_chunks = 0;
            l.alive = alive;
            l.partial = partial;
            l.logu0 = logu0;
            l.pair_old = pair_old;
            l.pair_new = pair_new;
            l.N = st.N;
            l.R = st.R.p;
            l.rho_commit = st.rho.p;
            l.accept = accept;
            l.n_accept = ctx->mc_naccept.p;
            l.n_actions = 0;
            for (pimc_action *a : acts) {
                if (!(a->use_long_range && n_k > 0)) continue;
                if (l.n_actions == kMaxLrActions) return Fail(PIMC_ERR_UNSUPPORTED, "more than 8 long-range actions on one species");
                const int partner = (a->sa == s) ? a->sb : a->sa;
                l.rho_other[l.n_actions] = (partner == s) ? nullptr : ctx->species[partner]->rho.p;
                l.wk[l.n_actions] = a->wk[WHICH_U].p;
                l.factor[l.n_actions] = a->ulong_scale * (partner == s ? 1.0 : 2.0);
                l.n_actions++;
            }
            const int tl = 2 * ctx->max_index + 1;
            {
                ScopedKernelTimer t(ctx, PIMC_KERNEL_KSUM);
                const size_t lr_smem = (size_t)kLrChunk * 6 * tl * sizeof(double2);
                if (lr_smem > 48 * 1024) PIMC_CUDA(cudaFuncSetAttribute(lr_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lr_smem));
                lr_window_kernel<<<C, 256, lr_smem, ctx->stream>>>(l);
            }
            ctx->launches++;
        }
        if (!any_lr) {  // with a long-range action the decision and the commit rode on lr_window_kernel
            bisect_decide_commit_kernel<<<C, 256, 0, ctx->stream>>>(pv, st.N, n_k, nb, alive, partial, logu0, pair_old, pair_new, lr_old, lr_new,
                                                                   st.P.p, st.P_particle.p, b0, nullptr, st.R.p, nullptr, accept, ctx->mc_naccept.p);
            ctx->launches++;
        }
    }
    PIMC_CUDA(cudaGetLastError());
    st.n_prop = 0;
    st.n_slots = 0;
    st.drho_valid = false;
    for (auto &sp : ctx->species) sp->need_update_rho_k = true;
    if (n_accept) {
        std::vector<long long> h(C);
        PIMC_CUDA(cudaMemcpyAsync(h.data(), ctx->mc_naccept.p, C * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        PIMC_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int c = 0; c < C; ++c) n_accept[c / W] += (int64_t)h[c];
    }
    return PIMC_OK;
}

int pimc_bisect_sweep(pimc_ctx *ctx, int32_t s, int32_t n_level, int32_t n_attempts, uint64_t seed, uint64_t attempt0,
                      int32_t with_kinetic, int64_t *n_accept) {
    return BisectSweepImpl(ctx, s, n_level, n_attempts, seed, attempt0, with_kinetic, n_accept, 0, nullptr);
}

int pimc_bisect_sweep_windows(pimc_ctx *ctx, int32_t s, int32_t n_level, int32_t n_rounds, uint64_t seed, uint64_t attempt0,
                              int32_t with_kinetic, int64_t *n_accept, int32_t *n_windows) {
    return BisectSweepImpl(ctx, s, n_level, n_rounds, seed, attempt0, with_kinetic, n_accept, 1, n_windows);
}


int pimc_perm_table(pimc_ctx *ctx, int32_t s, const int32_t *b0, int32_t n_bisect_beads, double epsilon, int32_t relative, double *t) {
    if (!ctx || !b0 || !t) return Fail(PIMC_ERR_INVALID, "null argument");
    if (s < 0 || s >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    if (ctx->sharded) return Fail(PIMC_ERR_UNSUPPORTED, "permutation table on a slice-sharded context");
    if (n_bisect_beads < 1 || n_bisect_beads > ctx->M) return Fail(PIMC_ERR_INVALID, "n_bisect_beads must be in 1..n_bead");
    if (!(epsilon > 0.)) return Fail(PIMC_ERR_INVALID, "epsilon must be positive");
    SpeciesState &st = *ctx->species[s];
    if (!(st.lambda > 0.)) return Fail(PIMC_ERR_INVALID, "permutation table of a species with lambda = 0");
    for (int c = 0; c < ctx->C; ++c)
        if (b0[c] < 0 || b0[c] >= ctx->M) return Fail(PIMC_ERR_INVALID, "window start out of range");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = EnsureI32(ctx, ctx->i32_c, b0, ctx->C)) != PIMC_OK) return rc;
    const size_t n = (size_t)ctx->C * st.N * st.N;
    if (ctx->stage.n < n) PIMC_CUDA(ctx->stage.Alloc(n));
    // Bisect: i_4_lambda_tau = 1 / (4 lambda tau), divided by n_bisect_beads (bisect_class.h:147-150)
    const double i_4_lambda_tau_n = (1. / (4. * st.lambda * ctx->tau)) / n_bisect_beads;
    perm_table_kernel<<<ctx->C * st.N, 128, 0, ctx->stream>>>(ctx->View(), st.R.p, st.N, ctx->i32_c.p, n_bisect_beads, i_4_lambda_tau_n, std::log(epsilon),
                                                              relative ? 1 : 0, ctx->stage.p);
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    return ToHost(ctx, ctx->stage.p, t, n);
}

// ------------------------------------------------------------------- permuting bisection
/// The species' seam permutation, created as the identity on first use.
static int EnsurePermutation(pimc_ctx *ctx, int s) {
    SpeciesState &st = *ctx->species[s];
    if (st.perm_tracked) return PIMC_OK;
    const size_t n = (size_t)ctx->C * st.N;
    PIMC_CUDA(st.perm_next.Alloc(n));
    perm_identity_kernel<<<(int)((n + 255) / 256), 256, 0, ctx->stream>>>(st.perm_next.p, ctx->C, st.N);
    ctx->launches++;
    PIMC_CUDA(cudaGetLastError());
    st.perm_tracked = true;
    return PIMC_OK;
}

}  // extern "C"

/// Entry points that read a particle's path by LABEL across the beta seam (Bisect, DisplaceParticle, Kinetic::GetAction
/// windows) are exact only on an unpermuted path: refuse loudly once a permuting move has changed the seam.
int pimc_host::RequireUnpermuted(pimc_ctx *ctx, int s, const char *what) {
    SpeciesState &st = *ctx->species[s];
    if (!st.perm_tracked) return PIMC_OK;
    std::vector<int32_t> h((size_t)ctx->C * st.N);
    PIMC_CUDA(cudaMemcpyAsync(h.data(), st.perm_next.p, h.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    PIMC_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < h.size(); ++i)
        if (h[i] != (int32_t)(i % st.N))
            return Fail(PIMC_ERR_UNSUPPORTED, std::string(what) + " on a species whose path is permuted at the beta seam: use pimc_perm_bisect_sweep "
                                                                  "(cycles of one particle are the plain bisection, links followed)");
    return PIMC_OK;
}

extern "C" {

int pimc_permutation_get(pimc_ctx *ctx, int32_t s, int32_t *next) {
    if (!ctx || !next) return Fail(PIMC_ERR_INVALID, "null argument");
    if (s < 0 || s >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    SpeciesState &st = *ctx->species[s];
    if (!st.perm_tracked) {
        for (int c = 0; c < ctx->C; ++c)
            for (int p = 0; p < st.N; ++p) next[(size_t)c * st.N + p] = p;
        return PIMC_OK;
    }
    PIMC_CUDA(cudaSetDevice(ctx->device));
    PIMC_CUDA(cudaMemcpyAsync(next, st.perm_next.p, (size_t)ctx->C * st.N * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    PIMC_CUDA(cudaStreamSynchronize(ctx->stream));
    return PIMC_OK;
}

int pimc_permutation_set(pimc_ctx *ctx, int32_t s, const int32_t *next) {
    if (!ctx || !next) return Fail(PIMC_ERR_INVALID, "null argument");
    if (s < 0 || s >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    if (ctx->sharded) return Fail(PIMC_ERR_UNSUPPORTED, "permutations on a slice-sharded context");
    SpeciesState &st = *ctx->species[s];
    std::vector<char> seen(st.N);
    for (int c = 0; c < ctx->C; ++c) {  // a permutation of the labels per walker
        std::fill(seen.begin(), seen.end(), 0);
        for (int p = 0; p < st.N; ++p) {
            const int32_t q = next[(size_t)c * st.N + p];
            if (q < 0 || q >= st.N || seen[q]) return Fail(PIMC_ERR_INVALID, "next[] is not a permutation of the particle labels");
            seen[q] = 1;
        }
    }
    PIMC_CUDA(cudaSetDevice(ctx->device));
    int rc = EnsurePermutation(ctx, s);
    if (rc != PIMC_OK) return rc;
    PIMC_CUDA(cudaMemcpyAsync(st.perm_next.p, next, (size_t)ctx->C * st.N * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    PIMC_CUDA(cudaStreamSynchronize(ctx->stream));
    return PIMC_OK;
}

int pimc_perm_last_cycle(pimc_ctx *ctx, int32_t *b0, int32_t *n_perm, int32_t *particles, int32_t *n_steps, int32_t *accept) {
    if (!ctx) return Fail(PIMC_ERR_INVALID, "null context");
    const int C = ctx->C;
    if (ctx->perm_i32.n < (size_t)C * (2 + kPermMaxLen) || ctx->mc_i32.n < (size_t)3 * C) return Fail(PIMC_ERR_INVALID, "no permuting bisection has run on this context");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    const int32_t *d_n_perm = ctx->perm_i32.p, *d_steps = d_n_perm + C, *d_part = d_steps + C;
    const int32_t *d_b0 = ctx->mc_i32.p + C, *d_accept = ctx->mc_i32.p + 2 * C;
    if (b0) PIMC_CUDA(cudaMemcpyAsync(b0, d_b0, C * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (n_perm) PIMC_CUDA(cudaMemcpyAsync(n_perm, d_n_perm, C * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (particles) PIMC_CUDA(cudaMemcpyAsync(particles, d_part, (size_t)C * kPermMaxLen * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (n_steps) PIMC_CUDA(cudaMemcpyAsync(n_steps, d_steps, C * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (accept) PIMC_CUDA(cudaMemcpyAsync(accept, d_accept, C * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    PIMC_CUDA(cudaStreamSynchronize(ctx->stream));
    return PIMC_OK;
}

int pimc_perm_bisect_sweep(pimc_ctx *ctx, int32_t s, int32_t n_level, int32_t n_attempts, uint64_t seed, uint64_t attempt0,
                           int32_t with_kinetic, double epsilon, int64_t *n_accept, int64_t *perm_attempt, int64_t *perm_accept) {
    if (!ctx) return Fail(PIMC_ERR_INVALID, "null context");
    if (s < 0 || s >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    if (n_level < 1 || (1 << n_level) > kMaxBisectBeads || (1 << n_level) > ctx->M)
        return Fail(PIMC_ERR_INVALID, "n_level must satisfy 2 <= 2^n_level <= min(32, n_bead)");
    if (n_attempts < 0) return Fail(PIMC_ERR_INVALID, "negative attempt count");
    if (!(epsilon > 0.)) return Fail(PIMC_ERR_INVALID, "epsilon must be positive");
    // a cycle relabels every slice after the window up to the seam: one decision for all shards of a path
    if (ctx->sharded) return Fail(PIMC_ERR_UNSUPPORTED, "permuting bisection on a slice-sharded context");
    SpeciesState &st = *ctx->species[s];
    if (!(st.lambda > 0.)) return Fail(PIMC_ERR_INVALID, "bisection of a species with lambda = 0");
    for (auto &sp : ctx->species)
        if (sp->n_prop > 0) return Fail(PIMC_ERR_INVALID, "a proposal is pending: call pimc_commit first");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = EnsurePermutation(ctx, s)) != PIMC_OK) return rc;
    const int C = ctx->C, nb = 1 << n_level, n_prop = nb - 1, n_k = ctx->n_k(), N = st.N;
    // move_class.h:27-31: the actions that involve this species (the kinetic action rides on the Levy construction)
    std::vector<pimc_action *> acts;
    bool any_lr = false;
    for (pimc_action *a : ctx->actions) {
        if (a->sa != s && a->sb != s) continue;
        if (a->is_constant || a->atype == ATYPE_KINETIC) continue;
        acts.push_back(a);
        any_lr = any_lr || (a->use_long_range && n_k > 0);
    }
    const int n_acts = (int)acts.size();
    if (st.P.n < (size_t)kMaxPropSlots * C * n_prop * 3) PIMC_CUDA(st.P.Alloc((size_t)kMaxPropSlots * C * n_prop * 3));
    if (st.P_particle.n < (size_t)kMaxPropSlots * C) PIMC_CUDA(st.P_particle.Alloc((size_t)kMaxPropSlots * C));
    if (st.P_first.n < (size_t)kMaxPropSlots * C) PIMC_CUDA(st.P_first.Alloc((size_t)kMaxPropSlots * C));
    if (ctx->mc_f64.n < (size_t)6 * C) PIMC_CUDA(ctx->mc_f64.Alloc((size_t)6 * C));
    if (ctx->mc_i32.n < (size_t)3 * C) PIMC_CUDA(ctx->mc_i32.Alloc((size_t)3 * C));
    if (ctx->mc_naccept.n < (size_t)C) PIMC_CUDA(ctx->mc_naccept.Alloc(C));
    if (ctx->perm_i32.n < (size_t)C * (2 + kPermMaxLen)) PIMC_CUDA(ctx->perm_i32.Alloc((size_t)C * (2 + kPermMaxLen)));
    const size_t n_f64 = (size_t)3 * C + (size_t)std::max(n_acts, 1) * 2 * C * nb;
    if (ctx->perm_f64.n < n_f64) PIMC_CUDA(ctx->perm_f64.Alloc(n_f64));
    if (ctx->perm_stage.n < (size_t)C * kPermMaxLen * 3 * ctx->M) PIMC_CUDA(ctx->perm_stage.Alloc((size_t)C * kPermMaxLen * 3 * ctx->M));
    if (ctx->perm_counts.n < (size_t)2 * C * kPermMaxLen) PIMC_CUDA(ctx->perm_counts.Alloc((size_t)2 * C * kPermMaxLen));
    PIMC_CUDA(cudaMemsetAsync(ctx->mc_naccept.p, 0, C * sizeof(long long), ctx->stream));
    PIMC_CUDA(cudaMemsetAsync(ctx->perm_counts.p, 0, (size_t)2 * C * kPermMaxLen * sizeof(long long), ctx->stream));
    if (any_lr) {
        const size_t need = (size_t)C * nb * n_k;
        if (st.drho.n < need) PIMC_CUDA(st.drho.Alloc(need));
    }
    double *partial = ctx->mc_f64.p, *logu0 = partial + C;
    int32_t *alive = ctx->mc_i32.p, *b0 = alive + C, *accept = b0 + C;
    int32_t *d_n_perm = ctx->perm_i32.p, *d_steps = d_n_perm + C, *d_part = d_steps + C;
    double *weight = ctx->perm_f64.p, *lr_old = weight + C, *lr_new = lr_old + C, *pair_parts = lr_new + C;
    long long *cnt_attempt = ctx->perm_counts.p, *cnt_accept = cnt_attempt + (size_t)C * kPermMaxLen;
    FreeSet *fs_move = nullptr, *fs_kin = nullptr;
    if ((rc = GetFreeSet(ctx, s, st.move_images, &fs_move)) != PIMC_OK) return rc;
    if ((rc = GetFreeSet(ctx, s, st.kinetic ? st.kinetic->n_images : 0, &fs_kin)) != PIMC_OK) return rc;
    if ((fs_move->view.n_images || fs_kin->view.n_images) && n_level + 1 > kMaxFreeSplines)
        return Fail(PIMC_ERR_UNSUPPORTED, "n_level too large for the tabulated free-particle splines");
    const PathView pv = ctx->View();
    const size_t select_smem = (size_t)2 * N * sizeof(double);
    if (select_smem > (size_t)ctx->smem_optin) return Fail(PIMC_ERR_UNSUPPORTED, "too many particles for the permutation-table row in shared memory");
    if (select_smem > 48 * 1024) PIMC_CUDA(cudaFuncSetAttribute(perm_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)select_smem));
    st.R2_valid = false;
    const int tl = 2 * ctx->max_index + 1;
    for (int it = 0; it < n_attempts; ++it) {
        const uint64_t attempt = attempt0 + (uint64_t)it;
        PermSelectArgs sel;
        sel.pv = pv;
        sel.R = st.R.p;
        sel.N = N;
        sel.next = st.perm_next.p;
        sel.n_bisect_beads = nb;
        sel.i_4_lambda_tau_n = (1. / (4. * st.lambda * ctx->tau)) / nb;  // bisect_class.h:147-150, perm_bisect_iterative_class.h:16
        sel.log_epsilon = std::log(epsilon);
        sel.seed_lo = (uint32_t)seed;
        sel.seed_hi = (uint32_t)(seed >> 32);
        sel.attempt_lo = (uint32_t)attempt;
        sel.attempt_hi = (uint32_t)(attempt >> 32);
        sel.b0 = b0;
        sel.n_perm = d_n_perm;
        sel.particles = d_part;
        sel.weight = weight;
        sel.n_steps = d_steps;
        perm_select_kernel<<<C, 32, select_smem, ctx->stream>>>(sel);
        ctx->launches++;
        PermSampleArgs sa;
        sa.pv = pv;
        sa.R = st.R.p;
        sa.N = N;
        sa.next = st.perm_next.p;
        sa.lambda = st.lambda;
        sa.tau = ctx->tau;
        sa.n_level = n_level;
        sa.with_kinetic = with_kinetic ? 1 : 0;
        sa.fs_move = fs_move->view;
        sa.fs_kin = fs_kin->view;
        sa.seed_lo = sel.seed_lo;
        sa.seed_hi = sel.seed_hi;
        sa.attempt_lo = sel.attempt_lo;
        sa.attempt_hi = sel.attempt_hi;
        sa.b0 = b0;
        sa.n_perm = d_n_perm;
        sa.particles = d_part;
        sa.weight = weight;
        sa.P = st.P.p;
        sa.P_particle = st.P_particle.p;
        sa.P_first = st.P_first.p;
        sa.partial = partial;
        sa.logu0 = logu0;
        sa.alive = alive;
        sa.perm_attempt = cnt_attempt;
        perm_sample_kernel<<<C, 32, 0, ctx->stream>>>(sa);
        ctx->launches++;
        // the moved species' view with the label windows as its pending proposal
        SpeciesView sv_new = ctx->SView(s, false);
        sv_new.n_prop = n_prop;
        sv_new.n_slots = kMaxPropSlots;
        for (int t = 0; t < n_acts; ++t) {
            pimc_action *a = acts[t];
            if (a->atype == ATYPE_ILKKA && a->fast_ok[WHICH_U] && !ctx->force_general) {  // OLD and NEW in one launch, pp-form tables
                PermPairArgs w;
                w.pv = pv;
                w.A_old = ctx->SView(a->sa, false);
                w.B_old = ctx->SView(a->sb, false);
                w.A_new = a->sa == s ? sv_new : w.A_old;
                w.B_new = a->sb == s ? sv_new : w.B_old;
                w.same = a->sa == a->sb;
                w.moved_is_a = a->sa == s;
                w.part = st.P_particle.p;
                w.b0 = b0;
                w.alive = alive;
                w.n_links = nb;
                w.FT = a->fast[WHICH_U];
                w.fast_tables = a->fast_tab[WHICH_U].p;
                w.partial_old = pair_parts + ((size_t)(2 * t) * C) * nb;
                w.partial_new = pair_parts + ((size_t)(2 * t + 1) * C) * nb;
                const int grid = (int)std::min<size_t>((size_t)C * nb, (size_t)ctx->n_sm * 16);
                {
                    ScopedKernelTimer tm(ctx, PIMC_KERNEL_PAIR_WINDOW);
                    perm_pair_fast_kernel<<<grid, 128, 0, ctx->stream>>>(w);
                }
                ctx->launches++;
                continue;
            }
            for (int mode = 0; mode < 2; ++mode) {
                PairWindowArgs w;
                w.pv = pv;
                w.A = (a->sa == s && mode) ? sv_new : ctx->SView(a->sa, false);
                w.B = (a->sb == s && mode) ? sv_new : ctx->SView(a->sb, false);
                w.same = a->sa == a->sb;
                w.n_a = a->sa == s ? kMaxPropSlots : 0;
                w.n_b = (a->sa != s && a->sb == s) ? kMaxPropSlots : 0;
                w.part_a = st.P_particle.p;
                w.part_b = st.P_particle.p;
                w.b0 = b0;
                w.n_links = nb;
                w.mode = mode;
                w.T = a->table[WHICH_U];
                w.blob = a->blob[WHICH_U].p;
                w.partial = pair_parts + ((size_t)(2 * t + mode) * C) * nb;
                w.alive = alive;
                const int grid = (int)std::min<size_t>((size_t)C * nb, (size_t)ctx->n_sm * 16);
                {
                    ScopedKernelTimer tm(ctx, PIMC_KERNEL_PAIR_WINDOW);
                    switch (a->atype) {
                        case ATYPE_ILKKA: pair_window_kernel<ATYPE_ILKKA><<<grid, 128, 0, ctx->stream>>>(w); break;
                        case ATYPE_BARE: pair_window_kernel<ATYPE_BARE><<<grid, 128, 0, ctx->stream>>>(w); break;
                        default: pair_window_kernel<ATYPE_DAVID><<<grid, 128, 0, ctx->stream>>>(w); break;
                    }
                }
                ctx->launches++;
            }
        }
        if (any_lr) {
            // Species::UpdateRhoK for every listed label (species_class.h:406-425), then CalcULong over the window
            // in OLD and NEW mode for every long-range action of the species (ilkka_pair_action_class.h:104-122)
            rhok_delta_kernel<<<GridFor(ctx, C * nb), 256, (size_t)6 * tl * sizeof(double2), ctx->stream>>>(pv, sv_new, ctx->KView(), b0, nb, st.drho.p);
            ctx->launches++;
            PIMC_CUDA(cudaMemsetAsync(lr_old, 0, (size_t)2 * C * sizeof(double), ctx->stream));
            for (pimc_action *a : acts) {
                if (!a->use_long_range) continue;
                for (int mode = 0; mode < 2; ++mode) {
                    KSumArgs k;
                    k.pv = pv;
                    k.n_k = n_k;
                    k.rho_a = ctx->species[a->sa]->rho.p;
                    k.rho_b = ctx->species[a->sb]->rho.p;
                    k.drho_a = (mode && a->sa == s) ? st.drho.p : nullptr;
                    k.drho_b = (mode && a->sb == s) ? st.drho.p : nullptr;
                    k.wk = a->wk[WHICH_U].p;
                    k.b0 = b0;
                    k.n_window = nb;
                    k.twice = a->sa != a->sb;
                    k.scale = a->ulong_scale;
                    k.accumulate = 1;
                    k.out = mode ? lr_new : lr_old;
                    ScopedKernelTimer tm(ctx, PIMC_KERNEL_KSUM);
                    ksum_kernel<<<C, 256, 0, ctx->stream>>>(k);
                    ctx->launches++;
                }
            }
        }
        PermDecideArgs d;
        d.pv = pv;
        d.N = N;
        d.n_k = n_k;
        d.nb = nb;
        d.alive = alive;
        d.partial = partial;
        d.logu0 = logu0;
        d.pair_parts = pair_parts;
        d.n_pair_actions = n_acts;
        d.lr_old = any_lr ? lr_old : nullptr;
        d.lr_new = any_lr ? lr_new : nullptr;
        d.P = st.P.p;
        d.P_particle = st.P_particle.p;
        d.b0 = b0;
        d.drho = any_lr ? st.drho.p : nullptr;
        d.R = st.R.p;
        d.rho = st.rho.p;
        d.n_perm = d_n_perm;
        d.accept = accept;
        d.n_accept = ctx->mc_naccept.p;
        d.perm_accept = cnt_accept;
        perm_decide_commit_kernel<<<C, 256, 0, ctx->stream>>>(d);
        ctx->launches++;
        PermApplyArgs ap;
        ap.pv = pv;
        ap.R = st.R.p;
        ap.N = N;
        ap.next = st.perm_next.p;
        ap.b0 = b0;
        ap.n_bisect_beads = nb;
        ap.accept = accept;
        ap.n_perm = d_n_perm;
        ap.particles = d_part;
        ap.scratch = ctx->perm_stage.p;
        perm_apply_kernel<<<C, 256, 0, ctx->stream>>>(ap);
        ctx->launches++;
    }
    PIMC_CUDA(cudaGetLastError());
    st.n_prop = 0;
    st.n_slots = 0;
    st.drho_valid = false;
    for (auto &sp : ctx->species) sp->need_update_rho_k = true;
    if (n_accept || perm_attempt || perm_accept) {
        std::vector<long long> h(C), hc((size_t)2 * C * kPermMaxLen);
        PIMC_CUDA(cudaMemcpyAsync(h.data(), ctx->mc_naccept.p, C * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        PIMC_CUDA(cudaMemcpyAsync(hc.data(), ctx->perm_counts.p, hc.size() * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        PIMC_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int c = 0; c < C; ++c) {
            if (n_accept) n_accept[c] += (int64_t)h[c];
            for (int i = 0; i < kPermMaxLen; ++i) {
                if (perm_attempt) perm_attempt[(size_t)c * kPermMaxLen + i] += (int64_t)hc[(size_t)c * kPermMaxLen + i];
                if (perm_accept) perm_accept[(size_t)c * kPermMaxLen + i] += (int64_t)hc[((size_t)C + c) * kPermMaxLen + i];
            }
        }
    }
    return PIMC_OK;
}

int pimc_displace_sweep(pimc_ctx *ctx, int32_t s, double step_size, int32_t n_attempts, uint64_t seed, uint64_t attempt0, int64_t *n_accept) {
    if (!ctx) return Fail(PIMC_ERR_INVALID, "null context");
    if (s < 0 || s >= (int)ctx->species.size()) return Fail(PIMC_ERR_INVALID, "species index out of range");
    if (!(step_size > 0.)) return Fail(PIMC_ERR_INVALID, "step_size must be positive");
    if (n_attempts < 0) return Fail(PIMC_ERR_INVALID, "negative attempt count");
    // a whole-path shift spans every shard: it would need one decision for all ranks
    if (ctx->sharded) return Fail(PIMC_ERR_UNSUPPORTED, "DisplaceParticle on a slice-sharded context");
    for (auto &sp : ctx->species)
        if (sp->n_prop > 0) return Fail(PIMC_ERR_INVALID, "a proposal is pending: call pimc_commit first");
    PIMC_CUDA(cudaSetDevice(ctx->device));
    {
        const int rc_p = RequireUnpermuted(ctx, s, "pimc_displace_sweep");
        if (rc_p != PIMC_OK) return rc_p;
    }
    SpeciesState &st = *ctx->species[s];
    const int C = ctx->C, M = ctx->M, n_k = ctx->n_k();
    if (ctx->mc_f64.n < (size_t)6 * C) PIMC_CUDA(ctx->mc_f64.Alloc((size_t)6 * C));
    if (ctx->mc_i32.n < (size_t)3 * C) PIMC_CUDA(ctx->mc_i32.Alloc((size_t)3 * C));
    if (ctx->mc_naccept.n < (size_t)C) PIMC_CUDA(ctx->mc_naccept.Alloc(C));
    PIMC_CUDA(cudaMemsetAsync(ctx->mc_naccept.p, 0, C * sizeof(long long), ctx->stream));
    if (st.P.n < (size_t)C * M * 3) PIMC_CUDA(st.P.Alloc((size_t)C * M * 3));
    double *dr = ctx->mc_f64.p, *logu = dr + 3 * C, *lr_old = logu + C, *lr_new = lr_old + C;
    int32_t *b0 = ctx->mc_i32.p, *accept = b0 + C;
    std::vector<pimc_action *> acts;  // move_class.h:27-31
    bool any_lr = false;
    for (pimc_action *a : ctx->actions) {
        if (a->sa != s && a->sb != s) continue;
        if (a->is_constant || a->atype == ATYPE_KINETIC) continue;  // the kinetic action is evaluated with the Levy construction
        acts.push_back(a);
        any_lr = any_lr || (a->use_long_range && n_k > 0);
    }
    if (any_lr) {
        const size_t need = (size_t)C * M * n_k;
        if (st.drho.n < need) PIMC_CUDA(st.drho.Alloc(need));
    }
    const PathView pv = ctx->View();
    const int n_chunks = (M + 31) / 32;
    if (ctx->partial.n < (size_t)C * n_chunks * 2) PIMC_CUDA(ctx->partial.Alloc((size_t)C * n_chunks * 2));
    const int grid = (int)std::min<size_t>((size_t)C * n_chunks, (size_t)ctx->n_sm);
    if (n_attempts > 0) st.R2_valid = false;
    for (int it = 0; it < n_attempts; ++it) {
        const uint64_t attempt = attempt0 + (uint64_t)it;
        DisplaceSampleArgs sa;
        sa.pv = pv;
        sa.R = st.R.p;
        sa.N = st.N;
        sa.step = step_size;
        sa.seed_lo = (uint32_t)seed;
        sa.seed_hi = (uint32_t)(seed >> 32);
        sa.attempt_lo = (uint32_t)attempt;
        sa.attempt_hi = (uint32_t)(attempt >> 32);
        sa.P = st.P.p;
        sa.P_particle = st.P_particle.p;
        sa.P_first = st.P_first.p;
        sa.b0 = b0;
        sa.dr = dr;
        sa.logu = logu;
        displace_sample_kernel<<<C, 128, 0, ctx->stream>>>(sa);
        ctx->launches++;
        if (acts.empty()) PIMC_CUDA(cudaMemsetAsync(ctx->partial.p, 0, (size_t)C * n_chunks * 2 * sizeof(double), ctx->stream));
        bool first = true;
        for (pimc_action *a : acts) {
            const int partner = (a->sa == s) ? a->sb : a->sa;
            DisplacePairArgs w;
            w.pv = pv;
            w.R_moved = st.R.p;
            w.N_moved = st.N;
            w.R_partner = ctx->species[partner]->R.p;
            w.N_partner = ctx->species[partner]->N;
            w.same = partner == s;
            w.particle = st.P_particle.p;
            w.dr = dr;
            w.n_chunks = n_chunks;
            w.FT = a->fast[WHICH_U];
            w.fast_tables = a->fast_tab[WHICH_U].p;
            w.T = a->table[WHICH_U];
            w.blob = a->blob[WHICH_U].p;
            w.accumulate = first ? 0 : 1;
            w.partial = ctx->partial.p;
            first = false;
            {
                ScopedKernelTimer t(ctx, PIMC_KERNEL_PAIR_WINDOW);
                if (a->atype == ATYPE_ILKKA && a->fast_ok[WHICH_U] && !ctx->force_general) {
                    const size_t smem = (((size_t)w.FT.n_bytes + 15) / 16 * 16) + kDispRingBytes;   // the ring takes the room K1's partner tile has
                    PIMC_CUDA(cudaFuncSetAttribute(displace_pair_kernel<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    displace_pair_kernel<-1><<<grid, kDispThreads, smem, ctx->stream>>>(w);
                } else if (a->atype == ATYPE_ILKKA)
                    displace_pair_kernel<ATYPE_ILKKA><<<grid, kDispThreads, 0, ctx->stream>>>(w);
                else if (a->atype == ATYPE_BARE)
                    displace_pair_kernel<ATYPE_BARE><<<grid, kDispThreads, 0, ctx->stream>>>(w);
                else
                    displace_pair_kernel<ATYPE_DAVID><<<grid, kDispThreads, 0, ctx->stream>>>(w);
            }
            ctx->launches++;
        }
        if (any_lr) {
            LrWindowArgs l;
            l.pv = pv;
            l.sv = ctx->SView(s, false);
            l.sv.n_prop = M;  // the proposal the sample kernel has just written: every bead of the particle
            l.ks = ctx->KView();
            l.b0 = b0;
            l.n_window = M;
            l.rho_self = st.rho.p;
            l.drho = st.drho.p;
            l.lr_old = lr_old;
            l.lr_new = lr_new;
            l.fuse_decide = 1;  // decision and commit ride on this launch (the pair sums arrive per 32-link chunk)
            l.recompute_commit = 1;  // the window is the whole path: no drho round trip through HBM
            l.alive = nullptr;
            l.partial = l.pair_old = l.pair_new = nullptr;
            l.logu0 = logu;
            l.chunk_partial = ctx->partial.p;
            l.n_chunks = n_chunks;
            l.N = st.N;
            l.R = st.R.p;
            l.rho_commit = st.rho.p;
            l.accept = accept;
            l.n_accept = ctx->mc_naccept.p;
            l.n_actions = 0;
            for (pimc_action *a : acts) {
                if (!(a->use_long_range && n_k > 0)) continue;
                if (l.n_actions == kMaxLrActions) return Fail(PIMC_ERR_UNSUPPORTED, "more than 8 long-range actions on one species");
                const int partner = (a->sa == s) ? a->sb : a->sa;
                l.rho_other[l.n_actions] = (partner == s) ? nullptr : ctx->species[partner]->rho.p;
                l.wk[l.n_actions] = a->wk[WHICH_U].p;
                l.factor[l.n_actions] = a->ulong_scale * (partner == s ? 1.0 : 2.0);
                l.n_actions++;
            }
            const int tl = 2 * ctx->max_index + 1;
            {
                ScopedKernelTimer t(ctx, PIMC_KERNEL_KSUM);
                const size_t lr_smem = (size_t)kLrChunk * 6 * tl * sizeof(double2);
                if (lr_smem > 48 * 1024) PIMC_CUDA(cudaFuncSetAttribute(lr_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lr_smem));
                lr_window_kernel<<<C, 256, lr_smem, ctx->stream>>>(l);
            }
            ctx->launches++;
        }
        if (!any_lr) {  // with a long-range action the decision and the commit rode on lr_window_kernel
            displace_decide_commit_kernel<<<C, 256, 0, ctx->stream>>>(pv, st.N, n_k, n_chunks, ctx->partial.p, 0, lr_old, lr_new, logu, st.P.p,
                                                                     st.P_particle.p, nullptr, st.R.p, nullptr, accept, ctx->mc_naccept.p);
            ctx->launches++;
        }
    }
    PIMC_CUDA(cudaGetLastError());
    st.n_prop = 0;
    st.n_slots = 0;
    st.drho_valid = false;
    for (auto &sp : ctx->species) sp->need_update_rho_k = true;
    if (n_accept) {
        std::vector<long long> h(C);
        PIMC_CUDA(cudaMemcpyAsync(h.data(), ctx->mc_naccept.p, C * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        PIMC_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int c = 0; c < C; ++c) n_accept[c] += (int64_t)h[c];
    }
    return PIMC_OK;
}

}  // extern "C"
