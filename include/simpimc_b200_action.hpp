// simpimc_b200_action.hpp -- the reference-side binding of the C ABI.
//
// `class GpuPairAction : public Action` is what a simpimc maintainer adds next to
// src/actions/pair_action/ilkka_pair_action_class.h: same constructor signature as every
// reference action (Path&, Input&, IO&), same XML attributes as PairAction /
// IlkkaPairAction / BarePairAction / DavidPairAction (pair_action_class.h:204-238,
// ilkka_pair_action_class.h:253-432, bare_pair_action_class.h:22-96, david_pair_action_class.h:192-360),
// same virtuals (action_class.h:37-73).  ActionFactory (src/actions/actions.h:13-35) returns it for
// type="IlkkaPairAction" / "BarePairAction" / "DavidPairAction";
// moves and estimators are untouched.  It compiles only inside the reference tree (it needs
// the reference's Path / Species / Bead / Input / IO); everything CUDA sits behind
// include/simpimc_b200.h.
//
// State mirroring: the reference keeps positions in Bead objects (r = NEW copy, r_c = OLD
// copy, bead_class.h:98-104).  One GpuPathMirror per Path holds a one-clone device context:
//   * committed positions are uploaded at construction and whenever the mirror is marked
//     stale (any Accept of a move that the adapter did not see the proposal of);
//   * the first NEW-mode GetAction of a move uploads the moved particle's beads b0..b1 as the
//     pending proposal (pimc_propose); OLD-mode calls read the committed copy;
//   * Accept()/Reject() of the FIRST action called after a proposal commits or drops it
//     (pimc_commit); the remaining actions' calls in the same move are no-ops.
#ifndef SIMPIMC_B200_ACTION_HPP_
#define SIMPIMC_B200_ACTION_HPP_

#include <cstdlib>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "simpimc_b200.h"

/// Device twin of one reference Path (n_clones = 1), shared by every GpuPairAction on it.
class GpuPathMirror {
   public:
    pimc_ctx *ctx = nullptr;
    Path &path;
    bool stale = true;             ///< committed device positions differ from the beads' r_c
    bool proposal_pending = false;  ///< pimc_propose issued, pimc_commit not yet
    bool committed_this_move = false;  ///< the pending proposal of the current move has been committed / dropped already
    std::vector<std::pair<int32_t, int32_t>> prop_particles;  ///< (species, particle) with a pending proposal

    /// One mirror per live Path.  The registry is guarded (simulations of different paths may be set up
    /// from different threads) and an entry dies with its mirror, so a Path allocated later at the same
    /// address gets a fresh device context.
    static std::shared_ptr<GpuPathMirror> Get(Path &path) {
        std::lock_guard<std::mutex> lock(RegistryMutex());
        std::shared_ptr<GpuPathMirror> m = Registry()[&path].lock();
        if (!m) {
            m = std::shared_ptr<GpuPathMirror>(new GpuPathMirror(path));
            Registry()[&path] = m;
        }
        return m;
    }

    ~GpuPathMirror() {
        {
            std::lock_guard<std::mutex> lock(RegistryMutex());
            auto it = Registry().find(&path);
            if (it != Registry().end() && it->second.expired()) Registry().erase(it);
        }
        pimc_ctx_destroy(ctx);
    }

    static void Check(int rc, const char *what) {
        if (rc != PIMC_OK) {  // reference error style: message + exit(1) (actions.h:32-33)
            std::cerr << "ERROR: simpimc_b200 " << what << ": " << pimc_last_error() << std::endl;
            exit(1);
        }
    }

    /// Beads' committed copies -> device (Species::InitPaths + InitRhoK on the device side).
    void UploadCommitted() {
        ModeType saved = path.GetMode();
        path.SetMode(OLD_MODE);
        const uint32_t M = path.GetNBead(), n_d = path.GetND();
        for (uint32_t s = 0; s < path.GetNSpecies(); ++s) {
            std::shared_ptr<Species> sp = path.GetSpecies()[s];
            std::vector<double> R((size_t)sp->GetNPart() * M * n_d);
            for (uint32_t p = 0; p < sp->GetNPart(); ++p)
                for (uint32_t b = 0; b < M; ++b) {
                    const vec<double> &r = sp->GetBead(p, b)->GetR();
                    for (uint32_t d = 0; d < n_d; ++d) R[((size_t)p * M + b) * n_d + d] = r(d);
                }
            Check(pimc_positions_upload(ctx, (int32_t)s, 0, 1, R.data()), "pimc_positions_upload");
        }
        path.SetMode(saved);
        stale = false;
        proposal_pending = false;
        prop_particles.clear();
    }

    /// NEW copies of beads b0..b1 of one particle -> pending proposal.
    void Propose(const std::shared_ptr<Species> &sp, uint32_t p, uint32_t b0, uint32_t b1) {
        const uint32_t M = path.GetNBead(), n_d = path.GetND();
        const uint32_t n = std::min(b1 - b0 + 1, M);
        ModeType saved = path.GetMode();
        path.SetMode(NEW_MODE);
        std::vector<double> newR((size_t)n * n_d);
        for (uint32_t i = 0; i < n; ++i) {
            const vec<double> &r = sp->GetBead(p, sp->bead_loop(b0 + i))->GetR();
            for (uint32_t d = 0; d < n_d; ++d) newR[(size_t)i * n_d + d] = r(d);
        }
        path.SetMode(saved);
        const int32_t particle = (int32_t)p, first = (int32_t)sp->bead_loop(b0);
        Check(pimc_propose(ctx, (int32_t)sp->GetId(), &particle, &first, (int32_t)n, newR.data()), "pimc_propose");
        proposal_pending = true;
        committed_this_move = false;
        prop_particles.push_back(std::make_pair((int32_t)sp->GetId(), (int32_t)p));
    }

    /// Accept() / Reject() of every action of a move arrive here (bisect_class.h:34-35,137-138): the
    /// first one commits or drops the pending proposal, the others of the same move are no-ops.  An
    /// Accept for a move this mirror never saw a proposal of (e.g. one whose pair actions all returned
    /// before evaluating) marks the device copy stale.
    void Finish(bool accept) {
        if (!proposal_pending) {
            if (accept && !committed_this_move) stale = true;
            return;
        }
        committed_this_move = true;
        const int32_t flag = accept ? 1 : 0;
        Check(pimc_commit(ctx, &flag), "pimc_commit");
        // several particles of one species = a permutation move: on acceptance the reference relabels
        // the permuted beads (PermBisect::AssignParticleLabels, perm_bisect_class.h:47-56), which this
        // mirror does not follow -- re-upload the committed beads before the next evaluation
        if (accept)
            for (size_t i = 0; i < prop_particles.size(); ++i)
                for (size_t j = i + 1; j < prop_particles.size(); ++j)
                    if (prop_particles[i].first == prop_particles[j].first) stale = true;
        proposal_pending = false;
        prop_particles.clear();
    }

   private:
    static std::map<Path *, std::weak_ptr<GpuPathMirror>> &Registry() {
        static std::map<Path *, std::weak_ptr<GpuPathMirror>> registry;
        return registry;
    }
    static std::mutex &RegistryMutex() {
        static std::mutex m;
        return m;
    }

    explicit GpuPathMirror(Path &p) : path(p) {
        std::vector<int32_t> n_part;
        std::vector<double> lambda;
        for (auto &sp : path.GetSpecies()) {
            n_part.push_back((int32_t)sp->GetNPart());
            lambda.push_back(sp->GetLambda());
        }
        pimc_config cfg;
        cfg.n_d = (int32_t)path.GetND();
        cfg.pbc = path.GetPBC() ? 1 : 0;
        cfg.L = path.GetL();
        cfg.beta = path.GetTau() * path.GetNBead();
        cfg.n_bead = (int32_t)path.GetNBead();
        cfg.n_species = (int32_t)n_part.size();
        cfg.n_part = n_part.data();
        cfg.lambda = lambda.data();
        cfg.n_clones = 1;
        const char *dev = getenv("SIMPIMC_B200_DEVICE");
        cfg.device = dev ? atoi(dev) : 0;
        cfg.slice_lo = 0;
        cfg.slice_hi = cfg.n_bead;
        Check(pimc_ctx_create(&cfg, &ctx), "pimc_ctx_create");
        if (path.GetPBC() && path.ks.cutoff > 0.) {
            int32_t n_k = 0;
            Check(pimc_kspace_setup(ctx, path.ks.cutoff, &n_k), "pimc_kspace_setup");
        }
    }
};

/// IlkkaPairAction / BarePairAction / DavidPairAction evaluated on the GPU behind the reference's Action API.
class GpuPairAction : public Action {
   private:
    std::shared_ptr<GpuPathMirror> mirror;
    pimc_action *act = nullptr;
    std::shared_ptr<Species> species_a, species_b;
    bool use_long_range = false;
    bool is_constant = false;
    double k_cut = 0.;

    static pimc_table_1d T1(const vec<double> &r, const vec<double> &f) {
        pimc_table_1d t;
        t.n = (int32_t)r.size();
        t.r = r.memptr();
        t.f = f.memptr();
        return t;
    }

    struct LongRangeData {
        vec<double> r, f_r, k, f_k;
        double f_r_0 = 0., f_k_0 = 0.;
        void Read(IO &pa_in, const std::string &obj) {  // <obj>/diag/* (ilkka...:286-305)
            uint32_t n_r, n_k;
            pa_in.Read(obj + "/diag/n_r_long", n_r);
            r.set_size(n_r);
            f_r.set_size(n_r);
            pa_in.Read(obj + "/diag/r_long", r);
            pa_in.Read(obj + "/diag/" + obj + "_long_r", f_r);
            pa_in.Read(obj + "/diag/" + obj + "_long_r_0", f_r_0);
            pa_in.Read(obj + "/diag/n_k", n_k);
            k.set_size(n_k);
            f_k.set_size(n_k);
            pa_in.Read(obj + "/diag/k", k);
            pa_in.Read(obj + "/diag/" + obj + "_long_k", f_k);
            pa_in.Read(obj + "/diag/" + obj + "_long_k_0", f_k_0);
        }
        pimc_long_range Pack() const {
            pimc_long_range lr;
            lr.f_r = T1(r, f_r);
            lr.f_r_0 = f_r_0;
            lr.n_k = (int32_t)k.size();
            lr.k = k.memptr();
            lr.f_k = f_k.memptr();
            lr.f_k_0 = f_k_0;
            return lr;
        }
    };

   public:
    GpuPairAction(Path &path, Input &in, IO &out) : Action(path, in, out) {
        // the attributes PairAction reads (pair_action_class.h:208-222)
        std::string species_a_name = in.GetAttribute<std::string>("species_a");
        std::string species_b_name = in.GetAttribute<std::string>("species_b");
        species_a = path.GetSpecies(species_a_name);
        species_b = path.GetSpecies(species_b_name);
        species_list.push_back(species_a);
        species_list.push_back(species_b);
        use_long_range = in.GetAttribute<bool>("use_long_range", 0);
        if (use_long_range) {
            k_cut = in.GetAttribute<double>("k_cut", path.ks.cutoff);
            path.ks.Setup(k_cut);  // keeps the reference's own KSpace consistent for its observables
        }
        is_constant = ((species_a == species_b) && (species_a->GetNPart() == 1 || species_a->GetLambda() == 0.));
        mirror = GpuPathMirror::Get(path);
        if (use_long_range) {  // an earlier action may have created the mirror with a smaller cutoff
            int32_t n_k = 0;
            GpuPathMirror::Check(pimc_kspace_setup(mirror->ctx, k_cut, &n_k), "pimc_kspace_setup");
        }
        std::string file_name = in.GetAttribute<std::string>("file");
        IO pa_in;
        pa_in.Load(file_name);
        const int32_t sa = (int32_t)species_a->GetId(), sb = (int32_t)species_b->GetId();
        if (type == "IlkkaPairAction") {
            uint32_t n_x, n_y, n_r;
            vec<double> x[2], y[2], r_v, v_r;
            mat<double> f[2];
            const char *obj[2] = {"u", "du"};
            for (int w = 0; w < 2; ++w) {
                const std::string o(obj[w]);
                pa_in.Read(o + "/off_diag/n_x", n_x);
                pa_in.Read(o + "/off_diag/n_y", n_y);
                x[w].set_size(n_x);
                y[w].set_size(n_y);
                f[w].set_size(n_x, n_y);
                pa_in.Read(o + "/off_diag/x", x[w]);
                pa_in.Read(o + "/off_diag/y", y[w]);
                pa_in.Read(o + "/off_diag/" + o + "_xy", f[w]);
            }
            pa_in.Read("v/diag/n_r", n_r);
            r_v.set_size(n_r);
            v_r.set_size(n_r);
            pa_in.Read("v/diag/r", r_v);
            pa_in.Read("v/diag/v_r", v_r);
            LongRangeData lr[3];
            if (use_long_range) {
                lr[0].Read(pa_in, "u");
                lr[1].Read(pa_in, "du");
                lr[2].Read(pa_in, "v");
            }
            pimc_ilkka_tables t;
            pimc_table_2d *xy[2] = {&t.u_xy, &t.du_xy};
            for (int w = 0; w < 2; ++w) {
                xy[w]->n_x = (int32_t)x[w].size();
                xy[w]->n_y = (int32_t)y[w].size();
                xy[w]->x = x[w].memptr();
                xy[w]->y = y[w].memptr();
                xy[w]->f = f[w].memptr();  // file bytes are row-major f[ix * n_y + iy] (ilkka...:272-280)
            }
            t.v_r = T1(r_v, v_r);
            t.u_long = lr[0].Pack();
            t.du_long = lr[1].Pack();
            t.v_long = lr[2].Pack();
            GpuPathMirror::Check(pimc_action_create_ilkka(mirror->ctx, sa, sb, &t, (int32_t)max_level, use_long_range ? 1 : 0, k_cut, &act),
                                 "pimc_action_create_ilkka");
        } else if (type == "BarePairAction") {
            uint32_t n_r;
            vec<double> r_v, v_r;
            pa_in.Read("v/diag/n_r", n_r);
            r_v.set_size(n_r);
            v_r.set_size(n_r);
            pa_in.Read("v/diag/r", r_v);
            pa_in.Read("v/diag/v_r", v_r);
            LongRangeData lr;
            if (use_long_range) lr.Read(pa_in, "v");
            pimc_bare_tables t;
            t.v_r = T1(r_v, v_r);
            t.v_long = lr.Pack();
            t.is_coulomb = in.GetAttribute<bool>("is_coulomb", 0) ? 1 : 0;
            GpuPathMirror::Check(pimc_action_create_bare(mirror->ctx, sa, sb, &t, (int32_t)max_level, use_long_range ? 1 : 0, k_cut, &act),
                                 "pimc_action_create_bare");
        } else if (type == "DavidPairAction") {
            // the datasets DavidPairAction's constructor reads (david_pair_action_class.h:194-336), in its order
            const uint32_t n_order = in.GetAttribute<uint32_t>("n_order", 0);  // pair_action_class.h:208: a misspelt attribute
                                                                                // (inputs/h-atom/h-david.xml "nOrder") silently reads order 0
            std::stringstream su, sdu;
            su << "/u_kj_" << n_order;
            sdu << "/du_kj_dbeta_" << n_order;
            const std::string u_str = su.str(), du_str = sdu.str();
            double r_start, r_end;
            uint32_t n_grid;
            std::string grid_type;
            vec<double> grid_points;
            pa_in.Read(u_str + "/grid/start", r_start);
            pa_in.Read(u_str + "/grid/end", r_end);
            pa_in.Read(u_str + "/grid/n_grid_points", n_grid);
            pa_in.Read(u_str + "/grid/type", grid_type);
            pimc_david_tables t;
            t.grid_points = nullptr;
            if ((grid_type.find("LOG") != std::string::npos) && (grid_type.find("LOGLIN") == std::string::npos)) {
                t.grid_type = PIMC_GRID_LOG;
            } else if (grid_type.find("LINEAR") != std::string::npos) {
                t.grid_type = PIMC_GRID_LINEAR;
            } else if (grid_type.find("LOGLIN") != std::string::npos) {
                std::cerr << "ERROR: GpuPairAction: LOGLIN grids (create_loglin_grid exists only in the reference's einspline fork) are not supported" << std::endl;
                exit(1);
            } else {
                grid_points.set_size(n_grid);
                pa_in.Read(u_str + "/grid/grid_points", grid_points);
                t.grid_type = PIMC_GRID_GENERAL;
                t.grid_points = grid_points.memptr();
            }
            const uint32_t n_tau = max_level + 1;
            vec<double> taus(n_tau);
            pa_in.Read(u_str + "/taus", taus);
            vec<double> V(n_grid);
            pa_in.Read("/potential/data", V);
            uint32_t n_val = 1;
            for (uint32_t i = 1; i <= n_order; ++i) n_val += 1 + i;
            cube<double> u_kj(n_val, n_grid, n_tau), du_kj(n_val, n_grid, n_tau);  // column-major: [tau][grid][val] in memory
            pa_in.Read(u_str + "/data", u_kj);
            pa_in.Read(du_str + "/data", du_kj);
            vec<double> k_v, u_k;
            double v_image = 0.;
            t.n_k = 0;
            t.k_points = t.u_k = nullptr;
            if (use_long_range) {
                uint32_t n_k_v;
                pa_in.Read("long_range/n_k", n_k_v);
                k_v.set_size(n_k_v);
                u_k.set_size(n_k_v);
                pa_in.Read("long_range/k_points", k_v);
                pa_in.Read("long_range/u_k", u_k);
                pa_in.Read("squarer/v_image", v_image);
                t.n_k = (int32_t)n_k_v;
                t.k_points = k_v.memptr();
                t.u_k = u_k.memptr();
            }
            t.r_start = r_start;
            t.r_end = r_end;
            t.n_grid = (int32_t)n_grid;
            t.n_order = (int32_t)n_order;
            t.n_tau = (int32_t)n_tau;
            t.taus = taus.memptr();
            t.u_kj = u_kj.memptr();
            t.du_kj_dbeta = du_kj.memptr();
            t.potential = V.memptr();
            t.v_image = v_image;
            // a tau missing from the table is PIMC_ERR_TABLE -> "ERROR: ..." + exit(1), as david...:241-245
            GpuPathMirror::Check(pimc_action_create_david(mirror->ctx, sa, sb, &t, (int32_t)max_level, use_long_range ? 1 : 0, &act),
                                 "pimc_action_create_david");
        } else {
            std::cerr << "ERROR: GpuPairAction does not implement " << type << std::endl;
            exit(1);
        }
        mirror->stale = true;
    }

    virtual double DActionDBeta() {
        if (mirror->stale) mirror->UploadCommitted();
        double v = 0.;
        GpuPathMirror::Check(pimc_action_dbeta(act, &v), "pimc_action_dbeta");
        return v;
    }

    virtual double Potential() {
        if (mirror->stale) mirror->UploadCommitted();
        double v = 0.;
        GpuPathMirror::Check(pimc_action_potential(act, &v), "pimc_action_potential");
        return v;
    }

    virtual double GetAction(const uint32_t b0, const uint32_t b1, const std::vector<std::pair<std::shared_ptr<Species>, uint32_t>> &particles,
                             const uint32_t level) {
        if (level > max_level || is_constant) return 0.;  // pair_action_class.h:269
        if (mirror->stale) mirror->UploadCommitted();
        if (!mirror->proposal_pending) mirror->committed_this_move = false;  // the first evaluation of the next move
        std::vector<int32_t> sp, pi;
        for (auto &p : particles) {
            sp.push_back((int32_t)p.first->GetId());
            pi.push_back((int32_t)p.second);
        }
        const int32_t mode = path.GetMode() ? PIMC_NEW : PIMC_OLD;
        if (mode == PIMC_NEW) {
            for (auto &p : particles) {
                bool have = false;
                for (auto &q : mirror->prop_particles) have = have || (q.first == (int32_t)p.first->GetId() && q.second == (int32_t)p.second);
                if (!have) mirror->Propose(p.first, p.second, b0, b1);
            }
        }
        const int32_t first = (int32_t)b0;
        double v = 0.;
        GpuPathMirror::Check(pimc_action_get(act, mode, &first, (int32_t)(b1 - b0), (int32_t)sp.size(), sp.data(), pi.data(), (int32_t)level, &v),
                             "pimc_action_get");
        return v;
    }

    /// Action::GetActionGradient / GetActionLaplacian (action_class.h:46,49): called by the
    /// contact-density and virial estimators between moves (no proposal pending).
    virtual vec<double> GetActionGradient(const uint32_t b0, const uint32_t b1,
                                          const std::vector<std::pair<std::shared_ptr<Species>, uint32_t>> &particles, const uint32_t level) {
        vec<double> g(zeros<vec<double>>(path.GetND()));
        if (level > max_level || is_constant) return g;  // pair_action_class.h:308
        if (mirror->stale) mirror->UploadCommitted();
        std::vector<int32_t> sp, pi;
        for (auto &p : particles) {
            sp.push_back((int32_t)p.first->GetId());
            pi.push_back((int32_t)p.second);
        }
        const int32_t first = (int32_t)(b0 % path.GetNBead());
        double v[3] = {0., 0., 0.};
        GpuPathMirror::Check(pimc_action_gradient(act, &first, (int32_t)(b1 - b0), (int32_t)sp.size(), sp.data(), pi.data(), (int32_t)level, v),
                             "pimc_action_gradient");
        for (uint32_t d = 0; d < path.GetND(); ++d) g(d) = v[d];
        return g;
    }

    virtual double GetActionLaplacian(const uint32_t b0, const uint32_t b1,
                                      const std::vector<std::pair<std::shared_ptr<Species>, uint32_t>> &particles, const uint32_t level) {
        if (level > max_level || is_constant) return 0.;  // pair_action_class.h:342
        if (mirror->stale) mirror->UploadCommitted();
        std::vector<int32_t> sp, pi;
        for (auto &p : particles) {
            sp.push_back((int32_t)p.first->GetId());
            pi.push_back((int32_t)p.second);
        }
        const int32_t first = (int32_t)(b0 % path.GetNBead());
        double v = 0.;
        GpuPathMirror::Check(pimc_action_laplacian(act, &first, (int32_t)(b1 - b0), (int32_t)sp.size(), sp.data(), pi.data(), (int32_t)level, &v),
                             "pimc_action_laplacian");
        return v;
    }

    /// PairAction::ImportanceWeight (pair_action_class.h:398-400).
    virtual double ImportanceWeight() { return is_importance_weight ? exp(DActionDBeta() / path.GetNBead()) : 1.; }

    virtual void Accept() {
        mirror->Finish(true);
        pimc_action_accept(act);
    }

    virtual void Reject() {
        mirror->Finish(false);
        pimc_action_reject(act);
    }

    virtual void Write() {}
};

#endif  // SIMPIMC_B200_ACTION_HPP_
