// TEST INFRASTRUCTURE ONLY -- builds oracle/_ref/libsimpimc_ref.so.
//
// This translation unit compiles the UNMODIFIED reference headers from /root/reference/src
// (included where they lie; nothing is copied into this repository) against the functional
// shims in oracle/shim/ for the three libraries the image lacks (Armadillo, einspline,
// HDF5), and exposes a small C ABI so Python (ctypes) can drive the reference's own
// Path / Species / PairAction / Bisect / observables objects.  It is the checker for the
// CPU restatement in oracle/pimc_oracle.cc and the generator of tests/golden/*.
// It is never linked into, imported by, or called from the product library.
//
// Build: see oracle/Makefile (g++ -std=c++11, no -ffast-math, USE_MPI undefined).

#include <assert.h>
#include <sys/time.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include <deque>

#include <scaffold/matrix/matrix.h>
#include <scaffold/algorithm/algorithm.h>
#include <scaffold/io/io_xml.h>
#include <scaffold/io/io_hdf5.h>

// Random-number injection (tests only).  The reference's RNG (scaffold/rng/rng.h) draws from
// std::uniform_real_distribution / std::normal_distribution over a private std::mt19937; a parallel
// generator cannot reproduce that stream, so the stream-exact tests go the other way: they hand the
// reference's moves the numbers the device's Philox stream drew.  While the two queues below are
// empty the stand-ins forward to the real distributions, so every other use of this library sees
// the unmodified std::mt19937 stream.  (Only the two distribution NAMES are redirected while rng.h
// is read; the reference's file is included where it lies, unmodified.)
namespace pimc_inject {
inline std::deque<double> &Uniforms() {
    static std::deque<double> q;
    return q;
}
inline std::deque<double> &Normals() {
    static std::deque<double> q;
    return q;
}
}  // namespace pimc_inject
namespace std {
template <class T>
class pimc_injected_uniform {
    std::uniform_real_distribution<T> real;

   public:
    pimc_injected_uniform() : real() {}
    pimc_injected_uniform(T a, T b) : real(a, b) {}
    template <class G>
    T operator()(G &g) {
        std::deque<double> &q = pimc_inject::Uniforms();
        if (!q.empty()) {
            const T v = (T)q.front();
            q.pop_front();
            return v;
        }
        return real(g);
    }
};
template <class T>
class pimc_injected_normal {
    std::normal_distribution<T> real;

   public:
    pimc_injected_normal() : real() {}
    template <class G>
    T operator()(G &g) {
        std::deque<double> &q = pimc_inject::Normals();
        if (!q.empty()) {
            const T v = (T)q.front();
            q.pop_front();
            return v;
        }
        return real(g);
    }
};
}  // namespace std
#define uniform_real_distribution pimc_injected_uniform
#define normal_distribution pimc_injected_normal
#include <scaffold/rng/rng.h>
#undef uniform_real_distribution
#undef normal_distribution
#include <einspline/nubspline.h>

using namespace scaffold::matrix;
using namespace scaffold::algorithm;
using namespace scaffold::io;
using namespace scaffold::rand;

// The driver needs to reach protected members (CalcU, affected_beads, histogram storage)
// of the reference classes without editing them.
#define private public
#define protected public
#include <src/actions/pair_action/bare_pair_action_class.h>
#include <src/actions/pair_action/david_pair_action_class.h>
#include <src/actions/pair_action/ilkka_pair_action_class.h>
#include <src/actions/single_action/kinetic_class.h>
#include <src/events/moves/single_species_move/bisect/bisect_class.h>
#include <src/events/moves/single_species_move/displace_particle_class.h>
#include <src/events/moves/single_species_move/bisect/perm_bisect/perm_bisect_iterative_class.h>
#include <src/events/observables/energy_class.h>
#include <src/events/observables/pair_correlation_class.h>
#include <src/events/observables/structure_factor_class.h>
#undef private
#undef protected

// integration/Makefile compiles this same driver with -DPIMC_DROPIN_GPU: the reference's moves
// and observables then run on top of the CUDA library through the adapter a maintainer would
// add to the reference tree (include/simpimc_b200_action.hpp).
#ifdef PIMC_DROPIN_GPU
#include "simpimc_b200_action.hpp"
#endif

namespace {

struct RefSim {
    Input in;
    IO out;
    std::unique_ptr<RNG> rng;
    std::unique_ptr<Path> path;
    std::vector<std::shared_ptr<Action>> actions;
    std::vector<std::shared_ptr<Move>> moves;
    std::vector<std::shared_ptr<Observable>> observables;
};

std::shared_ptr<Action> MakeAction(Input &in, IO &out, Path &path) {
    std::string type = in.GetAttribute<std::string>("type");
    if (type == "Kinetic") return std::make_shared<Kinetic>(path, in, out);
#ifdef PIMC_DROPIN_GPU
    if (type == "IlkkaPairAction" || type == "BarePairAction" || type == "DavidPairAction") return std::make_shared<GpuPairAction>(path, in, out);
#endif
    if (type == "BarePairAction") return std::make_shared<BarePairAction>(path, in, out);
    if (type == "DavidPairAction") return std::make_shared<DavidPairAction>(path, in, out);
    if (type == "IlkkaPairAction") return std::make_shared<IlkkaPairAction>(path, in, out);
    std::cerr << "ref_driver: action type " << type << " is outside the hot path" << std::endl;
    std::abort();
}

std::shared_ptr<Move> MakeMove(Input &in, IO &out, Path &path, RNG &rng, std::vector<std::shared_ptr<Action>> &actions) {
    std::string type = in.GetAttribute<std::string>("type");
    if (type == "Bisect") return std::make_shared<Bisect>(path, rng, actions, in, out);
    if (type == "DisplaceParticle") return std::make_shared<DisplaceParticle>(path, rng, actions, in, out);
    if (type == "PermBisectIterative") return std::make_shared<PermBisectIterative>(path, rng, actions, in, out);
    std::cerr << "ref_driver: move type " << type << " is outside the hot path" << std::endl;
    std::abort();
}

std::shared_ptr<Observable> MakeObservable(Input &in, IO &out, Path &path, std::vector<std::shared_ptr<Action>> &actions) {
    std::string type = in.GetAttribute<std::string>("type");
    if (type == "Energy") return std::make_shared<Energy>(path, actions, in, out);
    if (type == "PairCorrelation") return std::make_shared<PairCorrelation>(path, in, out);
    if (type == "StructureFactor") return std::make_shared<StructureFactor>(path, in, out);
    std::cerr << "ref_driver: observable type " << type << " is outside the hot path" << std::endl;
    std::abort();
}

PairAction *AsPair(RefSim *s, int a) {
    PairAction *p = dynamic_cast<PairAction *>(s->actions[a].get());
    if (!p) {
        std::cerr << "ref_driver: action " << a << " is not a pair action" << std::endl;
        std::abort();
    }
    return p;
}

}  // namespace

extern "C" {

/// Same construction order as Simulation's constructor (simulation_class.h:16-39).
void *ref_create(const char *xml_file, int seed, int quiet) {
    std::streambuf *old_cout = nullptr, *old_cerr = nullptr;
    std::ostringstream sink;
    if (quiet) {
        old_cout = std::cout.rdbuf(sink.rdbuf());
        old_cerr = std::cerr.rdbuf(sink.rdbuf());
    }
    RefSim *s = new RefSim;
    s->in.Load(xml_file);
    std::string out_name = std::string(xml_file) + ".capture";
    s->out.Load(out_name);
    s->out.Create();
    s->rng.reset(new RNG(seed));
    s->path.reset(new Path(0, s->in, s->out, *s->rng));
    for (auto &input : s->in.GetChild("Actions").GetChildList("Action")) s->actions.push_back(MakeAction(input, s->out, *s->path));
    for (auto &input : s->in.GetChild("Moves").GetChildList("Move")) s->moves.push_back(MakeMove(input, s->out, *s->path, *s->rng, s->actions));
    for (auto &input : s->in.GetChild("Observables").GetChildList("Observable"))
        s->observables.push_back(MakeObservable(input, s->out, *s->path, s->actions));
    if (quiet) {
        std::cout.rdbuf(old_cout);
        std::cerr.rdbuf(old_cerr);
    }
    return s;
}

void ref_destroy(void *h) { delete (RefSim *)h; }

int ref_n_species(void *h) { return ((RefSim *)h)->path->GetNSpecies(); }
int ref_n_part(void *h, int sp) { return ((RefSim *)h)->path->GetSpecies()[sp]->GetNPart(); }
int ref_n_bead(void *h) { return ((RefSim *)h)->path->GetNBead(); }
double ref_tau(void *h) { return ((RefSim *)h)->path->GetTau(); }

// ---- k space -------------------------------------------------------------------------
int ref_n_k(void *h) { return ((RefSim *)h)->path->ks.vecs.size(); }
void ref_kspace(void *h, int *indices, double *vecs, double *mags, int *max_index) {
    KSpace &ks = ((RefSim *)h)->path->ks;
    const uint32_t n_d = ks.n_d;
    for (size_t k = 0; k < ks.vecs.size(); ++k) {
        for (uint32_t d = 0; d < n_d; ++d) {
            indices[k * n_d + d] = ks.indices[k](d);
            vecs[k * n_d + d] = ks.vecs[k](d);
        }
        mags[k] = ks.mags[k];
    }
    for (uint32_t d = 0; d < n_d; ++d) max_index[d] = ks.max_index(d);
}

// ---- positions -----------------------------------------------------------------------
/// R is [particle][bead][dim]; sets r and r_c of every bead and rebuilds rho_k
/// (species_class.h:380-403).
void ref_set_positions(void *h, int sp, const double *R) {
    RefSim *s = (RefSim *)h;
    auto species = s->path->GetSpecies()[sp];
    const uint32_t n_d = s->path->GetND(), M = species->GetNBead(), N = species->GetNPart();
    s->path->SetMode(NEW_MODE);
    for (uint32_t p = 0; p < N; ++p)
        for (uint32_t b = 0; b < M; ++b) {
            vec<double> r(n_d);
            for (uint32_t d = 0; d < n_d; ++d) r(d) = R[(p * M + b) * n_d + d];
            species->GetBead(p, b)->SetR(r);
            species->GetBead(p, b)->StoreR();
        }
    species->InitRhoK();
    species->SetNeedUpdateRhoK(true);
}

void ref_get_positions(void *h, int sp, int mode, double *R) {
    RefSim *s = (RefSim *)h;
    auto species = s->path->GetSpecies()[sp];
    const uint32_t n_d = s->path->GetND(), M = species->GetNBead(), N = species->GetNPart();
    ModeType old = s->path->GetMode();
    s->path->SetMode(mode ? NEW_MODE : OLD_MODE);
    for (uint32_t p = 0; p < N; ++p)
        for (uint32_t b = 0; b < M; ++b) {
            const vec<double> &r = species->GetBead(p, b)->GetR();
            for (uint32_t d = 0; d < n_d; ++d) R[(p * M + b) * n_d + d] = r(d);
        }
    s->path->SetMode(old);
}

/// NEW-mode proposal: overwrite beads (p, b_first .. b_first+n-1) with newR[n][dim].
void ref_propose(void *h, int sp, int p, int b_first, int n, const double *newR) {
    RefSim *s = (RefSim *)h;
    auto species = s->path->GetSpecies()[sp];
    const uint32_t n_d = s->path->GetND();
    s->path->SetMode(NEW_MODE);
    for (int i = 0; i < n; ++i) {
        vec<double> r(n_d);
        for (uint32_t d = 0; d < n_d; ++d) r(d) = newR[i * n_d + d];
        species->GetBead(p, b_first + i)->SetR(r);
    }
}

/// What Bisect::Accept / Reject (bisect_class.h:24-36,127-139) and DisplaceParticle's do:
/// beads (p, b0..b1] commit or roll back r and rho_k, slices [b0,b1) commit / roll back the
/// species rho_k, and every action re-arms its rho_k flag.
void ref_finish_move(void *h, int sp, int p, int b0, int b1, int accept) {
    RefSim *s = (RefSim *)h;
    auto species = s->path->GetSpecies()[sp];
    for (int b = b0; b <= b1; ++b) {
        auto bead = species->GetBead(p, b);
        if (accept) {
            bead->StoreR();
            bead->StoreRhoK();
        } else {
            bead->RestoreR();
            bead->RestoreRhoK();
        }
    }
    for (int b = b0; b < b1; ++b) {
        if (accept)
            species->StoreRhoK(b);
        else
            species->RestoreRhoK(b);
    }
    for (auto &action : s->actions) {
        if (accept)
            action->Accept();
        else
            action->Reject();
    }
}

void ref_rhok(void *h, int sp, int mode, double *out) {
    RefSim *s = (RefSim *)h;
    auto species = s->path->GetSpecies()[sp];
    ModeType old = s->path->GetMode();
    s->path->SetMode(mode ? NEW_MODE : OLD_MODE);
    auto &rho = species->GetRhoK();
    const size_t n_k = s->path->ks.vecs.size();
    for (uint32_t b = 0; b < species->GetNBead(); ++b)
        for (size_t k = 0; k < n_k; ++k) {
            out[(b * n_k + k) * 2 + 0] = rho(b)(k).real();
            out[(b * n_k + k) * 2 + 1] = rho(b)(k).imag();
        }
    s->path->SetMode(old);
}

// ---- the Action API ------------------------------------------------------------------
int ref_n_actions(void *h) { return ((RefSim *)h)->actions.size(); }

double ref_dbeta(void *h, int a) {
    RefSim *s = (RefSim *)h;
    s->path->SetMode(NEW_MODE);
    return s->actions[a]->DActionDBeta();
}

double ref_potential(void *h, int a) {
    RefSim *s = (RefSim *)h;
    s->path->SetMode(NEW_MODE);
    return s->actions[a]->Potential();
}

double ref_get_action(void *h, int a, int mode, int b0, int b1, int n, const int *sp, const int *pi, int level) {
    RefSim *s = (RefSim *)h;
    std::vector<std::pair<std::shared_ptr<Species>, uint32_t>> particles;
    for (int i = 0; i < n; ++i) particles.push_back(std::make_pair(s->path->GetSpecies()[sp[i]], (uint32_t)pi[i]));
    s->path->SetMode(mode ? NEW_MODE : OLD_MODE);
    return s->actions[a]->GetAction(b0, b1, particles, level);
}

/// Action::GetActionGradient / GetActionLaplacian (pair_action_class.h:305-366).
void ref_action_gradient(void *h, int a, int mode, int b0, int b1, int n, const int *sp, const int *pi, int level, double *out) {
    RefSim *s = (RefSim *)h;
    std::vector<std::pair<std::shared_ptr<Species>, uint32_t>> particles;
    for (int i = 0; i < n; ++i) particles.push_back(std::make_pair(s->path->GetSpecies()[sp[i]], (uint32_t)pi[i]));
    s->path->SetMode(mode ? NEW_MODE : OLD_MODE);
    vec<double> g = s->actions[a]->GetActionGradient(b0, b1, particles, level);
    for (uint32_t d = 0; d < s->path->GetND(); ++d) out[d] = g(d);
}

double ref_action_laplacian(void *h, int a, int mode, int b0, int b1, int n, const int *sp, const int *pi, int level) {
    RefSim *s = (RefSim *)h;
    std::vector<std::pair<std::shared_ptr<Species>, uint32_t>> particles;
    for (int i = 0; i < n; ++i) particles.push_back(std::make_pair(s->path->GetSpecies()[sp[i]], (uint32_t)pi[i]));
    s->path->SetMode(mode ? NEW_MODE : OLD_MODE);
    return s->actions[a]->GetActionLaplacian(b0, b1, particles, level);
}

void ref_action_accept(void *h, int a, int accept) {
    RefSim *s = (RefSim *)h;
    if (accept)
        s->actions[a]->Accept();
    else
        s->actions[a]->Reject();
}

/// Per-pair kernels: which = 0 CalcU, 1 CalcdUdBeta, 2 CalcV.
void ref_calc_pair(void *h, int a, int which, int n, const double *r, const double *rp, const double *sv, int level, double *out) {
    PairAction *p = AsPair((RefSim *)h, a);
    for (int i = 0; i < n; ++i) {
        if (which == 0)
            out[i] = p->CalcU(r[i], rp[i], sv[i], level);
        else if (which == 1)
            out[i] = p->CalcdUdBeta(r[i], rp[i], sv[i], level);
        else
            out[i] = p->CalcV(r[i], rp[i], level);
    }
}

/// Path::DrDrpDrrp (path_class.h:137-150) for one pair and slice link.
void ref_dr_drp_drrp(void *h, int b0, int b1, int sa, int sb, int p0, int p1, double *out3) {
    RefSim *s = (RefSim *)h;
    s->path->DrDrpDrrp(b0, b1, s->path->GetSpecies()[sa], s->path->GetSpecies()[sb], p0, p1, out3[0], out3[1], out3[2]);
}

/// Long-range pieces: which = 0 CalcULong(b0,b1,level), 1 CalcdUdBetaLong, 2 CalcVLong.
double ref_calc_long(void *h, int a, int which, int b0, int b1, int level) {
    RefSim *s = (RefSim *)h;
    PairAction *p = AsPair(s, a);
    if (which == 0) return p->CalcULong(b0, b1, level);
    if (which == 1) return p->CalcdUdBetaLong();
    return p->CalcVLong();
}

void ref_set_mode(void *h, int mode) { ((RefSim *)h)->path->SetMode(mode ? NEW_MODE : OLD_MODE); }

// ---- observables ---------------------------------------------------------------------
int ref_n_observables(void *h) { return ((RefSim *)h)->observables.size(); }
void ref_observable_accumulate(void *h, int o) { ((RefSim *)h)->observables[o]->DoEvent(); }
void ref_observable_write(void *h, int o) { ((RefSim *)h)->observables[o]->Write(); }

/// Raw (un-normalised) g(r) counts currently accumulated.
int ref_gofr_counts(void *h, int o, double *y) {
    PairCorrelation *pc = dynamic_cast<PairCorrelation *>(((RefSim *)h)->observables[o].get());
    if (!pc) return -1;
    for (uint32_t i = 0; i < pc->gr.x.n_r; ++i) y[i] = pc->gr.y(i);
    return pc->gr.x.n_r;
}
/// LinearGrid::ReverseMap (observable_class.h:56-62) on arbitrary distances.
int ref_gofr_bins(void *h, int o, int n, const double *r, uint32_t *bins) {
    PairCorrelation *pc = dynamic_cast<PairCorrelation *>(((RefSim *)h)->observables[o].get());
    if (!pc) return -1;
    for (int i = 0; i < n; ++i) bins[i] = pc->gr.x.ReverseMap(r[i]);
    return 0;
}
int ref_sofk_sums(void *h, int o, double *sk) {
    StructureFactor *sf = dynamic_cast<StructureFactor *>(((RefSim *)h)->observables[o].get());
    if (!sf) return -1;
    for (size_t i = 0; i < sf->sk.size(); ++i) sk[i] = sf->sk(i);
    return sf->sk.size();
}
int ref_energy_sums(void *h, int o, double *energies, double *potentials) {
    Energy *e = dynamic_cast<Energy *>(((RefSim *)h)->observables[o].get());
    if (!e) return -1;
    for (size_t i = 0; i < e->energies.size(); ++i) energies[i] = e->energies(i);
    if (e->measure_potential)
        for (size_t i = 0; i < e->potentials.size(); ++i) potentials[i] = e->potentials(i);
    return e->energies.size();
}

// ---- moves (sampled runs; the reference's own RNG stream) ----------------------------
/// PermBisectIterative::UpdatePermTable (perm_bisect_iterative_class.h:10-30) for the window that starts at bead0.
int ref_perm_table(void *h, int m, int bead0, double *t_out) {
    RefSim *s = (RefSim *)h;
    PermBisectIterative *pb = dynamic_cast<PermBisectIterative *>(s->moves[m].get());
    if (!pb) return -1;
    s->path->SetMode(NEW_MODE);
    pb->bead0 = bead0;
    pb->UpdatePermTable();
    const uint32_t N = pb->species->GetNPart();
    for (uint32_t i = 0; i < N; ++i)
        for (uint32_t j = 0; j < N; ++j) t_out[i * N + j] = pb->t(i, j);
    return (int)N;
}

/// PermBisect's per-cycle-length counters (perm_bisect_class.h:26-27): attempts / accepts of 1-, 2-, ... particle cycles.
int ref_perm_counts(void *h, int m, int n_max, uint32_t *attempt, uint32_t *accept) {
    RefSim *s = (RefSim *)h;
    PermBisect *pb = dynamic_cast<PermBisect *>(s->moves[m].get());
    if (!pb) return -1;
    const int n = std::min<int>(n_max, pb->perm_attempt.size());
    for (int i = 0; i < n; ++i) {
        attempt[i] = pb->perm_attempt(i);
        accept[i] = pb->perm_accept(i);
    }
    return n;
}

/// The permutation as PathDump records it (path_dump_class.h:46-51): next[p] = label of the bead that follows
/// (p, n_bead - 1), prev[p] = label of the bead that precedes (p, 0).
void ref_permutation(void *h, int sp, int32_t *prev, int32_t *next) {
    RefSim *s = (RefSim *)h;
    std::shared_ptr<Species> species = s->path->GetSpecies()[sp];
    for (uint32_t p = 0; p < species->GetNPart(); ++p) {
        prev[p] = (int32_t)species->GetBead(p, 0)->GetPrevBead(1)->GetP();
        next[p] = (int32_t)species->GetBead(p, species->GetNBead() - 1)->GetNextBead(1)->GetP();
    }
}

int ref_n_moves(void *h) { return ((RefSim *)h)->moves.size(); }
void ref_move_do(void *h, int m, int n_times) {
    RefSim *s = (RefSim *)h;
    for (int i = 0; i < n_times; ++i) s->moves[m]->DoEvent();
}
/// Queue uniforms (each answers one RNG::UnifRand()) and normals (one RNG::NormRand() each) for the moves that follow.
void ref_inject_random(const double *uniforms, int n_u, const double *normals, int n_n) {
    for (int i = 0; i < n_u; ++i) pimc_inject::Uniforms().push_back(uniforms[i]);
    for (int i = 0; i < n_n; ++i) pimc_inject::Normals().push_back(normals[i]);
}
/// Numbers still queued (a move that rejected early leaves some); clear != 0 drops them.
void ref_inject_pending(int *n_u, int *n_n, int clear) {
    *n_u = (int)pimc_inject::Uniforms().size();
    *n_n = (int)pimc_inject::Normals().size();
    if (clear) {
        pimc_inject::Uniforms().clear();
        pimc_inject::Normals().clear();
    }
}
void ref_move_counts(void *h, int m, uint32_t *n_attempt, uint32_t *n_accept) {
    RefSim *s = (RefSim *)h;
    *n_attempt = s->moves[m]->n_attempt;
    *n_accept = s->moves[m]->n_accept;
}

// ---- captured output -----------------------------------------------------------------
/// Number of records and record length of a dataset the reference wrote/appended.
int ref_capture_shape(void *h, const char *name, int *n_records, int *record_len) {
    RefSim *s = (RefSim *)h;
    auto &file = CaptureStore()[s->out.file_name];
    auto it = file.find(PtabKey(name));
    if (it == file.end()) return -1;
    *n_records = it->second.size();
    *record_len = it->second.empty() ? 0 : it->second[0].size();
    return 0;
}
int ref_capture_read(void *h, const char *name, double *out) {
    RefSim *s = (RefSim *)h;
    auto &file = CaptureStore()[s->out.file_name];
    auto it = file.find(PtabKey(name));
    if (it == file.end()) return -1;
    size_t k = 0;
    for (auto &rec : it->second)
        for (double v : rec) out[k++] = v;
    return 0;
}

// ---- stand-alone pieces --------------------------------------------------------------
/// KSpace::Setup ordering (k_space_class.h:33-80) without building a Path.
int ref_kspace_standalone(int n_d, double L, double k_cut, int max_out, int *indices, double *mags) {
    KSpace ks;
    ks.n_d = n_d;
    ks.L = L;
    ks.cutoff = 0.;
    ks.Setup(k_cut);
    int n = ks.vecs.size();
    for (int k = 0; k < n && k < max_out; ++k) {
        for (int d = 0; d < n_d; ++d) indices[k * n_d + d] = ks.indices[k](d) - ks.max_index(d);
        mags[k] = ks.mags[k];
    }
    return n;
}

}  // extern "C"
