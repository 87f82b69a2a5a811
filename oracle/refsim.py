"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/_ref/libsimpimc_ref.so, i.e. the
reference's own classes (compiled in place from /root/reference/src against oracle/shim).

Used by oracle/make_golden.py to produce tests/golden/* and by tests that cross-check the
CPU restatement (oracle/pimc_oracle.cc) where the built library is present.  Never imported
by the product package.
"""
import ctypes as C
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libsimpimc_ref.so")
REF_LIB_FAST = os.path.join(HERE, "_ref", "libsimpimc_ref_fast.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_up = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


def available(fast=False):
    return os.path.exists(REF_LIB_FAST if fast else REF_LIB)


_libs = {}


DROPIN_LIB = os.path.join(os.path.dirname(HERE), "integration", "_build", "libsimpimc_dropin.so")


def dropin_available():
    return os.path.exists(DROPIN_LIB)


def _load(fast=False, dropin=False):
    key = "dropin" if dropin else bool(fast)
    if key in _libs:
        return _libs[key]
    lib = C.CDLL(DROPIN_LIB if dropin else (REF_LIB_FAST if fast else REF_LIB))
    lib.ref_create.restype = C.c_void_p
    lib.ref_create.argtypes = [C.c_char_p, C.c_int, C.c_int]
    lib.ref_destroy.argtypes = [C.c_void_p]
    for name in ("ref_n_species", "ref_n_bead", "ref_n_k", "ref_n_actions", "ref_n_observables", "ref_n_moves"):
        getattr(lib, name).restype = C.c_int
        getattr(lib, name).argtypes = [C.c_void_p]
    lib.ref_n_part.restype = C.c_int
    lib.ref_n_part.argtypes = [C.c_void_p, C.c_int]
    lib.ref_tau.restype = C.c_double
    lib.ref_tau.argtypes = [C.c_void_p]
    lib.ref_kspace.argtypes = [C.c_void_p, _ip, _dp, _dp, _ip]
    lib.ref_set_positions.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.ref_get_positions.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
    lib.ref_propose.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
    lib.ref_finish_move.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ref_rhok.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
    lib.ref_dbeta.restype = C.c_double
    lib.ref_dbeta.argtypes = [C.c_void_p, C.c_int]
    lib.ref_potential.restype = C.c_double
    lib.ref_potential.argtypes = [C.c_void_p, C.c_int]
    lib.ref_get_action.restype = C.c_double
    lib.ref_get_action.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, C.c_int]
    lib.ref_action_accept.argtypes = [C.c_void_p, C.c_int, C.c_int]
    if hasattr(lib, "ref_perm_counts"):
        lib.ref_perm_counts.restype = C.c_int
        lib.ref_perm_counts.argtypes = [C.c_void_p, C.c_int, C.c_int, _up, _up]
    if hasattr(lib, "ref_perm_table"):
        lib.ref_perm_table.restype = C.c_int
        lib.ref_perm_table.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
    if hasattr(lib, "ref_action_gradient"):
        lib.ref_action_gradient.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, C.c_int, _dp]
        lib.ref_action_laplacian.restype = C.c_double
        lib.ref_action_laplacian.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, C.c_int]
    lib.ref_calc_pair.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _dp]
    lib.ref_dr_drp_drrp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
    lib.ref_calc_long.restype = C.c_double
    lib.ref_calc_long.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ref_set_mode.argtypes = [C.c_void_p, C.c_int]
    lib.ref_observable_accumulate.argtypes = [C.c_void_p, C.c_int]
    lib.ref_observable_write.argtypes = [C.c_void_p, C.c_int]
    lib.ref_gofr_counts.restype = C.c_int
    lib.ref_gofr_counts.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.ref_gofr_bins.restype = C.c_int
    lib.ref_gofr_bins.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _up]
    lib.ref_sofk_sums.restype = C.c_int
    lib.ref_sofk_sums.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.ref_energy_sums.restype = C.c_int
    lib.ref_energy_sums.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
    lib.ref_move_do.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.ref_move_counts.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    if hasattr(lib, "ref_permutation"):
        lib.ref_permutation.argtypes = [C.c_void_p, C.c_int, np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS"),
                                        np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")]
    if hasattr(lib, "ref_inject_random"):
        lib.ref_inject_random.argtypes = [_dp, C.c_int, _dp, C.c_int]
        lib.ref_inject_pending.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
    lib.ref_capture_shape.restype = C.c_int
    lib.ref_capture_shape.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.ref_capture_read.restype = C.c_int
    lib.ref_capture_read.argtypes = [C.c_void_p, C.c_char_p, _dp]
    lib.ref_kspace_standalone.restype = C.c_int
    lib.ref_kspace_standalone.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, _ip, _dp]
    _libs[key] = lib
    return lib


def kspace_standalone(n_d, L, k_cut, max_out=100000):
    lib = _load()
    idx = np.zeros((max_out, n_d), dtype=np.int32)
    mags = np.zeros(max_out)
    n = lib.ref_kspace_standalone(n_d, L, k_cut, max_out, idx, mags)
    return idx[:n].copy(), mags[:n].copy()


def system_xml(cfg, table_files, workdir):
    """Reference-format input file for a simpimc_b200.system.SystemConfig (same attribute
    names as inputs/e-gas/e-gas.xml)."""
    lines = ["<Input>", '  <IO output_prefix="%s" />' % os.path.join(workdir, "out")]
    sysattrs = 'n_d="%d" n_bead="%d" beta="%.17g" pbc="%d"' % (cfg.n_d, cfg.n_bead, cfg.beta, 1 if cfg.pbc else 0)
    if cfg.pbc:
        sysattrs += ' L="%.17g"' % cfg.L
    if cfg.k_cut is not None:
        sysattrs += ' k_cut="%.17g"' % cfg.k_cut
    lines.append("  <System %s />" % sysattrs)
    lines.append("  <Particles>")
    for sp in cfg.species:
        lines.append('    <Species name="%s" n_part="%d" lambda="%.17g" fermi="0" fixed_node="0" init_type="Random" />'
                     % (sp.name, sp.n_part, sp.lam))
    lines.append("  </Particles>")
    lines.append("  <Actions>")
    for a in cfg.actions:
        if a.type == "Kinetic":
            lines.append('    <Action name="%s" type="Kinetic" species="%s" n_images="%d" />' % (a.name, a.species_a, a.n_images))
            continue
        attrs = 'name="%s" type="%s" file="%s" species_a="%s" species_b="%s" max_level="%d" use_long_range="%d" n_images="0"' % (
            a.name, a.type, table_files[a.name], a.species_a, a.species_b, a.max_level, 1 if a.use_long_range else 0)
        if a.use_long_range and a.k_cut is not None:
            attrs += ' k_cut="%.17g"' % a.k_cut
        if a.type == "DavidPairAction":
            attrs += ' n_order="%d"' % a.n_order
        if a.type == "BarePairAction" and a.is_coulomb:
            attrs += ' is_coulomb="1"'
        lines.append("    <Action %s />" % attrs)
    lines.append("  </Actions>")
    lines.append("  <Moves>")
    for m in cfg.moves:
        if m["type"] in ("Bisect", "PermBisectIterative"):
            lines.append('    <Move name="%s" type="%s" species="%s" n_level="%d" n_images="%d" />'
                         % (m["name"], m["type"], m["species"], m["n_level"], m.get("n_images", 0)))
        else:
            lines.append('    <Move name="%s" type="DisplaceParticle" species="%s" step_size="%.17g" />'
                         % (m["name"], m["species"], m["step_size"]))
    lines.append("  </Moves>")
    lines.append("  <Observables>")
    for o in cfg.observables:
        if o["type"] == "Energy":
            lines.append('    <Observable name="%s" type="Energy" measure_potential="%d" />' % (o["name"], o.get("measure_potential", 0)))
        elif o["type"] == "PairCorrelation":
            lines.append('    <Observable name="%s" type="PairCorrelation" species_a="%s" species_b="%s" r_min="%.17g" r_max="%.17g" n_r="%d" />'
                         % (o["name"], o["species_a"], o["species_b"], o["r_min"], o["r_max"], o["n_r"]))
        elif o["type"] == "StructureFactor":
            lines.append('    <Observable name="%s" type="StructureFactor" species_a="%s" species_b="%s" k_cut="%.17g" />'
                         % (o["name"], o["species_a"], o["species_b"], o["k_cut"]))
    lines.append("  </Observables>")
    lines.append("</Input>")
    return "\n".join(lines) + "\n"


class RefSim:
    """One reference Path + actions + moves + observables, built from a SystemConfig."""

    def __init__(self, cfg, seed=12345, workdir=None, fast=False, quiet=True, dropin=False):
        from simpimc_b200 import tables as T
        self.lib = _load(fast, dropin)
        self.cfg = cfg
        self._tmp = None
        if workdir is None:
            self._tmp = tempfile.TemporaryDirectory(prefix="refsim_")
            workdir = self._tmp.name
        table_files = {}
        for a in cfg.actions:
            if a.type == "Kinetic":
                continue
            fn = os.path.join(workdir, a.name + ".ptab")
            T.write_ptab(fn, a.table)
            table_files[a.name] = fn
        xml = system_xml(cfg, table_files, workdir)
        self.xml_file = os.path.join(workdir, "input.xml")
        with open(self.xml_file, "w") as f:
            f.write(xml)
        self.h = self.lib.ref_create(self.xml_file.encode(), seed, 1 if quiet else 0)
        self.n_d = cfg.n_d
        self.M = cfg.n_bead

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None

    # positions are [particle][bead][dim] on this side of the fence
    def set_positions(self, sp, R):
        self.lib.ref_set_positions(self.h, sp, np.ascontiguousarray(R, dtype=np.float64))

    def get_positions(self, sp, mode=1):
        N = self.lib.ref_n_part(self.h, sp)
        R = np.zeros((N, self.M, self.n_d))
        self.lib.ref_get_positions(self.h, sp, mode, R)
        return R

    def n_k(self):
        return self.lib.ref_n_k(self.h)

    def kspace(self):
        n = self.n_k()
        idx = np.zeros((n, self.n_d), dtype=np.int32)
        vecs = np.zeros((n, self.n_d))
        mags = np.zeros(n)
        mx = np.zeros(self.n_d, dtype=np.int32)
        self.lib.ref_kspace(self.h, idx, vecs, mags, mx)
        return idx, vecs, mags, mx

    def rhok(self, sp, mode=1):
        out = np.zeros((self.M, self.n_k(), 2))
        self.lib.ref_rhok(self.h, sp, mode, out)
        return out[..., 0] + 1j * out[..., 1]

    def dbeta(self, a):
        return self.lib.ref_dbeta(self.h, a)

    def potential(self, a):
        return self.lib.ref_potential(self.h, a)

    def get_action(self, a, mode, b0, b1, particles, level):
        sp = np.array([p[0] for p in particles], dtype=np.int32)
        pi = np.array([p[1] for p in particles], dtype=np.int32)
        return self.lib.ref_get_action(self.h, a, mode, b0, b1, len(particles), sp, pi, level)

    def action_gradient(self, a, mode, b0, b1, particles, level):
        sp = np.array([p[0] for p in particles], dtype=np.int32)
        pi = np.array([p[1] for p in particles], dtype=np.int32)
        out = np.zeros(3)
        self.lib.ref_action_gradient(self.h, a, mode, b0, b1, len(particles), sp, pi, level, out)
        return out

    def action_laplacian(self, a, mode, b0, b1, particles, level):
        sp = np.array([p[0] for p in particles], dtype=np.int32)
        pi = np.array([p[1] for p in particles], dtype=np.int32)
        return self.lib.ref_action_laplacian(self.h, a, mode, b0, b1, len(particles), sp, pi, level)

    def calc_pair(self, a, which, r, rp, s, level=0):
        r = np.ascontiguousarray(r, dtype=np.float64)
        rp = np.ascontiguousarray(rp, dtype=np.float64)
        s = np.ascontiguousarray(s, dtype=np.float64)
        out = np.zeros_like(r)
        self.lib.ref_calc_pair(self.h, a, which, len(r), r, rp, s, level, out)
        return out

    def dr_drp_drrp(self, b0, b1, sa, sb, p0, p1):
        out = np.zeros(3)
        self.lib.ref_dr_drp_drrp(self.h, b0, b1, sa, sb, p0, p1, out)
        return out

    def calc_long(self, a, which, b0=0, b1=0, level=0):
        return self.lib.ref_calc_long(self.h, a, which, b0, b1, level)

    def set_mode(self, mode):
        self.lib.ref_set_mode(self.h, mode)

    def propose(self, sp, p, b_first, newR):
        newR = np.ascontiguousarray(newR, dtype=np.float64)
        self.lib.ref_propose(self.h, sp, p, b_first, newR.shape[0], newR)

    def finish_move(self, sp, p, b0, b1, accept):
        self.lib.ref_finish_move(self.h, sp, p, b0, b1, 1 if accept else 0)

    def observable_accumulate(self, o):
        self.lib.ref_observable_accumulate(self.h, o)

    def observable_write(self, o):
        self.lib.ref_observable_write(self.h, o)

    def gofr_counts(self, o, n_r):
        y = np.zeros(n_r)
        n = self.lib.ref_gofr_counts(self.h, o, y)
        assert n == n_r
        return y

    def gofr_bins(self, o, r):
        r = np.ascontiguousarray(r, dtype=np.float64)
        bins = np.zeros(len(r), dtype=np.uint32)
        assert self.lib.ref_gofr_bins(self.h, o, len(r), r, bins) == 0
        return bins

    def sofk_sums(self, o):
        sk = np.zeros(self.n_k())
        n = self.lib.ref_sofk_sums(self.h, o, sk)
        return sk[:n]

    def energy_sums(self, o):
        n_a = self.lib.ref_n_actions(self.h)
        e = np.zeros(n_a)
        v = np.zeros(n_a)
        self.lib.ref_energy_sums(self.h, o, e, v)
        return e, v

    def perm_table(self, m, bead0, n_part):
        """PermBisectIterative::UpdatePermTable of move m for the window starting at bead0."""
        t = np.zeros((n_part, n_part))
        n = self.lib.ref_perm_table(self.h, m, bead0, t)
        assert n == n_part, n
        return t

    def perm_counts(self, m, n_max=8):
        """(attempts, accepts) per cycle length 1, 2, ... of a PermBisect move."""
        att = np.zeros(n_max, dtype=np.uint32)
        acc = np.zeros(n_max, dtype=np.uint32)
        n = self.lib.ref_perm_counts(self.h, m, n_max, att, acc)
        assert n >= 0
        return att[:n], acc[:n]

    def permutation(self, sp):
        """(prev, next): labels of the beads before (p, 0) and after (p, n_bead - 1) (path_dump_class.h:46-51)."""
        N = self.lib.ref_n_part(self.h, sp)
        prev, nxt = np.zeros(N, dtype=np.int32), np.zeros(N, dtype=np.int32)
        self.lib.ref_permutation(self.h, sp, prev, nxt)
        return prev, nxt

    def move_do(self, m, n_times=1):
        self.lib.ref_move_do(self.h, m, n_times)

    def inject_random(self, uniforms, normals):
        """Queue the numbers the next RNG::UnifRand() / NormRand() calls of the reference will return
        (process-wide queues; the real std::mt19937 stream resumes when they run dry)."""
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        n = np.ascontiguousarray(normals, dtype=np.float64)
        self.lib.ref_inject_random(u, len(u), n, len(n))

    def inject_pending(self, clear=True):
        a, b = C.c_int(), C.c_int()
        self.lib.ref_inject_pending(C.byref(a), C.byref(b), 1 if clear else 0)
        return a.value, b.value

    def move_counts(self, m):
        a, b = C.c_uint32(), C.c_uint32()
        self.lib.ref_move_counts(self.h, m, C.byref(a), C.byref(b))
        return a.value, b.value

    def capture(self, name):
        n, ln = C.c_int(), C.c_int()
        if self.lib.ref_capture_shape(self.h, name.encode(), C.byref(n), C.byref(ln)) != 0:
            return None
        out = np.zeros((n.value, ln.value))
        self.lib.ref_capture_read(self.h, name.encode(), out)
        return out
