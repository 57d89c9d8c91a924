"""``Potential`` -- host-side mirror of ``quippy.potential.Potential`` over libgapb200.so.

Mirrors the reference's Python interface for the GAP path (quippy/quippy/potential.py:64-331):
same constructor arguments (``args_str``, ``param_str`` / ``param_filename``, ``calc_args``), the ASE
``Calculator``-style ``calculate(atoms, properties)`` filling ``results`` with ``energy``, ``forces``
(N,3), ``stress`` (Voigt, ``-virial/V``, :281-284), ``energies``, ``stresses``, and the extras in
``extra_results``; errors surface as ``RuntimeError`` (quippy maps Fortran aborts the same way).

All arithmetic happens in the CUDA library behind the C ABI of include/gap_b200.h; this module only
marshals numpy arrays into the Fortran-compatible layouts.  There is no CPU fallback: if the shared
library or a CUDA device is missing, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBNAME = os.path.join(_HERE, "libgapb200.so")
_lib = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)

# every symbol include/gap_b200.h declares: (restype, argtypes)
ABI = {
    "gap_potential_filename_initialise": (C.c_int, [C.POINTER(C.c_void_p), C.c_char_p, C.c_char_p, C.c_int]),
    "gap_potential_initialise": (C.c_int, [C.POINTER(C.c_void_p), C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]),
    "gap_potential_finalise": (None, [C.c_void_p]),
    "gap_potential_cutoff": (C.c_double, [C.c_void_p]),
    "gap_potential_print": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "gap_potential_set_partition": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "gap_comm_get_unique_id": (C.c_int, [C.c_char_p]),
    "gap_potential_set_comm": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int]),
    "gap_potential_comm_info": (C.c_int, [C.c_void_p, c_ip, c_ip, C.c_char_p, C.c_size_t]),
    "gap_potential_comm_timing": (C.c_int, [C.c_void_p, c_dp, c_dp]),
    "gap_potential_set_deterministic": (C.c_int, [C.c_void_p, C.c_int]),
    "gap_potential_set_cutoff_skin": (C.c_int, [C.c_void_p, C.c_double]),
    "gap_potential_connect_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_long), C.POINTER(C.c_long)]),
    "gap_potential_calc": (C.c_int, [C.c_void_p, C.c_int, c_dp, c_ip, c_dp, c_ip, C.c_char_p, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "gap_potential_set_atom_mask": (C.c_int, [C.c_void_p, C.c_int, c_ip]),
    "gap_potential_set_resid": (C.c_int, [C.c_void_p, C.c_int, c_ip]),
    "gap_potential_get_energy_per_coordinate": (C.c_int, [C.c_void_p, c_dp]),
    "gap_potential_get_local_gap_variance": (C.c_int, [C.c_void_p, C.c_int, c_dp, c_dp]),
    "gap_potential_calc_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, c_dp, c_ip, C.c_char_p, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gap_md_run": (C.c_int, [C.c_void_p, C.c_int, c_dp, c_dp, c_ip, c_dp, c_dp, c_ip, C.c_double, C.c_int, C.c_char_p, c_dp, c_dp]),
    "gap_md_run_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_dp, c_ip, C.c_double, C.c_int,
                                    C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, c_dp, c_dp, C.c_void_p]),
    "quip_lammps_api_version": (C.c_int, []),
    "quip_lammps_potential_initialise": (None, [c_ip, c_ip, c_dp, C.c_char_p, c_ip, C.c_char_p, c_ip]),
    "quip_lammps_wrapper": (None, [c_ip, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip, c_dp, c_ip, c_ip, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "gap_potential_calc_device_enqueue": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, c_dp, c_ip, C.c_char_p, C.c_int,
                                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gap_potential_calc_device_verify": (C.c_int, [C.c_void_p, c_ip]),
    "gap_b200_wrapper_simple": (C.c_int, [C.c_char_p, c_ip, c_dp, c_ip, c_dp, c_dp, c_dp, c_dp]),
    "gap_calc_connect": (C.c_int, [C.c_void_p, C.c_int, c_dp, c_dp, c_ip, C.c_double, c_ip]),
    "gap_get_connect": (C.c_int, [C.c_void_p, c_ip, c_ip, c_ip, c_dp]),
    "gap_descriptor_calc": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dp, c_ip, c_dp, c_ip, c_ip, c_ip, c_dp, c_ip]),
    "gap_gp_predict": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dp, c_dp, c_dp]),
    "gap_potential_n_coordinate": (C.c_int, [C.c_void_p]),
    "gap_potential_launch_count": (C.c_long, [C.c_void_p]),
    "gap_potential_last_timings": (C.c_int, [C.c_void_p, c_dp]),
    "gap_potential_set_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "gap_model_describe": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]),
    "gap_last_error": (C.c_char_p, []),
}


def load_library():
    """Load libgapb200.so (built in-tree by ``__graft_entry__.build()``); fail loudly when absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIBNAME):
            raise RuntimeError("libgapb200.so is missing (%s): run `python -c 'import __graft_entry__ as g; g.build()'`; "
                               "quip_b200 has no CPU fallback" % _LIBNAME)
        lib = C.CDLL(_LIBNAME)
        for name, (res, args) in ABI.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_ip) if a is not None else None


def _check(rc):
    if rc != 0:
        raise RuntimeError(load_library().gap_last_error().decode("utf-8", "replace"))


def model_describe(param_str=None, param_filename=None, args_str="", base_dir="."):
    """Host-only parse of a GAP model through the product's C++ loader (no GPU needed); returns the description text."""
    if param_filename is not None:
        with open(param_filename) as fh:
            param_str = fh.read()
        base_dir = os.path.dirname(os.path.abspath(param_filename))
    buf = C.create_string_buffer(1 << 20)
    _check(load_library().gap_model_describe(args_str.encode(), param_str.encode(), base_dir.encode(), buf, len(buf)))
    return buf.value.decode()


# src/libAtoms/Units.f95:56-68 (the constants of the default unit system) and PeriodicTable.f95:90 (IUPAC 2013 masses)
_ELECTRONMASS_GPERMOL, _HARTREE, _BOHR, _HBAR_EVSEC = 5.48579903e-4, 27.2113961, 0.529177249, 6.5821220e-16
_AU_FS = 1.0 / (1.0 / (_HBAR_EVSEC / _HARTREE)) * 1e15
MASSCONVERT = 1.0 / _ELECTRONMASS_GPERMOL * _HARTREE * _AU_FS * _AU_FS / (_BOHR * _BOHR)
_AMU = {1: 1.008, 2: 4.002602, 3: 6.94, 4: 9.0121831, 5: 10.81, 6: 12.011, 7: 14.007, 8: 15.999, 9: 18.998403163, 10: 20.1797, 11: 22.98976928,
        12: 24.305, 13: 26.9815385, 14: 28.085, 15: 30.973761998, 16: 32.06, 17: 35.45, 18: 39.948, 22: 47.867, 23: 50.9415, 26: 55.845,
        28: 58.6934, 29: 63.546, 32: 72.63, 41: 92.90637, 42: 95.95, 73: 180.94788, 74: 183.84}


def element_masses(Z):
    """``ElementMass(Z)`` in QUIP's internal units (amu * MASSCONVERT)."""
    try:
        return np.array([_AMU[int(z)] for z in Z]) * MASSCONVERT
    except KeyError as e:
        raise RuntimeError("element_masses: no mass tabulated for Z=%s; pass masses= explicitly" % e)


def key_val_dict_to_str(d):
    return " ".join("%s=%s" % (k, v) if v is not None else str(k) for k, v in d.items())


def _calc_options(args_str):
    """key=value pairs of a calc args string (ParamReader grammar, no quoting needed for the keys used here)"""
    out = {}
    for tok in (args_str or "").split():
        if "=" in tok:
            k, v = tok.split("=", 1)
            out[k] = v.strip("{}'\"")
    return out


def _geometry(atoms):
    # the arrays are only read: use the Atoms object's own arrays when it exposes them (get_positions() returns a copy)
    pos = np.ascontiguousarray(atoms.positions if hasattr(atoms, "positions") else atoms.get_positions(), dtype=np.float64)
    Z = np.ascontiguousarray(atoms.numbers if hasattr(atoms, "numbers") else atoms.get_atomic_numbers(), dtype=np.int32)
    lat = np.ascontiguousarray(np.asarray(atoms.get_cell(), dtype=np.float64).reshape(9))  # rows = vectors = Fortran columns
    pbc = np.ascontiguousarray(np.asarray(atoms.get_pbc(), dtype=bool).astype(np.int32))
    return pos, Z, lat, pbc


class Potential:
    implemented_properties = ["energy", "free_energy", "forces", "stress", "energies", "stresses", "virial", "local_energy",
                              "local_virial", "force"]

    def __init__(self, args_str="", pot1=None, pot2=None, param_str=None, param_filename=None, atoms=None,
                 calculation_always_required=False, calc_args=None, device=0, **kwargs):
        if pot1 is not None or pot2 is not None:
            raise RuntimeError("Potential: Sum potentials (pot1/pot2) are outside the scope of the B200 GAP path")
        lib = load_library()
        self._h = C.c_void_p()
        if param_filename is not None:
            _check(lib.gap_potential_filename_initialise(C.byref(self._h), args_str.encode(), os.fspath(param_filename).encode(),
                                                         int(device)))
        elif param_str is not None:
            _check(lib.gap_potential_initialise(C.byref(self._h), args_str.encode(), param_str.encode(),
                                                kwargs.pop("base_dir", ".").encode(), int(device)))
        else:
            raise RuntimeError("Potential: one of param_filename / param_str is required")
        self.args_str = args_str
        if isinstance(calc_args, dict):
            calc_args = key_val_dict_to_str(calc_args)
        self.calc_args = calc_args or ""
        self.calculation_always_required = calculation_always_required
        self.results = {}
        self.extra_results = {"config": {}, "atoms": {}}
        self.atoms = None
        if atoms is not None:
            atoms.calc = self

    # ---- reference API ------------------------------------------------------------------------
    def cutoff(self):
        return load_library().gap_potential_cutoff(self._h)

    def print_(self):
        buf = C.create_string_buffer(16384)
        _check(load_library().gap_potential_print(self._h, buf, len(buf)))
        return buf.value.decode()

    def set_partition(self, rank, n_ranks):
        _check(load_library().gap_potential_set_partition(self._h, int(rank), int(n_ranks)))

    def set_deterministic(self, on=True):
        """Bitwise reproducible forces: pair forces are summed per receiving atom in a fixed order instead of with FP64 atomics."""
        _check(load_library().gap_potential_set_deterministic(self._h, int(bool(on))))

    def set_cutoff_skin(self, cutoff_skin):
        """``at%cutoff_skin`` (Connection.f95:1085-1128): build the neighbour list out to cutoff + skin and reuse it while no atom has moved
        more than skin / 2."""
        _check(load_library().gap_potential_set_cutoff_skin(self._h, float(cutoff_skin)))

    def connect_stats(self):
        a, b = C.c_long(0), C.c_long(0)
        load_library().gap_potential_connect_stats(self._h, C.byref(a), C.byref(b))
        return {"rebuilds": a.value, "reuses": b.value}

    def set_comm(self, comm_id, rank, n_ranks):
        """Collective: join the communicator ``comm_id`` (128 bytes from :func:`comm_unique_id` on one rank, distributed by the
        host); from then on every calc returns the totals over the ranks (``IPModel_GAP.f95:538-556`` inside the library)."""
        _check(load_library().gap_potential_set_comm(self._h, comm_id, int(rank), int(n_ranks)))

    def comm_info(self):
        r, n = C.c_int(0), C.c_int(1)
        buf = C.create_string_buffer(16)
        load_library().gap_potential_comm_info(self._h, C.byref(r), C.byref(n), buf, len(buf))
        w, t = C.c_double(0.0), C.c_double(0.0)
        load_library().gap_potential_comm_timing(self._h, C.byref(w), C.byref(t))
        return {"rank": r.value, "n_ranks": n.value, "transport": buf.value.decode(), "last_wait_us": w.value, "last_sum_us": t.value}

    def calc(self, atoms, energy=True, force=False, virial=False, local_energy=False, local_virial=False, args_str="", out_force=None):
        """``calc(pot, at, energy, force, virial, local_energy, local_virial, args_str)`` (Potential.f95:803).
        Returns a dict with the requested quantities; ``force`` is (N,3), ``virial`` (3,3), ``local_virial`` (N,9).
        ``out_force``: a caller-owned C-contiguous float64 (N,3) array the forces are written into (the Fortran caller owns its
        output arrays too); page-locked arrays -- inputs and this one -- are transferred without a staging copy."""
        pos, Z, lat, pbc = _geometry(atoms)
        N = len(Z)
        e = np.zeros(1)
        if force and out_force is not None:
            if out_force.shape != (N, 3) or out_force.dtype != np.float64 or not out_force.flags.c_contiguous:
                raise RuntimeError("calc: out_force must be a C-contiguous float64 array of shape (N, 3)")
            f = out_force
        else:
            f = np.zeros((N, 3)) if force else None
        v = np.zeros((3, 3), order="F") if virial else None
        le = np.zeros(N) if local_energy else None
        lv = np.zeros((N, 9)) if local_virial else None
        full_args = (self.calc_args + " " + (args_str or "")).strip()
        opts = _calc_options(full_args)
        lib = load_library()
        if "atom_mask_name" in opts and opts["atom_mask_name"] != "NONE":
            # the reference reads the logical Atoms property of that name (IPModel_GAP.f95:344-346)
            arrays = getattr(atoms, "arrays", {})
            if opts["atom_mask_name"] not in arrays:
                raise RuntimeError("IPModel_GAP_Calc did not find %s property in the atoms object." % opts["atom_mask_name"])
            mask = np.ascontiguousarray(np.asarray(arrays[opts["atom_mask_name"]]).astype(bool).astype(np.int32))
            _check(lib.gap_potential_set_atom_mask(self._h, N, _ip(mask)))
        arrays = getattr(atoms, "arrays", None)
        if arrays and "resid" in arrays:  # residue ids for distance_2b only_intra / only_inter (the property the descriptor's resid_name names)
            rid = np.ascontiguousarray(arrays["resid"], dtype=np.int32)
            _check(lib.gap_potential_set_resid(self._h, N, _ip(rid)))
        _check(lib.gap_potential_calc(self._h, N, _dp(pos), _ip(Z), _dp(lat), _ip(pbc), full_args.encode(), _dp(e),
                                      _dp(le), _dp(f), _dp(v), _dp(lv)))
        out = {"energy": float(e[0])}
        if opts.get("energy_per_coordinate"):  # returned inside the Atoms object by the reference (:573)
            epc = np.zeros(self.n_coordinate)
            _check(lib.gap_potential_get_energy_per_coordinate(self._h, _dp(epc)))
            out[opts["energy_per_coordinate"]] = epc
        if opts.get("local_gap_variance"):  # (:558-571)
            lgv = np.zeros(N)
            gvg = np.zeros((N, 3)) if (force or virial or local_virial) else None
            _check(lib.gap_potential_get_local_gap_variance(self._h, N, _dp(lgv), _dp(gvg)))
            out[opts["local_gap_variance"]] = lgv
            if gvg is not None:
                out["gap_variance_gradient"] = gvg
        if force:
            out["force"] = f
        if virial:
            out["virial"] = np.array(v)
        if local_energy:
            out["local_energy"] = le
        if local_virial:
            out["local_virial"] = lv
        return out

    def calculate(self, atoms=None, properties=None, system_changes=None, forces=None, virial=None, local_energy=None,
                  local_virial=None, vol_per_atom=None, calc_args=None, **kwargs):
        """ASE-calculator entry point (quippy/quippy/potential.py:177-331)."""
        properties = list(set(["energy", "forces"] + list(properties or [])))
        for p in properties:
            if p not in self.implemented_properties:
                raise RuntimeError("Don't know how to calculate property '%s'" % p)
        if atoms is not None:
            self.atoms = atoms
        args_str = ""
        if calc_args is not None:
            args_str += " " + (key_val_dict_to_str(calc_args) if isinstance(calc_args, dict) else calc_args)
        if kwargs:
            args_str += " " + key_val_dict_to_str(kwargs)
        want_v = "virial" in properties or "stress" in properties or virial is not None
        want_lv = "local_virial" in properties or "stresses" in properties or local_virial is not None
        want_le = "energies" in properties or "local_energy" in properties or local_energy is not None
        r = self.calc(self.atoms, energy=True, force=True, virial=want_v, local_energy=want_le, local_virial=want_lv,
                      args_str=args_str)
        self.results = {"energy": r["energy"], "free_energy": r["energy"], "forces": r["force"]}
        self.extra_results = {"config": {}, "atoms": {}}
        if want_v:
            stress = -r["virial"] / self.atoms.get_volume()
            self.results["stress"] = np.array([stress[0, 0], stress[1, 1], stress[2, 2], stress[1, 2], stress[0, 2], stress[0, 1]])
            self.extra_results["config"]["virial"] = r["virial"].copy()
        if want_le:
            self.results["energies"] = r["local_energy"].copy()
            self.extra_results["atoms"]["local_energy"] = r["local_energy"].copy()
        if want_lv:
            self.extra_results["atoms"]["local_virial"] = r["local_virial"].copy()
            if "stresses" in properties:
                v_atom = self.atoms.get_volume() / len(self.atoms) if vol_per_atom is None else float(vol_per_atom)
                self.results["stresses"] = -r["local_virial"].reshape((len(self.atoms), 3, 3), order="F") / v_atom
        return self.results

    def get_potential_energy(self, atoms=None):
        return self.calculate(atoms, ["energy"])["energy"]

    def get_forces(self, atoms=None):
        return self.calculate(atoms, ["forces"])["forces"]

    def get_stress(self, atoms=None):
        return self.calculate(atoms, ["stress"])["stress"]

    def get_virial(self, atoms=None):
        self.calculate(atoms, ["stress"])
        return self.extra_results["config"]["virial"]

    # ---- stage-level entry points (parity tests, profiling) ---------------------------------------
    def calc_connect(self, atoms, cutoff=None):
        """Full neighbour list of ``atoms`` as (offsets[N+1], j[n], shift[n,3], distance[n])."""
        pos, Z, lat, pbc = _geometry(atoms)
        N = len(Z)
        n = C.c_int(0)
        cutoff = self.cutoff() if cutoff is None else float(cutoff)
        _check(load_library().gap_calc_connect(self._h, N, _dp(pos), _dp(lat), _ip(pbc), cutoff, C.byref(n)))
        off = np.zeros(N + 1, dtype=np.int32)
        j = np.zeros(max(n.value, 1), dtype=np.int32)
        s = np.zeros((max(n.value, 1), 3), dtype=np.int32)
        d = np.zeros(max(n.value, 1))
        _check(load_library().gap_get_connect(self._h, _ip(off), _ip(j), _ip(s), _dp(d)))
        return off, j[:n.value], s[:n.value], d[:n.value]

    def descriptor_calc(self, atoms, i_coord=0):
        pos, Z, lat, pbc = _geometry(atoms)
        N = len(Z)
        nd, d = C.c_int(0), C.c_int(0)
        lib = load_library()
        _check(lib.gap_descriptor_calc(self._h, i_coord, N, _dp(pos), _ip(Z), _dp(lat), _ip(pbc), C.byref(nd), C.byref(d), None, None))
        x = np.zeros((max(nd.value, 1), d.value))
        ci = np.zeros(max(nd.value, 1), dtype=np.int32)
        _check(lib.gap_descriptor_calc(self._h, i_coord, N, _dp(pos), _ip(Z), _dp(lat), _ip(pbc), C.byref(nd), C.byref(d), _dp(x), _ip(ci)))
        return x[:nd.value], ci[:nd.value]

    def gp_predict(self, i_coord, x, grad=True):
        x = np.ascontiguousarray(x, dtype=np.float64)
        n, d = x.shape
        e = np.zeros(n)
        g = np.zeros((n, d)) if grad else None
        _check(load_library().gap_gp_predict(self._h, i_coord, n, _dp(x), _dp(e), _dp(g)))
        return e, g

    def run(self, atoms, velocities, dt=1.0, n_steps=10, masses=None, args_str=""):
        """``DynamicalSystem_run(ds, pot, dt, n_steps)`` (Potential.f95:2304): NVE velocity Verlet, everything resident on the GPU,
        neighbour list rebuilt every step.  Updates ``atoms.positions`` in place; returns (velocities, epot[n_steps+1], ekin[n_steps+1])."""
        pos, Z, lat, pbc = _geometry(atoms)
        N = len(Z)
        pos = pos.copy()
        vel = np.ascontiguousarray(velocities, dtype=np.float64).copy()
        m = np.ascontiguousarray(element_masses(Z) if masses is None else masses, dtype=np.float64)
        ep, ek = np.zeros(n_steps + 1), np.zeros(n_steps + 1)
        _check(load_library().gap_md_run(self._h, N, _dp(pos), _dp(vel), _ip(Z), _dp(m), _dp(lat), _ip(pbc), float(dt), int(n_steps),
                                         (self.calc_args + " " + args_str).strip().encode(), _dp(ep), _dp(ek)))
        atoms.positions[...] = pos
        return vel, ep, ek

    def calc_device(self, N, d_pos_ptr, d_Z_ptr, lattice9, pbc3, d_packed_ptr, want_grad=True, d_local_e_ptr=None,
                    d_local_virial_ptr=None, stream_ptr=None, args_str=""):
        """GPU-resident evaluation: all ``*_ptr`` are raw device addresses (e.g. ``tensor.data_ptr()``)."""
        lat = np.ascontiguousarray(lattice9, dtype=np.float64).reshape(9)
        pbc = np.ascontiguousarray(np.asarray(pbc3, dtype=bool).astype(np.int32))
        _check(load_library().gap_potential_calc_device(self._h, int(N), C.c_void_p(d_pos_ptr), C.c_void_p(d_Z_ptr), _dp(lat), _ip(pbc),
                                                        args_str.encode(), int(bool(want_grad)), C.c_void_p(d_packed_ptr),
                                                        C.c_void_p(d_local_e_ptr) if d_local_e_ptr else None,
                                                        C.c_void_p(d_local_virial_ptr) if d_local_virial_ptr else None,
                                                        C.c_void_p(stream_ptr) if stream_ptr else None))

    def calc_device_enqueue(self, N, d_pos_ptr, d_Z_ptr, lattice9, pbc3, d_packed_ptr, want_grad=True, stream_ptr=None, args_str=""):
        """Enqueue-only evaluation (see ``gap_potential_calc_device_enqueue``); pair with :meth:`calc_device_verify` after a stream sync."""
        lat = np.ascontiguousarray(lattice9, dtype=np.float64).reshape(9)
        pbc = np.ascontiguousarray(np.asarray(pbc3, dtype=bool).astype(np.int32))
        _check(load_library().gap_potential_calc_device_enqueue(self._h, int(N), C.c_void_p(d_pos_ptr), C.c_void_p(d_Z_ptr), _dp(lat), _ip(pbc),
                                                                args_str.encode(), int(bool(want_grad)), C.c_void_p(d_packed_ptr), None, None,
                                                                C.c_void_p(stream_ptr) if stream_ptr else None))

    def calc_device_verify(self):
        """True when the last enqueued evaluation is valid; False = enqueue it again (the neighbour list overflowed its speculative size)."""
        rep = C.c_int(0)
        _check(load_library().gap_potential_calc_device_verify(self._h, C.byref(rep)))
        return rep.value == 0

    @property
    def n_coordinate(self):
        return load_library().gap_potential_n_coordinate(self._h)

    @property
    def launch_count(self):
        return load_library().gap_potential_launch_count(self._h)

    def set_timing(self, on=True):
        """Record per-stage CUDA events during calc (off by default; see last_timings).  on=2: only the events around the two
        covariance GEMMs."""
        _check(load_library().gap_potential_set_timing(self._h, 2 if on == 2 else (1 if on else 0)))

    def last_timings(self):
        t = np.zeros(8)
        load_library().gap_potential_last_timings(self._h, _dp(t))
        return dict(zip(["connect", "soap_forward", "cov_gemm1", "cov_gemm2", "soap_adjoint", "distance_2b", "other", "total"], t))

    def finalise(self):
        if getattr(self, "_h", None) is not None and self._h:
            load_library().gap_potential_finalise(self._h)
            self._h = None

    def __del__(self):
        try:
            self.finalise()
        except Exception:
            pass


def partition_bounds(rank, n_ranks, N):
    """Centres [first, last) of ``rank``: the same contiguous blocks ``gap_potential_set_partition`` uses."""
    return (rank * N) // n_ranks, ((rank + 1) * N) // n_ranks


def pack_results(energy, virial, force):
    """[E | virial(9) column-major | F(3,N)] -- the one buffer the path all-reduces."""
    return np.concatenate([[float(energy)], np.asarray(virial, dtype=np.float64).reshape(9, order="F"),
                           np.asarray(force, dtype=np.float64).reshape(-1)])


def unpack_results(packed, N):
    packed = np.asarray(packed)
    return {"energy": float(packed[0]), "virial": packed[1:10].reshape(3, 3, order="F").copy(), "force": packed[10:10 + 3 * N].reshape(N, 3).copy()}


def reduce_packed(t, group=None):
    """Sum of per-rank partial packed buffers held in a torch tensor (in place): the host-side stand-in the CPU (gloo) test of the
    partition logic uses.  The product path reduces inside libgapb200.so (``gap_potential_set_comm``)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def comm_unique_id():
    """``gap_comm_get_unique_id``: 128 bytes that identify a new communicator; call on ONE rank and distribute."""
    buf = C.create_string_buffer(128)
    _check(load_library().gap_comm_get_unique_id(buf))
    return buf.raw


def broadcast_comm_id(group=None, src=0):
    """The host's share of the communicator set-up: rank ``src`` creates the id, ``torch.distributed`` (NCCL or gloo; a Fortran host
    would use MPI_Bcast) hands it to everybody."""
    import torch.distributed as dist

    box = [comm_unique_id() if dist.get_rank(group) == src else None]
    dist.broadcast_object_list(box, src=src, group=group)
    return box[0]


def pinned_copy(a):
    """A page-locked copy of a numpy array (what a host that wants zero staging copies hands to ``calc``)."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(a)).clone().pin_memory()
    return t.numpy()


class ShardedPotential:
    """One process per GPU: the reference's MPI-parallel ``calc`` (atom mask + ``sum_in_place``,
    src/GAP/descriptors.f95:1036-1051, src/Potentials/IPModel_GAP.f95:538-556).

    Every rank holds the whole configuration (positions replicated) and evaluates the centres of its contiguous block; the
    ONLY exchange is the sum of the packed ``[E | virial(9) | F(3,N)]`` partials, and it happens INSIDE libgapb200.so on the
    evaluation's stream (``gap_potential_set_comm``: ncclAllReduce, or the one-shot NVLink peer-memory kernel for latency-bound
    sizes).  torch.distributed is used once, to hand the 128-byte communicator id to every rank; torch is plumbing here
    (device memory, streams, the process group), all arithmetic and the collective are in the library.
    """

    def __init__(self, args_str="", param_filename=None, param_str=None, device=0, group=None, rank=None, world_size=None, comm_id=None):
        import torch

        self.torch = torch
        self.group = group
        import torch.distributed as dist

        live = dist.is_available() and dist.is_initialized()
        if rank is None or world_size is None:
            rank, world_size = (dist.get_rank(group), dist.get_world_size(group)) if live else (0, 1)
        self.rank, self.world_size = int(rank), int(world_size)
        self.pot = Potential(args_str, param_filename=param_filename, param_str=param_str, device=device)
        self.device = torch.device("cuda", device)
        if self.world_size > 1 and (comm_id is not None or (live and dist.get_world_size(group) == self.world_size)):
            with torch.cuda.device(self.device):
                self.pot.set_comm(comm_id if comm_id is not None else broadcast_comm_id(group), self.rank, self.world_size)
            self.reduces = True
        else:
            # no process group: this handle evaluates ONE block of a partitioned run and returns partial sums (profiling aid, tests)
            self.pot.set_partition(self.rank, self.world_size)
            self.reduces = False
        self.stream = torch.cuda.Stream(self.device)  # a real stream: the C ABI reads stream 0 / NULL as "the handle's own"
        self._h_e = torch.zeros(1, dtype=torch.float64, pin_memory=True)

    def run(self, atoms, velocities, dt=1.0, n_steps=10, masses=None, args_str=""):
        """Sharded ``DynamicalSystem_run`` (Potential.f95:2304; BASELINE config C): NVE velocity Verlet with the neighbour list
        rebuilt on the device every step.  Every rank keeps the whole state resident on its GPU and evaluates its block of
        centres; the library sums ``[E | virial | F]`` over the ranks after each evaluation, then all ranks integrate all atoms
        with the same forces.  Updates ``atoms.positions``; returns (velocities, epot, ekin)."""
        torch = self.torch
        pos, Z, lat, pbc = _geometry(atoms)
        N = len(Z)
        m = np.ascontiguousarray(element_masses(Z) if masses is None else masses, dtype=np.float64)
        dev = self.device
        d_pos = torch.tensor(pos, dtype=torch.float64, device=dev)
        d_vel = torch.tensor(np.ascontiguousarray(velocities, dtype=np.float64), device=dev)
        d_Z = torch.tensor(Z, dtype=torch.int32, device=dev)
        d_m = torch.tensor(m, dtype=torch.float64, device=dev)
        d_packed = torch.zeros(10 + 3 * N, dtype=torch.float64, device=dev)
        ep, ek = np.zeros(n_steps + 1), np.zeros(n_steps + 1)
        torch.cuda.current_stream(dev).synchronize()
        _check(load_library().gap_md_run_device(self.pot._h, N, d_pos.data_ptr(), d_vel.data_ptr(), d_Z.data_ptr(), d_m.data_ptr(), _dp(lat),
                                                _ip(pbc), float(dt), int(n_steps), (self.pot.calc_args + " " + args_str).strip().encode(),
                                                d_packed.data_ptr(), None, None, _dp(ep), _dp(ek), self.stream.cuda_stream))
        atoms.positions[...] = d_pos.cpu().numpy()
        return d_vel.cpu().numpy(), ep, ek

    def calc_resident(self, N, d_pos, d_Z, lattice9, pbc3, d_packed, want_grad=True):
        """Inputs and outputs are device tensors; work is enqueued on torch's current stream (on ``self.stream``, ordered
        after and before the current stream, when the current stream is the legacy default stream)."""
        torch = self.torch
        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream == 0:
            self.stream.wait_stream(cur)
            with torch.cuda.stream(self.stream):
                self.calc_resident(N, d_pos, d_Z, lattice9, pbc3, d_packed, want_grad)
            cur.wait_stream(self.stream)
            return
        for _ in range(2):
            self.calc_resident_enqueue(N, d_pos, d_Z, lattice9, pbc3, d_packed, want_grad)
            if self.calc_resident_finish():
                break

    def calc_resident_enqueue(self, N, d_pos, d_Z, lattice9, pbc3, d_packed, want_grad=True, read_energy=True):
        """First half of :meth:`calc_resident`: enqueue the evaluation, the reduction over ranks (inside the library) and the
        read-back of the energy word on torch's current (non-default) stream and return without waiting."""
        cur = self.torch.cuda.current_stream(self.device)
        if cur.cuda_stream == 0:
            raise RuntimeError("calc_resident_enqueue needs a real current stream (torch.cuda.set_stream(sp.stream))")
        self.pot.calc_device_enqueue(N, d_pos.data_ptr(), d_Z.data_ptr(), lattice9, pbc3, d_packed.data_ptr(), want_grad=want_grad,
                                     stream_ptr=cur.cuda_stream)
        # a rank whose list overflowed has poisoned its energy with NaN (k_finalize): after the reduction every rank sees it, so
        # all ranks repeat together without an extra collective
        if read_energy:
            self._h_e.copy_(d_packed[:1], non_blocking=True)

    def calc_resident_finish(self, h_energy=None):
        """Second half: synchronise once, then check the speculatively sized neighbour list.  False = enqueue again (the
        repeat sizes the list exactly).  h_energy: pinned host tensor whose first word received the reduced energy (default:
        the word calc_resident_enqueue read back)."""
        self.torch.cuda.current_stream(self.device).synchronize()
        ok = self.pot.calc_device_verify()
        e = float((self._h_e if h_energy is None else h_energy)[0])
        return bool(ok and e == e)

    def calc(self, atoms, force=True, virial=True, out_force=None):
        """Host in, host out, on every rank: the plain host-pointer entry point of the C ABI (``gap_potential_calc``) -- exactly what a
        Fortran host calls.  H2D of (pos, Z), evaluation of this rank's block, reduction over the ranks, D2H of what THIS rank
        asked for (a rank that passes ``force=False`` skips the force read-back)."""
        return self.pot.calc(atoms, force=force, virial=virial, out_force=out_force)
