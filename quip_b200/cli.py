"""``quip``-style command line for the GAP path (mirror of src/Programs/quip.f95 for the options that reach IP GAP).

    python -m quip_b200.cli atoms_filename=frames.xyz param_filename=gp.xml [init_args="IP GAP"] [calc_args="..."] E F V [local] [timing]

Same ``key=value`` grammar and flags as the reference (quip.f95:135-235: ``E``/``energy``, ``F``/``forces``, ``V``/``virial``,
``local``, ``calc_args``, ``init_args``, ``atoms_filename``, ``param_filename``, ``timing``), same printed keys per frame
(quip.f95:698-756: ``Energy=``, ``Virial``, ``Pressure eV/A^3 ... GPa``, ``Cell Volume:``) and the frame echoed as extended
XYZ lines with the ``AT`` prefix (quip.f95:821).  Everything is computed by libgapb200.so; options outside the GAP
evaluation path (minimisation, phonons, EVB, ...) are rejected.
"""
from __future__ import annotations

import sys
import time

import numpy as np

from .atoms import ELEMENT_NAMES, _parse_comment, read_xyz
from .potential import Potential

EV_A3_IN_GPA = 1.6022e-19 * 1.0e30 / 1.0e9  # src/libAtoms/Units.f95:76
_FLAGS = {"E": "E", "energy": "E", "F": "F", "forces": "F", "V": "V", "virial": "V", "local": "local", "timing": "timing"}
_KEYS = ("atoms_filename", "param_filename", "init_args", "calc_args", "cutoff_skin", "verbosity", "real_format", "at_file", "param_file")


def parse_cli(argv):
    quoted = []
    for a in argv:  # the shell has already removed the quotes of init_args="IP GAP": put grouping braces back (ParamReader.f95:422)
        k, eq, v = a.partition("=")
        quoted.append("%s={%s}" % (k, v) if eq and " " in v and v[:1] not in "{\"'" else a)
    opts = _parse_comment(" ".join(quoted))
    cfg = {"E": False, "F": False, "V": False, "local": False, "timing": False, "atoms_filename": "stdin", "param_filename": "quip_params.xml",
           "init_args": "", "calc_args": ""}
    for k, v in opts.items():
        if k in _FLAGS:
            cfg[_FLAGS[k]] = (v is True) or str(v) in ("T", "True", "true")
        elif k in _KEYS:
            cfg[{"at_file": "atoms_filename", "param_file": "param_filename"}.get(k, k)] = v
        else:
            raise RuntimeError("quip: option '%s' is outside the GAP evaluation path of this build" % k)
    return cfg


def _fmt(v):
    return "%.8f" % v


def write_frame(at, out, extra_arrays, prefix="AT"):
    props = "species:S:1:pos:R:3" + "".join(":%s:R:%d" % (k, 1 if a.ndim == 1 else a.shape[1]) for k, a in extra_arrays.items())
    info = " ".join('%s=%s' % (k, ('"%s"' % v) if " " in str(v) else v) for k, v in at.info.items() if not isinstance(v, np.ndarray))
    out.write("%s %d\n" % (prefix, len(at)))
    out.write('%s Lattice="%s" Properties=%s %s\n' % (prefix, " ".join(_fmt(v) for v in at.cell.reshape(-1)), props, info))
    for i in range(len(at)):
        cols = [ELEMENT_NAMES[at.numbers[i]]] + [_fmt(v) for v in at.positions[i]]
        for a in extra_arrays.values():
            cols += [_fmt(v) for v in np.atleast_1d(a[i])]
        out.write("%s %s\n" % (prefix, " ".join(cols)))


def main(argv=None, out=sys.stdout):
    cfg = parse_cli(sys.argv[1:] if argv is None else argv)
    if not (cfg["E"] or cfg["F"] or cfg["V"] or cfg["local"]):
        raise RuntimeError("Nothing to be calculated")  # quip.f95:819
    pot = Potential(cfg["init_args"], param_filename=cfg["param_filename"], calc_args=cfg["calc_args"])
    if cfg["timing"]:
        pot.set_timing(True)
    pot.set_cutoff_skin(float(cfg.get("cutoff_skin", 0.5)))  # quip.f95:217, 343-345: at%cutoff_skin, default 0.5 A
    frames = read_xyz(cfg["atoms_filename"])
    for at in frames:
        t0 = time.perf_counter()
        r = pot.calc(at, energy=True, force=cfg["F"], virial=cfg["V"], local_energy=cfg["local"] and cfg["E"], local_virial=cfg["local"] and cfg["V"])
        dt = time.perf_counter() - t0
        extra = {}
        if cfg["E"]:
            out.write("Energy=%.12f\n" % r["energy"])
            at.info["Energy"] = "%.12f" % r["energy"]
        if cfg["V"]:
            V0 = r["virial"]
            P0 = V0 / at.get_volume()
            for i in range(3):
                out.write("Virial %s\n" % " ".join("%.10f" % v for v in V0[i]))
            for i in range(3):
                out.write("Pressure eV/A^3 %s   GPa %s\n" % (" ".join("%.10f" % v for v in P0[i]), " ".join("%.10f" % v for v in P0[i] * EV_A3_IN_GPA)))
        out.write("Cell Volume: %.10f A^3\n" % at.get_volume())
        if cfg["F"]:
            extra["force"] = r["force"]
        if "local_energy" in r:
            extra["local_energy"] = r["local_energy"]
        if "local_virial" in r:
            extra["local_virial"] = r["local_virial"]
        if cfg["timing"]:
            t = pot.last_timings()
            out.write("TIMER: calc done in %.6f wall clock secs (device ms: %s)\n" % (dt, " ".join("%s=%.4f" % kv for kv in t.items())))
        write_frame(at, out, extra)
    return 0


if __name__ == "__main__":
    try:
        sys.exit(main())
    except RuntimeError as e:
        sys.stderr.write("SYSTEM ABORT: %s\n" % e)
        sys.exit(1)
