"""Minimal atomic-configuration container and extended-XYZ reader.

The container carries exactly the fields the GAP path reads from QUIP's
``type(Atoms)`` (src/libAtoms/Atoms_types.f95:224-334): ``N``, ``Z(N)``,
``pos(3,N)``, ``lattice(3,3)`` (columns are the cell vectors a, b, c) and
``is_periodic(3)``.  It duck-types the handful of ``ase.Atoms`` accessors
that ``quippy.potential.Potential.calculate`` relies on, so an ``ase.Atoms``
object can be passed wherever an :class:`Atoms` is accepted.

The extended-XYZ reader covers the subset of src/libAtoms/xyz.c needed for
the reference's own test fixtures (``Lattice=``, ``Properties=``, ``pbc=``,
quoted values, per-atom real/int/string/logical columns).
"""
from __future__ import annotations

import re

import numpy as np

ELEMENT_NAMES = (
    "X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr "
    "Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb "
    "Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac Th Pa U Np Pu Am Cm Bk Cf Es Fm Md No Lr Rf "
    "Db Sg Bh Hs Mt Ds Rg Cn Nh Fl Mc Lv Ts Og"
).split()
ATOMIC_NUMBER = {s: z for z, s in enumerate(ELEMENT_NAMES)}


class Atoms:
    """Positions are stored C-contiguous as (N, 3): the memory image of Fortran ``pos(3,N)``."""

    def __init__(self, numbers, positions, cell=None, pbc=True, info=None, arrays=None):
        self.numbers = np.ascontiguousarray(numbers, dtype=np.int32)
        self.positions = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
        if cell is None:
            cell = np.zeros((3, 3))
        cell = np.asarray(cell, dtype=np.float64)
        if cell.shape == (3,):
            cell = np.diag(cell)
        self.cell = np.ascontiguousarray(cell)  # rows are the cell vectors (ASE convention)
        if isinstance(pbc, (bool, np.bool_)):
            pbc = (pbc,) * 3
        self.pbc = np.asarray(pbc, dtype=bool)
        self.info = dict(info or {})
        self.arrays = dict(arrays or {})
        self.calc = None

    def __len__(self):
        return len(self.numbers)

    # -- ase.Atoms-compatible accessors --------------------------------
    def get_atomic_numbers(self):
        return self.numbers

    def get_positions(self):
        return self.positions

    def get_cell(self):
        return self.cell

    def get_pbc(self):
        return self.pbc

    def get_volume(self):
        return abs(float(np.linalg.det(self.cell)))

    def get_potential_energy(self):
        self.calc.calculate(self, ["energy"])
        return self.calc.results["energy"]

    def get_forces(self):
        self.calc.calculate(self, ["energy", "forces"])
        return self.calc.results["forces"]

    def get_stress(self):
        self.calc.calculate(self, ["energy", "stress"])
        return self.calc.results["stress"]

    # -- QUIP-side view --------------------------------------------------
    @property
    def lattice_fortran(self):
        """3x3 column-major lattice (columns a,b,c) flattened = ASE ``cell`` rows flattened."""
        return np.ascontiguousarray(self.cell.reshape(9))


_KV = re.compile(r'(\w[\w\-\.]*)\s*=\s*("([^"]*)"|\'([^\']*)\'|\{([^}]*)\}|(\S+))|(\w[\w\-\.]*)')


def _parse_comment(line):
    out = {}
    for m in _KV.finditer(line):
        if m.group(7) is not None:  # bare key => true (ParamReader.f95:447-449)
            out[m.group(7)] = True
            continue
        key = m.group(1)
        val = next(v for v in (m.group(3), m.group(4), m.group(5), m.group(6)) if v is not None)
        out[key] = val
    return out


def _convert_scalar(v):
    if not isinstance(v, str):
        return v
    if v.startswith("_JSON"):
        import json

        return np.array(json.loads(v[5:].strip()))
    toks = v.split()
    try:
        arr = [float(t) for t in toks]
        if len(arr) == 1:
            return int(toks[0]) if re.fullmatch(r"[+-]?\d+", toks[0]) else arr[0]
        return np.array(arr)
    except ValueError:
        pass
    if all(t in ("T", "F", "True", "False") for t in toks) and toks:
        b = [t in ("T", "True") for t in toks]
        return b[0] if len(b) == 1 else np.array(b)
    return v


def read_xyz(path, index=None):
    """Read all frames (``index=None``) or a single frame of an extended-XYZ file."""
    frames = []
    with open(path) as fh:
        lines = fh.read().splitlines()
    p = 0
    while p < len(lines):
        if not lines[p].strip():
            p += 1
            continue
        n = int(lines[p].split()[0])
        info_raw = _parse_comment(lines[p + 1])
        props = info_raw.pop("Properties", "species:S:1:pos:R:3")
        lattice = info_raw.pop("Lattice", None)
        pbc = info_raw.pop("pbc", None)
        fields = props.split(":")
        cols = [(fields[i], fields[i + 1], int(fields[i + 2])) for i in range(0, len(fields), 3)]
        rows = [lines[p + 2 + i].split() for i in range(n)]
        arrays, c0 = {}, 0
        for name, typ, nc in cols:
            raw = [r[c0:c0 + nc] for r in rows]
            if typ == "R":
                a = np.array(raw, dtype=np.float64)
            elif typ == "I":
                a = np.array(raw, dtype=np.int64)
            elif typ == "L":
                a = np.array([[t in ("T", "True") for t in r] for r in raw])
            else:
                a = np.array(raw, dtype=object)
            arrays[name] = a[:, 0] if nc == 1 else a
            c0 += nc
        if "Z" in arrays:
            numbers = arrays.pop("Z")
        else:
            numbers = np.array([ATOMIC_NUMBER[s] for s in arrays["species"]])
        arrays.pop("species", None)
        pos = arrays.pop("pos")
        if lattice is not None:
            cell = np.array([float(t) for t in lattice.split()]).reshape(3, 3)
            pbc_v = [True] * 3 if pbc is None else [t in ("T", "True") for t in pbc.split()]
        else:
            cell, pbc_v = np.zeros((3, 3)), [False] * 3
        info = {k: _convert_scalar(v) for k, v in info_raw.items()}
        frames.append(Atoms(numbers, pos, cell, pbc_v, info=info, arrays=arrays))
        p += 2 + n
    if index is None:
        return frames
    return frames[index]
