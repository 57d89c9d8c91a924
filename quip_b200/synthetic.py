"""Synthetic configurations and random-init GAP models of the shapes named in BASELINE.json / SURVEY.md 8(d).

Structures are seeded and deterministic.  Sparse points are SOAP vectors of atoms drawn from an independent
rattled/strained copy of the same structure type (unit norm, non-negative dot products -- i.i.d. Gaussian columns
would make c^zeta ~ 0 and hide errors); alphas ~ N(0,1) * (0.05/sqrt(M)) / delta^2 so |E_i| ~ O(0.1 eV).
The descriptor vectors are produced by the caller-supplied ``descriptor_fn`` (the CUDA library in the product and the
bench; the oracle in CPU-only tests), so this module has no compute of its own.
"""
from __future__ import annotations

import os

import numpy as np

from .atoms import Atoms
from .gap_xml import write_gap_xml


def _rattle(pos, sigma, rng):
    return pos + rng.normal(0.0, sigma, size=pos.shape)


def si_diamond(nx=8, ny=8, nz=8, a=5.431, rattle=0.05, seed=1, strain=0.0):
    """Config A: Si diamond, (nx,ny,nz) cubic cells (8 atoms each), positions + N(0, rattle)."""
    rng = np.random.default_rng(seed)
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0], [.25, .25, .25], [.25, .75, .75], [.75, .25, .75], [.75, .75, .25]])
    cells = np.array([[i, j, k] for i in range(nx) for j in range(ny) for k in range(nz)], dtype=np.float64)
    pos = (cells[:, None, :] + basis[None, :, :]).reshape(-1, 3) * a
    cell = np.diag([nx * a, ny * a, nz * a]) * (1.0 + strain)
    pos = _rattle(pos * (1.0 + strain), rattle, rng)
    return Atoms(np.full(len(pos), 14, dtype=np.int32), pos, cell, True)


def sic_zincblende(n=16, a=4.36, rattle=0.05, seed=2, strain=0.0):
    """Config B: 3C-SiC, n^3 cubic cells (4 Si + 4 C each)."""
    rng = np.random.default_rng(seed)
    fcc = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]])
    basis = np.concatenate([fcc, fcc + 0.25])
    Zb = np.array([14] * 4 + [6] * 4, dtype=np.int32)
    cells = np.array([[i, j, k] for i in range(n) for j in range(n) for k in range(n)], dtype=np.float64)
    pos = (cells[:, None, :] + basis[None, :, :]).reshape(-1, 3) * a * (1.0 + strain)
    Z = np.tile(Zb, len(cells))
    return Atoms(Z, _rattle(pos, rattle, rng), np.eye(3) * n * a * (1.0 + strain), True)


def amorphous_carbon(N=262144, density=3.0, min_dist=1.2, seed=3):
    """Config C: N carbon atoms uniformly random in a cubic box at `density` g/cm^3 with a minimum-distance rejection
    (batched random sequential insertion: candidates closer than min_dist to an accepted atom, or to an earlier
    candidate of the same batch, are rejected; periodic k-d trees do the searches)."""
    from scipy.spatial import cKDTree

    rng = np.random.default_rng(seed)
    mass = 12.011 * 1.66053906660  # g/mol -> 1e-24 g per atom ; A^3 = 1e-24 cm^3
    L = (N * mass / density) ** (1.0 / 3.0)
    pos = np.zeros((0, 3))
    while len(pos) < N:
        cand = rng.random((max(1024, int(1.3 * (N - len(pos)))), 3)) * L
        if len(pos):
            dist, _ = cKDTree(pos, boxsize=L).query(cand, k=1, distance_upper_bound=min_dist)
            cand = cand[np.isinf(dist)]
        if len(cand) > 1:
            pairs = cKDTree(cand, boxsize=L).query_pairs(min_dist, output_type="ndarray")
            if len(pairs):
                drop = np.zeros(len(cand), dtype=bool)
                # greedy in candidate order: the later member of a close pair goes unless the earlier one already went
                for i, j in pairs[np.lexsort((pairs[:, 0], pairs[:, 1]))]:
                    if not drop[i]:
                        drop[j] = True
                cand = cand[~drop]
        pos = np.concatenate([pos, cand])[:N]
    return Atoms(np.full(N, 6, dtype=np.int32), pos, np.eye(3) * L, True)


def si_slab(nx=64, ny=64, nz=32, a=5.431, vacuum=20.0, rattle=0.05, seed=4):
    """Config D: Si(001) slab, nx*ny*nz cells * 8 atoms, `vacuum` A of vacuum along z."""
    at = si_diamond(nx, ny, nz, a, rattle, seed)
    cell = at.cell.copy()
    cell[2, 2] += vacuum
    return Atoms(at.numbers, at.positions, cell, True)


SOAP_A = ("soap cutoff=5.0 cutoff_transition_width=0.5 n_max=8 l_max=8 atom_sigma=0.5 central_weight=1.0 n_species=1 Z=14 "
          "species_Z={14}")


SOAP_B = ("soap cutoff=5.0 cutoff_transition_width=0.5 n_max=10 l_max=6 atom_sigma=0.5 central_weight=1.0 n_species=2 "
          "species_Z={6 14} n_Z=1 Z=%d")
SOAP_C = ("soap cutoff=5.5 cutoff_transition_width=0.5 n_max=8 l_max=8 atom_sigma=0.5 central_weight=1.0 n_species=1 Z=6 "
          "species_Z={6}")
SOAP_D = ("soap cutoff=5.0 cutoff_transition_width=0.5 n_max=12 l_max=8 atom_sigma=0.5 central_weight=1.0 n_species=1 Z=14 "
          "species_Z={14}")
D2B_B = "distance_2b cutoff=5.0 cutoff_transition_width=0.5 covariance_type=ard_se delta=0.5 theta_uniform=1.0 Z1=%d Z2=%d"


def random_soap_coordinate(descriptor, X_source, M, delta=1.0, zeta=4.0, seed=101):
    """Pick M rows of X_source as sparse points and draw alphas."""
    rng = np.random.default_rng(seed)
    if len(X_source) < M:
        raise ValueError("need at least M=%d descriptor vectors, got %d" % (M, len(X_source)))
    rows = rng.choice(len(X_source), size=M, replace=False)
    alpha = rng.normal(0.0, 1.0, size=M) * (0.05 / np.sqrt(M)) / (delta * delta)
    return {"descriptor": descriptor, "covariance_type": 2, "delta": delta, "f0": 0.0, "zeta": zeta,
            "sparseX": np.ascontiguousarray(X_source[rows]), "alpha": alpha, "sparseCutoff": np.ones(M)}


def random_2b_coordinate(descriptor, r_min, r_max, M=20, delta=0.5, theta=1.0, seed=201):
    rng = np.random.default_rng(seed)
    return {"descriptor": descriptor, "covariance_type": 1, "delta": delta, "f0": 0.0, "theta": [theta],
            "sparseX": np.linspace(r_min, r_max, M).reshape(M, 1), "alpha": rng.normal(0.0, 0.05, size=M), "sparseCutoff": np.ones(M)}


def bootstrap_xml(path, descriptors, label="GAP_b200_bootstrap"):
    """A model with one zero sparse point per SOAP descriptor: enough to construct a handle and call descriptor_calc."""
    coords = []
    for desc, d in descriptors:
        coords.append({"descriptor": desc, "covariance_type": 2, "delta": 1.0, "zeta": 1.0, "sparseX": np.zeros((1, d)),
                       "alpha": np.zeros(1), "sparseCutoff": np.ones(1)})
    return write_gap_xml(path, coords, label=label, separate_files=False)


def soap_dimension(n_max, l_max, n_species=1):
    K1 = n_max * n_species
    return (l_max + 1) * K1 * (K1 + 1) // 2 + 1


def build_config_A(workdir, descriptor_fn, n_cells=8, M=2000, seed=1):
    """Config A of BASELINE.json: Si diamond n_cells^3*8 atoms, SOAP n_max=8 l_max=8 cutoff 5, zeta=4, M sparse points.
    descriptor_fn(desc_str, atoms) -> (n, d) array.  Returns (atoms, xml_path)."""
    os.makedirs(workdir, exist_ok=True)
    atoms = si_diamond(n_cells, n_cells, n_cells, seed=seed)
    nsrc = max(n_cells, int(np.ceil((M / 8.0) ** (1.0 / 3.0))))
    src = si_diamond(nsrc, nsrc, nsrc, rattle=0.08, seed=100 + seed, strain=0.01)
    X = descriptor_fn(SOAP_A, src)
    coord = random_soap_coordinate(SOAP_A, X, M, delta=1.0, zeta=4.0, seed=100 + seed)
    xml = write_gap_xml(os.path.join(workdir, "gap_config_A.xml"), [coord], e0={14: -158.54496821}, label="GAP_b200_config_A")
    return atoms, xml


def build_config_B(workdir, descriptor_fn, n_cells=16, M=4000, seed=2):
    """Config B: 3C-SiC n_cells^3*8 atoms, three distance_2b (Si-Si, Si-C, C-C; 20 sparse points each) + one SOAP per
    centre species (n_max=10 l_max=6, 2 species), M sparse points each.  descriptor_fn(desc_str, atoms) -> (n, d)."""
    os.makedirs(workdir, exist_ok=True)
    atoms = sic_zincblende(n_cells, seed=seed)
    nsrc = max(3, int(np.ceil((M / 4.0) ** (1.0 / 3.0))))
    src = sic_zincblende(nsrc, rattle=0.08, seed=100 + seed, strain=0.01)
    coords = [random_2b_coordinate(D2B_B % zz, 1.5, 5.0, M=20, delta=0.5, theta=1.0, seed=200 + k)
              for k, zz in enumerate(((14, 14), (14, 6), (6, 6)))]
    for k, Zc in enumerate((6, 14)):
        X = descriptor_fn(SOAP_B % Zc, src)
        coords.append(random_soap_coordinate(SOAP_B % Zc, X, min(M, len(X)), delta=1.0, zeta=4.0, seed=110 + seed + k))
    xml = write_gap_xml(os.path.join(workdir, "gap_config_B.xml"), coords, e0={14: -158.54496821, 6: -148.314002}, label="GAP_b200_config_B")
    return atoms, xml


def build_config_C(workdir, descriptor_fn, N=262144, M=9000, seed=3, n_src=None):
    """Config C: amorphous carbon, N atoms at 3.0 g/cm^3, SOAP cutoff 5.5 (n_max=8 l_max=8 -- BASELINE names neither; stated
    in the results), M sparse points."""
    os.makedirs(workdir, exist_ok=True)
    atoms = amorphous_carbon(N, seed=seed)
    src = amorphous_carbon(n_src or max(2 * M, 1024), seed=100 + seed)
    X = descriptor_fn(SOAP_C, src)
    coord = random_soap_coordinate(SOAP_C, X, min(M, len(X)), delta=1.0, zeta=4.0, seed=100 + seed)
    xml = write_gap_xml(os.path.join(workdir, "gap_config_C.xml"), [coord], e0={6: -148.314002}, label="GAP_b200_config_C")
    return atoms, xml


def build_config_D(workdir, descriptor_fn, nx=64, ny=64, nz=32, M=8000, seed=4):
    """Config D: Si(001) slab nx*ny*nz cells * 8 atoms with 20 A of vacuum, SOAP n_max=12 l_max=8, M sparse points."""
    os.makedirs(workdir, exist_ok=True)
    atoms = si_slab(nx, ny, nz, seed=seed)
    ns = max(3, int(np.ceil((M / 8.0) ** (1.0 / 3.0))))
    src = si_slab(ns, ns, ns, rattle=0.08, seed=100 + seed)
    X = descriptor_fn(SOAP_D, src)
    coord = random_soap_coordinate(SOAP_D, X, min(M, len(X)), delta=1.0, zeta=4.0, seed=100 + seed)
    xml = write_gap_xml(os.path.join(workdir, "gap_config_D.xml"), [coord], e0={14: -158.54496821}, label="GAP_b200_config_D")
    return atoms, xml
