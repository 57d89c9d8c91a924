"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the reference's golden fixtures.

Bars (BASELINE.json north_star): neighbour lists bit-exact (identical (i, j, shift) sets AND identical FP64 distances);
energies within 1e-8 eV/atom, forces within 1e-6 eV/A, virials within 1e-6 eV, all FP64."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as orc  # noqa: E402
from quip_b200 import Atoms, Potential, read_xyz  # noqa: E402
from quip_b200 import synthetic as syn  # noqa: E402
from quip_b200.gap_xml import write_gap_xml  # noqa: E402
from tests.models import SI_SOAP, si_two_descriptor_model  # noqa: E402

TOL_E_PER_ATOM = 1e-8
TOL_F = 1e-6
TOL_V = 1e-6


@pytest.fixture(scope="module")
def gap_xml_pot(golden):
    return Potential("IP GAP", param_filename=os.path.join(golden, "GAP.xml"))


@pytest.fixture(scope="module")
def si_model(tmp_path_factory):
    d = tmp_path_factory.mktemp("si_model")
    xml = si_two_descriptor_model(str(d))
    return Potential("IP GAP", param_filename=xml), orc.Model(xml), xml


@pytest.fixture(scope="module")
def si_frames(golden):
    return read_xyz(os.path.join(golden, "Si.np1.xyz"))


def canon(off, j, s, d):
    i = np.repeat(np.arange(len(off) - 1), np.diff(off))
    key = np.lexsort((s[:, 2], s[:, 1], s[:, 0], j, i))
    return np.stack([i[key], j[key], s[key, 0], s[key, 1], s[key, 2]], axis=1), d[key]


def assert_same_list(pot, atoms, cutoff):
    g = canon(*pot.calc_connect(atoms, cutoff))
    o = canon(*orc.Connect(atoms, cutoff).arrays())
    assert g[0].shape == o[0].shape, (g[0].shape, o[0].shape)
    assert np.array_equal(g[0], o[0])
    assert np.array_equal(g[1], o[1])  # bit-exact FP64 distances
    return len(g[1])


# ----------------------------------------------------------------------------------------------------
# neighbour list
# ----------------------------------------------------------------------------------------------------
def test_neighbour_list_bit_exact_fixtures(gap_xml_pot, golden, si_frames):
    n = assert_same_list(gap_xml_pot, read_xyz(os.path.join(golden, "gap_sample.xyz"), 0), 4.0)
    assert n > 0
    for a in si_frames:  # 1..96 atoms, triclinic cells smaller than the cutoff (many periodic images)
        for rc in (4.0, 6.0):
            assert_same_list(gap_xml_pot, a, rc)


def test_left_handed_and_skewed_cells(si_model, si_frames):
    # a left-handed lattice (two cell vectors exchanged: negative triple product) and a strongly sheared one: list bit-exact, E / F / V vs oracle
    pot, om, _ = si_model
    a = si_frames[8]
    swapped = Atoms(a.numbers, a.positions, a.cell[[1, 0, 2]], True)
    assert np.linalg.det(swapped.cell) * np.linalg.det(a.cell) < 0
    shear = np.eye(3)
    shear[0, 1], shear[1, 2] = 0.45, -0.35
    sheared = Atoms(a.numbers, a.positions @ shear.T, a.cell @ shear.T, True)
    for b in (swapped, sheared):
        assert_same_list(pot, b, 6.0)
        check_efv(pot, om, b)
    # exchanging two cell vectors describes the same crystal
    assert abs(pot.calc(swapped)["energy"] - pot.calc(a)["energy"]) < 1e-9


def test_neighbour_list_pbc_variants_and_offsets(gap_xml_pot, si_frames):
    # tests/test_neighbour_list.py idea: every pbc combination, and invariance under lattice-vector offsets
    a = si_frames[8]
    for pbc in ([1, 1, 1], [1, 1, 0], [1, 0, 0], [0, 0, 0], [0, 1, 1], [0, 1, 0]):
        b = Atoms(a.numbers, a.positions, a.cell, pbc)
        assert_same_list(gap_xml_pot, b, 5.0)
    rng = np.random.default_rng(5)
    shifts = rng.integers(-3, 4, size=(len(a), 3))
    moved = Atoms(a.numbers, a.positions + shifts @ a.cell, a.cell, True)
    assert_same_list(gap_xml_pot, moved, 5.0)
    off0 = gap_xml_pot.calc_connect(a, 5.0)[0]
    off1 = gap_xml_pot.calc_connect(moved, 5.0)[0]
    assert np.array_equal(off0, off1)


def test_neighbour_list_edge_cases(gap_xml_pot, golden):
    S = json.load(open(os.path.join(golden, "soap_reference_cases.json")))
    for name in ("mono_3", "quad_3"):  # non-periodic triclinic
        for d in S["datasets"][name]:
            a = Atoms(d["numbers"], np.array(d["scaled_positions"]) @ np.array(d["cell"]), d["cell"], False)
            assert_same_list(gap_xml_pot, a, 6.0)
    one = Atoms([14], [[0.1, 0.2, 0.3]], np.eye(3) * 3.0, True)  # single atom, only self images
    assert assert_same_list(gap_xml_pot, one, 5.0) > 0
    assert assert_same_list(gap_xml_pot, Atoms([14], [[0.1, 0.2, 0.3]], np.eye(3) * 30.0, True), 5.0) == 0
    off, j, s, d = gap_xml_pot.calc_connect(Atoms(np.zeros(0, dtype=np.int32), np.zeros((0, 3)), np.eye(3), True), 5.0)
    assert len(off) == 1 and len(j) == 0


def test_neighbour_list_config_A_size(gap_xml_pot):
    a = syn.si_diamond(8, 8, 8)  # 4,096 atoms
    n = assert_same_list(gap_xml_pot, a, 5.0)
    assert 27.5 < n / len(a) < 28.5
    b = syn.sic_zincblende(6)
    assert_same_list(gap_xml_pot, b, 5.0)
    slab = syn.si_slab(4, 4, 3, vacuum=20.0)
    assert_same_list(gap_xml_pot, slab, 5.0)
    assert_same_list(gap_xml_pot, Atoms(slab.numbers, slab.positions, slab.cell, [1, 1, 0]), 5.0)


def test_neighbour_list_counting_sort_path(gap_xml_pot):
    # above cub's single-tile sort limit (4,864 atoms) the cells are filled by a counting sort; the list must not notice
    a = syn.si_diamond(9, 9, 9)  # 5,832 atoms
    n = assert_same_list(gap_xml_pot, a, 5.0)
    assert 27.5 < n / len(a) < 28.5
    rng = np.random.default_rng(11)
    L = 40.0
    gas = Atoms(np.full(6000, 6), rng.uniform(0, L, size=(6000, 3)), np.eye(3) * L, [1, 0, 1])  # uneven cells, mixed pbc
    assert_same_list(gap_xml_pot, gas, 4.0)


def test_stage_timing_is_opt_in(si_model, si_frames):
    pot, _, xml = si_model
    p = Potential("IP GAP", param_filename=xml)
    p.calc(si_frames[8], force=True)
    assert all(v == 0.0 for v in p.last_timings().values())
    p.set_timing(True)
    p.calc(si_frames[8], force=True)
    t = p.last_timings()
    assert t["total"] > 0.0 and t["connect"] > 0.0
    p.set_timing(False)


# ----------------------------------------------------------------------------------------------------
# SOAP descriptor
# ----------------------------------------------------------------------------------------------------
def test_soap_vectors_vs_golden_and_oracle(si_model, si_frames, golden):
    pot, om, _ = si_model
    z = np.load(os.path.join(golden, "si_two_descriptors.npz"))
    X = np.concatenate([pot.descriptor_calc(a, 1)[0] for a in si_frames])
    assert X.shape == (439, 325)
    assert np.abs(X[z["index_soap"] - 1] - z["sparsex_soap"]).max() < 1e-12  # reference fixture (tol 1e-8 there)
    Xo = np.concatenate([orc.soap_descriptor(SI_SOAP, a)["data"] for a in si_frames])
    assert np.abs(X - Xo).max() < 1e-12


def multi_species_model(tmpdir, desc, datasets, M, seed, delta=1.3, zeta=4.0, extra=()):
    X = np.concatenate([orc.soap_descriptor(desc, a)["data"] for a in datasets])
    rng = np.random.default_rng(seed)
    rows = rng.choice(len(X), size=min(M, len(X)), replace=False)
    coord = {"descriptor": desc, "covariance_type": 2, "delta": delta, "zeta": zeta, "sparseX": X[rows],
             "alpha": rng.normal(size=len(rows)), "sparseCutoff": rng.uniform(0.5, 1.0, size=len(rows))}
    return write_gap_xml(os.path.join(tmpdir, "ms_%d.xml" % seed), [coord] + list(extra), e0={23: 0.1, 41: -0.2, 42: 0.3, 73: 0.4, 6: -1.0})


def quad_datasets(golden, pbc):
    S = json.load(open(os.path.join(golden, "soap_reference_cases.json")))
    return [Atoms(d["numbers"], np.array(d["scaled_positions"]) @ np.array(d["cell"]), d["cell"], pbc) for d in S["datasets"]["quad_3"]]


def test_descriptor_tutorial_outputs_of_the_reference_binary(golden, tmp_path):
    """The stored outputs of src/GAP/doc_src/quippy-descriptor-tutorial.ipynb (printed by the real QUIP binary) on the GPU: SOAP vectors of the
    2-atom diamond cell (one species; two species with an extra H atom) to the 9 printed digits, the 92 distance_2b instances (neighbour
    list inside the cutoff) and the sum of their covariance_cutoff values as the energy of a GAP whose every instance has energy 1."""
    from tests.test_oracle_golden import _count_model, _tutorial
    T, a, ah = _tutorial(golden)
    for key, at in (("soap_1", a), ("soap_2", ah)):
        desc = T[key]["descriptor"]
        X = np.array(T[key]["data"])
        coord = {"descriptor": desc, "covariance_type": 2, "delta": 1.0, "zeta": 2.0, "sparseX": X, "alpha": np.ones(len(X)), "sparseCutoff": np.ones(len(X))}
        pot = Potential("", param_filename=write_gap_xml(str(tmp_path / (key + ".xml")), [coord], e0={6: -1.0, 1: 0.5}))
        x, ci = pot.descriptor_calc(at, 0)
        assert np.array_equal(ci, np.arange(len(X)))
        assert np.abs(x - X).max() < 1e-9
    off, j, s, d = Potential("", param_filename=_count_model(tmp_path, T["distance_2b"]["descriptor"])).calc_connect(a, 4.0)
    ref = T["distance_2b"]
    assert len(d) == ref["count"] and np.abs(np.sort(d) - np.sort(ref["data"])).max() < 1e-8
    e = Potential("", param_filename=_count_model(tmp_path, ref["descriptor"])).calc(a)["energy"]
    assert abs(e - sum(ref["covariance_cutoff"])) < 2e-7


def test_soap_multispecies_descriptor(golden, tmp_path):
    desc = "soap n_Z=4 n_species=4 Z={23 41 42 73} species_Z={23 41 42 73} n_max=4 l_max=2 cutoff=5 atom_sigma=0.4 cutoff_transition_width=0.5 central_weight=1"
    for pbc in (False, True):
        ds = quad_datasets(golden, pbc)
        xml = multi_species_model(str(tmp_path), desc, ds, 8, seed=7 + pbc)
        pot = Potential("", param_filename=xml)
        for a in ds:
            x, ci = pot.descriptor_calc(a, 0)
            o = orc.soap_descriptor(desc, a)
            assert np.array_equal(ci, o["ci"])
            assert np.abs(x - o["data"]).max() < 1e-12


# ----------------------------------------------------------------------------------------------------
# covariance stage
# ----------------------------------------------------------------------------------------------------
def test_gp_predict_vs_oracle(si_model, si_frames):
    pot, om, _ = si_model
    X = np.concatenate([orc.soap_descriptor(SI_SOAP, a)["data"] for a in si_frames[:9]])
    e, g = pot.gp_predict(1, X)
    eo, go = om.predict(1, X)
    assert np.abs(e - eo).max() < 1e-10 * max(1.0, np.abs(eo).max())
    assert np.abs(g - go).max() < 1e-10 * max(1.0, np.abs(go).max())


# ----------------------------------------------------------------------------------------------------
# full energy / force / virial
# ----------------------------------------------------------------------------------------------------
def check_efv(pot, om, atoms, connect_cutoff=None):
    r = pot.calc(atoms, force=True, virial=True, local_energy=True, local_virial=True)
    o = om.calc(atoms, local_energy=True, local_virial=True, connect_cutoff=connect_cutoff)
    n = max(len(atoms), 1)
    assert abs(r["energy"] - o["energy"]) / n < TOL_E_PER_ATOM, (r["energy"], o["energy"])
    assert np.abs(r["local_energy"] - o["local_energy"]).max() < 1e-8
    assert np.abs(r["force"] - o["force"]).max() < TOL_F
    assert np.abs(r["virial"] - o["virial"]).max() < TOL_V
    assert np.abs(r["local_virial"] - o["local_virial"]).max() < TOL_V
    # energy-only call must agree with the gradient call (do_grad_descriptor = .false. path)
    assert abs(pot.calc(atoms)["energy"] - r["energy"]) < 1e-9 * max(1.0, abs(r["energy"]))
    return r, o


def test_gap_xml_known_answer(gap_xml_pot, golden):
    # tests/test_gappot.py:30-48
    at = read_xyz(os.path.join(golden, "gap_sample.xyz"), 0)
    assert gap_xml_pot.cutoff() == 4.0
    om = orc.Model(os.path.join(golden, "GAP.xml"))
    r, _ = check_efv(gap_xml_pot, om, at)
    assert abs(r["energy"] - at.info["energy"]) < 1e-8
    assert np.abs(r["force"] - at.arrays["force"]).max() < 1e-8
    # ASE-calculator semantics (potential.py:281-288)
    res = gap_xml_pot.calculate(at, ["energy", "forces", "stress"])
    assert res["forces"].shape == (81, 3) and res["stress"].shape == (6,)
    assert abs(res["stress"][0] + r["virial"][0, 0] / at.get_volume()) < 1e-12


def test_h2_cell_energies(gap_xml_pot, golden):
    # tests/test_potential_cell.py:30-54
    h = json.load(open(os.path.join(golden, "h2_cell_energies.json")))
    for c, e in zip(h["cell_sizes"], h["ref_energies"]):
        a = Atoms(h["numbers"], h["positions"], [c, c, c], True)
        assert abs(gap_xml_pot.calc(a)["energy"] - e) < 1e-8


def test_si_two_descriptor_model_all_frames(si_model, si_frames):
    # BASELINE config[0]: distance_2b + SOAP on Si.np1.xyz
    pot, om, _ = si_model
    assert pot.cutoff() == 6.0 and pot.n_coordinate == 2
    for a in si_frames:
        check_efv(pot, om, a)
    # only_descriptor selects one coordinate (IPModel_GAP.f95:399-404); the two parts add up (minus the double e0)
    a = si_frames[7]
    e1 = pot.calc(a, args_str="only_descriptor=1")["energy"]
    e2 = pot.calc(a, args_str="only_descriptor=2")["energy"]
    e0 = len(a) * (-158.54496821 + 2.0)
    assert abs((e1 - e0) + (e2 - e0) + e0 - pot.calc(a)["energy"]) < 1e-8


def test_multispecies_soap_plus_2b_efv(golden, tmp_path):
    desc = ("soap n_Z=2 n_species=4 Z={23 42} species_Z={23 41 42 73} n_max=5 l_max=3 cutoff=4.5 atom_sigma=0.45 "
            "cutoff_transition_width=0.7 central_weight=0.8")
    extra = [syn.random_2b_coordinate("distance_2b cutoff=5.0 covariance_type=ard_se delta=0.5 theta_uniform=1.0 Z1=23 Z2=41", 1.5, 5.0, M=12,
                                      seed=3)]
    for pbc in (True, False, [1, 0, 1]):
        ds = quad_datasets(golden, pbc)
        xml = multi_species_model(str(tmp_path), desc, ds, 10, seed=11, zeta=2.0, extra=extra)
        pot, om = Potential("", param_filename=xml), orc.Model(xml)
        for a in ds:
            check_efv(pot, om, a)


def test_species_outside_map_and_unnormalised(golden, tmp_path):
    # neighbours whose species is not in species_Z are ignored (descriptors.f95:8194-8195); normalise=F; zeta non-integer
    desc = "soap n_Z=1 n_species=2 Z=23 species_Z={23 73} n_max=4 l_max=4 cutoff=5.0 atom_sigma=0.5 normalise=F covariance_sigma0=0.1"
    ds = quad_datasets(golden, True)
    xml = multi_species_model(str(tmp_path), desc, ds, 6, seed=13, zeta=2.5, delta=0.2)
    pot, om = Potential("", param_filename=xml), orc.Model(xml)
    for a in ds:
        check_efv(pot, om, a)


def test_config_A_reduced_vs_oracle(tmp_path):
    # same shape as BASELINE config A (n_max=8 l_max=8 cutoff 5, zeta=4) at a size the oracle finishes in seconds
    atoms, xml = syn.build_config_A(str(tmp_path), lambda desc, at: orc.soap_descriptor(desc, at)["data"], n_cells=3, M=150, seed=1)
    pot, om = Potential("", param_filename=xml), orc.Model(xml)
    r, o = check_efv(pot, om, atoms)
    assert np.abs(r["force"]).max() > 1e-3  # a non-trivial configuration


def test_partition_partials_sum_to_total(si_model, si_frames):
    # the reference's MPI atom mask + Allreduce (descriptors.f95:1036-1051, IPModel_GAP.f95:538-556)
    pot, om, xml = si_model
    a = si_frames[8]
    full = pot.calc(a, force=True, virial=True, local_energy=True)
    parts = []
    for r in range(3):
        p = Potential("IP GAP", param_filename=xml)
        p.set_partition(r, 3)
        parts.append(p.calc(a, force=True, virial=True, local_energy=True))
    for k, tol in (("energy", 1e-9), ("force", 1e-10), ("virial", 1e-9), ("local_energy", 1e-10)):
        tot = sum(np.asarray(p[k]) for p in parts)
        assert np.abs(tot - np.asarray(full[k])).max() < tol * max(1.0, np.abs(np.asarray(full[k])).max())


def test_device_resident_entry_point(si_model, si_frames):
    import torch

    pot, om, _ = si_model
    a = si_frames[8]
    ref = pot.calc(a, force=True, virial=True)
    pos = torch.tensor(a.positions, dtype=torch.float64, device="cuda")
    Z = torch.tensor(a.numbers, dtype=torch.int32, device="cuda")
    packed = torch.empty(10 + 3 * len(a), dtype=torch.float64, device="cuda")
    pot.calc_device(len(a), pos.data_ptr(), Z.data_ptr(), a.lattice_fortran, a.pbc, packed.data_ptr(), want_grad=True,
                    stream_ptr=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = packed.cpu().numpy()
    assert abs(out[0] - ref["energy"]) < 1e-9 * abs(ref["energy"])
    assert np.abs(out[1:10].reshape(3, 3, order="F") - ref["virial"]).max() < 1e-9
    assert np.abs(out[10:].reshape(-1, 3) - ref["force"]).max() < 1e-9


# ----------------------------------------------------------------------------------------------------
# full-size config A: size-independent properties
# ----------------------------------------------------------------------------------------------------
def test_config_A_full_size_vs_oracle(tmp_path):
    """BASELINE configs[1] at FULL size (4,096 atoms, 2,000 sparse points): neighbour list bit-exact, E/F/virial/local_e/local_virial
    against the oracle's whole evaluation (0.3 s of CPU), then the size-independent properties."""
    boot = syn.bootstrap_xml(str(tmp_path / "boot.xml"), [(syn.SOAP_A, syn.soap_dimension(8, 8))])
    bp = Potential("", param_filename=boot)
    atoms, xml = syn.build_config_A(str(tmp_path), lambda desc, at: bp.descriptor_calc(at, 0)[0], n_cells=8, M=2000, seed=1)
    assert len(atoms) == 4096
    pot, om = Potential("", param_filename=xml), orc.Model(xml)
    assert_same_list(pot, atoms, pot.cutoff())
    r, o = check_efv(pot, om, atoms)
    assert np.abs(r["force"]).max() > 1e-3
    assert abs(r["local_energy"].sum() - r["energy"]) < 1e-7
    assert np.abs(r["force"].sum(axis=0)).max() < 1e-8          # translation invariance
    assert np.abs(r["virial"] - r["virial"].T).max() < 1e-7     # rotation invariance
    assert np.abs(r["local_virial"].sum(axis=0).reshape(3, 3, order="F") - r["virial"]).max() < 1e-8
    # central finite difference of the energy along a random displacement direction (Potential.f95:1374 idea)
    rng = np.random.default_rng(0)
    dirn = rng.normal(size=atoms.positions.shape)
    dirn /= np.linalg.norm(dirn)
    h = 1e-4
    ep = pot.calc(Atoms(atoms.numbers, atoms.positions + h * dirn, atoms.cell, True))["energy"]
    em = pot.calc(Atoms(atoms.numbers, atoms.positions - h * dirn, atoms.cell, True))["energy"]
    assert abs((ep - em) / (2 * h) + np.sum(r["force"] * dirn)) < 1e-6
    # strain derivative = virial (dE/d eps_ab = -virial_ab)
    eps = 1e-5
    for (aa, bb) in ((0, 0), (1, 2)):
        F = np.eye(3)
        F[aa, bb] += eps
        Fm = np.eye(3)
        Fm[aa, bb] -= eps
        ep = pot.calc(Atoms(atoms.numbers, atoms.positions @ F.T, atoms.cell @ F.T, True))["energy"]
        em = pot.calc(Atoms(atoms.numbers, atoms.positions @ Fm.T, atoms.cell @ Fm.T, True))["energy"]
        assert abs((ep - em) / (2 * eps) + r["virial"][aa, bb]) < 1e-5


def test_si_fit_reproduces_dft_energies(si_model, si_frames):
    # end-to-end pin of a SOAP dot-product GAP on the GPU: the fitted model reproduces the frames' dft_energy to fit accuracy
    # (tests/test_gapfit.py:82-100; see tests/test_oracle_golden.py::test_si_fit_reproduces_dft_energies)
    pot, om, _ = si_model
    n_checked = 0
    for a in si_frames:
        if "dft_energy" not in a.info:
            continue
        err = abs(pot.calc(a)["energy"] - a.info["dft_energy"]) / len(a)
        assert err < (0.01 if len(a) <= 2 else 2e-3), (a.info.get("config_type"), len(a), err)
        n_checked += 1
    assert n_checked == 16


def test_wrapper_simple_and_print(si_model, si_frames, golden):
    # quip_wrapper_simple_ (quip_unified_wrapper.f95:311-332): one-shot F77-style call, pbc T T T; IPModel_GAP_Print (IPModel_GAP.f95:952)
    import ctypes as C

    from quip_b200 import load_library

    pot, om, xml = si_model
    a = si_frames[8]
    ref = pot.calc(a, force=True, virial=True)
    lib = load_library()
    N = len(a)
    pos = np.ascontiguousarray(a.positions, dtype=np.float64)
    Z = np.ascontiguousarray(a.numbers, dtype=np.int32)
    lat = np.ascontiguousarray(a.cell.reshape(9))
    e, f, v = C.c_double(0.0), np.zeros((N, 3)), np.zeros(9)
    dp = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
    rc = lib.gap_b200_wrapper_simple(xml.encode(), C.byref(C.c_int(N)), dp(lat), Z.ctypes.data_as(C.POINTER(C.c_int)), dp(pos), C.byref(e), dp(f), dp(v))
    assert rc == 0, lib.gap_last_error()
    assert abs(e.value - ref["energy"]) < 1e-9 * abs(ref["energy"])
    assert np.abs(f - ref["force"]).max() < 1e-10
    assert np.abs(v.reshape(3, 3, order="F") - ref["virial"]).max() < 1e-9
    assert lib.gap_b200_wrapper_simple(b"/nonexistent.xml", C.byref(C.c_int(N)), dp(lat), Z.ctypes.data_as(C.POINTER(C.c_int)), dp(pos), C.byref(e), dp(f), dp(v)) != 0
    txt = pot.print_()
    assert "IPModel_GAP : label = GAP_Si_two_descriptors" in txt and "cutoff = 6" in txt
    assert txt.count("coordinate ") == 2 and "n_sparseX=100" in txt and "zeta=4" in txt
    small = C.create_string_buffer(16)
    assert lib.gap_potential_print(pot._h, small, len(small)) == 0 and len(small.value) == 15  # truncated, NUL terminated


def test_error_reporting(golden):
    with pytest.raises(RuntimeError, match="could not initialise GAP potential"):
        Potential("IP GAP label=nonexistent", param_filename=os.path.join(golden, "GAP.xml"))
    pot = Potential("IP GAP", param_filename=os.path.join(golden, "GAP.xml"))
    at = read_xyz(os.path.join(golden, "gap_sample.xyz"), 0)
    with pytest.raises(RuntimeError, match="not yet implemented"):  # IPModel_GAP.f95:348-350
        pot.calc(at, args_str="E_scale=2.0")
    with pytest.raises(RuntimeError, match="singular lattice|zero-length"):
        pot.calc(Atoms(at.numbers, at.positions, np.zeros((3, 3)), True))


def test_speculative_neighbour_list_overflow_repeats(si_model, si_frames):
    # the list of call k+1 is sized from call k (same N): a denser configuration of the same N must trigger the
    # transparent repeat and still give the exact result
    pot, om, xml = si_model
    a = si_frames[8]
    r0 = pot.calc(a, force=True, virial=True)
    r1 = pot.calc(a, force=True, virial=True)  # speculative path, same list
    assert r0["energy"] == r1["energy"] and np.abs(r0["force"] - r1["force"]).max() < 1e-12
    dense = Atoms(a.numbers, a.positions * 0.93, a.cell * 0.93, True)  # ~25% more neighbours... then much more:
    for scale in (0.93, 0.8):
        dense = Atoms(a.numbers, a.positions * scale, a.cell * scale, True)
        r = pot.calc(dense, force=True, virial=True)
        fresh = Potential("IP GAP", param_filename=xml).calc(dense, force=True, virial=True)
        assert abs(r["energy"] - fresh["energy"]) < 1e-9 * abs(fresh["energy"])
        assert np.abs(r["force"] - fresh["force"]).max() < 1e-9
        assert np.abs(r["virial"] - fresh["virial"]).max() < 1e-8
    o = om.calc(dense)
    assert abs(r["energy"] - o["energy"]) / len(a) < TOL_E_PER_ATOM
    assert np.abs(r["force"] - o["force"]).max() < TOL_F


# ----------------------------------------------------------------------------------------------------
# the other BASELINE configurations: reduced sizes against the oracle, full sizes through invariances
# ----------------------------------------------------------------------------------------------------
def _oracle_desc(desc, at):
    return orc.soap_descriptor(desc, at)["data"]


def test_config_B_reduced_vs_oracle(tmp_path):
    # SiC, 2 species: three distance_2b + one SOAP (n_max=10 l_max=6) per centre species
    atoms, xml = syn.build_config_B(str(tmp_path), _oracle_desc, n_cells=3, M=60)
    pot, om = Potential("", param_filename=xml), orc.Model(xml)
    assert pot.n_coordinate == 5
    check_efv(pot, om, atoms)


def test_config_C_reduced_vs_oracle(tmp_path):
    # amorphous carbon, cutoff 5.5 (about 105 neighbours: several shared-memory tiles per centre), small periodic box
    atoms, xml = syn.build_config_C(str(tmp_path), _oracle_desc, N=400, M=80, n_src=300)
    pot, om = Potential("", param_filename=xml), orc.Model(xml)
    r, o = check_efv(pot, om, atoms)
    off = pot.calc_connect(atoms)[0]
    assert np.diff(off).mean() > 90


def test_config_D_reduced_vs_oracle(tmp_path):
    # Si slab with vacuum, n_max=12 (two DMMA channel tiles), evaluated whole and as 8 centre blocks
    atoms, xml = syn.build_config_D(str(tmp_path), _oracle_desc, nx=3, ny=3, nz=2, M=60)
    pot, om = Potential("", param_filename=xml), orc.Model(xml)
    r, o = check_efv(pot, om, atoms)
    tot = {k: 0.0 for k in ("energy", "force", "virial")}
    for rank in range(8):
        p = Potential("", param_filename=xml)
        p.set_partition(rank, 8)
        pr = p.calc(atoms, force=True, virial=True)
        for k in tot:
            tot[k] = tot[k] + np.asarray(pr[k])
    assert abs(tot["energy"] - o["energy"]) / len(atoms) < TOL_E_PER_ATOM
    assert np.abs(tot["force"] - o["force"]).max() < TOL_F
    assert np.abs(tot["virial"] - o["virial"]).max() < TOL_V


def _invariance_checks(pot, atoms, r, fd_tol=2e-6):
    assert np.abs(r["force"].sum(axis=0)).max() < 1e-7          # translation invariance
    assert np.abs(r["virial"] - r["virial"].T).max() < 1e-6 * max(1.0, np.abs(r["virial"]).max())  # rotation invariance
    rng = np.random.default_rng(0)
    dirn = rng.normal(size=atoms.positions.shape)
    dirn /= np.linalg.norm(dirn)
    h = 1e-4
    ep = pot.calc(Atoms(atoms.numbers, atoms.positions + h * dirn, atoms.cell, atoms.pbc))["energy"]
    em = pot.calc(Atoms(atoms.numbers, atoms.positions - h * dirn, atoms.cell, atoms.pbc))["energy"]
    assert abs((ep - em) / (2 * h) + np.sum(r["force"] * dirn)) < fd_tol * max(1.0, np.linalg.norm(r["force"]))


def test_config_B_full_size_properties(tmp_path):
    boot = syn.bootstrap_xml(str(tmp_path / "boot.xml"), [(syn.SOAP_B % 6, syn.soap_dimension(10, 6, 2)), (syn.SOAP_B % 14, syn.soap_dimension(10, 6, 2))])
    bp = Potential("", param_filename=boot)
    ic = {syn.SOAP_B % 6: 0, syn.SOAP_B % 14: 1}
    atoms, xml = syn.build_config_B(str(tmp_path), lambda desc, at: bp.descriptor_calc(at, ic[desc])[0], n_cells=16, M=4000)
    assert len(atoms) == 32768
    pot = Potential("", param_filename=xml)
    r = pot.calc(atoms, force=True, virial=True, local_energy=True)
    assert abs(r["local_energy"].sum() - r["energy"]) < 1e-6
    _invariance_checks(pot, atoms, r)
    # oracle on a bounded sample of centres, SOAP coordinates only (distance_2b splits each pair energy between both ends, so
    # its partial local energies are not comparable centre by centre): a SOAP-only copy of the same model
    spec = orc.load_gap_xml(xml)
    spec["coordinates"] = spec["coordinates"][3:]
    om = orc.Model(model=spec)
    first, last = 5000, 5016
    o = om.calc(atoms, first=first, last=last, local_energy=True, force=False, virial=False)
    e_soap = sum(pot.calc(atoms, local_energy=True, args_str="only_descriptor=%d" % k)["local_energy"] for k in (4, 5))
    e0 = np.where(atoms.numbers == 14, -158.54496821, -148.314002)
    assert np.abs((e_soap - 2 * e0)[first:last] - (o["local_energy"] - e0)[first:last]).max() < 1e-8


def test_config_C_full_size_properties(tmp_path):
    boot = syn.bootstrap_xml(str(tmp_path / "boot.xml"), [(syn.SOAP_C, syn.soap_dimension(8, 8))])
    bp = Potential("", param_filename=boot)
    atoms, xml = syn.build_config_C(str(tmp_path), lambda desc, at: bp.descriptor_calc(at, 0)[0], N=262144, M=9000, n_src=12288)
    pot = Potential("", param_filename=xml)
    r = pot.calc(atoms, force=True, virial=True, local_energy=True)
    assert abs(r["local_energy"].sum() - r["energy"]) < 1e-5
    _invariance_checks(pot, atoms, r)
    om = orc.Model(xml)
    first, last = 1000, 1016
    o = om.calc(atoms, first=first, last=last, local_energy=True, force=False, virial=False)
    assert np.abs(r["local_energy"][first:last] - o["local_energy"][first:last]).max() < 1e-8


def test_config_D_full_size_one_of_eight_blocks(tmp_path):
    # the 1,048,576-atom slab as rank 0 of 8 sees it: all positions resident, 131,072 centres evaluated
    boot = syn.bootstrap_xml(str(tmp_path / "boot.xml"), [(syn.SOAP_D, syn.soap_dimension(12, 8))])
    bp = Potential("", param_filename=boot)
    atoms, xml = syn.build_config_D(str(tmp_path), lambda desc, at: bp.descriptor_calc(at, 0)[0], M=8000)
    assert len(atoms) == 1048576
    pot = Potential("", param_filename=xml)
    pot.set_partition(0, 8)
    r = pot.calc(atoms, force=True, virial=True, local_energy=True)
    assert np.abs(r["force"].sum(axis=0)).max() < 1e-7  # every centre's contributions sum to zero
    assert np.count_nonzero(r["local_energy"]) == 131072
    om = orc.Model(xml)
    first, last = 70000, 70008
    o = om.calc(atoms, first=first, last=last, local_energy=True, force=False, virial=False)
    assert np.abs(r["local_energy"][first:last] - o["local_energy"][first:last]).max() < 1e-8


# ----------------------------------------------------------------------------------------------------
# device-resident MD loop (DynamicalSystem_run, Potential.f95:2304-2369)
# ----------------------------------------------------------------------------------------------------
def test_md_run_matches_host_verlet_with_oracle_forces(si_model, si_frames):
    from quip_b200.potential import element_masses

    pot, om, _ = si_model
    a = si_frames[8]
    at = Atoms(a.numbers, a.positions.copy(), a.cell, True)
    rng = np.random.default_rng(3)
    v0 = rng.normal(scale=0.01, size=at.positions.shape)  # A/fs
    m = element_masses(at.numbers)
    dt, n_steps = 1.0, 4
    # host reference: the same velocity-Verlet recurrence (advance_verlet1/2) with forces from the oracle
    x, v = at.positions.copy(), v0.copy()
    o = om.calc(Atoms(at.numbers, x, at.cell, True))
    acc = o["force"] / m[:, None]
    ep = [o["energy"]]
    for _ in range(n_steps):
        v = v + 0.5 * acc * dt
        x = x + v * dt
        o = om.calc(Atoms(at.numbers, x, at.cell, True))
        acc = o["force"] / m[:, None]
        v = v + 0.5 * acc * dt
        ep.append(o["energy"])
    vel, epot, ekin = pot.run(at, v0, dt=dt, n_steps=n_steps)
    assert np.abs(at.positions - x).max() < 1e-9
    assert np.abs(vel - v).max() < 1e-9
    assert np.abs(epot - np.array(ep)).max() / len(at) < TOL_E_PER_ATOM
    assert abs(ekin[-1] - 0.5 * np.sum(m[:, None] * v * v)) < 1e-9


def test_md_energy_conservation_config_A_shape(tmp_path):
    from quip_b200.potential import element_masses

    atoms, xml = syn.build_config_A(str(tmp_path), _oracle_desc, n_cells=3, M=100, seed=1)
    pot = Potential("", param_filename=xml)
    rng = np.random.default_rng(7)
    m = element_masses(atoms.numbers)
    kT = 8.617385e-5 * 300.0
    v0 = rng.normal(size=atoms.positions.shape) * np.sqrt(kT / m)[:, None]  # Maxwell at 300 K
    v0 -= (m[:, None] * v0).sum(axis=0) / m.sum()
    vel, epot, ekin = pot.run(atoms, v0, dt=0.5, n_steps=40)
    etot = epot + ekin
    assert np.abs(etot - etot[0]).max() < 2e-4 * max(1.0, ekin[0])  # NVE: total energy conserved to O(dt^2)
    assert abs(epot[-1] - epot[0]) > 1e-6  # and something actually moved


def test_sharded_md_driver_single_rank_matches_md_run(si_model, si_frames):
    # gap_md_run_device (device-resident state + reduction hook; the hook is exercised at world size 2 by
    # tools/md_sharded_check.py under torchrun) against gap_md_run on the same trajectory
    from quip_b200 import ShardedPotential

    pot, om, xml = si_model
    a = si_frames[8]
    rng = np.random.default_rng(5)
    v0 = rng.normal(scale=0.01, size=a.positions.shape)
    at1 = Atoms(a.numbers, a.positions.copy(), a.cell, True)
    v1, ep1, ek1 = pot.run(at1, v0, dt=1.0, n_steps=5)
    sp = ShardedPotential("IP GAP", param_filename=xml, rank=0, world_size=1)
    at2 = Atoms(a.numbers, a.positions.copy(), a.cell, True)
    v2, ep2, ek2 = sp.run(at2, v0, dt=1.0, n_steps=5)
    assert np.abs(at1.positions - at2.positions).max() < 1e-10
    assert np.abs(v1 - v2).max() < 1e-10
    assert np.abs(ep1 - ep2).max() < 1e-9 and np.abs(ek1 - ek2).max() < 1e-9


def test_quip_cli_known_answer(golden):
    # `quip atoms_filename=gap_sample.xyz param_filename=GAP.xml E F V` (quip.f95:135-235, 698-821)
    import io

    from quip_b200 import cli

    buf = io.StringIO()
    cli.main(["atoms_filename=" + os.path.join(golden, "gap_sample.xyz"), "param_filename=" + os.path.join(golden, "GAP.xml"), "E", "F", "V"], out=buf)
    lines = buf.getvalue().splitlines()
    e = float([ln for ln in lines if ln.startswith("Energy=")][0].split("=")[1])
    assert abs(e - 30.855129585407724) < 1e-7  # tests/test_gappot.py:36
    assert sum(ln.startswith("Virial ") for ln in lines) == 3 and sum(ln.startswith("Pressure eV/A^3") for ln in lines) == 3
    at = [ln for ln in lines if ln.startswith("AT ")]
    assert at[0].split()[1] == "81" and "force:R:3" in at[1]
    f = np.array([[float(t) for t in ln.split()[5:8]] for ln in at[2:]])
    ref = read_xyz(os.path.join(golden, "gap_sample.xyz"), 0).arrays["force"]
    assert np.abs(f - ref).max() < 1e-6
    with pytest.raises(RuntimeError, match="Nothing to be calculated"):
        cli.main(["atoms_filename=x.xyz", "param_filename=y.xml"])


def test_lammps_pair_style_quip_abi(si_model, si_frames):
    # quip_lammps_wrapper (quip_lammps_wrapper.f95:30-156) driven the way LAMMPS' pair_quip.cpp drives it: local atoms + explicit
    # ghost images, a full neighbour list with skin, 1-based neighbour indices; ghost forces folded back (reverse communication)
    import ctypes as C
    import itertools

    from quip_b200 import load_library

    pot, om, xml = si_model
    a = si_frames[8]
    ref = pot.calc(a, force=True, virial=True, local_energy=True)
    lib = load_library()
    assert lib.quip_lammps_api_version() == 1
    n_h = C.c_int(0)
    cutoff = C.c_double(0.0)
    f, s = xml.encode(), b"IP GAP"
    handle = (C.c_int * 2)()
    lib.quip_lammps_potential_initialise(handle, C.byref(n_h), C.byref(cutoff), f, C.byref(C.c_int(len(f))), s, C.byref(C.c_int(len(s))))
    assert n_h.value == 2 and cutoff.value == 6.0
    lib.quip_lammps_potential_initialise(handle, C.byref(n_h), C.byref(cutoff), f, C.byref(C.c_int(len(f))), s, C.byref(C.c_int(len(s))))
    # ghosts: every periodic image within cutoff + skin of the cell
    rc = cutoff.value + 0.3
    N = len(a)
    frac = np.linalg.solve(a.cell.T, a.positions.T).T
    frac -= np.floor(frac)
    pos0 = frac @ a.cell
    gpos, gorig = [], []
    heights = 1.0 / np.linalg.norm(np.linalg.inv(a.cell), axis=0)
    R = [int(np.ceil(rc / h)) for h in heights]
    for sh in itertools.product(*[range(-r, r + 1) for r in R]):
        if sh == (0, 0, 0):
            continue
        p = pos0 + np.array(sh) @ a.cell
        f2 = np.linalg.solve(a.cell.T, p.T).T
        keep = np.all((f2 > -rc / heights) & (f2 < 1 + rc / heights), axis=1)
        gpos.append(p[keep])
        gorig.append(np.nonzero(keep)[0])
    gpos, gorig = np.concatenate(gpos), np.concatenate(gorig)
    x = np.ascontiguousarray(np.concatenate([pos0, gpos]))
    Zall = np.ascontiguousarray(np.concatenate([a.numbers, a.numbers[gorig]]).astype(np.int32))
    ntot = len(x)
    from scipy.spatial import cKDTree

    tree = cKDTree(x)
    lists = tree.query_ball_point(pos0, rc)
    ilist = np.arange(N, dtype=np.int32)
    numneigh = np.array([len([j for j in l if j != i]) for i, l in enumerate(lists)], dtype=np.int32)
    neigh = np.array([j + 1 for i, l in enumerate(lists) for j in sorted(l) if j != i], dtype=np.int32)
    e = C.c_double(0.0)
    le, vir, lv, frc = np.zeros(ntot), np.zeros(9), np.zeros(9 * ntot), np.zeros((ntot, 3))
    tag = np.arange(1, ntot + 1, dtype=np.int32)
    ip = lambda v: v.ctypes.data_as(C.POINTER(C.c_int))
    dp = lambda v: v.ctypes.data_as(C.POINTER(C.c_double))
    lat = np.ascontiguousarray(a.cell.reshape(9))
    lib.quip_lammps_wrapper(C.byref(C.c_int(N)), C.byref(C.c_int(ntot - N)), ip(Zall), ip(tag), C.byref(C.c_int(N)), C.byref(C.c_int(len(neigh))),
                            ip(ilist), ip(numneigh), ip(neigh), dp(lat), handle, C.byref(n_h), dp(x), C.byref(e), dp(le), dp(vir), dp(lv), dp(frc))
    assert abs(e.value - ref["energy"]) / N < TOL_E_PER_ATOM
    ftot = frc[:N].copy()
    np.add.at(ftot, gorig, frc[N:])
    assert np.abs(ftot - ref["force"]).max() < TOL_F
    assert np.abs(vir.reshape(3, 3, order="F") - ref["virial"]).max() < TOL_V
    letot = le[:N].copy()
    np.add.at(letot, gorig, le[N:])
    assert np.abs(letot - ref["local_energy"]).max() < 1e-8


# ----------------------------------------------------------------------------------------------------
# optional inputs / outputs of IPModel_GAP_Calc (IPModel_GAP.f95:324-337): atom mask, energy per coordinate, GAP variance
# ----------------------------------------------------------------------------------------------------
def test_atom_mask_matches_oracle(si_model, si_frames):
    pot, om, _ = si_model
    for a in (si_frames[4], si_frames[8]):
        N = len(a)
        mask = np.zeros(N, dtype=bool)
        mask[::3] = True
        am = Atoms(a.numbers, a.positions, a.cell, True, arrays={"sel": mask})
        r = pot.calc(am, force=True, virial=True, local_energy=True, local_virial=True, args_str="atom_mask_name=sel")
        o = om.calc(a, atom_mask=mask, local_energy=True, local_virial=True)
        assert abs(r["energy"] - o["energy"]) / N < TOL_E_PER_ATOM
        assert np.abs(r["local_energy"] - o["local_energy"]).max() < 1e-8   # the split between pair partners is the reference's
        assert np.abs(r["force"] - o["force"]).max() < TOL_F
        assert np.abs(r["virial"] - o["virial"]).max() < TOL_V
        assert np.abs(r["local_virial"] - o["local_virial"]).max() < TOL_V
        # complementary masks add up to the unmasked result
        am2 = Atoms(a.numbers, a.positions, a.cell, True, arrays={"sel": ~mask})
        r2 = pot.calc(am2, force=True, virial=True, args_str="atom_mask_name=sel")
        full = pot.calc(a, force=True, virial=True)
        assert abs(r["energy"] + r2["energy"] - full["energy"]) < 1e-8 * N
        assert np.abs(r["force"] + r2["force"] - full["force"]).max() < 1e-9
    with pytest.raises(RuntimeError, match="did not find"):
        pot.calc(si_frames[4], args_str="atom_mask_name=nope")


def test_energy_per_coordinate(si_model, si_frames):
    pot, om, _ = si_model
    a = si_frames[8]
    r = pot.calc(a, force=True, args_str="energy_per_coordinate=epc")
    o = om.calc(a, energy_per_coordinate=True)
    assert r["epc"].shape == (2,)
    assert np.abs(r["epc"] - o["energy_per_coordinate"]).max() < 1e-8 * len(a)
    e0 = len(a) * (-158.54496821 + 2.0)
    assert abs(r["epc"].sum() + e0 - r["energy"]) < 1e-8 * len(a)
    r1 = pot.calc(a, args_str="energy_per_coordinate=epc only_descriptor=2")
    assert r1["epc"][0] == 0.0 and abs(r1["epc"][1] - r["epc"][1]) < 1e-9 * len(a)


def test_local_gap_variance_and_gradient(si_model, si_frames):
    pot, om, _ = si_model
    for a, reg in ((si_frames[4], 0.001), (si_frames[8], 0.01)):
        args = "local_gap_variance=var gap_variance_regularisation=%g" % reg
        r = pot.calc(a, force=True, args_str=args)
        o = om.calc(a, local_gap_variance=True, gap_variance_regularisation=reg)
        scale = np.abs(o["local_gap_variance"]).max()
        # k_mm has a condition number ~ delta^2 / regularisation^2: the two Cholesky solves agree to ~1e-16 * cond
        tol = 1e-9 * scale * (0.001 / reg) ** 2 * 1e3
        assert np.abs(r["var"] - o["local_gap_variance"]).max() < tol, np.abs(r["var"] - o["local_gap_variance"]).max()
        gs = np.abs(o["gap_variance_gradient"]).max()
        assert np.abs(r["gap_variance_gradient"] - o["gap_variance_gradient"]).max() < 1e-6 * max(gs, 1.0)
        # energies and forces are untouched by the request
        plain = pot.calc(a, force=True)
        assert r["energy"] == plain["energy"] and np.abs(r["force"] - plain["force"]).max() < 1e-11  # (forces are scattered with atomics)
        # energy only: variance without gradient
        r0 = pot.calc(a, args_str=args)
        assert "gap_variance_gradient" not in r0 and np.abs(r0["var"] - r["var"]).max() < 1e-12 * scale


@pytest.mark.parametrize("extra", [" R_mix=T K=3", " radial_basis=GTO"])
def test_local_gap_variance_of_soap_variants(si_frames, tmp_path, extra):
    # the predictive variance and its gradient through the compression-mode (borrowed default kernels) and GTO (general kernels) pull-backs
    desc = "soap cutoff=4.0 cutoff_transition_width=1.0 n_max=6 l_max=4 atom_sigma=0.5 central_weight=1.0 n_species=1 Z=14 species_Z={14}" + extra
    X = np.concatenate([orc.soap_descriptor(desc, si_frames[k])["data"] for k in (3, 8)])
    rng = np.random.default_rng(77)
    rows = rng.choice(len(X), size=10, replace=False)
    coord = {"descriptor": desc, "covariance_type": 2, "delta": 1.3, "zeta": 2.0, "sparseX": X[rows], "alpha": rng.normal(size=10),
             "sparseCutoff": np.ones(10)}  # (unit sparse cutoffs, as gap_fit writes them: larger ones drive the estimate negative)
    xml = write_gap_xml(str(tmp_path / "var.xml"), [coord], e0={14: -1.0})
    pot, om = Potential("", param_filename=xml), orc.Model(xml)
    a = si_frames[8]
    r = pot.calc(a, force=True, args_str="local_gap_variance=var gap_variance_regularisation=0.01")
    o = om.calc(a, local_gap_variance=True, gap_variance_regularisation=0.01)
    scale = np.abs(o["local_gap_variance"]).max()
    assert np.abs(r["var"] - o["local_gap_variance"]).max() < 1e-7 * scale
    assert np.abs(r["gap_variance_gradient"] - o["gap_variance_gradient"]).max() < 1e-6 * max(np.abs(o["gap_variance_gradient"]).max(), 1.0)
    assert np.abs(r["force"] - o["force"]).max() < TOL_F
    # a negative variance is an error on both sides, as in gp_predict.f95:3876-3878 (k carries the sparse cutoffs, k_mm does not)
    bad = dict(coord, sparseCutoff=np.full(10, 3.0))
    xml_bad = write_gap_xml(str(tmp_path / "var_bad.xml"), [bad], e0={14: -1.0})
    with pytest.raises(RuntimeError, match="negative variance"):
        Potential("", param_filename=xml_bad).calc(a, args_str="local_gap_variance=var gap_variance_regularisation=0.01")
    with pytest.raises(RuntimeError, match="negative variance"):
        orc.Model(xml_bad).calc(a, local_gap_variance=True, gap_variance_regularisation=0.01)


# ----------------------------------------------------------------------------------------------------
# SOAP variants: compression modes, nu_R / nu_S, Z_map, diagonal_radial, GTO / POLY radial bases -- the general
# path of soap_general.cu against the reference's own golden vectors and, for the gradients, E/F/V against the oracle
# ----------------------------------------------------------------------------------------------------
def _variant_cases(golden):
    meta = json.load(open(os.path.join(golden, "soap_reference_all.json")))
    z = np.load(os.path.join(golden, "soap_reference_all.npz"))
    S = json.load(open(os.path.join(golden, "soap_reference_cases.json")))["datasets"]
    return meta, z, S


def test_soap_reference_data_all_variants_gpu(golden, tmp_path):
    """tests/test_SOAP.py:36-77 on the GPU: gap_descriptor_calc reproduces X of ALL 122 cases of SOAP_reference_data.json (mixing with
    QUIP's random weights, coupling=F, Z_map, nu_R / nu_S, diagonal_radial, GTO, POLY, the default path; average=F and average=T)."""
    meta, z, S = _variant_cases(golden)
    for i, m in enumerate(meta):
        qs = m["quippy_str"]
        d = z["X_%d" % i].shape[1]
        coord = {"descriptor": qs, "covariance_type": 2, "delta": 1.0, "zeta": 2.0, "sparseX": np.zeros((1, d)), "alpha": np.zeros(1)}
        pot = Potential("", param_filename=write_gap_xml(str(tmp_path / ("v%d.xml" % i)), [coord], separate_files=False))
        ds = [Atoms(dd["numbers"], np.array(dd["scaled_positions"]) @ np.array(dd["cell"]), dd["cell"], False) for dd in S[m["dataset_name"]]]
        X = np.concatenate([pot.descriptor_calc(a, 0)[0] for a in ds])[z["perm_%d" % i]]
        assert X.shape == z["X_%d" % i].shape, (i, qs)
        assert np.abs(X - z["X_%d" % i]).max() < 1e-9, (i, qs, np.abs(X - z["X_%d" % i]).max())
        pot.finalise()
    assert len(meta) == 122


VARIANT_EFV = [0, 2, 6, 12, 18, 22, 30, 38, 48, 53, 55, 60, 69, 73, 75, 101, 113, 115, 117, 118, 120, 121,
               1, 3, 15, 23, 29, 57, 59, 79, 82, 109]  # second row: average=T (one descriptor per configuration)


@pytest.mark.parametrize("case", VARIANT_EFV)
def test_soap_variants_efv_vs_oracle(golden, tmp_path, case):
    """Gradients of the general path: a GAP built on the variant descriptor (random sparse points drawn from the oracle's descriptors,
    non-trivial alphas and sparseCutoff), energy / force / virial / local quantities against the oracle, whose forward-mode grad_data
    for the same strings is pinned to the reference's golden vectors (tests/test_oracle_golden.py)."""
    meta, z, S = _variant_cases(golden)
    qs = meta[case]["quippy_str"]
    for pbc in (False, True):
        ds = [Atoms(dd["numbers"], np.array(dd["scaled_positions"]) @ np.array(dd["cell"]), dd["cell"], pbc) for dd in S[meta[case]["dataset_name"]]]
        xml = multi_species_model(str(tmp_path), qs, ds, 9, seed=100 + case + int(pbc), zeta=3.0 if case % 2 else 2.0)
        pot, om = Potential("", param_filename=xml), orc.Model(xml)
        for a in ds:
            check_efv(pot, om, a)
        pot.finalise()


@pytest.mark.parametrize("shape", [(12, 6), (12, 8), (10, 12), (6, 4)])
def test_default_power_spectrum_shapes_vs_oracle(si_frames, tmp_path, shape):
    """The default power spectrum at (n_max, l_max) = (12,6), (12,8), (10,12) (warp-per-centre / block specialisations) and (6,4) (run-time-shape
    kernels): descriptor and E / F / V on Si frames against the oracle."""
    n, l = shape
    desc = "soap cutoff=4.0 cutoff_transition_width=1.0 n_max=%d l_max=%d atom_sigma=0.5 central_weight=1.0 n_species=1 Z=14 species_Z={14}" % (n, l)
    frames = [si_frames[k] for k in (3, 8, 16)]
    xml = multi_species_model(str(tmp_path), desc, frames, 12, seed=400 + n + l, zeta=4.0)
    pot, om = Potential("", param_filename=xml), orc.Model(xml)
    for a in frames:
        x, ci = pot.descriptor_calc(a, 0)
        o = orc.soap_descriptor(desc, a)
        assert np.array_equal(ci, o["ci"]) and np.abs(x - o["data"]).max() < 1e-12
        check_efv(pot, om, a)


HYBRID_SHAPES = [
    # (descriptor options on top of the shape, frames) -- shapes with warp-per-centre specialisations: (8,8,1) and (10,6,2)
    ("soap cutoff=4.0 cutoff_transition_width=1.0 n_max=8 l_max=8 atom_sigma=0.5 central_weight=1.0 n_species=1 Z=14 species_Z={14} R_mix=T K=4", "si"),
    ("soap cutoff=4.0 cutoff_transition_width=1.0 n_max=8 l_max=8 atom_sigma=0.5 central_weight=1.0 n_species=1 Z=14 species_Z={14} nu_R=1", "si"),
    ("soap cutoff=4.0 cutoff_transition_width=1.0 n_max=8 l_max=8 atom_sigma=0.5 central_weight=1.0 n_species=1 Z=14 species_Z={14} diagonal_radial=T "
     "normalise=F", "si"),
    ("soap cutoff=4.5 n_max=10 l_max=6 atom_sigma=0.5 n_species=2 species_Z={23 41} n_Z=2 Z={23 41} Z_mix=T K=3 sym_mix=T", "quad"),
    ("soap cutoff=4.5 n_max=10 l_max=6 atom_sigma=0.5 n_species=2 species_Z={23 41} n_Z=2 Z={23 41} coupling=F", "quad"),
]


@pytest.mark.parametrize("spec", range(len(HYBRID_SHAPES)))
def test_compression_modes_on_specialised_shapes_vs_oracle(golden, si_frames, tmp_path, spec):
    """Compression modes on the EQUISPACED_GAUSS basis borrow the default path's kernels (density expansion with skip_power, neighbour phase
    with lambda_in): here on the shapes that have warp-per-centre specialisations, (8,8,1) and (10,6,2); E / F / V and the local
    quantities against the oracle, plus the descriptor itself."""
    desc, which = HYBRID_SHAPES[spec]
    if which == "si":
        frames = [si_frames[k] for k in (3, 8, 16)]
    else:  # four-species cells: the two mapped species are centres and neighbours, the others are ignored (descriptors.f95:8194-8195)
        frames = quad_datasets(golden, True)[:2] + quad_datasets(golden, False)[:1]
    xml = multi_species_model(str(tmp_path), desc, frames, 12, seed=300 + spec, zeta=2.0 + (spec % 2))
    pot, om = Potential("", param_filename=xml), orc.Model(xml)
    for a in frames:
        x, ci = pot.descriptor_calc(a, 0)
        o = orc.soap_descriptor(desc, a)
        assert np.array_equal(ci, o["ci"]) and np.abs(x - o["data"]).max() < 1e-12
        check_efv(pot, om, a)


# ----------------------------------------------------------------------------------------------------
# skin-based neighbour-list reuse (calc_connect with cutoff_skin, Connection.f95:1085-1128)
# ----------------------------------------------------------------------------------------------------
def test_cutoff_skin_reuses_list_with_identical_results(si_model, si_frames, tmp_path):
    pot, om, xml = si_model
    a = si_frames[8]
    rng = np.random.default_rng(21)
    skin_pot = Potential("IP GAP", param_filename=xml)
    skin_pot.set_cutoff_skin(0.5)
    pos = a.positions.copy()
    for step in range(6):
        # steps 1-3 move atoms by < skin / 2 in total (reuse), step 4 jumps one atom by 0.6 A (rebuild), step 5 is small again (reuse)
        if step in (1, 2, 3, 5):
            pos = pos + rng.uniform(-0.04, 0.04, size=pos.shape)
        if step == 4:
            pos = pos.copy()
            pos[3] += np.array([0.6, 0.0, 0.0])
        at = Atoms(a.numbers, pos, a.cell, True)
        r = skin_pot.calc(at, force=True, virial=True, local_energy=True, local_virial=True)
        o = om.calc(at, local_energy=True, local_virial=True)
        assert abs(r["energy"] - o["energy"]) / len(a) < TOL_E_PER_ATOM
        assert np.abs(r["force"] - o["force"]).max() < TOL_F
        assert np.abs(r["virial"] - o["virial"]).max() < TOL_V
        assert np.abs(r["local_energy"] - o["local_energy"]).max() < 1e-8
        assert np.abs(r["local_virial"] - o["local_virial"]).max() < TOL_V
    st = skin_pot.connect_stats()
    assert st == {"rebuilds": 2, "reuses": 4}, st
    # a changed lattice forces the rebuild (:1093-1096)
    at = Atoms(a.numbers, pos * 1.001, a.cell * 1.001, True)
    r = skin_pot.calc(at, force=True)
    assert abs(r["energy"] - om.calc(at)["energy"]) / len(a) < TOL_E_PER_ATOM
    assert skin_pot.connect_stats()["rebuilds"] == 3


def test_md_with_skin_matches_rebuild_every_step(tmp_path):
    # DynamicalSystem_run with cutoff_skin (the quip / md programs' default): same trajectory as rebuilding every step, fewer list builds
    from quip_b200.potential import element_masses

    atoms, xml = syn.build_config_A(str(tmp_path), _oracle_desc, n_cells=3, M=100, seed=1)
    rng = np.random.default_rng(7)
    m = element_masses(atoms.numbers)
    v0 = rng.normal(size=atoms.positions.shape) * np.sqrt(8.617385e-5 * 300.0 / m)[:, None]
    a1 = Atoms(atoms.numbers, atoms.positions.copy(), atoms.cell, True)
    a2 = Atoms(atoms.numbers, atoms.positions.copy(), atoms.cell, True)
    p1 = Potential("", param_filename=xml)
    p2 = Potential("", param_filename=xml)
    p2.set_cutoff_skin(0.5)
    v1, ep1, ek1 = p1.run(a1, v0, dt=1.0, n_steps=30)
    v2, ep2, ek2 = p2.run(a2, v0, dt=1.0, n_steps=30)
    assert np.abs(a1.positions - a2.positions).max() < 1e-9
    assert np.abs(ep1 - ep2).max() < 1e-8 * len(atoms)
    st = p2.connect_stats()
    assert st["reuses"] > 20 and st["rebuilds"] >= 1 and st["rebuilds"] + st["reuses"] == 31, st


# ----------------------------------------------------------------------------------------------------
# distance_2b options (descriptors.f95:1757-1815, 4735-4764): exponents, tail, only_intra / only_inter
# ----------------------------------------------------------------------------------------------------
def _d2b_variant_model(tmpdir, seed=5):
    rng = np.random.default_rng(seed)
    def coord(desc, d, M=9):
        return {"descriptor": desc, "covariance_type": 1, "delta": 0.7, "f0": 0.05, "theta": list(rng.uniform(0.3, 1.2, size=d)),
                "sparseX": rng.uniform(0.05, 1.0, size=(M, d)) if d > 1 else np.linspace(1.2, 4.5, M).reshape(M, 1),
                "alpha": rng.normal(0.0, 0.2, size=M), "sparseCutoff": rng.uniform(0.6, 1.0, size=M)}
    coords = [coord("distance_2b cutoff=4.5 cutoff_transition_width=0.8 Z1=23 Z2=41 n_exponents=3 exponents={-1 -2 -4}", 3),
              coord("distance_2b cutoff=5.0 Z1=42 Z2=0 tail_exponent=2 tail_range=0.7", 1),
              coord("distance_2b cutoff=4.0 Z1=0 Z2=0 n_exponents=2 tail_exponent=1 tail_range=1.3 only_inter resid_name=resid", 2),
              coord("distance_2b cutoff=4.0 Z1=0 Z2=0 only_intra resid_name=resid", 1)]
    return write_gap_xml(os.path.join(tmpdir, "d2b_variants.xml"), coords, e0={23: 0.1, 41: -0.2, 42: 0.3, 73: 0.4})


def test_distance_2b_options_vs_oracle(golden, tmp_path):
    xml = _d2b_variant_model(str(tmp_path))
    pot, om = Potential("", param_filename=xml), orc.Model(xml)
    for pbc in (True, False):
        for a in quad_datasets(golden, pbc):
            resid = np.arange(len(a)) // 2
            am = Atoms(a.numbers, a.positions, a.cell, pbc, arrays={"resid": resid})
            check_efv(pot, om, am)
    # the residue property is mandatory for only_intra / only_inter (:4660-4668)
    fresh = Potential("", param_filename=xml)
    with pytest.raises(RuntimeError, match="residue id"):
        fresh.calc(quad_datasets(golden, True)[0])


# ----------------------------------------------------------------------------------------------------
# angle_3b (descriptors.f95:1886-1911, 4932-5112): three-body descriptor with ARD_SE covariance
# ----------------------------------------------------------------------------------------------------
def _angle_3b_model(tmpdir, seed=11, M=12, with_si=False):
    rng = np.random.default_rng(seed)
    def coord(desc, M=M):
        X = np.column_stack([rng.uniform(3.0, 7.5, size=M), rng.uniform(0.0, 2.0, size=M), rng.uniform(1.5, 6.0, size=M)])
        return {"descriptor": desc, "covariance_type": 1, "delta": 0.6, "f0": 0.02, "theta": list(rng.uniform(0.8, 2.0, size=3)),
                "sparseX": X, "alpha": rng.normal(0.0, 0.3, size=M), "sparseCutoff": rng.uniform(0.7, 1.0, size=M)}
    if with_si:
        return write_gap_xml(os.path.join(tmpdir, "angle_3b_si.xml"), [coord("angle_3b cutoff=3.7 cutoff_transition_width=0.6 Z_center=14 Z1=14 Z2=14", M=40)],
                             e0={14: -0.3})
    coords = [coord("angle_3b cutoff=4.2 cutoff_transition_width=0.7 Z_center=0 Z1=0 Z2=0"),
              coord("angle_3b cutoff=4.6 Z=23 Z1=41 Z2=42"),
              coord("angle_3b cutoff=4.0 Z_center=41 Z1=23 Z2=23")]
    return write_gap_xml(os.path.join(tmpdir, "angle_3b.xml"), coords, e0={23: 0.1, 41: -0.2, 42: 0.3, 73: 0.4})


def test_angle_3b_vs_oracle(golden, tmp_path):
    xml = _angle_3b_model(str(tmp_path))
    pot, om = Potential("", param_filename=xml), orc.Model(xml)
    assert pot.cutoff() == 4.6 and pot.n_coordinate == 3
    for pbc in (True, False):
        for a in quad_datasets(golden, pbc):
            check_efv(pot, om, a)
    a = quad_datasets(golden, True)[1]
    # atom mask: only masked atoms are centres, their neighbours still take forces (IPModel_GAP.f95:344-346, 472-491)
    mask = np.arange(len(a)) % 2 == 0
    am = Atoms(a.numbers, a.positions, a.cell, True, arrays={"sel": mask})
    r = pot.calc(am, force=True, virial=True, local_energy=True, args_str="atom_mask_name=sel")
    o = om.calc(a, local_energy=True, atom_mask=mask)
    assert abs(r["energy"] - o["energy"]) < 1e-9 and np.abs(r["force"] - o["force"]).max() < TOL_F and np.abs(r["virial"] - o["virial"]).max() < TOL_V
    # energy_per_coordinate and only_descriptor
    r = pot.calc(a, args_str="energy_per_coordinate=epc")
    o = om.calc(a, energy_per_coordinate=True, force=False, virial=False)
    assert np.abs(r["epc"] - o["energy_per_coordinate"]).max() < 1e-9
    r1 = pot.calc(a, args_str="only_descriptor=2")
    assert abs(r1["energy"] - (o["energy_per_coordinate"][1] + sum({23: 0.1, 41: -0.2, 42: 0.3, 73: 0.4}[int(z)] for z in a.numbers))) < 1e-9
    # the variance of an angle_3b coordinate is refused, not approximated
    with pytest.raises(RuntimeError, match="angle_3b"):
        pot.calc(a, args_str="local_gap_variance=var")


def test_angle_3b_silicon_frames_partition_and_deterministic(si_frames, tmp_path):
    # a Si three-body coordinate on the reference's Si.np1.xyz frames (up to 96 atoms, 30-odd neighbours inside 3.7 A pairs), whole and as
    # the sum of three centre blocks; deterministic mode: two calls bit-identical and equal to the default within rounding
    xml = _angle_3b_model(str(tmp_path), with_si=True)
    pot, om = Potential("", param_filename=xml), orc.Model(xml)
    for k in (3, 8, 16):
        check_efv(pot, om, si_frames[k])
    a = si_frames[16]
    full = pot.calc(a, force=True, virial=True)
    N = len(a)
    acc = {"energy": 0.0, "force": np.zeros((N, 3)), "virial": np.zeros((3, 3))}
    for r in range(3):
        p = Potential("", param_filename=xml)
        p.set_partition(r, 3)
        part = p.calc(a, force=True, virial=True)
        for key in acc:
            acc[key] = acc[key] + part[key]
    assert abs(acc["energy"] - full["energy"]) < 1e-9 * max(1.0, abs(full["energy"]))
    assert np.abs(acc["force"] - full["force"]).max() < 1e-10 and np.abs(acc["virial"] - full["virial"]).max() < 1e-9
    det = Potential("", param_filename=xml)
    det.set_deterministic(True)
    d1 = det.calc(a, force=True, virial=True)
    d2 = det.calc(a, force=True, virial=True)
    assert np.array_equal(d1["force"], d2["force"]) and d1["energy"] == d2["energy"]
    assert np.abs(d1["force"] - full["force"]).max() < 1e-10 and np.abs(d1["virial"] - full["virial"]).max() < 1e-9


def test_c_example_program_matches_python_binding(gap_xml_pot, golden, tmp_path):
    # examples/wrapper_simple_example.c (a plain-C caller of the library, the role of the reference's quip_wrapper_simple_example_C.c), compiled with gcc and run
    import subprocess

    from tests.test_host_cpu import _build_c_example
    exe = _build_c_example(tmp_path)
    r = subprocess.run([exe, os.path.join(golden, "GAP.xml")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = dict(line.split(" = ") for line in r.stdout.strip().splitlines())
    a = Atoms([1, 1], [[1.50, 2.25, 3.00], [2.35, 2.75, 3.40]], np.eye(3) * 12.0, True)
    ref = gap_xml_pot.calc(a, force=True, local_energy=True)
    assert abs(float(out["Energy"]) - ref["energy"]) < 1e-10 and abs(float(out["Energy2"]) - ref["energy"]) < 1e-10
    assert np.abs(np.array(out["Force0"].split(), dtype=float) - ref["force"][0]).max() < 1e-10
    assert np.abs(np.array(out["LocalE"].split(), dtype=float) - ref["local_energy"]).max() < 1e-10
    assert abs(float(out["Cutoff"]) - 4.0) < 1e-12


def test_md_run_device_reduce_hook_is_called(si_model, si_frames):
    # gap_md_run_device's reduction hook (a host without the library's communicator, e.g. one that reduces with MPI-aware CUDA): called once
    # after every force evaluation, on the evaluation's stream; the trajectory equals gap_md_run's
    import ctypes as C

    import torch

    from quip_b200 import load_library
    from quip_b200.potential import element_masses

    pot, om, xml = si_model
    a = si_frames[8]
    N = len(a)
    rng = np.random.default_rng(5)
    v0 = rng.normal(scale=0.01, size=a.positions.shape)
    at1 = Atoms(a.numbers, a.positions.copy(), a.cell, True)
    v1, ep1, ek1 = pot.run(at1, v0, dt=1.0, n_steps=4)
    calls = []
    hook = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)(lambda ctx, stream: calls.append(stream))
    dev = torch.device("cuda", 0)
    d_pos = torch.tensor(a.positions, dtype=torch.float64, device=dev)
    d_vel = torch.tensor(v0, dtype=torch.float64, device=dev)
    d_Z = torch.tensor(a.numbers, dtype=torch.int32, device=dev)
    d_m = torch.tensor(element_masses(a.numbers), dtype=torch.float64, device=dev)
    d_packed = torch.zeros(10 + 3 * N, dtype=torch.float64, device=dev)
    ep, ek = np.zeros(5), np.zeros(5)
    lat = np.ascontiguousarray(a.cell.reshape(9))
    pbc = np.ones(3, dtype=np.int32)
    torch.cuda.synchronize()
    p2 = Potential("IP GAP", param_filename=xml)
    rc = load_library().gap_md_run_device(p2._h, N, d_pos.data_ptr(), d_vel.data_ptr(), d_Z.data_ptr(), d_m.data_ptr(), lat.ctypes.data_as(C.POINTER(C.c_double)),
                                          pbc.ctypes.data_as(C.POINTER(C.c_int)), 1.0, 4, b"", d_packed.data_ptr(), C.cast(hook, C.c_void_p), None,
                                          ep.ctypes.data_as(C.POINTER(C.c_double)), ek.ctypes.data_as(C.POINTER(C.c_double)), None)
    assert rc == 0, load_library().gap_last_error()
    assert len(calls) == 5 and all(s == calls[0] for s in calls)  # initial evaluation + 4 steps, always the same stream
    assert np.abs(d_pos.cpu().numpy() - at1.positions).max() < 1e-10
    assert np.abs(ep - ep1).max() < 1e-9


def test_deterministic_scatter_is_bitwise_reproducible(tmp_path, si_model, si_frames, golden):
    # gap_potential_set_deterministic: forces identical to the last bit between runs and between handles, equal to the atomic path within
    # rounding and to the oracle within the usual bar; also with skin-based list reuse, multi-species (general path) and the LAMMPS entry
    atoms, xml = syn.build_config_A(str(tmp_path), _oracle_desc, n_cells=4, M=120, seed=1)
    om = orc.Model(xml)
    o = om.calc(atoms)
    ref = Potential("", param_filename=xml).calc(atoms, force=True, virial=True)
    runs = []
    for k in range(3):
        p = Potential("", param_filename=xml)
        p.set_deterministic(True)
        for _ in range(2 + k):  # different call histories: exact list on the first call, speculative rows afterwards
            r = p.calc(atoms, force=True, virial=True)
        runs.append(r["force"].copy())
    assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2])
    assert np.abs(runs[0] - ref["force"]).max() < 1e-12
    assert np.abs(runs[0] - o["force"]).max() < TOL_F
    # skin reuse + deterministic
    p = Potential("", param_filename=xml)
    p.set_deterministic(True)
    p.set_cutoff_skin(0.4)
    rng = np.random.default_rng(2)
    pos = atoms.positions + rng.uniform(-0.03, 0.03, size=atoms.positions.shape)
    p.calc(atoms, force=True)
    r1 = p.calc(Atoms(atoms.numbers, pos, atoms.cell, True), force=True)
    assert p.connect_stats()["reuses"] == 1
    o1 = om.calc(Atoms(atoms.numbers, pos, atoms.cell, True))
    assert np.abs(r1["force"] - o1["force"]).max() < TOL_F
    # two-descriptor model (distance_2b + SOAP) and a general-path multi-species model
    pot, om2, xml2 = si_model
    q = Potential("IP GAP", param_filename=xml2)
    q.set_deterministic(True)
    a = si_frames[8]
    f1, f2 = q.calc(a, force=True)["force"], q.calc(a, force=True)["force"]
    assert np.array_equal(f1, f2) and np.abs(f1 - om2.calc(a)["force"]).max() < TOL_F
    desc = "soap n_Z=2 n_species=4 Z={23 42} species_Z={23 41 42 73} n_max=4 l_max=3 cutoff=4.5 atom_sigma=0.45 Z_mix=T K=3 coupling=F"
    ds = quad_datasets(golden, True)
    xml3 = multi_species_model(str(tmp_path), desc, ds, 8, seed=31, zeta=2.0)
    g = Potential("", param_filename=xml3)
    g.set_deterministic(True)
    om3 = orc.Model(xml3)
    for b in ds:
        f1, f2 = g.calc(b, force=True)["force"], g.calc(b, force=True)["force"]
        assert np.array_equal(f1, f2) and np.abs(f1 - om3.calc(b)["force"]).max() < TOL_F
