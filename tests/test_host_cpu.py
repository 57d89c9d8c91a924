"""CPU-side checks of the product's host logic (no compute, no GPU): the C ABI exports what include/gap_b200.h
declares, the C++ model loader agrees with the independent Python reading used by the oracle, the XML writer
round-trips, and the reference's error conditions are reported."""
import os
import re

import numpy as np
import pytest

from oracle import oracle as orc
from quip_b200 import potential as P
from quip_b200.gap_xml import write_gap_xml
from tests.models import si_two_descriptor_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse_describe(text):
    out = {"coordinate": {}, "soap": {}, "distance_2b": {}, "vec": []}
    cur = None
    for line in text.splitlines():
        t = line.split()
        if t[0] in ("coordinate", "soap", "distance_2b"):
            cur = int(t[1])
            out[t[0]][cur] = {t[k]: float(t[k + 1]) for k in range(2, len(t) - 1, 2)}
        elif t[0] in ("species_Z", "Z", "r_basis", "transform_basis", "cholesky_overlap"):
            out["soap"][cur][t[0]] = np.array([float(v) for v in t[1:]])
        elif t[0] == "e0":
            out["e0"] = {int(a.split(":")[0]): float(a.split(":")[1]) for a in t[1:]}
        else:
            out[t[0]] = t[1] if len(t) > 1 else ""
    return out


def test_abi_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "gap_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b((?:gap|quip_lammps)_\w+)\s*\(", hdr))
    assert declared == set(P.ABI), declared ^ set(P.ABI)
    lib = P.load_library()
    for name in declared:
        assert hasattr(lib, name)


def test_initialise_fails_loudly_without_gpu(golden):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        P.Potential("IP GAP", param_filename=os.path.join(golden, "GAP.xml"))


def test_loader_matches_independent_reader(golden, tmp_path):
    for xml in (os.path.join(golden, "GAP.xml"), si_two_descriptor_model(str(tmp_path)),
                si_two_descriptor_model(str(tmp_path / "inline"), separate_files=False) if (tmp_path / "inline").mkdir() is None else None):
        d = parse_describe(P.model_describe(param_filename=xml))
        m = orc.load_gap_xml(xml)
        assert d["label"] == m["label"]
        assert int(d["xml_version"]) == m["xml_version"]
        assert int(d["n_coordinate"]) == len(m["coordinates"])
        for z, v in d["e0"].items():
            assert m["e0"][z] == v
        assert np.count_nonzero(m["e0"]) == len(d["e0"])
        for i, co in enumerate(m["coordinates"]):
            c = d["coordinate"][i]
            assert c["d"] == co["d"] and c["M"] == co["M"] and c["covariance_type"] == co["covariance_type"]
            assert c["delta"] == co["delta"] and c["f0"] == co["f0"]
            w = (np.arange(co["sparseX"].size) % 7 + 1.0)
            assert abs(c["sparseX_wsum"] - float(np.sum(co["sparseX"].reshape(-1, order="F") * w))) <= 1e-9 * abs(c["sparseX_wsum"])
            assert abs(c["alpha_sum"] - co["alpha"].sum()) <= 1e-9 * max(1.0, abs(c["alpha_sum"]))
            assert abs(c["sparseCutoff_sum"] - co["sparseCutoff"].sum()) < 1e-12
            if co["covariance_type"] == 2:
                assert c["zeta"] == co["zeta"]
                sp = orc.soap_params(co["descriptor"] + " xml_version=%d" % m["xml_version"], calc_xml_version=m["xml_version"])
                s = d["soap"][i]
                for k in ("l_max", "n_max", "n_species", "n_Z"):
                    assert s[k] == sp[k]
                assert bool(s["cras"]) == sp["central_reference_all_species"] and bool(s["two_lp1"]) == sp["do_two_l_plus_one"]
                r, T, ch = orc.soap_basis(sp)
                assert np.abs(s["r_basis"] - r).max() < 1e-14
                assert np.abs(s["transform_basis"] - T.reshape(-1, order="F")).max() < 1e-11 * np.abs(T).max()
                assert np.abs(s["cholesky_overlap"] - ch.reshape(-1, order="F")).max() < 1e-13
            else:
                assert c["theta0"] == co["theta"][0]
                a = orc.parse_args(co["descriptor"])
                assert d["distance_2b"][i]["Z1"] == int(a["Z1"]) and d["distance_2b"][i]["Z2"] == int(a["Z2"])
                assert c["cutoff"] == float(a["cutoff"])


def test_writer_roundtrip_sliced_inline(tmp_path):
    rng = np.random.default_rng(0)
    X = rng.random((7, 51))
    c = {"descriptor": "soap cutoff=3.0 l_max=4 n_max=4 atom_sigma=0.5 n_Z=1 Z=6 n_species=1 species_Z={6}", "covariance_type": 2,
         "delta": 0.7, "zeta": 2.0, "sparseX": X, "alpha": rng.normal(size=7), "sparseCutoff": rng.random(7)}
    for sep in (True, False):
        xml = write_gap_xml(str(tmp_path / ("m%d.xml" % sep)), [c], e0={6: -1.5}, separate_files=sep)
        m = orc.load_gap_xml(xml)
        assert np.array_equal(m["coordinates"][0]["sparseX"].T, X)
        assert np.array_equal(m["coordinates"][0]["alpha"], c["alpha"])
        d = parse_describe(P.model_describe(param_filename=xml))
        w = (np.arange(X.size) % 7 + 1.0)
        assert abs(d["coordinate"][0]["sparseX_wsum"] - float(np.sum(X.reshape(-1) * w))) < 1e-10
        assert d["e0"] == {6: -1.5}


def test_reference_error_conditions(golden, tmp_path):
    xml = open(os.path.join(golden, "GAP.xml")).read()
    with pytest.raises(RuntimeError, match="could not initialise GAP potential"):  # IPModel_GAP.f95:939-940
        P.model_describe(param_str=xml, args_str="IP GAP label=nonexistent", base_dir=golden)
    with pytest.raises(RuntimeError, match="does not exist"):  # gp_predict.f95:4719
        P.model_describe(param_str=xml, args_str="IP GAP", base_dir=str(tmp_path))
    with pytest.raises(RuntimeError, match="IP GAP"):
        P.model_describe(param_str=xml, args_str="IP SW", base_dir=golden)
    bad = si_two_descriptor_model(str(tmp_path))
    side = [f for f in os.listdir(tmp_path) if "sparseX" in f][0]
    with open(tmp_path / side, "a") as fh:
        fh.write("1.0\n")
    with pytest.raises(RuntimeError, match="md5 check sum failed"):  # gp_predict.f95:4736-4739
        P.model_describe(param_filename=bad)
    unfitted = xml.replace('n_coordinate="3"', 'n_coordinate="3" fitted="F"')
    with pytest.raises(RuntimeError, match="has not been fitted"):  # IPModel_GAP.f95:179
        P.model_describe(param_str=unfitted, base_dir=golden)
    other = xml.replace("distance_2b cutoff=4.0", "co_angle_3b cutoff=4.0", 1)
    with pytest.raises(RuntimeError, match="not supported"):
        P.model_describe(param_str=other, base_dir=golden)


def test_key_value_grammar_agrees():
    # ParamReader.f95:393-518 grouping rules, product (via describe of a soap string) vs oracle splitter
    s = "soap cutoff=4.0 l_max=2 n_max=3 atom_sigma=0.5 n_Z=2 Z={1 8} n_species=2 species_Z={8 1} normalise=F"
    assert orc.split_fields(s) == ["soap", "cutoff=4.0", "l_max=2", "n_max=3", "atom_sigma=0.5", "n_Z=2", "Z=1 8", "n_species=2",
                                   "species_Z=8 1", "normalise=F"]
    sp = orc.soap_params(s)
    assert sp["Z"] == [1, 8] and sp["species_Z"] == [8, 1] and not sp["normalise"] and not sp["central_reference_all_species"]


def test_cli_argument_grammar():
    from quip_b200 import cli

    c = cli.parse_cli(["atoms_filename=a.xyz", "param_filename=p.xml", "E", "forces", "V=T", 'calc_args={only_descriptor=2}', "init_args=IP GAP"])
    assert c["E"] and c["F"] and c["V"] and not c["local"]
    assert c["atoms_filename"] == "a.xyz" and c["calc_args"] == "only_descriptor=2"
    import pytest

    with pytest.raises(RuntimeError, match="outside the GAP evaluation path"):
        cli.parse_cli(["relax=T"])


def test_general_soap_setup_matches_oracle(golden, tmp_path):
    """The C++ loader's general-SOAP tables (mixing matrices with QUIP's random weights, power-spectrum element list, GTO / POLY
    radial maps; gap_model.cpp soap_general_setup) against the oracle's independent numpy derivation, for every descriptor string of
    the reference's SOAP_reference_data.json that takes the general path (average=T strings included)."""
    import json

    meta = json.load(open(os.path.join(golden, "soap_reference_all.json")))
    n_general = 0
    for m in meta:
        qs = m["quippy_str"]
        p = orc.soap_params(qs)
        W1, W2, _, pairs = orc.soap_mixing(p)
        d = (p["l_max"] + 1) * len(pairs) + 1
        coord = {"descriptor": qs, "covariance_type": 2, "delta": 1.0, "zeta": 2.0, "sparseX": np.zeros((1, d)), "alpha": np.zeros(1)}
        xml = write_gap_xml(str(tmp_path / "g.xml"), [coord], separate_files=False)
        text = P.model_describe(param_filename=xml)
        if not p["general"]:
            assert "soap_general" not in text
            continue
        n_general += 1
        rows = {ln.split()[0]: ln.split()[1:] for ln in text.splitlines() if ln.split()[0] in ("soap_general", "r_grid", "P", "c0", "W1", "W2", "pairs")}
        hdr = dict(zip(rows["soap_general"][1::2], rows["soap_general"][2::2]))
        assert int(hdr["Ka"]) == W1.shape[1] and int(hdr["Kb"]) == W2.shape[1] and int(hdr["n_pairs"]) == len(pairs), qs
        assert np.abs(np.array(rows["W1"], dtype=float).reshape(W1.shape) - W1).max() < 1e-14, qs
        assert np.abs(np.array(rows["W2"], dtype=float).reshape(W2.shape) - W2).max() < 1e-14, qs
        got = [tuple(float(v) for v in t.split(":")) for t in rows["pairs"]]
        assert all(g[0] == q[0] and g[1] == q[1] and abs(g[2] - q[2]) < 1e-15 for g, q in zip(got, pairs)), qs
        r, Pm, c0 = orc.soap_radial(p)
        assert int(hdr["n_grid"]) == len(r)
        assert np.abs(np.array(rows["r_grid"], dtype=float) - r).max() < 1e-13
        Pc = np.array(rows["P"], dtype=float).reshape(Pm.shape)
        assert np.abs(Pc - Pm).max() < 1e-9 * max(1.0, np.abs(Pm).max()), (qs, np.abs(Pc - Pm).max(), np.abs(Pm).max())
        assert np.abs(np.array(rows["c0"], dtype=float) - c0).max() < 1e-9 * max(1.0, np.abs(c0).max()), qs
    assert n_general >= 110


def _build_c_example(tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "wrapper_simple_example")
    subprocess.run(["gcc", "-Wall", "-Werror", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", "wrapper_simple_example.c"),
                    "-L" + os.path.join(root, "quip_b200"), "-lgapb200", "-Wl,-rpath," + os.path.join(root, "quip_b200"), "-o", exe], check=True)
    return exe


def test_c_example_compiles_against_the_header_and_fails_loudly_without_a_gpu(golden, tmp_path):
    # the C caller of INTEGRATION.md section 4 (the reference's quip_wrapper_simple_example_C.c): plain C, include/gap_b200.h, -lgapb200
    import subprocess
    exe = _build_c_example(tmp_path)
    r = subprocess.run([exe, os.path.join(golden, "GAP.xml")], capture_output=True, text=True)
    if r.returncode == 0:  # a CUDA device is present: the example ran
        assert "Energy =" in r.stdout
    else:
        assert "no CPU fallback" in r.stderr


def test_xml_with_comments_processing_instructions_and_embedded_training_data(golden):
    # a GAP XML as gap_fit writes it with do_copy_at_file=T: prolog, comments, and the training set inside a CDATA block whose text contains
    # angle brackets, quotes and ampersands (IPModel_GAP.f95:618-700 goes through FoX; the scanner here has to skip all of it)
    xml = open(os.path.join(golden, "GAP.xml")).read()
    wrapped = ('<?xml version="1.0"?>\n<!-- fitted by gap_fit <test> -->\n<GAP_wrapper>\n' + xml +
               '\n<XYZ_data compression="none"><![CDATA[2\nLattice="10 0 0 0 10 0 0 0 10" Properties=species:S:1:pos:R:3 note="a<b>c & d"\n'
               'H 0 0 0\nH 0 0 0.74\n]]></XYZ_data>\n</GAP_wrapper>\n')
    assert P.model_describe(param_str=wrapped, base_dir=golden) == P.model_describe(param_str=xml, base_dir=golden)
