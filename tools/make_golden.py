#!/usr/bin/env python
"""Extract the golden vectors that pin the oracle from the reference's OWN test fixtures.

Run in the build container only (``/root/reference`` does not exist on the GPU
box); the outputs are committed under tests/golden/.  Nothing here executes
reference code (it is Fortran and cannot be built in this image, SURVEY.md
section 0): the numbers are the known answers stored in the reference's test
suite, cited per item.

  tests/GAP.xml (+3 sparseX side files), tests/gap_sample.xyz
      -> known-answer E/F of a distance_2b/ARD_SE GAP (tests/test_gappot.py:30-48)
  tests/test_potential_cell.py:38-42 -> five H2 energies
  tests/Si.np1.xyz, tests/Si.two_descriptors.json
      -> 100 SOAP sparse vectors (n_max=8,l_max=8) + 20 distance_2b sparse points,
         alphas (tests/test_gapfit.py:75-80,231-239)
  tests/SOAP_reference_data.json cases 112/114/116/119 (+ structures embedded in
      tests/test_SOAP.py:38) -> SOAP X and grad_data on the default path
  tests/SOAP_reference_data.json, all 122 cases -> soap_reference_all.{json,npz}
  tests/test_descriptor.py:63-226 -> C2H cell gradient index table + 2 gradient blocks
  src/GAP/doc_src/quippy-descriptor-tutorial.ipynb, the STORED OUTPUTS of cells 19-22, 31, 38-41, 45-52 (a run of the real QUIP binary,
      git f538cd9fe): distance_2b instances (count, distances, covariance_cutoff) and SOAP vectors (n_max=4, l_max=4; one species
      2 x 51, two species 3 x 181) of the 2-atom diamond cell (+ one H) -> descriptor_tutorial.json
"""
import ast
import json
import os
import shutil
import sys

import numpy as np

REF = "/root/reference/tests"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    for f in ["GAP.xml", "gap_sample.xyz", "Si.np1.xyz"] + [
            "GAP.xml.sparseX.GAP_2018_10_7_60_14_59_24_970%d" % i for i in (1, 2, 3)]:
        shutil.copyfile(os.path.join(REF, f), os.path.join(OUT, f))
        os.chmod(os.path.join(OUT, f), 0o644)

    # --- Si two-descriptor fit outputs ---------------------------------
    d = json.load(open(os.path.join(REF, "Si.two_descriptors.json")))
    c2b, csoap = d["coords"]
    np.savez_compressed(
        os.path.join(OUT, "si_two_descriptors.npz"),
        config=np.array(d["config"]),
        index_2b=np.array(d["index"][0]), index_soap=np.array(d["index"][1]),
        sparsex_2b=np.array(c2b["sparsex"]), alpha_2b=np.array(c2b["alpha"]), cutoff_2b=np.array(c2b["cutoff"]),
        sparsex_soap=np.array(csoap["sparsex"]).reshape(100, 325), alpha_soap=np.array(csoap["alpha"]),
        cutoff_soap=np.array(csoap["cutoff"]))

    # --- SOAP_reference_data.json, default-path cases --------------------
    ref = json.load(open(os.path.join(REF, "SOAP_reference_data.json")))
    src = open(os.path.join(REF, "test_SOAP.py")).read()
    tree = ast.parse(src)
    dataset_info = None
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "dataset_info":
            dataset_info = ast.literal_eval(node.value)
    assert dataset_info is not None
    cases = {}
    for i in (112, 114, 116, 119):
        e = ref[i]
        cases[str(i)] = {k: e[k] for k in ("quippy_str", "perm", "X", "grad_data", "grad_index_0based", "grad_perm",
                                           "dataset_name")}
    json.dump({"datasets": dataset_info, "cases": cases}, open(os.path.join(OUT, "soap_reference_cases.json"), "w"))

    # --- SOAP_reference_data.json, ALL 122 cases (compression modes, nu_R/nu_S, Z_map, diagonal_radial, GTO / POLY bases, average):
    #     descriptor strings as JSON, the numbers as one compressed npz (tests/test_SOAP.py:36-77 compares X and grad_data) ----------
    arrays, meta = {}, []
    for i, e in enumerate(ref):
        meta.append({"quippy_str": e["quippy_str"], "dataset_name": e["dataset_name"]})
        arrays["X_%d" % i] = np.array(e["X"], dtype=np.float64)
        arrays["G_%d" % i] = np.array(e["grad_data"], dtype=np.float64)
        arrays["GI_%d" % i] = np.array(e["grad_index_0based"], dtype=np.int32)
        arrays["perm_%d" % i] = np.array(e["perm"], dtype=np.int32)
        arrays["gperm_%d" % i] = np.array(e["grad_perm"], dtype=np.int32)
    json.dump(meta, open(os.path.join(OUT, "soap_reference_all.json"), "w"))
    np.savez_compressed(os.path.join(OUT, "soap_reference_all.npz"), **arrays)

    # --- test_descriptor.py C2H --------------------------------------------
    src = open(os.path.join(REF, "test_descriptor.py")).read()
    tree = ast.parse(src)
    vals = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Attribute):
            name = node.targets[0].attr
            if name in ("ref_grad_index_0based", "ref_grad_array") and isinstance(node.value, ast.Call):
                vals[name] = ast.literal_eval(node.value.args[0])
    json.dump({
        "positions": [[0., 0., 0.], [0.875, 0.875, 0.875], [0.2, 0.2, 0.1]], "numbers": [6, 6, 1],
        "cell": [[0.0, 1.75, 1.75], [1.75, 0.0, 1.75], [1.75, 1.75, 0.0]], "pbc": True,
        "descriptor": "soap cutoff=1.3 l_max=4 n_max=4 atom_sigma=0.5 n_Z=2 Z={1 6}",
        "ref_grad_index_0based": vals["ref_grad_index_0based"], "ref_grad_array": vals["ref_grad_array"],
        "shapes": {"descriptor": [3, 51], "grad": [7, 3, 51]}}, open(os.path.join(OUT, "c2h_descriptor.json"), "w"))

    # --- test_potential_cell.py -------------------------------------------
    json.dump({"cell_sizes": list(np.linspace(2.5, 4.5, 5)), "positions": [[0., 0., 0.], [1., 1., 1.]],
               "numbers": [1, 1],
               "ref_energies": [0.36747083829015637, 2.8715032700273735, 4.10632306979403, 5.518256035535996,
                                5.885656871424537]}, open(os.path.join(OUT, "h2_cell_energies.json"), "w"))
    # --- quippy-descriptor-tutorial.ipynb: outputs the reference binary printed -------------------------
    nb = json.load(open("/root/reference/src/GAP/doc_src/quippy-descriptor-tutorial.ipynb"))

    def cell_output(i):
        txt = ""
        for o in nb["cells"][i].get("outputs", []):
            if "text" in o:
                txt += "".join(o["text"])
            elif "data" in o and "text/plain" in o["data"]:
                txt += "".join(o["data"]["text/plain"])
        return eval(txt, {"array": np.array, "int32": np.int32, "__builtins__": {}})  # numpy reprs of dicts of arrays

    d2b, s1, s2 = cell_output(22), cell_output(41), cell_output(52)
    json.dump({
        "source": "src/GAP/doc_src/quippy-descriptor-tutorial.ipynb (stored cell outputs; structure: ase.build.bulk('C', 'diamond', 3.5), cells 5-7)",
        "cell": [[0.0, 1.75, 1.75], [1.75, 0.0, 1.75], [1.75, 1.75, 0.0]], "positions": [[0.0, 0.0, 0.0], [0.875, 0.875, 0.875]], "numbers": [6, 6],
        "extra_atom": {"number": 1, "position": [0.2, 0.2, 0.2]},
        "distance_2b": {"descriptor": "distance_2b Z1=6 Z2=6 cutoff=4", "count": 92, "n_cross": 184,
                        "data": [float(v) for v in d2b["data"][:, 0]], "covariance_cutoff": [float(v) for v in d2b["covariance_cutoff"]]},
        "soap_1": {"descriptor": "soap cutoff=3 l_max=4 n_max=4 atom_sigma=0.5 n_Z=1 Z={6} ", "sizes": [2, 58], "data": s1["data"].tolist()},
        "soap_2": {"descriptor": "soap cutoff=3 l_max=4 n_max=4 atom_sigma=0.5 n_Z=2 Z={1 6} n_species=2 species_Z={1 6}", "sizes": [3, 123],
                   "data": s2["data"].tolist()}}, open(os.path.join(OUT, "descriptor_tutorial.json"), "w"))
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print("  %-50s %8d B" % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == "__main__":
    sys.exit(main())
