"""Stage times of the general SOAP path (soap_general.cu: compression modes, GTO / POLY) next to the specialised default path on the config-A
cell: 4,096 Si atoms, n_max = l_max = 8, 2,000 sparse points (random rows of the descriptors themselves; only the timings matter here).
    python tools/soap_general_timing.py     (needs a B200)"""
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quip_b200 import Potential  # noqa: E402
from quip_b200 import synthetic as syn  # noqa: E402
from quip_b200.gap_xml import write_gap_xml  # noqa: E402

atoms = syn.si_diamond()
base = "soap cutoff=5.0 cutoff_transition_width=0.5 n_max=8 l_max=8 atom_sigma=0.5 central_weight=1.0 n_species=1 Z=14 species_Z={14}"
variants = {"default": "", "GTO": " radial_basis=GTO", "POLY": " radial_basis=POLY", "R_mix K=4": " R_mix=T K=4", "nu_R=1": " nu_R=1", "diagonal_radial": " diagonal_radial=T"}
out = {}
with tempfile.TemporaryDirectory() as tmp:
    for name, extra in variants.items():
        desc = base + extra
        dim = 325  # the library's own error message gives the dimension of a variant
        try:
            pot0 = Potential("", param_filename=syn.bootstrap_xml(os.path.join(tmp, "b_%s.xml" % len(out)), [(desc, dim)]))
        except RuntimeError as e:  # "... does not match soap descriptor dimension N"
            dim = int(str(e).strip().split()[-1])
            pot0 = Potential("", param_filename=syn.bootstrap_xml(os.path.join(tmp, "b_%s.xml" % len(out)), [(desc, dim)]))
        X = pot0.descriptor_calc(atoms, 0)[0]
        coord = syn.random_soap_coordinate(desc, X, 2000, delta=1.0, zeta=4.0, seed=3)
        pot = Potential("", param_filename=write_gap_xml(os.path.join(tmp, "m_%s.xml" % len(out)), [coord], e0={14: 0.0}))
        pot.calc(atoms, force=True, virial=True)
        pot.set_timing(1)
        acc = {}
        for _ in range(5):
            pot.calc(atoms, force=True, virial=True)
            for k, v in pot.last_timings().items():
                acc.setdefault(k, []).append(v)
        out[name] = {"d": dim, **{k: round(float(np.median(v)), 4) for k, v in acc.items() if k in ("soap_forward", "soap_adjoint", "cov_gemm1", "cov_gemm2", "total")}}
print(json.dumps(out))
