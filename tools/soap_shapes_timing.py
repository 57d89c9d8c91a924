"""SOAP stage times of the default path for several (n_max, l_max) on the config-A cell (4,096 Si atoms, cutoff 5, 2,000 sparse points): the
shapes with specialised warp-per-centre kernels -- (8,8), (12,8), and (10,6) with two species -- against the run-time-shape block kernels.
    python tools/soap_shapes_timing.py [n_max,l_max ...]     (needs a B200)"""
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quip_b200 import Potential  # noqa: E402
from quip_b200 import synthetic as syn  # noqa: E402
from quip_b200.gap_xml import write_gap_xml  # noqa: E402

shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(8, 8), (12, 8), (12, 6), (10, 12), (8, 4), (6, 6)]
atoms = syn.si_diamond()
out = {}
with tempfile.TemporaryDirectory() as tmp:
    for n, l in shapes:
        desc = "soap cutoff=5.0 cutoff_transition_width=0.5 n_max=%d l_max=%d atom_sigma=0.5 central_weight=1.0 n_species=1 Z=14 species_Z={14}" % (n, l)
        d = syn.soap_dimension(n, l)
        pot0 = Potential("", param_filename=syn.bootstrap_xml(os.path.join(tmp, "b%d_%d.xml" % (n, l)), [(desc, d)]))
        X = pot0.descriptor_calc(atoms, 0)[0]
        coord = syn.random_soap_coordinate(desc, X, 2000, delta=1.0, zeta=4.0, seed=3)
        pot = Potential("", param_filename=write_gap_xml(os.path.join(tmp, "m%d_%d.xml" % (n, l)), [coord], e0={14: 0.0}))
        pot.calc(atoms, force=True, virial=True)
        pot.set_timing(1)
        acc = {}
        for _ in range(5):
            pot.calc(atoms, force=True, virial=True)
            for k, v in pot.last_timings().items():
                acc.setdefault(k, []).append(v)
        out["n%d_l%d" % (n, l)] = {"d": d, **{k: round(float(np.median(v)), 4) for k, v in acc.items() if k in ("soap_forward", "soap_adjoint", "cov_gemm1", "cov_gemm2", "total")}}
print(json.dumps(out))
