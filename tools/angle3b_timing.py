"""k_angle3b on a production-size cell: 4,096 Si atoms (BASELINE config A geometry), angle_3b cutoff 4.0 A (16 neighbours, 120 pairs per
centre) with 100 and 300 sparse points, default and deterministic mode; a 64-centre sample of local energies against the oracle.
    python tools/angle3b_timing.py     (needs a B200)"""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc  # noqa: E402  (checker only)
from quip_b200 import Potential  # noqa: E402
from quip_b200 import synthetic as syn  # noqa: E402
from quip_b200.gap_xml import write_gap_xml  # noqa: E402

atoms = syn.si_diamond()
rng = np.random.default_rng(5)
out = {"atoms": len(atoms)}
with tempfile.TemporaryDirectory() as tmp:
    for M in (100, 300):
        X = np.column_stack([rng.uniform(4.0, 7.5, size=M), rng.uniform(0.0, 2.0, size=M), rng.uniform(2.0, 6.5, size=M)])
        coord = {"descriptor": "angle_3b cutoff=4.0 cutoff_transition_width=0.5 Z_center=14 Z1=14 Z2=14", "covariance_type": 1, "delta": 0.5, "f0": 0.0,
                 "theta": [1.0, 0.8, 1.2], "sparseX": X, "alpha": rng.normal(0.0, 0.1, size=M), "sparseCutoff": np.ones(M)}
        xml = write_gap_xml(os.path.join(tmp, "a3b_%d.xml" % M), [coord], e0={14: 0.0})
        for det in (False, True):
            pot = Potential("", param_filename=xml)
            pot.set_deterministic(det)
            pot.calc(atoms, force=True, virial=True)
            pot.set_timing(1)
            ts = []
            for _ in range(5):
                r = pot.calc(atoms, force=True, virial=True, local_energy=True)
                ts.append(pot.last_timings()["distance_2b"])
            out["M%d_%s_ms" % (M, "deterministic" if det else "atomics")] = float(np.median(ts))
        o = orc.Model(xml).calc(atoms, local_energy=True, first=1000, last=1064, force=False, virial=False)
        out["M%d_max_dlocal_e_64_centres" % M] = float(np.abs(r["local_energy"][1000:1064] - o["local_energy"][1000:1064]).max())
print(json.dumps(out))
