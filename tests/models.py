"""Model builders shared by the tests, smoke() and bench (fixtures only -- no compute)."""
import os

import numpy as np

from quip_b200.gap_xml import write_gap_xml

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SI_SOAP = ("soap atom_sigma=0.5 central_weight=1.0 covariance_type=dot_product cutoff=4.0 cutoff_transition_width=1.0 "
           "delta=3.0 l_max=8 n_max=8 n_sparse=100 zeta=4 sparse_method=cur_points n_species=1 Z=14 species_Z={14}")
SI_2B = ("distance_2b covariance_type=ard_se cutoff=6.0 delta=1.0 n_sparse=20 theta_uniform=0.1 sparse_method=uniform "
         "Z1=14 Z2=14")


def si_two_descriptor_model(workdir, separate_files=True):
    """BASELINE config[0]: the distance_2b + SOAP Si model the reference's own gap_fit test produces
    (tests/Si.two_descriptors.json; descriptor strings from tests/test_gapfit.py:75-80 plus the species keys gap_fit
    appends, descriptors.f95:3439), e0 = isolated-atom energy + e0_offset (gap_fit_module.f95:1519-1531)."""
    z = np.load(os.path.join(GOLDEN, "si_two_descriptors.npz"))
    coords = [
        {"descriptor": SI_2B, "covariance_type": 1, "delta": 1.0, "f0": 0.0, "theta": [0.1], "sparseX": z["sparsex_2b"].reshape(-1, 1),
         "alpha": z["alpha_2b"], "sparseCutoff": z["cutoff_2b"]},
        {"descriptor": SI_SOAP, "covariance_type": 2, "delta": 3.0, "f0": 0.0, "zeta": 4.0, "sparseX": z["sparsex_soap"],
         "alpha": z["alpha_soap"], "sparseCutoff": z["cutoff_soap"]},
    ]
    return write_gap_xml(os.path.join(workdir, "Si_two_descriptors.xml"), coords, e0={14: -158.54496821 + 2.0},
                         label="GAP_Si_two_descriptors", separate_files=separate_files)
