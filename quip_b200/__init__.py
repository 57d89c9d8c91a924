"""quip_b200 -- B200-native (sm_100a) drop-in for the GAP energy/force/virial path of libAtoms/QUIP.

Only what the path needs lives here: ``csrc/`` (hand-written CUDA kernels + the C ABI of include/gap_b200.h, built
into ``libgapb200.so``) and the host-side mirror of the reference's Python interface (``Potential``, ``Atoms``,
ext-XYZ reader, GAP XML writer, synthetic configurations).  There is no CPU fallback.
"""
from .atoms import Atoms, read_xyz  # noqa: F401
from .potential import Potential, ShardedPotential, load_library  # noqa: F401
