"""Writer for GAP XML model files in the reference's format.

Produces what ``gap_fit_print_xml`` (src/GAP/gap_fit_module.f95:1586-1701) and ``gpCoordinates_printXML``
(src/GAP/gp_predict.f95:4196-4411) write, so that a synthetic model generated here (random-init sparse points and
alphas of the named descriptor hyper-parameters, BASELINE.json) can be fed unchanged to real QUIP elsewhere:
``<Potential>``, ``<GAP_params>`` with ``<GAP_data>/<e0>``, ``<gpSparse>`` with one ``<gpCoordinates>`` per
descriptor, sparse points either inline / sliced in chunks of 50 (:4289-4302) or in ``<xml>.sparseX.<label>``
side files, one ``%.20e`` per line, column-major (src/libAtoms/cutil.c:195-214) with their md5 sum.
"""
from __future__ import annotations

import hashlib
import os

import numpy as np

DEFAULT_GAP_VERSION = 1527075646  # the version tests/GAP.xml carries; any value >= 1426512068 selects current behaviour


def _r(v):
    return "%.17g" % float(v)


def write_gap_xml(path, coordinates, e0=None, label="GAP_b200_synthetic", gap_version=DEFAULT_GAP_VERSION, separate_files=True):
    """coordinates: list of dicts with keys
         descriptor (str), covariance_type (1 ard_se | 2 dot_product), delta, f0 (default 0), zeta (dot_product),
         theta (ard_se, sequence of d), sparseX (M, d) array, alpha (M,), sparseCutoff (M,) (default ones)
       e0: dict {Z: value}."""
    e0 = e0 or {}
    path = os.fspath(path)
    base = os.path.basename(path)
    out = []
    out.append("<%s>" % label)
    out.append('<Potential label="%s" init_args="IP GAP label=%s"/>' % (label, label))
    out.append('<GAP_params label="%s" gap_version="%d">' % (label, gap_version))
    out.append('  <GAP_data do_core="F">')
    for z in range(1, 117):
        out.append('    <e0 Z="%d" value="%s"/>' % (z, _r(e0.get(z, 0.0))))
    out.append("  </GAP_data>")
    out.append('  <gpSparse label="%s" n_coordinate="%d" fitted="T">' % (label, len(coordinates)))
    for i, c in enumerate(coordinates, start=1):
        X = np.ascontiguousarray(c["sparseX"], dtype=np.float64)
        M, d = X.shape
        alpha = np.asarray(c["alpha"], dtype=np.float64)
        cut = np.asarray(c.get("sparseCutoff", np.ones(M)), dtype=np.float64)
        clabel = "%s%d" % (label, i)
        attrs = ['label="%s"' % clabel, 'dimensions="%d"' % d, 'signal_variance="%s"' % _r(c["delta"]),
                 'signal_mean="%s"' % _r(c.get("f0", 0.0)), 'sparsified="T"', 'n_permutations="1"',
                 'covariance_type="%d"' % c["covariance_type"]]
        if c["covariance_type"] == 2:
            attrs.append('zeta="%s"' % _r(c["zeta"]))
        attrs.append('n_sparseX="%d"' % M)
        side = None
        if separate_files:
            side = "%s.sparseX.%s" % (base, clabel)
            text = "".join("%.20e\n" % v for v in X.reshape(-1))  # vector k = lines (k-1)d+1 .. kd
            with open(os.path.join(os.path.dirname(path) or ".", side), "w") as fh:
                fh.write(text)
            attrs.append('sparseX_filename="%s"' % side)
            attrs.append('sparseX_md5sum="%s"' % hashlib.md5(text.encode()).hexdigest())
        out.append("    <gpCoordinates %s>" % " ".join(attrs))
        if c["covariance_type"] == 1:
            out.append("      <theta>%s</theta>" % " ".join(_r(t) for t in np.atleast_1d(c["theta"])))
        out.append("      <descriptor>%s</descriptor>" % c["descriptor"])
        out.append('      <permutation i="1">%s</permutation>' % (" ".join(str(k) for k in range(1, d + 1)) if c["covariance_type"] == 1 else "1"))
        for k in range(M):
            head = '      <sparseX i="%d" alpha="%s" sparseCutoff="%s"' % (k + 1, _r(alpha[k]), _r(cut[k]))
            if side is not None:
                out.append(head + "/>")
            elif d > 50:
                out.append(head + ' sliced="T">')
                for s in range(0, d, 50):
                    e = min(s + 50, d)
                    out.append('        <sparseX_slice start="%d" end="%d">%s</sparseX_slice>' % (s + 1, e, " ".join(_r(v) for v in X[k, s:e])))
                out.append("      </sparseX>")
            else:
                out.append(head + ">%s</sparseX>" % " ".join(_r(v) for v in X[k]))
        out.append("    </gpCoordinates>")
    out.append("  </gpSparse>")
    out.append("</GAP_params>")
    out.append("</%s>" % label)
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")
    return path
