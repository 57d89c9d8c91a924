"""ctypes front end of the CPU oracle (oracle/gap_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / ``--impl reference`` legs of bench.py.  Nothing under quip_b200/
imports this module.

Besides the bindings it holds an *independent* (Python, ElementTree) reading of
the GAP XML format and of QUIP's ``key=value`` descriptor strings, so that the
product's C++ loader (quip_b200/csrc/gap_model.cpp) is checked against a second
implementation rather than against itself.  Reference lines followed:
  * XML tags/attributes: src/GAP/gp_predict.f95:4584-5057, 5200-5273;
    src/Potentials/IPModel_GAP.f95:618-700
  * sparseX side files: src/libAtoms/cutil.c:195-214
  * argument grammar: src/libAtoms/ParamReader.f95:393-518
  * SOAP string defaults / version switches: src/GAP/descriptors.f95:2502-2596
  * distance_2b defaults: src/GAP/descriptors.f95:1771-1782
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import xml.etree.ElementTree as ET

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libgaporacle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "gap_oracle.c")):
            build()
        L = C.CDLL(path)
        L.orc_connect_new.restype = C.c_void_p
        L.orc_connect_new.argtypes = [C.c_int, c_dp, c_dp, c_ip, C.c_double]
        L.orc_connect_free.argtypes = [C.c_void_p]
        L.orc_connect_total.argtypes = [C.c_void_p]
        L.orc_connect_cells.argtypes = [C.c_void_p, c_ip]
        L.orc_connect_get.argtypes = [C.c_void_p, c_ip, c_ip, c_ip, c_dp]
        L.orc_soap_new.restype = C.c_void_p
        L.orc_soap_new.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                   C.c_double, C.c_int, C.c_int, c_ip, C.c_int, c_ip, C.c_int, C.c_int, C.c_double,
                                   C.c_double, C.c_double]
        L.orc_soap_set_general.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp, c_dp, C.c_int, c_dp, C.c_int, c_dp, C.c_int, c_ip, c_ip, c_dp]
        L.orc_soap_set_global.argtypes = [C.c_void_p, C.c_int]
        L.orc_soap_is_global.argtypes = [C.c_void_p]
        L.orc_soap_calc_global.argtypes = [C.c_void_p, C.c_void_p, C.c_int, c_dp, c_ip, c_dp, c_ip, C.c_int, c_ip, c_dp, c_ip, c_dp, c_ip, c_dp, c_ip]
        L.orc_soap_free.argtypes = [C.c_void_p]
        L.orc_soap_dim.argtypes = [C.c_void_p]
        L.orc_soap_get_basis.argtypes = [C.c_void_p, c_dp, c_dp, c_dp]
        L.orc_soap_calc.argtypes = [C.c_void_p, C.c_void_p, C.c_int, c_dp, c_ip, c_dp, C.c_int, c_ip, c_dp, c_ip, c_ip,
                                    c_dp, c_ip, c_dp, c_ip]
        L.orc_model_new.restype = C.c_void_p
        L.orc_model_free.argtypes = [C.c_void_p]
        L.orc_model_set_e0.argtypes = [C.c_void_p, c_dp, C.c_int]
        L.orc_model_set_E_scale.argtypes = [C.c_void_p, C.c_double]
        L.orc_model_n_coord.argtypes = [C.c_void_p]
        L.orc_model_cutoff.restype = C.c_double
        L.orc_model_cutoff.argtypes = [C.c_void_p]
        L.orc_model_add_soap.argtypes = [C.c_void_p, C.c_void_p, C.c_int, c_dp, c_dp, c_dp, C.c_double, C.c_double]
        L.orc_model_add_distance_2b.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, c_dp,
                                                c_dp, c_dp, C.c_double, C.c_double, C.c_int, c_dp, c_dp, C.c_int, C.c_double, C.c_int]
        L.orc_model_add_angle_3b.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp,
                                             C.c_double, C.c_double, c_dp]
        L.orc_model_set_resid.argtypes = [C.c_void_p, c_ip]
        L.orc_model_calc.argtypes = [C.c_void_p, C.c_int, c_dp, c_ip, c_dp, c_ip, C.c_double, C.c_int, C.c_int, C.c_int,
                                     c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
        L.orc_model_set_extras.argtypes = [C.c_void_p, c_ip, c_dp, c_dp, c_dp, C.c_double]
        L.orc_model_predict.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, c_dp, c_dp]
        L.orc_set_blas.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
        _LIB = L
    return _LIB


def use_openblas(on=True):
    """Switch the two per-atom products of gp_predict between the oracle's vectorised loops and dgemv from scipy's bundled OpenBLAS
    (``scipy_dgemv_``; BASELINE.md section 3: the reference links a BLAS).  Returns True when the BLAS flavour is active."""
    if not on:
        lib().orc_set_blas(b"", b"", b"")
        return False
    import glob

    import scipy

    cands = glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas*.so"))
    for path in cands:
        if lib().orc_set_blas(path.encode(), b"scipy_dgemv_", b"scipy_openblas_set_num_threads") == 0:
            return True
    return False


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_ip) if a is not None else None


# ----------------------------------------------------------------------
# key=value strings (ParamReader.f95:393-518)
# ----------------------------------------------------------------------
def split_fields(line):
    """split on space/comma with {} "" '' grouping (split_string(..., matching=.true.))"""
    fields, cur, depth, quote = [], "", 0, None
    for ch in line:
        if quote:
            if ch == quote:
                quote = None
            else:
                cur += ch
        elif ch in "\"'" and depth == 0:
            quote = ch
        elif ch == "{":
            if depth > 0:
                cur += ch
            depth += 1
        elif ch == "}":
            depth -= 1
            if depth > 0:
                cur += ch
        elif ch in " ," and depth == 0:
            if cur:
                fields.append(cur)
            cur = ""
        else:
            cur += ch
    if cur:
        fields.append(cur)
    return fields


def parse_args(line):
    out = {}
    for fld in split_fields(line):
        if "=" in fld:
            k, v = fld.split("=", 1)
            out[k] = v
        else:
            out[fld] = "T"  # bare key => true (:447-449)
    return out


def _f(v):
    return float(str(v).strip().lower().replace("d", "e"))


def _b(v):
    return str(v).strip().upper() in ("T", "TRUE", ".TRUE.", "1")


def _ilist(v):
    return [int(t) for t in str(v).replace(",", " ").split()]


def soap_params(desc_str, calc_xml_version=None):
    """soap_initialise argument handling, descriptors.f95:2502-2596.

    ``calc_xml_version`` is the ``xml_version`` soap_calc sees in its args_str
    (IPModel_GAP.f95:361); None = the descriptor-only default 1423143769 (:7800)."""
    a = parse_args(desc_str)
    p = {}
    p["cutoff"] = _f(a["cutoff"])
    p["cutoff_transition_width"] = _f(a.get("cutoff_transition_width", 0.5))
    p["cutoff_dexp"] = int(a.get("cutoff_dexp", 0))
    p["cutoff_scale"] = _f(a.get("cutoff_scale", 1.0))
    p["cutoff_rate"] = _f(a.get("cutoff_rate", 1.0))
    p["l_max"] = int(a["l_max"])
    p["n_max"] = int(a["n_max"])
    p["atom_sigma"] = _f(a["atom_gaussian_width"] if "atom_gaussian_width" in a else a["atom_sigma"])
    p["central_weight"] = _f(a.get("central_weight", 1.0))
    has_cras = "central_reference_all_species" in a
    p["central_reference_all_species"] = _b(a.get("central_reference_all_species", "F"))
    p["covariance_sigma0"] = _f(a.get("covariance_sigma0", 0.0))
    p["normalise"] = _b(a.get("normalise", a.get("normalize", "T")))
    p["basis_error_exponent"] = _f(a.get("basis_error_exponent", 10.0))
    p["n_Z"] = int(a.get("n_Z", 1))
    has_n_species = "n_species" in a
    p["n_species"] = int(a.get("n_species", 1))
    xml_version = int(a.get("xml_version", 1426512068))
    if xml_version < 1426512068:
        p["central_reference_all_species"] = True
    if p["n_species"] == 1:
        p["species_Z"] = [int(a.get("species_Z", "0").split()[0])] if a.get("species_Z", "").strip() else [0]
    else:
        p["species_Z"] = _ilist(a["species_Z"])
    if not has_n_species and "species_Z" in a and a["species_Z"].strip():
        raise ValueError("soap: species_Z given without n_species")
    if not has_cras and p["n_species"] == 1:
        p["central_reference_all_species"] = True
    p["Z"] = _ilist(a.get("Z", "0")) if p["n_Z"] > 1 else [int(str(a.get("Z", "0")).split()[0])]
    p["global"] = _b(a.get("average", "F"))  # average=T: ONE descriptor per configuration from the summed density expansions (:2516)
    # ---- compression modes and radial bases (descriptors.f95:2529-2537): handled by the general path ----
    p["diagonal_radial"] = _b(a.get("diagonal_radial", "F"))
    p["Z_mix"], p["R_mix"], p["sym_mix"] = _b(a.get("Z_mix", "F")), _b(a.get("R_mix", "F")), _b(a.get("sym_mix", "F"))
    p["coupling"] = _b(a.get("coupling", "T"))
    p["nu_R"], p["nu_S"] = int(a.get("nu_R", 2)), int(a.get("nu_S", 2))
    p["K"], p["mix_shift"] = int(a.get("K", 0)), int(a.get("mix_shift", 0))
    p["Z_map"] = a.get("Z_map", "").strip()
    p["radial_basis"] = a.get("radial_basis", "") or "EQUISPACED_GAUSS"
    if p["radial_basis"] not in ("EQUISPACED_GAUSS", "GTO", "POLY"):
        raise ValueError("soap_initialise: radial_basis not recognised: EQUISPACED_GAUSS, POLY or GTO")
    p["general"] = (p["global"] or p["diagonal_radial"] or p["Z_mix"] or p["R_mix"] or p["sym_mix"] or not p["coupling"] or p["nu_R"] != 2 or p["nu_S"] != 2
                    or bool(p["Z_map"]) or p["radial_basis"] != "EQUISPACED_GAUSS")
    cv = 1423143769 if calc_xml_version is None else calc_xml_version
    p["do_two_l_plus_one"] = cv >= 1423143769
    return p


# ---- QUIP's own random numbers (src/libAtoms/System.f95:2458-2489, 2659-2701): the channel-mixing weights are drawn from them ----
class QuipRNG:
    A, M, Q, R = 16807, 2147483647, 127773, 2836  # Park-Miller minimal standard (System.f95:142-145)

    def __init__(self, seed):  # system_reseed_rng -> system_set_random_seeds: idum = seed, then 100 draws are discarded (:2474, :2484-2488)
        self.idum = int(seed)
        for _ in range(100):
            self.ran()

    def ran(self):  # :2659-2675
        k = self.idum // self.Q
        self.idum = self.A * (self.idum - k * self.Q) - self.R * k
        if self.idum < 0:
            self.idum += self.M
        return self.idum

    def uniform(self):  # :2678-2689 (RAN_MAX = huge(1))
        while True:
            u = self.ran() / 2147483647.0
            if u <= 1.0:
                return u

    def normal(self):  # :2692-2701
        while True:
            v1, v2 = 2.0 * self.uniform() - 1.0, 2.0 * self.uniform() - 1.0
            r = v1 * v1 + v2 * v2
            if r <= 1.0:
                return v1 * np.sqrt(-2.0 * np.log(r) / r)


def soap_mixing(p):
    """form_W (descriptors.f95:7618-7670) -> W(1), W(2), sym_desc, and the list of power-spectrum elements (ia, jb, factor) the
    unpacking loops produce (:8402-8416 for coupling=T, form_coupling_inds :7274-7350 for coupling=F)."""
    n, ns, K = p["n_max"], p["n_species"], p["K"]
    K1 = n * ns
    mixing = p["R_mix"] or p["Z_mix"] or p["sym_mix"]
    if (p["nu_R"] != 2 or p["nu_S"] != 2) and (mixing or p["diagonal_radial"] or p["Z_map"]):
        raise ValueError("(nu_R, nu_S) = (2,2) required to use channel mixing / diagonal radial / Zmap")
    if mixing and p["Z_map"]:
        raise ValueError("cant' using mixing and Zmap at the same time")
    if mixing:  # form_mix_W :7430-7533
        sym_desc = p["sym_mix"]
        W = []
        for j in (1, 2):
            if sym_desc and j == 2:
                W.append(W[0].copy())
                continue
            if p["R_mix"] and p["Z_mix"]:
                w = np.zeros((K1, K))
                for i_s in range(ns):
                    rng = QuipRNG(p["species_Z"][i_s] + p["mix_shift"] + j * 200)
                    for r in range(i_s * n, (i_s + 1) * n):
                        for c in range(K):
                            w[r, c] = rng.normal()
            elif p["Z_mix"]:
                w = np.zeros((K1, K * n))
                R = np.zeros((ns, K))
                for i_s in range(ns):
                    rng = QuipRNG(p["species_Z"][i_s] + p["mix_shift"] + j * 200)
                    for c in range(K):
                        R[i_s, c] = rng.normal()
                for i_s in range(ns):
                    for a in range(n):
                        for k in range(K):
                            w[i_s * n + a, k * n + a] = R[i_s, k]
            elif p["R_mix"]:
                w = np.zeros((K1, K * ns))
                rng = QuipRNG(n + p["mix_shift"] + j * 200)
                R = np.zeros((n, K))
                for r in range(n):
                    for c in range(K):
                        R[r, c] = rng.normal()
                for i_s in range(ns):
                    for a in range(n):
                        for k in range(K):
                            w[i_s * n + a, i_s * K + k] = R[a, k]
            else:
                raise ValueError("form_mix_W: not mixing anything")
            W.append(w)
    elif p["Z_map"]:  # form_Zmap_W :7536-7616
        zs = p["Z_map"]
        n_groups, dens = [1, 1], 0
        for ch in zs:
            if ch == ",":
                n_groups[dens] += 1
            if ch == ":":
                dens += 1
        two = dens == 1
        sym_desc = not two
        W = [np.zeros((K1, n * n_groups[0])), np.zeros((K1, n * (n_groups[1] if two else n_groups[0])))]
        i_group, i_density, tok = 0, 0, ""
        for ch in zs + " ":
            if ch.isdigit():
                tok += ch
                continue
            if tok:
                i_sp = p["species_Z"].index(int(tok))
                for a in range(n):
                    W[i_density][i_sp * n + a, i_group * n + a] = 1.0
                tok = ""
            if ch == ",":
                i_group += 1
            if ch == ":":
                i_density += 1
                i_group = 0
        if sym_desc:
            W[1] = W[0].copy()
    else:  # form_nu_W :7352-7424
        sym_desc = not (p["nu_R"] == 1 or p["nu_S"] == 1)
        nu_R, nu_S = p["nu_R"], p["nu_S"]
        if not (0 <= nu_R <= 2 and 0 <= nu_S <= 2):
            raise ValueError("nu_R / nu_S outside allowed range of 0-2")
        W = []
        for _ in (1, 2):
            dn = ds = 0
            n2_max = s2_max = 1
            if nu_R > 0:
                nu_R -= 1
                dn, n2_max = 1, n
            if nu_S > 0:
                nu_S -= 1
                ds, s2_max = 1, ns
            w = np.zeros((K1, n2_max * s2_max))
            for s_ in range(1, ns + 1):
                for a in range(1, n + 1):
                    ic = 0
                    for s2 in range(1, s2_max + 1):
                        for n2 in range(1, n2_max + 1):
                            if ds * s_ == ds * s2 and dn * a == dn * n2:
                                w[(s_ - 1) * n + a - 1, ic] = 1.0
                            ic += 1
            W.append(w)
    Ka, Kb = W[0].shape[1], W[1].shape[1]
    original = p["coupling"] and p["nu_R"] == 2 and p["nu_S"] == 2 and not mixing and not p["Z_map"]
    pairs = []
    if p["coupling"]:
        if p["diagonal_radial"] and not original:
            raise ValueError("soap_dimensions: can't combine diagonal radial with any other compression strategies")
        for ia in range(Ka):
            for jb in range(ia + 1 if sym_desc else Kb):
                if p["diagonal_radial"] and (ia % n) != (jb % n):  # rs_index(1, .) = radial index of the channel
                    continue
                pairs.append((ia, jb, np.sqrt(2.0) if (sym_desc and ia != jb) else 1.0))
    else:  # form_coupling_inds; sym_facs is declared `real` (single precision, :7283): its SQRT_TWO is rounded to float32
        sqrt2_f32 = float(np.float32(np.sqrt(2.0)))
        if Ka != Kb:
            raise ValueError("require K1=K2 to use elementwise coupling")
        if p["Z_mix"] and not p["R_mix"]:
            for k in range(K):
                for a in range(n):
                    for b in range(a + 1 if p["sym_mix"] else n):
                        pairs.append((k * n + a, k * n + b, sqrt2_f32 if (a != b and p["sym_mix"]) else 1.0))
        elif p["R_mix"] and not p["Z_mix"]:
            for i_s in range(ns):
                for j_s in range(i_s + 1 if p["sym_mix"] else ns):
                    for k in range(K):
                        pairs.append((i_s * K + k, j_s * K + k, sqrt2_f32 if (i_s != j_s and p["sym_mix"]) else 1.0))
        else:
            pairs = [(i, i, 1.0) for i in range(Ka)]
    return W[0], W[1], sym_desc, pairs


def soap_radial(p):
    """Radial points and the linear map radial_fun(l, :) -> radial_coefficient(l, :), per l, plus the central atom's coefficients.
    EQUISPACED_GAUSS: r_basis and transform_basis (descriptors.f95:2603-2642).  GTO / POLY (:2643-2770): a grid of 3 n_max points, the
    functions of the basis tabulated on it (B), orthonormalised with the Cholesky factor of their overlap, and the coefficients
    found by the least-squares (QR) solve  B (L^T)^-1 c = radial_fun  (:8264-8278), i.e. c = pinv(B L^-T) radial_fun."""
    from scipy.special import gamma, gammaincc

    n, L = p["n_max"], p["l_max"]
    alpha = 0.5 / p["atom_sigma"] ** 2
    cutoff_basis = p["cutoff"] + p["atom_sigma"] * np.sqrt(2.0 * p["basis_error_exponent"] * np.log(10.0))
    if p["radial_basis"] == "EQUISPACED_GAUSS":
        r, t, ch = soap_basis(p)
        return r, np.repeat(t[None, :, :], L + 1, axis=0), ch[0, :].copy()
    ng = 3 * n
    r = np.arange(ng) * (cutoff_basis / ng)
    P = np.zeros((L + 1, ng, n))
    l_ub = 0 if p["radial_basis"] == "POLY" else L
    for l in range(l_ub + 1):
        if p["radial_basis"] == "POLY":
            idx = np.arange(1, n + 1, dtype=np.float64)
            N_a = np.sqrt(cutoff_basis ** (2 * idx + 7) / ((idx + 3) * (2 * idx + 5) * (2 * idx + 7)))
            i, j = np.meshgrid(idx, idx, indexing="ij")
            S = 2 * cutoff_basis ** (i + j + 7) / ((5 + i + j) * (6 + i + j) * (7 + i + j)) / np.outer(N_a, N_a)
            B = (cutoff_basis - r[:, None]) ** (idx[None, :] + 2) / N_a[None, :]
        else:
            Rg = (cutoff_basis / n) * np.arange(1, n + 1)
            a_ln = -Rg ** (-2.0) * (np.log(0.001) - l * np.log(Rg))
            ag = a_ln[:, None] + a_ln[None, :]
            u, t = ag * cutoff_basis ** 2, l + 1.5
            S = 0.5 * cutoff_basis ** (2 * t) * u ** (-t) * (gamma(t) - gammaincc(t, u) * gamma(t))
            B = r[:, None] ** l * np.exp(-a_ln[None, :] * r[:, None] ** 2)
        Lc = np.linalg.cholesky(S)
        A = B @ np.linalg.inv(Lc.T)
        P[l] = np.linalg.pinv(A).T
    for l in range(l_ub + 1, L + 1):
        P[l] = P[0]
    c0 = np.exp(-alpha * r ** 2) @ P[0]
    return r, P, c0


def new_soap(p):
    Z = np.array(p["Z"], dtype=np.int32)
    sZ = np.array(p["species_Z"], dtype=np.int32)
    h = lib().orc_soap_new(p["cutoff"], p["cutoff_transition_width"], p["l_max"], p["n_max"], p["atom_sigma"],
                           p["central_weight"], int(p["central_reference_all_species"]), p["covariance_sigma0"],
                           int(p["normalise"]), len(Z), _ip(Z), len(sZ), _ip(sZ), int(p["do_two_l_plus_one"]),
                           p["cutoff_dexp"], p["cutoff_scale"], p["cutoff_rate"], p["basis_error_exponent"])
    if not h:
        raise RuntimeError("orc_soap_new failed")
    if p.get("general"):
        W1, W2, _, pairs = soap_mixing(p)
        r, P, c0 = soap_radial(p)
        W1, W2, P, r, c0 = (np.ascontiguousarray(x, dtype=np.float64) for x in (W1, W2, P, r, c0))
        ia = np.array([q[0] for q in pairs], dtype=np.int32)
        jb = np.array([q[1] for q in pairs], dtype=np.int32)
        fac = np.array([q[2] for q in pairs], dtype=np.float64)
        lib().orc_soap_set_general(h, len(r), _dp(r), _dp(P), _dp(c0), W1.shape[1], _dp(W1), W2.shape[1], _dp(W2), len(pairs), _ip(ia), _ip(jb), _dp(fac))
        lib().orc_soap_set_global(h, int(bool(p.get("global"))))
    return h


def soap_basis(p):
    h = new_soap({**p, "general": False})
    n = p["n_max"]
    r, t, ch = np.zeros(n), np.zeros((n, n), order="F"), np.zeros((n, n), order="F")
    lib().orc_soap_get_basis(h, _dp(r), _dp(t), _dp(ch))
    lib().orc_soap_free(h)
    return r, t, ch


# ----------------------------------------------------------------------
# geometry helpers
# ----------------------------------------------------------------------
def _geom(atoms):
    pos = np.ascontiguousarray(atoms.get_positions(), dtype=np.float64)
    Z = np.ascontiguousarray(atoms.get_atomic_numbers(), dtype=np.int32)
    lat = np.ascontiguousarray(np.asarray(atoms.get_cell(), dtype=np.float64).reshape(9))
    pbc = np.ascontiguousarray(np.asarray(atoms.get_pbc(), dtype=bool).astype(np.int32))
    return pos, Z, lat, pbc


class Connect:
    def __init__(self, atoms, cutoff):
        self.pos, self.Z, self.lat, self.pbc = _geom(atoms)
        self.N = len(self.Z)
        self.h = lib().orc_connect_new(self.N, _dp(self.pos), _dp(self.lat), _ip(self.pbc), float(cutoff))

    def arrays(self):
        tot = lib().orc_connect_total(self.h)
        off = np.zeros(self.N + 1, dtype=np.int32)
        j = np.zeros(max(tot, 1), dtype=np.int32)
        s = np.zeros((max(tot, 1), 3), dtype=np.int32)
        d = np.zeros(max(tot, 1))
        lib().orc_connect_get(self.h, _ip(off), _ip(j), _ip(s), _dp(d))
        return off, j[:tot], s[:tot], d[:tot]

    def cells(self):
        out = np.zeros(6, dtype=np.int32)
        lib().orc_connect_cells(self.h, _ip(out))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_connect_free(self.h)
            self.h = None


def soap_descriptor(desc_str, atoms, grad=False, cutoff=None):
    """quippy ``Descriptor(desc_str).calc(atoms, grad=...)`` analogue (quippy/descriptors.py:141-235):
    the neighbour list is built with ``descriptor cutoff + 1`` unless ``cutoff`` is given."""
    p = soap_params(desc_str)
    hs = new_soap(p)
    con = Connect(atoms, p["cutoff"] + 1.0 if cutoff is None else cutoff)
    d = lib().orc_soap_dim(hs)
    nrows = C.c_int(0)
    if p.get("global"):  # one descriptor per configuration; grad rows: (centre i, then its neighbours) for every centre
        nc = lib().orc_soap_calc_global(hs, con.h, con.N, _dp(con.pos), _ip(con.Z), _dp(con.lat), None, int(grad), C.byref(nrows), None, None,
                                        None, None, None, None)
        rows = nrows.value
        x, ci = np.zeros((1, d)), np.zeros(max(nc, 1), dtype=np.int32)
        g = np.zeros((max(rows, 1), 3, d)) if grad else None
        ii, gpos, hg = np.zeros(max(rows, 1), dtype=np.int32), np.zeros((max(rows, 1), 3)), np.zeros(max(rows, 1), dtype=np.int32)
        lib().orc_soap_calc_global(hs, con.h, con.N, _dp(con.pos), _ip(con.Z), _dp(con.lat), None, int(grad), C.byref(nrows), _dp(x), _ip(ci),
                                   _dp(g), _ip(ii), _dp(gpos), _ip(hg))
        lib().orc_soap_free(hs)
        out = {"data": x, "ci": ci[:nc], "row_off": np.array([0, rows], dtype=np.int32)}
        if grad:
            out.update(grad_data=g[:rows], ii=ii[:rows], pos=gpos[:rows], has_grad_data=hg[:rows].astype(bool),
                       grad_index_0based=np.stack([np.zeros(rows, dtype=np.int32), ii[:rows]], axis=1))
        return out
    nd = lib().orc_soap_calc(hs, con.h, con.N, _dp(con.pos), _ip(con.Z), _dp(con.lat), int(grad), C.byref(nrows),
                             None, None, None, None, None, None, None)
    rows = nrows.value
    x = np.zeros((nd, d))
    ci = np.zeros(nd, dtype=np.int32)
    row_off = np.zeros(nd + 1, dtype=np.int32)
    out = {}
    if grad:
        g = np.zeros((rows, 3, d))
        ii = np.zeros(rows, dtype=np.int32)
        gpos = np.zeros((rows, 3))
        hg = np.zeros(rows, dtype=np.int32)
        lib().orc_soap_calc(hs, con.h, con.N, _dp(con.pos), _ip(con.Z), _dp(con.lat), 1, C.byref(nrows), _dp(x), _ip(ci),
                            _ip(row_off), _dp(g), _ip(ii), _dp(gpos), _ip(hg))
        out.update(grad_data=g, ii=ii, pos=gpos, has_grad_data=hg.astype(bool))
        out["grad_index_0based"] = np.stack([np.repeat(ci, np.diff(row_off)), ii], axis=1)
    else:
        lib().orc_soap_calc(hs, con.h, con.N, _dp(con.pos), _ip(con.Z), _dp(con.lat), 0, C.byref(nrows), _dp(x), _ip(ci),
                            _ip(row_off), None, None, None, None)
    out.update(data=x, ci=ci, row_off=row_off)
    lib().orc_soap_free(hs)
    return out


# ----------------------------------------------------------------------
# GAP XML (independent reading)
# ----------------------------------------------------------------------
def _floats(text):
    return np.array([_f(t) for t in text.split()], dtype=np.float64) if text and text.strip() else np.zeros(0)


def load_gap_xml(path=None, xml_string=None, label=None):
    if xml_string is None:
        with open(path) as fh:
            xml_string = fh.read()
    base = os.path.dirname(os.path.abspath(path)) if path else os.getcwd()
    root = ET.fromstring(xml_string)
    params = None
    cands = [root] if root.tag == "GAP_params" else list(root.iter("GAP_params"))
    for gp in cands:  # first exact label match, else the first stanza (IPModel_GAP.f95:629-653)
        if label and gp.get("label") == label:
            params = gp
            break
    if params is None:
        if label and not any(True for _ in cands):
            raise ValueError("no GAP_params")
        params = cands[0] if (not label) else None
    if params is None:
        raise ValueError("GAP_params label %r not found" % label)
    model = {"label": params.get("label", ""), "xml_version": int(params.get("gap_version", 0)), "e0": np.zeros(128),
             "coordinates": []}
    gd = params.find("GAP_data")
    if gd is not None:
        if gd.get("e0") is not None:
            model["e0"][:] = _f(gd.get("e0"))
        for e in gd.findall("e0"):
            model["e0"][int(e.get("Z"))] = _f(e.get("value"))
    gs = params.find("gpSparse")
    if gs is None:
        raise ValueError("no gpSparse")
    if gs.get("fitted") is not None and not _b(gs.get("fitted")):
        raise ValueError("GAP model has not been fitted")
    n_coord = int(gs.get("n_coordinate"))
    coords = {c.get("label"): c for c in gs.findall("gpCoordinates")}
    for i in range(1, n_coord + 1):
        c = coords[gs.get("label") + str(i)]
        d, M = int(c.get("dimensions")), int(c.get("n_sparseX"))
        co = {"d": d, "M": M, "delta": _f(c.get("signal_variance")), "f0": _f(c.get("signal_mean")),
              "covariance_type": int(c.get("covariance_type")), "n_permutations": int(c.get("n_permutations")),
              "descriptor": "".join(c.find("descriptor").itertext()).strip(), "zeta": None}
        if c.get("zeta") is not None:
            co["zeta"] = _f(c.get("zeta"))
        th = c.find("theta")
        co["theta"] = _floats(th.text) if th is not None else np.zeros(0)
        if co["covariance_type"] == 2 and co["zeta"] is None:  # legacy: theta holds zeta (:4972-4977)
            co["zeta"] = float(co["theta"][0])
        alpha, cut = np.zeros(M), np.zeros(M)
        X = np.zeros((d, M), order="F")
        fn = c.get("sparseX_filename")
        if fn is not None:
            with open(os.path.join(base, fn)) as fh:
                X[:] = np.array([_f(t) for t in fh.read().split()]).reshape((d, M), order="F")
        for sx in c.findall("sparseX"):
            k = int(sx.get("i")) - 1
            alpha[k], cut[k] = _f(sx.get("alpha")), _f(sx.get("sparseCutoff"))
            if fn is None:
                if sx.get("sliced") is not None and _b(sx.get("sliced")):
                    for sl in sx.findall("sparseX_slice"):
                        X[int(sl.get("start")) - 1:int(sl.get("end")), k] = _floats(sl.text)
                else:
                    X[:, k] = _floats(sx.text)
        co.update(sparseX=X, alpha=alpha, sparseCutoff=cut)
        model["coordinates"].append(co)
    return model


class Model:
    """Oracle GAP model: ``Potential('IP GAP', param_filename=...)`` + ``calc`` analogue."""

    def __init__(self, path=None, xml_string=None, label=None, E_scale=1.0, model=None):
        self.spec = model if model is not None else load_gap_xml(path, xml_string, label)
        self.resid_names = []  # the resid_name properties the distance_2b coordinates with only_intra / only_inter read
        L = lib()
        self.h = L.orc_model_new()
        e0 = np.ascontiguousarray(self.spec["e0"], dtype=np.float64)
        L.orc_model_set_e0(self.h, _dp(e0), len(e0))
        L.orc_model_set_E_scale(self.h, float(E_scale))
        v = self.spec["xml_version"]
        for co in self.spec["coordinates"]:
            desc = co["descriptor"] + " xml_version=%d" % v  # IPModel_GAP.f95:184
            kind = split_fields(desc)[0]
            X = np.asfortranarray(co["sparseX"], dtype=np.float64)
            al = np.ascontiguousarray(co["alpha"])
            cu = np.ascontiguousarray(co["sparseCutoff"])
            if kind == "soap":
                if co["covariance_type"] != 2:
                    raise NotImplementedError("soap with covariance_type %d" % co["covariance_type"])
                p = soap_params(desc, calc_xml_version=v)
                hs = new_soap(p)
                assert L.orc_soap_dim(hs) == co["d"], (L.orc_soap_dim(hs), co["d"])
                L.orc_model_add_soap(self.h, hs, co["M"], _dp(X), _dp(al), _dp(cu), co["delta"], co["zeta"])
            elif kind == "distance_2b":
                a = parse_args(desc)  # distance_2b_initialise, descriptors.f95:1757-1815
                if co["covariance_type"] != 1 or co["n_permutations"] != 1:
                    raise NotImplementedError("distance_2b variant")
                n_exp = int(a.get("n_exponents", 1))
                if co["d"] != n_exp or n_exp > 8:
                    raise NotImplementedError("distance_2b dimensions")
                if "exponents" in a:
                    expo = [_f(t) for t in str(a["exponents"]).replace(",", " ").split()]
                else:
                    expo = [1.0] if n_exp == 1 else [-float(i) for i in range(1, n_exp + 1)]  # :1808-1812
                assert len(expo) == n_exp
                intra, inter = _b(a.get("only_intra", "F")), _b(a.get("only_inter", "F"))
                if intra and inter:
                    raise ValueError("distance_2b_initialise: cannot specify both only_inter AND only_intra")
                if (intra or inter) and "resid_name" not in a:
                    raise ValueError("distance_2b_initialise: only_intra and only_inter require resid_name to be given as well")
                self.resid_names.append(a.get("resid_name")) if (intra or inter) else None
                th = np.ascontiguousarray(co["theta"], dtype=np.float64)
                ex = np.ascontiguousarray(expo, dtype=np.float64)
                L.orc_model_add_distance_2b(self.h, _f(a.get("cutoff", 0.0)), _f(a.get("cutoff_transition_width", 0.5)),
                                            int(a.get("Z1", 0)), int(a.get("Z2", 0)), co["M"], _dp(X), _dp(al), _dp(cu),
                                            co["delta"], co["f0"], n_exp, _dp(th), _dp(ex), int(a.get("tail_exponent", 0)),
                                            _f(a.get("tail_range", 1.0)), 1 if intra else (2 if inter else 0))
            elif kind == "angle_3b":
                a = parse_args(desc)  # angle_3b_initialise, descriptors.f95:1886-1911 (Z_center has the alternative key Z)
                if co["covariance_type"] != 1 or co["n_permutations"] != 1 or co["d"] != 3:
                    raise NotImplementedError("angle_3b variant")
                th = np.ascontiguousarray(co["theta"], dtype=np.float64)
                L.orc_model_add_angle_3b(self.h, _f(a.get("cutoff", 0.0)), _f(a.get("cutoff_transition_width", 0.5)),
                                         int(a.get("Z_center", a.get("Z", 0))), int(a.get("Z1", 0)), int(a.get("Z2", 0)), co["M"],
                                         _dp(X), _dp(al), _dp(cu), co["delta"], co["f0"], _dp(th))
            else:
                raise NotImplementedError("descriptor %s" % kind)

    @property
    def cutoff(self):
        return lib().orc_model_cutoff(self.h)

    def calc(self, atoms, energy=True, force=True, virial=True, local_energy=False, local_virial=False,
             connect_cutoff=None, first=0, last=None, nthreads=0, atom_mask=None, energy_per_coordinate=False,
             local_gap_variance=False, gap_variance_regularisation=0.001):
        """atom_mask / energy_per_coordinate / local_gap_variance: the optional calc args of IPModel_GAP_Calc
        (IPModel_GAP.f95:324-337); the variance gradient is returned when forces or virials are requested (:560-564)."""
        pos, Z, lat, pbc = _geom(atoms)
        N = len(Z)
        last = N if last is None else last
        mask = None if atom_mask is None else np.ascontiguousarray(np.asarray(atom_mask, dtype=bool).astype(np.int32))
        epc = np.zeros(lib().orc_model_n_coord(self.h)) if energy_per_coordinate else None
        lgv = np.zeros(N) if local_gap_variance else None
        gvg = np.zeros((N, 3)) if (local_gap_variance and (force or virial or local_virial)) else None
        if mask is not None or epc is not None or lgv is not None:
            lib().orc_model_set_extras(self.h, _ip(mask), _dp(epc), _dp(lgv), _dp(gvg), float(gap_variance_regularisation))
        resid = None
        if self.resid_names:  # (one residue property serves all coordinates here; the reference reads each coordinate's own)
            arrays = getattr(atoms, "arrays", {})
            if self.resid_names[0] not in arrays:
                raise RuntimeError("distance_2b_calc did not find %s property (residue id) in the atoms object." % self.resid_names[0])
            resid = np.ascontiguousarray(arrays[self.resid_names[0]], dtype=np.int32)
        lib().orc_model_set_resid(self.h, _ip(resid))
        e = np.zeros(1)
        f = np.zeros((N, 3)) if force else None
        v = np.zeros((3, 3), order="F") if virial else None
        le = np.zeros(N) if local_energy else None
        lv = np.zeros((N, 9)) if local_virial else None
        t = np.zeros(3)
        rc = lib().orc_model_calc(self.h, N, _dp(pos), _ip(Z), _dp(lat), _ip(pbc),
                                  float(connect_cutoff if connect_cutoff is not None else self.cutoff), first, last,
                                  nthreads, _dp(e), _dp(le), _dp(f), _dp(v), _dp(lv), _dp(t))
        if rc:
            raise RuntimeError("gpCoordinates_Predict: variance_estimate: negative variance predicted" if rc == 5 else "orc_model_calc failed (rc=%d)" % rc)
        out = {"energy": float(e[0]), "timings": t}
        if epc is not None:
            out["energy_per_coordinate"] = epc
        if lgv is not None:
            out["local_gap_variance"] = lgv
        if gvg is not None:
            out["gap_variance_gradient"] = gvg
        if force:
            out["force"] = f
        if virial:
            out["virial"] = np.array(v)
        if local_energy:
            out["local_energy"] = le
        if local_virial:
            out["local_virial"] = lv
        return out

    def predict(self, icoord, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        n, d = x.shape
        e, g = np.zeros(n), np.zeros((n, d))
        lib().orc_model_predict(self.h, icoord, n, _dp(x), _dp(e), _dp(g))
        return e, g

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_model_free(self.h)
            self.h = None
