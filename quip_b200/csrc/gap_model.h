// gap_model.h -- host-side GAP model: XML loader, key=value grammar, SOAP radial-basis set-up.
//
// Replaces (init path only, milliseconds, host):
//   IPModel_GAP_Initialise_str      src/Potentials/IPModel_GAP.f95:149-192
//   IPModel_GAP XML handlers        src/Potentials/IPModel_GAP.f95:618-944
//   gpSparse/gpCoordinates readXML  src/GAP/gp_predict.f95:4562-5057, 5200-5273
//   fread_array_d_                  src/libAtoms/cutil.c:195-214
//   param_read_line grammar         src/libAtoms/ParamReader.f95:393-518
//   soap_initialise                 src/GAP/descriptors.f95:2476-2642
//   distance_2b_initialise          src/GAP/descriptors.f95:1757-1818
//   angle_3b_initialise             src/GAP/descriptors.f95:1886-1911
#pragma once
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace gapb200 {

struct GapError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// ---- key=value argument strings -------------------------------------------------
std::vector<std::string> split_fields(const std::string& line);
struct ArgDict {
  std::map<std::string, std::string> kv;
  explicit ArgDict(const std::string& line);
  bool has(const std::string& k) const { return kv.count(k) != 0; }
  std::string str(const std::string& k, const std::string& def) const;
  double real(const std::string& k, double def) const;
  long integer(const std::string& k, long def) const;
  bool logical(const std::string& k, bool def) const;
  std::vector<int> int_list(const std::string& k) const;
};
double parse_real(const std::string& s);  // Fortran list-directed real: ".5", "1.0D0", "1E-001"

// ---- descriptors -----------------------------------------------------------------
struct SoapSpec {
  double cutoff = 0, cutoff_transition_width = 0.5, atom_sigma = 0, alpha = 0, central_weight = 1, covariance_sigma0 = 0;
  int cutoff_dexp = 0;
  double cutoff_scale = 1, cutoff_rate = 1;
  int l_max = 0, n_max = 0, n_Z = 1, n_species = 1;
  bool central_reference_all_species = false, normalise = true, do_two_l_plus_one = true;
  std::vector<int> Z, species_Z;
  std::vector<double> r_basis;          // n_max
  std::vector<double> transform_basis;  // n_max x n_max, column-major T(a,a')
  std::vector<double> cholesky_overlap; // n_max x n_max, column-major, lower
  int d = 0;
  int K1() const { return n_species * n_max; }
  // ---- general path: compression modes (Z_mix / R_mix / sym_mix / coupling / nu_R, nu_S / Z_map / diagonal_radial,
  //      descriptors.f95:7274-7670) and the GTO / POLY radial bases (:2643-2770, :8264-8278); general == false is the reference's
  //      "original" power spectrum (:7772-7775) on EQUISPACED_GAUSS
  bool general = false;
  bool global = false;          // average=T: ONE descriptor per configuration (descriptors.f95:2516, 8357-8367, 8738-9008)
  std::string radial_basis = "EQUISPACED_GAUSS";
  int n_grid = 0;               // radial points: n_max, or 3 n_max for GTO / POLY
  std::vector<double> r_grid;   // n_grid
  std::vector<double> P;        // [l][g][a]: radial_coefficient(l, a) = sum_g radial_fun(l, g) P[l][g][a]
  std::vector<double> c0;       // n_max: radial coefficients of the central atom's own Gaussian
  int Ka = 0, Kb = 0;           // columns of the mixing matrices W(1), W(2)
  std::vector<double> W1, W2;   // [K1][Ka], [K1][Kb] row-major
  std::vector<int> pair_ia, pair_jb;  // power-spectrum elements per l: tlpo_l * fac * sum_m Y1_lm(ia) Y2_lm(jb), index l + (l_max+1) k
  std::vector<double> pair_fac;
};
struct Distance2bSpec {
  double cutoff = 0, cutoff_transition_width = 0.5;
  int Z1 = 0, Z2 = 0;
  // descriptors.f95:1771-1815: data = r^exponents (one component per exponent), covariance_cutoff *= (erf(tail_range r) / r)^tail_exponent,
  // only_intra / only_inter against the residue ids of the atoms (the integer property resid_name)
  std::vector<double> exponents;  // n_exponents entries
  int tail_exponent = 0;
  double tail_range = 1.0;
  int intra_mode = 0;             // 0: all pairs, 1: only_intra, 2: only_inter
  std::string resid_name;
};
struct Angle3bSpec {  // angle_3b_initialise, descriptors.f95:1886-1911 (Z_center has the alternative key Z)
  double cutoff = 0, cutoff_transition_width = 0.5;
  int Zc = 0, Z1 = 0, Z2 = 0;
};
// soap_initialise; calc_xml_version = the xml_version soap_calc would see (<0: descriptor-only default)
SoapSpec soap_from_string(const std::string& desc, long calc_xml_version);
Distance2bSpec distance_2b_from_string(const std::string& desc);
Angle3bSpec angle_3b_from_string(const std::string& desc);

// ---- model --------------------------------------------------------------------------
enum { COVARIANCE_ARD_SE = 1, COVARIANCE_DOT_PRODUCT = 2 };
enum { DESC_DISTANCE_2B = 1, DESC_SOAP = 2, DESC_ANGLE_3B = 3 };

struct Coordinate {
  int kind = 0, covariance_type = 0, d = 0, M = 0, n_permutations = 1;
  double delta = 0, f0 = 0, zeta = 0;
  std::vector<double> theta, sparseX /* d x M column-major */, alpha, sparseCutoff;
  std::string label, descriptor_str;
  SoapSpec soap;
  Distance2bSpec d2b;
  Angle3bSpec a3b;
  double cutoff() const { return kind == DESC_SOAP ? soap.cutoff : (kind == DESC_ANGLE_3B ? a3b.cutoff : d2b.cutoff); }
};

struct GapModel {
  std::string label;
  long xml_version = 0;
  double e0[128];
  double E_scale = 1.0;
  double cutoff = 0.0;
  std::vector<Coordinate> coord;
};

// args_str: "IP GAP [label=...] [E_scale=...]" ; param_str: whole XML file ; base_dir: where sparseX side files live
GapModel load_gap_model(const std::string& args_str, const std::string& param_str, const std::string& base_dir);

std::string md5_hex(const std::string& bytes);

}  // namespace gapb200
