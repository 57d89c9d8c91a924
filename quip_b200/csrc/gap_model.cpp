// gap_model.cpp -- see gap_model.h for the reference lines each piece replaces.
#include "gap_model.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <sstream>

namespace gapb200 {

// ======================================================================================
// MD5 (RFC 1321) -- the reference verifies sparseX side files against sparseX_md5sum
// (gp_predict.f95:4721-4739, src/libAtoms/md5.c)
// ======================================================================================
std::string md5_hex(const std::string& msg) {
  static const uint32_t K[64] = {
      0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af,
      0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa,
      0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8,
      0x676f02d9, 0x8d2a4c8a, 0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
      0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97,
      0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1,
      0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
  static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9,  14, 20, 5, 9,
                            14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                            4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
  uint32_t a0 = 0x67452301, b0 = 0xefcdab89, c0 = 0x98badcfe, d0 = 0x10325476;
  std::string m = msg;
  uint64_t bitlen = (uint64_t)msg.size() * 8;
  m.push_back((char)0x80);
  while (m.size() % 64 != 56) m.push_back(0);
  for (int i = 0; i < 8; i++) m.push_back((char)((bitlen >> (8 * i)) & 0xff));
  for (size_t off = 0; off < m.size(); off += 64) {
    uint32_t w[16];
    for (int i = 0; i < 16; i++)
      w[i] = (uint32_t)(uint8_t)m[off + 4 * i] | ((uint32_t)(uint8_t)m[off + 4 * i + 1] << 8) |
             ((uint32_t)(uint8_t)m[off + 4 * i + 2] << 16) | ((uint32_t)(uint8_t)m[off + 4 * i + 3] << 24);
    uint32_t A = a0, B = b0, Cc = c0, D = d0;
    for (int i = 0; i < 64; i++) {
      uint32_t F;
      int g;
      if (i < 16) { F = (B & Cc) | (~B & D); g = i; }
      else if (i < 32) { F = (D & B) | (~D & Cc); g = (5 * i + 1) % 16; }
      else if (i < 48) { F = B ^ Cc ^ D; g = (3 * i + 5) % 16; }
      else { F = Cc ^ (B | ~D); g = (7 * i) % 16; }
      F = F + A + K[i] + w[g];
      A = D; D = Cc; Cc = B;
      B = B + ((F << S[i]) | (F >> (32 - S[i])));
    }
    a0 += A; b0 += B; c0 += Cc; d0 += D;
  }
  char out[33];
  uint32_t v[4] = {a0, b0, c0, d0};
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) snprintf(out + 8 * i + 2 * j, 3, "%02x", (v[i] >> (8 * j)) & 0xff);
  return std::string(out, 32);
}

// ======================================================================================
// key=value grammar (ParamReader.f95:393-518; split_string with {} "" '' grouping :422)
// ======================================================================================
std::vector<std::string> split_fields(const std::string& line) {
  std::vector<std::string> out;
  std::string cur;
  int depth = 0;
  char quote = 0;
  for (char ch : line) {
    if (quote) {
      if (ch == quote) quote = 0; else cur.push_back(ch);
    } else if ((ch == '"' || ch == '\'') && depth == 0) {
      quote = ch;
    } else if (ch == '{') {
      if (depth > 0) cur.push_back(ch);
      depth++;
    } else if (ch == '}') {
      depth--;
      if (depth > 0) cur.push_back(ch);
      if (depth < 0) throw GapError("param_read_line: unmatched '}' in '" + line + "'");
    } else if ((ch == ' ' || ch == ',' || ch == '\t' || ch == '\n' || ch == '\r') && depth == 0) {
      if (!cur.empty()) out.push_back(cur);
      cur.clear();
    } else {
      cur.push_back(ch);
    }
  }
  if (!cur.empty()) out.push_back(cur);
  return out;
}

ArgDict::ArgDict(const std::string& line) {
  for (const std::string& f : split_fields(line)) {
    size_t eq = f.find('=');
    if (eq == std::string::npos) kv[f] = "T";  // bare key => true (:447-449)
    else if (eq == 0) throw GapError("Malformed field '" + f + "'");
    else kv[f.substr(0, eq)] = f.substr(eq + 1);
  }
}
std::string ArgDict::str(const std::string& k, const std::string& def) const {
  auto it = kv.find(k);
  return it == kv.end() ? def : it->second;
}
double parse_real(const std::string& s0) {
  std::string s;
  for (char c : s0) {
    if (c == 'd' || c == 'D') s.push_back('e');
    else if (!isspace((unsigned char)c)) s.push_back(c);
  }
  if (s.empty()) throw GapError("cannot parse empty string as real");
  char* end = nullptr;
  double v = strtod(s.c_str(), &end);
  if (end == s.c_str() || *end != 0) throw GapError("cannot parse '" + s0 + "' as real");
  return v;
}
static std::vector<double> parse_reals(const std::string& text) {
  std::vector<double> out;
  std::istringstream is(text);
  std::string tok;
  while (is >> tok) out.push_back(parse_real(tok));
  return out;
}
double ArgDict::real(const std::string& k, double def) const { return has(k) ? parse_real(kv.at(k)) : def; }
long ArgDict::integer(const std::string& k, long def) const {
  if (!has(k)) return def;
  const std::string& s = kv.at(k);
  char* end = nullptr;
  long v = strtol(s.c_str(), &end, 10);
  if (end == s.c_str()) throw GapError("cannot parse '" + s + "' as integer for key " + k);
  return v;
}
bool ArgDict::logical(const std::string& k, bool def) const {
  if (!has(k)) return def;
  std::string s = kv.at(k);
  if (s.empty()) return true;
  char c = s[0] == '.' && s.size() > 1 ? s[1] : s[0];
  return c == 'T' || c == 't' || c == '1';
}
std::vector<int> ArgDict::int_list(const std::string& k) const {
  std::vector<int> out;
  if (!has(k)) return out;
  std::string s = kv.at(k);
  for (char& c : s)
    if (c == ',' || c == '{' || c == '}') c = ' ';
  std::istringstream is(s);
  int v;
  while (is >> v) out.push_back(v);
  return out;
}

// ======================================================================================
// SOAP set-up (descriptors.f95:2476-2642, EQUISPACED_GAUSS branch)
// ======================================================================================
static void chol_lower(std::vector<double>& a, int n, const char* what) {
  for (int j = 0; j < n; j++) {
    double s = a[j + n * j];
    for (int k = 0; k < j; k++) s -= a[j + n * k] * a[j + n * k];
    if (!(s > 0.0)) throw GapError(std::string("LA_Matrix_Factorise: cannot factorise ") + what);
    double ljj = std::sqrt(s);
    a[j + n * j] = ljj;
    for (int i = j + 1; i < n; i++) {
      double t = a[i + n * j];
      for (int k = 0; k < j; k++) t -= a[i + n * k] * a[j + n * k];
      a[i + n * j] = t / ljj;
    }
  }
}


// ======================================================================================
// general SOAP set-up: mixing matrices, power-spectrum element list, GTO / POLY radial maps
// ======================================================================================
namespace {
// QUIP's random numbers (src/libAtoms/System.f95): Park-Miller minimal standard generator (:142-145, 2659-2675), seeded by
// system_reseed_rng -> idum = seed, 100 draws discarded (:2474, 2484-2488); ran_uniform = ran / huge(1) (:2678-2689); ran_normal by the
// polar method (:2692-2701).  The channel-mixing weights of form_mix_W are drawn from it, so a model fitted with QUIP needs these numbers.
struct QuipRng {
  long idum;
  explicit QuipRng(long seed) : idum(seed) {
    if (idum == 0) throw GapError("function ran(): linear-congruential random number generators fail with seed idum=0");
    for (int i = 0; i < 100; i++) ran();
  }
  long ran() {
    const long k = idum / 127773;
    idum = 16807 * (idum - k * 127773) - 2836 * k;
    if (idum < 0) idum += 2147483647;
    return idum;
  }
  double uniform() {
    double u = 1.1;
    while (u > 1.0) u = (double)ran() / 2147483647.0;
    return u;
  }
  double normal() {
    double r = 2.0, v1 = 0.0, v2 = 0.0;
    while (r > 1.0) {
      v1 = 2.0 * uniform() - 1.0;
      v2 = 2.0 * uniform() - 1.0;
      r = v1 * v1 + v2 * v2;
    }
    return v1 * std::sqrt(-2.0 * std::log(r) / r);
  }
};

// upper incomplete gamma function Gamma(a, x) = Q(a, x) Gamma(a) (gamma_incomplete_upper, src/libAtoms/gamma_functions.f95:43-63:
// series for x < a + 1, continued fraction otherwise)
double gamma_incomplete_upper(double a, double x) {
  if (x < 0.0 || a <= 0.0) throw GapError("bad arguments in gamma_incomplete_upper");
  const double gln = std::lgamma(a);
  double q;
  if (x < a + 1.0) {
    double p = 0.0;
    if (x > 0.0) {
      double ap = a, sum = 1.0 / a, del = sum;
      for (int i = 0; i < 100000; i++) {
        ap += 1.0;
        del *= x / ap;
        sum += del;
        if (std::fabs(del) < std::fabs(sum) * 1e-17) break;
      }
      p = sum * std::exp(-x + a * std::log(x) - gln);
    }
    q = 1.0 - p;
  } else {
    const double tiny = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
    for (int i = 1; i < 100000; i++) {
      const double an = -i * (i - a);
      b += 2.0;
      d = an * d + b;
      if (std::fabs(d) < tiny) d = tiny;
      c = b + an / c;
      if (std::fabs(c) < tiny) c = tiny;
      d = 1.0 / d;
      const double del = d * c;
      h *= del;
      if (std::fabs(del - 1.0) < 1e-16) break;
    }
    q = std::exp(-x + a * std::log(x) - gln) * h;
  }
  return q * std::exp(gln);
}

// least-squares solution operator of the m x n (m >= n, full rank) column-major matrix A by Householder QR (LA_Matrix_QR_Factorise +
// Matrix_QR_Solve, linearalgebra.f95): returns Pinv[n][m] with x = Pinv b
std::vector<double> pinv_qr(std::vector<double> A, int m, int n) {
  std::vector<double> Q((size_t)m * m, 0.0);  // accumulate Q^T as a dense m x m matrix (m <= 48)
  for (int i = 0; i < m; i++) Q[i + (size_t)m * i] = 1.0;
  for (int k = 0; k < n; k++) {
    double nrm = 0.0;
    for (int i = k; i < m; i++) nrm += A[i + (size_t)m * k] * A[i + (size_t)m * k];
    nrm = std::sqrt(nrm);
    if (nrm == 0.0) throw GapError("LA_Matrix_QR_Factorise: rank-deficient radial basis");
    const double alpha = A[k + (size_t)m * k] > 0 ? -nrm : nrm;
    std::vector<double> v(m, 0.0);
    for (int i = k; i < m; i++) v[i] = A[i + (size_t)m * k];
    v[k] -= alpha;
    double vn = 0.0;
    for (int i = k; i < m; i++) vn += v[i] * v[i];
    if (vn == 0.0) continue;
    for (int j = 0; j < n; j++) {  // A <- H A
      double t = 0.0;
      for (int i = k; i < m; i++) t += v[i] * A[i + (size_t)m * j];
      t *= 2.0 / vn;
      for (int i = k; i < m; i++) A[i + (size_t)m * j] -= t * v[i];
    }
    for (int j = 0; j < m; j++) {  // Qt <- H Qt
      double t = 0.0;
      for (int i = k; i < m; i++) t += v[i] * Q[i + (size_t)m * j];
      t *= 2.0 / vn;
      for (int i = k; i < m; i++) Q[i + (size_t)m * j] -= t * v[i];
    }
  }
  // x = R^-1 (Q^T b)[0:n]  ->  Pinv = R^-1 Qt[0:n, :]
  std::vector<double> Pinv((size_t)n * m, 0.0);
  for (int col = 0; col < m; col++)
    for (int i = n - 1; i >= 0; i--) {
      double t = Q[i + (size_t)m * col];
      for (int k = i + 1; k < n; k++) t -= A[i + (size_t)m * k] * Pinv[(size_t)k * m + col];
      Pinv[(size_t)i * m + col] = t / A[i + (size_t)m * i];
    }
  return Pinv;
}

void soap_general_setup(SoapSpec& s, bool diagonal_radial, bool Z_mix, bool R_mix, bool sym_mix, bool coupling, int nu_R, int nu_S, int K, int mix_shift,
                        const std::string& Z_map, double cutoff_basis) {
  const int n = s.n_max, ns = s.n_species, L = s.l_max, K1 = n * ns;
  const bool mixing = R_mix || Z_mix || sym_mix, using_Zmap = !Z_map.empty();
  // form_W :7618-7670
  if ((nu_R != 2 || nu_S != 2) && mixing) throw GapError("(nu_R, nu_S) = (2,2) required to use channel mixing");
  if ((nu_R != 2 || nu_S != 2) && diagonal_radial) throw GapError("(nu_R, nu_S) = (2,2) required to use diagonal radial");
  if ((nu_R != 2 || nu_S != 2) && using_Zmap) throw GapError("(nu_R, nu_S) = (2,2) required to use Zmap");
  if (mixing && using_Zmap) throw GapError("cant' using mixing and Zmap at the same time");
  bool sym_desc;
  std::vector<double> W[2];
  int Kw[2] = {0, 0};
  if (mixing) {  // form_mix_W :7430-7533
    sym_desc = sym_mix;
    if (K < 1) throw GapError("form_mix_W: K (number of mixing channels) must be positive");
    for (int j = 1; j <= 2; j++) {
      std::vector<double>& w = W[j - 1];
      if (sym_desc && j == 2) { w = W[0]; Kw[1] = Kw[0]; continue; }
      if (R_mix && Z_mix) {
        Kw[j - 1] = K;
        w.assign((size_t)K1 * K, 0.0);
        for (int is = 0; is < ns; is++) {
          QuipRng rng(s.species_Z[is] + mix_shift + j * 200);
          for (int r = is * n; r < (is + 1) * n; r++)
            for (int c = 0; c < K; c++) w[(size_t)r * K + c] = rng.normal();
        }
      } else if (Z_mix) {
        Kw[j - 1] = K * n;
        w.assign((size_t)K1 * K * n, 0.0);
        for (int is = 0; is < ns; is++) {
          QuipRng rng(s.species_Z[is] + mix_shift + j * 200);
          for (int k = 0; k < K; k++) {
            const double rv = rng.normal();
            for (int a = 0; a < n; a++) w[(size_t)(is * n + a) * (K * n) + k * n + a] = rv;
          }
        }
      } else if (R_mix) {
        Kw[j - 1] = K * ns;
        w.assign((size_t)K1 * K * ns, 0.0);
        QuipRng rng(n + mix_shift + j * 200);
        std::vector<double> R((size_t)n * K);
        for (int r = 0; r < n; r++)
          for (int c = 0; c < K; c++) R[(size_t)r * K + c] = rng.normal();
        for (int is = 0; is < ns; is++)
          for (int a = 0; a < n; a++)
            for (int k = 0; k < K; k++) w[(size_t)(is * n + a) * (K * ns) + is * K + k] = R[(size_t)a * K + k];
      } else {
        throw GapError("form_mix_W: not mixing anything");
      }
    }
  } else if (using_Zmap) {  // form_Zmap_W :7536-7616
    int n_groups[2] = {1, 1}, dens = 0;
    for (char ch : Z_map) {
      if (ch == ',') n_groups[dens < 2 ? dens : 1]++;
      if (ch == ':') dens++;
    }
    if (dens > 1) throw GapError("form_Zmap_W: at most one ':' in Z_map");
    sym_desc = dens == 0;
    Kw[0] = n * n_groups[0];
    Kw[1] = n * (dens == 1 ? n_groups[1] : n_groups[0]);
    W[0].assign((size_t)K1 * Kw[0], 0.0);
    W[1].assign((size_t)K1 * Kw[1], 0.0);
    int i_group = 0, i_density = 0;
    std::string tok;
    for (size_t q = 0; q <= Z_map.size(); q++) {
      const char ch = q < Z_map.size() ? Z_map[q] : ' ';
      if (isdigit((unsigned char)ch)) { tok.push_back(ch); continue; }
      if (!tok.empty()) {
        const int Zv = atoi(tok.c_str());
        int isp = -1;
        for (int k = 0; k < ns; k++)
          if (s.species_Z[k] == Zv) isp = k;
        if (isp < 0) throw GapError("form_Zmap_W: Z_map names Z=" + tok + " which is not in species_Z");
        for (int a = 0; a < n; a++) W[i_density][(size_t)(isp * n + a) * Kw[i_density] + i_group * n + a] = 1.0;
        tok.clear();
      }
      if (ch == ',') i_group++;
      if (ch == ':') { i_density++; i_group = 0; }
    }
    if (sym_desc) W[1] = W[0];
  } else {  // form_nu_W :7352-7424
    if (nu_R < 0 || nu_R > 2) throw GapError("nu_R outside allowed range of 0-2");
    if (nu_S < 0 || nu_S > 2) throw GapError("nu_S outside allowed range of 0-2");
    sym_desc = !(nu_R == 1 || nu_S == 1);
    int r = nu_R, sp = nu_S;
    for (int i = 0; i < 2; i++) {
      int dn = 0, ds = 0, n2_max = 1, s2_max = 1;
      if (r > 0) { r--; dn = 1; n2_max = n; }
      if (sp > 0) { sp--; ds = 1; s2_max = ns; }
      Kw[i] = n2_max * s2_max;
      W[i].assign((size_t)K1 * Kw[i], 0.0);
      for (int s1 = 1; s1 <= ns; s1++)
        for (int a = 1; a <= n; a++) {
          int ic = 0;
          for (int s2 = 1; s2 <= s2_max; s2++)
            for (int n2 = 1; n2 <= n2_max; n2++, ic++)
              if (ds * s1 == ds * s2 && dn * a == dn * n2) W[i][(size_t)((s1 - 1) * n + a - 1) * Kw[i] + ic] = 1.0;
        }
    }
  }
  s.Ka = Kw[0]; s.Kb = Kw[1];
  s.W1 = W[0]; s.W2 = W[1];
  // the power-spectrum elements, in the order the unpacking loops write them (:8402-8416; form_coupling_inds :7274-7350)
  const bool original = coupling && nu_R == 2 && nu_S == 2 && !mixing && !using_Zmap;
  const double sqrt2 = std::sqrt(2.0);
  s.pair_ia.clear(); s.pair_jb.clear(); s.pair_fac.clear();
  auto push = [&](int ia, int jb, double f) { s.pair_ia.push_back(ia); s.pair_jb.push_back(jb); s.pair_fac.push_back(f); };
  if (coupling) {
    if (diagonal_radial && !original) throw GapError("soap_dimensions: can't combine diagonal radial with any other compression strategies");
    for (int ia = 0; ia < s.Ka; ia++)
      for (int jb = 0; jb < (sym_desc ? ia + 1 : s.Kb); jb++) {
        if (diagonal_radial && (ia % n) != (jb % n)) continue;
        push(ia, jb, (sym_desc && ia != jb) ? sqrt2 : 1.0);
      }
  } else {
    if (s.Ka != s.Kb) throw GapError("require K1=K2 to use elementwise coupling");
    const double sqrt2_f32 = (double)(float)sqrt2;  // sym_facs is declared `real` (single precision, :7283)
    if (Z_mix && !R_mix) {
      for (int k = 0; k < K; k++)
        for (int a = 0; a < n; a++)
          for (int b = 0; b < (sym_mix ? a + 1 : n); b++) push(k * n + a, k * n + b, (a != b && sym_mix) ? sqrt2_f32 : 1.0);
    } else if (R_mix && !Z_mix) {
      for (int is = 0; is < ns; is++)
        for (int js = 0; js < (sym_mix ? is + 1 : ns); js++)
          for (int k = 0; k < K; k++) push(is * K + k, js * K + k, (is != js && sym_mix) ? sqrt2_f32 : 1.0);
    } else {
      for (int i = 0; i < s.Ka; i++) push(i, i, 1.0);
    }
  }
  s.d = (L + 1) * (int)s.pair_ia.size() + 1;  // soap_dimensions :10929-10967

  // radial map per l
  if (s.radial_basis == "EQUISPACED_GAUSS") {
    s.n_grid = n;
    s.r_grid = s.r_basis;
    s.P.assign((size_t)(L + 1) * n * n, 0.0);
    for (int l = 0; l <= L; l++)
      for (int g = 0; g < n; g++)
        for (int a = 0; a < n; a++) s.P[((size_t)l * n + g) * n + a] = s.transform_basis[g + n * a];
    s.c0.assign(n, 0.0);
    for (int a = 0; a < n; a++) s.c0[a] = s.cholesky_overlap[0 + n * a];  // radial_fun(0,:) = e_1 times the Cholesky factor :8151-8154
    return;
  }
  const int ng = 3 * n;  // :2644-2651
  s.n_grid = ng;
  s.r_grid.assign(ng, 0.0);
  for (int i = 1; i < ng; i++) s.r_grid[i] = s.r_grid[i - 1] + cutoff_basis / ng;
  s.P.assign((size_t)(L + 1) * ng * n, 0.0);
  const int l_ub = s.radial_basis == "POLY" ? 0 : L;
  for (int l = 0; l <= l_ub; l++) {
    std::vector<double> S((size_t)n * n), B((size_t)ng * n);  // overlap (n x n) and the basis functions on the grid (ng x n), column-major
    if (s.radial_basis == "POLY") {  // :2661-2678
      auto Nn = [&](int i) { return std::sqrt(std::pow(cutoff_basis, 2 * i + 7) / ((i + 3.0) * (2 * i + 5.0) * (2 * i + 7.0))); };
      for (int i = 1; i <= n; i++)
        for (int j = 1; j <= n; j++)
          S[(j - 1) + (size_t)n * (i - 1)] = 2.0 * std::pow(cutoff_basis, i + j + 7) / ((5.0 + i + j) * (6.0 + i + j) * (7.0 + i + j)) / (Nn(i) * Nn(j));
      for (int i = 1; i <= n; i++)
        for (int g = 0; g < ng; g++) B[g + (size_t)ng * (i - 1)] = std::pow(cutoff_basis - s.r_grid[g], i + 2) / Nn(i);
    } else {  // GTO :2680-2712
      std::vector<double> aln(n);
      for (int k = 1; k <= n; k++) {
        const double Rg = (cutoff_basis / n) * k;
        aln[k - 1] = -std::pow(Rg, -2.0) * (std::log(0.001) - l * std::log(Rg));
      }
      const double t = l + 1.5;
      for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
          const double u = (aln[i] + aln[j]) * cutoff_basis * cutoff_basis;
          S[i + (size_t)n * j] = 0.5 * std::pow(cutoff_basis, 2.0 * t) * std::pow(u, -t) * (std::tgamma(t) - gamma_incomplete_upper(t, u));
        }
      for (int i = 0; i < n; i++)
        for (int g = 0; g < ng; g++) B[g + (size_t)ng * i] = std::pow(s.r_grid[g], l) * std::exp(-aln[i] * s.r_grid[g] * s.r_grid[g]);
    }
    chol_lower(S, n, "overlap_basis");  // :2724-2731 (lower factor Lc, S = Lc Lc^T)
    // A = B (Lc^T)^-1 :2733-2739 : solve A Lc^T = B column by column (Lc^T upper triangular)
    std::vector<double> A((size_t)ng * n, 0.0);
    for (int g = 0; g < ng; g++)
      for (int j = 0; j < n; j++) {
        double v = B[g + (size_t)ng * j];
        for (int k = 0; k < j; k++) v -= A[g + (size_t)ng * k] * S[j + (size_t)n * k];  // (Lc^T)(k, j) = Lc(j, k)
        A[g + (size_t)ng * j] = v / S[j + (size_t)n * j];
      }
    const std::vector<double> Pinv = pinv_qr(A, ng, n);  // [a][g]
    for (int g = 0; g < ng; g++)
      for (int a = 0; a < n; a++) s.P[((size_t)l * ng + g) * n + a] = Pinv[(size_t)a * ng + g];
  }
  for (int l = l_ub + 1; l <= L; l++)  // POLY: the l = 0 factors serve every l (:2751-2756)
    for (size_t k = 0; k < (size_t)ng * n; k++) s.P[(size_t)l * ng * n + k] = s.P[k];
  s.c0.assign(n, 0.0);  // central atom: radial_fun(0, g) = exp(-alpha r_g^2) through the l = 0 map (:8156-8165)
  for (int g = 0; g < ng; g++) {
    const double rf = std::exp(-s.alpha * s.r_grid[g] * s.r_grid[g]);
    for (int a = 0; a < n; a++) s.c0[a] += rf * s.P[(size_t)g * n + a];
  }
}
}  // namespace

SoapSpec soap_from_string(const std::string& desc, long calc_xml_version) {
  ArgDict a(desc);
  SoapSpec s;
  auto need = [&](const char* k) {
    if (!a.has(k)) throw GapError(std::string("soap_initialise: missing mandatory parameter ") + k + " in '" + desc + "'");
  };
  need("cutoff"); need("l_max"); need("n_max");
  s.cutoff = a.real("cutoff", 0);
  s.cutoff_transition_width = a.real("cutoff_transition_width", 0.5);
  s.cutoff_dexp = (int)a.integer("cutoff_dexp", 0);
  s.cutoff_scale = a.real("cutoff_scale", 1.0);
  s.cutoff_rate = a.real("cutoff_rate", 1.0);
  s.l_max = (int)a.integer("l_max", 0);
  s.n_max = (int)a.integer("n_max", 0);
  if (a.has("atom_gaussian_width")) s.atom_sigma = a.real("atom_gaussian_width", 0);
  else if (a.has("atom_sigma")) s.atom_sigma = a.real("atom_sigma", 0);
  else throw GapError("soap_initialise: missing mandatory parameter atom_gaussian_width/atom_sigma");
  s.central_weight = a.real("central_weight", 1.0);
  bool has_cras = a.has("central_reference_all_species");
  s.central_reference_all_species = a.logical("central_reference_all_species", false);
  s.covariance_sigma0 = a.real("covariance_sigma0", 0.0);
  s.normalise = a.has("normalise") ? a.logical("normalise", true) : a.logical("normalize", true);
  double basis_error_exponent = a.real("basis_error_exponent", 10.0);
  s.n_Z = (int)a.integer("n_Z", 1);
  bool has_n_species = a.has("n_species");
  s.n_species = (int)a.integer("n_species", 1);
  long xml_version = a.integer("xml_version", 1426512068L);
  s.global = a.logical("average", false);  // ONE descriptor per configuration (global SOAP): the general path's element list on the summed X
  const bool diagonal_radial = a.logical("diagonal_radial", false), Z_mix = a.logical("Z_mix", false), R_mix = a.logical("R_mix", false),
             sym_mix = a.logical("sym_mix", false), coupling = a.logical("coupling", true);
  const int nu_R = (int)a.integer("nu_R", 2), nu_S = (int)a.integer("nu_S", 2), K_mix = (int)a.integer("K", 0), mix_shift = (int)a.integer("mix_shift", 0);
  const std::string Z_map = a.str("Z_map", "");
  s.radial_basis = a.str("radial_basis", "");
  if (s.radial_basis.empty()) s.radial_basis = "EQUISPACED_GAUSS";  // :2548-2551
  if (s.radial_basis != "EQUISPACED_GAUSS" && s.radial_basis != "GTO" && s.radial_basis != "POLY")
    throw GapError("soap_initialise: radial_basis not recognised: EQUISPACED_GAUSS, POLY or GTO");
  s.general = s.global || diagonal_radial || Z_mix || R_mix || sym_mix || !coupling || nu_R != 2 || nu_S != 2 || !Z_map.empty() || s.radial_basis != "EQUISPACED_GAUSS";
  if (s.cutoff_dexp < 0) throw GapError("soap_initialise: cutoff_dexp may not be less than 0");
  if (s.cutoff_scale <= 0.0) throw GapError("soap_initialise: cutoff_scale must be greater than 0");
  if (s.cutoff_rate < 0.0) throw GapError("soap_initialise: cutoff_rate may not be less than 0");
  if (s.n_max < 1 || s.l_max < 0 || s.n_species < 1 || s.n_Z < 1) throw GapError("soap_initialise: bad n_max/l_max/n_species/n_Z");

  if (xml_version < 1426512068L) s.central_reference_all_species = true;  // :2545
  bool has_species_Z = a.has("species_Z") && !a.str("species_Z", "").empty();
  if (has_species_Z && !has_n_species) throw GapError("soap_initialise: is species_Z is present, n_species must be present, too.");
  s.species_Z = a.int_list("species_Z");
  if (s.n_species == 1) {
    if (s.species_Z.empty()) s.species_Z.push_back(0);
    s.species_Z.resize(1);
  } else if ((int)s.species_Z.size() != s.n_species) {
    throw GapError("soap_initialise: species_Z must list n_species atomic numbers");
  }
  if (!has_cras && s.n_species == 1) s.central_reference_all_species = true;  // :2585
  s.Z = a.int_list("Z");
  if (s.n_Z == 1) {
    if (s.Z.empty()) s.Z.push_back(0);
    s.Z.resize(1);
  } else if ((int)s.Z.size() != s.n_Z) {
    throw GapError("soap_initialise: Z must list n_Z atomic numbers");
  }
  s.do_two_l_plus_one = (calc_xml_version < 0 ? 1423143769L : calc_xml_version) >= 1423143769L;  // :7800,:7837

  const int n = s.n_max;
  s.alpha = 0.5 / (s.atom_sigma * s.atom_sigma);
  const double al = s.alpha;
  double cutoff_basis = s.cutoff + s.atom_sigma * std::sqrt(2.0 * basis_error_exponent * std::log(10.0));
  double spacing = cutoff_basis / n;
  s.r_basis.assign(n, 0.0);
  for (int i = 1; i < n; i++) s.r_basis[i] = s.r_basis[i - 1] + spacing;
  std::vector<double> cov(n * n), ovl(n * n);
  const double pi = 3.14159265358979323846264338327950288;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      double ri = s.r_basis[i], rj = s.r_basis[j];
      cov[j + n * i] = std::exp(-al * (ri - rj) * (ri - rj));
      ovl[j + n * i] = (std::exp(-al * (ri * ri + rj * rj)) * std::sqrt(2.0) * std::pow(al, 1.5) * (ri + rj) +
                        al * std::exp(-0.5 * al * (ri - rj) * (ri - rj)) * std::sqrt(pi) * (1.0 + al * (ri + rj) * (ri + rj)) *
                            (1.0 + std::erf(std::sqrt(al / 2.0) * (ri + rj)))) /
                       std::sqrt(128.0 * std::pow(al, 5));
    }
  chol_lower(ovl, n, "overlap_basis");
  for (int i = 0; i < n; i++)
    for (int j = 0; j < i; j++) ovl[j + n * i] = 0.0;
  s.cholesky_overlap = ovl;
  chol_lower(cov, n, "covariance_basis");
  s.transform_basis.assign(n * n, 0.0);
  for (int col = 0; col < n; col++) {
    double* x = &s.transform_basis[n * col];
    for (int i = 0; i < n; i++) {
      double t = ovl[i + n * col];
      for (int k = 0; k < i; k++) t -= cov[i + n * k] * x[k];
      x[i] = t / cov[i + n * i];
    }
    for (int i = n - 1; i >= 0; i--) {
      double t = x[i];
      for (int k = i + 1; k < n; k++) t -= cov[k + n * i] * x[k];
      x[i] = t / cov[i + n * i];
    }
  }
  int K1 = s.K1();
  s.d = (s.l_max + 1) * K1 * (K1 + 1) / 2 + 1;  // soap_dimensions :10953-10958
  if (s.general) soap_general_setup(s, diagonal_radial, Z_mix, R_mix, sym_mix, coupling, nu_R, nu_S, K_mix, mix_shift, Z_map, cutoff_basis);
  return s;
}

Distance2bSpec distance_2b_from_string(const std::string& desc) {
  ArgDict a(desc);
  Distance2bSpec s;
  s.cutoff = a.real("cutoff", 0.0);
  s.cutoff_transition_width = a.real("cutoff_transition_width", 0.5);
  s.Z1 = (int)a.integer("Z1", 0);
  s.Z2 = (int)a.integer("Z2", 0);
  const bool intra = a.logical("only_intra", false), inter = a.logical("only_inter", false);
  if (intra && inter) throw GapError("distance_2b_initialise: cannot specify both only_inter AND only_intra");
  if ((intra || inter) && !a.has("resid_name")) throw GapError("distance_2b_initialise: only_intra and only_inter require resid_name to be given as well");
  s.intra_mode = intra ? 1 : (inter ? 2 : 0);
  s.resid_name = a.str("resid_name", "");
  const int n_exp = (int)a.integer("n_exponents", 1);
  if (n_exp < 1 || n_exp > 4) throw GapError("distance_2b: n_exponents must be 1..4 on the B200 path");
  if (a.has("exponents")) {
    std::string t = a.str("exponents", "");
    for (char& ch : t)
      if (ch == ',') ch = ' ';
    s.exponents = parse_reals(t);
    if ((int)s.exponents.size() != n_exp) throw GapError("distance_2b_initialise: exponents must list n_exponents values");
  } else if (n_exp == 1) {
    s.exponents.assign(1, 1.0);
  } else {  // :1808-1812
    for (int i = 1; i <= n_exp; i++) s.exponents.push_back(-(double)i);
  }
  s.tail_exponent = a.has("tail_exponent") ? (int)a.integer("tail_exponent", 0) : 0;
  s.tail_range = a.real("tail_range", 1.0);
  return s;
}

Angle3bSpec angle_3b_from_string(const std::string& desc) {
  ArgDict a(desc);
  Angle3bSpec s;
  s.cutoff = a.real("cutoff", 0.0);
  s.cutoff_transition_width = a.real("cutoff_transition_width", 0.5);
  s.Zc = (int)a.integer("Z_center", a.integer("Z", 0));
  s.Z1 = (int)a.integer("Z1", 0);
  s.Z2 = (int)a.integer("Z2", 0);
  return s;
}

// ======================================================================================
// minimal SAX-style XML scanner (stands in for FoX, src/fox)
// ======================================================================================
namespace {
struct XmlAttrs {
  std::vector<std::pair<std::string, std::string>> a;
  bool get(const std::string& k, std::string& v) const {
    for (auto& p : a)
      if (p.first == k) { v = p.second; return true; }
    return false;
  }
};
std::string decode_entities(const std::string& s) {
  if (s.find('&') == std::string::npos) return s;
  std::string o;
  for (size_t i = 0; i < s.size(); i++) {
    if (s[i] != '&') { o.push_back(s[i]); continue; }
    size_t e = s.find(';', i);
    if (e == std::string::npos) { o.push_back('&'); continue; }
    std::string ent = s.substr(i + 1, e - i - 1);
    if (ent == "lt") o.push_back('<'); else if (ent == "gt") o.push_back('>'); else if (ent == "amp") o.push_back('&');
    else if (ent == "quot") o.push_back('"'); else if (ent == "apos") o.push_back('\'');
    else if (!ent.empty() && ent[0] == '#') o.push_back((char)(ent.size() > 1 && ent[1] == 'x' ? strtol(ent.c_str() + 2, 0, 16) : strtol(ent.c_str() + 1, 0, 10)));
    else o += "&" + ent + ";";
    i = e;
  }
  return o;
}
struct XmlHandler {
  std::function<void(const std::string&, const XmlAttrs&)> start;
  std::function<void(const std::string&)> end;
  std::function<void(const std::string&)> chars;
};
void xml_scan(const std::string& s, const XmlHandler& h) {
  size_t p = 0, n = s.size();
  while (p < n) {
    size_t lt = s.find('<', p);
    if (lt == std::string::npos) lt = n;
    if (lt > p && h.chars) h.chars(decode_entities(s.substr(p, lt - p)));
    if (lt >= n) break;
    if (s.compare(lt, 4, "<!--") == 0) {
      size_t e = s.find("-->", lt + 4);
      if (e == std::string::npos) throw GapError("XML: unterminated comment");
      p = e + 3;
    } else if (s.compare(lt, 9, "<![CDATA[") == 0) {
      size_t e = s.find("]]>", lt + 9);
      if (e == std::string::npos) throw GapError("XML: unterminated CDATA");
      if (h.chars) h.chars(s.substr(lt + 9, e - lt - 9));
      p = e + 3;
    } else if (s.compare(lt, 2, "<?") == 0) {
      size_t e = s.find("?>", lt + 2);
      if (e == std::string::npos) throw GapError("XML: unterminated processing instruction");
      p = e + 2;
    } else if (s.compare(lt, 2, "<!") == 0) {
      size_t e = s.find('>', lt + 2);
      if (e == std::string::npos) throw GapError("XML: unterminated declaration");
      p = e + 1;
    } else if (s.compare(lt, 2, "</") == 0) {
      size_t e = s.find('>', lt + 2);
      if (e == std::string::npos) throw GapError("XML: unterminated end tag");
      std::string name = s.substr(lt + 2, e - lt - 2);
      while (!name.empty() && isspace((unsigned char)name.back())) name.pop_back();
      if (h.end) h.end(name);
      p = e + 1;
    } else {
      size_t q = lt + 1;
      while (q < n && !isspace((unsigned char)s[q]) && s[q] != '>' && s[q] != '/') q++;
      std::string name = s.substr(lt + 1, q - lt - 1);
      XmlAttrs attrs;
      bool selfclose = false;
      while (q < n) {
        while (q < n && isspace((unsigned char)s[q])) q++;
        if (q >= n) throw GapError("XML: unterminated start tag <" + name);
        if (s[q] == '>') { q++; break; }
        if (s[q] == '/') { selfclose = true; q++; continue; }
        size_t k0 = q;
        while (q < n && s[q] != '=' && !isspace((unsigned char)s[q]) && s[q] != '>') q++;
        std::string key = s.substr(k0, q - k0);
        while (q < n && isspace((unsigned char)s[q])) q++;
        if (q >= n || s[q] != '=') throw GapError("XML: attribute without value in <" + name);
        q++;
        while (q < n && isspace((unsigned char)s[q])) q++;
        if (q >= n || (s[q] != '"' && s[q] != '\'')) throw GapError("XML: unquoted attribute value in <" + name);
        char qc = s[q++];
        size_t v0 = q;
        while (q < n && s[q] != qc) q++;
        if (q >= n) throw GapError("XML: unterminated attribute value in <" + name);
        attrs.a.emplace_back(key, decode_entities(s.substr(v0, q - v0)));
        q++;
      }
      if (h.start) h.start(name, attrs);
      if (selfclose && h.end) h.end(name);
      p = q;
    }
  }
}
std::string read_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw GapError("cannot open file " + path);
  std::ostringstream ss;
  ss << f.rdbuf();
  return ss.str();
}
std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) a++;
  while (b > a && isspace((unsigned char)s[b - 1])) b--;
  return s.substr(a, b - a);
}
}  // namespace

// ======================================================================================
// model loader
// ======================================================================================
GapModel load_gap_model(const std::string& args_str_in, const std::string& param_str, const std::string& base_dir) {
  if (trim(param_str).empty()) throw GapError("IPModel_GAP_read_params_xml: invalid param_str length 0");
  std::string args_str = args_str_in;
  // Potential_initialise: with empty args_str take init_args of the first <Potential> element (Potential.f95:526-552)
  if (trim(args_str).empty()) {
    bool found = false;
    XmlHandler hp;
    hp.start = [&](const std::string& name, const XmlAttrs& at) {
      if (name == "Potential" && !found) {
        std::string v;
        if (at.get("init_args", v)) { args_str = v; found = true; }
      }
    };
    xml_scan(param_str, hp);
    if (!found) throw GapError("Potential_initialise: no args_str given and no <Potential init_args=...> in the XML");
  }
  ArgDict args(args_str);
  if (args.has("Sum") || args.has("ForceMixing") || args.has("EVB") || args.has("ONIOM") || args.has("Cluster"))
    throw GapError("Potential_initialise: only simple 'IP GAP' potentials are supported by the B200 path, got '" + args_str + "'");
  if (!(args.has("IP") && args.has("GAP")))
    throw GapError("Potential_initialise: init args must be 'IP GAP [label=...]', got '" + args_str + "'");
  // Potential-level rescaling (Potential.f95:568-746: r_scale / E_scale wrapped around calc, target_vol / target_B fitted at
  // initialise) is not part of the GAP path: refuse it rather than return unscaled results
  for (const char* k : {"do_rescale_r", "do_rescale_E", "r_scale", "target_vol", "target_B", "minimise_bulk"})
    if (args.has(k))
      throw GapError(std::string("Potential_initialise: ") + k + " (rescaling of the potential, Potential.f95:568-746) is not supported by the B200 path");

  GapModel m;
  for (double& v : m.e0) v = 0.0;
  m.label = args.str("label", "");
  m.E_scale = args.real("E_scale", 1.0);

  // ---- pass 1: <GAP_params>/<GAP_data>/<e0>  (IPModel_GAP.f95:618-700) ----
  {
    bool in_ip = false, matched = false, done = false, in_gap_data = false;
    long version = 0;
    bool any_params = false;
    XmlHandler h;
    h.start = [&](const std::string& name, const XmlAttrs& at) {
      std::string v;
      if (name == "GAP_params") {
        any_params = true;
        if (matched) return;
        std::string lab;
        if (!at.get("label", lab)) lab = "";
        if (!m.label.empty()) {
          if (lab == m.label) { matched = true; in_ip = true; } else in_ip = false;
        } else {
          in_ip = true;
          m.label = lab;
        }
        if (in_ip) {
          version = at.get("gap_version", v) ? strtol(v.c_str(), 0, 10) : 0;
          for (double& e : m.e0) e = 0.0;
        }
      } else if (in_ip && name == "GAP_data") {
        if (at.get("e0", v)) {
          double e = parse_real(v);
          for (double& x : m.e0) x = e;
        }
        if (at.get("do_pca", v) && (v.find('T') != std::string::npos || v.find('t') != std::string::npos))
          throw GapError("IPModel_GAP: do_pca=T is not supported by the B200 path");
        in_gap_data = true;
      } else if (in_ip && in_gap_data && name == "e0") {
        if (!at.get("Z", v)) throw GapError("IPModel_GAP_read_params_xml cannot find Z");
        long Z = strtol(v.c_str(), 0, 10);
        if (Z < 1 || Z > 116) throw GapError("IPModel_GAP_read_params_xml: attribute Z = " + v + " > 116");
        if (!at.get("value", v)) throw GapError("IPModel_GAP_read_params_xml cannot find value in e0");
        m.e0[Z] = parse_real(v);
      }
    };
    h.end = [&](const std::string& name) {
      if (!in_ip) return;
      if (name == "GAP_params") { in_ip = false; done = true; }
      else if (name == "GAP_data") in_gap_data = false;
    };
    xml_scan(param_str, h);
    if (!done) throw GapError("IPModel_GAP_read_params_xml: could not initialise GAP potential. No GAP_params present?");
    m.xml_version = version;
    (void)any_params;
  }

  // ---- pass 2: <gpSparse> (gp_predict.f95:5200-5250) ----
  int n_coordinate = -1;
  std::string gp_label;
  {
    bool matched = false, in_gp = false;
    bool fitted = true;
    XmlHandler h;
    h.start = [&](const std::string& name, const XmlAttrs& at) {
      if (name != "gpSparse" || matched) return;
      std::string lab, v;
      if (!at.get("label", lab)) lab = "";
      if (!m.label.empty()) {
        if (lab == m.label) { matched = true; in_gp = true; } else { in_gp = false; return; }
      } else in_gp = true;
      if (!at.get("n_coordinate", v)) throw GapError("gpSparse_startElement_handler did not find the n_coordinate attribute.");
      n_coordinate = (int)strtol(v.c_str(), 0, 10);
      gp_label = lab;
      fitted = at.get("fitted", v) ? (v.find('T') != std::string::npos || v.find('t') != std::string::npos) : true;
    };
    xml_scan(param_str, h);
    (void)in_gp;
    if (n_coordinate < 0) throw GapError("gp_readXML: no gpSparse element with label '" + m.label + "'");
    if (!fitted) throw GapError("IPModel_GAP_Initialise_str: GAP model has not been fitted.");
  }

  // ---- pass 3: each <gpCoordinates label=<label>//i> (gp_predict.f95:4596-5047) ----
  for (int ic = 1; ic <= n_coordinate; ic++) {
    Coordinate c;
    c.label = gp_label + std::to_string(ic);
    bool in_c = false, matched = false, found = false, separate_file = false, sliced = false, in_sparseX = false;
    bool has_zeta = false;
    int i_sparseX = 0, slice_start = 0, slice_end = 0;
    std::string cur;
    XmlHandler h;
    h.start = [&](const std::string& name, const XmlAttrs& at) {
      std::string v;
      if (name == "gpCoordinates") {
        if (matched) return;
        std::string lab;
        if (!at.get("label", lab)) lab = "";
        if (trim(lab) != c.label) { in_c = false; return; }
        matched = in_c = found = true;
        auto req = [&](const char* k) {
          if (!at.get(k, v)) throw GapError(std::string("gpCoordinates_startElement_handler did not find the ") + k + " attribute.");
          return v;
        };
        c.d = (int)strtol(req("dimensions").c_str(), 0, 10);
        c.delta = parse_real(req("signal_variance"));
        c.f0 = parse_real(req("signal_mean"));
        std::string sp = req("sparsified");
        if (sp.find('T') == std::string::npos && sp.find('t') == std::string::npos)
          throw GapError("gpCoordinates: sparsified=F models cannot be used for prediction");
        c.n_permutations = (int)strtol(req("n_permutations").c_str(), 0, 10);
        c.covariance_type = (int)strtol(req("covariance_type").c_str(), 0, 10);
        if (at.get("zeta", v)) {
          if (c.covariance_type != COVARIANCE_DOT_PRODUCT)
            throw GapError("gpCoordinates_startElement_handler found zeta attribute but the covariance is not dot product.");
          c.zeta = parse_real(v);
          has_zeta = true;
        }
        c.M = (int)strtol(req("n_sparseX").c_str(), 0, 10);
        if (c.d < 1 || c.M < 0) throw GapError("gpCoordinates: bad dimensions/n_sparseX");
        c.sparseX.assign((size_t)c.d * c.M, 0.0);
        c.alpha.assign(c.M, 0.0);
        c.sparseCutoff.assign(c.M, 0.0);
        c.theta.assign(c.d, 0.0);
        if (at.get("sparseX_filename", v)) {
          std::string path = (!v.empty() && v[0] == '/') ? v : base_dir + "/" + v;
          std::string bytes;
          try { bytes = read_file(path); } catch (GapError&) {
            throw GapError("gpCoordinates_startElement_handler: sparseX file " + v + " does not exist.");
          }
          std::string md5;
          if (at.get("sparseX_md5sum", md5) && trim(md5).size() == 32 && trim(md5) != md5_hex(bytes))
            throw GapError("gpCoordinates_startElement_handler: md5 check sum failed. Sparse file (" + md5_hex(bytes) +
                           ") does not match record in XML (" + trim(md5) + ")");
          // fread_array_d_: one "%lf" per entry, column-major (cutil.c:205-214)
          const char* p = bytes.c_str();
          for (size_t k = 0; k < c.sparseX.size(); k++) {
            char* e = nullptr;
            c.sparseX[k] = strtod(p, &e);
            if (e == p) throw GapError("fread_array_d: sparseX file " + v + " holds fewer than dimensions*n_sparseX numbers");
            p = e;
          }
          separate_file = true;
        }
      } else if (!in_c) {
        return;
      } else if (name == "theta" || name == "descriptor" || name == "permutation") {
        cur.clear();
      } else if (name == "sparseX") {
        in_sparseX = true;
        if (!at.get("i", v)) throw GapError("gpCoordinates_startElement_handler did not find the i attribute.");
        i_sparseX = (int)strtol(v.c_str(), 0, 10);
        if (i_sparseX < 1 || i_sparseX > c.M)
          throw GapError("gpCoordinates_endElement_handler: parse_i_sparseX (" + v + ") greater than n_sparseX (" + std::to_string(c.M) + ")");
        if (!at.get("alpha", v)) throw GapError("gpCoordinates_startElement_handler did not find the alpha attribute.");
        c.alpha[i_sparseX - 1] = parse_real(v);
        if (!at.get("sparseCutoff", v)) throw GapError("gpCoordinates_startElement_handler did not find the cutoff attribute.");
        c.sparseCutoff[i_sparseX - 1] = parse_real(v);
        sliced = at.get("sliced", v) && (v.find('T') != std::string::npos || v.find('t') != std::string::npos);
        cur.clear();
      } else if (in_sparseX && name == "sparseX_slice") {
        if (!at.get("start", v)) throw GapError("gpCoordinates_startElement_handler did not find the start attribute.");
        slice_start = (int)strtol(v.c_str(), 0, 10);
        if (!at.get("end", v)) throw GapError("gpCoordinates_startElement_handler did not find the end attribute.");
        slice_end = (int)strtol(v.c_str(), 0, 10);
        cur.clear();
      }
    };
    h.chars = [&](const std::string& t) {
      if (in_c) cur += t;
    };
    h.end = [&](const std::string& name) {
      if (!in_c) return;
      if (name == "gpCoordinates") {
        in_c = false;
      } else if (name == "theta") {
        std::vector<double> th = parse_reals(cur);
        for (size_t k = 0; k < th.size() && k < c.theta.size(); k++) c.theta[k] = th[k];
        if (c.covariance_type == COVARIANCE_DOT_PRODUCT && !th.empty()) {  // legacy files: theta holds zeta (:4972-4977)
          c.zeta = th[0];
          has_zeta = true;
        }
      } else if (name == "descriptor") {
        c.descriptor_str = trim(cur);
      } else if (name == "sparseX") {
        if (!separate_file && !sliced) {
          std::vector<double> v = parse_reals(cur);
          if ((int)v.size() != c.d) throw GapError("gpCoordinates: sparseX " + std::to_string(i_sparseX) + " does not hold 'dimensions' numbers");
          for (int k = 0; k < c.d; k++) c.sparseX[(size_t)(i_sparseX - 1) * c.d + k] = v[k];
        }
        in_sparseX = false;
      } else if (name == "sparseX_slice") {
        if (slice_start < 1) throw GapError("gpCoordinates_endElement_handler: slice start less than 1");
        if (slice_end > c.d) throw GapError("gpCoordinates_endElement_handler: slice start greater than dimension");
        if (!separate_file && sliced) {
          std::vector<double> v = parse_reals(cur);
          if ((int)v.size() != slice_end - slice_start + 1) throw GapError("gpCoordinates: sparseX_slice length mismatch");
          for (int k = slice_start; k <= slice_end; k++) c.sparseX[(size_t)(i_sparseX - 1) * c.d + (k - 1)] = v[k - slice_start];
        }
        cur.clear();
      }
    };
    xml_scan(param_str, h);
    if (!found) throw GapError("gp_readXML: no gpCoordinates element with label '" + c.label + "'");
    if (c.covariance_type == COVARIANCE_DOT_PRODUCT && !has_zeta)
      throw GapError("gpCoordinates: dot_product covariance without zeta (neither attribute nor theta element)");

    // descriptor initialise (IPModel_GAP.f95:183-188; descriptors.f95:797)
    std::string desc = c.descriptor_str + " xml_version=" + std::to_string(m.xml_version);
    std::vector<std::string> f = split_fields(desc);
    if (f.empty()) throw GapError("descriptor_initialise: empty descriptor string");
    if (f[0] == "soap") {
      c.kind = DESC_SOAP;
      c.soap = soap_from_string(desc, m.xml_version);
      if (c.soap.d != c.d)
        throw GapError("gpCoordinates dimensions=" + std::to_string(c.d) + " does not match soap descriptor dimension " + std::to_string(c.soap.d));
      if (c.covariance_type != COVARIANCE_DOT_PRODUCT)
        throw GapError("soap with covariance_type=" + std::to_string(c.covariance_type) + " is not supported by the B200 path (dot_product only)");
    } else if (f[0] == "distance_2b") {
      c.kind = DESC_DISTANCE_2B;
      c.d2b = distance_2b_from_string(desc);
      if (c.d != (int)c.d2b.exponents.size())
        throw GapError("gpCoordinates dimensions=" + std::to_string(c.d) + " does not match distance_2b n_exponents=" + std::to_string(c.d2b.exponents.size()));
      if (c.covariance_type != COVARIANCE_ARD_SE || c.n_permutations != 1)
        throw GapError("distance_2b is supported with covariance_type=ard_se and n_permutations=1 only");
      if ((int)c.theta.size() < c.d) throw GapError("gpCoordinates: ard_se covariance needs one theta per dimension");
      for (int k = 0; k < c.d; k++)
        if (c.theta[k] == 0.0) throw GapError("gpCoordinates: ard_se covariance with theta = 0");
    } else if (f[0] == "angle_3b") {
      c.kind = DESC_ANGLE_3B;
      c.a3b = angle_3b_from_string(desc);
      if (c.d != 3) throw GapError("gpCoordinates dimensions=" + std::to_string(c.d) + " does not match angle_3b (3)");
      if (c.covariance_type != COVARIANCE_ARD_SE || c.n_permutations != 1)
        throw GapError("angle_3b is supported with covariance_type=ard_se and n_permutations=1 only");
      if ((int)c.theta.size() < c.d) throw GapError("gpCoordinates: ard_se covariance needs one theta per dimension");
      for (int k = 0; k < c.d; k++)
        if (c.theta[k] == 0.0) throw GapError("gpCoordinates: ard_se covariance with theta = 0");
    } else {
      throw GapError("descriptor '" + f[0] + "' is not supported by the B200 path (soap, distance_2b and angle_3b only)");
    }
    if (c.cutoff() > m.cutoff) m.cutoff = c.cutoff();
    m.coord.push_back(std::move(c));
  }
  return m;
}

}  // namespace gapb200
