/*
 * hyperion_oracle.cpp -- CPU restatement of the reference's photon path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: a plain, scalar,
 * single-stream fp64 restatement of the Fortran algorithm, including the
 * reference's Marsaglia-Tsang generator and its exact draw order, so that it
 * can be pinned bit-for-bit (to the reference's own 1000-ULP criterion) against
 * the golden .rtout files in hyperion/model/tests/data/.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product (hyperion_b200/csrc) never links or calls it.
 *
 * Every routine cites the reference file:line it restates (paths relative to
 * the reference checkout).  Arithmetic is written in the same operand order as
 * the Fortran and must be compiled with -ffp-contract=off (no FMA), because
 * the goldens were produced by gfortran on baseline x86-64.
 */
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../include/hyperion_b200.h"

namespace {

const double PI = 3.14159265358979323846;
const double TWOPI = PI + PI;  // lib_random.f90:51

// ---------------------------------------------------------------------------
// RNG: fortranlib/src/lib_random.f90
// ---------------------------------------------------------------------------
struct Rng {
  double u[98];
  int i = 97, j = 33;
  double c = 0.0;
  uint64_t n_draws = 0;

  // set_seed (lib_random.f90:100-107) -> set_seed_64 (:109-127)
  void set_seed(int seed) {
    int a = seed < 0 ? -seed : seed;
    int32_t x = a, y = 987654321;
    for (int ii = 1; ii <= 97; ii++) {
      double s = 0.0, t = 0.5;
      for (int jj = 1; jj <= 53; jj++) {
        x = (6969 * x) % 65543;
        y = (8888 * x) % 65579;
        if (((x ^ y) & 32) > 0) s = s + t;
        t = 0.5 * t;
      }
      u[ii] = s;
    }
    (void)y;
  }

  // random_dp (lib_random.f90:172-197)
  double random() {
    const double r = 9007199254740881.0 / 9007199254740992.0;
    const double d = 362436069876.0 / 9007199254740992.0;
    double x = u[i] - u[j];
    if (x < 0.0) x = x + 1.0;
    u[i] = x;
    i = i - 1;
    if (i == 0) i = 97;
    j = j - 1;
    if (j == 0) j = 97;
    c = c - d;
    if (c < 0.0) c = c + r;
    x = x - c;
    n_draws++;
    if (x < 0.0) return x + 1.0;
    return x;
  }

  // random_uni_dp (lib_random.f90:200-207)
  double random_uni(double a, double b) {
    double xi = random();
    return a + (b - a) * xi;
  }

  // random_exp_dp (lib_random.f90:227-236)
  double random_exp() {
    double xi;
    do {
      xi = random();
    } while (!(xi < 1.0));
    return -std::log(1.0 - xi);
  }

  // random_planck_frequency_dp (lib_random.f90:297-347)
  double random_planck_frequency(double T) {
    const double k = 1.3806503e-23;
    const double h = 6.626068e-34;
    double r;
    for (;;) {
      double r1 = random();
      double r2 = random();
      double r3 = random();
      double r4 = random();
      r = r1 * r2 * r3 * r4;
      if (r > 0.0) break;
    }
    double x = -std::log(r);
    double a = 1.0, y = 1.0, z = 1.0;
    double r1 = random();
    for (;;) {
      if (1.08232 * r1 <= a) break;
      y = y + 1.0;
      z = 1.0 / y;
      a = a + z * z * z * z;
    }
    x = x * z;
    return x * k * T / h;
  }
};

// ---------------------------------------------------------------------------
// lib_array.f90 numerics (1-based indices returned, as in the Fortran)
// ---------------------------------------------------------------------------

// locate_dp (lib_array.f90:917-950)
int locate(const double *xx, int n, double x) {
  bool ascnd = (xx[n - 1] >= xx[0]);
  int jl = 0, ju = n + 1;
  for (;;) {
    if (ju - jl <= 1) break;
    int jm = (ju + jl) / 2;
    if (ascnd == (x >= xx[jm - 1]))
      jl = jm;
    else
      ju = jm;
  }
  if (x == xx[0]) return 1;
  if (x == xx[n - 1]) return n - 1;
  if (ascnd && (x > xx[n - 1] || x < xx[0])) return -1;
  if (!ascnd && (x < xx[n - 1] || x > xx[0])) return -1;
  return jl;
}

// ipos_dp (lib_array.f90:954-998)
int ipos(double xmin, double xmax, double x, int nbin) {
  if (xmax > xmin) {
    if (x < xmin) return 0;
    if (x > xmax) return nbin + 1;
    if (x < xmax) {
      double frac = (x - xmin) / (xmax - xmin);
      return (int)(frac * (double)nbin) + 1;
    }
    return nbin;
  } else {
    if (x > xmin) return 0;
    if (x < xmax) return nbin + 1;
    if (x > xmax) {
      double frac = (x - xmin) / (xmax - xmin);
      return (int)(frac * (double)nbin) + 1;
    }
    return nbin;
  }
}

// trapezium_dp (lib_array.f90:528-532)
double trapezium(double x1, double y1, double x2, double y2) {
  return 0.5 * (y1 + y2) * (x2 - x1);
}

// trapezium_linlog_dp (lib_array.f90:548-560)
double trapezium_linlog(double x1, double y1, double x2, double y2) {
  if (x1 == x2) return 0.0;
  if (y1 == y2) return y1 * (x2 - x1);
  return (y2 - y1) * (x2 - x1) / std::log(10.0) / std::log10(y2 / y1);
}

// trapezium_loglog_dp (lib_array.f90:562-578)
double trapezium_loglog(double x1, double y1, double x2, double y2) {
  if (x1 == x2) return 0.0;
  if (y1 == 0.0 || y2 == 0.0) return 0.0;
  double b = std::log10(y1 / y2) / std::log10(x1 / x2);
  // note: the Fortran compares with the single-precision literal 1e-10
  if (std::fabs(b + 1.0) < (double)1e-10f) return x1 * y1 * std::log(x2 / x1);
  return y1 * (x2 * std::pow(x2 / x1, b) - x1) / (b + 1);
}

typedef double (*chunk_fn)(double, double, double, double);

// integral_general_dp (lib_array.f90:413-429)
double integral_general(const double *x, const double *y, int n, chunk_fn f) {
  double sum = 0.0;
  for (int j = 0; j < n - 1; j++) sum = sum + f(x[j], y[j], x[j + 1], y[j + 1]);
  return sum;
}

// cumulative_integral_general_dp (lib_array.f90:431-448)
void cumulative_integral_general(const double *x, const double *y, int n, chunk_fn f, double *c) {
  c[0] = 0.0;
  for (int j = 0; j < n - 1; j++) c[j + 1] = c[j] + f(x[j], y[j], x[j + 1], y[j + 1]);
}

// interp1d_single_loglog_dp (lib_array.f90:605-614)
double interp1d_single_loglog(double x1, double y1, double x2, double y2, double xval) {
  if (y1 == 0.0 || y2 == 0.0) return 0.0;
  double frac = (std::log10(xval) - std::log10(x1)) / (std::log10(x2) - std::log10(x1));
  return std::pow(10.0, std::log10(y1) + frac * (std::log10(y2) - std::log10(y1)));
}

// interp1d_single_linlog_dp (lib_array.f90:587-596)
double interp1d_single_linlog(double x1, double y1, double x2, double y2, double xval) {
  if (y1 == 0.0 || y2 == 0.0) return 0.0;
  double frac = (xval - x1) / (x2 - x1);
  return std::pow(10.0, std::log10(y1) + frac * (std::log10(y2) - std::log10(y1)));
}

struct OracleError {
  std::string msg;
};

// interp1d_general_dp with loglog chunks (lib_array.f90:682-689,704-778)
double interp1d_loglog(const double *x, const double *y, int n, double xval, bool bounds_error = true,
                       double fill_value = 0.0) {
  int ip = locate(x, n, xval);
  if (ip == -1) {
    if (bounds_error) throw OracleError{"Interpolation out of bounds"};
    return fill_value;
  }
  if (ip < n && ip > 0) return interp1d_single_loglog(x[ip - 1], y[ip - 1], x[ip], y[ip], xval);
  if (ip == n) return y[n - 1];
  if (ip == 0) return y[0];
  throw OracleError{"Unexpected value of ipos"};
}

// interp1d_dp = interp1d_general_dp with linear chunks (lib_array.f90:578-585,616-624,704-778)
double interp1d_lin(const double *x, const double *y, int n, double xval, bool bounds_error = true,
                    double fill_value = 0.0) {
  int ip = locate(x, n, xval);
  if (ip == -1) {
    if (bounds_error) throw OracleError{"Interpolation out of bounds"};
    return fill_value;
  }
  if (ip < n && ip > 0) {
    double frac = (xval - x[ip - 1]) / (x[ip] - x[ip - 1]);
    return y[ip - 1] + frac * (y[ip] - y[ip - 1]);
  }
  if (ip == n) return y[n - 1];
  if (ip == 0) return y[0];
  throw OracleError{"Unexpected value of ipos"};
}

// interp2d_dp (lib_array.f90:780-846); array(i,j) stored as a[(j-1)*nx + (i-1)]
double interp2d(const double *x, int nx, const double *y, int ny, const double *a, double x0, double y0) {
  int i1 = locate(x, nx, x0), i2 = i1 + 1;
  int j1 = locate(y, ny, y0), j2 = j1 + 1;
  if (i1 == -1 || j1 == -1) throw OracleError{"Interpolation out of bounds"};
  double norm = 1.0 / (x[i2 - 1] - x[i1 - 1]) / (y[j2 - 1] - y[j1 - 1]);
#define A(i, j) a[(size_t)((j)-1) * nx + ((i)-1)]
  double value = A(i1, j1) * (x[i2 - 1] - x0) * (y[j2 - 1] - y0) * norm +
                 A(i2, j1) * (x0 - x[i1 - 1]) * (y[j2 - 1] - y0) * norm +
                 A(i1, j2) * (x[i2 - 1] - x0) * (y0 - y[j1 - 1]) * norm +
                 A(i2, j2) * (x0 - x[i1 - 1]) * (y0 - y[j1 - 1]) * norm;
#undef A
  return value;
}

// ---------------------------------------------------------------------------
// type_pdf.f90
// ---------------------------------------------------------------------------
struct PdfDiscrete {
  int n = 0;
  std::vector<double> pdf, cdf;

  // find_cdf_discrete_dp (type_pdf.f90:167-180)
  void find_cdf() {
    cdf[0] = pdf[0];
    for (int i = 1; i < n; i++) cdf[i] = cdf[i - 1] + pdf[i];
    double norm = cdf[n - 1];
    if (norm == 0.0) throw OracleError{"[find_cdf_discrete] all PDF elements are zero"};
    for (int i = 0; i < n; i++) cdf[i] = cdf[i] / norm;
  }
  // set_pdf_discrete_dp (type_pdf.f90:222-231) + normalize (:200-209)
  void set(const double *y, int nn) {
    n = nn;
    pdf.assign(y, y + nn);
    cdf.assign(nn, 0.0);
    double norm = 0.0;
    for (int i = 0; i < n; i++) norm = norm + pdf[i];
    if (norm == 0.0) throw OracleError{"[normalize_pdf_discrete] all PDF elements are zero"};
    for (int i = 0; i < n; i++) pdf[i] = pdf[i] / norm;
    find_cdf();
  }
  // sample_pdf_discrete_dp (type_pdf.f90:313-337)
  int sample(Rng &rng) const {
    double xi = rng.random();
    if (xi <= cdf[0]) return 1;
    if (xi >= cdf[n - 1]) return n;
    int jmin = 1, jmax = n;
    for (;;) {
      int j = (jmax + jmin) / 2;
      if (xi > cdf[j - 1])
        jmin = j;
      else
        jmax = j;
      if (jmax == jmin + 1) break;
    }
    return jmax;
  }
};

struct PdfCont {
  int n = 0;
  bool log = false;
  std::vector<double> x, pdf, cdf, a, b, r, rx, rc;

  // set_pdf_cont_dp (type_pdf.f90:233-248) -> normalize (:211-220) -> find_cdf (:279-311)
  void set(const double *xs, const double *ys, int nn, bool is_log) {
    n = nn;
    log = is_log;
    x.assign(xs, xs + nn);
    pdf.assign(ys, ys + nn);
    cdf.assign(nn, 0.0);
    double norm = log ? integral_general(x.data(), pdf.data(), n, trapezium_loglog)
                      : integral_general(x.data(), pdf.data(), n, trapezium);
    for (int i = 0; i < n; i++) pdf[i] = pdf[i] / norm;
    for (int i = 1; i < n; i++)
      if (!(x[i] > x[i - 1])) throw OracleError{"[check_pdf] PDF x array is not sorted"};
    if (log)
      cumulative_integral_general(x.data(), pdf.data(), n, trapezium_loglog, cdf.data());
    else
      cumulative_integral_general(x.data(), pdf.data(), n, trapezium, cdf.data());
    double last = cdf[n - 1];
    for (int i = 0; i < n; i++) cdf[i] = cdf[i] / last;
    if (log) {
      b.assign(n - 1, 0.0);
      r.assign(n - 1, 0.0);
      for (int i = 0; i < n - 1; i++) {
        b[i] = std::log10(pdf[i] / pdf[i + 1]) / std::log10(x[i] / x[i + 1]);
        r[i] = std::pow(x[i + 1] / x[i], b[i] + 1.0);
      }
    } else {
      a.assign(n - 1, 0.0);
      b.assign(n - 1, 0.0);
      rx.assign(n - 1, 0.0);
      rc.assign(n - 1, 0.0);
      for (int i = 0; i < n - 1; i++) {
        a[i] = (pdf[i] - pdf[i + 1]) / (x[i] - x[i + 1]);
        b[i] = pdf[i] - a[i] * x[i];
        rx[i] = x[i + 1] / x[i];
        rc[i] = b[i] / a[i];
      }
    }
  }

  // sample_pdf_cont_dp (type_pdf.f90:339-381), non-"simple" branch
  double sample_xi(double xi) const {
    if (xi <= cdf[0]) return x[0];
    if (xi >= cdf[n - 1]) return x[n - 1];
    int i = locate(cdf.data(), n, xi);
    xi = (xi - cdf[i - 1]) / (cdf[i] - cdf[i - 1]);
    if (log) {
      return std::pow(xi * (r[i - 1] - 1.0) + 1.0, 1.0 / (b[i - 1] + 1.0)) * x[i - 1];
    } else {
      double A = a[i - 1], RC = rc[i - 1], X = x[i - 1], X1 = x[i], RX = rx[i - 1];
      if (A == 0.0) return xi * (X1 - X) + X;
      if (X == 0.0) return -RC + std::copysign(std::sqrt(RC * RC + xi * X1 * X1 + 2.0 * RC * xi * X1), A);
      return -RC + std::copysign(std::sqrt(RC * RC + X * X * (xi * (RX * RX - 1.0) + 1.0) +
                                           2.0 * RC * X * (xi * (RX - 1.0) + 1.0)),
                                 A);
    }
  }
  double sample(Rng &rng) const { return sample_xi(rng.random()); }
  // interpolate_pdf_cont_dp (type_pdf.f90:402-414) with bounds_error=.false., fill_value=0
  double interpolate(double xv) const {
    return log ? interp1d_loglog(x.data(), pdf.data(), n, xv, false, 0.0) : interp1d_lin(x.data(), pdf.data(), n, xv, false, 0.0);
  }
  // sample_pdf_cont_log_dp (type_pdf.f90:383-400): x interpolated in the log against the cdf
  double sample_log(Rng &rng) const {
    double xi = rng.random();
    if (xi <= cdf[0]) return x[0];
    if (xi >= cdf[n - 1]) return x[n - 1];
    int ip = locate(cdf.data(), n, xi);
    if (ip == -1) throw OracleError{"Interpolation out of bounds"};
    if (ip < n && ip > 0) return interp1d_single_linlog(cdf[ip - 1], x[ip - 1], cdf[ip], x[ip], xi);
    return ip == n ? x[n - 1] : x[0];
  }
};

// ---------------------------------------------------------------------------
// type_angle3d.f90 / type_vector3d.f90 / type_stokes
// ---------------------------------------------------------------------------
struct Angle {
  double cost, sint, cosp, sinp;
};
struct Vec {
  double x, y, z;
};
struct Stokes {
  double I, Q, U, V;
};

// random_sphere_angle3d_dp (type_angle3d.f90:431-441); random_sphere_dp (lib_random.f90:239-245)
Angle random_sphere_angle3d(Rng &rng) {
  Angle a;
  a.cost = rng.random_uni(-1.0, +1.0);
  double phi = rng.random_uni(0.0, TWOPI);
  a.sint = std::sqrt(1.0 - a.cost * a.cost);
  a.cosp = std::cos(phi);
  a.sinp = std::sin(phi);
  return a;
}

// angle3d_to_vector3d_dp (type_vector3d.f90:301-317)
Vec angle3d_to_vector3d(const Angle &a) { return Vec{a.sint * a.cosp, a.sint * a.sinp, a.cost}; }

// sin2cos_dp (type_angle3d.f90:421-429)
double sin2cos(double x) { return (x * x < 1.0) ? std::sqrt(1.0 - x * x) : 0.0; }

Angle angle3d_deg(double theta, double phi) {
  const double deg2rad = PI / 180.0;
  return Angle{std::cos(theta * deg2rad), std::sin(theta * deg2rad), std::cos(phi * deg2rad),
               std::sin(phi * deg2rad)};
}

// rotate_angle3d_dp (type_angle3d.f90:160-281)
Angle rotate_angle3d(const Angle &a_local, const Angle &a_coord) {
  Angle f;
  if (std::fabs(a_coord.sint) < 1.e-10) {
    f = a_local;
    if (a_coord.cost > 0.0) {
      f.cosp = +a_local.cosp * a_coord.cosp + a_local.sinp * a_coord.sinp;
      f.sinp = +a_local.cosp * a_coord.sinp - a_local.sinp * a_coord.cosp;
    } else {
      f.cost = -a_local.cost;
      f.cosp = +a_local.cosp * a_coord.cosp - a_local.sinp * a_coord.sinp;
      f.sinp = +a_local.cosp * a_coord.sinp + a_local.sinp * a_coord.cosp;
    }
    return f;
  }
  double cos_a = a_coord.cost, sin_a = a_coord.sint;
  double cos_b = a_local.cost, sin_b = a_local.sint;
  double cos_big_c, sin_big_c;
  if (a_local.sinp < 0.0) {
    cos_big_c = +a_local.cosp;
    sin_big_c = -a_local.sinp;
  } else {
    cos_big_c = +a_local.cosp;
    sin_big_c = +a_local.sinp;
  }
  bool same_sign;
  double delta;
  if (std::fabs(sin_a) > std::fabs(cos_a)) {
    same_sign = (sin_a > 0.0) == (sin_b > 0.0);
    delta = cos_b - cos_a;
  } else {
    same_sign = (cos_a > 0.0) == (cos_b > 0.0);
    delta = sin_b - sin_a;
  }
  double cos_c, sin_c;
  if (same_sign && std::fabs(delta) < 1.e-5 && sin_big_c < 1.e-5 && cos_big_c > 0.0) {
    if (std::fabs(sin_a) > std::fabs(cos_a)) {
      double q = cos_a / sin_a;
      sin_c = std::sqrt(delta * delta * (1.0 + q * q) + sin_a * sin_b * sin_big_c * sin_big_c);
    } else {
      double q = sin_a / cos_a;
      sin_c = std::sqrt(delta * delta * (1.0 + q * q) + sin_a * sin_b * sin_big_c * sin_big_c);
    }
    cos_c = sin2cos(sin_c);
  } else {
    cos_c = cos_a * cos_b + sin_a * sin_b * cos_big_c;
    sin_c = sin2cos(cos_c);
  }
  if (std::fabs(sin_c) < 1.e-10) {
    if (cos_c > 0.0) return angle3d_deg(0.0, 0.0);
    return angle3d_deg(180.0, 0.0);
  }
  double cos_big_b = (cos_b - cos_a * cos_c) / (sin_a * sin_c);
  double sin_big_b = +sin_big_c * sin_b / sin_c;
  f.cost = cos_c;
  f.sint = sin_c;
  if (a_local.sinp < 0.0) {
    f.cosp = +cos_big_b * a_coord.cosp + sin_big_b * a_coord.sinp;
    f.sinp = +cos_big_b * a_coord.sinp - sin_big_b * a_coord.cosp;
  } else {
    f.cosp = +cos_big_b * a_coord.cosp - sin_big_b * a_coord.sinp;
    f.sinp = +cos_big_b * a_coord.sinp + sin_big_b * a_coord.cosp;
  }
  return f;
}

// ---------------------------------------------------------------------------
// dust: src/dust/dust_type_4elem.f90
// ---------------------------------------------------------------------------
struct Dust {
  int version = 2, sublimation_mode = 0;
  double sublimation_specific_energy = 0.0;
  int n_nu = 0, n_mu = 0, n_e = 0, n_jnu = 0;
  std::vector<double> nu, albedo_nu, chi_nu, kappa_nu;
  std::vector<double> mu;
  std::vector<double> P1, P2, P3, P4, P1_cdf, P2_cdf, P3_cdf, P4_cdf;  // (imu, inu) at [inu*n_mu+imu]
  double mu_min = -1, mu_max = 1;
  std::vector<double> specific_energy, chi_planck, kappa_planck, chi_inv_planck, kappa_inv_planck,
      chi_rosseland, kappa_rosseland;
  std::vector<double> j_nu_var, log10_j_nu_var;
  std::vector<PdfCont> j_nu, b_nu;  // b_nu: emissivity divided by opacity (dust_type_4elem.f90:286-291)
  bool zero_p2 = true;
  bool have_b_nu = false;

  // dust_setup (dust_type_4elem.f90:78-293)
  void setup(const hyp_dust_tables &t) {
    version = t.version;
    sublimation_mode = t.sublimation_mode;
    sublimation_specific_energy = t.sublimation_specific_energy;
    n_nu = t.n_nu;
    n_mu = t.n_mu;
    nu.assign(t.nu, t.nu + n_nu);
    albedo_nu.assign(t.albedo, t.albedo + n_nu);
    chi_nu.assign(t.chi, t.chi + n_nu);
    size_t np = (size_t)n_nu * n_mu;
    P1.assign(t.P1, t.P1 + np);
    P2.assign(t.P2, t.P2 + np);
    P3.assign(t.P3, t.P3 + np);
    P4.assign(t.P4, t.P4 + np);
    zero_p2 = true;
    for (size_t i = 0; i < np; i++)
      if (P2[i] != 0.0) zero_p2 = false;
    kappa_nu.resize(n_nu);
    for (int j = 0; j < n_nu; j++) kappa_nu[j] = chi_nu[j] * (1.0 - albedo_nu[j]);
    mu.assign(t.mu, t.mu + n_mu);
    mu_min = mu[0];
    mu_max = mu[n_mu - 1];
    double dmu = mu_max - mu_min;
    // normalise so that the integral over mu is dmu (:193-203)
    for (int j = 0; j < n_nu; j++) {
      double *p1 = &P1[(size_t)j * n_mu];
      double norm = integral_general(mu.data(), p1, n_mu, trapezium_linlog);
      if (norm == 0.0) throw OracleError{"P1 matrix normalization is zero"};
      for (int i = 0; i < n_mu; i++) {
        size_t k = (size_t)j * n_mu + i;
        P1[k] = P1[k] / norm * dmu;
        P2[k] = P2[k] / norm * dmu;
        P3[k] = P3[k] / norm * dmu;
        P4[k] = P4[k] / norm * dmu;
      }
    }
    P1_cdf.assign(np, 0.0);
    P2_cdf.assign(np, 0.0);
    P3_cdf.assign(np, 0.0);
    P4_cdf.assign(np, 0.0);
    std::vector<double> *Ps[4] = {&P1, &P2, &P3, &P4};
    std::vector<double> *Cs[4] = {&P1_cdf, &P2_cdf, &P3_cdf, &P4_cdf};
    for (int j = 0; j < n_nu; j++) {
      for (int q = 0; q < 4; q++) {
        double *c = &(*Cs[q])[(size_t)j * n_mu];
        cumulative_integral_general(mu.data(), &(*Ps[q])[(size_t)j * n_mu], n_mu, trapezium, c);
        bool all_zero = true;
        for (int i = 0; i < n_mu; i++)
          if (c[i] != 0.0) all_zero = false;
        if (!all_zero) {
          double last = c[n_mu - 1];
          for (int i = 0; i < n_mu; i++) c[i] = c[i] / last;
        }
      }
    }
    n_e = t.n_e;
    specific_energy.assign(t.specific_energy, t.specific_energy + n_e);
    chi_planck.assign(t.chi_planck, t.chi_planck + n_e);
    kappa_planck.assign(t.kappa_planck, t.kappa_planck + n_e);
    chi_inv_planck.assign(t.chi_inv_planck, t.chi_inv_planck + n_e);
    kappa_inv_planck.assign(t.kappa_inv_planck, t.kappa_inv_planck + n_e);
    chi_rosseland.assign(t.chi_rosseland, t.chi_rosseland + n_e);
    kappa_rosseland.assign(t.kappa_rosseland, t.kappa_rosseland + n_e);
    for (int i = 1; i < n_e; i++)
      if (specific_energy[i] < specific_energy[i - 1])
        throw OracleError{"energy per unit mass is not monotonically increasing"};
    n_jnu = t.n_jnu;
    j_nu_var.assign(t.jnu_var, t.jnu_var + n_jnu);
    log10_j_nu_var.resize(n_jnu);
    for (int i = 0; i < n_jnu; i++) log10_j_nu_var[i] = std::log10(j_nu_var[i]);
    j_nu.resize(n_jnu);
    std::vector<double> col(t.n_emiss_nu);
    for (int i = 0; i < n_jnu; i++) {
      for (int k = 0; k < t.n_emiss_nu; k++) col[k] = t.emiss_jnu[(size_t)k * n_jnu + i];
      j_nu[i].set(t.emiss_nu, col.data(), t.n_emiss_nu, true);
    }
    // b_nu = j_nu / kappa_nu on the emissivity frequency grid, used by the modified random walk
    b_nu.resize(n_jnu);
    std::vector<double> kap(t.n_emiss_nu);
    bool kap_ok = true;
    for (int k = 0; k < t.n_emiss_nu; k++) {
      if (t.emiss_nu[k] < nu[0] || t.emiss_nu[k] > nu[n_nu - 1]) {
        kap_ok = false;
        break;
      }
      kap[k] = interp1d_loglog(nu.data(), kappa_nu.data(), n_nu, t.emiss_nu[k]);
    }
    have_b_nu = kap_ok;
    if (kap_ok)
      for (int i = 0; i < n_jnu; i++) {
        for (int k = 0; k < t.n_emiss_nu; k++) col[k] = t.emiss_jnu[(size_t)k * n_jnu + i] / kap[k];
        try {
          b_nu[i].set(t.emiss_nu, col.data(), t.n_emiss_nu, true);
        } catch (OracleError &) {
          have_b_nu = false;
          break;
        }
      }
  }

  // dust_jnu_var_pos_frac (dust_type_4elem.f90:295-320)
  void jnu_var_pos_frac(double se, int &id, double &frac) const {
    double v = se;
    if (v < j_nu_var[0]) {
      id = 1;
      frac = 0.0;
    } else if (v > j_nu_var[n_jnu - 1]) {
      id = n_jnu - 1;
      frac = 1.0;
    } else {
      id = locate(j_nu_var.data(), n_jnu, v);
      frac = (std::log10(v) - log10_j_nu_var[id - 1]) / (log10_j_nu_var[id] - log10_j_nu_var[id - 1]);
    }
  }

  // dust_sample_b_nu (dust_type_4elem.f90:400-419)
  double sample_b_nu(Rng &rng, int id, double frac) const {
    double xi = rng.random();
    double nu1 = b_nu[id - 1].sample_xi(xi);
    double nu2 = b_nu[id].sample_xi(xi);
    double v = std::log10(nu1) + frac * (std::log10(nu2) - std::log10(nu1));
    return std::pow(10.0, v);
  }

  // dust_sample_j_nu (dust_type_4elem.f90:379-398)
  double sample_j_nu(Rng &rng, int id, double frac) const {
    double xi = rng.random();
    double nu1 = j_nu[id - 1].sample_xi(xi);
    double nu2 = j_nu[id].sample_xi(xi);
    double v = std::log10(nu1) + frac * (std::log10(nu2) - std::log10(nu1));
    return std::pow(10.0, v);
  }
};

// scatter_stokes (dust_type_4elem.f90:603-690)
void scatter_stokes(Stokes &s, const Angle &a_coord, const Angle &a_scat, const Angle &a_final, double P1,
                    double P2, double P3, double P4) {
  const double tiny = std::numeric_limits<double>::min();
  double cos_a = a_coord.cost, sin_a = a_coord.sint;
  double cos_b = a_scat.cost, sin_b = a_scat.sint;
  double cos_c = a_final.cost, sin_c = a_final.sint;
  double cos_big_b = a_coord.cosp * a_final.cosp + a_coord.sinp * a_final.sinp;
  double cos_big_c = a_scat.cosp;
  double sin_big_c = std::fabs(a_scat.sinp);
  double cos_big_a, sin_big_a;
  if (sin_big_c < 10. * tiny && sin_c < 10. * tiny) {
    cos_big_a = -cos_big_b * cos_big_c;
    sin_big_a = std::sqrt(1.0 - cos_big_a * cos_big_a);
  } else {
    cos_big_a = (cos_a - cos_b * cos_c) / (sin_b * sin_c);
    sin_big_a = +sin_big_c * sin_a / sin_c;
  }
  double cos_i2 = cos_big_a, sin_i2 = sin_big_a;
  double cos_2_i2 = 1.0 - 2.0 * sin_i2 * sin_i2;
  double sin_2_i2 = 2.0 * sin_i2 * cos_i2;
  double cos_2_alpha = 1.0 - 2.0 * a_scat.sinp * a_scat.sinp;
  double sin_2_alpha = -2.0 * a_scat.sinp * a_scat.cosp;
  double cos_2_beta, sin_2_beta;
  if (a_scat.sinp < 0.) {
    cos_2_beta = cos_2_i2;
    sin_2_beta = sin_2_i2;
  } else {
    cos_2_beta = cos_2_i2;
    sin_2_beta = -sin_2_i2;
  }
  double RLS1 = P1 * s.I + P2 * (+cos_2_alpha * s.Q + sin_2_alpha * s.U);
  double RLS2 = P2 * s.I + P1 * (+cos_2_alpha * s.Q + sin_2_alpha * s.U);
  double RLS3 = -P4 * s.V + P3 * (-sin_2_alpha * s.Q + cos_2_alpha * s.U);
  double RLS4 = P3 * s.V + P4 * (-sin_2_alpha * s.Q + cos_2_alpha * s.U);
  s.I = RLS1;
  s.Q = +cos_2_beta * RLS2 + sin_2_beta * RLS3;
  s.U = -sin_2_beta * RLS2 + cos_2_beta * RLS3;
  s.V = RLS4;
}

// dust_scatter (dust_type_4elem.f90:446-566)
void dust_scatter(const Dust &d, Rng &rng, double nu, Angle &a, Stokes &s) {
  Angle a_scat = random_sphere_angle3d(rng);
  double sin_2_i1 = 2.0 * a_scat.sinp * a_scat.cosp;
  double cos_2_i1 = 1.0 - 2.0 * a_scat.sinp * a_scat.sinp;
  double c1 = s.I;
  double c2 = (cos_2_i1 * s.Q - sin_2_i1 * s.U);
  double ctot = c1 + c2;
  c1 = c1 / ctot;
  c2 = c2 / ctot;
  int imin = 1, imax = d.n_mu;
  int inu = locate(d.nu.data(), d.n_nu, nu);
  double P1, P2, P3, P4;
  if (inu == -1) {
    P1 = 1.0;
    P2 = 0.0;
    P3 = 1.0;
    P4 = 0.0;
  } else {
    double xi = rng.random();
    int imu = 0;
    double cdf1 = 0, cdf2 = 0;
    const double *c1p = &d.P1_cdf[(size_t)(inu - 1) * d.n_mu];
    const double *c2p = &d.P2_cdf[(size_t)(inu - 1) * d.n_mu];
    for (int iter = 1; iter <= 1000000; iter++) {
      imu = (imax + imin) / 2;
      if (d.zero_p2) {
        cdf1 = c1p[imu - 1];
        cdf2 = c1p[imu];
      } else {
        cdf1 = c1 * c1p[imu - 1] + c2 * c2p[imu - 1];
        cdf2 = c1 * c1p[imu] + c2 * c2p[imu];
      }
      if (xi > cdf2)
        imin = imu;
      else if (xi < cdf1)
        imax = imu;
      else
        break;
      if (imin == imax) throw OracleError{"ERROR: in sampling mu for scattering"};
    }
    a_scat.cost = (xi - cdf1) / (cdf2 - cdf1) * (d.mu[imu] - d.mu[imu - 1]) + d.mu[imu - 1];
    a_scat.sint = std::sqrt(1.0 - a_scat.cost * a_scat.cost);
    P1 = interp2d(d.mu.data(), d.n_mu, d.nu.data(), d.n_nu, d.P1.data(), a_scat.cost, nu);
    P2 = interp2d(d.mu.data(), d.n_mu, d.nu.data(), d.n_nu, d.P2.data(), a_scat.cost, nu);
    P3 = interp2d(d.mu.data(), d.n_mu, d.nu.data(), d.n_nu, d.P3.data(), a_scat.cost, nu);
    P4 = interp2d(d.mu.data(), d.n_mu, d.nu.data(), d.n_nu, d.P4.data(), a_scat.cost, nu);
  }
  Angle a_final = rotate_angle3d(a_scat, a);
  scatter_stokes(s, a, a_scat, a_final, P1, P2, P3, P4);
  a = a_final;
  double norm = 1.0 / s.I;
  s.I = 1.0;
  s.Q = s.Q * norm;
  s.U = s.U * norm;
  s.V = s.V * norm;
}

// ---------------------------------------------------------------------------
// photon: src/core/type_photon.f90:14-73
// ---------------------------------------------------------------------------
struct WallId {
  int w1 = 0, w2 = 0, w3 = 0;
};
struct Cell {
  int i1 = 0, i2 = 0, i3 = 0, ic = 0;
  int ilevel = 0, igrid = 0;  // AMR only (type_cell_id_amr.f90:14-19)
};
const int MAX_DUST = 16;
struct Photon {
  Vec r{0, 0, 0}, v{0, 0, 0};
  Angle a{0, 0, 0, 0};
  Stokes s{0, 0, 0, 0};
  double nu = 0, energy = 0;
  bool in_cell = false, on_wall = false;
  WallId on_wall_id;
  bool killed = false, reabsorbed = false;
  int reabsorbed_id = 0;
  double current_chi[MAX_DUST], current_albedo[MAX_DUST], current_kappa[MAX_DUST];
  Cell icell;
  bool last_isotropic = false, scattered = false, reprocessed = false;
  int n_scat = 0, source_id = 0, dust_id = 0;
  int64_t id = 0;  // photon_counter at emission (source.f90:171-172)
  char last[3] = "  ";
  // state before the last interaction, used by peel-off (type_photon.f90:44-46)
  Angle a_prev{0, 0, 0, 0};
  Stokes s_prev{0, 0, 0, 0};
  Vec v_prev{0, 0, 0};
  int emiss_type = 0, emiss_var_id = 0;
  double emiss_var_frac = 0.0;
  int inu = 0;   // monochromatic mode: index of the packet's frequency (1-based)
  Angle source_a{0, 0, 0, 0};  // position angle on a stellar surface (type_photon.f90, emit_from_sphere)
  int face_id = 0;             // face of an external box source the packet came from (emit_from_extern_box)
};

struct Source {
  int type = 1;
  bool peeloff = true;
  double luminosity = 0;
  Vec position{0, 0, 0};
  double radius = 0;
  bool limb_darkening = false;
  int freq_type = 2;
  double temperature = 0;
  PdfCont spectrum;
  bool intersect = false;
  double xmin = 0, xmax = 0, ymin = 0, ymax = 0, zmin = 0, zmax = 0;  // extern_box
  PdfDiscrete face;                                                    // extern_box: area of the six faces
  Angle direction{1.0, 0.0, 1.0, 0.0};                                 // plane_parallel
  std::vector<Vec> position_collection;                                // point_collection
  PdfDiscrete collection_pdf;
  PdfDiscrete luminosity_map;                                          // map: one entry per cell
  // spots (source_type.f90:27-33,150-188): spot_pdf has n_spots + 1 entries, the last one is the star itself
  struct Spot {
    Angle a;
    double cost = 0;
    int freq_type = 2;
    double temperature = 0;
    PdfCont spectrum;
  };
  std::vector<Spot> spot;
  PdfDiscrete spot_pdf;
};

// ---------------------------------------------------------------------------
// final iteration, peel-off and raytracing
// ---------------------------------------------------------------------------

// integral_general_subset_dp with loglog pieces (lib_array.f90:362-366,450-526)
double integral_loglog_subset(const double *x, const double *y, int n, double x1, double x2) {
  if (x1 > x[n - 1] || x2 < x[0]) return 0.0;
  int i1, i2;
  double f1, f2, xx1, xx2;
  if (x1 > x[0]) {
    i1 = locate(x, n, x1);
    f1 = interp1d_loglog(x, y, n, x1);
    xx1 = x1;
  } else {
    i1 = 0;
    f1 = y[0];
    xx1 = x[0];
  }
  if (x2 < x[n - 1]) {
    i2 = locate(x, n, x2);
    f2 = interp1d_loglog(x, y, n, x2);
    xx2 = x2;
  } else {
    i2 = n;
    f2 = y[n - 1];
    xx2 = x[n - 1];
  }
  double sum;
  if (i2 > i1) {
    if (i2 > i1 + 1)
      sum = integral_general(x + i1, y + i1, i2 - i1, trapezium_loglog);  // x(i1+1:i2)
    else
      sum = 0.0;
    sum = sum + trapezium_loglog(xx1, f1, x[i1], y[i1]);          // x(i1+1)
    sum = sum + trapezium_loglog(x[i2 - 1], y[i2 - 1], xx2, f2);  // x(i2)
  } else {
    sum = trapezium_loglog(xx1, f1, xx2, f2);
  }
  return sum;
}

// difference_angle3d_dp (type_angle3d.f90:283-419)
Angle difference_angle3d(const Angle &a_coord, const Angle &a_final) {
  Angle l;
  if (std::fabs(a_coord.sint) < 1.e-10) {
    l = a_final;
    if (a_coord.cost > 0.0) {
      l.cosp = +a_coord.cosp * a_final.cosp + a_coord.sinp * a_final.sinp;
      l.sinp = -a_coord.cosp * a_final.sinp + a_coord.sinp * a_final.cosp;
    } else {
      l.cost = -l.cost;
      l.cosp = +a_coord.cosp * a_final.cosp + a_coord.sinp * a_final.sinp;
      l.sinp = +a_coord.cosp * a_final.sinp - a_coord.sinp * a_final.cosp;
    }
    return l;
  }
  double cos_a = a_coord.cost, sin_a = a_coord.sint, cos_c = a_final.cost, sin_c = a_final.sint;
  double cos_big_b = a_coord.cosp * a_final.cosp + a_coord.sinp * a_final.sinp;
  double sin_big_b = a_coord.sinp * a_final.cosp - a_coord.cosp * a_final.sinp;
  double cos_b = cos_a * cos_c + sin_a * sin_c * cos_big_b;
  double sin_b = sin2cos(cos_b);  // cos2sin is the same expression
  if (std::fabs(cos_b + 1.0) < 1.e-10) return angle3d_deg(180.0, 0.0);
  if (std::fabs(cos_b - 1.0) < 1.e-10) return angle3d_deg(0.0, 0.0);
  // Fortran precedence: .eqv. binds weaker than .and.:  A .eqv. (B .and. C) .eqv. D
  bool same_sign = ((cos_a > 0.0) == ((cos_b > 0.0) && (sin_a > 0.0))) == (sin_b > 0.0);
  double delta;
  if (std::fabs(sin_a) > std::fabs(cos_a))
    delta = cos_b - cos_a;
  else
    delta = sin_b - sin_a;
  double sin_big_c, cos_big_c;
  if (same_sign && std::fabs(delta) < 1.e-5 && sin_c < 1.e-5) {
    double diff;
    if (std::fabs(sin_a) > std::fabs(cos_a))
      diff = (sin_c * sin_c - delta * delta * (1.0 + (cos_a / sin_a) * (cos_a / sin_a))) / (sin_a * sin_b);
    else
      diff = (sin_c * sin_c - delta * delta * (1.0 + (sin_a / cos_a) * (sin_a / cos_a))) / (sin_a * sin_b);
    sin_big_c = diff >= 0.0 ? std::sqrt(diff) : 0.0;
    cos_big_c = cos_c > 0.0 ? sin2cos(sin_big_c) : -sin2cos(sin_big_c);
  } else {
    sin_big_c = +std::fabs(sin_big_b) * sin_c / sin_b;
    cos_big_c = (cos_c - cos_a * cos_b) / (sin_a * sin_b);
  }
  if (sin_big_c == 0.0) sin_big_c = std::numeric_limits<double>::min();
  l.cost = cos_b;
  l.sint = sin_b;
  l.cosp = cos_big_c;
  l.sinp = sin_big_b < 0.0 ? sin_big_c : -sin_big_c;
  return l;
}

// type image (image_type.f90:44-115); arrays in Fortran order, first index fastest
struct Image {
  hyp_image_conf c;
  int n_nu = 0, n_stokes = 4, n_orig = 1, n_sources = 0, n_dust = 0;
  double nu_min = 0, nu_max = 0, log10_nu_min = 0, log10_nu_max = 0, log10_ap_min = 0, log10_ap_max = 0;
  std::vector<double> img, img2, imgn, sed, sed2, sedn;
  std::vector<std::vector<double>> filt_nu, filt_tr;   // use_filters: one transmission curve per channel
  std::vector<double> filt_nu0;
  bool use_exact_nu = false;                // monochromatic mode (image_type.f90:243-258)
  std::vector<double> nu;                   // the frequencies of the channels, inu_min .. inu_max
  size_t img_index(int inu, int ix, int iy, int iv, int io, int is) const {
    return (size_t)(inu - 1) +
           (size_t)n_nu * ((ix - 1) + (size_t)c.n_x * ((iy - 1) + (size_t)c.n_y * ((iv - 1) + (size_t)c.n_view * ((io - 1) + (size_t)n_orig * is))));
  }
  size_t sed_index(int inu, int ir, int iv, int io, int is) const {
    return (size_t)(inu - 1) + (size_t)n_nu * ((ir - 1) + (size_t)c.n_ap * ((iv - 1) + (size_t)c.n_view * ((io - 1) + (size_t)n_orig * is)));
  }
};

struct PeeledState {
  std::vector<Image> image;                 // per group
  std::vector<int> group_id, view_id;       // per peeled view, 1-based
  std::vector<Angle> viewing_angles;
  std::vector<Vec> r_peeloff;               // per group
  // raytracing caches (images_peeled.f90:423-530), per (group, source / dust)
  std::vector<std::vector<std::vector<double>>> source_spectra, dust_extinction;
  std::vector<std::vector<std::vector<double>>> dust_log10_emissivity;  // [ig][id][ijnu*n_nu + inu]
};

// image_setup (image_type.f90:153-335)
void image_setup(Image &im, const hyp_image_conf &c, int n_sources, int n_dust, const std::vector<double> &frequencies) {
  im.c = c;
  im.n_sources = n_sources;
  im.n_dust = n_dust;
  im.n_nu = c.n_wav;
  if (im.n_nu < 1) throw OracleError{"n_nu should be >= 1"};
  im.n_stokes = c.compute_stokes ? 4 : 1;
  if (c.compute_sed) {
    im.log10_ap_min = std::log10(c.ap_min);
    im.log10_ap_max = std::log10(c.ap_max);
  }
  switch (c.track_origin) {
    case HYP_TRACK_SCATTERINGS: im.n_orig = 4 + 2 * c.track_n_scat; break;
    case HYP_TRACK_DETAILED: im.n_orig = 2 * (n_sources + n_dust); break;
    case HYP_TRACK_BASIC: im.n_orig = 4; break;
    case HYP_TRACK_NO: im.n_orig = 1; break;
    default: throw OracleError{"unknown track_origin flag"};
  }
  if (c.use_filters) {
    // image_type.f90:274-284
    if (!c.filt_n || !c.filt_nu || !c.filt_tr || !c.filt_nu0) throw OracleError{"filter tables are missing"};
    size_t off = 0;
    for (int i = 0; i < im.n_nu; i++) {
      im.filt_nu.emplace_back(c.filt_nu + off, c.filt_nu + off + c.filt_n[i]);
      im.filt_tr.emplace_back(c.filt_tr + off, c.filt_tr + off + c.filt_n[i]);
      im.filt_nu0.push_back(c.filt_nu0[i]);
      off += c.filt_n[i];
    }
    im.c.filt_n = nullptr;
    im.c.filt_nu = im.c.filt_tr = im.c.filt_nu0 = nullptr;
  }
  const double c_cgs = 2.99792458e10;
  // "1.e-4" is a default-real (single precision) literal in image_type.f90:262-263
  const double micron = (double)1.e-4f;
  if (c.inu_min > 0) {
    // image_type.f90:243-258
    if (c.use_filters) throw OracleError{"cannot use filters in monochromatic mode"};
    const int nf = (int)frequencies.size();
    if (c.inu_min < 1 || c.inu_min > nf) throw OracleError{"inu_min value is out of range"};
    if (c.inu_max < 1 || c.inu_max > nf) throw OracleError{"inu_max value is out of range"};
    im.use_exact_nu = true;
    im.nu.assign(frequencies.begin() + (c.inu_min - 1), frequencies.begin() + c.inu_max);
    if (im.n_nu != (int)im.nu.size()) throw OracleError{"n_nu should match length of frequencies array"};
  } else if (!c.use_filters) {
    im.nu_min = c_cgs / (c.wav_max * micron);
    im.nu_max = c_cgs / (c.wav_min * micron);
    im.log10_nu_min = std::log10(im.nu_min);
    im.log10_nu_max = std::log10(im.nu_max);
  }
  if (c.compute_image) {
    size_t n = (size_t)im.n_nu * c.n_x * c.n_y * c.n_view * im.n_orig * im.n_stokes;
    im.img.assign(n, 0.0);
    if (c.uncertainties) {
      im.img2.assign(n, 0.0);
      im.imgn.assign(n, 0.0);
    }
  }
  if (c.compute_sed) {
    size_t n = (size_t)im.n_nu * c.n_ap * c.n_view * im.n_orig * im.n_stokes;
    im.sed.assign(n, 0.0);
    if (c.uncertainties) {
      im.sed2.assign(n, 0.0);
      im.sedn.assign(n, 0.0);
    }
  }
  if (c.io_bytes != 4 && c.io_bytes != 8) throw OracleError{"unexpected value of io_bytes (should be 4 or 8)"};
}

// orig (image_type.f90:117-134)
int orig(const Photon &p) {
  if (p.scattered) return p.reprocessed ? 4 : 3;
  return p.reprocessed ? 2 : 1;
}

// origin slice shared by image_bin / image_bin_raytraced (image_type.f90:440-460)
int origin_slice(const Image &im, const Photon &p) {
  if (im.c.track_origin == HYP_TRACK_DETAILED) {
    int iorig = orig(p);
    // mod(iorig,2), mod(iorig+1,2) as in the Fortran
    int io = ((iorig - iorig % 2) * im.n_sources + (iorig - (iorig + 1) % 2 - 1) * im.n_dust) / 2;
    if (iorig % 2 == 0)
      io = io + p.dust_id;
    else
      io = io + p.source_id;
    return io;
  }
  if (im.c.track_origin == HYP_TRACK_SCATTERINGS) {
    int io = p.n_scat > im.c.track_n_scat ? im.c.track_n_scat + 2 : p.n_scat + 1;
    if (p.reprocessed) io = io + (im.c.track_n_scat + 2);
    return io;
  }
  if (im.c.track_origin == HYP_TRACK_BASIC) return orig(p);
  return 1;
}

// find_sed_bin (image_type.f90:337-354)
int find_sed_bin(const Image &im, double x_image, double y_image) {
  double log10_r = std::log10(std::sqrt(x_image * x_image + y_image * y_image));
  if (log10_r < im.log10_ap_min || im.c.n_ap == 1) return 1;
  return ipos(im.log10_ap_min, im.log10_ap_max, log10_r, im.c.n_ap - 1) + 1;
}

// in_image (image_type.f90:369-406)
bool in_image(const Image &im, double x, double y) {
  const hyp_image_conf &c = im.c;
  if (c.compute_image) {
    if ((x >= c.x_min && x <= c.x_max) || (x <= c.x_min && x >= c.x_max))
      if ((y >= c.y_min && y <= c.y_max) || (y <= c.y_min && y >= c.y_max)) return true;
  }
  if (c.compute_sed) {
    if (x * x + y * y <= c.ap_max * c.ap_max) return true;
  }
  return false;
}

// image_bin_single (image_type.f90:478-524)
void image_bin_single(Image &im, const Photon &p, double x_image, double y_image, int iv, int inu, int io,
                      double transmission);

// image_bin (image_type.f90:408-476)
void image_bin(Image &im, const Photon &p, double x_image, double y_image, int iv) {
  if (std::isnan(p.energy) || std::isnan(p.s.I)) return;
  int io = origin_slice(im, p);
  if (im.c.use_filters) {
    for (int ifilt = 1; ifilt <= im.n_nu; ifilt++) {
      const auto &fn = im.filt_nu[ifilt - 1];
      double transmission = interp1d_lin(fn.data(), im.filt_tr[ifilt - 1].data(), (int)fn.size(), p.nu, false, 0.0);
      if (transmission > 0.0) image_bin_single(im, p, x_image, y_image, iv, ifilt, io, transmission);
    }
  } else if (im.use_exact_nu) {
    image_bin_single(im, p, x_image, y_image, iv, p.inu - im.c.inu_min + 1, io, 1.0);   // image_type.f90:435-436
  } else {
    int inu = ipos(im.log10_nu_min, im.log10_nu_max, std::log10(p.nu), im.n_nu);
    image_bin_single(im, p, x_image, y_image, iv, inu, io, 1.0);
  }
}

void image_bin_single(Image &im, const Photon &p, double x_image, double y_image, int iv, int inu, int io,
                      double transmission) {
  const double st[4] = {p.s.I, p.s.Q, p.s.U, p.s.V};
  if (inu >= 1 && inu <= im.n_nu) {
    if (im.c.compute_image) {
      int ix = ipos(im.c.x_min, im.c.x_max, x_image, im.c.n_x);
      int iy = ipos(im.c.y_min, im.c.y_max, y_image, im.c.n_y);
      if (ix >= 1 && ix <= im.c.n_x && iy >= 1 && iy <= im.c.n_y) {
        for (int is = 0; is < im.n_stokes; is++) {
          size_t k = im.img_index(inu, ix, iy, iv, io, is);
          double v = st[is] * p.energy * transmission;
          im.img[k] = im.img[k] + v;
          if (im.c.uncertainties) {
            im.img2[k] = im.img2[k] + v * v;
            im.imgn[k] = im.imgn[k] + 1.0;
          }
        }
      }
    }
    if (im.c.compute_sed) {
      int ir = find_sed_bin(im, x_image, y_image);
      if (ir >= 1 && ir <= im.c.n_ap) {
        for (int is = 0; is < im.n_stokes; is++) {
          size_t k = im.sed_index(inu, ir, iv, io, is);
          double v = st[is] * p.energy * transmission;
          im.sed[k] = im.sed[k] + v;
          if (im.c.uncertainties) {
            im.sed2[k] = im.sed2[k] + v * v;
            im.sedn[k] = im.sedn[k] + 1.0;
          }
        }
      }
    }
  }
}

// image_bin_raytraced (image_type.f90:526-606)
void image_bin_raytraced(Image &im, const Photon &p, double x_image, double y_image, int iv,
                         const std::vector<double> &spectrum) {
  if (std::isnan(p.energy) || std::isnan(p.s.I)) return;
  int io = origin_slice(im, p);
  if (im.c.compute_image) {
    int ix = ipos(im.c.x_min, im.c.x_max, x_image, im.c.n_x);
    int iy = ipos(im.c.y_min, im.c.y_max, y_image, im.c.n_y);
    if (ix >= 1 && ix <= im.c.n_x && iy >= 1 && iy <= im.c.n_y) {
      for (int iw = 1; iw <= im.n_nu; iw++) {
        size_t k = im.img_index(iw, ix, iy, iv, io, 0);
        im.img[k] = im.img[k] + spectrum[iw - 1];
        if (im.c.uncertainties) {
          im.img2[k] = im.img2[k] + spectrum[iw - 1] * spectrum[iw - 1];
          im.imgn[k] = im.imgn[k] + 1.0;
        }
      }
    }
  }
  if (im.c.compute_sed) {
    int ir = find_sed_bin(im, x_image, y_image);
    if (ir >= 1 && ir <= im.c.n_ap) {
      for (int iw = 1; iw <= im.n_nu; iw++) {
        size_t k = im.sed_index(iw, ir, iv, io, 0);
        im.sed[k] = im.sed[k] + spectrum[iw - 1];
        if (im.c.uncertainties) {
          im.sed2[k] = im.sed2[k] + spectrum[iw - 1] * spectrum[iw - 1];
          im.sedn[k] = im.sedn[k] + 1.0;
        }
      }
    }
  }
}

// image_scale (image_type.f90:136-151)
void image_scale(Image &im, double scale) {
  for (auto &v : im.img) v = v * scale;
  for (auto &v : im.sed) v = v * scale;
  for (auto &v : im.img2) v = v * (scale * scale);
  for (auto &v : im.sed2) v = v * (scale * scale);
}

}  // namespace

// ---------------------------------------------------------------------------
// context: the module-level state of the Fortran program
// ---------------------------------------------------------------------------
struct orc_ctx {
  Rng rng;
  hyp_run_conf conf;
  // geometry (grid_geometry_cartesian_3d.f90:77-135)
  int n1 = 0, n2 = 0, n3 = 0, n_cells = 0;
  std::vector<double> w1, w2, w3, ew1, ew2, ew3, dx, dy, dz, volume;
  // find_wall scratch (:29-30)
  double tmin = 0, emin = 0;
  WallId imin, iext;
  // geometry kind: 0 Cartesian, 1 spherical polar (grid_geometry_spherical_3d.f90),
  // 2 cylindrical polar (grid_geometry_cylindrical_3d.f90), 3 octree (grid_geometry_octree.f90),
  // 4 AMR (grid_geometry_amr.f90)
  int grid_type = 0;
  bool radial = false;  // the 'radial' argument of the spherical find_wall (grid_propagate_3d.f90:73)
  std::vector<double> wr2, wtanp, wtant, wcost, wsint, wtant2;  // (:169-185)
  int midplane = -1;
  // octree (grid_geometry_octree.f90:160-260): one entry per node, 1-based ids stored 0-based
  std::vector<double> ox, oy, oz, odx, ody, odz;
  std::vector<char> orefined;
  std::vector<int> ochildren, oparent, oparent_subcell;  // ochildren[8*ic + k]
  double oct_eps = 0.0;
  // AMR: levels of grids (type_grid_amr.f90); goto arrays are (0:n1+1, 0:n2+1, 0:n3+1), first index fastest
  struct AmrGrid {
    int n1, n2, n3, n_cells, start_id;
    double xmin, xmax, ymin, ymax, zmin, zmax, width[3], volume;
    std::vector<double> w1, w2, w3;
    std::vector<int> goto_grid, goto_level;
    size_t gidx(int i1, int i2, int i3) const { return (size_t)i1 + (size_t)(n1 + 2) * (i2 + (size_t)(n2 + 2) * i3); }
  };
  std::vector<std::vector<AmrGrid>> levels;
  std::vector<int> cell_ilevel, cell_igrid, cell_i1, cell_i2, cell_i3;  // preset_cell_id
  double amr_eps = 0.0;
  // Voronoi mesh (grid_geometry_voronoi.f90, type_grid_voronoi.f90): sites, neighbour lists in the file's C-style
  // numbering (>= 0: cell, -1 .. -6: the walls of the box), bounding boxes, the box; grid_type 5
  std::vector<double> vx, vy, vz, vbb;      // vbb: xmin, xmax, ymin, ymax, zmin, zmax per cell
  std::vector<int> vidx, vneigh;
  double vbox[6] = {0, 0, 0, 0, 0, 0};
  // nearest-site search (the reference's kd-tree): sites bucketed on a uniform grid over the box
  int vg[3] = {1, 1, 1};
  std::vector<int> vg_start, vg_sites;
  // cells that hold physical quantities (geo%mask / mask_map); empty = all cells
  std::vector<int> mask_map;
  int n_masked = 0;
  bool specific_energy_from_file = false;
  // modified random walk (grid_mrw_3d.f90:21-26, grid_physics_3d.f90:60)
  std::vector<double> alpha_inv_planck, diff_coeff;
  double mrw_xcdf[100], mrw_ycdf[100];
  int64_t n_mrw_steps = 0;
  // dust
  int n_dust = 0;
  std::vector<Dust> d;
  // grid physics (grid_physics_3d.f90:34-63): (ic, id) stored at [id*n_cells + ic]
  std::vector<double> density, specific_energy, specific_energy_sum, jnu_var_frac, minimum_specific_energy;
  std::vector<double> specific_energy_additional;  // specific_energy_type = 'additional' (grid_physics_3d.f90:55)
  // frequency-resolved specific energy (grid_physics_3d.f90:41-56): [bin][dust][cell], the file's order;
  // empty unless bin edges were given (compute_specific_energy_spectrum)
  std::vector<double> nu_bin_edges, log_nu_bin_edges, specific_energy_spectrum, specific_energy_sum_spectrum;
  std::vector<double> j_nu_bin_frac;               // [dust][state][bin] (setup_j_nu_bin_fractions, :325-348)
  int n_nu_bins = 0, n_jnu_max = 0;
  std::vector<int> jnu_var_id;
  std::vector<double> energy_abs_tot;
  PdfDiscrete absorption;
  // sources (source.f90:25-44)
  std::vector<Source> s;
  PdfDiscrete luminosity;
  double energy_total = 0, energy_current = 0;
  // monochromatic mode (settings.f90:29-31)
  std::vector<double> frequencies;
  double monochromatic_energy_threshold = 1e-10;
  std::vector<double> mono_mean_prob;            // grid_monochromatic.f90:35-36
  std::vector<PdfDiscrete> mono_emiss_pdf;
  bool mono_run = false;
  bool any_intersect = false;
  // number of packets that visited each cell (grid_physics_3d.f90:38-39,308-317), kept when the partial diffusion
  // approximation or the n_photons output needs it
  std::vector<int64_t> n_photons, last_photon_id;
  int64_t photon_counter = 0;
  int pda_exact_limit = 10000;   // grid_pda_3d.f90:126: below this many PDA cells the exact solver is used
  // counters
  int64_t killed_photons_geo = 0, killed_photons_int = 0;
  int64_t n_crossings = 0, n_absorptions = 0, n_scatterings = 0, n_escaped = 0, n_photons_run = 0;
  int64_t n_peel_crossings = 0, n_peeloffs = 0, n_reabsorptions = 0;
  PeeledState peeled;
  bool setup_done = false;
  std::string error;
};

namespace {

thread_local std::string g_last_error;

double spacing(double x) {
  // Fortran SPACING(x): distance to the next representable number of larger magnitude
  if (x == 0.0) return std::numeric_limits<double>::min();
  double ax = std::fabs(x);
  return std::nextafter(ax, std::numeric_limits<double>::infinity()) - ax;
}

// new_grid_cell_3d (type_cell_id_3d.f90:67-75,97-102)
Cell new_grid_cell(const orc_ctx &g, int i1, int i2, int i3) {
  Cell c;
  c.i1 = i1;
  c.i2 = i2;
  c.i3 = i3;
  c.ic = (i3 - 1) * g.n1 * g.n2 + (i2 - 1) * g.n1 + i1;
  return c;
}

// escaped_cell (grid_geometry_cartesian_3d.f90:267-275)
bool escaped(const orc_ctx &g, const Cell &c) {
  if (g.grid_type == 4) return c.ic == -2;  // outside_cell (grid_geometry_amr.f90:592-597)
  if (g.grid_type == 3) return c.ic == g.n_cells + 1;  // grid_geometry_octree.f90:318-325
  if (g.grid_type == 5) return c.ic == g.n_cells + 1;  // grid_geometry_voronoi.f90:249-255
  if (c.i1 < 1 || c.i1 > g.n1) return true;
  if (g.grid_type == 1) return false;  // spherical: radial escape only (grid_geometry_spherical_3d.f90:493-500)
  if (g.grid_type == 2) return c.i2 < 1 || c.i2 > g.n2;  // cylindrical: w and z (grid_geometry_cylindrical_3d.f90:375-384)
  if (c.i2 < 1 || c.i2 > g.n2) return true;
  if (c.i3 < 1 || c.i3 > g.n3) return true;
  return false;
}

// find_cell (grid_geometry_cartesian_3d.f90:143-167); returns false for invalid_cell
bool sph_find_cell(const orc_ctx &g, const Photon &p, Cell &out);
void sph_adjust_wall(const orc_ctx &g, Photon &p);
bool sph_in_correct_cell(const orc_ctx &g, const Photon &p);
void sph_find_wall(orc_ctx &g, const Photon &p, double &tnearest, WallId &id_min);
bool cyl_find_cell(const orc_ctx &g, const Photon &p, Cell &out);
bool oct_find_cell(const orc_ctx &g, const Photon &p, Cell &out);
bool amr_find_cell(const orc_ctx &g, const Photon &p, Cell &out);
bool amr_in_correct_cell(const orc_ctx &g, const Photon &p);
void amr_find_wall(orc_ctx &g, const Photon &p, double &tnearest, WallId &id_min);
Cell amr_next_cell(const orc_ctx &g, const Cell &c, const WallId &dir, const Vec &r);
bool oct_in_correct_cell(const orc_ctx &g, const Photon &p);
void oct_find_wall(orc_ctx &g, const Photon &p, double &tnearest, WallId &id_min);
Cell oct_next_cell(const orc_ctx &g, const Cell &c, const WallId &dir, const Vec &r);
void cyl_adjust_wall(const orc_ctx &g, Photon &p);
bool cyl_in_correct_cell(const orc_ctx &g, const Photon &p);
void cyl_find_wall(orc_ctx &g, const Photon &p, double &tnearest, WallId &id_min);

bool vor_find_cell(const orc_ctx &g, const Photon &p, Cell &out);
bool vor_in_correct_cell(const orc_ctx &g, const Photon &p);
void vor_find_wall(orc_ctx &g, const Photon &p, double &tnearest, WallId &id_min);

bool find_cell(const orc_ctx &g, const Photon &p, Cell &out) {
  if (g.grid_type == 5) return vor_find_cell(g, p, out);
  if (g.grid_type == 1) return sph_find_cell(g, p, out);
  if (g.grid_type == 2) return cyl_find_cell(g, p, out);
  if (g.grid_type == 3) return oct_find_cell(g, p, out);
  if (g.grid_type == 4) return amr_find_cell(g, p, out);
  int i1 = locate(g.w1.data(), g.n1 + 1, p.r.x);
  int i2 = locate(g.w2.data(), g.n2 + 1, p.r.y);
  int i3 = locate(g.w3.data(), g.n3 + 1, p.r.z);
  if (i1 < 1 || i1 > g.n1) return false;
  if (i2 < 1 || i2 > g.n2) return false;
  if (i3 < 1 || i3 > g.n3) return false;
  out = new_grid_cell(g, i1, i2, i3);
  return true;
}

// adjust_wall (grid_geometry_cartesian_3d.f90:169-237)
void adjust_wall(const orc_ctx &g, Photon &p) {
  if (g.grid_type == 1) {
    sph_adjust_wall(g, p);
    return;
  }
  if (g.grid_type == 2) {
    cyl_adjust_wall(g, p);
    return;
  }
  if (g.grid_type >= 3) return;  // octree / AMR / Voronoi place_in_cell have no adjust_wall
  p.on_wall = false;
  p.on_wall_id = WallId();
#define ADJ(V, R, W, I, WID)                   \
  if (V > 0.0) {                               \
    if (R == W[I - 1]) {                       \
      WID = -1;                                \
    } else if (R == W[I]) {                    \
      WID = -1;                                \
      I = I + 1;                               \
    }                                          \
  } else if (V < 0.0) {                        \
    if (R == W[I - 1]) {                       \
      WID = +1;                                \
      I = I - 1;                               \
    } else if (R == W[I]) {                    \
      WID = +1;                                \
    }                                          \
  }
  ADJ(p.v.x, p.r.x, g.w1, p.icell.i1, p.on_wall_id.w1)
  ADJ(p.v.y, p.r.y, g.w2, p.icell.i2, p.on_wall_id.w2)
  ADJ(p.v.z, p.r.z, g.w3, p.icell.i3, p.on_wall_id.w3)
#undef ADJ
  // NB: the reference does not refresh icell%ic here (only i1/i2/i3 change).
  p.on_wall = p.on_wall_id.w1 != 0 || p.on_wall_id.w2 != 0 || p.on_wall_id.w3 != 0;
}

// place_in_cell (grid_geometry_cartesian_3d.f90:239-259)
void place_in_cell(orc_ctx &g, Photon &p) {
  Cell c;
  if (!find_cell(g, p, c)) {
    g.killed_photons_geo++;
    p.killed = true;
  } else {
    p.icell = c;
    p.in_cell = true;
    adjust_wall(g, p);
  }
}

// in_correct_cell (grid_geometry_cartesian_3d.f90:330-381)
bool in_correct_cell(const orc_ctx &g, const Photon &p) {
  if (g.grid_type == 1) return sph_in_correct_cell(g, p);
  if (g.grid_type == 2) return cyl_in_correct_cell(g, p);
  if (g.grid_type == 3) return oct_in_correct_cell(g, p);
  if (g.grid_type == 4) return amr_in_correct_cell(g, p);
  if (g.grid_type == 5) return vor_in_correct_cell(g, p);
  const double threshold = 1.e-3;
  Cell act;
  bool valid = find_cell(g, p, act);
  if (!valid) act = Cell{-1, -1, -1, -1};  // invalid_cell
  if (p.on_wall) {
    bool ok = true;
    double frac;
#define CHK(WID, R, W, I, IA)                              \
  if (WID == -1) {                                         \
    frac = (R - W[I - 1]) / (W[I] - W[I - 1]);             \
    ok = ok && std::fabs(frac) < threshold;                \
  } else if (WID == +1) {                                  \
    frac = (R - W[I]) / (W[I] - W[I - 1]);                 \
    ok = ok && std::fabs(frac) < threshold;                \
  } else {                                                 \
    ok = ok && IA == I;                                    \
  }
    CHK(p.on_wall_id.w1, p.r.x, g.w1, p.icell.i1, act.i1)
    CHK(p.on_wall_id.w2, p.r.y, g.w2, p.icell.i2, act.i2)
    CHK(p.on_wall_id.w3, p.r.z, g.w3, p.icell.i3, act.i3)
#undef CHK
    return ok;
  }
  return act.i1 == p.icell.i1 && act.i2 == p.icell.i2 && act.i3 == p.icell.i3;
}

// insert_t (grid_geometry_cartesian_3d.f90:482-512)
inline void insert_t(orc_ctx &g, double t, int iw, int i, double e) {
  if (t > 0.0) {
    double emax = e > g.emin ? e : g.emin;
    if (t < g.tmin - emax) {
      g.tmin = t;
      g.imin = WallId();
      g.emin = emax;
      if (iw == 1)
        g.imin.w1 = i;
      else if (iw == 2)
        g.imin.w2 = i;
      else
        g.imin.w3 = i;
    } else if (t < g.tmin + emax) {
      g.emin = emax;
      if (iw == 1)
        g.imin.w1 = i;
      else if (iw == 2)
        g.imin.w2 = i;
      else
        g.imin.w3 = i;
    }
  }
}

// find_wall (grid_geometry_cartesian_3d.f90:424-472), reset_t (:474-480), find_next_wall (:514-521)
void find_wall(orc_ctx &g, const Photon &p, double &tnearest, WallId &id_min) {
  if (g.grid_type == 1) {
    sph_find_wall(g, p, tnearest, id_min);
    return;
  }
  if (g.grid_type == 2) {
    cyl_find_wall(g, p, tnearest, id_min);
    return;
  }
  if (g.grid_type == 3) {
    oct_find_wall(g, p, tnearest, id_min);
    return;
  }
  if (g.grid_type == 4) {
    amr_find_wall(g, p, tnearest, id_min);
    return;
  }
  if (g.grid_type == 5) {
    vor_find_wall(g, p, tnearest, id_min);
    return;
  }
  g.tmin = std::numeric_limits<double>::max();
  g.emin = 0.0;
  g.imin = WallId();
  if (p.on_wall_id.w1 != -1) {
    double t1 = (g.w1[p.icell.i1 - 1] - p.r.x) / p.v.x;
    insert_t(g, t1, 1, -1, g.ew1[p.icell.i1 - 1]);
  }
  if (p.on_wall_id.w1 != +1) {
    double t2 = (g.w1[p.icell.i1] - p.r.x) / p.v.x;
    insert_t(g, t2, 1, +1, g.ew1[p.icell.i1]);
  }
  if (p.on_wall_id.w2 != -1) {
    double t1 = (g.w2[p.icell.i2 - 1] - p.r.y) / p.v.y;
    insert_t(g, t1, 2, -1, g.ew2[p.icell.i2 - 1]);
  }
  if (p.on_wall_id.w2 != +1) {
    double t2 = (g.w2[p.icell.i2] - p.r.y) / p.v.y;
    insert_t(g, t2, 2, +1, g.ew2[p.icell.i2]);
  }
  if (p.on_wall_id.w3 != -1) {
    double t1 = (g.w3[p.icell.i3 - 1] - p.r.z) / p.v.z;
    insert_t(g, t1, 3, -1, g.ew3[p.icell.i3 - 1]);
  }
  if (p.on_wall_id.w3 != +1) {
    double t2 = (g.w3[p.icell.i3] - p.r.z) / p.v.z;
    insert_t(g, t2, 3, +1, g.ew3[p.icell.i3]);
  }
  tnearest = g.tmin;
  id_min = g.imin;
}

// next_cell_wall_id (grid_geometry_cartesian_3d.f90:303-328)
Cell next_cell(const orc_ctx &g, const Cell &c, const WallId &dir, const Vec &r) {
  // next_cell_wall_id (grid_geometry_voronoi.f90:266-272): the wall id is the id of the cell behind it
  if (g.grid_type == 5) return Cell{0, 0, 0, dir.w1};
  if (g.grid_type == 3) return oct_next_cell(g, c, dir, r);
  if (g.grid_type == 4) return amr_next_cell(g, c, dir, r);
  int i1 = c.i1, i2 = c.i2, i3 = c.i3;
  if (dir.w1 == -1)
    i1 = i1 - 1;
  else if (dir.w1 == +1)
    i1 = i1 + 1;
  if (dir.w2 == -1)
    i2 = i2 - 1;
  else if (dir.w2 == +1)
    i2 = i2 + 1;
  if (dir.w3 == -1)
    i3 = i3 - 1;
  else if (dir.w3 == +1)
    i3 = i3 + 1;
  if (g.grid_type == 1 || g.grid_type == 2) {  // phi is periodic (grid_geometry_spherical_3d.f90:549-555)
    if (i3 == 0) i3 = g.n3;
    if (i3 == g.n3 + 1) i3 = 1;
  }
  return new_grid_cell(g, i1, i2, i3);
}

// ---------------------------------------------------------------------------
// spherical polar geometry (src/grid/grid_geometry_spherical_3d.f90)
// ---------------------------------------------------------------------------
const double PI_F = 3.14159265358979323846;  // lib_constants.f90:66
const double TWOPI_F = PI_F + PI_F;

// equal_nulp (:48-57)
bool equal_nulp(double x, double y, int n) {
  if (x == y) return true;
  return std::fabs(x - y) <= n * spacing(std::max(x, y));
}

// theta / phi of a photon as find_cell and adjust_wall compute them (:224-245)
void sph_angles(const Photon &p, double &r_sq, double &w_sq, double &theta, double &phi) {
  r_sq = p.r.x * p.r.x + p.r.y * p.r.y + p.r.z * p.r.z;
  w_sq = p.r.x * p.r.x + p.r.y * p.r.y;
  if (r_sq == 0.0)
    theta = std::atan2(std::sqrt(p.v.x * p.v.x + p.v.y * p.v.y), p.v.z);
  else
    theta = std::atan2(std::sqrt(p.r.x * p.r.x + p.r.y * p.r.y), p.r.z);
  if (w_sq == 0.0) {
    phi = std::atan2(p.v.y, p.v.x);
    if (phi < 0.0) phi = phi + TWOPI_F;
  } else {
    phi = std::atan2(p.r.y, p.r.x);
    if (phi < 0.0) phi = phi + TWOPI_F;
  }
}

// find_cell (:212-276)
bool sph_find_cell(const orc_ctx &g, const Photon &p, Cell &out) {
  double r_sq, w_sq, theta, phi;
  sph_angles(p, r_sq, w_sq, theta, phi);
  int i1 = locate(g.wr2.data(), g.n1 + 1, r_sq);
  int i2 = locate(g.w2.data(), g.n2 + 1, theta);
  int i3 = locate(g.w3.data(), g.n3 + 1, phi);
  if (i1 < 1 || i1 > g.n1) return false;
  if (i2 < 1 || i2 > g.n2) return false;
  if (i3 < 1 || i3 > g.n3) return false;
  out = new_grid_cell(g, i1, i2, i3);
  return true;
}

// adjust_wall (:278-469)
void sph_adjust_wall(const orc_ctx &g, Photon &p) {
  const int eps = 3;
  p.on_wall = false;
  p.on_wall_id = WallId();
  double r_sq, w_sq, theta, phi;
  sph_angles(p, r_sq, w_sq, theta, phi);
  const double rdotv = p.r.x * p.v.x + p.r.y * p.v.y + p.r.z * p.v.z;
  // radial walls
  if (rdotv >= 0.0) {
    if (equal_nulp(r_sq, g.wr2[p.icell.i1 - 1], eps)) {
      p.on_wall_id.w1 = -1;
    } else if (equal_nulp(r_sq, g.wr2[p.icell.i1], eps)) {
      p.on_wall_id.w1 = -1;
      p.icell.i1 = p.icell.i1 + 1;
    }
  } else {
    if (equal_nulp(r_sq, g.wr2[p.icell.i1 - 1], eps)) {
      p.on_wall_id.w1 = +1;
      p.icell.i1 = p.icell.i1 - 1;
    } else if (equal_nulp(r_sq, g.wr2[p.icell.i1], eps)) {
      p.on_wall_id.w1 = +1;
    }
  }
  // theta walls
  if (r_sq == 0.0) {
    if (std::fabs(p.v.z) < 1.0) {
      double theta_v = std::atan2(std::sqrt(p.v.x * p.v.x + p.v.y * p.v.y), p.v.z);
      if (equal_nulp(theta_v, g.w2[p.icell.i2 - 1], eps))
        p.on_wall_id.w2 = -1;
      else if (equal_nulp(theta_v, g.w2[p.icell.i2], eps))
        p.on_wall_id.w2 = +1;
    }
  } else if (p.icell.i2 > 1 && equal_nulp(theta, g.w2[p.icell.i2 - 1], eps)) {
    if (p.icell.i2 == g.midplane) {
      if (p.v.z > 0.0) {
        p.on_wall_id.w2 = +1;
        p.icell.i2 = p.icell.i2 - 1;
      } else {
        p.on_wall_id.w2 = -1;
      }
    } else {
      bool lhs = (std::sqrt(w_sq) * p.v.z * g.wtant[p.icell.i2 - 1] - (p.r.x * p.v.x + p.r.y * p.v.y) < 0.0);
      if (lhs == (p.r.z > 0.0)) {
        p.on_wall_id.w2 = -1;
      } else {
        p.on_wall_id.w2 = +1;
        p.icell.i2 = p.icell.i2 - 1;
      }
    }
  } else if (p.icell.i2 + 1 < g.n2 + 1 && equal_nulp(theta, g.w2[p.icell.i2], eps)) {
    if (p.icell.i2 + 1 == g.midplane) {
      if (p.v.z > 0.0) {
        p.on_wall_id.w2 = +1;
      } else {
        p.on_wall_id.w2 = -1;
        p.icell.i2 = p.icell.i2 + 1;
      }
    } else {
      bool lhs = (std::sqrt(w_sq) * p.v.z * g.wtant[p.icell.i2] - (p.r.x * p.v.x + p.r.y * p.v.y) < 0.0);
      if (lhs == (p.r.z > 0.0)) {
        p.on_wall_id.w2 = -1;
        p.icell.i2 = p.icell.i2 + 1;
      } else {
        p.on_wall_id.w2 = +1;
      }
    }
  }
  // phi walls
  if (p.r.x == 0.0 && p.r.y == 0.0 && p.v.x == 0.0 && p.v.y == 0.0) {
    // on all phi walls at once: leave alone
  } else if (equal_nulp(phi, g.w3[p.icell.i3 - 1], eps)) {
    double phi_v = std::atan2(p.v.y, p.v.x);
    double dphi = phi_v - g.w3[p.icell.i3 - 1];
    if (dphi < -PI_F) dphi = dphi + TWOPI_F;
    if (dphi > 0.0) {
      p.on_wall_id.w3 = -1;
    } else {
      p.on_wall_id.w3 = +1;
      p.icell.i3 = p.icell.i3 - 1;
      if (p.icell.i3 == 0) p.icell.i3 = g.n3;
    }
  } else if (equal_nulp(phi, g.w3[p.icell.i3], eps)) {
    double phi_v = std::atan2(p.v.y, p.v.x);
    double dphi = phi_v - g.w3[p.icell.i3];
    if (dphi < -PI_F) dphi = dphi + TWOPI_F;
    if (dphi > 0.0) {
      p.on_wall_id.w3 = -1;
      p.icell.i3 = p.icell.i3 + 1;
      if (p.icell.i3 == g.n3 + 1) p.icell.i3 = 1;
    } else {
      p.on_wall_id.w3 = +1;
    }
  }
  p.on_wall = p.on_wall_id.w1 != 0 || p.on_wall_id.w2 != 0 || p.on_wall_id.w3 != 0;
}

// in_correct_cell (:557-643)
bool sph_in_correct_cell(const orc_ctx &g, const Photon &p) {
  const double threshold = 1.e-3;
  Cell act;
  bool valid = sph_find_cell(g, p, act);
  if (!valid) act = Cell{-1, -1, -1, -1};
  if (!p.on_wall) return act.i1 == p.icell.i1 && act.i2 == p.icell.i2 && act.i3 == p.icell.i3;
  bool ok = true;
  double r_sq, w_sq, theta, phi, frac, dphi;
  r_sq = p.r.x * p.r.x + p.r.y * p.r.y + p.r.z * p.r.z;
  if (r_sq == 0.0) return true;
  sph_angles(p, r_sq, w_sq, theta, phi);
  if (p.on_wall_id.w1 == -1) {
    if (g.w1[p.icell.i1 - 1] != std::sqrt(r_sq)) {
      frac = std::sqrt(r_sq) / g.w1[p.icell.i1 - 1] - 1.0;
      ok = ok && std::fabs(frac) < threshold;
    }
  } else if (p.on_wall_id.w1 == +1) {
    if (g.w1[p.icell.i1] != std::sqrt(r_sq)) {
      frac = std::sqrt(r_sq) / g.w1[p.icell.i1] - 1.0;
      ok = ok && std::fabs(frac) < threshold;
    }
  } else {
    ok = ok && act.i1 == p.icell.i1;
  }
  if (p.on_wall_id.w2 == -1) {
    frac = theta / g.w2[p.icell.i2 - 1] - 1.0;
    ok = ok && std::fabs(frac) < threshold;
  } else if (p.on_wall_id.w2 == +1) {
    frac = theta / g.w2[p.icell.i2] - 1.0;
    ok = ok && std::fabs(frac) < threshold;
  } else {
    ok = ok && act.i2 == p.icell.i2;
  }
  if (p.on_wall_id.w3 == -1) {
    dphi = phi - g.w3[p.icell.i3 - 1];
    if (dphi > PI_F) dphi = dphi - TWOPI_F;
    if (dphi < -PI_F) dphi = dphi + TWOPI_F;
    frac = dphi / (g.w3[p.icell.i3] - g.w3[p.icell.i3 - 1]);
    ok = ok && std::fabs(frac) < threshold;
  } else if (p.on_wall_id.w3 == +1) {
    dphi = phi - g.w3[p.icell.i3];
    if (dphi > PI_F) dphi = dphi - TWOPI_F;
    if (dphi < -PI_F) dphi = dphi + TWOPI_F;
    frac = dphi / (g.w3[p.icell.i3] - g.w3[p.icell.i3 - 1]);
    ok = ok && std::fabs(frac) < threshold;
  } else {
    ok = ok && act.i3 == p.icell.i3;
  }
  return ok;
}

// quadratic_dp (fortranlib/src/lib_algebra.f90:107-122)
void quadratic(double a, double b, double c, double &x1, double &x2) {
  const double huge = std::numeric_limits<double>::max();
  double delta = b * b - 4.0 * a * c;
  if (delta > 0) {
    delta = std::sqrt(delta);
    double factor = 0.5 / a;
    x1 = (-b - delta) * factor;
    x2 = (-b + delta) * factor;
  } else {
    x1 = huge;
    x2 = huge;
  }
}

void quadratic_pascal_reduced(double b, double c, double &x1, double &x2);

// one cone wall of find_wall (:822-962); iw2 is the wall index (1-based), side -1 lower / +1 upper
void sph_cone_wall(orc_ctx &g, const Photon &p, int iw2, int side, double v2_xy, double v2_z, double rv_xy,
                   double rv_z, double r2_xy, double r2_z) {
  const double huge = std::numeric_limits<double>::max();
  const double wtant = g.wtant[iw2 - 1], wtant2 = g.wtant2[iw2 - 1], ew = g.ew2[iw2 - 1];
  if (p.on_wall_id.w2 == side && equal_nulp(wtant, std::sqrt(v2_xy) / p.v.z, 10) &&
      equal_nulp(std::sqrt(r2_xy) * p.v.z * wtant, rv_xy, 10)) {
    g.iext.w2 = side;
    return;
  }
  if (iw2 == g.midplane && p.v.z != 0) {
    if (p.on_wall_id.w2 != side) insert_t(g, -p.r.z / p.v.z, 2, side, ew);
    return;
  }
  double pA = v2_xy - v2_z * wtant2;
  double pB = rv_xy - rv_z * wtant2;
  pB = pB + pB;
  double pC = r2_xy - r2_z * wtant2;
  if (std::fabs(pA) > 0.0) {
    double t1, t2;
    quadratic(pA, pB, pC, t1, t2);
    double z1 = p.r.z + p.v.z * t1;
    if ((z1 > 0.0) != (wtant > 0.0)) t1 = huge;
    double z2 = p.r.z + p.v.z * t2;
    if ((z2 > 0.0) != (wtant > 0.0)) t2 = huge;
    if (p.on_wall_id.w2 == side) {
      if (std::fabs(t1) < std::fabs(t2))
        insert_t(g, t2, 2, side, ew);
      else
        insert_t(g, t1, 2, side, ew);
    } else {
      insert_t(g, t1, 2, side, ew);
      insert_t(g, t2, 2, side, ew);
    }
  } else if (std::fabs(pB) > 0.0) {
    if (p.on_wall_id.w2 != side) insert_t(g, -pC / pB, 2, side, ew);
  }
}

// find_wall (:741-1073), reset_t (:1075-1081), find_next_wall (:1113-1120)
void sph_find_wall(orc_ctx &g, const Photon &p, double &tnearest, WallId &id_min) {
  g.tmin = std::numeric_limits<double>::max();
  g.emin = 0.0;
  g.imin = WallId();
  g.iext = WallId();
  double v2_xy = p.v.x * p.v.x + p.v.y * p.v.y;
  double v2_z = p.v.z * p.v.z;
  double rv_xy = p.r.x * p.v.x + p.r.y * p.v.y;
  double rv_z = p.r.z * p.v.z;
  double r2_xy = p.r.x * p.r.x + p.r.y * p.r.y;
  double r2_z = p.r.z * p.r.z;
  double pB = rv_xy + rv_z;
  pB = pB + pB;
  double pC = r2_xy + r2_z;
  double t1, t2;
  // spherical walls
  if (!g.radial) {
    double pC_1 = pC - g.wr2[p.icell.i1 - 1];
    quadratic_pascal_reduced(pB, pC_1, t1, t2);
    if (p.on_wall_id.w1 == -1) {
      if (std::fabs(t1) < std::fabs(t2))
        insert_t(g, t2, 1, -1, g.ew1[p.icell.i1 - 1]);
      else
        insert_t(g, t1, 1, -1, g.ew1[p.icell.i1 - 1]);
    } else {
      insert_t(g, t1, 1, -1, g.ew1[p.icell.i1 - 1]);
      insert_t(g, t2, 1, -1, g.ew1[p.icell.i1 - 1]);
    }
  }
  double pC_2 = pC - g.wr2[p.icell.i1];
  quadratic_pascal_reduced(pB, pC_2, t1, t2);
  if (p.on_wall_id.w1 == +1) {
    if (std::fabs(t1) < std::fabs(t2))
      insert_t(g, t2, 1, +1, g.ew1[p.icell.i1]);
    else
      insert_t(g, t1, 1, +1, g.ew1[p.icell.i1]);
  } else {
    insert_t(g, t1, 1, +1, g.ew1[p.icell.i1]);
    insert_t(g, t2, 1, +1, g.ew1[p.icell.i1]);
  }
  // cone walls (theta = 0 and theta = pi are not walls)
  if (p.icell.i2 > 1) sph_cone_wall(g, p, p.icell.i2, -1, v2_xy, v2_z, rv_xy, rv_z, r2_xy, r2_z);
  if (p.icell.i2 < g.n2) sph_cone_wall(g, p, p.icell.i2 + 1, +1, v2_xy, v2_z, rv_xy, rv_z, r2_xy, r2_z);
  // phi walls
  if (g.n3 > 1) {
    double dphi = 0.0;
    if (p.on_wall_id.w3 == -1) {
      dphi = std::atan2(p.v.y, p.v.x) - g.w3[p.icell.i3 - 1];
      if (dphi > PI_F) dphi = dphi - TWOPI_F;
      if (dphi < -PI_F) dphi = dphi + TWOPI_F;
    }
    if (p.on_wall_id.w3 == +1) {
      dphi = std::atan2(p.v.y, p.v.x) - g.w3[p.icell.i3];
      if (dphi > PI_F) dphi = dphi - TWOPI_F;
      if (dphi < -PI_F) dphi = dphi + TWOPI_F;
    }
    if (p.on_wall_id.w3 == +1 && std::fabs(dphi) < g.ew3[p.icell.i3]) {
      g.iext.w3 = +1;
    } else if (p.on_wall_id.w3 == -1 && std::fabs(dphi) < g.ew3[p.icell.i3 - 1]) {
      g.iext.w3 = -1;
    } else if (r2_xy > 0.0) {
      if (p.on_wall_id.w3 != -1) {
        double tp = g.wtanp[p.icell.i3 - 1];
        t1 = -(tp * p.r.x - p.r.y) / (tp * p.v.x - p.v.y);
        double x_i = p.r.x + p.v.x * t1, y_i = p.r.y + p.v.y * t1;
        double phi_i = std::atan2(y_i, x_i);
        double dp = std::fabs(phi_i - g.w3[p.icell.i3 - 1]);
        if (dp > PI_F) dp = std::fabs(dp - TWOPI_F);
        if (dp < 0.5 * PI_F) insert_t(g, t1, 3, -1, 0.0);
      }
      if (p.on_wall_id.w3 != +1) {
        double tp = g.wtanp[p.icell.i3];
        t2 = -(tp * p.r.x - p.r.y) / (tp * p.v.x - p.v.y);
        double x_i = p.r.x + p.v.x * t2, y_i = p.r.y + p.v.y * t2;
        double phi_i = std::atan2(y_i, x_i);
        double dp = std::fabs(phi_i - g.w3[p.icell.i3]);
        if (dp > PI_F) dp = std::fabs(dp - TWOPI_F);
        if (dp < 0.5 * PI_F) insert_t(g, t2, 3, +1, 0.0);
      }
    }
  }
  tnearest = g.tmin;
  id_min = WallId{g.imin.w1 + g.iext.w1, g.imin.w2 + g.iext.w2, g.imin.w3 + g.iext.w3};
}

// random_position_cell (:645-677)
Vec sph_random_position_cell(orc_ctx &g, const Cell &c) {
  double r = g.rng.random(), t = g.rng.random(), ph = g.rng.random();
  double w1a = g.w1[c.i1 - 1], w1b = g.w1[c.i1];
  r = std::pow(r * (w1b * w1b * w1b - w1a * w1a * w1a) + w1a * w1a * w1a, 1.0 / 3.0);
  t = std::acos(t * (g.wcost[c.i2] - g.wcost[c.i2 - 1]) + g.wcost[c.i2 - 1]);
  ph = ph * (g.w3[c.i3] - g.w3[c.i3 - 1]) + g.w3[c.i3 - 1];
  if (r <= w1a || r >= w1b) r = 0.5 * (w1a + w1b);
  if (t <= g.w2[c.i2 - 1] || t >= g.w2[c.i2]) t = 0.5 * (g.w2[c.i2 - 1] + g.w2[c.i2]);
  if (ph <= g.w3[c.i3 - 1] || ph >= g.w3[c.i3]) ph = 0.5 * (g.w3[c.i3 - 1] + g.w3[c.i3]);
  return Vec{r * std::sin(t) * std::cos(ph), r * std::sin(t) * std::sin(ph), r * std::cos(t)};
}

// ---------------------------------------------------------------------------
// cylindrical polar geometry (src/grid/grid_geometry_cylindrical_3d.f90): walls w (cylinders), z (planes),
// phi (half-planes)
// ---------------------------------------------------------------------------
void cyl_phi(const Photon &p, double &w_sq, double &phi) {
  w_sq = p.r.x * p.r.x + p.r.y * p.r.y;
  if (w_sq == 0.0) {
    phi = std::atan2(p.v.y, p.v.x);
    if (phi < 0.0) phi = phi + TWOPI_F;
  } else {
    phi = std::atan2(p.r.y, p.r.x);
    if (phi < 0.0) phi = phi + TWOPI_F;
  }
}

// find_cell (:184-237)
bool cyl_find_cell(const orc_ctx &g, const Photon &p, Cell &out) {
  double w_sq, phi;
  cyl_phi(p, w_sq, phi);
  int i1 = locate(g.wr2.data(), g.n1 + 1, w_sq);
  int i2 = locate(g.w2.data(), g.n2 + 1, p.r.z);
  int i3 = locate(g.w3.data(), g.n3 + 1, phi);
  if (i1 < 1 || i1 > g.n1) return false;
  if (i2 < 1 || i2 > g.n2) return false;
  if (i3 < 1 || i3 > g.n3) return false;
  out = new_grid_cell(g, i1, i2, i3);
  return true;
}

// adjust_wall (:239-346)
void cyl_adjust_wall(const orc_ctx &g, Photon &p) {
  const int eps = 3;
  p.on_wall = false;
  p.on_wall_id = WallId();
  double w_sq, phi;
  cyl_phi(p, w_sq, phi);
  if ((p.r.x * p.v.x + p.r.y * p.v.y) >= 0.0) {
    if (equal_nulp(w_sq, g.wr2[p.icell.i1 - 1], eps)) {
      p.on_wall_id.w1 = -1;
    } else if (equal_nulp(w_sq, g.wr2[p.icell.i1], eps)) {
      p.on_wall_id.w1 = -1;
      p.icell.i1 = p.icell.i1 + 1;
    }
  } else {
    if (equal_nulp(w_sq, g.wr2[p.icell.i1 - 1], eps)) {
      p.on_wall_id.w1 = +1;
      p.icell.i1 = p.icell.i1 - 1;
    } else if (equal_nulp(w_sq, g.wr2[p.icell.i1], eps)) {
      p.on_wall_id.w1 = +1;
    }
  }
  if (p.v.z > 0.0) {
    if (equal_nulp(p.r.z, g.w2[p.icell.i2 - 1], eps)) {
      p.on_wall_id.w2 = -1;
    } else if (equal_nulp(p.r.z, g.w2[p.icell.i2], eps)) {
      p.on_wall_id.w2 = -1;
      p.icell.i2 = p.icell.i2 + 1;
    }
  } else if (p.v.z < 0.0) {
    if (equal_nulp(p.r.z, g.w2[p.icell.i2 - 1], eps)) {
      p.on_wall_id.w2 = +1;
      p.icell.i2 = p.icell.i2 - 1;
    } else if (equal_nulp(p.r.z, g.w2[p.icell.i2], eps)) {
      p.on_wall_id.w2 = +1;
    }
  }
  if (p.r.x == 0.0 && p.r.y == 0.0 && p.v.x == 0.0 && p.v.y == 0.0) {
    // on every phi wall at once
  } else if (equal_nulp(phi, g.w3[p.icell.i3 - 1], eps)) {
    double phi_v = std::atan2(p.v.y, p.v.x);
    double dphi = phi_v - g.w3[p.icell.i3 - 1];
    if (dphi < -PI_F) dphi = dphi + TWOPI_F;
    if (dphi > 0.0) {
      p.on_wall_id.w3 = -1;
    } else {
      p.on_wall_id.w3 = +1;
      p.icell.i3 = p.icell.i3 - 1;
      if (p.icell.i3 == 0) p.icell.i3 = g.n3;
    }
  } else if (equal_nulp(phi, g.w3[p.icell.i3], eps)) {
    double phi_v = std::atan2(p.v.y, p.v.x);
    double dphi = phi_v - g.w3[p.icell.i3];
    if (dphi < -PI_F) dphi = dphi + TWOPI_F;
    if (dphi > 0.0) {
      p.on_wall_id.w3 = -1;
      p.icell.i3 = p.icell.i3 + 1;
      if (p.icell.i3 == g.n3 + 1) p.icell.i3 = 1;
    } else {
      p.on_wall_id.w3 = +1;
    }
  }
  p.on_wall = p.on_wall_id.w1 != 0 || p.on_wall_id.w2 != 0 || p.on_wall_id.w3 != 0;
}

// in_correct_cell (:443-514)
bool cyl_in_correct_cell(const orc_ctx &g, const Photon &p) {
  const double threshold = 1.e-3;
  Cell act;
  bool valid = cyl_find_cell(g, p, act);
  if (!valid) act = Cell{-1, -1, -1, -1};
  if (!p.on_wall) return act.i1 == p.icell.i1 && act.i2 == p.icell.i2 && act.i3 == p.icell.i3;
  bool ok = true;
  double w_sq, phi, frac, dphi;
  cyl_phi(p, w_sq, phi);
  if (p.on_wall_id.w1 == -1) {
    if (g.w1[p.icell.i1 - 1] != std::sqrt(w_sq)) {
      frac = std::sqrt(w_sq) / g.w1[p.icell.i1 - 1] - 1.0;
      ok = ok && std::fabs(frac) < threshold;
    }
  } else if (p.on_wall_id.w1 == +1) {
    if (g.w1[p.icell.i1] != std::sqrt(w_sq)) {
      frac = std::sqrt(w_sq) / g.w1[p.icell.i1] - 1.0;
      ok = ok && std::fabs(frac) < threshold;
    }
  } else {
    ok = ok && act.i1 == p.icell.i1;
  }
  if (p.on_wall_id.w2 == -1) {
    frac = (p.r.z - g.w2[p.icell.i2 - 1]) / (g.w2[p.icell.i2] - g.w2[p.icell.i2 - 1]);
    ok = ok && std::fabs(frac) < threshold;
  } else if (p.on_wall_id.w2 == +1) {
    frac = (p.r.z - g.w2[p.icell.i2]) / (g.w2[p.icell.i2] - g.w2[p.icell.i2 - 1]);
    ok = ok && std::fabs(frac) < threshold;
  } else {
    ok = ok && act.i2 == p.icell.i2;
  }
  if (p.on_wall_id.w3 == -1) {
    dphi = phi - g.w3[p.icell.i3 - 1];
    if (dphi > PI_F) dphi = dphi - TWOPI_F;
    if (dphi < -PI_F) dphi = dphi + TWOPI_F;
    frac = dphi / (g.w3[p.icell.i3] - g.w3[p.icell.i3 - 1]);
    ok = ok && std::fabs(frac) < threshold;
  } else if (p.on_wall_id.w3 == +1) {
    dphi = phi - g.w3[p.icell.i3];
    if (dphi > PI_F) dphi = dphi - TWOPI_F;
    if (dphi < -PI_F) dphi = dphi + TWOPI_F;
    frac = dphi / (g.w3[p.icell.i3] - g.w3[p.icell.i3 - 1]);
    ok = ok && std::fabs(frac) < threshold;
  } else {
    ok = ok && act.i3 == p.icell.i3;
  }
  return ok;
}

// find_wall (:592-770)
void cyl_find_wall(orc_ctx &g, const Photon &p, double &tnearest, WallId &id_min) {
  g.tmin = std::numeric_limits<double>::max();
  g.emin = 0.0;
  g.imin = WallId();
  g.iext = WallId();
  double v2_xy = p.v.x * p.v.x + p.v.y * p.v.y;
  double rv_xy = p.r.x * p.v.x + p.r.y * p.v.y;
  double r2_xy = p.r.x * p.r.x + p.r.y * p.r.y;
  double pB = rv_xy / v2_xy;
  pB = pB + pB;
  double pC = r2_xy / v2_xy;
  double t1, t2;
  double pC_1 = pC - g.wr2[p.icell.i1 - 1] / v2_xy;
  quadratic_pascal_reduced(pB, pC_1, t1, t2);
  if (p.on_wall_id.w1 == -1) {
    if (std::fabs(t1) < std::fabs(t2))
      insert_t(g, t2, 1, -1, g.ew1[p.icell.i1 - 1]);
    else
      insert_t(g, t1, 1, -1, g.ew1[p.icell.i1 - 1]);
  } else {
    insert_t(g, t1, 1, -1, g.ew1[p.icell.i1 - 1]);
    insert_t(g, t2, 1, -1, g.ew1[p.icell.i1 - 1]);
  }
  double pC_2 = pC - g.wr2[p.icell.i1] / v2_xy;
  quadratic_pascal_reduced(pB, pC_2, t1, t2);
  if (p.on_wall_id.w1 == +1) {
    if (std::fabs(t1) < std::fabs(t2))
      insert_t(g, t2, 1, +1, g.ew1[p.icell.i1]);
    else
      insert_t(g, t1, 1, +1, g.ew1[p.icell.i1]);
  } else {
    insert_t(g, t1, 1, +1, g.ew1[p.icell.i1]);
    insert_t(g, t2, 1, +1, g.ew1[p.icell.i1]);
  }
  if (p.on_wall_id.w2 != -1) {
    t1 = (g.w2[p.icell.i2 - 1] - p.r.z) / p.v.z;
    insert_t(g, t1, 2, -1, 0.0);
  }
  if (p.on_wall_id.w2 != +1) {
    t2 = (g.w2[p.icell.i2] - p.r.z) / p.v.z;
    insert_t(g, t2, 2, +1, 0.0);
  }
  if (g.n3 > 1) {
    double dphi = 0.0;
    if (p.on_wall_id.w3 == -1) {
      dphi = std::atan2(p.v.y, p.v.x) - g.w3[p.icell.i3 - 1];
      if (dphi > PI_F) dphi = dphi - TWOPI_F;
      if (dphi < -PI_F) dphi = dphi + TWOPI_F;
    }
    if (p.on_wall_id.w3 == +1) {
      dphi = std::atan2(p.v.y, p.v.x) - g.w3[p.icell.i3];
      if (dphi > PI_F) dphi = dphi - TWOPI_F;
      if (dphi < -PI_F) dphi = dphi + TWOPI_F;
    }
    if (p.on_wall_id.w3 == +1 && std::fabs(dphi) < g.ew3[p.icell.i3]) {
      g.iext.w3 = +1;
    } else if (p.on_wall_id.w3 == -1 && std::fabs(dphi) < g.ew3[p.icell.i3 - 1]) {
      g.iext.w3 = -1;
    } else if (r2_xy > 0.0) {
      if (p.on_wall_id.w3 != -1) {
        double tp = g.wtanp[p.icell.i3 - 1];
        t1 = -(tp * p.r.x - p.r.y) / (tp * p.v.x - p.v.y);
        double x_i = p.r.x + p.v.x * t1, y_i = p.r.y + p.v.y * t1;
        double dp = std::fabs(std::atan2(y_i, x_i) - g.w3[p.icell.i3 - 1]);
        if (dp > PI_F) dp = std::fabs(dp - TWOPI_F);
        if (dp < 0.5 * PI_F) insert_t(g, t1, 3, -1, 0.0);
      }
      if (p.on_wall_id.w3 != +1) {
        double tp = g.wtanp[p.icell.i3];
        t2 = -(tp * p.r.x - p.r.y) / (tp * p.v.x - p.v.y);
        double x_i = p.r.x + p.v.x * t2, y_i = p.r.y + p.v.y * t2;
        double dp = std::fabs(std::atan2(y_i, x_i) - g.w3[p.icell.i3]);
        if (dp > PI_F) dp = std::fabs(dp - TWOPI_F);
        if (dp < 0.5 * PI_F) insert_t(g, t2, 3, +1, 0.0);
      }
    }
  }
  tnearest = g.tmin;
  id_min = WallId{g.imin.w1 + g.iext.w1, g.imin.w2 + g.iext.w2, g.imin.w3 + g.iext.w3};
}

// random_position_cell (:516-547)
Vec cyl_random_position_cell(orc_ctx &g, const Cell &c) {
  double r = g.rng.random(), z = g.rng.random(), ph = g.rng.random();
  double a = g.w1[c.i1 - 1], b = g.w1[c.i1];
  r = std::sqrt(r * (b * b - a * a) + a * a);
  z = z * (g.w2[c.i2] - g.w2[c.i2 - 1]) + g.w2[c.i2 - 1];
  ph = ph * (g.w3[c.i3] - g.w3[c.i3 - 1]) + g.w3[c.i3 - 1];
  if (r <= a || r >= b) r = 0.5 * (a + b);
  if (z <= g.w2[c.i2 - 1] || z >= g.w2[c.i2]) z = 0.5 * (g.w2[c.i2 - 1] + g.w2[c.i2]);
  if (ph <= g.w3[c.i3 - 1] || ph >= g.w3[c.i3]) ph = 0.5 * (g.w3[c.i3 - 1] + g.w3[c.i3]);
  return Vec{r * std::cos(ph), r * std::sin(ph), z};
}

// ---------------------------------------------------------------------------
// octree geometry (src/grid/grid_geometry_octree.f90); cell ids are 1-based as in the Fortran
// ---------------------------------------------------------------------------
// subcell_id (:98-133): 1..8, x fastest
int oct_subcell_id(const orc_ctx &g, int cell_id, const Vec &r) {
  const int k = cell_id - 1;
  int s = 1;
  if (!(r.x < g.ox[k])) s += 1;
  if (!(r.y < g.oy[k])) s += 2;
  if (!(r.z < g.oz[k])) s += 4;
  return s;
}

// locate_cell (:135-146)
int oct_locate_cell(const orc_ctx &g, const Vec &r, int cell_id) {
  while (g.orefined[cell_id - 1]) cell_id = g.ochildren[(size_t)8 * (cell_id - 1) + oct_subcell_id(g, cell_id, r) - 1];
  return cell_id;
}

// find_cell (:277-297)
bool oct_find_cell(const orc_ctx &g, const Photon &p, Cell &out) {
  if (p.r.x < g.ox[0] - g.odx[0] || p.r.x > g.ox[0] + g.odx[0]) return false;
  if (p.r.y < g.oy[0] - g.ody[0] || p.r.y > g.oy[0] + g.ody[0]) return false;
  if (p.r.z < g.oz[0] - g.odz[0] || p.r.z > g.oz[0] + g.odz[0]) return false;
  out = Cell{0, 0, 0, oct_locate_cell(g, p.r, 1)};
  return true;
}

// next_cell_int (:328-347) + next_cell_wall_id (:349-367)
Cell oct_next_cell(const orc_ctx &g, const Cell &c, const WallId &dir, const Vec &r) {
  // opposite_cell(subcell, wall) (:53-59)
  static const int opposite_cell[6][8] = {{0, 1, 0, 3, 0, 5, 0, 7}, {2, 0, 4, 0, 6, 0, 8, 0}, {0, 0, 1, 2, 0, 0, 5, 6},
                                          {3, 4, 0, 0, 7, 8, 0, 0}, {0, 0, 0, 0, 1, 2, 3, 4}, {5, 6, 7, 8, 0, 0, 0, 0}};
  int direction;
  if (dir.w1 == -1) direction = 1;
  else if (dir.w1 == +1) direction = 2;
  else if (dir.w2 == -1) direction = 3;
  else if (dir.w2 == +1) direction = 4;
  else if (dir.w3 == -1) direction = 5;
  else if (dir.w3 == +1) direction = 6;
  else return c;
  int ic = c.ic;
  for (;;) {
    if (ic == 1) return Cell{0, 0, 0, g.n_cells + 1};
    int sub = opposite_cell[direction - 1][g.oparent_subcell[ic - 1] - 1];
    int parent = g.oparent[ic - 1];
    if (sub > 0) return Cell{0, 0, 0, oct_locate_cell(g, r, g.ochildren[(size_t)8 * (parent - 1) + sub - 1])};
    ic = parent;
  }
}

// in_correct_cell (:369-394)
bool oct_in_correct_cell(const orc_ctx &g, const Photon &p) {
  Cell act;
  bool valid = oct_find_cell(g, p, act);
  if (!valid) act = Cell{0, 0, 0, -1};
  if (!p.on_wall) return act.ic == p.icell.ic;
  const int k = p.icell.ic - 1;
  double frac1 = std::fabs(p.r.x - g.ox[k]) / g.odx[k];
  double frac2 = std::fabs(p.r.y - g.oy[k]) / g.ody[k];
  double frac3 = std::fabs(p.r.z - g.oz[k]) / g.odz[k];
  double frac = 0.0;
  bool ok = true;
  if (std::abs(p.on_wall_id.w1) == 1) {
    frac = frac1 - 1.0;
    ok = frac2 < 1.0 && frac3 < 1.0;
  }
  if (std::abs(p.on_wall_id.w2) == 1) {
    frac = frac2 - 1.0;
    ok = frac1 < 1.0 && frac3 < 1.0;
  }
  if (std::abs(p.on_wall_id.w3) == 1) {
    frac = frac3 - 1.0;
    ok = frac1 < 1.0 && frac2 < 1.0;
  }
  return std::fabs(frac) < 1.e-3 && ok;
}

// find_wall (:438-537)
void oct_find_wall(orc_ctx &g, const Photon &p, double &tmin, WallId &id_min) {
  const double huge = std::numeric_limits<double>::max();
  const int k = p.icell.ic - 1;
  id_min = WallId();
  bool pos_vx = p.v.x > 0.0, pos_vy = p.v.y > 0.0, pos_vz = p.v.z > 0.0;
  double tx, ty, tz;
  if (pos_vx) tx = (g.ox[k] + g.odx[k] - p.r.x) / p.v.x;
  else if (p.v.x < 0.0) tx = (g.ox[k] - g.odx[k] - p.r.x) / p.v.x;
  else tx = huge;
  if (pos_vy) ty = (g.oy[k] + g.ody[k] - p.r.y) / p.v.y;
  else if (p.v.y < 0.0) ty = (g.oy[k] - g.ody[k] - p.r.y) / p.v.y;
  else ty = huge;
  if (pos_vz) tz = (g.oz[k] + g.odz[k] - p.r.z) / p.v.z;
  else if (p.v.z < 0.0) tz = (g.oz[k] - g.odz[k] - p.r.z) / p.v.z;
  else tz = huge;
  if (tx < tz) {
    if (tx < ty) {
      id_min.w1 = pos_vx ? +1 : -1;
      tmin = tx;
    } else {
      id_min.w2 = pos_vy ? +1 : -1;
      tmin = ty;
    }
  } else {
    if (tz < ty) {
      id_min.w3 = pos_vz ? +1 : -1;
      tmin = tz;
    } else {
      id_min.w2 = pos_vy ? +1 : -1;
      tmin = ty;
    }
  }
  if (tmin < 0.0) {
    if (tmin > -10 * g.oct_eps)
      tmin = 0.0;
    else
      id_min = WallId();
  }
}

// random_position_cell (:396-408)
Vec vor_random_position_cell(orc_ctx &g, const Cell &c);
Vec oct_random_position_cell(orc_ctx &g, const Cell &c) {
  double x = g.rng.random(), y = g.rng.random(), z = g.rng.random();
  const int k = c.ic - 1;
  return Vec{(2.0 * x - 1.0) * g.odx[k] + g.ox[k], (2.0 * y - 1.0) * g.ody[k] + g.oy[k], (2.0 * z - 1.0) * g.odz[k] + g.oz[k]};
}

// ---------------------------------------------------------------------------
// Voronoi geometry (src/grid/grid_geometry_voronoi.f90); cell ids 1-based as in the Fortran, n_cells + 1 = outside
// ---------------------------------------------------------------------------
// The k <= 2 nearest sites of a point (kdtree2_n_nearest): buckets of a uniform grid are searched in shells around
// the point's bucket until no unvisited bucket can hold a closer site.
void vor_nearest(const orc_ctx &g, double x, double y, double z, int k, int *idx) {
  double best[2] = {std::numeric_limits<double>::max(), std::numeric_limits<double>::max()};
  idx[0] = idx[1] = 0;
  double w[3], lo[3] = {g.vbox[0], g.vbox[2], g.vbox[4]}, pos[3] = {x, y, z};
  int b[3];
  for (int a = 0; a < 3; a++) {
    w[a] = (g.vbox[2 * a + 1] - g.vbox[2 * a]) / g.vg[a];
    b[a] = std::min(std::max((int)((pos[a] - lo[a]) / w[a]), 0), g.vg[a] - 1);
  }
  const double wmin = std::min(w[0], std::min(w[1], w[2]));
  const int rmax = std::max(g.vg[0], std::max(g.vg[1], g.vg[2]));
  for (int r = 0; r <= rmax; r++) {
    // every site in a bucket of shell r or beyond is at least (r - 1) * wmin away (the point lies anywhere in its bucket)
    if (r > 0) {
      const double reach = (double)(r - 1) * wmin;
      if (best[k - 1] <= reach * reach) break;
    }
    for (int k3 = b[2] - r; k3 <= b[2] + r; k3++) {
      if (k3 < 0 || k3 >= g.vg[2]) continue;
      for (int k2 = b[1] - r; k2 <= b[1] + r; k2++) {
        if (k2 < 0 || k2 >= g.vg[1]) continue;
        for (int k1 = b[0] - r; k1 <= b[0] + r; k1++) {
          if (k1 < 0 || k1 >= g.vg[0]) continue;
          if (std::max(std::abs(k1 - b[0]), std::max(std::abs(k2 - b[1]), std::abs(k3 - b[2]))) != r) continue;
          const int cell = (k3 * g.vg[1] + k2) * g.vg[0] + k1;
          for (int q = g.vg_start[cell]; q < g.vg_start[cell + 1]; q++) {
            const int i = g.vg_sites[q];
            const double dx = g.vx[i] - x, dy = g.vy[i] - y, dz = g.vz[i] - z;
            const double d2 = dx * dx + dy * dy + dz * dz;
            if (d2 < best[0]) {
              best[1] = best[0]; idx[1] = idx[0];
              best[0] = d2; idx[0] = i + 1;
            } else if (d2 < best[1]) {
              best[1] = d2; idx[1] = i + 1;
            }
          }
        }
      }
    }
  }
}

// find_cell (:195-228)
bool vor_find_cell(const orc_ctx &g, const Photon &p, Cell &out) {
  if (p.r.x < g.vbox[0] || p.r.x > g.vbox[1]) return false;
  if (p.r.y < g.vbox[2] || p.r.y > g.vbox[3]) return false;
  if (p.r.z < g.vbox[4] || p.r.z > g.vbox[5]) return false;
  int idx[2];
  vor_nearest(g, p.r.x, p.r.y, p.r.z, 1, idx);
  out = Cell{0, 0, 0, idx[0]};
  return true;
}

// in_correct_cell (:274-283): the cell is one of the two nearest sites
bool vor_in_correct_cell(const orc_ctx &g, const Photon &p) {
  int idx[2];
  vor_nearest(g, p.r.x, p.r.y, p.r.z, 2, idx);
  return p.icell.ic == idx[0] || p.icell.ic == idx[1];
}

// find_wall (:322-402): the planes that bisect the segments to the neighbouring sites, and the walls of the box
void vor_find_wall(orc_ctx &g, const Photon &p, double &tmin, WallId &id_min) {
  g.tmin = std::numeric_limits<double>::max();
  g.emin = 0.0;
  g.imin = WallId();
  const int ic = p.icell.ic - 1;
  for (int q = g.vidx[ic]; q < g.vidx[ic + 1]; q++) {
    const int nb = g.vneigh[q] + 1;   // the Fortran numbering: cells from 1, walls 0 .. -5
    if (nb <= 0 && nb >= -5) {
      switch (nb) {
        case 0: if (p.v.x < 0.0) insert_t(g, (g.vbox[0] - p.r.x) / p.v.x, 1, g.n_cells + 1, 0.0); break;
        case -1: if (p.v.x > 0.0) insert_t(g, (g.vbox[1] - p.r.x) / p.v.x, 1, g.n_cells + 1, 0.0); break;
        case -2: if (p.v.y < 0.0) insert_t(g, (g.vbox[2] - p.r.y) / p.v.y, 1, g.n_cells + 1, 0.0); break;
        case -3: if (p.v.y > 0.0) insert_t(g, (g.vbox[3] - p.r.y) / p.v.y, 1, g.n_cells + 1, 0.0); break;
        case -4: if (p.v.z < 0.0) insert_t(g, (g.vbox[4] - p.r.z) / p.v.z, 1, g.n_cells + 1, 0.0); break;
        default: if (p.v.z > 0.0) insert_t(g, (g.vbox[5] - p.r.z) / p.v.z, 1, g.n_cells + 1, 0.0); break;
      }
      continue;
    }
    if (nb == -p.on_wall_id.w2) continue;   // the wall the packet stands on (w2 keeps the previous cell)
    const double nx = g.vx[nb - 1] - g.vx[ic], ny = g.vy[nb - 1] - g.vy[ic], nz = g.vz[nb - 1] - g.vz[ic];
    const double mx = 0.5 * (g.vx[nb - 1] + g.vx[ic]), my = 0.5 * (g.vy[nb - 1] + g.vy[ic]), mz = 0.5 * (g.vz[nb - 1] + g.vz[ic]);
    const double t = (nx * (mx - p.r.x) + ny * (my - p.r.y) + nz * (mz - p.r.z)) / (nx * p.v.x + ny * p.v.y + nz * p.v.z);
    insert_t(g, t, 1, nb, 0.0);
  }
  tmin = g.tmin;
  id_min = g.imin;
  id_min.w2 = p.icell.ic;   // "use w2 to store previous cell"
}

// random_position_cell (:285-312): rejection sampling in the cell's bounding box
Vec vor_random_position_cell(orc_ctx &g, const Cell &c) {
  const double *bb = &g.vbb[(size_t)6 * (c.ic - 1)];
  for (int i = 0; i < 1000000; i++) {
    Vec pos;
    pos.x = g.rng.random() * (bb[1] - bb[0]) + bb[0];   // random_uni (lib_random.f90)
    pos.y = g.rng.random() * (bb[3] - bb[2]) + bb[2];
    pos.z = g.rng.random() * (bb[5] - bb[4]) + bb[4];
    int idx[2];
    vor_nearest(g, pos.x, pos.y, pos.z, 1, idx);
    if (idx[0] == c.ic) return pos;
  }
  throw OracleError{"too many samples"};
}

// ---------------------------------------------------------------------------
// AMR geometry (src/grid/grid_geometry_amr.f90); level / grid ids 1-based as in the Fortran
// ---------------------------------------------------------------------------
bool amr_in_grid(const orc_ctx::AmrGrid &G, const Vec &r) {
  if (r.x < G.xmin) return false;
  if (r.x > G.xmax) return false;
  if (r.y < G.ymin) return false;
  if (r.y > G.ymax) return false;
  if (r.z < G.zmin) return false;
  if (r.z > G.zmax) return false;
  return true;
}

// new_grid_cell_5d (type_cell_id_amr.f90:104-117)
Cell amr_cell(const orc_ctx &g, int i1, int i2, int i3, int ilevel, int igrid) {
  const orc_ctx::AmrGrid &G = g.levels[ilevel - 1][igrid - 1];
  Cell c;
  c.ic = (G.start_id - 1) + (i3 - 1) * G.n1 * G.n2 + (i2 - 1) * G.n1 + i1;
  c.ilevel = ilevel;
  c.igrid = igrid;
  c.i1 = i1;
  c.i2 = i2;
  c.i3 = i3;
  return c;
}

// new_grid_cell_1d (:119-129)
Cell amr_cell_1d(const orc_ctx &g, int ic) {
  Cell c;
  c.ic = ic;
  c.ilevel = g.cell_ilevel[ic - 1];
  c.igrid = g.cell_igrid[ic - 1];
  c.i1 = g.cell_i1[ic - 1];
  c.i2 = g.cell_i2[ic - 1];
  c.i3 = g.cell_i3[ic - 1];
  return c;
}

// ipos2 (:510-519)
int amr_ipos2(double xmin, double xmax, double x, int nbin) {
  double eps = (xmax - xmin) * 1.e-10;
  int i = ipos(xmin, xmax, x, nbin);
  if (i == 0 && std::fabs(x - xmin) < eps) i = 1;
  if (i == nbin + 1 && std::fabs(x - xmax) < eps) i = nbin;
  return i;
}

// find_position_in_grid (:521-545); false = invalid_cell
bool amr_find_position_in_grid(const orc_ctx &g, const Vec &r, int ilevel, int igrid, Cell &out) {
  for (;;) {
    const orc_ctx::AmrGrid &G = g.levels[ilevel - 1][igrid - 1];
    int i1 = amr_ipos2(G.xmin, G.xmax, r.x, G.n1);
    int i2 = amr_ipos2(G.ymin, G.ymax, r.y, G.n2);
    int i3 = amr_ipos2(G.zmin, G.zmax, r.z, G.n3);
    int ilevel_new = G.goto_level[G.gidx(i1, i2, i3)];
    int igrid_new = G.goto_grid[G.gidx(i1, i2, i3)];
    if (ilevel_new == 0) {
      if (i1 < 1 || i1 > G.n1 || i2 < 1 || i2 > G.n2 || i3 < 1 || i3 > G.n3) return false;
      out = amr_cell(g, i1, i2, i3, ilevel, igrid);
      return true;
    }
    ilevel = ilevel_new;
    igrid = igrid_new;
  }
}

// find_cell / find_cell_position (:553-573)
bool amr_find_cell(const orc_ctx &g, const Photon &p, Cell &out) {
  int igrid = -1;
  for (size_t k = 0; k < g.levels[0].size(); k++)
    if (amr_in_grid(g.levels[0][k], p.r)) {
      igrid = (int)k + 1;
      break;
    }
  if (igrid == -1) return false;
  return amr_find_position_in_grid(g, p.r, 1, igrid, out);
}

// next_cell_int (:599-655) + next_cell_wall_id (:657-675)
Cell amr_next_cell(const orc_ctx &g, const Cell &cell, const WallId &dir, const Vec &intersection) {
  int direction;
  if (dir.w1 == -1) direction = 1;
  else if (dir.w1 == +1) direction = 2;
  else if (dir.w2 == -1) direction = 3;
  else if (dir.w2 == +1) direction = 4;
  else if (dir.w3 == -1) direction = 5;
  else if (dir.w3 == +1) direction = 6;
  else return cell;
  const orc_ctx::AmrGrid &G = g.levels[cell.ilevel - 1][cell.igrid - 1];
  int i1 = cell.i1, i2 = cell.i2, i3 = cell.i3;
  switch (direction) {
    case 1: i1 = i1 - 1; break;
    case 2: i1 = i1 + 1; break;
    case 3: i2 = i2 - 1; break;
    case 4: i2 = i2 + 1; break;
    case 5: i3 = i3 - 1; break;
    case 6: i3 = i3 + 1; break;
  }
  if (G.goto_level[G.gidx(i1, i2, i3)] == 0) {
    if (i1 == 0 || i1 == G.n1 + 1 || i2 == 0 || i2 == G.n2 + 1 || i3 == 0 || i3 == G.n3 + 1) {
      Cell out;
      out.ic = out.i1 = out.i2 = out.i3 = out.ilevel = out.igrid = -2;  // outside_cell
      return out;
    }
    return amr_cell(g, i1, i2, i3, cell.ilevel, cell.igrid);
  }
  int ilevel = G.goto_level[G.gidx(i1, i2, i3)], igrid = G.goto_grid[G.gidx(i1, i2, i3)];
  Vec r = intersection;
  switch (direction) {
    case 1: r.x = r.x - g.amr_eps; break;
    case 2: r.x = r.x + g.amr_eps; break;
    case 3: r.y = r.y - g.amr_eps; break;
    case 4: r.y = r.y + g.amr_eps; break;
    case 5: r.z = r.z - g.amr_eps; break;
    case 6: r.z = r.z + g.amr_eps; break;
  }
  Cell out;
  if (!amr_find_position_in_grid(g, r, ilevel, igrid, out)) out.ic = out.i1 = out.i2 = out.i3 = out.ilevel = out.igrid = -1;
  return out;
}

// in_correct_cell (:677-727)
bool amr_in_correct_cell(const orc_ctx &g, const Photon &p) {
  const orc_ctx::AmrGrid &G = g.levels[p.icell.ilevel - 1][p.icell.igrid - 1];
  int a1 = ipos(G.xmin, G.xmax, p.r.x, G.n1), a2 = ipos(G.ymin, G.ymax, p.r.y, G.n2), a3 = ipos(G.zmin, G.zmax, p.r.z, G.n3);
  if (!p.on_wall) return a1 == p.icell.i1 && a2 == p.icell.i2 && a3 == p.icell.i3;
  bool ok = true;
  double frac;
#define CHK(WID, R, W, I, IA)                              \
  if (WID == -1) {                                         \
    frac = (R - W[I - 1]) / (W[I] - W[I - 1]);             \
    ok = ok && std::fabs(frac) < 1.e-3;                    \
  } else if (WID == +1) {                                  \
    frac = (R - W[I]) / (W[I] - W[I - 1]);                 \
    ok = ok && std::fabs(frac) < 1.e-3;                    \
  } else {                                                 \
    ok = ok && IA == I;                                    \
  }
  CHK(p.on_wall_id.w1, p.r.x, G.w1, p.icell.i1, a1)
  CHK(p.on_wall_id.w2, p.r.y, G.w2, p.icell.i2, a2)
  CHK(p.on_wall_id.w3, p.r.z, G.w3, p.icell.i3, a3)
#undef CHK
  return ok;
}

// find_wall (:775-871)
void amr_find_wall(orc_ctx &g, const Photon &p, double &tmin, WallId &id_min) {
  const double huge = std::numeric_limits<double>::max();
  const orc_ctx::AmrGrid &G = g.levels[p.icell.ilevel - 1][p.icell.igrid - 1];
  id_min = WallId();
  bool pos_vx = p.v.x > 0.0, pos_vy = p.v.y > 0.0, pos_vz = p.v.z > 0.0;
  double tx, ty, tz;
  if (pos_vx) tx = (G.w1[p.icell.i1] - p.r.x) / p.v.x;
  else if (p.v.x < 0.0) tx = (G.w1[p.icell.i1 - 1] - p.r.x) / p.v.x;
  else tx = huge;
  if (pos_vy) ty = (G.w2[p.icell.i2] - p.r.y) / p.v.y;
  else if (p.v.y < 0.0) ty = (G.w2[p.icell.i2 - 1] - p.r.y) / p.v.y;
  else ty = huge;
  if (pos_vz) tz = (G.w3[p.icell.i3] - p.r.z) / p.v.z;
  else if (p.v.z < 0.0) tz = (G.w3[p.icell.i3 - 1] - p.r.z) / p.v.z;
  else tz = huge;
  if (std::min(tx, std::min(ty, tz)) < 0.0) throw OracleError{"negative t"};
  if (tx < tz) {
    if (tx < ty) {
      id_min.w1 = pos_vx ? +1 : -1;
      tmin = tx;
    } else {
      id_min.w2 = pos_vy ? +1 : -1;
      tmin = ty;
    }
  } else {
    if (tz < ty) {
      id_min.w3 = pos_vz ? +1 : -1;
      tmin = tz;
    } else {
      id_min.w2 = pos_vy ? +1 : -1;
      tmin = ty;
    }
  }
}

// random_position_cell (:729-741)
Vec amr_random_position_cell(orc_ctx &g, const Cell &c) {
  const orc_ctx::AmrGrid &G = g.levels[c.ilevel - 1][c.igrid - 1];
  double x = g.rng.random(), y = g.rng.random(), z = g.rng.random();
  return Vec{x * (G.w1[c.i1] - G.w1[c.i1 - 1]) + G.w1[c.i1 - 1], y * (G.w2[c.i2] - G.w2[c.i2 - 1]) + G.w2[c.i2 - 1],
             z * (G.w3[c.i3] - G.w3[c.i3 - 1]) + G.w3[c.i3 - 1]};
}

// update_optconsts (dust.f90:64-79)
void update_optconsts(orc_ctx &g, Photon &p) {
  for (int id = 0; id < g.n_dust; id++) {
    const Dust &d = g.d[id];
    if (p.nu < d.nu[0] || p.nu > d.nu[d.n_nu - 1]) {
      char buf[256];
      snprintf(buf, sizeof buf,
               "photon frequency (%10.4E Hz) is outside the range defined for the dust optical properties "
               "(%10.4E to %10.4E Hz)",
               p.nu, d.nu[0], d.nu[d.n_nu - 1]);
      throw OracleError{buf};
    }
    p.current_chi[id] = interp1d_loglog(d.nu.data(), d.chi_nu.data(), d.n_nu, p.nu);
    p.current_albedo[id] = interp1d_loglog(d.nu.data(), d.albedo_nu.data(), d.n_nu, p.nu);
    p.current_kappa[id] = p.current_chi[id] * (1.0 - p.current_albedo[id]);
  }
}

// quadratic_pascal_reduced_dp (fortranlib/src/lib_algebra.f90:145-164): x^2 + b x + c = 0
void quadratic_pascal_reduced(double b, double c, double &x1, double &x2) {
  const double huge = std::numeric_limits<double>::max();
  double delta = b * b - 4.0 * c;
  if (delta > 0) {
    delta = std::sqrt(delta);
    delta = std::copysign(delta, b);
    double q = -0.5 * (b + delta);
    x1 = q;
    x2 = c / q;
  } else if (delta < 0) {
    x1 = -huge;
    x2 = -huge;
  } else {
    x1 = -2.0 * c / b;
    x2 = -huge;
  }
}

// source_distance (source_type.f90:324-357)
double source_distance(const Source &src, const Vec &r, const Vec &v) {
  double d = std::numeric_limits<double>::infinity();
  if (src.type == HYP_SOURCE_SPHERE) {
    const double tol = (double)1.e-8f;
    Vec dr{r.x - src.position.x, r.y - src.position.y, r.z - src.position.z};
    double pB = 2.0 * (dr.x * v.x + dr.y * v.y + dr.z * v.z);
    double pC = (dr.x * dr.x + dr.y * dr.y + dr.z * dr.z) - src.radius * src.radius;
    double t1, t2;
    quadratic_pascal_reduced(pB, pC, t1, t2);
    if (t1 < d && t1 > tol * src.radius) d = t1;
    if (t2 < d && t2 > tol * src.radius) d = t2;
  }
  return d;
}

// find_nearest_source (source.f90:206-227)
void find_nearest_source(const orc_ctx &g, const Vec &r, const Vec &v, double &nearest, int &nearest_id) {
  nearest_id = 0;
  nearest = std::numeric_limits<double>::infinity();
  if (!g.any_intersect) return;
  for (size_t is = 0; is < g.s.size(); is++)
    if (g.s[is].intersect) {
      if (source_distance(g.s[is], r, v) < nearest) {
        nearest = source_distance(g.s[is], r, v);
        nearest_id = (int)is + 1;
      }
    }
}

// cbrt_dp (lib_algebra.f90:79-88)
double cbrt_f(double x) {
  const double alpha = 1.0 / 3.0;
  return x >= 0. ? std::pow(x, alpha) : -std::pow(std::fabs(x), alpha);
}

// ran_mu_limb (source_type.f90:982-1084): P = a mu^2 + b mu
double ran_mu_limb(Rng &rng, double a, double b) {
  double s1 = a * (1.0 / 3.0), t1 = b * (1.0 / 2.0);
  double norm = s1 + t1;
  s1 = s1 / norm;
  t1 = t1 / norm;
  double xi = rng.random();
  xi = -xi;
  // cubic_real_root_v2: x^3 + bb x^2 + dd = 0
  double bb = t1 / s1, dd = xi / s1;
  const double alpha = 1.0 / 3.0, gamma = 1.0 / 27.0;
  double pp = -bb * bb * alpha * alpha;
  double q = (dd + 2.0 * bb * bb * bb * gamma) * 0.5;
  double p3 = pp * pp * pp, q2 = q * q;
  double delta = q2 + p3;
  if (delta < 0) {
    double phi = std::acos(-q / std::sqrt(std::fabs(p3)));
    double y = +2 * std::sqrt(std::fabs(pp)) * std::cos(phi * alpha);
    return y - bb * alpha;
  }
  delta = std::sqrt(delta);
  double u = cbrt_f(-q + delta), v = cbrt_f(-q - delta);
  return u + v - bb * alpha;
}

// emit_from_sphere (source_type.f90:604-690), no spots
void emit_from_sphere(orc_ctx &g, const Source &src, Photon &p, int spot = 0) {
  Angle a_coord;
  if (spot > 0) {
    // rejection sampling of a point inside the spot (source_type.f90:632-637)
    const Source::Spot &sp = src.spot[spot - 1];
    for (;;) {
      a_coord = random_sphere_angle3d(g.rng);
      const double dot = a_coord.sint * a_coord.cosp * sp.a.sint * sp.a.cosp + a_coord.sint * a_coord.sinp * sp.a.sint * sp.a.sinp +
                         a_coord.cost * sp.a.cost;
      if (dot > sp.cost) break;
    }
  } else {
    a_coord = random_sphere_angle3d(g.rng);
  }
  double phi_local = 0.0 + (TWOPI_F - 0.0) * g.rng.random();  // random_uni(phi_local, zero, twopi)
  Angle a_local;
  a_local.cosp = std::cos(phi_local);
  a_local.sinp = std::sin(phi_local);
  if (src.limb_darkening) {
    a_local.cost = ran_mu_limb(g.rng, 1.5, 1.0);
  } else {
    double xi = g.rng.random();
    a_local.cost = std::sqrt(xi);
  }
  a_local.sint = std::sqrt(1.0 - a_local.cost * a_local.cost);
  p.a = rotate_angle3d(a_local, a_coord);
  p.s = Stokes{1.0, 0.0, 0.0, 0.0};
  Vec u = angle3d_to_vector3d(a_coord);
  p.r = Vec{u.x * src.radius, u.y * src.radius, u.z * src.radius};
  p.r = Vec{p.r.x + src.position.x, p.r.y + src.position.y, p.r.z + src.position.z};
  p.last_isotropic = false;
  p.source_a = a_coord;
}

// emit_from_point (source_type.f90:539-564)
void emit_from_point(orc_ctx &g, const Source &src, Photon &p) {
  p.r = src.position;
  p.a = random_sphere_angle3d(g.rng);
  p.s = Stokes{1.0, 0.0, 0.0, 0.0};
  p.last_isotropic = true;
}

// Fortran MODULO(a, p) for reals: a - floor(a / p) * p (result has the sign of p)
static double f_modulo(double a, double p) { return a - std::floor(a / p) * p; }

// minus_angle_dp (type_angle3d.f90:443-447)
static Angle minus_angle(const Angle &a) { return Angle{-a.cost, a.sint, -a.cosp, -a.sinp}; }

// emit_from_point_collection (source_type.f90:570-598)
void emit_from_point_collection(orc_ctx &g, const Source &src, Photon &p) {
  const int i_source = src.collection_pdf.sample(g.rng);
  p.r = src.position_collection[i_source - 1];
  p.a = random_sphere_angle3d(g.rng);
  p.s = Stokes{1.0, 0.0, 0.0, 0.0};
  p.last_isotropic = true;
}

// emit_from_extern_sph (source_type.f90:748-802): from the sphere inwards, cosine law about the inward normal
void emit_from_extern_sph(orc_ctx &g, const Source &src, Photon &p) {
  Angle a_coord = random_sphere_angle3d(g.rng);
  double phi_local = 0.0 + (TWOPI_F - 0.0) * g.rng.random();
  Angle a_local;
  a_local.cosp = std::cos(phi_local);
  a_local.sinp = std::sin(phi_local);
  double xi = g.rng.random();
  a_local.cost = std::sqrt(xi);
  a_local.sint = std::sqrt(1.0 - a_local.cost * a_local.cost);
  p.a = rotate_angle3d(a_local, a_coord);
  p.a = minus_angle(p.a);
  p.s = Stokes{1.0, 0.0, 0.0, 0.0};
  Vec u = angle3d_to_vector3d(a_coord);
  p.r = Vec{u.x * src.radius, u.y * src.radius, u.z * src.radius};
  p.r = Vec{p.r.x + src.position.x, p.r.y + src.position.y, p.r.z + src.position.z};
  p.last_isotropic = false;
  p.source_a = minus_angle(a_coord);
}

// normal of face `face` of an external box pointing into the box (source_type.f90:866-896, 916-929)
static Angle extern_box_normal(int face) {
  switch (face) {
    case 1: return Angle{0.0, 1.0, 1.0, 0.0};
    case 2: return Angle{0.0, -1.0, 1.0, 0.0};
    case 3: return Angle{0.0, 1.0, 0.0, 1.0};
    case 4: return Angle{0.0, -1.0, 0.0, 1.0};
    case 5: return Angle{1.0, 0.0, 1.0, 0.0};
    default: return Angle{-1.0, 0.0, 1.0, 0.0};
  }
}

// emit_from_extern_box (source_type.f90:822-904)
void emit_from_extern_box(orc_ctx &g, const Source &src, Photon &p) {
  const int face = src.face.sample(g.rng);
  double phi_local = 0.0 + (TWOPI_F - 0.0) * g.rng.random();
  Angle a_local;
  a_local.cosp = std::cos(phi_local);
  a_local.sinp = std::sin(phi_local);
  double xi = g.rng.random();
  a_local.cost = std::sqrt(xi);
  a_local.sint = std::sqrt(1.0 - a_local.cost * a_local.cost);
  auto uni = [&](double a, double b) { return a + (b - a) * g.rng.random(); };
  switch (face) {
    case 1:
      p.r.x = src.xmin; p.r.y = uni(src.ymin, src.ymax); p.r.z = uni(src.zmin, src.zmax);
      break;
    case 2:
      p.r.x = src.xmax; p.r.y = uni(src.ymin, src.ymax); p.r.z = uni(src.zmin, src.zmax);
      break;
    case 3:
      p.r.x = uni(src.xmin, src.xmax); p.r.y = src.ymin; p.r.z = uni(src.zmin, src.zmax);
      break;
    case 4:
      p.r.x = uni(src.xmin, src.xmax); p.r.y = src.ymax; p.r.z = uni(src.zmin, src.zmax);
      break;
    case 5:
      p.r.x = uni(src.xmin, src.xmax); p.r.y = uni(src.ymin, src.ymax); p.r.z = src.zmin;
      break;
    default:
      p.r.x = uni(src.xmin, src.xmax); p.r.y = uni(src.ymin, src.ymax); p.r.z = src.zmax;
      break;
  }
  p.a = rotate_angle3d(a_local, extern_box_normal(face));
  p.s = Stokes{1.0, 0.0, 0.0, 0.0};
  p.last_isotropic = false;
  p.face_id = face;
}

// emit_from_plane_parallel (source_type.f90:935-980): a beam of radius `radius` travelling along `direction`
void emit_from_plane_parallel(orc_ctx &g, const Source &src, Photon &p) {
  double xi = g.rng.random();
  double r = std::pow(xi, 0.5) * src.radius;
  double phi = 0.0 + (360.0 - 0.0) * g.rng.random();
  Angle a_local = angle3d_deg(90.0, phi);
  Angle a_final = rotate_angle3d(a_local, src.direction);
  Vec u = angle3d_to_vector3d(a_final);
  p.r = Vec{u.x * r, u.y * r, u.z * r};
  p.r = Vec{p.r.x + src.position.x, p.r.y + src.position.y, p.r.z + src.position.z};
  p.a = src.direction;
  p.s = Stokes{1.0, 0.0, 0.0, 0.0};
  p.last_isotropic = false;
}

// new_grid_cell(ic, geo) + random_position_cell of the geometry module for the 1-based cell id `ic`
void place_at_random_position_in_cell(orc_ctx &g, Photon &p, int ic) {
  if (g.grid_type == 5) {
    p.icell = Cell{0, 0, 0, ic};
    p.r = vor_random_position_cell(g, p.icell);
  } else if (g.grid_type == 3) {
    p.icell = Cell{0, 0, 0, ic};
    p.r = oct_random_position_cell(g, p.icell);
  } else if (g.grid_type == 4) {
    p.icell = amr_cell_1d(g, ic);
    p.r = amr_random_position_cell(g, p.icell);
  } else {
    int i3 = (ic - 1) / (g.n1 * g.n2) + 1;
    int i2 = (ic - 1 - (i3 - 1) * g.n1 * g.n2) / g.n1 + 1;
    int i1 = ic - (i3 - 1) * g.n1 * g.n2 - (i2 - 1) * g.n1;
    p.icell = new_grid_cell(g, i1, i2, i3);
    if (g.grid_type == 1) {
      p.r = sph_random_position_cell(g, p.icell);
    } else if (g.grid_type == 2) {
      p.r = cyl_random_position_cell(g, p.icell);
    } else {
      // grid_geometry_cartesian_3d.f90:383-394
      double x = g.rng.random(), y = g.rng.random(), z = g.rng.random();
      p.r.x = x * (g.w1[i1] - g.w1[i1 - 1]) + g.w1[i1 - 1];
      p.r.y = y * (g.w2[i2] - g.w2[i2 - 1]) + g.w2[i2 - 1];
      p.r.z = z * (g.w3[i3] - g.w3[i3 - 1]) + g.w3[i3 - 1];
    }
  }
}

// emit_from_map (source_type.f90:713-746): cell from the luminosity map (grid_sample_pdf_map,
// grid_geometry_common_3d.f90:65-71), uniform position inside it, isotropic direction
void emit_from_map(orc_ctx &g, const Source &src, Photon &p) {
  const int ic = src.luminosity_map.sample(g.rng);
  p.in_cell = true;
  place_at_random_position_in_cell(g, p, ic);
  p.a = random_sphere_angle3d(g.rng);
  p.s = Stokes{1.0, 0.0, 0.0, 0.0};
  p.last_isotropic = true;
}

// select_dust_specific_energy_rho (grid_physics_3d.f90:101-109): always draws, even for one dust type
int select_dust_specific_energy_rho(orc_ctx &g, const Cell &c) {
  for (int id = 0; id < g.n_dust; id++) {
    size_t k = (size_t)id * g.n_cells + c.ic - 1;
    g.absorption.pdf[id] = g.specific_energy[k] * g.density[k];
  }
  g.absorption.find_cdf();
  return g.absorption.sample(g.rng);
}

// dust_sample_emit_probability (dust_type_4elem.f90:356-377)
double dust_sample_emit_probability(const Dust &d, int jnu_var_id, double jnu_var_frac, double nu) {
  const double prob1 = d.j_nu[jnu_var_id - 1].interpolate(nu);
  const double prob2 = d.j_nu[jnu_var_id].interpolate(nu);
  if (prob1 == 0.0 || prob2 == 0.0) return 0.0;
  const double l = std::log10(prob1) + jnu_var_frac * (std::log10(prob2) - std::log10(prob1));
  return std::pow(10.0, l);
}

double normalized_B_nu(double nu, double T);

// emit (source.f90:100-179) + source_emit (source_type.f90:398-511)
void emit(orc_ctx &g, Photon &p, bool reemit = false, int reemit_id = 0, double reemit_energy = 0.0, int inu = 0) {
  p = Photon();
  int n_sources = (int)g.s.size();
  if (n_sources == 0) throw OracleError{"no sources to emit from"};
  if (n_sources > 1) {
    if (g.conf.sample_sources_evenly) {
      double xi = g.rng.random();
      p.source_id = (int)(xi * n_sources) + 1;
    } else {
      p.source_id = g.luminosity.sample(g.rng);
    }
  } else {
    p.source_id = 1;
  }
  if (reemit) p.source_id = reemit_id;  // source.f90:134-140
  const Source &src = g.s[p.source_id - 1];
  int ispot = 0;
  switch (src.type) {
    case HYP_SOURCE_POINT:
      emit_from_point(g, src, p);
      break;
    case HYP_SOURCE_SPHERE:
      if (!src.spot.empty()) {
        // source type 3 (source_type.f90:421-427)
        ispot = src.spot_pdf.sample(g.rng);
        if (ispot == (int)src.spot.size() + 1)
          emit_from_sphere(g, src, p);
        else
          emit_from_sphere(g, src, p, ispot);
      } else {
        emit_from_sphere(g, src, p);
      }
      break;
    case HYP_SOURCE_EXTERN_SPH:
      emit_from_extern_sph(g, src, p);
      break;
    case HYP_SOURCE_EXTERN_BOX:
      emit_from_extern_box(g, src, p);
      break;
    case HYP_SOURCE_PLANE_PARALLEL:
      emit_from_plane_parallel(g, src, p);
      break;
    case HYP_SOURCE_POINT_COLLECTION:
      emit_from_point_collection(g, src, p);
      break;
    case HYP_SOURCE_MAP:
      emit_from_map(g, src, p);
      break;
    default:
      throw OracleError{"source type not restated in the oracle"};
  }
  p.energy = 1.0;
  if (inu > 0) {
    // the frequency is fixed; the energy reflects the probability of emission there (source_type.f90:436-474)
    const double nu = g.frequencies[inu - 1];
    p.nu = nu;
    p.inu = inu;
    if (ispot >= 1 && ispot <= (int)src.spot.size()) {
      const Source::Spot &sp = src.spot[ispot - 1];
      if (sp.freq_type == 1)
        p.energy = sp.spectrum.interpolate(nu);
      else if (sp.freq_type == 2)
        p.energy = normalized_B_nu(nu, sp.temperature);
      else
        throw OracleError{"Spot cannot have LTE spectrum"};
    } else if (src.freq_type == 1) {
      p.energy = src.spectrum.interpolate(nu);
    } else if (src.freq_type == 2) {
      p.energy = normalized_B_nu(nu, src.temperature);
    } else if (src.freq_type == 3) {
      p.dust_id = select_dust_specific_energy_rho(g, p.icell);
      size_t k = (size_t)(p.dust_id - 1) * g.n_cells + p.icell.ic - 1;
      p.emiss_var_id = g.jnu_var_id[k];
      p.emiss_var_frac = g.jnu_var_frac[k];
      p.energy = dust_sample_emit_probability(g.d[p.dust_id - 1], p.emiss_var_id, p.emiss_var_frac, nu);
    } else {
      throw OracleError{"unknown spectrum type"};
    }
  } else
  if (ispot >= 1 && ispot <= (int)src.spot.size()) {
    // the spot's own spectrum; tabulated ones are sampled with sample_pdf_log (source_type.f90:480-486)
    const Source::Spot &sp = src.spot[ispot - 1];
    if (sp.freq_type == 1)
      p.nu = sp.spectrum.sample_log(g.rng);
    else if (sp.freq_type == 2)
      p.nu = g.rng.random_planck_frequency(sp.temperature);
    else
      throw OracleError{"Spot cannot have LTE spectrum"};
  } else if (src.freq_type == 1) {
    p.nu = src.spectrum.sample(g.rng);
  } else if (src.freq_type == 2) {
    p.nu = g.rng.random_planck_frequency(src.temperature);
  } else if (src.freq_type == 3) {
    // LTE spectrum of the dust in the cell (source_type.f90:500-505)
    p.dust_id = select_dust_specific_energy_rho(g, p.icell);
    size_t k = (size_t)(p.dust_id - 1) * g.n_cells + p.icell.ic - 1;
    p.emiss_var_id = g.jnu_var_id[k];
    p.emiss_var_frac = g.jnu_var_frac[k];
    p.nu = g.d[p.dust_id - 1].sample_j_nu(g.rng, p.emiss_var_id, p.emiss_var_frac);
  } else {
    throw OracleError{"unknown spectrum type"};
  }
  p.v = angle3d_to_vector3d(p.a);
  if (reemit) {
    p.energy = reemit_energy;
  } else {
    if (inu > 0) p.energy = p.energy * g.energy_total;   // source.f90:161
    if (g.conf.sample_sources_evenly) p.energy = p.energy * g.luminosity.pdf[p.source_id - 1] * n_sources;
    g.energy_current = g.energy_current + p.energy;
  }
  update_optconsts(g, p);
  p.emiss_type = src.freq_type;
  p.last[0] = 's';
  p.last[1] = 'r';
  g.photon_counter++;
  p.id = g.photon_counter;
  place_in_cell(g, p);
  if (p.killed)
    throw OracleError{
        "photon was not emitted inside a cell - this usually indicates that a source is not inside the grid"};
}

// grid_integrate (grid_propagate_3d.f90:35-234)
void grid_integrate(orc_ctx &g, Photon &p, double tau_required, double &tau_achieved) {
  const double frac_check = g.conf.propagation_check_frequency;
  tau_achieved = 0.0;
  // grid_propagate_3d.f90:61-71: the frequency bin of the packet, found once per call; -1 outside the edges
  int idx = -1;
  if (g.n_nu_bins > 0) idx = locate(g.log_nu_bin_edges.data(), g.n_nu_bins + 1, std::log10(p.nu));
  const size_t nsp = (size_t)g.n_dust * g.n_cells;
  if (!p.in_cell) throw OracleError{"photon has not been placed in a cell"};
  if (escaped(g, p.icell)) return;
  // grid_propagate_3d.f90:90-95: a packet counts once per cell for as long as no other packet enters it
  auto count_visit = [&]() {
    if (g.n_photons.empty()) return;
    if (g.last_photon_id[p.icell.ic - 1] != p.id) {
      g.n_photons[p.icell.ic - 1]++;
      g.last_photon_id[p.icell.ic - 1] = p.id;
    }
  };
  count_visit();
  if (tau_required == 0.0) return;
  double t_source;
  int source_id;
  find_nearest_source(g, p.r, p.v, t_source, source_id);
  g.radial = (p.r.x * p.v.x + p.r.y * p.v.y + p.r.z * p.v.z) > 0.;  // grid_propagate_3d.f90:73,262
  double t_achieved = 0.0;
  const int nc = g.n_cells;
  for (;;) {
    double xi = g.rng.random();
    if (xi < frac_check) {
      if (!in_correct_cell(g, p)) {
        g.killed_photons_geo++;
        p.killed = true;
        return;
      }
    }
    double tau_needed = tau_required - tau_achieved;
    double tmin;
    WallId id_min;
    find_wall(g, p, tmin, id_min);
    if (id_min.w1 == 0 && id_min.w2 == 0 && id_min.w3 == 0) {
      g.killed_photons_geo++;
      p.killed = true;
      return;
    }
    // density(p%icell%ic, id): the STORED 1-D id.  adjust_wall changes i1/i2/i3 without
    // refreshing ic (grid_geometry_cartesian_3d.f90:184-232), so for a photon emitted exactly
    // on a wall the first segment reads/deposits in the cell find_cell returned.  Restated as is.
    int ic = p.icell.ic;
    double chi_rho_total = 0.0;
    for (int id = 0; id < g.n_dust; id++)
      chi_rho_total = chi_rho_total + p.current_chi[id] * g.density[(size_t)id * nc + ic - 1];
    double tau_cell = chi_rho_total * tmin;
    g.n_crossings++;
    if (tau_cell < tau_needed) {
      t_achieved = t_achieved + tmin;
      if (t_achieved > t_source) {
        p.reabsorbed = true;
        p.reabsorbed_id = source_id;
        return;
      }
      p.r.x = p.r.x + tmin * p.v.x;
      p.r.y = p.r.y + tmin * p.v.y;
      p.r.z = p.r.z + tmin * p.v.z;
      tau_achieved = tau_achieved + tau_cell;
      for (int id = 0; id < g.n_dust; id++) {
        size_t k = (size_t)id * nc + ic - 1;
        if (g.density[k] > 0.0) {
          g.specific_energy_sum[k] = g.specific_energy_sum[k] + tmin * p.current_kappa[id] * p.energy;
          if (idx > 0) {   // :155-158
            double &b = g.specific_energy_sum_spectrum[(size_t)(idx - 1) * nsp + k];
            b = b + tmin * p.current_kappa[id] * p.energy;
          }
        }
      }
      p.on_wall = true;
      p.icell = next_cell(g, p.icell, id_min, p.r);
      p.on_wall_id = WallId{-id_min.w1, -id_min.w2, -id_min.w3};  // opposite_wall (grid_geometry_common_3d.f90:40-45)
      if (escaped(g, p.icell)) return;
      count_visit();   // grid_propagate_3d.f90:175-180
    } else {
      double tact = tmin * (tau_needed / tau_cell);
      t_achieved = t_achieved + tact;
      if (t_achieved > t_source) {
        p.reabsorbed = true;
        p.reabsorbed_id = source_id;
        return;
      }
      p.r.x = p.r.x + tact * p.v.x;
      p.r.y = p.r.y + tact * p.v.y;
      p.r.z = p.r.z + tact * p.v.z;
      tau_achieved = tau_achieved + tau_needed;
      p.on_wall = false;
      p.on_wall_id = WallId();
      for (int id = 0; id < g.n_dust; id++) {
        size_t k = (size_t)id * nc + ic - 1;
        if (g.density[k] > 0.0)
          g.specific_energy_sum[k] = g.specific_energy_sum[k] + tact * p.current_kappa[id] * p.energy;
      }
      if (idx > 0)   // :217-225
        for (int id = 0; id < g.n_dust; id++) {
          size_t k = (size_t)id * nc + ic - 1;
          if (g.density[k] > 0.0) {
            double &b = g.specific_energy_sum_spectrum[(size_t)(idx - 1) * nsp + k];
            b = b + tact * p.current_kappa[id] * p.energy;
          }
        }
      return;
    }
  }
}

// select_dust_chi_rho (grid_physics_3d.f90:87-99)
int select_dust_chi_rho(orc_ctx &g, const Photon &p) {
  if (g.n_dust == 1) return 1;
  int ic = p.icell.ic;
  for (int id = 0; id < g.n_dust; id++)
    g.absorption.pdf[id] = p.current_chi[id] * g.density[(size_t)id * g.n_cells + ic - 1];
  g.absorption.find_cdf();
  return g.absorption.sample(g.rng);
}

// interact (dust_interact.f90:22-79)
void interact(orc_ctx &g, Photon &p, bool force_scatter = false) {
  int id = select_dust_chi_rho(g, p);
  double albedo = p.current_albedo[id - 1];
  p.a_prev = p.a;
  p.v_prev = p.v;
  p.s_prev = p.s;
  // dust_interact.f90:47-52: a forced scattering draws no number
  double xi = force_scatter ? 0.0 : g.rng.random();
  const Dust &d = g.d[id - 1];
  if (xi > albedo) {
    // dust_emit (dust_type_4elem.f90:334-354)
    int ic = p.icell.ic;
    size_t k = (size_t)(id - 1) * g.n_cells + ic - 1;
    p.nu = d.sample_j_nu(g.rng, g.jnu_var_id[k], g.jnu_var_frac[k]);
    p.s = Stokes{1.0, 0.0, 0.0, 0.0};
    p.a = random_sphere_angle3d(g.rng);
    p.energy = p.energy * 1.0;
    update_optconsts(g, p);
    p.scattered = false;
    p.reprocessed = true;
    p.last_isotropic = true;
    p.dust_id = id;
    p.last[0] = 'd';
    p.last[1] = 'e';
    g.n_absorptions++;
  } else {
    dust_scatter(d, g.rng, p.nu, p.a, p.s);
    p.scattered = true;
    p.last_isotropic = false;
    p.dust_id = id;
    p.last[0] = 'd';
    p.last[1] = 's';
    p.n_scat = p.n_scat + 1;
    g.n_scatterings++;
  }
  p.v = angle3d_to_vector3d(p.a);
  if (force_scatter) p.energy = p.energy * albedo;   // dust_interact.f90:75-77
}

// update_energy_abs_tot (grid_physics_3d.f90:605-611)
void update_energy_abs_tot(orc_ctx &g) {
  for (int id = 0; id < g.n_dust; id++) {
    double sum = 0.0;
    for (int ic = 0; ic < g.n_cells; ic++) {
      size_t k = (size_t)id * g.n_cells + ic;
      sum = sum + g.specific_energy[k] * g.density[k] * g.volume[ic];
    }
    g.energy_abs_tot[id] = sum;
  }
}

// check_energy_abs (grid_physics_3d.f90:555-603)
void check_energy_abs(orc_ctx &g) {
  for (int id = 0; id < g.n_dust; id++) {
    const Dust &d = g.d[id];
    double *e = &g.specific_energy[(size_t)id * g.n_cells];
    double mn = g.minimum_specific_energy[id];
    for (int ic = 0; ic < g.n_cells; ic++)
      if (e[ic] < mn) e[ic] = mn;
    if (g.conf.enforce_energy_range) {
      double lo = d.specific_energy[0], hi = d.specific_energy[d.n_e - 1];
      for (int ic = 0; ic < g.n_cells; ic++)
        if (e[ic] < lo) e[ic] = lo;
      for (int ic = 0; ic < g.n_cells; ic++)
        if (e[ic] > hi) e[ic] = hi;
    }
  }
  update_energy_abs_tot(g);
}

// update_energy_abs (grid_physics_3d.f90:500-553)
void update_energy_abs(orc_ctx &g, double scale) {
  for (int id = 0; id < g.n_dust; id++)
    for (int ic = 0; ic < g.n_cells; ic++) {
      size_t k = (size_t)id * g.n_cells + ic;
      g.specific_energy[k] = g.specific_energy_sum[k] * scale / g.volume[ic];
      if (g.volume[ic] == 0.0) g.specific_energy[k] = 0.0;
      for (int ib = 0; ib < g.n_nu_bins; ib++) {   // :516-524
        const size_t kb = (size_t)ib * g.n_dust * g.n_cells + k;
        g.specific_energy_spectrum[kb] = g.specific_energy_sum_spectrum[kb] * scale / g.volume[ic];
        if (g.volume[ic] == 0.0) g.specific_energy_spectrum[kb] = 0.0;
      }
    }
  // (specific_energy_additional_spectrum is a copy of the spectrum array at a moment when it is all zeros,
  // grid_physics_3d.f90:143,223-225: adding it, :543-545, changes nothing)
  // additional source of heating (grid_physics_3d.f90:537-545)
  if (!g.specific_energy_additional.empty())
    for (size_t k = 0; k < g.specific_energy.size(); k++)
      g.specific_energy[k] = g.specific_energy[k] + g.specific_energy_additional[k];
  update_energy_abs_tot(g);
  check_energy_abs(g);
}

// scale_specific_energy_spectrum (grid_physics_3d.f90:350-365); ic 0-based
void scale_specific_energy_spectrum(orc_ctx &g, int ic, int id, double factor) {
  for (int ib = 0; ib < g.n_nu_bins; ib++) {
    double &b = g.specific_energy_spectrum[((size_t)ib * g.n_dust + id) * g.n_cells + ic];
    b = b * factor;
  }
}

// sublimate_dust (grid_physics_3d.f90:420-498)
void sublimate_dust(orc_ctx &g) {
  for (int id = 0; id < g.n_dust; id++) {
    const Dust &d = g.d[id];
    double *e = &g.specific_energy[(size_t)id * g.n_cells];
    double *rho = &g.density[(size_t)id * g.n_cells];
    switch (d.sublimation_mode) {
      case 1:
        for (int ic = 0; ic < g.n_cells; ic++)
          if (e[ic] > d.sublimation_specific_energy) {
            rho[ic] = 0.;
            e[ic] = g.minimum_specific_energy[id];
            for (int ib = 0; ib < g.n_nu_bins; ib++)   // :442-446
              g.specific_energy_spectrum[((size_t)ib * g.n_dust + id) * g.n_cells + ic] = g.minimum_specific_energy[id];
          }
        break;
      case 2:
        for (int ic = 0; ic < g.n_cells; ic++)
          if (e[ic] > d.sublimation_specific_energy) {
            double cr1 = interp1d_loglog(d.specific_energy.data(), d.chi_rosseland.data(), d.n_e, e[ic]);
            double cr2 = interp1d_loglog(d.specific_energy.data(), d.chi_rosseland.data(), d.n_e,
                                         d.sublimation_specific_energy);
            double q = cr1 / cr2;
            rho[ic] = rho[ic] * d.sublimation_specific_energy / e[ic] * (q * q);
            scale_specific_energy_spectrum(g, ic, id, d.sublimation_specific_energy / e[ic]);
            e[ic] = d.sublimation_specific_energy;
          }
        break;
      case 3:
        for (int ic = 0; ic < g.n_cells; ic++)
          if (e[ic] > d.sublimation_specific_energy) {
            scale_specific_energy_spectrum(g, ic, id, d.sublimation_specific_energy / e[ic]);
            e[ic] = d.sublimation_specific_energy;
          }
        break;
      default:
        break;
    }
  }
  update_energy_abs_tot(g);
  check_energy_abs(g);
}

// ---------------------------------------------------------------------------
// Partial diffusion approximation (src/grid/grid_pda_3d.f90, grid_pda_{cartesian,spherical,cylindrical}_3d.f90)
// ---------------------------------------------------------------------------
double kappa_planck(const Dust &d, double e) { return interp1d_loglog(d.specific_energy.data(), d.kappa_planck.data(), d.n_e, e); }
double chi_rosseland(const Dust &d, double e) { return interp1d_loglog(d.specific_energy.data(), d.chi_rosseland.data(), d.n_e, e); }

struct PdaCell {
  int i1, i2, i3, ic;  // 1-based, as type_grid_cell
};

int pda_cell_id(const orc_ctx &g, int i1, int i2, int i3) { return (i3 - 1) * g.n1 * g.n2 + (i2 - 1) * g.n1 + i1; }

// cell_width (grid_geometry_cartesian_3d.f90:49-61, _spherical_3d.f90:60-72, _cylindrical_3d.f90:60-72)
double pda_cell_width(const orc_ctx &g, const PdaCell &c, int idir) {
  auto mid_log = [](const std::vector<double> &w, int i) {   // geo%r / geo%w (:139-143 / :130-134)
    return w[i - 1] == 0.0 ? w[i] / 2.0 : std::pow(10.0, (std::log10(w[i - 1]) + std::log10(w[i])) / 2.0);
  };
  if (g.grid_type == 0) {
    const std::vector<double> &w = idir == 1 ? g.w1 : (idir == 2 ? g.w2 : g.w3);
    const int i = idir == 1 ? c.i1 : (idir == 2 ? c.i2 : c.i3);
    return w[i] - w[i - 1];
  }
  if (g.grid_type == 1) {
    const double r = mid_log(g.w1, c.i1);
    if (idir == 1) return g.w1[c.i1] - g.w1[c.i1 - 1];
    if (idir == 2) return r * (g.w2[c.i2] - g.w2[c.i2 - 1]);
    const double t = (g.w2[c.i2 - 1] + g.w2[c.i2]) / 2.0;
    return r * std::sin(t) * (g.w3[c.i3] - g.w3[c.i3 - 1]);
  }
  if (idir == 1) return g.w1[c.i1] - g.w1[c.i1 - 1];
  if (idir == 2) return g.w2[c.i2] - g.w2[c.i2 - 1];
  return mid_log(g.w1, c.i1) * (g.w3[c.i3] - g.w3[c.i3 - 1]);
}

// geometrical_factor (grid_pda_*_3d.f90)
double pda_geometrical_factor(const orc_ctx &g, int wall, const PdaCell &c) {
  if (g.grid_type == 1) {
    const double a = g.w1[c.i1 - 1], b = g.w1[c.i1];
    if (wall == 1) return 4. * a * a / ((a + b) * (a + b));
    if (wall == 2) return 4. * b * b / ((a + b) * (a + b));
    const double sa = std::sin(g.w2[c.i2 - 1]), sb = std::sin(g.w2[c.i2]);
    if (wall == 3) return 2. * sa / (sa + sb);
    if (wall == 4) return 2. * sb / (sa + sb);
    return 1.0;
  }
  if (g.grid_type == 2) {
    const double a = g.w1[c.i1 - 1], b = g.w1[c.i1];
    if (wall == 1) return 2. * a / (a + b);
    if (wall == 2) return 2. * b / (a + b);
    return 1.0;
  }
  return 1.0;
}

// next_cell(cell, wall) (next_cell_int; the polar grids wrap in phi)
PdaCell pda_next_cell(const orc_ctx &g, const PdaCell &c, int wall) {
  PdaCell n = c;
  switch (wall) {
    case 1: n.i1--; break;
    case 2: n.i1++; break;
    case 3: n.i2--; break;
    case 4: n.i2++; break;
    case 5:
      n.i3--;
      if (g.grid_type != 0 && n.i3 == 0) n.i3 = g.n3;
      break;
    default:
      n.i3++;
      if (g.grid_type != 0 && n.i3 == g.n3 + 1) n.i3 = 1;
  }
  n.ic = pda_cell_id(g, n.i1, n.i2, n.i3);
  return n;
}

struct PdaState {
  std::vector<double> e_mean;
  int n_dim = 3;
};

double pda_density_sum(const orc_ctx &g, int ic) {
  double s = 0.0;
  for (int id = 0; id < g.n_dust; id++) s = s + g.density[(size_t)id * g.n_cells + ic - 1];
  return s;
}

// update_e_mean (grid_pda_3d.f90:92-103)
void pda_update_e_mean(const orc_ctx &g, PdaState &P, int ic) {
  P.e_mean[ic - 1] = 0.;
  const double rs = pda_density_sum(g, ic);
  if (rs > 0.0) {
    for (int id = 0; id < g.n_dust; id++) {
      const size_t k = (size_t)id * g.n_cells + ic - 1;
      P.e_mean[ic - 1] = P.e_mean[ic - 1] + g.density[k] * g.specific_energy[k] / kappa_planck(g.d[id], g.specific_energy[k]);
    }
    P.e_mean[ic - 1] = P.e_mean[ic - 1] / rs;
  }
}

// update_specific_energy (grid_pda_3d.f90:52-90)
void pda_update_specific_energy(orc_ctx &g, const PdaState &P, int ic) {
  for (int id = 0; id < g.n_dust; id++) {
    const Dust &d = g.d[id];
    const size_t k = (size_t)id * g.n_cells + ic - 1;
    double s = g.specific_energy[k];
    const double s_old = s;
    const double smin = d.specific_energy[0], smax = d.specific_energy[d.n_e - 1];
    if (P.e_mean[ic - 1] < smin / kappa_planck(d, smin)) {
      s = smin;
    } else if (P.e_mean[ic - 1] > smax / kappa_planck(d, smax)) {
      s = smax;
    } else {
      for (;;) {
        const double s_prev = s;
        s = P.e_mean[ic - 1] * kappa_planck(d, s);
        if (std::max(s / s_prev, s_prev / s) - 1.0 < 1.e-5) break;
      }
    }
    g.specific_energy[k] = s;
    if (s_old > 0.0) scale_specific_energy_spectrum(g, ic - 1, id, s / s_old);   // :63-67
  }
}

// dtau_rosseland (grid_pda_3d.f90:171-181)
double pda_dtau_rosseland(const orc_ctx &g, const PdaCell &c, int idir) {
  double t = 0.0;
  for (int id = 0; id < g.n_dust; id++) {
    const size_t k = (size_t)id * g.n_cells + c.ic - 1;
    t = t + g.density[k] * chi_rosseland(g.d[id], g.specific_energy[k]) * pda_cell_width(g, c, idir);
  }
  return t;
}

// lineq_gausselim_dp (fortranlib/src/lib_algebra.f90:166-200); a is stored column-major, a(i, j) at [i + n * j]
void lineq_gausselim(std::vector<double> &a, std::vector<double> &b, int n) {
  auto A = [&](int i, int j) -> double & { return a[(size_t)(i - 1) + (size_t)n * (j - 1)]; };
  for (int i = 1; i <= n - 1; i++) {
    if (A(i, i) == 0.0) throw OracleError{"Zero pivot value"};
    for (int j = i + 1; j <= n; j++)
      if (A(i, j) != 0.0) {
        const double frac = A(i, j) / A(i, i);
        b[j - 1] = b[j - 1] - frac * b[i - 1];
        for (int k = i; k <= n; k++) A(k, j) = A(k, j) - frac * A(k, i);
      }
  }
  for (int i = n; i >= 2; i--)
    for (int j = i - 1; j >= 1; j--)
      if (A(i, j) != 0.0) {
        const double frac = A(i, j) / A(i, i);
        b[j - 1] = b[j - 1] - frac * b[i - 1];
        for (int k = i; k <= n; k++) A(k, j) = A(k, j) - frac * A(k, i);
      }
  for (int i = 1; i <= n; i++) b[i - 1] = b[i - 1] / A(i, i);
}

// solve_pda_indiv_exact (grid_pda_3d.f90:183-248)
void solve_pda_indiv_exact(orc_ctx &g, PdaState &P, const std::vector<PdaCell> &cells, const std::vector<int> &id_pda_cell) {
  const int n = (int)cells.size();
  for (const PdaCell &c : cells) pda_update_e_mean(g, P, c.ic);
  std::vector<double> a((size_t)n * n, 0.0), b(n, 0.0);
  for (int id_curr = 1; id_curr <= n; id_curr++) {
    const PdaCell &curr = cells[id_curr - 1];
    for (int wall = 1; wall <= P.n_dim * 2; wall++) {
      const int direction = (wall + 1) / 2;
      const PdaCell next = pda_next_cell(g, curr, wall);
      double dtau_sum = pda_dtau_rosseland(g, curr, direction) + pda_dtau_rosseland(g, next, direction);
      if (dtau_sum < 1e-100) dtau_sum = 1e-100;
      double coefficient = 1. / dtau_sum / pda_cell_width(g, curr, direction);
      coefficient = coefficient * pda_geometrical_factor(g, wall, curr);
      a[(size_t)(id_curr - 1) + (size_t)n * (id_curr - 1)] -= coefficient;
      if (id_pda_cell[next.ic - 1] > 0) {
        const int id_next = id_pda_cell[next.ic - 1];
        a[(size_t)(id_next - 1) + (size_t)n * (id_curr - 1)] += coefficient;
      } else {
        b[id_curr - 1] = b[id_curr - 1] - coefficient * P.e_mean[next.ic - 1];
      }
    }
  }
  lineq_gausselim(a, b, n);
  for (int id_curr = 1; id_curr <= n; id_curr++) {
    const int ic = cells[id_curr - 1].ic;
    P.e_mean[ic - 1] = b[id_curr - 1];
    pda_update_specific_energy(g, P, ic);
  }
}

// solve_pda_indiv_iterative (grid_pda_3d.f90:250-325): Gauss-Seidel sweeps in cell order
void solve_pda_indiv_iterative(orc_ctx &g, PdaState &P, const std::vector<PdaCell> &cells) {
  const double tolerance_iter = 1.e-4;
  for (const PdaCell &c : cells) pda_update_e_mean(g, P, c.ic);
  for (;;) {
    double max_e_diff = 0.;
    for (const PdaCell &curr : cells) {
      double a = 0.0, b = 0.0;
      for (int wall = 1; wall <= P.n_dim * 2; wall++) {
        const int direction = (wall + 1) / 2;
        const PdaCell next = pda_next_cell(g, curr, wall);
        double coefficient = 1. / (pda_dtau_rosseland(g, curr, direction) + pda_dtau_rosseland(g, next, direction)) /
                             pda_cell_width(g, curr, direction);
        coefficient = coefficient * pda_geometrical_factor(g, wall, curr);
        a = a - coefficient;
        b = b - coefficient * P.e_mean[next.ic - 1];
      }
      const double e_new = b / a;
      const double e_diff = std::fabs(e_new - P.e_mean[curr.ic - 1]) / P.e_mean[curr.ic - 1];
      if (e_diff > max_e_diff) max_e_diff = e_diff;
      P.e_mean[curr.ic - 1] = e_new;
    }
    if (max_e_diff < tolerance_iter) break;
  }
  for (const PdaCell &c : cells) pda_update_specific_energy(g, P, c.ic);
}

// solve_pda (grid_pda_3d.f90:105-169).  Returns the number of PDA cells; exact_limit is the reference's 10000.
int solve_pda(orc_ctx &g, int exact_limit = 10000) {
  if (g.grid_type > 2) throw OracleError{"PDA is not available for this grid type"};   // grid_pda_disabled.f90
  if (g.n_photons.empty()) throw OracleError{"n_photons array is not allocated"};
  const int nc = g.n_cells;
  double tot = 0.0;
  for (int64_t v : g.n_photons) tot = tot + (double)v;
  const double mean_n_photons = tot / (double)nc;
  const double threshold_pda = 0.005;
  const int64_t limit = std::max<int64_t>(30, (int64_t)std::ceil(threshold_pda * mean_n_photons));
  std::vector<char> do_pda(nc, 0);
  for (int ic = 1; ic <= nc; ic++) do_pda[ic - 1] = g.n_photons[ic - 1] < limit && pda_density_sum(g, ic) > 0.0;
  // check_allowed_pda: no PDA in the cells on the edge of the grid (the polar grids are periodic in phi)
  for (int i3 = 1; i3 <= g.n3; i3++)
    for (int i2 = 1; i2 <= g.n2; i2++)
      for (int i1 = 1; i1 <= g.n1; i1++)
        if (i1 == 1 || i1 == g.n1 || i2 == 1 || i2 == g.n2 || (g.grid_type == 0 && (i3 == 1 || i3 == g.n3)))
          do_pda[pda_cell_id(g, i1, i2, i3) - 1] = 0;
  int n_pda = 0;
  for (char c : do_pda) n_pda += c;
  if (n_pda == 0) return 0;
  const double tolerance = n_pda < exact_limit ? 1.e-5 : 1.e-4;
  PdaState P;
  P.n_dim = (g.grid_type != 0 && g.n3 == 1) ? 2 : 3;
  P.e_mean.assign(nc, 0.0);
  for (int ic = 1; ic <= nc; ic++) pda_update_e_mean(g, P, ic);
  std::vector<PdaCell> cells;
  std::vector<int> id_pda_cell(nc, -1);
  for (int i3 = 1; i3 <= g.n3; i3++)
    for (int i2 = 1; i2 <= g.n2; i2++)
      for (int i1 = 1; i1 <= g.n1; i1++) {
        const int ic = pda_cell_id(g, i1, i2, i3);
        if (do_pda[ic - 1]) {
          cells.push_back(PdaCell{i1, i2, i3, ic});
          id_pda_cell[ic - 1] = (int)cells.size();
        }
      }
  for (;;) {
    const std::vector<double> prev = g.specific_energy;
    if (n_pda < exact_limit)
      solve_pda_indiv_exact(g, P, cells, id_pda_cell);
    else
      solve_pda_indiv_iterative(g, P, cells);
    double maxdiff = 0.0;
    for (size_t k = 0; k < prev.size(); k++) maxdiff = std::max(maxdiff, std::fabs(g.specific_energy[k] - prev[k]) / prev[k]);
    if (maxdiff < tolerance) break;
  }
  update_energy_abs_tot(g);
  check_energy_abs(g);
  return n_pda;
}

// precompute_jnu_var (grid_physics_3d.f90:613-629)
void precompute_jnu_var(orc_ctx &g) {
  for (int ic = 0; ic < g.n_cells; ic++)
    for (int id = 0; id < g.n_dust; id++) {
      size_t k = (size_t)id * g.n_cells + ic;
      g.d[id].jnu_var_pos_frac(g.specific_energy[k], g.jnu_var_id[k], g.jnu_var_frac[k]);
    }
}

// ---------------------------------------------------------------------------
// modified random walk (src/grid/grid_mrw_3d.f90; Min et al. 2009, Robitaille 2010)
// ---------------------------------------------------------------------------
// distance_to_closest_wall of each geometry module
double distance_to_closest_wall(const orc_ctx &g, const Photon &p) {
  double d;
  if (g.grid_type == 0) {  // grid_geometry_cartesian_3d.f90:396-422
    double d1 = p.r.x - g.w1[p.icell.i1 - 1], d2 = g.w1[p.icell.i1] - p.r.x;
    double d3 = p.r.y - g.w2[p.icell.i2 - 1], d4 = g.w2[p.icell.i2] - p.r.y;
    double d5 = p.r.z - g.w3[p.icell.i3 - 1], d6 = g.w3[p.icell.i3] - p.r.z;
    d = std::min(std::min(std::min(d1, d2), std::min(d3, d4)), std::min(d5, d6));
  } else if (g.grid_type == 1) {  // grid_geometry_spherical_3d.f90:679-739
    double r = std::sqrt(p.r.x * p.r.x + p.r.y * p.r.y + p.r.z * p.r.z);
    double d1 = r - g.w1[p.icell.i1 - 1], d2 = g.w1[p.icell.i1] - r;
    if (std::fabs(d1) < g.ew1[p.icell.i1 - 1]) d1 = 0.0;
    if (std::fabs(d2) < g.ew1[p.icell.i1]) d2 = 0.0;
    double rcyl = std::sqrt(p.r.x * p.r.x + p.r.y * p.r.y);
    double ta = g.wtant[p.icell.i2 - 1], tb = g.wtant[p.icell.i2];
    double d3 = std::fabs(-rcyl + ta * p.r.z) / std::sqrt(1 + ta * ta);
    double d4 = std::fabs(-rcyl + tb * p.r.z) / std::sqrt(1 + tb * tb);
    double d5 = std::numeric_limits<double>::max(), d6 = d5;
    if (g.n3 > 1) {
      double pa = g.wtanp[p.icell.i3 - 1], pb = g.wtanp[p.icell.i3];
      d5 = std::fabs(pa * p.r.x - p.r.y) / std::sqrt(pa * pa + 1.0);
      d6 = std::fabs(pb * p.r.x - p.r.y) / std::sqrt(pb * pb + 1.0);
    }
    d = std::min(std::min(std::min(d1, d2), std::min(d3, d4)), std::min(d5, d6));
  } else if (g.grid_type == 2) {  // grid_geometry_cylindrical_3d.f90:549-590
    double r = std::sqrt(p.r.x * p.r.x + p.r.y * p.r.y);
    double d1 = r - g.w1[p.icell.i1 - 1], d2 = g.w1[p.icell.i1] - r;
    double d3 = p.r.z - g.w2[p.icell.i2 - 1], d4 = g.w2[p.icell.i2] - p.r.z;
    double d5 = std::numeric_limits<double>::max(), d6 = d5;
    if (g.n3 > 1) {
      double pa = g.wtanp[p.icell.i3 - 1], pb = g.wtanp[p.icell.i3];
      d5 = std::fabs(pa * p.r.x - p.r.y) / std::sqrt(pa * pa + 1.0);
      d6 = std::fabs(pb * p.r.x - p.r.y) / std::sqrt(pb * pb + 1.0);
    }
    d = std::min(std::min(std::min(d1, d2), std::min(d3, d4)), std::min(d5, d6));
  } else if (g.grid_type == 5) {
    throw OracleError{"not implemented for Voronoi grid"};   // distance_to_closest_wall, grid_geometry_voronoi.f90:314-320
  } else if (g.grid_type == 3) {  // grid_geometry_octree.f90:410-436
    const int k = p.icell.ic - 1;
    double d1 = p.r.x - g.ox[k] + g.odx[k], d2 = g.ox[k] + g.odx[k] - p.r.x;
    double d3 = p.r.y - g.oy[k] + g.ody[k], d4 = g.oy[k] + g.ody[k] - p.r.y;
    double d5 = p.r.z - g.oz[k] + g.odz[k], d6 = g.oz[k] + g.odz[k] - p.r.z;
    d = std::min(std::min(std::min(d1, d2), std::min(d3, d4)), std::min(d5, d6));
  } else {  // grid_geometry_amr.f90:743-773
    const orc_ctx::AmrGrid &G = g.levels[p.icell.ilevel - 1][p.icell.igrid - 1];
    double d1 = p.r.x - G.w1[p.icell.i1 - 1], d2 = G.w1[p.icell.i1] - p.r.x;
    double d3 = p.r.y - G.w2[p.icell.i2 - 1], d4 = G.w2[p.icell.i2] - p.r.y;
    double d5 = p.r.z - G.w3[p.icell.i3 - 1], d6 = G.w3[p.icell.i3] - p.r.z;
    d = std::min(std::min(std::min(d1, d2), std::min(d3, d4)), std::min(d5, d6));
  }
  if (d < 0.0) d = 0.0;
  return d;
}

// update_alpha_inv_planck (grid_physics_3d.f90:397-418) + prepare_mrw (grid_mrw_3d.f90:29-54)
void prepare_mrw(orc_ctx &g) {
  // initialize_cumulative (:158-196): P(y) = 2 sum_n (-1)^(n+1) y^(n^2)
  const int ncdf = 100;
  for (int i = 1; i <= ncdf; i++) {
    g.mrw_xcdf[i - 1] = (double)(i - 1) / (double)(ncdf - 1);
    double y = 0.0;
    if (i == ncdf) {
      y = 0.5;
    } else {
      for (int64_t j = 1;; j++) {
        double term = std::pow(g.mrw_xcdf[i - 1], (double)(j * j));
        if (term == 0.0) break;
        if (j % 2 == 0)
          y = y - term;
        else
          y = y + term;
      }
    }
    g.mrw_ycdf[i - 1] = y * 2.0;
  }
  const int nc = g.n_cells;
  g.alpha_inv_planck.assign(nc, 0.0);
  g.diff_coeff.assign(nc, 0.0);
  for (int ic = 0; ic < nc; ic++) {
    for (int id = 0; id < g.n_dust; id++) {
      size_t k = (size_t)id * nc + ic;
      if (g.density[k] > 0.0) {
        const Dust &d = g.d[id];
        g.alpha_inv_planck[ic] = g.alpha_inv_planck[ic] +
                                 g.density[k] * interp1d_loglog(d.specific_energy.data(), d.chi_inv_planck.data(), d.n_e, g.specific_energy[k]);
      }
    }
    double total = 0.0;
    for (int id = 0; id < g.n_dust; id++) {
      size_t k = (size_t)id * nc + ic;
      const Dust &d = g.d[id];
      // cells without dust would divide by zero below; they are never entered by the random walk
      if (g.density[k] > 0.0)
        total = total + g.density[k] * interp1d_loglog(d.specific_energy.data(), d.chi_inv_planck.data(), d.n_e, g.specific_energy[k]);
    }
    g.diff_coeff[ic] = 1.0 / 3.0 / total;
  }
}

// sample_cumulative (:198-203): interp1d(ycdf, xcdf, xi), linear
double mrw_sample_cumulative(orc_ctx &g) {
  double xi = g.rng.random();
  int ip = locate(g.mrw_ycdf, 100, xi);
  if (ip == -1) throw OracleError{"Interpolation out of bounds"};
  const double *x = g.mrw_ycdf, *y = g.mrw_xcdf;
  return (xi - x[ip - 1]) / (x[ip] - x[ip - 1]) * (y[ip] - y[ip - 1]) + y[ip - 1];
}

// grid_do_mrw (:56-111) / grid_do_mrw_noenergy (:113-149)
void grid_do_mrw(orc_ctx &g, Photon &p, bool deposit) {
  const int nc = g.n_cells, ic = p.icell.ic;
  double R0 = distance_to_closest_wall(g, p);
  if (deposit) {
    double y = mrw_sample_cumulative(g);
    double ct = -std::log(y) / g.diff_coeff[ic - 1] * ((R0 / PI_F) * (R0 / PI_F));
    for (int id = 0; id < g.n_dust; id++) {
      size_t k = (size_t)id * nc + ic - 1;
      if (g.density[k] > 0.0) {
        const Dust &d = g.d[id];
        double e = p.energy * ct * interp1d_loglog(d.specific_energy.data(), d.kappa_planck.data(), d.n_e, g.specific_energy[k]);
        g.specific_energy_sum[k] = g.specific_energy_sum[k] + e;
        // deposit_specific_energy_spectrum (grid_physics_3d.f90:367-395): spread over the bins like the local
        // emissivity, interpolated between the two adjacent emissivity states
        if (g.n_nu_bins > 0) {
          const int iv = g.jnu_var_id[k];
          const double fr = g.jnu_var_frac[k];
          const double *f1 = &g.j_nu_bin_frac[((size_t)id * g.n_jnu_max + iv - 1) * g.n_nu_bins];
          const double *f2 = f1 + g.n_nu_bins;
          for (int ib = 0; ib < g.n_nu_bins; ib++) {
            double &b = g.specific_energy_sum_spectrum[(size_t)ib * g.n_dust * nc + k];
            b = b + e * ((1.0 - fr) * f1[ib] + fr * f2[ib]);
          }
        }
      }
    }
  }
  // random_sphere_vector3d (type_vector3d.f90:323-333)
  double mu = -1.0 + 2.0 * g.rng.random();
  double phi = TWOPI_F * g.rng.random();
  double cut = std::sqrt(1.0 - mu * mu);
  p.r.x = p.r.x + cut * std::cos(phi) * R0;
  p.r.y = p.r.y + cut * std::sin(phi) * R0;
  p.r.z = p.r.z + mu * R0;
  p.a_prev = p.a;
  p.v_prev = p.v;
  p.s_prev = p.s;
  p.a = random_sphere_angle3d(g.rng);
  p.v = angle3d_to_vector3d(p.a);
  int id = select_dust_chi_rho(g, p);
  size_t k = (size_t)(id - 1) * nc + ic - 1;
  p.nu = g.d[id - 1].sample_b_nu(g.rng, g.jnu_var_id[k], g.jnu_var_frac[k]);
  // NB: the reference does not refresh the opacities for the new frequency here
  p.last_isotropic = true;
  p.dust_id = id;
  p.last[0] = deposit ? 'd' : 'm';
  p.last[1] = 'e';
  g.n_mrw_steps++;
}

// the MRW block at the top of the interaction loop (iter_lucy.f90:134-148, iter_final.f90:166-185);
// returns false if the packet had to be killed
bool mrw_steps(orc_ctx &g, Photon &p, bool deposit, bool peel);

// the photon loop of do_lucy (iter_lucy.f90:127-205)
void lucy_photons(orc_ctx &g, int64_t n_photons) {
  Photon p;
  const int64_t n_inter_max = g.conf.n_inter_max;
  for (int64_t ip = 1; ip <= n_photons; ip++) {
    emit(g, p);
    g.n_photons_run++;
    for (int64_t interactions = 1; interactions <= n_inter_max + 1; interactions++) {
      if (g.conf.use_mrw && interactions > 1) {
        if (!mrw_steps(g, p, true, false)) break;
      }
      double tau = g.rng.random_exp();
      double tau_achieved;
      grid_integrate(g, p, tau, tau_achieved);
      if (p.reabsorbed) {
        // loop until the packet finally escapes interacting with sources (iter_lucy.f90:158-185)
        int64_t ia;
        for (ia = 1; ia <= g.conf.n_reabs_max; ia++) {
          emit(g, p, true, p.reabsorbed_id, p.energy);
          g.n_reabsorptions++;
          tau = g.rng.random_exp();
          grid_integrate(g, p, tau, tau_achieved);
          if (!p.reabsorbed) break;
        }
        if (ia == g.conf.n_reabs_max + 1) {
          g.killed_photons_int++;
          p.killed = true;
          break;
        }
      }
      if (p.killed || escaped(g, p.icell)) {
        if (!p.killed) g.n_escaped++;
        break;
      }
      if (interactions == n_inter_max + 1) {
        g.killed_photons_int++;
        p.killed = true;
        break;
      }
      interact(g, p);
      p.killed = (g.conf.kill_on_scatter && p.scattered) || (g.conf.kill_on_absorb && !p.scattered);
      if (p.killed) break;
    }
  }
}


// grid_integrate_noenergy (grid_propagate_3d.f90:237-375)
void grid_integrate_noenergy(orc_ctx &g, Photon &p, double tau_required, double &tau_achieved) {
  const double frac_check = g.conf.propagation_check_frequency;
  tau_achieved = 0.0;
  if (!p.in_cell) throw OracleError{"photon has not been placed in a cell"};
  if (escaped(g, p.icell)) return;
  if (tau_required == 0.0) return;
  double t_source;
  int source_id;
  find_nearest_source(g, p.r, p.v, t_source, source_id);
  g.radial = (p.r.x * p.v.x + p.r.y * p.v.y + p.r.z * p.v.z) > 0.;  // grid_propagate_3d.f90:73,262
  double t_achieved = 0.0;
  const int nc = g.n_cells;
  for (;;) {
    double xi = g.rng.random();
    if (xi < frac_check) {
      if (!in_correct_cell(g, p)) {
        g.killed_photons_geo++;
        p.killed = true;
        return;
      }
    }
    double tau_needed = tau_required - tau_achieved;
    double tmin;
    WallId id_min;
    find_wall(g, p, tmin, id_min);
    if (id_min.w1 == 0 && id_min.w2 == 0 && id_min.w3 == 0) {
      g.killed_photons_geo++;
      p.killed = true;
      return;
    }
    int ic = p.icell.ic;
    double chi_rho_total = 0.0;
    for (int id = 0; id < g.n_dust; id++)
      chi_rho_total = chi_rho_total + p.current_chi[id] * g.density[(size_t)id * nc + ic - 1];
    double tau_cell = chi_rho_total * tmin;
    g.n_crossings++;
    if (tau_cell < tau_needed) {
      t_achieved = t_achieved + tmin;
      if (t_achieved > t_source) {
        p.reabsorbed = true;
        p.reabsorbed_id = source_id;
        return;
      }
      p.r.x = p.r.x + tmin * p.v.x;
      p.r.y = p.r.y + tmin * p.v.y;
      p.r.z = p.r.z + tmin * p.v.z;
      tau_achieved = tau_achieved + tau_cell;
      p.on_wall = true;
      p.icell = next_cell(g, p.icell, id_min, p.r);
      p.on_wall_id = WallId{-id_min.w1, -id_min.w2, -id_min.w3};
      if (escaped(g, p.icell)) return;
    } else {
      double tact = tmin * (tau_needed / tau_cell);
      t_achieved = t_achieved + tact;
      if (t_achieved > t_source) {
        p.reabsorbed = true;
        p.reabsorbed_id = source_id;
        return;
      }
      p.r.x = p.r.x + tact * p.v.x;
      p.r.y = p.r.y + tact * p.v.y;
      p.r.z = p.r.z + tact * p.v.z;
      tau_achieved = tau_achieved + tau_needed;
      p.on_wall = false;
      p.on_wall_id = WallId();
      return;
    }
  }
}

// grid_escape_tau (grid_propagate_3d.f90:377-480) and grid_escape_column_density (:482-582):
// march a COPY of the photon to tmax / the grid edge.  column == nullptr: optical depth.
void grid_escape(orc_ctx &g, const Photon &p_orig, double tmax, double &tau, double *column, bool &killed) {
  const double frac_check = g.conf.propagation_check_frequency;
  Photon p = p_orig;
  killed = false;
  if (!p.in_cell) throw OracleError{"photon has not been placed in a cell"};
  if (escaped(g, p.icell)) {
    tau = 0.0;
    if (column)
      for (int id = 0; id < g.n_dust; id++) column[id] = 0.0;
    return;
  }
  double t_source;
  int source_id;
  find_nearest_source(g, p.r, p.v, t_source, source_id);
  g.radial = (p.r.x * p.v.x + p.r.y * p.v.y + p.r.z * p.v.z) > 0.;  // grid_propagate_3d.f90:400,505
  if (t_source < tmax) {
    killed = true;
    return;
  }
  tau = 0.0;
  if (column)
    for (int id = 0; id < g.n_dust; id++) column[id] = 0.0;
  double t_current = 0.0;
  bool finished = false;
  const int nc = g.n_cells;
  for (;;) {
    double xi = g.rng.random();
    if (xi < frac_check) {
      if (!in_correct_cell(g, p)) {
        g.killed_photons_geo++;
        killed = true;
        return;
      }
    }
    double tmin;
    WallId id_min;
    find_wall(g, p, tmin, id_min);
    if (id_min.w1 == 0 && id_min.w2 == 0 && id_min.w3 == 0) {
      g.killed_photons_geo++;
      killed = true;
      return;
    }
    if (t_current + tmin > tmax) {
      tmin = tmax - t_current;
      finished = true;
    }
    p.r.x = p.r.x + tmin * p.v.x;
    p.r.y = p.r.y + tmin * p.v.y;
    p.r.z = p.r.z + tmin * p.v.z;
    t_current = t_current + tmin;
    g.n_peel_crossings++;
    int ic = p.icell.ic;
    if (column) {
      for (int id = 0; id < g.n_dust; id++) column[id] = column[id] + g.density[(size_t)id * nc + ic - 1] * tmin;
    } else {
      for (int id = 0; id < g.n_dust; id++)
        tau = tau + p.current_chi[id] * g.density[(size_t)id * nc + ic - 1] * tmin;
    }
    if (finished) return;
    p.on_wall = true;
    p.icell = next_cell(g, p.icell, id_min, p.r);
    p.on_wall_id = WallId{-id_min.w1, -id_min.w2, -id_min.w3};
    if (escaped(g, p.icell)) return;
  }
}

// forced_interaction_wr99 / _baes16 (src/main/forced_interaction.f90:23-133)
void forced_interaction(orc_ctx &g, double tau_escape, double &tau, double &weight) {
  const double TAU_THRES = 1.e-7;
  if (g.conf.forced_first_interaction_algorithm == HYP_FFI_BAES16) {
    double one_minus_exp = tau_escape > TAU_THRES ? 1.0 - std::exp(-tau_escape) : tau_escape;
    double alpha = (1.0 - g.conf.baes16_xi) / one_minus_exp;
    double beta = g.conf.baes16_xi / tau_escape;
    double tau_min = 0.0, tau_max = tau_escape;
    double xi = g.rng.random();
    for (int i = 1; i <= 60; i++) {
      tau = 0.5 * (tau_min + tau_max);
      double xi_test;
      if (tau > TAU_THRES)
        xi_test = alpha * (1.0 - std::exp(-tau)) + beta * tau;
      else
        xi_test = alpha * tau + beta * tau;
      if (xi_test > xi)
        tau_max = tau;
      else
        tau_min = tau;
    }
    tau = 0.5 * (tau_min + tau_max);
    weight = 1.0 / (alpha + beta * std::exp(tau));
  } else {
    double xi = g.rng.random();
    double one_minus_exp = tau_escape > TAU_THRES ? (1.0 - std::exp(-tau_escape)) : tau_escape;
    tau = -std::log(1.0 - xi * one_minus_exp);
    weight = one_minus_exp;
  }
}

// dust_scatter_peeloff (dust_type_4elem.f90:421-444)
void dust_scatter_peeloff(const Dust &d, double nu, Angle &a, Stokes &s, const Angle &a_req) {
  Angle a_scat = difference_angle3d(a, a_req);
  if (a_scat.cost < d.mu_min || a_scat.cost > d.mu_max) {
    s = Stokes{0.0, 0.0, 0.0, 0.0};
  } else {
    double P1 = interp2d(d.mu.data(), d.n_mu, d.nu.data(), d.n_nu, d.P1.data(), a_scat.cost, nu);
    double P2 = interp2d(d.mu.data(), d.n_mu, d.nu.data(), d.n_nu, d.P2.data(), a_scat.cost, nu);
    double P3 = interp2d(d.mu.data(), d.n_mu, d.nu.data(), d.n_nu, d.P3.data(), a_scat.cost, nu);
    double P4 = interp2d(d.mu.data(), d.n_mu, d.nu.data(), d.n_nu, d.P4.data(), a_scat.cost, nu);
    scatter_stokes(s, a, a_scat, a_req, P1, P2, P3, P4);
  }
  a = a_req;
}

// normalized_B_nu (source_type.f90:1088-1096)
double normalized_B_nu(double nu, double T) {
  const double h_cgs = 6.6260689633e-27, c_cgs = 2.99792458e10, k_cgs = 1.380650424e-16, stef_boltz = 5.670400e-5;  // lib_constants.f90:52-96
  const double a = 2.0 * h_cgs / c_cgs / c_cgs / stef_boltz * PI;
  const double b = h_cgs / k_cgs;
  double T4 = T * T * T * T;
  return a * nu * nu * nu / (std::exp(b * nu / T) - 1.0) / T4;
}

void bin_edges(const Image &im, int inu, double &numin, double &numax) {
  numin = std::pow(10.0, im.log10_nu_min + (im.log10_nu_max - im.log10_nu_min) * (double)(inu - 1) / (double)im.n_nu);
  numax = std::pow(10.0, im.log10_nu_min + (im.log10_nu_max - im.log10_nu_min) * (double)inu / (double)im.n_nu);
}

// get_spectrum_binned (source_type.f90:1118-1165)
std::vector<double> get_spectrum_binned(const Source &src, const Image &im) {
  std::vector<double> nu, fnu;
  if (src.freq_type == 1) {
    nu = src.spectrum.x;
    fnu = src.spectrum.pdf;
  } else {
    double lmin = std::log10(3.e9), lmax = std::log10(3.e16);
    int n = (int)std::ceil((lmax - lmin) * 100000);
    nu.resize(n);
    fnu.resize(n);
    for (int i = 1; i <= n; i++) {
      nu[i - 1] = std::pow(10.0, (double)(i - 1) / (double)(n - 1) * (lmax - lmin) + lmin);
      fnu[i - 1] = normalized_B_nu(nu[i - 1], src.temperature);
    }
  }
  std::vector<double> sp(im.n_nu);
  for (int inu = 1; inu <= im.n_nu; inu++) {
    double numin, numax;
    bin_edges(im, inu, numin, numax);
    sp[inu - 1] = integral_loglog_subset(nu.data(), fnu.data(), (int)nu.size(), numin, numax);
  }
  double tot = integral_general(nu.data(), fnu.data(), (int)nu.size(), trapezium_loglog);
  for (auto &v : sp) v = v / tot;
  return sp;
}

// get_j_nu_binned (dust_type_4elem.f90:722-741)
std::vector<double> get_j_nu_binned(const Dust &d, const Image &im, int ijnu) {
  const PdfCont &j = d.j_nu[ijnu - 1];
  std::vector<double> sp(im.n_nu);
  for (int inu = 1; inu <= im.n_nu; inu++) {
    double numin, numax;
    bin_edges(im, inu, numin, numax);
    sp[inu - 1] = integral_loglog_subset(j.x.data(), j.pdf.data(), j.n, numin, numax);
  }
  double tot = integral_general(j.x.data(), j.pdf.data(), j.n, trapezium_loglog);
  for (auto &v : sp) v = v / tot;
  return sp;
}

// get_chi_nu_binned (dust_type_4elem.f90:793-811)
std::vector<double> get_chi_nu_binned(const Dust &d, const Image &im) {
  std::vector<double> chi(im.n_nu);
  for (int inu = 1; inu <= im.n_nu; inu++) {
    double numin, numax;
    bin_edges(im, inu, numin, numax);
    chi[inu - 1] = integral_loglog_subset(d.nu.data(), d.chi_nu.data(), d.n_nu, numin, numax) / (numax - numin);
  }
  return chi;
}

// peeloff_photon (images_peeled.f90:95-270); inside observers are not restated
// get_spectrum_interp (source_type.f90:1098-1116), get_j_nu_interp (dust_type_4elem.f90:708-720),
// get_chi_nu_interp (:780-791): the raytracing spectra at the exact frequencies of a monochromatic image
std::vector<double> get_spectrum_interp(const Source &src, const Image &im) {
  std::vector<double> sp(im.n_nu);
  for (int i = 0; i < im.n_nu; i++) {
    if (src.freq_type == 1)
      sp[i] = interp1d_loglog(src.spectrum.x.data(), src.spectrum.pdf.data(), src.spectrum.n, im.nu[i], false, 0.0);
    else if (src.freq_type == 2)
      sp[i] = normalized_B_nu(im.nu[i], src.temperature);
    else
      throw OracleError{"cannot get spectrum"};
  }
  return sp;
}
std::vector<double> get_j_nu_interp(const Dust &d, const Image &im, int ijnu) {
  const PdfCont &j = d.j_nu[ijnu - 1];
  std::vector<double> sp(im.n_nu);
  for (int i = 0; i < im.n_nu; i++) sp[i] = interp1d_loglog(j.x.data(), j.pdf.data(), j.n, im.nu[i], false, 0.0);
  return sp;
}
std::vector<double> get_chi_nu_interp(const Dust &d, const Image &im) {
  std::vector<double> sp(im.n_nu);
  for (int i = 0; i < im.n_nu; i++) sp[i] = interp1d_loglog(d.nu.data(), d.chi_nu.data(), d.n_nu, im.nu[i], false, 0.0);
  return sp;
}

void peeloff_photon(orc_ctx &g, const Photon &p_orig, bool polychromatic) {
  PeeledState &P = g.peeled;
  const int n_peeled = (int)P.group_id.size();
  std::vector<double> column(g.n_dust), spectrum;
  for (int ip = 1; ip <= n_peeled; ip++) {
    const int ig = P.group_id[ip - 1], iv = P.view_id[ip - 1];
    Image &im = P.image[ig - 1];
    Photon p = p_orig;
    p.s = p.s_prev;
    p.a = p.a_prev;
    p.v = p.v_prev;
    // a_peeloff (images_peeled.f90:410-421): an inside observer is looked at from the event itself
    const bool inside = im.c.inside_observer != 0;
    Angle a_req = P.viewing_angles[ip - 1];
    if (inside) {
      // vector3d_to_angle3d (type_vector3d.f90:276-299) of r_peeloff - r
      const Vec rp0 = P.r_peeloff[ig - 1];
      const Vec w{rp0.x - p.r.x, rp0.y - p.r.y, rp0.z - p.r.z};
      const double small_r = std::sqrt(w.x * w.x + w.y * w.y);
      const double big_r = std::sqrt(w.x * w.x + w.y * w.y + w.z * w.z);
      a_req.cosp = w.x / small_r;
      a_req.sinp = w.y / small_r;
      a_req.cost = w.z / big_r;
      a_req.sint = small_r / big_r;
    }
    const Vec v_req = angle3d_to_vector3d(a_req);
    if (p.last_isotropic) {
      p.s = Stokes{1.0, 0.0, 0.0, 0.0};
      p.a = a_req;
      p.v = v_req;
    } else {
      if (p.last[0] == 's' && p.last[1] == 'r') {
        // source_emit_peeloff (source_type.f90:513-537) -> emit_from_sphere_peeloff (:692-707)
        const Source &src = g.s[p.source_id - 1];
        if (src.peeloff) {
          if (src.type != HYP_SOURCE_SPHERE && src.type != HYP_SOURCE_EXTERN_SPH && src.type != HYP_SOURCE_EXTERN_BOX)
            throw OracleError{"Should not be here, all other source types are isotropic"};
          // emit_from_sphere_peeloff (:692-707), emit_from_extern_sph_peeloff (:804-820),
          // emit_from_extern_box_peeloff (:906-933)
          const Angle b = src.type == HYP_SOURCE_EXTERN_BOX ? extern_box_normal(p.face_id) : p.source_a;
          double mu = a_req.sint * a_req.cosp * b.sint * b.cosp + a_req.sint * a_req.sinp * b.sint * b.sinp + a_req.cost * b.cost;
          mu = std::max(mu, 0.0);
          if (src.limb_darkening)
            p.s = Stokes{2. * (1.5 * mu * mu + mu), 0.0, 0.0, 0.0};
          else
            p.s = Stokes{4. * mu, 0.0, 0.0, 0.0};
          p.a = a_req;
        } else {
          p.s = Stokes{0.0, 0.0, 0.0, 0.0};
          p.a = a_req;
        }
        p.v = angle3d_to_vector3d(p.a);
      } else if (p.last[0] == 'd' && p.last[1] == 's') {
        dust_scatter_peeloff(g.d[p.dust_id - 1], p.nu, p.a, p.s, a_req);
        p.v = angle3d_to_vector3d(p.a);
      } else if (p.last[0] == 'd' && p.last[1] == 'e') {
        p.a = a_req;  // dust_emit_peeloff (dust_type_4elem.f90:322-332)
        p.v = angle3d_to_vector3d(p.a);
      } else {
        throw OracleError{"unexpected p%last flag"};
      }
    }
    place_in_cell(g, p);
    const Vec rp = P.r_peeloff[ig - 1];
    double d, tmax;
    Vec dr{p.r.x - rp.x, p.r.y - rp.y, p.r.z - rp.z};
    if (inside) {
      d = std::sqrt(dr.x * dr.x + dr.y * dr.y + dr.z * dr.z);
      tmax = d;
    } else {
      d = -(v_req.x * p.r.x + v_req.y * p.r.y + v_req.z * p.r.z);
      tmax = std::numeric_limits<double>::max();
    }
    if (d < im.c.d_min || d > im.c.d_max) continue;
    double x_image, y_image;
    if (inside) {
      // longitude / latitude of the arrival direction in the observer's frame (images_peeled.f90:168-183)
      const double rad2deg = 180.0 / PI;
      const Angle a_view = P.viewing_angles[ip - 1];
      const Vec v_a = angle3d_to_vector3d(p.a);
      Vec v_sky;
      v_sky.x = (v_a.x * a_view.cosp + v_a.y * a_view.sinp) * a_view.sint + v_a.z * a_view.cost;
      v_sky.y = -v_a.x * a_view.sinp + v_a.y * a_view.cosp;
      v_sky.z = -(v_a.x * a_view.cosp + v_a.y * a_view.sinp) * a_view.cost + v_a.z * a_view.sint;
      x_image = std::atan2(v_sky.y, v_sky.x) * rad2deg;
      y_image = std::atan2(std::sqrt(v_sky.x * v_sky.x + v_sky.y * v_sky.y), v_sky.z) * rad2deg - 90.0;
      x_image = im.c.x_max + f_modulo(x_image - im.c.x_max, 360.0);
      y_image = im.c.y_min + f_modulo(y_image - im.c.y_min, 360.0);
    } else {
      x_image = dr.y * p.a.cosp - dr.x * p.a.sinp;
      y_image = dr.z * p.a.sint - dr.y * p.a.cost * p.a.sinp - dr.x * p.a.cost * p.a.cosp;
    }
    if (!in_image(im, x_image, y_image)) continue;
    double tau = 0.0;
    bool killed = false;
    if (im.c.ignore_optical_depth) {
      for (auto &c : column) c = 0.0;
    } else {
      grid_escape(g, p, tmax, tau, polychromatic ? column.data() : nullptr, killed);
    }
    if (killed) continue;
    g.n_peeloffs++;
    if (inside) {
      // flux through a unit area at the observer (images_peeled.f90:207)
      const double f = 4.0 * PI * std::pow(d, 2.0);
      p.s = Stokes{p.s.I / f, p.s.Q / f, p.s.U / f, p.s.V / f};
    }
    if (polychromatic) {
      if (p.emiss_type == 1 || p.emiss_type == 2) {
        auto &cache = P.source_spectra[ig - 1][p.source_id - 1];
        if (cache.empty())
          cache = im.use_exact_nu ? get_spectrum_interp(g.s[p.source_id - 1], im) : get_spectrum_binned(g.s[p.source_id - 1], im);
        spectrum = cache;
      } else if (p.emiss_type == 3) {
        // get_dust_emissivity (images_peeled.f90:454-500)
        auto &cache = P.dust_log10_emissivity[ig - 1][p.dust_id - 1];
        const Dust &dd = g.d[p.dust_id - 1];
        if (cache.empty()) {
          cache.resize((size_t)dd.n_jnu * im.n_nu);
          for (int ij = 1; ij <= dd.n_jnu; ij++) {
            std::vector<double> sp = im.use_exact_nu ? get_j_nu_interp(dd, im, ij) : get_j_nu_binned(dd, im, ij);
            for (int inu = 0; inu < im.n_nu; inu++) cache[(size_t)(ij - 1) * im.n_nu + inu] = std::log10(sp[inu]);
          }
        }
        spectrum.resize(im.n_nu);
        for (int inu = 0; inu < im.n_nu; inu++) {
          double l0 = cache[(size_t)(p.emiss_var_id - 1) * im.n_nu + inu];
          double l1 = cache[(size_t)p.emiss_var_id * im.n_nu + inu];
          double v = std::pow(10.0, (l1 - l0) * p.emiss_var_frac + l0);
          spectrum[inu] = std::isnan(v) ? 0.0 : v;
        }
      } else {
        throw OracleError{"unknown emiss_type"};
      }
      for (auto &v : spectrum) v = v * p.s.I * p.energy;
      for (int id = 1; id <= g.n_dust; id++) {
        auto &chi = P.dust_extinction[ig - 1][id - 1];
        if (chi.empty()) chi = im.use_exact_nu ? get_chi_nu_interp(g.d[id - 1], im) : get_chi_nu_binned(g.d[id - 1], im);
        for (int inu = 0; inu < im.n_nu; inu++) spectrum[inu] = spectrum[inu] * std::exp(-column[id - 1] * chi[inu]);
      }
      image_bin_raytraced(im, p, x_image, y_image, iv, spectrum);
    } else {
      double e = std::exp(-tau);
      p.s = Stokes{p.s.I * e, p.s.Q * e, p.s.U * e, p.s.V * e};
      image_bin(im, p, x_image, y_image, iv);
    }
  }
}

bool mrw_steps(orc_ctx &g, Photon &p, bool deposit, bool peel) {
  int64_t steps;
  for (steps = 1; steps <= g.conf.n_mrw_max; steps++) {
    if (g.alpha_inv_planck[p.icell.ic - 1] * distance_to_closest_wall(g, p) > g.conf.mrw_gamma) {
      grid_do_mrw(g, p, deposit);
      if (peel) peeloff_photon(g, p, false);
    } else {
      break;
    }
  }
  if (steps == g.conf.n_mrw_max + 1) {
    g.killed_photons_int++;
    p.killed = true;
    return false;
  }
  return true;
}

// propagate (iter_final.f90:147-273); MRW and source re-absorption are not restated
void propagate_final(orc_ctx &g, Photon &p, bool peeloff_scattering_only) {
  const int64_t n_inter_max = g.conf.n_inter_max;
  const bool make_peeled = !g.peeled.group_id.empty();
  for (int64_t interactions = 1; interactions <= n_inter_max + 1; interactions++) {
    if (interactions > 1 && g.conf.use_mrw) {
      if (!mrw_steps(g, p, false, make_peeled && !peeloff_scattering_only)) break;
    }
    double tau;
    if (interactions == 1 && g.conf.forced_first_interaction) {
      double tau_escape;
      bool killed;
      grid_escape(g, p, std::numeric_limits<double>::max(), tau_escape, nullptr, killed);
      if (tau_escape > 1.e-10 && !killed) {
        double weight;
        forced_interaction(g, tau_escape, tau, weight);
        p.energy = p.energy * weight;
      } else {
        tau = g.rng.random_exp();
      }
    } else {
      tau = g.rng.random_exp();
    }
    double tau_achieved;
    grid_integrate_noenergy(g, p, tau, tau_achieved);
    if (p.reabsorbed) {
      // iter_final.f90:212-242: re-emission from the stellar surface is always peeled off
      int64_t ia;
      for (ia = 1; ia <= g.conf.n_reabs_max; ia++) {
        emit(g, p, true, p.reabsorbed_id, p.energy);
        g.n_reabsorptions++;
        if (make_peeled) peeloff_photon(g, p, false);
        tau = g.rng.random_exp();
        grid_integrate_noenergy(g, p, tau, tau_achieved);
        if (!p.reabsorbed) break;
      }
      if (ia == g.conf.n_reabs_max + 1) {
        g.killed_photons_int++;
        p.killed = true;
        break;
      }
    }
    if (p.killed || escaped(g, p.icell)) {
      if (!p.killed) g.n_escaped++;
      break;
    }
    if (interactions == n_inter_max + 1) {
      g.killed_photons_int++;
      p.killed = true;
      break;
    }
    interact(g, p);
    p.killed = (g.conf.kill_on_scatter && p.scattered) || (g.conf.kill_on_absorb && !p.scattered);
    if (p.killed) break;
    if (make_peeled) {
      if (p.scattered || !peeloff_scattering_only) peeloff_photon(g, p, false);
    }
  }
}

// binned_images_bin_photon (images_binned.f90:57-77): an escaping packet goes into the view its own
// direction falls in (n_theta bins of cos theta, n_phi bins of phi)
void binned_images_bin_photon(Image &im, const Photon &p) {
  const int n_theta = im.c.n_theta, n_phi = im.c.n_phi;
  double phi = std::atan2(p.a.sinp, p.a.cosp);
  if (phi < 0.) phi = phi + TWOPI_F;
  const int it = ipos(-1.0, +1.0, p.a.cost, n_theta);
  const int ip = ipos(0.0, TWOPI_F, phi, n_phi);
  const double x_image = p.r.y * p.a.cosp - p.r.x * p.a.sinp;
  const double y_image = p.r.z * p.a.sint - p.r.y * p.a.cost * p.a.sinp - p.r.x * p.a.cost * p.a.cosp;
  // image_bin indexes img(:, :, :, iv, :, :); the Fortran would fault on an out-of-range view, which
  // ipos only returns for NaN directions
  if (it < 1 || it > n_theta || ip < 1 || ip > n_phi) return;
  image_bin(im, p, x_image, y_image, n_phi * (it - 1) + ip);
}

// do_final (iter_final.f90:60-145), photon loop only
void final_photons(orc_ctx &g, int64_t n_photons, bool peeloff_scattering_only) {
  const bool make_peeled = !g.peeled.group_id.empty();
  Photon p;
  for (int64_t ip = 1; ip <= n_photons; ip++) {
    emit(g, p);
    g.n_photons_run++;
    if (make_peeled && !peeloff_scattering_only) peeloff_photon(g, p, false);
    propagate_final(g, p, peeloff_scattering_only);
    // iter_final.f90:126-129
    if (!p.killed)
      for (auto &im : g.peeled.image)
        if (im.c.binned) binned_images_bin_photon(im, p);
  }
}

// emit_from_grid (grid_physics_3d.f90:691-753), polychromatic form
Photon emit_from_grid(orc_ctx &g) {
  Photon p;
  double xi = g.rng.random();
  p.dust_id = std::max((int)std::ceil(xi * (double)g.n_dust), 1);
  // random_masked_cell (grid_geometry_common_3d.f90:104-115)
  xi = g.rng.random();
  int ic;
  const int n_masked = g.mask_map.empty() ? g.n_cells : g.n_masked;
  if (!g.mask_map.empty())
    ic = g.mask_map[std::max((int)std::ceil(xi * g.n_masked), 1) - 1];
  else
    ic = std::max((int)std::ceil(xi * g.n_cells), 1);
  p.in_cell = true;
  place_at_random_position_in_cell(g, p, ic);
  p.a = random_sphere_angle3d(g.rng);
  p.v = angle3d_to_vector3d(p.a);
  p.s = Stokes{1.0, 0.0, 0.0, 0.0};
  size_t k = (size_t)(p.dust_id - 1) * g.n_cells + ic - 1;
  if (g.energy_abs_tot[p.dust_id - 1] > 0.0) {
    double mass = g.density[k] * g.volume[ic - 1];
    p.energy = g.specific_energy[k] * mass * (double)n_masked / g.energy_abs_tot[p.dust_id - 1];
  } else {
    p.energy = 0.0;
  }
  p.emiss_type = 3;
  p.emiss_var_id = g.jnu_var_id[k];
  p.emiss_var_frac = g.jnu_var_frac[k];
  p.scattered = false;
  p.reprocessed = true;
  p.last_isotropic = true;
  p.last[0] = 'd';
  p.last[1] = 'e';
  return p;
}

// do_raytracing (iter_raytracing.f90:31-141)
void raytracing_photons(orc_ctx &g, int64_t n_sources, int64_t n_thermal) {
  for (const auto &im : g.peeled.image)
    if (im.c.use_filters && !im.c.binned) throw OracleError{"filter convolution cannot be used with raytracing"};
  precompute_jnu_var(g);
  Photon p;
  for (int64_t ip = 1; ip <= n_sources; ip++) {
    emit(g, p);
    p.energy = p.energy * g.energy_total / (double)n_sources;
    peeloff_photon(g, p, true);
  }
  if (g.n_dust == 0) return;
  for (int64_t ip = 1; ip <= n_thermal; ip++) {
    p = emit_from_grid(g);
    if (p.energy > 0.0) {
      p.energy = p.energy * g.energy_abs_tot[p.dust_id - 1] / (double)n_thermal * (double)g.n_dust;
      peeloff_photon(g, p, true);
    }
  }
}

// ---------------------------------------------------------------------------
// Monochromatic final iteration (src/main/iter_final_mono.f90, src/grid/grid_monochromatic.f90)
// ---------------------------------------------------------------------------

// setup_monochromatic_grid_pdfs (grid_monochromatic.f90:50-118); returns `empty`
bool setup_monochromatic_grid_pdfs(orc_ctx &g, int inu) {
  const double nu = g.frequencies[inu - 1];
  const int nc = g.n_cells;
  g.mono_mean_prob.assign(g.n_dust, 0.0);
  g.mono_emiss_pdf.assign(g.n_dust, PdfDiscrete());
  std::vector<double> energy(nc), prob(nc), pe(nc);
  for (int id = 1; id <= g.n_dust; id++) {
    const size_t o = (size_t)(id - 1) * nc;
    for (int ic = 0; ic < nc; ic++) {
      if (g.energy_abs_tot[id - 1] > 0.0)
        energy[ic] = g.specific_energy[o + ic] * g.density[o + ic] * g.volume[ic] * (double)nc / g.energy_abs_tot[id - 1];
      else
        energy[ic] = 0.0;
    }
    if (!g.mask_map.empty()) {
      std::vector<char> in(nc, 0);
      for (int ic : g.mask_map) in[ic - 1] = 1;
      for (int ic = 0; ic < nc; ic++)
        if (!in[ic]) energy[ic] = 0.0;
    }
    for (int ic = 0; ic < nc; ic++)
      prob[ic] = dust_sample_emit_probability(g.d[id - 1], g.jnu_var_id[o + ic], g.jnu_var_frac[o + ic], nu);
    double sum = 0.0;
    for (int ic = 0; ic < nc; ic++) {
      pe[ic] = prob[ic] * energy[ic];
      sum = sum + pe[ic];
    }
    g.mono_mean_prob[id - 1] = sum / (double)nc;
    if (g.mono_mean_prob[id - 1] > 0.0) g.mono_emiss_pdf[id - 1].set(pe.data(), nc);
  }
  double tot = 0.0;
  for (double m : g.mono_mean_prob) tot = tot + m;
  return tot == 0.0;
}

// emit_from_monochromatic_grid_pdf (grid_monochromatic.f90:120-174)
Photon emit_from_monochromatic_grid_pdf(orc_ctx &g, int inu) {
  Photon p;
  p.nu = g.frequencies[inu - 1];
  p.inu = inu;
  update_optconsts(g, p);
  const double xi = g.rng.random();
  const int dust_id = std::max((int)std::ceil(xi * (double)g.n_dust), 1);
  if (g.mono_mean_prob[dust_id - 1] == 0.0) {
    p.energy = 0.0;
    return p;
  }
  const int ic = g.mono_emiss_pdf[dust_id - 1].sample(g.rng);
  p.in_cell = true;
  place_at_random_position_in_cell(g, p, ic);
  p.a = random_sphere_angle3d(g.rng);
  p.v = angle3d_to_vector3d(p.a);
  p.s = Stokes{1.0, 0.0, 0.0, 0.0};
  p.energy = g.mono_mean_prob[dust_id - 1];
  p.scattered = false;
  p.reprocessed = true;
  p.last_isotropic = true;
  p.dust_id = dust_id;
  p.last[0] = 'd';
  p.last[1] = 'e';
  return p;
}

// propagate of iter_final_mono.f90:231-341: every interaction is a forced scattering
void propagate_mono(orc_ctx &g, Photon &p) {
  const int64_t n_inter_max = g.conf.n_inter_max;
  const bool make_peeled = !g.peeled.group_id.empty();
  const double energy_initial = p.energy;
  for (int64_t interactions = 1; interactions <= n_inter_max + 1; interactions++) {
    double tau;
    if (interactions == 1 && g.conf.forced_first_interaction) {
      double tau_escape;
      bool killed;
      grid_escape(g, p, std::numeric_limits<double>::max(), tau_escape, nullptr, killed);
      if (tau_escape > 1.e-10 && !killed) {
        double weight;
        forced_interaction(g, tau_escape, tau, weight);
        p.energy = p.energy * weight;
      } else {
        tau = g.rng.random_exp();
      }
    } else {
      tau = g.rng.random_exp();
    }
    double tau_achieved;
    grid_integrate_noenergy(g, p, tau, tau_achieved);
    if (p.reabsorbed) {
      int64_t ia;
      for (ia = 1; ia <= g.conf.n_reabs_max; ia++) {
        emit(g, p, true, p.reabsorbed_id, p.energy, p.inu);
        g.n_reabsorptions++;
        if (make_peeled) peeloff_photon(g, p, false);
        tau = g.rng.random_exp();
        grid_integrate_noenergy(g, p, tau, tau_achieved);
        if (!p.reabsorbed) break;
      }
      if (ia == g.conf.n_reabs_max + 1) {
        g.killed_photons_int++;
        p.killed = true;
        break;
      }
    }
    if (p.killed || escaped(g, p.icell)) {
      if (!p.killed) g.n_escaped++;
      break;
    }
    if (interactions == n_inter_max + 1) {
      g.killed_photons_int++;
      p.killed = true;
      break;
    }
    interact(g, p, true);
    p.killed = (g.conf.kill_on_scatter && p.scattered) || (p.energy < energy_initial * g.monochromatic_energy_threshold);
    if (p.killed) break;
    if (make_peeled) peeloff_photon(g, p, false);
  }
}

// do_final_mono (iter_final_mono.f90:58-229) for ONE frequency: n_sources of n_total_sources source packets,
// then n_thermal of n_total_thermal thermal packets (the weights divide by the totals)
void final_mono_photons(orc_ctx &g, int inu, int64_t n_sources, int64_t n_total_sources, int64_t n_thermal,
                        int64_t n_total_thermal, bool peeloff_scattering_only) {
  if (inu < 1 || inu > (int)g.frequencies.size()) throw OracleError{"incorrect inu"};
  const bool make_peeled = !g.peeled.group_id.empty();
  Photon p;
  for (int64_t ip = 1; ip <= n_sources; ip++) {
    emit(g, p, false, 0, 0.0, inu);
    g.n_photons_run++;
    p.energy = p.energy / (double)n_total_sources;
    if (make_peeled && !peeloff_scattering_only) peeloff_photon(g, p, false);
    propagate_mono(g, p);
  }
  if (n_thermal > 0 && g.n_dust > 0) {
    const bool empty = setup_monochromatic_grid_pdfs(g, inu);
    if (!empty) {
      for (int64_t ip = 1; ip <= n_thermal; ip++) {
        p = emit_from_monochromatic_grid_pdf(g, inu);
        g.n_photons_run++;
        if (p.energy > 0.0) {
          p.energy = p.energy * g.energy_abs_tot[p.dust_id - 1] / (double)n_total_thermal * (double)g.n_dust;
          if (make_peeled && !peeloff_scattering_only) peeloff_photon(g, p, false);
          propagate_mono(g, p);
        }
      }
    }
  }
}

// image_write (image_type.f90:608-788): the arrays as written, in memory order
void image_written(const Image &im, bool sed, std::vector<double> &out, std::vector<double> &unc) {
  // with filters the flux stays in F_nu dnu: the filter carries the normalisation (image_type.f90:649-657)
  double dnunorm = (im.c.use_filters || im.use_exact_nu) ? 1.0
                                    : std::pow(im.nu_max / im.nu_min, +0.5 / (double)im.n_nu) -
                                          std::pow(im.nu_max / im.nu_min, -0.5 / (double)im.n_nu);
  out = sed ? im.sed : im.img;
  if (im.c.uncertainties) {
    const std::vector<double> &s2 = sed ? im.sed2 : im.img2;
    unc.resize(s2.size());
    for (size_t i = 0; i < s2.size(); i++) unc[i] = std::sqrt(s2[i]);
  } else {
    unc.clear();
  }
  if (!im.use_exact_nu) {
    for (auto &v : out) v = v / dnunorm;
    for (auto &v : unc) v = v / dnunorm;
  } else {
    // image_type.f90:679-682,737-740: nu F_nu at the exact frequencies
    for (size_t i = 0; i < out.size(); i++) out[i] = out[i] * im.nu[i % (size_t)im.n_nu];
    for (size_t i = 0; i < unc.size(); i++) unc[i] = unc[i] * im.nu[i % (size_t)im.n_nu];
  }
  if (sed) {
    const size_t n_nu = im.n_nu, n_ap = im.c.n_ap;
    const size_t outer = out.size() / (n_nu * n_ap);
    for (size_t o = 0; o < outer; o++)
      for (size_t ia = 1; ia < n_ap; ia++)
        for (size_t inu = 0; inu < n_nu; inu++) {
          size_t k = inu + n_nu * (ia + n_ap * o), km = inu + n_nu * (ia - 1 + n_ap * o);
          out[k] = out[km] + out[k];
          if (!unc.empty()) unc[k] = std::sqrt(unc[km] * unc[km] + unc[k] * unc[k]);
        }
  }
}

int fail(orc_ctx *g, const std::string &m) {
  g_last_error = m;
  if (g) g->error = m;
  return HYP_ERR_INVALID;
}

}  // namespace

// ---------------------------------------------------------------------------
// C API (mirrors include/hyperion_b200.h with the orc_ prefix)
// ---------------------------------------------------------------------------
extern "C" {

const char *orc_last_error(void) { return g_last_error.c_str(); }

int orc_ctx_create(orc_ctx **out) {
  *out = new orc_ctx();
  hyp_run_conf &c = (*out)->conf;
  memset(&c, 0, sizeof c);
  c.seed = -124902;
  c.n_inter_max = 1000000;
  c.n_reabs_max = 1000000;
  c.enforce_energy_range = 1;
  c.propagation_check_frequency = 1.e-3;
  return 0;
}

void orc_ctx_destroy(orc_ctx *g) { delete g; }

int orc_set_grid_cartesian(orc_ctx *g, int32_t n1, int32_t n2, int32_t n3, const double *w1, const double *w2,
                           const double *w3) {
  g->n1 = n1;
  g->n2 = n2;
  g->n3 = n3;
  g->n_cells = n1 * n2 * n3;
  g->w1.assign(w1, w1 + n1 + 1);
  g->w2.assign(w2, w2 + n2 + 1);
  g->w3.assign(w3, w3 + n3 + 1);
  g->dx.resize(n1);
  g->dy.resize(n2);
  g->dz.resize(n3);
  for (int i = 0; i < n1; i++) g->dx[i] = w1[i + 1] - w1[i];
  for (int i = 0; i < n2; i++) g->dy[i] = w2[i + 1] - w2[i];
  for (int i = 0; i < n3; i++) g->dz[i] = w3[i + 1] - w3[i];
  g->volume.resize(g->n_cells);
  for (int i3 = 0; i3 < n3; i3++)
    for (int i2 = 0; i2 < n2; i2++)
      for (int i1 = 0; i1 < n1; i1++)
        g->volume[(size_t)i3 * n1 * n2 + i2 * n1 + i1] = g->dx[i1] * g->dy[i2] * g->dz[i3];
  g->ew1.resize(n1 + 1);
  g->ew2.resize(n2 + 1);
  g->ew3.resize(n3 + 1);
  for (int i = 0; i <= n1; i++) g->ew1[i] = 3 * spacing(w1[i]);
  for (int i = 0; i <= n2; i++) g->ew2[i] = 3 * spacing(w2[i]);
  for (int i = 0; i <= n3; i++) g->ew3[i] = 3 * spacing(w3[i]);
  return 0;
}

// setup_grid_geometry (grid_geometry_spherical_3d.f90:92-203): w1 = r, w2 = theta, w3 = phi walls
int orc_set_grid_spherical(orc_ctx *g, int32_t n1, int32_t n2, int32_t n3, const double *w1, const double *w2,
                           const double *w3) {
  g->grid_type = 1;
  g->n1 = n1;
  g->n2 = n2;
  g->n3 = n3;
  g->n_cells = n1 * n2 * n3;
  g->w1.assign(w1, w1 + n1 + 1);
  g->w2.assign(w2, w2 + n2 + 1);
  g->w3.assign(w3, w3 + n3 + 1);
  for (int i = 0; i <= n1; i++)
    if (w1[i] < 0.) return fail(g, "r walls should be positive");
  for (int i = 0; i <= n2; i++)
    if (w2[i] < 0. || w2[i] > PI_F) return fail(g, "theta walls should be between 0 and pi");
  for (int i = 0; i <= n3; i++)
    if (w3[i] < 0. || w3[i] > TWOPI_F) return fail(g, "phi walls should be between 0 and 2*pi");
  std::vector<double> dr3(n1), dcost(n2), dphi(n3);
  for (int i = 0; i < n1; i++) dr3[i] = w1[i + 1] * w1[i + 1] * w1[i + 1] - w1[i] * w1[i] * w1[i];
  for (int i = 0; i < n2; i++) dcost[i] = std::cos(w2[i]) - std::cos(w2[i + 1]);
  for (int i = 0; i < n3; i++) dphi[i] = w3[i + 1] - w3[i];
  g->volume.resize(g->n_cells);
  for (int i3 = 0; i3 < n3; i3++)
    for (int i2 = 0; i2 < n2; i2++)
      for (int i1 = 0; i1 < n1; i1++)
        g->volume[(size_t)i3 * n1 * n2 + i2 * n1 + i1] = dr3[i1] * dcost[i2] * dphi[i3] / 3.0;
  for (double v : g->volume)
    if (v == 0.0) return fail(g, "all volumes should be greater than zero");
  g->wr2.resize(n1 + 1);
  for (int i = 0; i <= n1; i++) g->wr2[i] = w1[i] * w1[i];
  g->wtanp.resize(n3 + 1);
  for (int i = 0; i <= n3; i++) g->wtanp[i] = std::tan(w3[i]);
  g->wtant.resize(n2 + 1);
  g->wcost.resize(n2 + 1);
  g->wsint.resize(n2 + 1);
  g->wtant2.resize(n2 + 1);
  for (int i = 0; i <= n2; i++) {
    g->wtant[i] = std::tan(w2[i]);
    g->wcost[i] = std::cos(w2[i]);
    g->wsint[i] = std::sin(w2[i]);
    g->wtant2[i] = g->wtant[i] * g->wtant[i];
  }
  g->midplane = -1;
  bool any = false;
  for (int i = 0; i <= n2; i++)
    if (std::fabs(w2[i] - PI_F / 2.0) < (double)1.e-6f) any = true;
  if (any) {
    int best = 0;
    for (int i = 1; i <= n2; i++)
      if (std::fabs(w2[i] - PI_F / 2.0) < std::fabs(w2[best] - PI_F / 2.0)) best = i;
    g->midplane = best + 1;
  }
  g->ew1.resize(n1 + 1);
  g->ew2.assign(n2 + 1, 3 * spacing(1.0));
  g->ew3.assign(n3 + 1, 3 * spacing(1.0));
  for (int i = 0; i <= n1; i++) g->ew1[i] = 3 * spacing(w1[i]);
  return 0;
}

// setup_grid_geometry (grid_geometry_cylindrical_3d.f90:90-177): w1 = w, w2 = z, w3 = phi walls
int orc_set_grid_cylindrical(orc_ctx *g, int32_t n1, int32_t n2, int32_t n3, const double *w1, const double *w2,
                             const double *w3) {
  g->grid_type = 2;
  g->n1 = n1;
  g->n2 = n2;
  g->n3 = n3;
  g->n_cells = n1 * n2 * n3;
  g->w1.assign(w1, w1 + n1 + 1);
  g->w2.assign(w2, w2 + n2 + 1);
  g->w3.assign(w3, w3 + n3 + 1);
  for (int i = 0; i <= n1; i++)
    if (w1[i] < 0.) return fail(g, "w walls should be positive");
  for (int i = 0; i <= n3; i++)
    if (w3[i] < 0. || w3[i] > TWOPI_F) return fail(g, "phi walls should be between 0 and 2*pi");
  g->volume.resize(g->n_cells);
  for (int i3 = 0; i3 < n3; i3++)
    for (int i2 = 0; i2 < n2; i2++)
      for (int i1 = 0; i1 < n1; i1++)
        g->volume[(size_t)i3 * n1 * n2 + i2 * n1 + i1] =
            (w1[i1 + 1] * w1[i1 + 1] - w1[i1] * w1[i1]) * (w2[i2 + 1] - w2[i2]) * (w3[i3 + 1] - w3[i3]) / 2.0;
  for (double v : g->volume)
    if (v == 0.0) return fail(g, "all volumes should be greater than zero");
  g->wr2.resize(n1 + 1);
  for (int i = 0; i <= n1; i++) g->wr2[i] = w1[i] * w1[i];
  g->wtanp.resize(n3 + 1);
  for (int i = 0; i <= n3; i++) g->wtanp[i] = std::tan(w3[i]);
  g->ew1.resize(n1 + 1);
  g->ew2.resize(n2 + 1);
  g->ew3.assign(n3 + 1, 3 * spacing(1.0));
  for (int i = 0; i <= n1; i++) g->ew1[i] = 3 * spacing(w1[i]);
  for (int i = 0; i <= n2; i++) g->ew2[i] = 3 * spacing(w2[i]);
  return 0;
}

// setup_grid_geometry + octree_setup_indiv (grid_geometry_octree.f90:148-262): refined is the depth-first
// list of refinement flags, (x, y, z) the centre and (dx, dy, dz) the HALF-widths of the root cell
// setup_grid_geometry (grid_geometry_voronoi.f90:92-187)
int orc_set_grid_voronoi(orc_ctx *g, int32_t n_cells, const double *coords, const double *bb_min, const double *bb_max,
                         const double *volume, const int32_t *sparse_idx, const int32_t *sparse_neighs, const double *box) {
  g->grid_type = 5;
  g->n_cells = n_cells;
  g->n1 = n_cells;
  g->n2 = g->n3 = 1;
  g->vx.resize(n_cells); g->vy.resize(n_cells); g->vz.resize(n_cells);
  g->vbb.resize((size_t)6 * n_cells);
  g->volume.assign(n_cells, 0.0);
  g->mask_map.clear();
  for (int i = 0; i < n_cells; i++) {
    g->vx[i] = coords[3 * i]; g->vy[i] = coords[3 * i + 1]; g->vz[i] = coords[3 * i + 2];
    for (int a = 0; a < 3; a++) {
      g->vbb[(size_t)6 * i + 2 * a] = bb_min[3 * i + a];
      g->vbb[(size_t)6 * i + 2 * a + 1] = bb_max[3 * i + a];
    }
    if (volume[i] > 0.0) g->mask_map.push_back(i + 1);       // geo%mask = geo%volume > 0
    g->volume[i] = volume[i] < 0.0 ? 0.0 : volume[i];
  }
  g->n_masked = (int)g->mask_map.size();
  g->vidx.assign(sparse_idx, sparse_idx + n_cells + 1);
  g->vneigh.assign(sparse_neighs, sparse_neighs + sparse_idx[n_cells]);
  for (int a = 0; a < 6; a++) g->vbox[a] = box[a];
  // buckets of about two sites
  const int per_axis = std::max(1, (int)std::cbrt((double)n_cells / 2.0));
  for (int a = 0; a < 3; a++) g->vg[a] = per_axis;
  const int nb = per_axis * per_axis * per_axis;
  std::vector<int> bucket(n_cells);
  g->vg_start.assign(nb + 1, 0);
  for (int i = 0; i < n_cells; i++) {
    const double pos[3] = {g->vx[i], g->vy[i], g->vz[i]};
    int b[3];
    for (int a = 0; a < 3; a++)
      b[a] = std::min(std::max((int)((pos[a] - box[2 * a]) / ((box[2 * a + 1] - box[2 * a]) / per_axis)), 0), per_axis - 1);
    bucket[i] = (b[2] * per_axis + b[1]) * per_axis + b[0];
    g->vg_start[bucket[i] + 1]++;
  }
  for (int k = 0; k < nb; k++) g->vg_start[k + 1] += g->vg_start[k];
  g->vg_sites.assign(n_cells, 0);
  std::vector<int> fill(g->vg_start.begin(), g->vg_start.end() - 1);
  for (int i = 0; i < n_cells; i++) g->vg_sites[fill[bucket[i]]++] = i;
  return 0;
}

int orc_set_grid_octree(orc_ctx *g, int32_t n_cells, const int32_t *refined, double x, double y, double z, double dx,
                        double dy, double dz) {
  g->grid_type = 3;
  g->n_cells = n_cells;
  g->n1 = n_cells;
  g->n2 = g->n3 = 1;
  g->ox.assign(n_cells, 0.0);
  g->oy = g->oz = g->odx = g->ody = g->odz = g->ox;
  g->orefined.assign(n_cells, 0);
  g->ochildren.assign((size_t)8 * n_cells, 0);
  g->oparent.assign(n_cells, 0);
  g->oparent_subcell.assign(n_cells, 0);
  for (int i = 0; i < n_cells; i++) g->orefined[i] = refined[i] == 1;
  g->mask_map.clear();
  for (int i = 0; i < n_cells; i++)
    if (refined[i] == 0) g->mask_map.push_back(i + 1);
  g->n_masked = (int)g->mask_map.size();
  g->ox[0] = x; g->oy[0] = y; g->oz[0] = z;
  g->odx[0] = dx; g->ody[0] = dy; g->odz[0] = dz;
  int n_filled = 1;
  // iterative form of the recursion: a stack of (parent, next child index)
  std::vector<std::pair<int, int>> stack;
  if (g->orefined[0]) stack.push_back({1, 1});
  while (!stack.empty()) {
    int parent = stack.back().first, ic = stack.back().second;
    if (ic > 8) {
      stack.pop_back();
      continue;
    }
    stack.back().second = ic + 1;
    n_filled = n_filled + 1;
    if (n_filled > n_cells) return fail(g, "refined array is not self-consistent");
    int child = n_filled;
    g->ochildren[(size_t)8 * (parent - 1) + ic - 1] = child;
    int sx = 1, sy = 1, sz = 1;
    if ((ic - 1) % 2 == 0) sx = -sx;
    if (((ic - 1) / 2) % 2 == 0) sy = -sy;
    if (((ic - 1) / 4) % 2 == 0) sz = -sz;
    g->ox[child - 1] = g->ox[parent - 1] + sx * g->odx[parent - 1] / 2.0;
    g->oy[child - 1] = g->oy[parent - 1] + sy * g->ody[parent - 1] / 2.0;
    g->oz[child - 1] = g->oz[parent - 1] + sz * g->odz[parent - 1] / 2.0;
    g->odx[child - 1] = g->odx[parent - 1] / 2.0;
    g->ody[child - 1] = g->ody[parent - 1] / 2.0;
    g->odz[child - 1] = g->odz[parent - 1] / 2.0;
    g->oparent[child - 1] = parent;
    g->oparent_subcell[child - 1] = ic;
    if (g->orefined[child - 1]) stack.push_back({child, 1});
  }
  if (n_filled != n_cells) return fail(g, "refined array is not self-consistent");
  g->volume.resize(n_cells);
  for (int i = 0; i < n_cells; i++) g->volume[i] = g->odx[i] * g->ody[i] * g->odz[i] * 8.0;
  for (double v : g->volume)
    if (v == 0.0) return fail(g, "all volumes should be greater than zero");
  g->oct_eps = spacing(std::max(dx, std::max(dy, dz))) * 3.0;
  return 0;
}

// read_grid / setup_grid_geometry (grid_geometry_amr.f90:111-507).  n_grids[n_levels]; per grid (level-major):
// dims[3] = n1, n2, n3 and bounds[6] = xmin, xmax, ymin, ymax, zmin, zmax
int orc_set_grid_amr(orc_ctx *g, int32_t n_levels, const int32_t *n_grids, const int32_t *dims, const double *bounds) {
  typedef orc_ctx::AmrGrid Grid;
  g->grid_type = 4;
  g->levels.assign(n_levels, std::vector<Grid>());
  int k = 0, start_id = 1;
  double min_width = std::numeric_limits<double>::max();
  for (int il = 0; il < n_levels; il++) {
    g->levels[il].resize(n_grids[il]);
    for (int ig = 0; ig < n_grids[il]; ig++, k++) {
      Grid &G = g->levels[il][ig];
      G.n1 = dims[3 * k];
      G.n2 = dims[3 * k + 1];
      G.n3 = dims[3 * k + 2];
      G.n_cells = G.n1 * G.n2 * G.n3;
      G.xmin = bounds[6 * k]; G.xmax = bounds[6 * k + 1];
      G.ymin = bounds[6 * k + 2]; G.ymax = bounds[6 * k + 3];
      G.zmin = bounds[6 * k + 4]; G.zmax = bounds[6 * k + 5];
      // linspace_dp (lib_array.f90:284-302)
      auto linspace = [](double a, double b, int n, std::vector<double> &x) {
        x.resize(n);
        for (int i = 1; i <= n; i++) x[i - 1] = (b - a) * (double)(i - 1) / (double)(n - 1) + a;
      };
      linspace(G.xmin, G.xmax, G.n1 + 1, G.w1);
      linspace(G.ymin, G.ymax, G.n2 + 1, G.w2);
      linspace(G.zmin, G.zmax, G.n3 + 1, G.w3);
      G.width[0] = (G.xmax - G.xmin) / (double)G.n1;
      G.width[1] = (G.ymax - G.ymin) / (double)G.n2;
      G.width[2] = (G.zmax - G.zmin) / (double)G.n3;
      G.volume = G.width[0] * G.width[1] * G.width[2];
      G.goto_grid.assign((size_t)(G.n1 + 2) * (G.n2 + 2) * (G.n3 + 2), 0);
      G.goto_level = G.goto_grid;
      G.start_id = start_id;
      start_id += G.n_cells;
      for (int a = 0; a < 3; a++)
        if (G.width[a] < min_width) min_width = G.width[a];
    }
  }
  g->n_cells = start_id - 1;
  g->n1 = g->n_cells;
  g->n2 = g->n3 = 1;
  g->volume.resize(g->n_cells);
  g->cell_ilevel.resize(g->n_cells);
  g->cell_igrid.resize(g->n_cells);
  g->cell_i1.resize(g->n_cells);
  g->cell_i2.resize(g->n_cells);
  g->cell_i3.resize(g->n_cells);
  int ic = 0;
  for (int il = 0; il < n_levels; il++)
    for (size_t ig = 0; ig < g->levels[il].size(); ig++) {
      const Grid &G = g->levels[il][ig];
      for (int i3 = 1; i3 <= G.n3; i3++)
        for (int i2 = 1; i2 <= G.n2; i2++)
          for (int i1 = 1; i1 <= G.n1; i1++) {
            g->volume[ic] = G.volume;
            g->cell_ilevel[ic] = il + 1;
            g->cell_igrid[ic] = (int)ig + 1;
            g->cell_i1[ic] = i1;
            g->cell_i2[ic] = i2;
            g->cell_i3[ic] = i3;
            ic++;
          }
    }
  for (double v : g->volume)
    if (v == 0.0) return fail(g, "all volumes should be greater than zero");
  g->amr_eps = min_width / 2.0;
  auto intersect = [](const Grid &a, const Grid &b) {
    return !(a.xmax < b.xmin || a.xmin > b.xmax || a.ymax < b.ymin || a.ymin > b.ymax || a.zmax < b.zmin || a.zmin > b.zmax);
  };
  auto close = [](const Grid &a, const Grid &b) {
    return !(a.xmax < b.xmin - b.width[0] * 0.5 || a.xmin > b.xmax + b.width[0] * 0.5 || a.ymax < b.ymin - b.width[1] * 0.5 ||
             a.ymin > b.ymax + b.width[1] * 0.5 || a.zmax < b.zmin - b.width[2] * 0.5 || a.zmin > b.zmax + b.width[2] * 0.5);
  };
  // cells covered by a grid of the next level point to it (:355-381)
  for (int il1 = n_levels - 2; il1 >= 0; il1--)
    for (Grid &G1 : g->levels[il1])
      for (size_t ig2 = 0; ig2 < g->levels[il1 + 1].size(); ig2++) {
        const Grid &G2 = g->levels[il1 + 1][ig2];
        if (!intersect(G1, G2)) continue;
        for (int i1 = 1; i1 <= G1.n1; i1++)
          for (int i2 = 1; i2 <= G1.n2; i2++)
            for (int i3 = 1; i3 <= G1.n3; i3++) {
              Vec r{0.5 * (G1.w1[i1 - 1] + G1.w1[i1]), 0.5 * (G1.w2[i2 - 1] + G1.w2[i2]), 0.5 * (G1.w3[i3 - 1] + G1.w3[i3])};
              if (amr_in_grid(G2, r)) {
                G1.goto_grid[G1.gidx(i1, i2, i3)] = (int)ig2 + 1;
                G1.goto_level[G1.gidx(i1, i2, i3)] = il1 + 2;
              }
            }
      }
  // ghost cells: which grid does a packet land in one step outside the grid (:383-487)
  for (int il1 = 0; il1 < n_levels; il1++)
    for (size_t ig1 = 0; ig1 < g->levels[il1].size(); ig1++) {
      Grid &G1 = g->levels[il1][ig1];
      for (int il2 = il1; il2 >= 0; il2--)
        for (size_t ig2 = 0; ig2 < g->levels[il2].size(); ig2++) {
          const Grid &G2 = g->levels[il2][ig2];
          if (!(close(G1, G2) && (ig1 != ig2 || il1 != il2))) continue;
          auto mark = [&](int i1, int i2, int i3, const Vec &r) {
            if (amr_in_grid(G2, r) && G1.goto_grid[G1.gidx(i1, i2, i3)] == 0) {
              G1.goto_grid[G1.gidx(i1, i2, i3)] = (int)ig2 + 1;
              G1.goto_level[G1.gidx(i1, i2, i3)] = il2 + 1;
            }
          };
          auto cx = [&](int i) { return 0.5 * (G1.w1[i - 1] + G1.w1[i]); };
          auto cy = [&](int i) { return 0.5 * (G1.w2[i - 1] + G1.w2[i]); };
          auto cz = [&](int i) { return 0.5 * (G1.w3[i - 1] + G1.w3[i]); };
          for (int i2 = 1; i2 <= G1.n2; i2++)
            for (int i3 = 1; i3 <= G1.n3; i3++) mark(0, i2, i3, Vec{G1.xmin - G1.width[0] * 0.5, cy(i2), cz(i3)});
          for (int i2 = 1; i2 <= G1.n2; i2++)
            for (int i3 = 1; i3 <= G1.n3; i3++) mark(G1.n1 + 1, i2, i3, Vec{G1.xmax + G1.width[0] * 0.5, cy(i2), cz(i3)});
          for (int i1 = 1; i1 <= G1.n1; i1++)
            for (int i3 = 1; i3 <= G1.n3; i3++) mark(i1, 0, i3, Vec{cx(i1), G1.ymin - G1.width[1] * 0.5, cz(i3)});
          for (int i1 = 1; i1 <= G1.n1; i1++)
            for (int i3 = 1; i3 <= G1.n3; i3++) mark(i1, G1.n2 + 1, i3, Vec{cx(i1), G1.ymax + G1.width[1] * 0.5, cz(i3)});
          for (int i1 = 1; i1 <= G1.n1; i1++)
            for (int i2 = 1; i2 <= G1.n2; i2++) mark(i1, i2, 0, Vec{cx(i1), cy(i2), G1.zmin - G1.width[2] * 0.5});
          for (int i1 = 1; i1 <= G1.n1; i1++)
            for (int i2 = 1; i2 <= G1.n2; i2++) mark(i1, i2, G1.n3 + 1, Vec{cx(i1), cy(i2), G1.zmax + G1.width[2] * 0.5});
        }
    }
  // valid cells: not covered by a finer grid (:489-506)
  g->mask_map.clear();
  for (int icell = 1; icell <= g->n_cells; icell++) {
    Cell c = amr_cell_1d(*g, icell);
    const Grid &G = g->levels[c.ilevel - 1][c.igrid - 1];
    if (G.goto_grid[G.gidx(c.i1, c.i2, c.i3)] == 0) g->mask_map.push_back(icell);
  }
  g->n_masked = (int)g->mask_map.size();
  return 0;
}

int orc_add_dust(orc_ctx *g, const hyp_dust_tables *t) {
  try {
    g->d.emplace_back();
    g->d.back().setup(*t);
    g->n_dust = (int)g->d.size();
    if (g->n_dust > MAX_DUST) return fail(g, "too many dust types");
  } catch (OracleError &e) {
    return fail(g, e.msg);
  }
  return 0;
}

int orc_add_source(orc_ctx *g, const hyp_source *s) {
  try {
    Source src;
    src.type = s->type;
    src.peeloff = s->peeloff != 0;
    src.luminosity = s->luminosity;
    src.position = Vec{s->x, s->y, s->z};
    src.radius = s->radius;
    src.limb_darkening = s->limb_darkening != 0;
    src.freq_type = s->spectrum_type;
    src.temperature = s->temperature;
    if (s->spectrum_type == HYP_SPECTRUM_LTE && s->type != HYP_SOURCE_MAP)
      return fail(g, "only map sources can have an LTE spectrum");
    if (s->spectrum_type == HYP_SPECTRUM_TABLE) {
      for (int i = 0; i + 1 < s->n_spec; i++)
        if (s->spec_nu[i + 1] < s->spec_nu[i])
          return fail(g, "spectrum frequency should be monotonically increasing");
      src.spectrum.set(s->spec_nu, s->spec_fnu, s->n_spec, true);
    }
    if (s->type == HYP_SOURCE_EXTERN_BOX) {
      src.xmin = s->box[0]; src.xmax = s->box[1];
      src.ymin = s->box[2]; src.ymax = s->box[3];
      src.zmin = s->box[4]; src.zmax = s->box[5];
      const double dx = src.xmax - src.xmin, dy = src.ymax - src.ymin, dz = src.zmax - src.zmin;
      const double areas[6] = {dy * dz, dy * dz, dz * dx, dz * dx, dx * dy, dx * dy};  // source_type.f90:229
      src.face.set(areas, 6);
    } else if (s->type == HYP_SOURCE_PLANE_PARALLEL) {
      src.direction = angle3d_deg(s->theta, s->phi);
    } else if (s->type == HYP_SOURCE_POINT_COLLECTION) {
      if (s->n_points < 1 || !s->points_xyz || !s->points_lum) return fail(g, "point collection is empty");
      for (int64_t i = 0; i < s->n_points; i++)
        src.position_collection.push_back(Vec{s->points_xyz[3 * i], s->points_xyz[3 * i + 1], s->points_xyz[3 * i + 2]});
      src.luminosity = 0.0;
      for (int64_t i = 0; i < s->n_points; i++) src.luminosity += s->points_lum[i];  // sum() (source_type.f90:268)
      src.collection_pdf.set(s->points_lum, (int)s->n_points);
    } else if (s->type == HYP_SOURCE_MAP) {
      // grid_load_pdf_map (grid_geometry_common_3d.f90:47-63)
      if (!s->map || s->n_map != g->n_cells) return fail(g, "luminosity map should have one entry per cell");
      src.luminosity_map.set(s->map, (int)s->n_map);
    } else if (s->type != HYP_SOURCE_POINT && s->type != HYP_SOURCE_SPHERE && s->type != HYP_SOURCE_EXTERN_SPH) {
      return fail(g, "source type not restated in the oracle");
    }
    if (s->type == HYP_SOURCE_SPHERE && s->n_spots > 0) {
      // source_read (source_type.f90:150-188)
      std::vector<double> pdf((size_t)s->n_spots + 1);
      pdf[s->n_spots] = s->luminosity;
      for (int i = 0; i < s->n_spots; i++) {
        const hyp_spot &q = s->spots[i];
        Source::Spot sp;
        pdf[i] = q.luminosity;
        src.luminosity = src.luminosity + q.luminosity;
        sp.a = angle3d_deg(q.longitude, q.latitude);
        sp.cost = std::cos(q.radius * (PI / 180.0));
        sp.freq_type = q.spectrum_type;
        sp.temperature = q.temperature;
        if (q.spectrum_type == HYP_SPECTRUM_TABLE) sp.spectrum.set(q.spec_nu, q.spec_fnu, q.n_spec, true);
        else if (q.spectrum_type != HYP_SPECTRUM_BLACKBODY) return fail(g, "Spot cannot have LTE spectrum");
        src.spot.push_back(sp);
      }
      src.spot_pdf.n = s->n_spots + 1;
      src.spot_pdf.pdf = pdf;
      src.spot_pdf.cdf.assign(pdf.size(), 0.0);
      src.spot_pdf.find_cdf();   // allocate_pdf + find_cdf: the pdf itself is not normalised (:160-188)
    }
    src.intersect = (s->type == HYP_SOURCE_SPHERE);
    if (src.intersect) g->any_intersect = true;
    g->s.push_back(src);
  } catch (OracleError &e) {
    return fail(g, e.msg);
  }
  return 0;
}

int orc_set_run_conf(orc_ctx *g, const hyp_run_conf *c) {
  g->conf = *c;
  return 0;
}

int orc_set_density(orc_ctx *g, int32_t n_dust, const double *density) {
  if (n_dust != g->n_dust) return fail(g, "density array has wrong number of dust types");
  g->density.assign(density, density + (size_t)n_dust * g->n_cells);
  return 0;
}

int orc_set_specific_energy(orc_ctx *g, const double *se, const double *min_e) {
  size_t n = (size_t)g->n_dust * g->n_cells;
  g->minimum_specific_energy.assign(g->n_dust, 0.0);
  if (min_e)
    for (int id = 0; id < g->n_dust; id++) g->minimum_specific_energy[id] = min_e[id];
  g->specific_energy.resize(n);
  g->specific_energy_from_file = se != nullptr;
  if (se) {
    g->specific_energy.assign(se, se + n);
  } else {
    for (int id = 0; id < g->n_dust; id++)
      for (int ic = 0; ic < g->n_cells; ic++)
        g->specific_energy[(size_t)id * g->n_cells + ic] = g->minimum_specific_energy[id];
  }
  return 0;
}

// the remaining steps of setup_initial (setup_rt.f90:27-304) and main.f90:157-167
int orc_finalize_setup(orc_ctx *g, int32_t rank) {
  try {
    size_t n = (size_t)g->n_dust * g->n_cells;
    if (g->density.size() != n) return fail(g, "density not set");
    // setup_grid_physics applies the mask (refined octree nodes hold no dust): grid_physics_3d.f90:156-164
    if (!g->mask_map.empty() && !g->setup_done) {
      std::vector<char> valid(g->n_cells, 0);
      for (int ic : g->mask_map) valid[ic - 1] = 1;
      for (int id = 0; id < g->n_dust; id++)
        for (int ic = 0; ic < g->n_cells; ic++)
          if (!valid[ic]) g->density[(size_t)id * g->n_cells + ic] = 0.0;
      if (g->specific_energy_from_file)
        for (int id = 0; id < g->n_dust; id++)
          for (int ic = 0; ic < g->n_cells; ic++)
            if (!valid[ic]) g->specific_energy[(size_t)id * g->n_cells + ic] = 0.0;
    }
    if (g->specific_energy.size() != n) orc_set_specific_energy(g, nullptr, nullptr);
    if (g->conf.specific_energy_additional && !g->setup_done) {
      // grid_physics_3d.f90:213-235,241-243
      if (!g->specific_energy_from_file)
        return fail(g, "cannot specify specific_energy_type since specific_energy was not given");
      g->specific_energy_additional = g->specific_energy;
      for (int id = 0; id < g->n_dust; id++)
        for (int ic = 0; ic < g->n_cells; ic++)
          g->specific_energy[(size_t)id * g->n_cells + ic] = g->minimum_specific_energy[id];
    }
    g->specific_energy_sum.assign(n, 0.0);
    if (g->n_nu_bins > 0 && !g->setup_done) {
      // grid_physics_3d.f90:135-143,199-207,229-233,249-252: zeros when the specific energy comes from the file,
      // the minimum specific energy otherwise (and with specific_energy_type = 'additional')
      g->specific_energy_spectrum.assign(n * g->n_nu_bins, 0.0);
      g->specific_energy_sum_spectrum.assign(n * g->n_nu_bins, 0.0);
      if (!g->specific_energy_from_file || g->conf.specific_energy_additional)
        for (int ib = 0; ib < g->n_nu_bins; ib++)
          for (int id = 0; id < g->n_dust; id++)
            for (int ic = 0; ic < g->n_cells; ic++)
              g->specific_energy_spectrum[((size_t)ib * g->n_dust + id) * g->n_cells + ic] = g->minimum_specific_energy[id];
      // setup_j_nu_bin_fractions (:325-348), get_j_nu_bin_fractions (dust_type_4elem.f90:752-778)
      g->n_jnu_max = 0;
      for (int id = 0; id < g->n_dust; id++) g->n_jnu_max = std::max(g->n_jnu_max, g->d[id].n_jnu);
      g->j_nu_bin_frac.assign((size_t)g->n_dust * g->n_jnu_max * g->n_nu_bins, 0.0);
      for (int id = 0; id < g->n_dust; id++)
        for (int iv = 0; iv < g->d[id].n_jnu; iv++) {
          const PdfCont &j = g->d[id].j_nu[iv];
          double *frac = &g->j_nu_bin_frac[((size_t)id * g->n_jnu_max + iv) * g->n_nu_bins];
          for (int ib = 0; ib < g->n_nu_bins; ib++)
            frac[ib] = integral_loglog_subset(j.x.data(), j.pdf.data(), j.n, g->nu_bin_edges[ib], g->nu_bin_edges[ib + 1]);
          const double norm = integral_general(j.x.data(), j.pdf.data(), j.n, trapezium_loglog);
          if (norm > 0.0)
            for (int ib = 0; ib < g->n_nu_bins; ib++) frac[ib] = frac[ib] / norm;
        }
    }
    g->jnu_var_id.assign(n, 0);
    g->jnu_var_frac.assign(n, 0.0);
    g->energy_abs_tot.assign(g->n_dust, 0.0);
    g->absorption.n = g->n_dust;
    g->absorption.pdf.assign(g->n_dust, 0.0);
    g->absorption.cdf.assign(g->n_dust, 0.0);
    check_energy_abs(*g);  // setup_grid_physics (grid_physics_3d.f90:291)
    if (!g->s.empty()) {
      std::vector<double> lum;
      for (auto &s : g->s) lum.push_back(s.luminosity);
      g->luminosity.set(lum.data(), (int)lum.size());
      double tot = 0.0;
      for (double l : lum) tot = tot + l;
      g->energy_total = tot;
    }
    g->rng.set_seed((int)(g->conf.seed + rank));  // mp_set_random_seed (mpi_routines.f90:266-270)
    g->setup_done = true;
  } catch (OracleError &e) {
    return fail(g, e.msg);
  }
  return 0;
}

// do_lucy, first part (iter_lucy.f90:99-112)
int orc_lucy_begin(orc_ctx *g) {
  std::fill(g->specific_energy_sum.begin(), g->specific_energy_sum.end(), 0.0);
  std::fill(g->specific_energy_sum_spectrum.begin(), g->specific_energy_sum_spectrum.end(), 0.0);   // grid_generic.f90:26
  g->energy_current = 0.0;
  g->killed_photons_geo = g->killed_photons_int = 0;
  g->n_crossings = g->n_absorptions = g->n_scatterings = g->n_escaped = g->n_photons_run = 0;
  // grid_reset_energy (grid_generic.f90:18-25); the arrays exist with the PDA or the n_photons output
  // (grid_physics_3d.f90:308-317)
  if (g->conf.use_pda || g->conf.count_photons) {
    g->n_photons.assign(g->n_cells, 0);
    g->last_photon_id.assign(g->n_cells, 0);
  } else {
    g->n_photons.clear();
    g->last_photon_id.clear();
  }
  try {
    precompute_jnu_var(*g);
    if (g->conf.use_mrw) prepare_mrw(*g);  // iter_lucy.f90:109-112
  } catch (OracleError &e) {
    return fail(g, e.msg);
  }
  return 0;
}

int orc_lucy_photons(orc_ctx *g, int64_t n_photons) {
  try {
    lucy_photons(*g, n_photons);
  } catch (OracleError &e) {
    g_last_error = e.msg;
    g->error = e.msg;
    return HYP_ERR_PHYSICS;
  }
  return 0;
}

// accessors used to emulate mp_collect_physical_arrays / mp_sync across oracle "ranks"
double *orc_energy_sum_ptr(orc_ctx *g) { return g->specific_energy_sum.data(); }
double orc_get_energy_current(orc_ctx *g) { return g->energy_current; }
void orc_set_energy_current(orc_ctx *g, double e) { g->energy_current = e; }

// do_lucy, last part (iter_lucy.f90:224-235)
int orc_lucy_finish(orc_ctx *g, hyp_iter_stats *st) {
  try {
    update_energy_abs(*g, g->energy_total / g->energy_current);
    if (g->conf.use_pda) solve_pda(*g, g->pda_exact_limit);   // iter_lucy.f90:227
    sublimate_dust(*g);
  } catch (OracleError &e) {
    return fail(g, e.msg);
  }
  if (st) {
    memset(st, 0, sizeof *st);
    st->energy_emitted = g->energy_current;
    st->n_photons = g->n_photons_run;
    st->killed_geo = g->killed_photons_geo;
    st->killed_int = g->killed_photons_int;
    st->n_crossings = g->n_crossings;
    st->n_absorptions = g->n_absorptions;
    st->n_scatterings = g->n_scatterings;
    st->n_escaped = g->n_escaped;
  }
  return 0;
}

int orc_run_lucy_iteration(orc_ctx *g, int64_t n_photons, hyp_iter_stats *st) {
  int rc = orc_lucy_begin(g);
  if (rc) return rc;
  rc = orc_lucy_photons(g, n_photons);
  if (rc) return rc;
  return orc_lucy_finish(g, st);
}

int orc_get_specific_energy(orc_ctx *g, double *out) {
  memcpy(out, g->specific_energy.data(), g->specific_energy.size() * sizeof(double));
  return 0;
}
// n_photons (grid_physics_3d.f90:38), zeros when it is not kept; the arrays of emulated ranks add up
// (mp_collect_physical_arrays, mpi_routines.f90:303-311)
int orc_get_n_photons(orc_ctx *g, int64_t *out) {
  for (int ic = 0; ic < g->n_cells; ic++) out[ic] = g->n_photons.empty() ? 0 : g->n_photons[ic];
  return 0;
}
int orc_set_n_photons(orc_ctx *g, const int64_t *in) {
  g->n_photons.assign(in, in + g->n_cells);
  return 0;
}
int orc_put_specific_energy(orc_ctx *g, const double *se) {
  g->specific_energy.assign(se, se + (size_t)g->n_dust * g->n_cells);
  return 0;
}
// test hook: the reference's switch between the exact and the iterative solver (10000 PDA cells)
int orc_set_pda_exact_limit(orc_ctx *g, int32_t n) {
  g->pda_exact_limit = n;
  return 0;
}
// solve_pda on the current state; returns the number of PDA cells, or -1
int orc_solve_pda(orc_ctx *g) {
  try {
    return solve_pda(*g, g->pda_exact_limit);
  } catch (OracleError &e) {
    fail(g, e.msg);
    return -1;
  }
}
int orc_get_density(orc_ctx *g, double *out) {
  memcpy(out, g->density.data(), g->density.size() * sizeof(double));
  return 0;
}
int orc_get_energy_sum(orc_ctx *g, double *out) {
  memcpy(out, g->specific_energy_sum.data(), g->specific_energy_sum.size() * sizeof(double));
  return 0;
}
int orc_set_energy_sum(orc_ctx *g, const double *in) {
  memcpy(g->specific_energy_sum.data(), in, g->specific_energy_sum.size() * sizeof(double));
  return 0;
}
// specific_energy_spectrum_bin_edges (setup_rt.f90:98-104, grid_physics_3d.f90:124-129,278-283)
int orc_set_specific_energy_spectrum_bins(orc_ctx *g, int32_t n_edges, const double *edges) {
  if (n_edges < 2) return fail(g, "specific_energy_spectrum_bin_edges should have at least two values");
  for (int i = 1; i < n_edges; i++)
    if (edges[i] <= edges[i - 1]) return fail(g, "specific_energy_spectrum_bin_edges should be strictly increasing");
  g->nu_bin_edges.assign(edges, edges + n_edges);
  g->log_nu_bin_edges.resize(n_edges);
  for (int i = 0; i < n_edges; i++) g->log_nu_bin_edges[i] = std::log10(edges[i]);
  g->n_nu_bins = n_edges - 1;
  return 0;
}
// [n_bins][n_dust][n_cells], as output_grid writes it (grid_generic.f90:68-84)
int orc_get_specific_energy_spectrum(orc_ctx *g, double *out) {
  if (g->n_nu_bins == 0) return fail(g, "specific_energy_spectrum array is not allocated");
  memcpy(out, g->specific_energy_spectrum.data(), g->specific_energy_spectrum.size() * sizeof(double));
  return 0;
}
// the sums of emulated ranks add up (mp_collect_physical_arrays, mpi_routines.f90:292-301)
int orc_get_energy_sum_spectrum(orc_ctx *g, double *out) {
  memcpy(out, g->specific_energy_sum_spectrum.data(), g->specific_energy_sum_spectrum.size() * sizeof(double));
  return 0;
}
int orc_set_energy_sum_spectrum(orc_ctx *g, const double *in) {
  memcpy(g->specific_energy_sum_spectrum.data(), in, g->specific_energy_sum_spectrum.size() * sizeof(double));
  return 0;
}

// unit-test hooks for the numerics

// ---- final / raytracing iterations ------------------------------------------------------------
// peeled_images_setup (images_peeled.f90:272-382)
int orc_add_peeled_group(orc_ctx *g, const hyp_image_conf *c) {
  try {
    if (!(c->n_view > 0)) return fail(g, "n_view should be a positive integer");
    if (c->binned) {
      // setup_final_iteration (setup_rt.f90:318-331) + binned_images_setup (images_binned.f90:41-55)
      for (auto &o : g->peeled.image)
        if (o.c.binned) return fail(g, "can't have more than one binned image group");
      if (g->conf.forced_first_interaction) return fail(g, "can't use binned images with forced first interaction");
      if (c->n_view != c->n_theta * c->n_phi) return fail(g, "binned images: n_view should be n_theta * n_phi");
    }
    Image im;
    image_setup(im, *c, (int)g->s.size(), g->n_dust, g->frequencies);
    PeeledState &P = g->peeled;
    P.image.push_back(im);
    const int ig = (int)P.image.size();
    P.r_peeloff.push_back(Vec{c->peeloff_x, c->peeloff_y, c->peeloff_z});
    for (int iv = 1; iv <= (c->binned ? 0 : c->n_view); iv++) {
      P.group_id.push_back(ig);
      P.view_id.push_back(iv);
      P.viewing_angles.push_back(angle3d_deg(c->theta[iv - 1], c->phi[iv - 1]));
    }
    P.image.back().c.theta = nullptr;
    P.image.back().c.phi = nullptr;
    P.source_spectra.emplace_back(g->s.size());
    P.dust_extinction.emplace_back(g->n_dust);
    P.dust_log10_emissivity.emplace_back(g->n_dust);
  } catch (OracleError &e) {
    return fail(g, e.msg);
  }
  return 0;
}

// do_final, first part (iter_final.f90:84-100)
int orc_final_begin(orc_ctx *g) {
  g->energy_current = 0.0;
  g->killed_photons_geo = g->killed_photons_int = 0;
  g->n_crossings = g->n_absorptions = g->n_scatterings = g->n_escaped = g->n_photons_run = 0;
  g->n_peel_crossings = g->n_peeloffs = 0;
  try {
    precompute_jnu_var(*g);
    if (g->conf.use_mrw) prepare_mrw(*g);  // iter_final.f90:93-96
  } catch (OracleError &e) {
    return fail(g, e.msg);
  }
  return 0;
}

int orc_final_photons(orc_ctx *g, int64_t n_photons, int32_t peeloff_scattering_only) {
  try {
    final_photons(*g, n_photons, peeloff_scattering_only != 0);
  } catch (OracleError &e) {
    return fail(g, e.msg);
  }
  return 0;
}

static void fill_stats(orc_ctx *g, hyp_iter_stats *st) {
  if (!st) return;
  std::memset(st, 0, sizeof *st);
  st->energy_emitted = g->energy_current;
  st->n_photons = g->n_photons_run;
  st->killed_geo = g->killed_photons_geo;
  st->killed_int = g->killed_photons_int;
  st->n_crossings = g->n_crossings;
  st->n_absorptions = g->n_absorptions;
  st->n_scatterings = g->n_scatterings;
  st->n_escaped = g->n_escaped;
  st->n_peel_crossings = g->n_peel_crossings;
  st->n_peeloffs = g->n_peeloffs;
}

int orc_set_monochromatic(orc_ctx *g, int32_t n_nu, const double *frequencies, double energy_threshold) {
  if (n_nu < 1 || !frequencies) return fail(g, "monochromatic mode needs at least one frequency");
  g->frequencies.assign(frequencies, frequencies + n_nu);
  g->monochromatic_energy_threshold = energy_threshold;
  return 0;
}

// do_final_mono for one frequency; first_*_id are unused (one sequential stream per emulated rank)
int orc_final_mono_photons(orc_ctx *g, int32_t inu, int64_t, int64_t n_sources, int64_t n_total_sources, int64_t,
                           int64_t n_dust, int64_t n_total_dust, int32_t peeloff_scattering_only) {
  try {
    g->mono_run = true;
    update_energy_abs_tot(*g);
    final_mono_photons(*g, inu, n_sources, n_total_sources, n_dust, n_total_dust, peeloff_scattering_only != 0);
  } catch (OracleError &e) {
    return fail(g, e.msg);
  }
  return 0;
}

// do_final, last part (iter_final.f90:136-143)
int orc_final_finish(orc_ctx *g, hyp_iter_stats *st) {
  if (g->mono_run) {
    // do_final_mono scales every packet itself (iter_final_mono.f90:118,187)
    g->mono_run = false;
    fill_stats(g, st);
    return 0;
  }
  if (!(g->energy_current > 0.0)) return fail(g, "no photons were emitted in this iteration");
  // peeled_images_adjust_scale / binned_images_adjust_scale (iter_final.f90:142-143, images_binned.f90:35-39)
  for (auto &im : g->peeled.image)
    image_scale(im, im.c.binned ? g->energy_total / g->energy_current * (double)im.c.n_theta * (double)im.c.n_phi
                                : g->energy_total / g->energy_current);
  fill_stats(g, st);
  return 0;
}

int orc_raytracing_photons(orc_ctx *g, int64_t n_sources, int64_t n_thermal, hyp_iter_stats *st) {
  g->killed_photons_geo = g->killed_photons_int = 0;
  g->n_peel_crossings = g->n_peeloffs = 0;
  try {
    raytracing_photons(*g, n_sources, n_thermal);
  } catch (OracleError &e) {
    return fail(g, e.msg);
  }
  fill_stats(g, st);
  return 0;
}

int orc_image_shape(orc_ctx *g, int32_t group, int32_t which, int64_t dims[6], int32_t *ndim) {
  if (group < 0 || group >= (int)g->peeled.image.size()) return fail(g, "no such image group");
  const Image &im = g->peeled.image[group];
  if (which == 0) {
    if (!im.c.compute_sed) return fail(g, "group has no SED");
    int64_t d[5] = {im.n_stokes, im.n_orig, im.c.n_view, im.c.n_ap, im.n_nu};
    for (int i = 0; i < 5; i++) dims[i] = d[i];
    *ndim = 5;
  } else {
    if (!im.c.compute_image) return fail(g, "group has no image");
    int64_t d[6] = {im.n_stokes, im.n_orig, im.c.n_view, im.c.n_y, im.c.n_x, im.n_nu};
    for (int i = 0; i < 6; i++) dims[i] = d[i];
    *ndim = 6;
  }
  return 0;
}

static int get_written(orc_ctx *g, int32_t group, bool sed, double *out, double *unc) {
  if (group < 0 || group >= (int)g->peeled.image.size()) return fail(g, "no such image group");
  std::vector<double> a, u;
  image_written(g->peeled.image[group], sed, a, u);
  std::copy(a.begin(), a.end(), out);
  if (unc && !u.empty()) std::copy(u.begin(), u.end(), unc);
  return 0;
}
int orc_get_sed(orc_ctx *g, int32_t group, double *sed, double *unc) { return get_written(g, group, true, sed, unc); }
int orc_get_image(orc_ctx *g, int32_t group, double *img, double *unc) { return get_written(g, group, false, img, unc); }

double orc_test_random(orc_ctx *g) { return g->rng.random(); }
int orc_test_locate(const double *xx, int n, double x) { return locate(xx, n, x); }
double orc_test_interp1d_loglog(const double *x, const double *y, int n, double xv) {
  try {
    return interp1d_loglog(x, y, n, xv);
  } catch (OracleError &) {
    return std::nan("");
  }
}
double orc_test_planck(orc_ctx *g, double T) { return g->rng.random_planck_frequency(T); }
uint64_t orc_rng_draws(orc_ctx *g) { return g->rng.n_draws; }

}  // extern "C"
