// tables.h -- host-side construction of the sampling tables the photon kernels read.
//
// Takes the raw columns of a Hyperion dust file / source spectrum (the hyp_dust_tables and
// hyp_source structs of include/hyperion_b200.h) and produces flat fp64 arrays laid out for
// the device: everything a kernel needs for one dust type lives in ONE contiguous buffer so
// that it is uploaded with a single copy and stays resident in L2.
//
// What is computed follows dust_setup (reference src/dust/dust_type_4elem.f90:78-293) and
// set_pdf / find_cdf (fortranlib/src/type_pdf.f90:233-311): phase-matrix normalisation,
// cumulative phase functions, and for each emissivity state the normalised power-law PDF, its
// CDF and the per-interval exponents used for exact inverse-CDF sampling.
#pragma once

#include <cmath>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/hyperion_b200.h"

namespace hyp {

// Offsets (in doubles) into a dust type's device buffer.
struct DustLayout {
  int32_t n_nu, n_mu, n_e, n_jnu, n_enu;
  int32_t zero_p2, sublimation_mode, pad0;
  double sublimation_specific_energy;
  double nu_min, nu_max, mu_min, mu_max;
  double e_min, e_max;        // range of the mean-opacity table (check_energy_abs clamp)
  double jvar_min, jvar_max;  // range of the emissivity variable
  // per-frequency optical properties (length n_nu)
  int64_t o_nu, o_lognu, o_logchi, o_logalb;
  // scattering matrix, each [n_nu][n_mu]
  int64_t o_mu, o_P1, o_P2, o_P3, o_P4, o_C1, o_C2;
  // mean opacities (length n_e), log10 of each
  int64_t o_loge, o_logchi_ross, o_logchi_invp, o_logkap_planck;
  // emissivities
  int64_t o_logjvar;           // [n_jnu] log10 of the emissivity variable
  int64_t o_jvar;              // [n_jnu]
  int64_t o_enu;               // [n_enu] frequencies of the emissivity PDFs
  int64_t o_ecdf;              // [n_jnu][n_enu]
  int64_t o_einvb;             // [n_jnu][n_enu-1]  1/(b+1)
  int64_t o_erm1;              // [n_jnu][n_enu-1]  r-1
  int64_t total;
};

// Samplers of b_nu = j_nu / kappa_nu for every emissivity state (dust_type_4elem.f90:286-291), used by the
// modified random walk (dust_sample_b_nu, :400-419).  Layout: cdf [n_jnu][n_enu], 1/(b+1) and r-1
// [n_jnu][n_enu-1] each, in a buffer of its own (only models that enable the MRW pay for it).
struct DustMrwLayout {
  int64_t o_bcdf, o_binvb, o_brm1, total;
};

struct SpectrumLayout {
  int32_t n;
  int32_t pad;
  int64_t o_x, o_cdf, o_invb, o_rm1;  // same sampling form as an emissivity state
  int64_t total;
};

namespace detail {

inline double seg_loglog(double x1, double y1, double x2, double y2) {
  // exact integral of the power law through (x1,y1),(x2,y2)
  if (x1 == x2 || y1 == 0.0 || y2 == 0.0) return 0.0;
  const double b = std::log10(y1 / y2) / std::log10(x1 / x2);
  if (std::fabs(b + 1.0) < 1e-10) return x1 * y1 * std::log(x2 / x1);
  return y1 * (x2 * std::pow(x2 / x1, b) - x1) / (b + 1.0);
}

inline double seg_linlog(double x1, double y1, double x2, double y2) {
  // integral of an exponential through the two points (linear x, log y)
  if (x1 == x2) return 0.0;
  if (y1 == y2) return y1 * (x2 - x1);
  return (y2 - y1) * (x2 - x1) / std::log(10.0) / std::log10(y2 / y1);
}

// Builds cdf / 1/(b+1) / r-1 for a log-log PDF given on x.  Returns false if the PDF has no weight.
inline void build_powerlaw_sampler(const double *x, const double *y, int n, double *cdf, double *invb,
                                   double *rm1) {
  std::vector<double> seg(n > 1 ? n - 1 : 0);
  double total = 0.0;
  for (int i = 0; i + 1 < n; ++i) {
    if (!(x[i + 1] > x[i])) throw std::runtime_error("[check_pdf] PDF x array is not sorted");
    seg[i] = seg_loglog(x[i], y[i], x[i + 1], y[i + 1]);
    total += seg[i];
  }
  if (!(total > 0.0)) throw std::runtime_error("PDF has zero integral");
  // normalise first (as set_pdf does), then accumulate, then renormalise the cdf by its last value
  std::vector<double> p(n);
  for (int i = 0; i < n; ++i) p[i] = y[i] / total;
  cdf[0] = 0.0;
  for (int i = 0; i + 1 < n; ++i) cdf[i + 1] = cdf[i] + seg_loglog(x[i], p[i], x[i + 1], p[i + 1]);
  const double last = cdf[n - 1];
  for (int i = 0; i < n; ++i) cdf[i] /= last;
  for (int i = 0; i + 1 < n; ++i) {
    const double b = std::log10(p[i] / p[i + 1]) / std::log10(x[i] / x[i + 1]);
    const double r = std::pow(x[i + 1] / x[i], b + 1.0);
    invb[i] = 1.0 / (b + 1.0);
    rm1[i] = r - 1.0;
  }
}

inline double safe_log10(double v) {
  return v > 0.0 ? std::log10(v) : -std::numeric_limits<double>::infinity();
}

}  // namespace detail

inline void check_finite(const double *a, size_t n, const char *what) {
  for (size_t i = 0; i < n; ++i)
    if (std::isnan(a[i])) throw std::runtime_error(std::string(what) + " array contains NaN values");
}

// Build the flat table for one dust type.
inline void build_dust(const hyp_dust_tables &t, DustLayout &L, std::vector<double> &buf) {
  if (t.n_nu < 2 || t.n_mu < 2 || t.n_e < 1 || t.n_jnu < 2 || t.n_emiss_nu < 2)
    throw std::runtime_error("dust tables are too small");
  const int n_nu = t.n_nu, n_mu = t.n_mu, n_e = t.n_e, n_jnu = t.n_jnu, n_enu = t.n_emiss_nu;
  const size_t np = (size_t)n_nu * n_mu;
  check_finite(t.nu, n_nu, "nu");
  check_finite(t.albedo, n_nu, "albedo_nu");
  check_finite(t.chi, n_nu, "chi_nu");
  check_finite(t.P1, np, "P1 matrix");
  check_finite(t.P2, np, "P2 matrix");
  check_finite(t.P3, np, "P3 matrix");
  check_finite(t.P4, np, "P4 matrix");
  check_finite(t.mu, n_mu, "mu");
  check_finite(t.specific_energy, n_e, "specific_energy");
  check_finite(t.emiss_nu, n_enu, "emiss_nu");
  check_finite(t.emiss_jnu, (size_t)n_enu * n_jnu, "emiss_jnu");
  check_finite(t.jnu_var, n_jnu, "emissivity variable");
  for (int i = 1; i < n_e; ++i)
    if (t.specific_energy[i] < t.specific_energy[i - 1])
      throw std::runtime_error("energy per unit mass is not monotonically increasing");

  L = DustLayout();
  L.n_nu = n_nu;
  L.n_mu = n_mu;
  L.n_e = n_e;
  L.n_jnu = n_jnu;
  L.n_enu = n_enu;
  L.sublimation_mode = t.sublimation_mode;
  L.sublimation_specific_energy = t.sublimation_specific_energy;
  L.nu_min = t.nu[0];
  L.nu_max = t.nu[n_nu - 1];
  L.mu_min = t.mu[0];
  L.mu_max = t.mu[n_mu - 1];
  L.e_min = t.specific_energy[0];
  L.e_max = t.specific_energy[n_e - 1];
  L.jvar_min = t.jnu_var[0];
  L.jvar_max = t.jnu_var[n_jnu - 1];

  int64_t off = 0;
  auto take = [&](int64_t n) {
    int64_t o = off;
    off += (n + 1) & ~int64_t(1);  // keep 16-byte alignment of every sub-array
    return o;
  };
  L.o_nu = take(n_nu);
  L.o_lognu = take(n_nu);
  L.o_logchi = take(n_nu);
  L.o_logalb = take(n_nu);
  L.o_mu = take(n_mu);
  L.o_P1 = take(np);
  L.o_P2 = take(np);
  L.o_P3 = take(np);
  L.o_P4 = take(np);
  L.o_C1 = take(np);
  L.o_C2 = take(np);
  L.o_loge = take(n_e);
  L.o_logchi_ross = take(n_e);
  L.o_logchi_invp = take(n_e);
  L.o_logkap_planck = take(n_e);
  L.o_logjvar = take(n_jnu);
  L.o_jvar = take(n_jnu);
  L.o_enu = take(n_enu);
  L.o_ecdf = take((int64_t)n_jnu * n_enu);
  L.o_einvb = take((int64_t)n_jnu * (n_enu - 1));
  L.o_erm1 = take((int64_t)n_jnu * (n_enu - 1));
  L.total = off;
  buf.assign((size_t)off, 0.0);
  double *B = buf.data();

  for (int j = 0; j < n_nu; ++j) {
    B[L.o_nu + j] = t.nu[j];
    B[L.o_lognu + j] = std::log10(t.nu[j]);
    B[L.o_logchi + j] = detail::safe_log10(t.chi[j]);
    B[L.o_logalb + j] = detail::safe_log10(t.albedo[j]);
  }
  for (int i = 0; i < n_mu; ++i) B[L.o_mu + i] = t.mu[i];

  // phase matrix: normalise each frequency row so that the mu-integral of P1 equals the mu range
  const double dmu = L.mu_max - L.mu_min;
  bool zero_p2 = true;
  for (size_t k = 0; k < np; ++k)
    if (t.P2[k] != 0.0) zero_p2 = false;
  L.zero_p2 = zero_p2 ? 1 : 0;
  const double *Pin[4] = {t.P1, t.P2, t.P3, t.P4};
  const int64_t Pout[4] = {L.o_P1, L.o_P2, L.o_P3, L.o_P4};
  for (int j = 0; j < n_nu; ++j) {
    const double *p1 = t.P1 + (size_t)j * n_mu;
    double norm = 0.0;
    for (int i = 0; i + 1 < n_mu; ++i) norm += detail::seg_linlog(t.mu[i], p1[i], t.mu[i + 1], p1[i + 1]);
    if (norm == 0.0) throw std::runtime_error("P1 matrix normalization is zero");
    for (int q = 0; q < 4; ++q)
      for (int i = 0; i < n_mu; ++i) B[Pout[q] + (size_t)j * n_mu + i] = Pin[q][(size_t)j * n_mu + i] / norm * dmu;
    // cumulative (trapezoid) of the normalised P1 and P2, each scaled by its last element
    const int64_t Cout[2] = {L.o_C1, L.o_C2};
    for (int q = 0; q < 2; ++q) {
      const double *P = B + Pout[q] + (size_t)j * n_mu;
      double *Cq = B + Cout[q] + (size_t)j * n_mu;
      Cq[0] = 0.0;
      bool any = false;
      for (int i = 0; i + 1 < n_mu; ++i) {
        Cq[i + 1] = Cq[i] + 0.5 * (P[i] + P[i + 1]) * (t.mu[i + 1] - t.mu[i]);
        if (Cq[i + 1] != 0.0) any = true;
      }
      if (any) {
        const double last = Cq[n_mu - 1];
        for (int i = 0; i < n_mu; ++i) Cq[i] /= last;
      }
    }
  }

  for (int i = 0; i < n_e; ++i) {
    B[L.o_loge + i] = detail::safe_log10(t.specific_energy[i]);
    B[L.o_logchi_ross + i] = detail::safe_log10(t.chi_rosseland[i]);
    B[L.o_logchi_invp + i] = detail::safe_log10(t.chi_inv_planck[i]);
    B[L.o_logkap_planck + i] = detail::safe_log10(t.kappa_planck[i]);
  }
  for (int i = 0; i < n_jnu; ++i) {
    B[L.o_jvar + i] = t.jnu_var[i];
    B[L.o_logjvar + i] = std::log10(t.jnu_var[i]);
  }
  for (int k = 0; k < n_enu; ++k) B[L.o_enu + k] = t.emiss_nu[k];
  std::vector<double> col(n_enu);
  for (int s = 0; s < n_jnu; ++s) {
    for (int k = 0; k < n_enu; ++k) col[k] = t.emiss_jnu[(size_t)k * n_jnu + s];
    detail::build_powerlaw_sampler(t.emiss_nu, col.data(), n_enu, B + L.o_ecdf + (size_t)s * n_enu,
                                   B + L.o_einvb + (size_t)s * (n_enu - 1),
                                   B + L.o_erm1 + (size_t)s * (n_enu - 1));
  }
}

inline void build_spectrum(const double *nu, const double *fnu, int n, SpectrumLayout &L,
                           std::vector<double> &buf) {
  if (n < 2) throw std::runtime_error("spectrum needs at least two points");
  for (int i = 0; i + 1 < n; ++i)
    if (nu[i + 1] < nu[i]) throw std::runtime_error("spectrum frequency should be monotonically increasing");
  L = SpectrumLayout();
  L.n = n;
  int64_t off = 0;
  auto take = [&](int64_t m) {
    int64_t o = off;
    off += (m + 1) & ~int64_t(1);
    return o;
  };
  L.o_x = take(n);
  L.o_cdf = take(n);
  L.o_invb = take(n - 1);
  L.o_rm1 = take(n - 1);
  L.total = off;
  buf.assign((size_t)off, 0.0);
  for (int i = 0; i < n; ++i) buf[L.o_x + i] = nu[i];
  detail::build_powerlaw_sampler(nu, fnu, n, buf.data() + L.o_cdf, buf.data() + L.o_invb, buf.data() + L.o_rm1);
}


// ---------------------------------------------------------------------------------------------
// spectra on an image group's frequency grid, used by the raytracing iteration
// (get_spectrum_binned src/sources/source_type.f90:1118-1165, get_j_nu_binned / get_chi_nu_binned
//  src/dust/dust_type_4elem.f90:722-741,793-811, on top of integral_loglog with limits,
//  fortranlib/src/lib_array.f90:362-366,450-526)
// ---------------------------------------------------------------------------------------------
namespace detail {

// value of the piecewise power law through (x, y) at xv (x[0] <= xv <= x[n-1]); j = interval of xv
inline double powerlaw_at(const double *x, const double *y, int j, double xv) {
  if (y[j] == 0.0 || y[j + 1] == 0.0) return 0.0;
  const double f = (std::log10(xv) - std::log10(x[j])) / (std::log10(x[j + 1]) - std::log10(x[j]));
  return std::pow(10.0, std::log10(y[j]) + f * (std::log10(y[j + 1]) - std::log10(y[j])));
}

// interval j in [0, n-2] with x[j] <= v < x[j+1] (top edge belongs to the last interval)
inline int interval_of(const double *x, int n, double v) {
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) / 2;
    if (x[mid] <= v) lo = mid; else hi = mid;
  }
  return lo;
}

inline double seg_loglog_f32tol(double x1, double y1, double x2, double y2) {
  // as seg_loglog, with the reference's single-precision tolerance on b = -1 (lib_array.f90:562-578)
  if (x1 == x2 || y1 == 0.0 || y2 == 0.0) return 0.0;
  const double b = std::log10(y1 / y2) / std::log10(x1 / x2);
  if (std::fabs(b + 1.0) < (double)1e-10f) return x1 * y1 * std::log(x2 / x1);
  return y1 * (x2 * std::pow(x2 / x1, b) - x1) / (b + 1.0);
}

// integral of the piecewise power law over [a, b] clipped to the table
inline double integral_loglog_range(const double *x, const double *y, int n, double a, double b) {
  if (a > x[n - 1] || b < x[0]) return 0.0;
  // first node strictly above the lower limit, last node strictly below the upper limit
  int k_lo, k_hi;
  double xa, ya, xb, yb;
  if (a > x[0]) {
    const int j = (a == x[n - 1]) ? n - 2 : interval_of(x, n, a);
    k_lo = j + 1;
    xa = a;
    ya = powerlaw_at(x, y, j, a);
  } else {
    k_lo = 0;
    xa = x[0];
    ya = y[0];
  }
  if (b < x[n - 1]) {
    const int j = interval_of(x, n, b);
    k_hi = j;
    xb = b;
    yb = powerlaw_at(x, y, j, b);
  } else {
    k_hi = n - 1;
    xb = x[n - 1];
    yb = y[n - 1];
  }
  if (k_hi < k_lo) return seg_loglog_f32tol(xa, ya, xb, yb);  // both limits inside one interval
  double sum = 0.0;
  for (int k = k_lo; k < k_hi; ++k) sum += seg_loglog_f32tol(x[k], y[k], x[k + 1], y[k + 1]);
  sum += seg_loglog_f32tol(xa, ya, x[k_lo], y[k_lo]);
  sum += seg_loglog_f32tol(x[k_hi], y[k_hi], xb, yb);
  return sum;
}

inline double integral_loglog_all(const double *x, const double *y, int n) {
  double sum = 0.0;
  for (int k = 0; k + 1 < n; ++k) sum += seg_loglog_f32tol(x[k], y[k], x[k + 1], y[k + 1]);
  return sum;
}

}  // namespace detail

// edges of frequency bin inu (0-based) of an image group with n_nu log-spaced bins
inline void image_bin_edges(double log10_nu_min, double log10_nu_max, int n_nu, int inu, double &lo, double &hi) {
  lo = std::pow(10.0, log10_nu_min + (log10_nu_max - log10_nu_min) * (double)inu / (double)n_nu);
  hi = std::pow(10.0, log10_nu_min + (log10_nu_max - log10_nu_min) * (double)(inu + 1) / (double)n_nu);
}

// fraction of a spectrum (nu, fnu) falling in each bin
inline void binned_fraction(const double *nu, const double *fnu, int n, double l0, double l1, int n_nu, double *out) {
  const double tot = detail::integral_loglog_all(nu, fnu, n);
  for (int i = 0; i < n_nu; ++i) {
    double lo, hi;
    image_bin_edges(l0, l1, n_nu, i, lo, hi);
    out[i] = detail::integral_loglog_range(nu, fnu, n, lo, hi) / tot;
  }
}

// normalized_B_nu (source_type.f90:1088-1096) tabulated as get_spectrum_binned does for a blackbody
inline void blackbody_table(double T, std::vector<double> &nu, std::vector<double> &fnu) {
  const double h_cgs = 6.6260689633e-27, c_cgs = 2.99792458e10, k_cgs = 1.380650424e-16, stef_boltz = 5.670400e-5;
  const double pi = 3.14159265358979323846264338327950288419;
  const double a = 2.0 * h_cgs / c_cgs / c_cgs / stef_boltz * pi, b = h_cgs / k_cgs;
  const double lmin = std::log10(3.e9), lmax = std::log10(3.e16);
  const int n = (int)std::ceil((lmax - lmin) * 100000);
  nu.resize(n);
  fnu.resize(n);
  const double T4 = T * T * T * T;
  for (int i = 0; i < n; ++i) {
    nu[i] = std::pow(10.0, (double)i / (double)(n - 1) * (lmax - lmin) + lmin);
    fnu[i] = a * nu[i] * nu[i] * nu[i] / (std::exp(b * nu[i] / T) - 1.0) / T4;
  }
}

// normalized_B_nu (source_type.f90:1088-1096) at one frequency
inline double normalized_B_nu(double nu, double T) {
  const double h_cgs = 6.6260689633e-27, c_cgs = 2.99792458e10, k_cgs = 1.380650424e-16, stef_boltz = 5.670400e-5;
  const double pi = 3.14159265358979323846264338327950288419;
  const double a = 2.0 * h_cgs / c_cgs / c_cgs / stef_boltz * pi, b = h_cgs / k_cgs;
  const double T4 = T * T * T * T;
  return a * nu * nu * nu / (std::exp(b * nu / T) - 1.0) / T4;
}

// interp1d_loglog(x, y, xval, bounds_error=.false., fill_value=0) (lib_array.f90:588-614): the piecewise power
// law through an ascending table, zero outside it and where either node is zero
inline double interp_loglog_fill0(const double *x, const double *y, int n, double xv) {
  if (!(xv >= x[0] && xv <= x[n - 1])) return 0.0;
  if (xv == x[n - 1]) return y[n - 1];
  if (xv == x[0]) return y[0];
  return detail::powerlaw_at(x, y, detail::interval_of(x, n, xv), xv);
}

// interpolate_pdf (type_pdf.f90:402-420) of a log-log PDF: the table normalised to unit integral, at xv
inline double pdf_loglog_at(const double *x, const double *y, int n, double xv) {
  return interp_loglog_fill0(x, y, n, xv) / detail::integral_loglog_all(x, y, n);
}

// mean extinction in each bin
inline void binned_chi(const double *nu, const double *chi, int n, double l0, double l1, int n_nu, double *out) {
  for (int i = 0; i < n_nu; ++i) {
    double lo, hi;
    image_bin_edges(l0, l1, n_nu, i, lo, hi);
    out[i] = detail::integral_loglog_range(nu, chi, n, lo, hi) / (hi - lo);
  }
}

inline void build_dust_mrw(const double *nu, const double *chi, const double *albedo, int n_nu, const double *emiss_nu,
                           const double *emiss_jnu, int n_enu, int n_jnu, DustMrwLayout &L, std::vector<double> &buf) {
  L.o_bcdf = 0;
  L.o_binvb = (int64_t)n_jnu * n_enu;
  L.o_brm1 = L.o_binvb + (int64_t)n_jnu * (n_enu - 1);
  L.total = L.o_brm1 + (int64_t)n_jnu * (n_enu - 1);
  buf.assign((size_t)L.total, 0.0);
  // kappa_nu on the emissivity grid, log-log interpolation of chi (1 - albedo)
  std::vector<double> kap(n_enu), col(n_enu);
  for (int k = 0; k < n_enu; ++k) {
    const double x = emiss_nu[k];
    if (x < nu[0] || x > nu[n_nu - 1]) throw std::runtime_error("Interpolation out of bounds");
    int j = detail::interval_of(nu, n_nu, x);
    const double k0 = chi[j] * (1.0 - albedo[j]), k1 = chi[j + 1] * (1.0 - albedo[j + 1]);
    if (k0 == 0.0 || k1 == 0.0) {
      kap[k] = 0.0;
    } else {
      const double f = (std::log10(x) - std::log10(nu[j])) / (std::log10(nu[j + 1]) - std::log10(nu[j]));
      kap[k] = std::pow(10.0, std::log10(k0) + f * (std::log10(k1) - std::log10(k0)));
    }
  }
  for (int s = 0; s < n_jnu; ++s) {
    for (int k = 0; k < n_enu; ++k) col[k] = emiss_jnu[(size_t)k * n_jnu + s] / kap[k];
    detail::build_powerlaw_sampler(emiss_nu, col.data(), n_enu, buf.data() + L.o_bcdf + (size_t)s * n_enu,
                                   buf.data() + L.o_binvb + (size_t)s * (n_enu - 1),
                                   buf.data() + L.o_brm1 + (size_t)s * (n_enu - 1));
  }
}

// P(y) = 2 sum_n (-1)^(n+1) y^(n^2) tabulated at 100 points (initialize_cumulative, grid_mrw_3d.f90:158-196)
inline void build_mrw_cumulative(double *xcdf, double *ycdf) {
  const int ncdf = 100;
  for (int i = 1; i <= ncdf; ++i) {
    xcdf[i - 1] = (double)(i - 1) / (double)(ncdf - 1);
    double y = 0.0;
    if (i == ncdf) {
      y = 0.5;
    } else {
      for (long long j = 1;; ++j) {
        const double term = std::pow(xcdf[i - 1], (double)(j * j));
        if (term == 0.0) break;
        y += (j % 2 == 0) ? -term : term;
      }
    }
    ycdf[i - 1] = y * 2.0;
  }
}


}  // namespace hyp
