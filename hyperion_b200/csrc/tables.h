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
  int64_t o_loge, o_logchi_ross, o_logchi_invp;
  // emissivities
  int64_t o_logjvar;           // [n_jnu] log10 of the emissivity variable
  int64_t o_jvar;              // [n_jnu]
  int64_t o_enu;               // [n_enu] frequencies of the emissivity PDFs
  int64_t o_ecdf;              // [n_jnu][n_enu]
  int64_t o_einvb;             // [n_jnu][n_enu-1]  1/(b+1)
  int64_t o_erm1;              // [n_jnu][n_enu-1]  r-1
  int64_t total;
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

}  // namespace hyp
