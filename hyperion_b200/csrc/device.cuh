// device.cuh -- device-side building blocks of the photon-packet loop (sm_100a).
//
// Everything here is fp64 IEEE arithmetic (no fast-math): the traversal relies on +-Inf for
// rays parallel to a wall, exactly as the reference relies on IEEE semantics in find_wall
// (src/grid/grid_geometry_cartesian_3d.f90:443-468).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "tables.h"

namespace hyp {

constexpr int MAX_DUST = 4;      // dust types per model handled by the kernels
constexpr int MAX_SOURCES = 4095;   // 12 bits of the packet tag (imaging.cuh)

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG.  key = run seed, counter = (photon id lo, hi, block index, iteration):
// every packet owns a private stream, so results do not depend on how packets are spread over
// threads or GPUs.  Replaces the single sequential Marsaglia-Tsang stream of the reference
// (fortranlib/src/lib_random.f90:172-197).
// ---------------------------------------------------------------------------------------------
struct Rng {
  uint32_t k0, k1;        // key
  uint32_t c0, c1, c3;    // fixed counter words
  uint32_t blk;           // next block index
  double spare;
  bool has_spare;

  __device__ __forceinline__ void init(uint64_t seed, uint64_t photon_id, uint32_t iteration) {
    k0 = (uint32_t)seed;
    k1 = (uint32_t)(seed >> 32);
    c0 = (uint32_t)photon_id;
    c1 = (uint32_t)(photon_id >> 32);
    c3 = iteration;
    blk = 0;
    has_spare = false;
    spare = 0.0;
  }

  __device__ __forceinline__ void block(uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t x0 = c0, x1 = c1, x2 = blk, x3 = c3, a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0 = __umulhi(M0, x0), lo0 = M0 * x0;
      uint32_t hi1 = __umulhi(M1, x2), lo1 = M1 * x2;
      uint32_t y0 = hi1 ^ x1 ^ a, y1 = lo1, y2 = hi0 ^ x3 ^ b, y3 = lo0;
      x0 = y0; x1 = y1; x2 = y2; x3 = y3;
      a += W0; b += W1;
    }
    out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
    ++blk;
  }

  // uniform in [0,1) with 53 random bits
  __device__ __forceinline__ double next() {
    if (has_spare) {
      has_spare = false;
      return spare;
    }
    uint32_t w[4];
    block(w);
    const double s = 1.0 / 9007199254740992.0;
    uint64_t u = ((uint64_t)w[0] << 32 | w[1]) >> 11;
    uint64_t v = ((uint64_t)w[2] << 32 | w[3]) >> 11;
    spare = (double)v * s;
    has_spare = true;
    return (double)u * s;
  }
};

// ---------------------------------------------------------------------------------------------
// table look-ups
// ---------------------------------------------------------------------------------------------

// Largest j in [0, n-2] with x[j] <= v (x ascending); callers guarantee x[0] <= v <= x[n-1].
__device__ __forceinline__ int lower_interval(const double *__restrict__ x, int n, double v) {
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(x + mid) <= v) lo = mid; else hi = mid;
  }
  return lo;
}

struct DustDev {
  DustLayout L;
  const double *B;   // device buffer
  DustMrwLayout Lm;  // b_nu samplers (modified random walk), nullptr unless the MRW is enabled
  const double *Bm;
};

struct SpectrumDev {
  SpectrumLayout L;
  const double *B;
};

struct SourceDev {
  int32_t type, freq_type;
  double x, y, z, radius, temperature;
  int32_t limb, spectrum;  // index into spectra
  int32_t peeloff, pad;    // whether the source is peeled off (source_emit_peeloff, source_type.f90:513-537)
  double pdf;              // normalised luminosity (for even sampling weights)
  double cdf;              // cumulative normalised luminosity
  double box[6];           // extern_box: xmin, xmax, ymin, ymax, zmin, zmax
  double face_cdf[6];      // extern_box: cumulative face areas (set_pdf, source_type.f90:229)
  double dir_cost, dir_sint, dir_cosp, dir_sinp;  // plane_parallel: direction of travel (angle3d_deg(theta, phi))
  int64_t coll_off;        // point_collection: first entry in ModelDev::coll_xyz / coll_cdf
  int64_t coll_n;
  int64_t map_off;         // map: first entry of this source's cumulative luminosity map in ModelDev::map_cdf
  int32_t spot_off, n_spots;  // sphere with spots: entries [spot_off, spot_off + n_spots] of ModelDev::spots (the last
                              // entry is the star itself and only carries the cdf)
};

// One spot of a spherical source (source_type.f90:27-33,150-188)
struct SpotDev {
  double cdf;                            // cumulative luminosity over the spots and, last, the star
  double a_cost, a_sint, a_cosp, a_sinp; // angle3d_deg(longitude, latitude)
  double cost;                           // cos(radius)
  double temperature;
  int32_t freq_type, spectrum;
};

// sample_pdf_discrete_dp (type_pdf.f90:313-337): 0-based index of the first entry with cdf >= xi
__device__ __forceinline__ int64_t sample_discrete(const double *__restrict__ cdf, int64_t n, double xi) {
  if (xi <= cdf[0]) return 0;
  if (xi >= cdf[n - 1]) return n - 1;
  int64_t jmin = 1, jmax = n;
  for (;;) {
    const int64_t j = (jmax + jmin) / 2;
    if (xi > cdf[j - 1]) jmin = j; else jmax = j;
    if (jmax == jmin + 1) break;
  }
  return jmax - 1;
}

// log-log interpolation on a pre-logged table (update_optconsts, src/dust/dust.f90:64-79 +
// interp1d_loglog, fortranlib/src/lib_array.f90:605-614): returns 0 where either node is 0.
__device__ __forceinline__ double loglog_at(const double *__restrict__ logy, int j, double frac) {
  double y1 = __ldg(logy + j), y2 = __ldg(logy + j + 1);
  if (isinf(y1) || isinf(y2)) return 0.0;
  return pow(10.0, y1 + frac * (y2 - y1));
}

// Inverse-CDF sample of a piecewise power-law PDF (sample_pdf_cont_dp, type_pdf.f90:339-381).
__device__ __forceinline__ double sample_powerlaw(const double *__restrict__ x, const double *__restrict__ cdf,
                                                  const double *__restrict__ invb, const double *__restrict__ rm1,
                                                  int n, double xi) {
  if (xi <= __ldg(cdf)) return __ldg(x);
  if (xi >= __ldg(cdf + n - 1)) return __ldg(x + n - 1);
  int i = lower_interval(cdf, n, xi);
  double c0 = __ldg(cdf + i), c1 = __ldg(cdf + i + 1);
  double f = (xi - c0) / (c1 - c0);
  return pow(f * __ldg(rm1 + i) + 1.0, __ldg(invb + i)) * __ldg(x + i);
}

// The same draw xi from two samplers that share x and lie row by row in one table (the emissivity states jid and
// jid + 1 of dust_sample_j_nu, dust_type_4elem.f90:379-398): the two bisections advance in lock-step, so that their
// dependent loads (the latency of a re-emission) are in flight two at a time instead of one after the other.
__device__ __forceinline__ void sample_powerlaw_pair(const double *__restrict__ x, const double *__restrict__ cdfA,
                                                     const double *__restrict__ invbA, const double *__restrict__ rm1A,
                                                     const double *__restrict__ cdfB, const double *__restrict__ invbB,
                                                     const double *__restrict__ rm1B, int n, double xi, double &outA,
                                                     double &outB) {
  const double a0 = __ldg(cdfA), a1 = __ldg(cdfA + n - 1), b0 = __ldg(cdfB), b1 = __ldg(cdfB + n - 1);
  int loA = 0, hiA = n - 1, loB = 0, hiB = n - 1;
  while (hiA - loA > 1 || hiB - loB > 1) {
    const int mA = (loA + hiA) >> 1, mB = (loB + hiB) >> 1;
    const double a = __ldg(cdfA + mA), b = __ldg(cdfB + mB);
    if (hiA - loA > 1) {
      if (a <= xi) loA = mA; else hiA = mA;
    }
    if (hiB - loB > 1) {
      if (b <= xi) loB = mB; else hiB = mB;
    }
  }
  const double cA0 = __ldg(cdfA + loA), cA1 = __ldg(cdfA + loA + 1), cB0 = __ldg(cdfB + loB), cB1 = __ldg(cdfB + loB + 1);
  const double rA = __ldg(rm1A + loA), iA = __ldg(invbA + loA), xA = __ldg(x + loA);
  const double rB = __ldg(rm1B + loB), iB = __ldg(invbB + loB), xB = __ldg(x + loB);
  const double fA = (xi - cA0) / (cA1 - cA0), fB = (xi - cB0) / (cB1 - cB0);
  outA = xi <= a0 ? __ldg(x) : (xi >= a1 ? __ldg(x + n - 1) : pow(fA * rA + 1.0, iA) * xA);
  outB = xi <= b0 ? __ldg(x) : (xi >= b1 ? __ldg(x + n - 1) : pow(fB * rB + 1.0, iB) * xB);
}

// Planck-law frequency sampling (random_planck_frequency_dp, lib_random.f90:297-347).
__device__ __forceinline__ double sample_planck(Rng &rng, double T) {
  const double k = 1.3806503e-23, h = 6.626068e-34;
  double r;
  do {
    r = rng.next() * rng.next() * rng.next() * rng.next();
  } while (!(r > 0.0));
  double x = -log(r);
  double a = 1.0, y = 1.0, z = 1.0;
  double r1 = rng.next();
  while (!(1.08232 * r1 <= a)) {
    y += 1.0;
    z = 1.0 / y;
    a += z * z * z * z;
  }
  x = x * z;
  return x * k * T / h;
}

// ---------------------------------------------------------------------------------------------
// direction algebra on (cos t, sin t, cos p, sin p) -- fortranlib/src/type_angle3d.f90
// ---------------------------------------------------------------------------------------------
struct Angle {
  double cost, sint, cosp, sinp;
};

__device__ __forceinline__ Angle random_sphere_angle(Rng &rng) {
  const double TWOPI = 6.283185307179586476925286766559;
  Angle a;
  a.cost = -1.0 + 2.0 * rng.next();
  double phi = TWOPI * rng.next();
  a.sint = sqrt(1.0 - a.cost * a.cost);
  sincos(phi, &a.sinp, &a.cosp);
  return a;
}

__device__ __forceinline__ double sin2cos(double x) { return (x * x < 1.0) ? sqrt(1.0 - x * x) : 0.0; }

// rotate_angle3d_dp (type_angle3d.f90:160-281): add the local (scattering) angle to a_coord.
__device__ inline Angle rotate_angle(const Angle &l, const Angle &c) {
  Angle f;
  if (fabs(c.sint) < 1.e-10) {
    f = l;
    if (c.cost > 0.0) {
      f.cosp = l.cosp * c.cosp + l.sinp * c.sinp;
      f.sinp = l.cosp * c.sinp - l.sinp * c.cosp;
    } else {
      f.cost = -l.cost;
      f.cosp = l.cosp * c.cosp - l.sinp * c.sinp;
      f.sinp = l.cosp * c.sinp + l.sinp * c.cosp;
    }
    return f;
  }
  const double cos_a = c.cost, sin_a = c.sint, cos_b = l.cost, sin_b = l.sint;
  const double cos_C = l.cosp, sin_C = fabs(l.sinp);
  bool same_sign;
  double delta;
  if (fabs(sin_a) > fabs(cos_a)) {
    same_sign = (sin_a > 0.0) == (sin_b > 0.0);
    delta = cos_b - cos_a;
  } else {
    same_sign = (cos_a > 0.0) == (cos_b > 0.0);
    delta = sin_b - sin_a;
  }
  double cos_c, sin_c;
  if (same_sign && fabs(delta) < 1.e-5 && sin_C < 1.e-5 && cos_C > 0.0) {
    double q = (fabs(sin_a) > fabs(cos_a)) ? cos_a / sin_a : sin_a / cos_a;
    sin_c = sqrt(delta * delta * (1.0 + q * q) + sin_a * sin_b * sin_C * sin_C);
    cos_c = sin2cos(sin_c);
  } else {
    cos_c = cos_a * cos_b + sin_a * sin_b * cos_C;
    sin_c = sin2cos(cos_c);
  }
  if (fabs(sin_c) < 1.e-10) {
    f.cost = cos_c > 0.0 ? 1.0 : -1.0;
    f.sint = 0.0;
    f.cosp = 1.0;
    f.sinp = 0.0;
    return f;
  }
  const double cos_B = (cos_b - cos_a * cos_c) / (sin_a * sin_c);
  const double sin_B = sin_C * sin_b / sin_c;
  f.cost = cos_c;
  f.sint = sin_c;
  if (l.sinp < 0.0) {
    f.cosp = cos_B * c.cosp + sin_B * c.sinp;
    f.sinp = cos_B * c.sinp - sin_B * c.cosp;
  } else {
    f.cosp = cos_B * c.cosp - sin_B * c.sinp;
    f.sinp = cos_B * c.sinp + sin_B * c.cosp;
  }
  return f;
}

struct Stokes {
  double I, Q, U, V;
};

// scatter_stokes (src/dust/dust_type_4elem.f90:603-690): S = L(pi - i2) R L(-i1) S'
__device__ inline void scatter_stokes(Stokes &s, const Angle &ac, const Angle &as, const Angle &af, double P1,
                                      double P2, double P3, double P4) {
  const double tiny10 = 10.0 * 2.2250738585072014e-308;
  const double cos_a = ac.cost, sin_a = ac.sint, cos_b = as.cost, sin_b = as.sint, cos_c = af.cost, sin_c = af.sint;
  const double cos_B = ac.cosp * af.cosp + ac.sinp * af.sinp;
  const double cos_C = as.cosp, sin_C = fabs(as.sinp);
  double cos_A, sin_A;
  if (sin_C < tiny10 && sin_c < tiny10) {
    cos_A = -cos_B * cos_C;
    sin_A = sqrt(1.0 - cos_A * cos_A);
  } else {
    cos_A = (cos_a - cos_b * cos_c) / (sin_b * sin_c);
    sin_A = sin_C * sin_a / sin_c;
  }
  const double cos_2_i2 = 1.0 - 2.0 * sin_A * sin_A;
  const double sin_2_i2 = 2.0 * sin_A * cos_A;
  const double cos_2_alpha = 1.0 - 2.0 * as.sinp * as.sinp;
  const double sin_2_alpha = -2.0 * as.sinp * as.cosp;
  const double cos_2_beta = cos_2_i2;
  const double sin_2_beta = (as.sinp < 0.0) ? sin_2_i2 : -sin_2_i2;
  const double RLS1 = P1 * s.I + P2 * (cos_2_alpha * s.Q + sin_2_alpha * s.U);
  const double RLS2 = P2 * s.I + P1 * (cos_2_alpha * s.Q + sin_2_alpha * s.U);
  const double RLS3 = -P4 * s.V + P3 * (-sin_2_alpha * s.Q + cos_2_alpha * s.U);
  const double RLS4 = P3 * s.V + P4 * (-sin_2_alpha * s.Q + cos_2_alpha * s.U);
  s.I = RLS1;
  s.Q = cos_2_beta * RLS2 + sin_2_beta * RLS3;
  s.U = -sin_2_beta * RLS2 + cos_2_beta * RLS3;
  s.V = RLS4;
}

// bilinear (mu, nu) interpolation of a phase-matrix element (interp2d_dp, lib_array.f90:780-846)
__device__ __forceinline__ double interp_phase(const double *__restrict__ P, int n_mu, int i, int j, double wx0,
                                               double wx1, double wy0, double wy1, double norm) {
  const double *r0 = P + (size_t)j * n_mu + i;
  const double *r1 = r0 + n_mu;
  return __ldg(r0) * wx1 * wy1 * norm + __ldg(r0 + 1) * wx0 * wy1 * norm + __ldg(r1) * wx1 * wy0 * norm +
         __ldg(r1 + 1) * wx0 * wy0 * norm;
}

}  // namespace hyp
