// geometry_sph.cuh -- spherical polar grid traversal on the device.
//
// Restates src/grid/grid_geometry_spherical_3d.f90 (find_cell :212-276, adjust_wall :278-469,
// next_cell :530-557, escaped :493-500, find_wall :741-1073, insert_t :1083-1111) for a ray kept in
// "origin + path length" form: the reference re-derives the quadratic coefficients from the current
// position at every wall, here they are fixed per flight (v2_xy, rv_xy, r2_xy, ... at the origin) and the
// roots are absolute path lengths, so rounding does not accumulate along a flight.  The wall bookkeeping
// (on_wall ids, "discard the root closest to the current position", ULP tolerances that let a ray cross
// two walls at once) is the reference's.
#pragma once

namespace hyp {

// Tables of a spherical polar grid in one device buffer (offsets in doubles).
// The same structure also carries a cylindrical polar grid (kind == POLAR_CYL,
// src/grid/grid_geometry_cylindrical_3d.f90): w1 = cylinder radii, w2 = z planes, w3 = phi half-planes;
// o_ew2 then holds 3*spacing(z walls) and the theta tables are unused.
enum { POLAR_SPH = 1, POLAR_CYL = 2 };

struct SphGrid {
  const double *T;
  int32_t kind;                  // POLAR_SPH / POLAR_CYL
  int32_t n1, n2, n3, midplane;  // midplane: 0-based index of the theta wall at pi/2, or -1
  int32_t o_w1, o_wr2, o_ew1;    // [n1+1] r walls, squared, 3*spacing
  int32_t o_w2, o_wtant, o_wtant2;  // [n2+1] theta walls, tan, tan^2
  int32_t o_w3, o_wtanp, o_wcosp, o_wsinp;  // [n3+1] phi walls, tan, cos, sin
  int32_t o_dr3, o_dcost, o_dphi;   // [n1], [n2], [n3] for the cell volumes
  int32_t o_wcost;                  // [n2+1] cos(theta walls)
  int32_t o_ew2;                    // [n2+1] cylindrical: 3*spacing(z walls)
};

constexpr double SPH_PI = 3.14159265358979323846;
constexpr double SPH_TWOPI = SPH_PI + SPH_PI;
constexpr double SPH_EW_ANGLE = 3.0 * 2.220446049250313e-16;  // 3*spacing(1): ew2, ew3 (:199-200)
constexpr double SPH_HUGE = 1.7976931348623157e308;

__device__ __forceinline__ double spacing_dev(double x) {
  if (x == 0.0) return 2.2250738585072014e-308;
  const double ax = fabs(x);
  return __longlong_as_double(__double_as_longlong(ax) + 1) - ax;
}

__device__ __forceinline__ bool equal_nulp(double x, double y, int n) {
  if (x == y) return true;
  return fabs(x - y) <= n * spacing_dev(fmax(x, y));
}

// locate_dp (fortranlib/src/lib_array.f90:917-950) on an ascending array, 0-based result, -1 outside
__device__ __forceinline__ int locate0(const double *__restrict__ x, int n, double v) {
  if (!(v >= __ldg(x) && v <= __ldg(x + n - 1))) return -1;
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(x + mid) <= v) lo = mid; else hi = mid;
  }
  return lo;
}

struct SphRay {
  double r0x, r0y, r0z, vx, vy, vz;
  double v2_xy, v2_z, rv_xy, rv_z, r2_xy, r2_z;  // at the origin of the flight
  double t;                                       // path length travelled
  int i1, i2, i3, ic;                             // 0-based cell; ic = id used for density / deposits
  int ow1, ow2, ow3;                              // on_wall_id
  bool radial;                                    // moving outwards at the start (grid_propagate_3d.f90:73)
};

__device__ __forceinline__ void sph_angles(double rx, double ry, double rz, double vx, double vy, double vz,
                                           double &r_sq, double &w_sq, double &theta, double &phi) {
  r_sq = rx * rx + ry * ry + rz * rz;
  w_sq = rx * rx + ry * ry;
  theta = r_sq == 0.0 ? atan2(sqrt(vx * vx + vy * vy), vz) : atan2(sqrt(rx * rx + ry * ry), rz);
  phi = w_sq == 0.0 ? atan2(vy, vx) : atan2(ry, rx);
  if (phi < 0.0) phi += SPH_TWOPI;
}

// find_cell: false if outside the grid
__device__ inline bool sph_find_cell(const SphGrid &G, double rx, double ry, double rz, double vx, double vy, double vz,
                                     int &i1, int &i2, int &i3) {
  double r_sq, w_sq, theta, phi;
  sph_angles(rx, ry, rz, vx, vy, vz, r_sq, w_sq, theta, phi);
  if (G.kind == POLAR_CYL) {
    // grid_geometry_cylindrical_3d.f90:184-237
    i1 = locate0(G.T + G.o_wr2, G.n1 + 1, w_sq);
    i2 = locate0(G.T + G.o_w2, G.n2 + 1, rz);
  } else {
    i1 = locate0(G.T + G.o_wr2, G.n1 + 1, r_sq);
    i2 = locate0(G.T + G.o_w2, G.n2 + 1, theta);
  }
  i3 = locate0(G.T + G.o_w3, G.n3 + 1, phi);
  return i1 >= 0 && i2 >= 0 && i3 >= 0;
}

// Start a ray at (r, v) in the cell find_cell reported: computes the per-flight constants and applies
// adjust_wall.  ic keeps the id of the cell find_cell reported (the reference does not refresh it).
__device__ inline void sph_start(const SphGrid &G, SphRay &R, double rx, double ry, double rz, double vx, double vy,
                                 double vz, int i1, int i2, int i3) {
  R.r0x = rx; R.r0y = ry; R.r0z = rz;
  R.vx = vx; R.vy = vy; R.vz = vz;
  R.v2_xy = vx * vx + vy * vy;
  R.v2_z = vz * vz;
  R.rv_xy = rx * vx + ry * vy;
  R.rv_z = rz * vz;
  R.r2_xy = rx * rx + ry * ry;
  R.r2_z = rz * rz;
  R.t = 0.0;
  R.ic = (i3 * G.n2 + i2) * G.n1 + i1;
  R.ow1 = R.ow2 = R.ow3 = 0;
  R.radial = (rx * vx + ry * vy + rz * vz) > 0.0;
  const int eps = 3;
  double r_sq, w_sq, theta, phi;
  sph_angles(rx, ry, rz, vx, vy, vz, r_sq, w_sq, theta, phi);
  const double *wr2 = G.T + G.o_wr2, *w2 = G.T + G.o_w2, *w3 = G.T + G.o_w3, *wtant = G.T + G.o_wtant;
  if (G.kind == POLAR_CYL) {
    // adjust_wall of the cylindrical grid (grid_geometry_cylindrical_3d.f90:239-346)
    R.radial = false;  // the cylindrical find_wall always tests the inner wall
    if (rx * vx + ry * vy >= 0.0) {
      if (equal_nulp(w_sq, wr2[i1], eps)) { R.ow1 = -1; }
      else if (equal_nulp(w_sq, wr2[i1 + 1], eps)) { R.ow1 = -1; i1 += 1; }
    } else {
      if (equal_nulp(w_sq, wr2[i1], eps)) { R.ow1 = +1; i1 -= 1; }
      else if (equal_nulp(w_sq, wr2[i1 + 1], eps)) { R.ow1 = +1; }
    }
    if (vz > 0.0) {
      if (equal_nulp(rz, w2[i2], eps)) { R.ow2 = -1; }
      else if (equal_nulp(rz, w2[i2 + 1], eps)) { R.ow2 = -1; i2 += 1; }
    } else if (vz < 0.0) {
      if (equal_nulp(rz, w2[i2], eps)) { R.ow2 = +1; i2 -= 1; }
      else if (equal_nulp(rz, w2[i2 + 1], eps)) { R.ow2 = +1; }
    }
  } else {
  // radial walls
  if (rx * vx + ry * vy + rz * vz >= 0.0) {
    if (equal_nulp(r_sq, wr2[i1], eps)) {
      R.ow1 = -1;
    } else if (equal_nulp(r_sq, wr2[i1 + 1], eps)) {
      R.ow1 = -1;
      i1 += 1;
    }
  } else {
    if (equal_nulp(r_sq, wr2[i1], eps)) {
      R.ow1 = +1;
      i1 -= 1;
    } else if (equal_nulp(r_sq, wr2[i1 + 1], eps)) {
      R.ow1 = +1;
    }
  }
  // theta walls
  if (r_sq == 0.0) {
    if (fabs(vz) < 1.0) {
      const double theta_v = atan2(sqrt(vx * vx + vy * vy), vz);
      if (equal_nulp(theta_v, w2[i2], eps)) R.ow2 = -1;
      else if (equal_nulp(theta_v, w2[i2 + 1], eps)) R.ow2 = +1;
    }
  } else if (i2 > 0 && equal_nulp(theta, w2[i2], eps)) {
    if (i2 == G.midplane) {
      if (vz > 0.0) { R.ow2 = +1; i2 -= 1; } else { R.ow2 = -1; }
    } else {
      const bool lhs = sqrt(w_sq) * vz * wtant[i2] - (rx * vx + ry * vy) < 0.0;
      if (lhs == (rz > 0.0)) { R.ow2 = -1; } else { R.ow2 = +1; i2 -= 1; }
    }
  } else if (i2 + 1 < G.n2 && equal_nulp(theta, w2[i2 + 1], eps)) {
    if (i2 + 1 == G.midplane) {
      if (vz > 0.0) { R.ow2 = +1; } else { R.ow2 = -1; i2 += 1; }
    } else {
      const bool lhs = sqrt(w_sq) * vz * wtant[i2 + 1] - (rx * vx + ry * vy) < 0.0;
      if (lhs == (rz > 0.0)) { R.ow2 = -1; i2 += 1; } else { R.ow2 = +1; }
    }
  }
  }
  // phi walls
  if (rx == 0.0 && ry == 0.0 && vx == 0.0 && vy == 0.0) {
    // on the axis moving along it: on every phi wall at once, leave alone
  } else if (equal_nulp(phi, w3[i3], eps)) {
    double dphi = atan2(vy, vx) - w3[i3];
    if (dphi < -SPH_PI) dphi += SPH_TWOPI;
    if (dphi > 0.0) { R.ow3 = -1; } else { R.ow3 = +1; i3 -= 1; if (i3 < 0) i3 = G.n3 - 1; }
  } else if (equal_nulp(phi, w3[i3 + 1], eps)) {
    double dphi = atan2(vy, vx) - w3[i3 + 1];
    if (dphi < -SPH_PI) dphi += SPH_TWOPI;
    if (dphi > 0.0) { R.ow3 = -1; i3 += 1; if (i3 == G.n3) i3 = 0; } else { R.ow3 = +1; }
  }
  R.i1 = i1; R.i2 = i2; R.i3 = i3;
}

__device__ __forceinline__ bool sph_escaped(const SphGrid &G, const SphRay &R) {
  // spherical: radial only (:493-500); cylindrical: w and z (grid_geometry_cylindrical_3d.f90:375-384)
  return (unsigned)R.i1 >= (unsigned)G.n1 || (G.kind == POLAR_CYL && (unsigned)R.i2 >= (unsigned)G.n2);
}

// nearest-wall search state (reset_t / insert_t / find_next_wall)
struct WallSearch {
  double tmin, emin;  // tmin is relative to the current position, as in the reference
  int w1, w2, w3;     // imin
  __device__ __forceinline__ void reset() {
    tmin = SPH_HUGE;
    emin = 0.0;
    w1 = w2 = w3 = 0;
  }
  __device__ __forceinline__ void insert(double t, int iw, int i, double e) {
    if (t > 0.0) {
      const double emax = fmax(e, emin);
      if (t < tmin - emax) {
        tmin = t;
        emin = emax;
        w1 = iw == 1 ? i : 0;
        w2 = iw == 2 ? i : 0;
        w3 = iw == 3 ? i : 0;
      } else if (t < tmin + emax) {
        emin = emax;
        if (iw == 1) w1 = i; else if (iw == 2) w2 = i; else w3 = i;
      }
    }
  }
};

// quadratic_pascal_reduced_dp (fortranlib/src/lib_algebra.f90:145-164): x^2 + b x + c = 0
__device__ __forceinline__ void quad_pascal_reduced(double b, double c, double &x1, double &x2) {
  double delta = b * b - 4.0 * c;
  if (delta > 0.0) {
    delta = copysign(sqrt(delta), b);
    const double q = -0.5 * (b + delta);
    x1 = q;
    x2 = c / q;
  } else if (delta < 0.0) {
    x1 = -SPH_HUGE;
    x2 = -SPH_HUGE;
  } else {
    x1 = -2.0 * c / b;
    x2 = -SPH_HUGE;
  }
}

// quadratic_dp (lib_algebra.f90:107-122)
__device__ __forceinline__ void quad_plain(double a, double b, double c, double &x1, double &x2) {
  double delta = b * b - 4.0 * a * c;
  if (delta > 0.0) {
    delta = sqrt(delta);
    const double f = 0.5 / a;
    x1 = (-b - delta) * f;
    x2 = (-b + delta) * f;
  } else {
    x1 = SPH_HUGE;
    x2 = SPH_HUGE;
  }
}

// One cone wall (find_wall :822-962).  iw = 0-based theta wall, side = -1 lower / +1 upper.
__device__ __forceinline__ void sph_cone(const SphGrid &G, const SphRay &R, WallSearch &S, int &iext2, int iw, int side) {
  const double wtant = __ldg(G.T + G.o_wtant + iw), wtant2 = __ldg(G.T + G.o_wtant2 + iw);
  const double t0 = R.t;
  if (R.ow2 == side) {
    // moving along the wall?  (quantities at the current position)
    const double rv_xy = R.rv_xy + t0 * R.v2_xy;
    const double cx = R.r0x + t0 * R.vx, cy = R.r0y + t0 * R.vy;
    if (equal_nulp(wtant, sqrt(R.v2_xy) / R.vz, 10) && equal_nulp(sqrt(cx * cx + cy * cy) * R.vz * wtant, rv_xy, 10)) {
      iext2 = side;
      return;
    }
  }
  if (iw == G.midplane && R.vz != 0.0) {
    if (R.ow2 != side) S.insert(-R.r0z / R.vz - t0, 2, side, SPH_EW_ANGLE);
    return;
  }
  const double pA = R.v2_xy - R.v2_z * wtant2;
  double pB = R.rv_xy - R.rv_z * wtant2;
  pB = pB + pB;
  const double pC = R.r2_xy - R.r2_z * wtant2;
  if (fabs(pA) > 0.0) {
    double t1, t2;
    quad_plain(pA, pB, pC, t1, t2);
    // keep only the nappe of this wall
    if ((R.r0z + R.vz * t1 > 0.0) != (wtant > 0.0)) t1 = SPH_HUGE;
    if ((R.r0z + R.vz * t2 > 0.0) != (wtant > 0.0)) t2 = SPH_HUGE;
    t1 -= t0;
    t2 -= t0;
    if (R.ow2 == side) {
      // the root closest to the current position is the wall we sit on
      S.insert(fabs(t1) < fabs(t2) ? t2 : t1, 2, side, SPH_EW_ANGLE);
    } else {
      S.insert(t1, 2, side, SPH_EW_ANGLE);
      S.insert(t2, 2, side, SPH_EW_ANGLE);
    }
  } else if (fabs(pB) > 0.0) {
    if (R.ow2 != side) S.insert(-pC / pB - t0, 2, side, SPH_EW_ANGLE);
  }
}

// find_wall: distance from the current position to the next wall and the wall ids crossed there.
// Returns false if no wall was found (the reference kills such packets).
__device__ inline bool sph_find_wall(const SphGrid &G, const SphRay &R, double &dt, int &d1, int &d2, int &d3) {
  WallSearch S;
  S.reset();
  int iext2 = 0, iext3 = 0;
  const double t0 = R.t;
  const double *wr2 = G.T + G.o_wr2, *ew1 = G.T + G.o_ew1;
  double t1, t2;
  if (G.kind == POLAR_CYL) {
    // cylinders and z planes (grid_geometry_cylindrical_3d.f90:592-680)
    double qB = R.rv_xy / R.v2_xy;
    qB = qB + qB;
    const double qC = R.r2_xy / R.v2_xy;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int side = s == 0 ? -1 : +1;
      quad_pascal_reduced(qB, qC - __ldg(wr2 + R.i1 + s) / R.v2_xy, t1, t2);
      t1 -= t0;
      t2 -= t0;
      const double e = __ldg(ew1 + R.i1 + s);
      if (R.ow1 == side) {
        S.insert(fabs(t1) < fabs(t2) ? t2 : t1, 1, side, e);
      } else {
        S.insert(t1, 1, side, e);
        S.insert(t2, 1, side, e);
      }
    }
    const double *w2 = G.T + G.o_w2;
    if (R.ow2 != -1) S.insert((__ldg(w2 + R.i2) - R.r0z) / R.vz - t0, 2, -1, 0.0);
    if (R.ow2 != +1) S.insert((__ldg(w2 + R.i2 + 1) - R.r0z) / R.vz - t0, 2, +1, 0.0);
  } else {
  // spheres: |r0 + t v|^2 = R^2
  double pB = R.rv_xy + R.rv_z;
  pB = pB + pB;
  const double pC = R.r2_xy + R.r2_z;
  if (!R.radial) {
    quad_pascal_reduced(pB, pC - __ldg(wr2 + R.i1), t1, t2);
    t1 -= t0;
    t2 -= t0;
    const double e = __ldg(ew1 + R.i1);
    if (R.ow1 == -1) {
      S.insert(fabs(t1) < fabs(t2) ? t2 : t1, 1, -1, e);
    } else {
      S.insert(t1, 1, -1, e);
      S.insert(t2, 1, -1, e);
    }
  }
  {
    quad_pascal_reduced(pB, pC - __ldg(wr2 + R.i1 + 1), t1, t2);
    t1 -= t0;
    t2 -= t0;
    const double e = __ldg(ew1 + R.i1 + 1);
    if (R.ow1 == +1) {
      S.insert(fabs(t1) < fabs(t2) ? t2 : t1, 1, +1, e);
    } else {
      S.insert(t1, 1, +1, e);
      S.insert(t2, 1, +1, e);
    }
  }
  // cones (theta = 0 and theta = pi are not walls)
  if (R.i2 > 0) sph_cone(G, R, S, iext2, R.i2, -1);
  if (R.i2 < G.n2 - 1) sph_cone(G, R, S, iext2, R.i2 + 1, +1);
  }
  // half-planes of constant phi
  if (G.n3 > 1) {
    const double *w3 = G.T + G.o_w3;
    double dphi = 0.0;
    if (R.ow3 == -1) dphi = atan2(R.vy, R.vx) - __ldg(w3 + R.i3);
    if (R.ow3 == +1) dphi = atan2(R.vy, R.vx) - __ldg(w3 + R.i3 + 1);
    if (dphi > SPH_PI) dphi -= SPH_TWOPI;
    if (dphi < -SPH_PI) dphi += SPH_TWOPI;
    const double cx = R.r0x + t0 * R.vx, cy = R.r0y + t0 * R.vy;
    if (R.ow3 == +1 && fabs(dphi) < SPH_EW_ANGLE) {
      iext3 = +1;
    } else if (R.ow3 == -1 && fabs(dphi) < SPH_EW_ANGLE) {
      iext3 = -1;
    } else if (cx * cx + cy * cy > 0.0) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int side = s == 0 ? -1 : +1;
        if (R.ow3 == side) continue;
        const int iw = R.i3 + s;
        const double tp = __ldg(G.T + G.o_wtanp + iw);
        const double tt = -(tp * R.r0x - R.r0y) / (tp * R.vx - R.vy);
        // the intersection must lie on the half-plane of this wall, not on its continuation through
        // the axis (the reference tests |atan2(y_i, x_i) - phi_wall| < pi/2)
        const double xi = R.r0x + R.vx * tt, yi = R.r0y + R.vy * tt;
        if (xi * __ldg(G.T + G.o_wcosp + iw) + yi * __ldg(G.T + G.o_wsinp + iw) > 0.0) S.insert(tt - t0, 3, side, 0.0);
      }
    }
  }
  dt = S.tmin;
  d1 = S.w1;
  d2 = S.w2 + iext2;
  d3 = S.w3 + iext3;
  return (d1 | d2 | d3) != 0;
}

// next_cell_wall_id + opposite_wall: step across the walls found by sph_find_wall
__device__ __forceinline__ void sph_step(const SphGrid &G, SphRay &R, int d1, int d2, int d3) {
  R.i1 += d1;
  R.i2 += d2;
  R.i3 += d3;
  if (R.i3 < 0) R.i3 = G.n3 - 1;
  if (R.i3 >= G.n3) R.i3 = 0;
  R.ow1 = -d1;
  R.ow2 = -d2;
  R.ow3 = -d3;
  R.ic = (R.i3 * G.n2 + R.i2) * G.n1 + R.i1;
}

// spherical: dr3 * dcost * dphi / 3; cylindrical: the same slots hold dw2, dz, dphi and the divisor is 2
__device__ __forceinline__ double sph_volume(const SphGrid &G, int64_t ic) {
  const int i1 = (int)(ic % G.n1), i2 = (int)((ic / G.n1) % G.n2), i3 = (int)(ic / ((int64_t)G.n1 * G.n2));
  const double v = __ldg(G.T + G.o_dr3 + i1) * __ldg(G.T + G.o_dcost + i2) * __ldg(G.T + G.o_dphi + i3);
  return G.kind == POLAR_CYL ? v / 2.0 : v / 3.0;
}

}  // namespace hyp
