// imaging.cuh -- device side of the final (imaging) and raytracing iterations.
//
// Reference routines restated here for the GPU: do_final / propagate (src/main/iter_final.f90:60-273),
// forced first interaction (src/main/forced_interaction.f90:23-133), peeloff_photon
// (src/images/images_peeled.f90:95-270), grid_escape_tau / grid_escape_column_density
// (src/grid/grid_propagate_3d.f90:377-582), image_bin / image_bin_raytraced
// (src/images/image_type.f90:408-606), do_raytracing (src/main/iter_raytracing.f90:31-141) and
// emit_from_grid (src/grid/grid_physics_3d.f90:691-753).
//
// The imaging iteration reuses the packet pool of the Lucy iteration.  A round is
//   emit_final -> peel -> flight_final -> interact_final
// Every event that the reference peels off (an emission, and each interaction that survives) is
// written as a self-contained PeelJob; the peel kernel runs one thread per (job, viewing angle),
// marches the ray to the grid edge (8 B of density per crossing, no deposit) and adds the
// attenuated packet into the image / SED cubes with fp64 REDs.
//
// Included by hyperion_b200.cu after the Lucy kernels (it uses ModelDev, Slot, Photon, Lane, ...).
#pragma once

// ---------------------------------------------------------------------------------------------
// image cubes on the device
// ---------------------------------------------------------------------------------------------
struct ImageDev {
  int32_t n_view, n_nu, n_x, n_y, n_ap, n_orig, n_stokes;
  int32_t compute_image, compute_sed, track_origin, track_n_scat, uncertainties, ignore_optical_depth;
  int32_t n_sources, n_dust;
  int32_t inside_observer;  // the observer sits at (rpx, rpy, rpz) inside the grid (images_peeled.f90:95-270)
  int32_t inu_min;          // monochromatic mode (image_type.f90:243-258): channel k is frequency inu_min + k; 0 otherwise
  double x_min, x_max, y_min, y_max, ap_min, ap_max;
  double log10_ap_min, log10_ap_max, log10_nu_min, log10_nu_max;
  double d_min, d_max;
  double rpx, rpy, rpz;  // peeloff origin
  // accumulators, Fortran order (first index fastest): sed(n_nu, n_ap, n_view, n_orig, n_stokes),
  // img(n_nu, n_x, n_y, n_view, n_orig, n_stokes)  (image_type.f90:291,299); 2 = sum of squares, n = counts
  double *sed, *sed2, *sedn, *img, *img2, *imgn;
  // raytracing spectra on this group's frequency grid (images_peeled.f90:423-530)
  const double *src_spec;               // [n_sources][n_nu]
  const double *dust_chi;               // [n_dust][n_nu]
  const double *dust_logj[MAX_DUST];    // [n_jnu][n_nu] log10 of the binned emissivities
  // filter convolution (image_type.f90:274-284,467-476): channel k has points [filt_off[k], filt_off[k+1])
  int32_t use_filters, pad1;
  const int32_t *filt_off;
  const double *filt_nu, *filt_tr;
};

// interp1d_dp(x, y, xval, bounds_error=.false., fill_value=0) (lib_array.f90:616-624,704-778) with the
// bisection of locate_dp (:917-950): x may increase or decrease
__device__ inline double interp1d_fill0(const double *__restrict__ x, const double *__restrict__ y, int n, double xval) {
  const bool ascnd = x[n - 1] >= x[0];
  int jl = 0, ju = n + 1;
  while (ju - jl > 1) {
    const int jm = (ju + jl) / 2;
    if (ascnd == (xval >= x[jm - 1])) jl = jm; else ju = jm;
  }
  int ip = jl;
  if (xval == x[0]) ip = 1;
  else if (xval == x[n - 1]) ip = n - 1;
  else if (ascnd ? (xval > x[n - 1] || xval < x[0]) : (xval < x[n - 1] || xval > x[0])) return 0.0;
  if (ip < n && ip > 0) {
    const double frac = (xval - x[ip - 1]) / (x[ip] - x[ip - 1]);
    return y[ip - 1] + frac * (y[ip] - y[ip - 1]);
  }
  return ip == n ? y[n - 1] : y[0];
}

struct ViewDev {
  int32_t group, view;  // 0-based group, 0-based view inside the group
  Angle a;
};

struct ImagingDev {
  const ImageDev *images;
  const ViewDev *views;
  int32_t n_groups, n_views;
  // Column densities from every point source to the grid edge towards every view: [n_sources][n_views]
  // records of ND columns + the number of cells crossed (-1: line of sight blocked by a star).  All
  // packets a point source emits share these rays, so their peel-offs need no march of their own
  // (tau = sum_d chi_d(nu) * column_d): the reference marches each of them, which on a GPU also makes
  // every SM hammer the same cells at the same time.
  const double *src_columns;
};

// One peel-off event.  kind 0: the last event was isotropic (emission from a point source, thermal
// re-emission); kind 1: a scattering -- the direction and Stokes vector BEFORE the scattering are kept so
// that the phase matrix can be evaluated towards each observer (dust_scatter_peeloff,
// src/dust/dust_type_4elem.f90:421-444).
template <int ND>
struct PeelJob {
  double rx, ry, rz, nu, energy;
  double chi[ND];
  double vpx, vpy, vpz, sQ, sU, sV;
  double emiss_var_frac;
  int32_t kind, source_id, dust_id, n_scat;      // ids are 1-based as in the reference
  int32_t scattered, reprocessed, emiss_type, emiss_var_id;
  int32_t point_src, pad;  // 1-based id of the point source this job was just emitted from, else 0
};

// bits of Slot::tag / Photon::tag during the final iteration
// bits 0-11 source id (1-based, up to MAX_SOURCES), 12 scattered, 13 reprocessed, 14-19 successive re-absorptions
// (hyperion_b200.cu), 20-29 number of scatterings (saturating), 30-31 dust type of the last interaction
constexpr uint32_t TAG_SRC_MASK = 0xfffu, TAG_SCATTERED = 0x1000u, TAG_REPROCESSED = 0x2000u;
// the packet is in the middle of a modified random walk that was interrupted because the peel-off queue of the
// round was full: its next "flight" has optical depth zero and the interaction kernel resumes the walk
constexpr uint32_t TAG_MRW_PAUSED = 0x4000u;
constexpr int TAG_NSCAT_SHIFT = 20;
constexpr uint32_t TAG_LOW_MASK = (1u << TAG_NSCAT_SHIFT) - 1u;

// ipos_dp (fortranlib/src/lib_array.f90:954-998): 1-based bin, 0 / nbin+1 outside
__device__ __forceinline__ int ipos_bin(double xmin, double xmax, double x, int nbin) {
  if (xmax > xmin) {
    if (x < xmin) return 0;
    if (x > xmax) return nbin + 1;
    if (x < xmax) return (int)((x - xmin) / (xmax - xmin) * (double)nbin) + 1;
    return nbin;
  }
  if (x > xmin) return 0;
  if (x < xmax) return nbin + 1;
  if (x > xmax) return (int)((x - xmin) / (xmax - xmin) * (double)nbin) + 1;
  return nbin;
}

__device__ __forceinline__ Angle angle_of(double vx, double vy, double vz) {
  Angle a;
  a.cost = vz;
  a.sint = sqrt(vx * vx + vy * vy);
  if (a.sint > 0.0) {
    a.cosp = vx / a.sint;
    a.sinp = vy / a.sint;
  } else {
    a.cosp = 1.0;
    a.sinp = 0.0;
  }
  return a;
}

// difference_angle3d_dp (fortranlib/src/type_angle3d.f90:283-419): the local angle that takes
// a_coord into a_final, i.e. the inverse of rotate_angle.
__device__ inline Angle difference_angle(const Angle &c, const Angle &f) {
  Angle l;
  if (fabs(c.sint) < 1.e-10) {
    l = f;
    if (c.cost > 0.0) {
      l.cosp = c.cosp * f.cosp + c.sinp * f.sinp;
      l.sinp = -c.cosp * f.sinp + c.sinp * f.cosp;
    } else {
      l.cost = -l.cost;
      l.cosp = c.cosp * f.cosp + c.sinp * f.sinp;
      l.sinp = c.cosp * f.sinp - c.sinp * f.cosp;
    }
    return l;
  }
  const double cos_a = c.cost, sin_a = c.sint, cos_c = f.cost, sin_c = f.sint;
  const double cos_B = c.cosp * f.cosp + c.sinp * f.sinp;
  const double sin_B = c.sinp * f.cosp - c.cosp * f.sinp;
  const double cos_b = cos_a * cos_c + sin_a * sin_c * cos_B;
  const double sin_b = sin2cos(cos_b);
  if (fabs(cos_b + 1.0) < 1.e-10) return Angle{-1.0, 0.0, 1.0, 0.0};  // angle3d_deg(180, 0) up to rounding of sin(pi)
  if (fabs(cos_b - 1.0) < 1.e-10) return Angle{1.0, 0.0, 1.0, 0.0};
  // the reference's logical expression parses as  A .eqv. (B .and. C) .eqv. D
  const bool same_sign = ((cos_a > 0.0) == ((cos_b > 0.0) && (sin_a > 0.0))) == (sin_b > 0.0);
  const bool sin_dom = fabs(sin_a) > fabs(cos_a);
  const double delta = sin_dom ? cos_b - cos_a : sin_b - sin_a;
  double sin_C, cos_C;
  if (same_sign && fabs(delta) < 1.e-5 && sin_c < 1.e-5) {
    const double q = sin_dom ? cos_a / sin_a : sin_a / cos_a;
    const double diff = (sin_c * sin_c - delta * delta * (1.0 + q * q)) / (sin_a * sin_b);
    sin_C = diff >= 0.0 ? sqrt(diff) : 0.0;
    cos_C = cos_c > 0.0 ? sin2cos(sin_C) : -sin2cos(sin_C);
  } else {
    sin_C = fabs(sin_B) * sin_c / sin_b;
    cos_C = (cos_c - cos_a * cos_b) / (sin_a * sin_b);
  }
  if (sin_C == 0.0) sin_C = 2.2250738585072014e-308;
  l.cost = cos_b;
  l.sint = sin_b;
  l.cosp = cos_C;
  l.sinp = sin_B < 0.0 ? sin_C : -sin_C;
  return l;
}

// find_cell + adjust_wall for all three axes (grid_geometry_cartesian_3d.f90:143-259).  ic is the 1-D id
// find_cell reports; the reference keeps it for the first segment even when adjust_wall moved an index.
__device__ __forceinline__ bool place_in_grid(const ModelDev &M, double rx, double ry, double rz, double vx, double vy,
                                              double vz, int &ix, int &iy, int &iz, int &ic) {
  int fx, fy, fz;
  bool ok = place_axis(M.w1, M.n1, rx, vx, ix, fx);
  ok = place_axis(M.w2, M.n2, ry, vy, iy, fy) && ok;
  ok = place_axis(M.w3, M.n3, rz, vz, iz, fz) && ok;
  if (!ok) return false;
  ic = (fz * M.n2 + fy) * M.n1 + fx;
  return true;
}

__device__ __forceinline__ bool lane_outside(int ix, int iy, int iz, int n1, int n2, int n3) {
  return (unsigned)ix >= (unsigned)n1 || (unsigned)iy >= (unsigned)n2 || (unsigned)iz >= (unsigned)n3;
}

// March a ray from its current cell to the edge of the grid (grid_escape_tau with tmax = huge /
// grid_escape_column_density).  COLUMN: accumulate the column density of every dust type instead of the
// optical depth.  D crossings are resolved geometrically before their densities are consumed, so D loads
// are in flight per lane (the cell sequence does not depend on the density).
// At most max_groups groups are marched per call; returns true once the ray has left the grid.
template <int ND, bool COLUMN, int D>
__device__ __forceinline__ bool escape_march(Lane<ND> &L, const double *__restrict__ W,
                                             const double *__restrict__ rho, const int n1, const int n2,
                                             const int n3, double &tau, double (&col)[ND], uint32_t &n_cross,
                                             const int max_groups = 0x7fffffff,
                                             const double tmax = 1.7976931348623157e308) {
  // tmax: the march ends after that path length (inside observers: grid_propagate_3d.f90:440-443)
  const int o2 = n1 + 1, o3 = n1 + n2 + 2;
  bool dead = lane_outside(L.ix, L.iy, L.iz, n1, n2, n3);
  for (int g = 0; g < max_groups && !dead; ++g) {
    double ds_s[D], rho_s[D][ND];
#pragma unroll
    for (int j = 0; j < D; ++j) {
#pragma unroll
      for (int id = 0; id < ND; ++id) rho_s[j][id] = __ldg(rho + (size_t)L.ic * ND + id);
      const bool bx = (L.tnx <= L.tny) & (L.tnx <= L.tnz);
      const bool by = (!bx) & (L.tny <= L.tnz);
      const double t_exit = bx ? L.tnx : (by ? L.tny : L.tnz);
      const double iv_ax = bx ? L.ivx : (by ? L.ivy : L.ivz);
      const int fwd = iv_ax > 0.0 ? 1 : 0;
      const int i_new = (bx ? L.ix : (by ? L.iy : L.iz)) + 2 * fwd - 1;
      const int n_ax = bx ? n1 : (by ? n2 : n3);
      const bool out = (unsigned)i_new >= (unsigned)n_ax;
      const int woff = bx ? 0 : (by ? o2 : o3);
      const double wall = W[woff + (out ? 0 : i_new + fwd)];
      const double tn_new = (wall - (bx ? L.r0x : (by ? L.r0y : L.r0z))) * iv_ax;
      const bool live = !dead;
      const bool past = t_exit > tmax;
      const bool moved = live & !out & !past;
      const double t_end = past ? tmax : t_exit;
      ds_s[j] = live ? t_end - L.t : -1.0;
      L.t = live ? t_end : L.t;
      L.ix = (moved & bx) ? i_new : L.ix;
      L.iy = (moved & by) ? i_new : L.iy;
      L.iz = (moved & !(bx | by)) ? i_new : L.iz;
      L.tnx = (moved & bx) ? tn_new : L.tnx;
      L.tny = (moved & by) ? tn_new : L.tny;
      L.tnz = (moved & !(bx | by)) ? tn_new : L.tnz;
      L.ic = moved ? (L.iz * n2 + L.iy) * n1 + L.ix : L.ic;
      dead |= out | past;
    }
#pragma unroll
    for (int j = 0; j < D; ++j) {
      if (ds_s[j] >= 0.0) {
        ++n_cross;
#pragma unroll
        for (int id = 0; id < ND; ++id) {
          if (COLUMN)
            col[id] = col[id] + rho_s[j][id] * ds_s[j];
          else
            tau = tau + L.chi[id] * rho_s[j][id] * ds_s[j];
        }
      }
    }
  }
  return dead;
}

// ---------------------------------------------------------------------------------------------
// binning (image_type.f90:117-134, 337-606)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int origin_slice(const ImageDev &im, int scattered, int reprocessed, int source_id,
                                            int dust_id, int n_scat) {
  const int iorig = scattered ? (reprocessed ? 4 : 3) : (reprocessed ? 2 : 1);
  if (im.track_origin == HYP_TRACK_DETAILED) {
    int io = ((iorig - iorig % 2) * im.n_sources + (iorig - (iorig + 1) % 2 - 1) * im.n_dust) / 2;
    return io + (iorig % 2 == 0 ? dust_id : source_id);
  }
  if (im.track_origin == HYP_TRACK_SCATTERINGS) {
    int io = n_scat > im.track_n_scat ? im.track_n_scat + 2 : n_scat + 1;
    if (reprocessed) io += im.track_n_scat + 2;
    return io;
  }
  if (im.track_origin == HYP_TRACK_BASIC) return iorig;
  return 1;
}

__device__ __forceinline__ bool in_image(const ImageDev &im, double x, double y) {
  if (im.compute_image) {
    if ((x >= im.x_min && x <= im.x_max) || (x <= im.x_min && x >= im.x_max))
      if ((y >= im.y_min && y <= im.y_max) || (y <= im.y_min && y >= im.y_max)) return true;
  }
  if (im.compute_sed && x * x + y * y <= im.ap_max * im.ap_max) return true;
  return false;
}

__device__ __forceinline__ int find_sed_bin(const ImageDev &im, double x, double y) {
  const double log10_r = log10(sqrt(x * x + y * y));
  if (log10_r < im.log10_ap_min || im.n_ap == 1) return 1;
  return ipos_bin(im.log10_ap_min, im.log10_ap_max, log10_r, im.n_ap - 1) + 1;
}

__device__ __forceinline__ void bin_add(double *a, double *a2, double *an, size_t k, double v, bool unc) {
  atomicAdd(a + k, v);
  if (unc) {
    atomicAdd(a2 + k, v * v);
    atomicAdd(an + k, 1.0);
  }
}

// ---------------------------------------------------------------------------------------------
// the peel-off kernel: one thread per (job, view); threads of a warp share the view
// ---------------------------------------------------------------------------------------------
constexpr int PEEL_THREADS = 256;
constexpr int PEEL_LOOKAHEAD = 4;

// Stokes vector a peel-off carries towards view direction a_req before attenuation:
// isotropic events 1, stellar surfaces the limb law, scatterings the phase matrix.
template <int ND>
__device__ inline Stokes peel_stokes(const ModelDev &M, const PeelJob<ND> &J, const Angle &a_req) {
  Stokes S{1.0, 0.0, 0.0, 0.0};
  if (J.kind == 2) {
    // emit_from_sphere_peeloff (source_type.f90:692-707): (vpx, vpy, vpz) is the outward normal of the
    // stellar surface at the emission point; the weights integrate to 4 pi over the sphere
    const SourceDev &src = M.sources[J.source_id - 1];
    const double vxr = a_req.sint * a_req.cosp, vyr = a_req.sint * a_req.sinp, vzr = a_req.cost;
    const double mu = fmax(vxr * J.vpx + vyr * J.vpy + vzr * J.vpz, 0.0);
    S.I = !src.peeloff ? 0.0 : (src.limb ? 2.0 * (1.5 * mu * mu + mu) : 4.0 * mu);
  } else if (J.kind == 1) {
    // dust_scatter_peeloff (dust_type_4elem.f90:421-444)
    const DustDev &d = M.dust[J.dust_id - 1];
    const Angle a_prev = angle_of(J.vpx, J.vpy, J.vpz);
    const Angle as = difference_angle(a_prev, a_req);
    S = Stokes{1.0, J.sQ, J.sU, J.sV};
    if (as.cost < d.L.mu_min || as.cost > d.L.mu_max) {
      S = Stokes{0.0, 0.0, 0.0, 0.0};
    } else {
      const double *nu = d.B + d.L.o_nu;
      const double *mu = d.B + d.L.o_mu;
      const int n_mu = d.L.n_mu;
      const int j = lower_interval(nu, d.L.n_nu, J.nu);
      const int i = lower_interval(mu, n_mu, as.cost);
      const double x0 = __ldg(mu + i), x1 = __ldg(mu + i + 1);
      const double y0 = __ldg(nu + j), y1 = __ldg(nu + j + 1);
      const double norm = 1.0 / (x1 - x0) / (y1 - y0);
      const double wx0 = as.cost - x0, wx1 = x1 - as.cost, wy0 = J.nu - y0, wy1 = y1 - J.nu;
      const double P1 = interp_phase(d.B + d.L.o_P1, n_mu, i, j, wx0, wx1, wy0, wy1, norm);
      const double P2 = interp_phase(d.B + d.L.o_P2, n_mu, i, j, wx0, wx1, wy0, wy1, norm);
      const double P3 = interp_phase(d.B + d.L.o_P3, n_mu, i, j, wx0, wx1, wy0, wy1, norm);
      const double P4 = interp_phase(d.B + d.L.o_P4, n_mu, i, j, wx0, wx1, wy0, wy1, norm);
      scatter_stokes(S, a_prev, as, a_req, P1, P2, P3, P4);
    }
  }
  return S;
}

// image-plane coordinates of a peel-off (images_peeled.f90:196-211)
template <int ND>
__device__ __forceinline__ void peel_image_xy(const PeelJob<ND> &J, const ImageDev &im, const Angle &a, double &x_image,
                                              double &y_image) {
  const double dx = J.rx - im.rpx, dy = J.ry - im.rpy, dz = J.rz - im.rpz;
  x_image = dy * a.cosp - dx * a.sinp;
  y_image = dz * a.sint - dy * a.cost * a.sinp - dx * a.cost * a.cosp;
}

// Inside observer (images_peeled.f90:131-183, 410-421): the peel-off travels from the event to the observer at
// (rpx, rpy, rpz); the image axes are the longitude and latitude of the arrival direction in the frame of
// the viewing direction a_view.  Returns the direction, the distance and the image coordinates.
template <int ND>
__device__ __forceinline__ void peel_inside(const PeelJob<ND> &J, const ImageDev &im, const Angle &a_view, Angle &a_req,
                                            double &dist, double &x_image, double &y_image) {
  // vector3d_to_angle3d (type_vector3d.f90:276-299) of r_peeloff - r
  const double wx = im.rpx - J.rx, wy = im.rpy - J.ry, wz = im.rpz - J.rz;
  const double small_r = sqrt(wx * wx + wy * wy), big_r = sqrt(wx * wx + wy * wy + wz * wz);
  a_req.cosp = wx / small_r;
  a_req.sinp = wy / small_r;
  a_req.cost = wz / big_r;
  a_req.sint = small_r / big_r;
  const double dx = J.rx - im.rpx, dy = J.ry - im.rpy, dz = J.rz - im.rpz;
  dist = sqrt(dx * dx + dy * dy + dz * dz);
  const double vax = a_req.sint * a_req.cosp, vay = a_req.sint * a_req.sinp, vaz = a_req.cost;
  const double sx = (vax * a_view.cosp + vay * a_view.sinp) * a_view.sint + vaz * a_view.cost;
  const double sy = -vax * a_view.sinp + vay * a_view.cosp;
  const double sz = -(vax * a_view.cosp + vay * a_view.sinp) * a_view.cost + vaz * a_view.sint;
  const double rad2deg = 180.0 / 3.14159265358979323846;
  x_image = atan2(sy, sx) * rad2deg;
  y_image = atan2(sqrt(sx * sx + sy * sy), sz) * rad2deg - 90.0;
  // Fortran MODULO: a - floor(a / p) * p
  const double ax = x_image - im.x_max, ay = y_image - im.y_min;
  x_image = im.x_max + (ax - floor(ax / 360.0) * 360.0);
  y_image = im.y_min + (ay - floor(ay / 360.0) * 360.0);
}

// image_bin / image_bin_raytraced (image_type.f90:408-606) for one finished ray
template <int ND, bool POLY>
__device__ inline void peel_bin(const PeelJob<ND> &J, const ImageDev &im, const ViewDev &V, const Stokes &S, const double tau,
                                const double (&col)[ND], const int mono_inu) {
  if (isnan(J.energy) || isnan(S.I)) return;
  double x_image, y_image;
  if (im.inside_observer) {
    Angle a_req;
    double dist;
    peel_inside<ND>(J, im, V.a, a_req, dist, x_image, y_image);
  } else {
    peel_image_xy<ND>(J, im, V.a, x_image, y_image);
  }
  const int io = origin_slice(im, J.scattered, J.reprocessed, J.source_id, J.dust_id, J.n_scat);
  const bool unc = im.uncertainties != 0;
  int ixp = 0, iyp = 0, ir = 0;
  bool in_img = false, in_sed = false;
  if (im.compute_image) {
    ixp = ipos_bin(im.x_min, im.x_max, x_image, im.n_x);
    iyp = ipos_bin(im.y_min, im.y_max, y_image, im.n_y);
    in_img = ixp >= 1 && ixp <= im.n_x && iyp >= 1 && iyp <= im.n_y;
  }
  if (im.compute_sed) {
    ir = find_sed_bin(im, x_image, y_image);
    in_sed = ir >= 1 && ir <= im.n_ap;
  }
  const size_t nn = (size_t)im.n_nu;
  // offsets of (inu = 1, stokes 0) in the two cubes
  const size_t k_img = nn * ((ixp - 1) + (size_t)im.n_x * ((iyp - 1) + (size_t)im.n_y * (V.view + (size_t)im.n_view * (io - 1))));
  const size_t s_img = nn * im.n_x * im.n_y * im.n_view * im.n_orig;
  const size_t k_sed = nn * ((ir - 1) + (size_t)im.n_ap * (V.view + (size_t)im.n_view * (io - 1)));
  const size_t s_sed = nn * im.n_ap * im.n_view * im.n_orig;
  if (POLY) {
    // image_bin_raytraced: the whole spectrum of the ray goes into the cube, attenuated per bin
    const double w = S.I * J.energy;
    const double *base0 = nullptr, *base1 = nullptr;
    if (J.emiss_type == 3) {
      base0 = im.dust_logj[J.dust_id - 1] + (size_t)(J.emiss_var_id) * nn;  // emiss_var_id is 0-based here
      base1 = base0 + nn;
    } else {
      base0 = im.src_spec + (size_t)(J.source_id - 1) * nn;
    }
    for (int inu = 0; inu < im.n_nu; ++inu) {
      double v;
      if (J.emiss_type == 3) {
        const double l0 = __ldg(base0 + inu), l1 = __ldg(base1 + inu);
        v = pow(10.0, (l1 - l0) * J.emiss_var_frac + l0);
        if (isnan(v)) v = 0.0;
      } else {
        v = __ldg(base0 + inu);
      }
      v = v * w;
#pragma unroll
      for (int id = 0; id < ND; ++id) v = v * exp(-col[id] * __ldg(im.dust_chi + (size_t)id * nn + inu));
      if (in_img) bin_add(im.img, im.img2, im.imgn, k_img + inu, v, unc);
      if (in_sed) bin_add(im.sed, im.sed2, im.sedn, k_sed + inu, v, unc);
    }
  } else {
    const double e = exp(-tau);
    const double st[4] = {S.I * e, S.Q * e, S.U * e, S.V * e};
    // without filters one channel, with filters every channel that transmits at this frequency
    // exact frequencies: the channel of the run's frequency (image_type.f90:435-436)
    const int inu0 = im.use_filters ? 1
                     : im.inu_min > 0 ? mono_inu - im.inu_min + 1
                                      : ipos_bin(im.log10_nu_min, im.log10_nu_max, log10(J.nu), im.n_nu);
    const int inu1 = im.use_filters ? im.n_nu : inu0;
    for (int inu = inu0; inu <= inu1; ++inu) {
      if (inu < 1 || inu > im.n_nu) return;
      double transmission = 1.0;
      if (im.use_filters) {
        const int o = im.filt_off[inu - 1];
        transmission = interp1d_fill0(im.filt_nu + o, im.filt_tr + o, im.filt_off[inu] - o, J.nu);
        if (!(transmission > 0.0)) continue;
      }
      for (int is = 0; is < im.n_stokes; ++is) {
        const double v = st[is] * J.energy * transmission;
        if (in_img) bin_add(im.img, im.img2, im.imgn, k_img + (inu - 1) + is * s_img, v, unc);
        if (in_sed) bin_add(im.sed, im.sed2, im.sedn, k_sed + (inu - 1) + is * s_sed, v, unc);
      }
    }
  }
}

constexpr int PEEL_GROUPS = 8;  // look-ahead groups (Cartesian) / crossings x4 (other grids) between two refill votes

// Persistent threads, one (job, view) ray per lane.  A lane that finishes its ray bins it and takes the
// next one from a global cursor while the other lanes of the warp keep marching, so rays of very different
// length (a peel-off from the near or the far side of the grid) do not idle the warp.
// (the polar grids get the register allocation of PEEL_SPH_MIN_BLOCKS resident blocks: their find_wall is bound by
// instruction issue at the 2 blocks of 256 threads that 128 registers allow, as in flight_geo.cuh)
#ifndef PEEL_MIN_BLOCKS
#define PEEL_MIN_BLOCKS 1
#endif
#ifndef PEEL_SPH_MIN_BLOCKS
#define PEEL_SPH_MIN_BLOCKS 3
#endif
template <int ND, bool POLY, int GEO>
__global__ void __launch_bounds__(PEEL_THREADS, GEO == GEO_SPH ? PEEL_SPH_MIN_BLOCKS : PEEL_MIN_BLOCKS)
peel_kernel(const ModelDev M, const ImagingDev I, const PeelJob<ND> *__restrict__ jobs,
            const uint32_t *__restrict__ n_jobs_ptr, unsigned long long *cursor, const int walls_in_smem) {
  extern __shared__ double s_walls[];
  constexpr int GG = GEO == GEO_CAR ? GEO_SPH : GEO;  // the generic traits are not used for Cartesian grids
  const int n1 = M.n1, n2 = M.n2, n3 = M.n3;
  const double *__restrict__ W = GEO == GEO_CAR ? stage_walls(M, s_walls, walls_in_smem) : nullptr;
  const uint32_t n_jobs = *n_jobs_ptr;
  const unsigned long long total = (unsigned long long)n_jobs * (unsigned)I.n_views;
  const unsigned lane = threadIdx.x & 31;
  uint32_t n_cross = 0, n_peel = 0, n_killed = 0;
  unsigned long long n_cached = 0;
  bool active = false, exhausted = false;
  uint32_t ij = 0, ip = 0;
  Stokes S{0.0, 0.0, 0.0, 0.0};
  double tau = 0.0, col[ND], tmax = 1.7976931348623157e308;
  Lane<ND> L;
  typename Geo<GG>::Ray R;
  L.ic = 0;
  for (;;) {
    // ---------------- refill ----------------
    const bool need = !active && !exhausted;
    const unsigned m_need = __ballot_sync(0xffffffffu, need);
    if (m_need) {
      const int leader = __ffs(m_need) - 1;
      unsigned long long base = 0;
      if ((int)lane == leader) base = atomicAdd(cursor, (unsigned long long)__popc(m_need));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (need) {
        const unsigned long long idx = base + __popc(m_need & ((1u << lane) - 1u));
        if (idx >= total) {
          exhausted = true;
        } else {
          // view-major order: neighbouring lanes look towards the same observer
          ip = (uint32_t)(idx / n_jobs);
          ij = (uint32_t)(idx % n_jobs);
          const PeelJob<ND> &J = jobs[ij];
          const ViewDev &V = I.views[ip];
          const ImageDev &im = I.images[V.group];
          Angle a_req = V.a;
          const bool inside = im.inside_observer != 0;
          double dist = 0.0, x_in = 0.0, y_in = 0.0;
          tmax = 1.7976931348623157e308;
          if (inside) {
            peel_inside<ND>(J, im, V.a, a_req, dist, x_in, y_in);
            tmax = dist;
          }
          const double vx = a_req.sint * a_req.cosp, vy = a_req.sint * a_req.sinp, vz = a_req.cost;
          int ix = 0, iy = 0, iz = 0, ic = 0;
          bool ok = GEO == GEO_CAR ? place_in_grid(M, J.rx, J.ry, J.rz, vx, vy, vz, ix, iy, iz, ic)
                                   : Geo<GG>::find_cell(M, J.rx, J.ry, J.rz, vx, vy, vz, ix, iy, iz, ic);
          if (ok) {
            // depth along the line of sight and image-plane coordinates (images_peeled.f90:196-211)
            const double depth = inside ? dist : -(vx * J.rx + vy * J.ry + vz * J.rz);
            double x_image = x_in, y_image = y_in;
            if (!inside) peel_image_xy<ND>(J, im, a_req, x_image, y_image);
            ok = !(depth < im.d_min || depth > im.d_max) && in_image(im, x_image, y_image);
          }
          if (ok && !im.ignore_optical_depth && M.any_sphere) {
            // grid_escape_*: a source on the line of sight kills the peel-off (grid_propagate_3d.f90:410-415)
            int hit;
            const double t_source = nearest_source(M, J.rx, J.ry, J.rz, vx, vy, vz, hit);
            ok = !(t_source < tmax);
          }
          if (ok && J.point_src > 0 && !im.ignore_optical_depth && !inside) {
            // the ray from this point source towards this view has been marched once already
            const double *rec = I.src_columns + ((size_t)(J.point_src - 1) * I.n_views + ip) * (ND + 1);
            const double nc = __ldg(rec + ND);
            if (nc >= 0.0) {
              tau = 0.0;
#pragma unroll
              for (int id = 0; id < ND; ++id) {
                col[id] = __ldg(rec + id);
                tau = tau + J.chi[id] * col[id];
              }
              n_cached += (unsigned long long)nc;
              ++n_peel;
              peel_bin<ND, POLY>(J, im, V, Stokes{1.0, 0.0, 0.0, 0.0}, tau, col, M.mono_inu);
            }
            ok = false;
          }
          if (ok) {
            S = peel_stokes<ND>(M, J, a_req);
            if (inside) {
              // flux at the observer (images_peeled.f90:207)
              const double f = 4.0 * 3.14159265358979323846 * (dist * dist);
              S.I = S.I / f; S.Q = S.Q / f; S.U = S.U / f; S.V = S.V / f;
            }
            tau = 0.0;
#pragma unroll
            for (int id = 0; id < ND; ++id) col[id] = 0.0;
            if (GEO == GEO_CAR) {
              init_lane<ND>(L, J.rx, J.ry, J.rz, vx, vy, vz, ix, iy, iz, ic, W, n1 + 1, n1 + n2 + 2);
#pragma unroll
              for (int id = 0; id < ND; ++id) L.chi[id] = J.chi[id];
            } else {
              Geo<GG>::start(M, R, J.rx, J.ry, J.rz, vx, vy, vz, ix, iy, iz, ic);
            }
            active = true;
          }
        }
      }
    }
    if (__ballot_sync(0xffffffffu, active) == 0) {
      if (__ballot_sync(0xffffffffu, !exhausted) == 0) break;
      continue;
    }
    // ---------------- march ----------------
    if (active) {
      const PeelJob<ND> &J = jobs[ij];
      const ViewDev &V = I.views[ip];
      const ImageDev &im = I.images[V.group];
      int done;  // 0 still marching, 1 reached the edge, -1 killed
      if (im.ignore_optical_depth) {
        done = 1;
      } else if (GEO == GEO_CAR) {
        done = escape_march<ND, POLY, PEEL_LOOKAHEAD>(L, W, M.rho, n1, n2, n3, tau, col, n_cross, PEEL_GROUPS, tmax) ? 1 : 0;
      } else {
        done = geo_escape<GG, ND, POLY>(M, R, J.chi, M.rho, tau, col, n_cross, 4 * PEEL_GROUPS, tmax);
      }
      if (done) {
        active = false;
        if (done > 0) {
          ++n_peel;
          peel_bin<ND, POLY>(J, im, V, S, tau, col, M.mono_inu);
        } else {
          ++n_killed;  // no wall found: the reference counts the packet as killed and drops the peel-off
        }
      }
    }
  }
  // SC_PEEL_CROSS counts what the reference would have marched; SC_PEEL_CACHED the part served by the cache
  warp_add_scalar(M.scalars + SC_PEEL_CROSS, (double)n_cross + (double)n_cached);
  warp_add_scalar(M.scalars + SC_PEEL_CACHED, (double)n_cached);
  warp_add_scalar(M.scalars + SC_PEELOFFS, (double)n_peel);
  if (GEO != GEO_CAR) warp_add_scalar(M.scalars + SC_KILLED_GEO, (double)n_killed);
}

// One thread per (point source, view): column density of every dust type from the source to the grid edge.
template <int ND, int GEO>
__global__ void source_columns_kernel(const ModelDev M, const ImagingDev I, double *__restrict__ out) {
  constexpr int GG = GEO == GEO_CAR ? GEO_SPH : GEO;
  const int total = M.n_sources * I.n_views;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
    const int is = k / I.n_views, ip = k % I.n_views;
    const SourceDev &src = M.sources[is];
    double *rec = out + (size_t)k * (ND + 1);
    double col[ND], tau = 0.0;
#pragma unroll
    for (int id = 0; id < ND; ++id) col[id] = 0.0;
    uint32_t n_cross = 0;
    bool ok = src.type == HYP_SOURCE_POINT;
    const Angle a = I.views[ip].a;
    const double vx = a.sint * a.cosp, vy = a.sint * a.sinp, vz = a.cost;
    int ix = 0, iy = 0, iz = 0, ic = 0;
    if (ok) ok = GEO == GEO_CAR ? place_in_grid(M, src.x, src.y, src.z, vx, vy, vz, ix, iy, iz, ic)
                                : Geo<GG>::find_cell(M, src.x, src.y, src.z, vx, vy, vz, ix, iy, iz, ic);
    if (ok && M.any_sphere) {
      int hit;
      nearest_source(M, src.x, src.y, src.z, vx, vy, vz, hit);
      ok = hit < 0;
    }
    if (ok) {
      double chi[ND];
#pragma unroll
      for (int id = 0; id < ND; ++id) chi[id] = 0.0;
      if (GEO == GEO_CAR) {
        Lane<ND> L;
        init_lane<ND>(L, src.x, src.y, src.z, vx, vy, vz, ix, iy, iz, ic, M.w1, M.n1 + 1, M.n1 + M.n2 + 2);
#pragma unroll
        for (int id = 0; id < ND; ++id) L.chi[id] = 0.0;
        escape_march<ND, true, PEEL_LOOKAHEAD>(L, M.w1, M.rho, M.n1, M.n2, M.n3, tau, col, n_cross);
      } else {
        typename Geo<GG>::Ray R;
        Geo<GG>::start(M, R, src.x, src.y, src.z, vx, vy, vz, ix, iy, iz, ic);
        ok = geo_escape<GG, ND, true>(M, R, chi, M.rho, tau, col, n_cross) > 0;
      }
    }
#pragma unroll
    for (int id = 0; id < ND; ++id) rec[id] = col[id];
    rec[ND] = ok ? (double)n_cross : -1.0;
  }
}

// ---------------------------------------------------------------------------------------------
// kernels of one round of the final iteration
// ---------------------------------------------------------------------------------------------
struct FinalArgs {
  void *jobs;            // PeelJob<ND>[job_capacity]
  uint32_t *n_jobs;
  uint32_t job_capacity;
  uint32_t job_margin;       // jobs the threads of an interaction kernel can add between a look at n_jobs and their append
  int32_t scattering_only;   // main.f90:274: with raytracing on, only scattered light is peeled here
  int32_t forced, algorithm; // forced first interaction (iter_final.f90:191-209)
  double baes16_xi;
  int32_t make_peeled;
  // binned images (images_binned.f90): the group escaping packets are binned into, or nullptr
  const ImageDev *binned;
  int32_t n_theta, n_phi;
  // monochromatic mode: thermal = the packets of this call come from emit_from_monochromatic_grid_pdf
  int32_t thermal;
};

constexpr uint32_t TAG_NSCAT_MASK = 0x3ffu;    // n_scat saturates at 1023

// binned_images_bin_photon (images_binned.f90:57-77) + image_bin (image_type.f90:408-524) for a packet
// that has just left the grid at path length t along its flight.
template <int ND>
__device__ inline void bin_escaped_packet(const FinalArgs &F, const Slot<ND> *__restrict__ s, const double t,
                                          const int mono_inu) {
  const ImageDev &im = *F.binned;
  const double energy = s->energy;
  if (isnan(energy)) return;
  const double vx = s->vx, vy = s->vy, vz = s->vz;
  const double rx = s->r0x + t * vx, ry = s->r0y + t * vy, rz = s->r0z + t * vz;
  const Angle a = angle_of(vx, vy, vz);
  double phi = atan2(a.sinp, a.cosp);
  if (phi < 0.0) phi = phi + 6.283185307179586476925286766559;
  const int it = ipos_bin(-1.0, 1.0, a.cost, F.n_theta);
  const int ip = ipos_bin(0.0, 6.283185307179586476925286766559, phi, F.n_phi);
  if (it < 1 || it > F.n_theta || ip < 1 || ip > F.n_phi) return;
  const int iv = F.n_phi * (it - 1) + ip - 1;  // image_id, 0-based
  const double x_image = ry * a.cosp - rx * a.sinp;
  const double y_image = rz * a.sint - ry * a.cost * a.sinp - rx * a.cost * a.cosp;
  const int inu0 = im.use_filters ? 1
                   : im.inu_min > 0 ? mono_inu - im.inu_min + 1
                                    : ipos_bin(im.log10_nu_min, im.log10_nu_max, log10(s->nu), im.n_nu);
  const int inu1 = im.use_filters ? im.n_nu : inu0;
  if (inu0 < 1 || inu0 > im.n_nu) return;
  const uint32_t tag = s->tag;
  const int io = origin_slice(im, (tag & TAG_SCATTERED) ? 1 : 0, (tag & TAG_REPROCESSED) ? 1 : 0, (int)(tag & TAG_SRC_MASK),
                              (int)(tag >> TAG_DUST_SHIFT) + 1, (int)((tag >> TAG_NSCAT_SHIFT) & TAG_NSCAT_MASK));
  const bool unc = im.uncertainties != 0;
  const size_t nn = (size_t)im.n_nu;
  const double st[4] = {1.0, s->sQ, s->sU, s->sV};
  for (int inu = inu0; inu <= inu1; ++inu) {
    double transmission = 1.0;
    if (im.use_filters) {
      const int o = im.filt_off[inu - 1];
      transmission = interp1d_fill0(im.filt_nu + o, im.filt_tr + o, im.filt_off[inu] - o, s->nu);
      if (!(transmission > 0.0)) continue;
    }
    if (im.compute_image) {
      const int ixp = ipos_bin(im.x_min, im.x_max, x_image, im.n_x), iyp = ipos_bin(im.y_min, im.y_max, y_image, im.n_y);
      if (ixp >= 1 && ixp <= im.n_x && iyp >= 1 && iyp <= im.n_y) {
        const size_t k = nn * ((ixp - 1) + (size_t)im.n_x * ((iyp - 1) + (size_t)im.n_y * (iv + (size_t)im.n_view * (io - 1))));
        const size_t stride = nn * im.n_x * im.n_y * im.n_view * im.n_orig;
        for (int is = 0; is < im.n_stokes; ++is)
          bin_add(im.img, im.img2, im.imgn, k + (inu - 1) + is * stride, st[is] * energy * transmission, unc);
      }
    }
    if (im.compute_sed) {
      const int ir = find_sed_bin(im, x_image, y_image);
      if (ir >= 1 && ir <= im.n_ap) {
        const size_t k = nn * ((ir - 1) + (size_t)im.n_ap * (iv + (size_t)im.n_view * (io - 1)));
        const size_t stride = nn * im.n_ap * im.n_view * im.n_orig;
        for (int is = 0; is < im.n_stokes; ++is)
          bin_add(im.sed, im.sed2, im.sedn, k + (inu - 1) + is * stride, st[is] * energy * transmission, unc);
      }
    }
  }
}

// Append one job per lane with pred set; whole-warp call.
template <int ND>
__device__ __forceinline__ PeelJob<ND> *job_append(bool pred, const FinalArgs &F) {
  const unsigned m = __ballot_sync(0xffffffffu, pred);
  if (m == 0) return nullptr;
  const unsigned lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if ((int)lane == leader) base = atomicAdd(F.n_jobs, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (!pred) return nullptr;
  const uint32_t k = base + __popc(m & ((1u << lane) - 1u));
  return k < F.job_capacity ? (PeelJob<ND> *)F.jobs + k : nullptr;
}

template <int ND>
__device__ __forceinline__ void fill_job(PeelJob<ND> *J, const Photon<ND> &p, int kind, double vpx, double vpy,
                                         double vpz, double sQ, double sU, double sV, int dust_id) {
  J->rx = p.r0x; J->ry = p.r0y; J->rz = p.r0z;
  J->nu = p.nu;
  J->energy = p.energy;
#pragma unroll
  for (int k = 0; k < ND; ++k) J->chi[k] = p.chi[k];
  J->vpx = vpx; J->vpy = vpy; J->vpz = vpz;
  J->sQ = sQ; J->sU = sU; J->sV = sV;
  J->emiss_var_frac = 0.0;
  J->kind = kind;
  J->source_id = (int)(p.tag & TAG_SRC_MASK);
  J->dust_id = dust_id;
  J->n_scat = (int)((p.tag >> TAG_NSCAT_SHIFT) & TAG_NSCAT_MASK);
  J->scattered = (p.tag & TAG_SCATTERED) ? 1 : 0;
  J->reprocessed = (p.tag & TAG_REPROCESSED) ? 1 : 0;
  J->emiss_type = 0;
  J->emiss_var_id = 0;
  J->point_src = 0;
  J->pad = 0;
}

// kind of the peel-off of a packet that has just been emitted: 2 from a stellar surface, 0 isotropic
template <int ND>
__device__ __forceinline__ int surface_kind(const Photon<ND> &p) {
  return (p.nx != 0.0 || p.ny != 0.0 || p.nz != 0.0) ? 2 : 0;
}

// emit_from_monochromatic_grid_pdf (grid_monochromatic.f90:120-174): dust type uniformly, cell from the cumulative
// emission probability x energy of that type, uniform position in the cell, isotropic direction; every packet of
// a type carries the same energy (mean_prob x energy_abs_tot / n_photons x n_dust, iter_final_mono.f90:187).
// Returns false for a type that does not emit at this frequency (the packet counts as run).
template <int ND>
__device__ bool emit_mono_thermal(const ModelDev &M, Photon<ND> &p, Rng &rng) {
  p.nu = M.mono_nu;
  const int id = max((int)ceil(rng.next() * (double)ND), 1) - 1;
  double w = 0.0;
#pragma unroll
  for (int k = 0; k < ND; ++k)
    if (k == id) w = M.mono_thermal_w[k];
  if (!(w > 0.0)) return false;
  const int64_t ic = sample_discrete(M.mono_cdf + (size_t)id * M.n_cells, M.n_cells, rng.next());
  random_position_cell(M, ic, rng, p.r0x, p.r0y, p.r0z);
  const Angle a = random_sphere_angle(rng);
  set_dir(p, a);
  p.nx = p.ny = p.nz = 0.0;
  p.sQ = p.sU = p.sV = 0.0;
  p.energy = p.energy0 = w;
  p.tag = TAG_REPROCESSED | ((uint32_t)id << TAG_DUST_SHIFT);
  return place_emitted<ND>(M, p);
}

// emit + the peel-off of the fresh packet (iter_final.f90:113-123)
template <int ND>
__global__ void __launch_bounds__(SERVICE_THREADS)
emit_final_kernel(const ModelDev M, Pool P, const FinalArgs F, const unsigned long long first_id,
                  const unsigned long long n_photons, const uint32_t iteration) {
  const uint32_t n = P.counts[C_NE];
  const unsigned lane = threadIdx.x & 31;
  Slot<ND> *slots = (Slot<ND> *)P.slots;
  double energy_emitted = 0.0;
  uint32_t n_run = 0;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
    const uint32_t i = base + lane;
    const bool valid = i < n;
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    const int leader = __ffs(m) - 1;
    unsigned long long k0 = 0;
    if ((int)lane == leader) k0 = atomicAdd(P.next_photon, (unsigned long long)__popc(m));
    k0 = __shfl_sync(0xffffffffu, k0, leader);
    const unsigned long long k = k0 + __popc(m & ((1u << lane) - 1u));
    bool go = valid && k < n_photons;
    uint32_t slot = 0;
    Photon<ND> p;
    Rng rng;
    unsigned long long id = 0;
    if (go) {
      slot = P.q_emit[i];
      id = first_id + (k & ~(unsigned long long)(P.window - 1)) + P.perm[k & (2ull * P.window - 1)];
      rng.init(M.seed, id, iteration);
      ++n_run;
      go = F.thermal ? emit_mono_thermal<ND>(M, p, rng) : emit_photon<ND>(M, p, rng, energy_emitted);
    }
    PeelJob<ND> *J = job_append<ND>(go && F.make_peeled && !F.scattering_only, F);
    if (J) {
      if (F.thermal) {
        fill_job<ND>(J, p, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, (int)(p.tag >> TAG_DUST_SHIFT) + 1);
      } else {
        fill_job<ND>(J, p, surface_kind(p), p.nx, p.ny, p.nz, 0.0, 0.0, 0.0, 0);
        if (J->kind == 0 && M.sources[J->source_id - 1].type == HYP_SOURCE_POINT) J->point_src = J->source_id;
      }
    }
    if (go) {
      // with a forced first interaction the optical depth is drawn by the flight kernel once the
      // optical depth to the grid edge is known
      p.tau_left = F.forced ? -1.0 : -log(1.0 - rng.next());
      store_photon<ND>(slots + slot, p, rng, id);
    }
    queue_append(go, P.q_beam, P.counts + C_NB, slot);
  }
  warp_add_scalar(M.scalars + SC_ENERGY, energy_emitted);
  warp_add_scalar(M.scalars + SC_PHOTONS, (double)n_run);
}

// interact + the peel-off of the surviving packet (iter_final.f90:247-269)
template <int ND>
__global__ void __launch_bounds__(SERVICE_THREADS)
interact_final_kernel(const ModelDev M, Pool P, const FinalArgs F, uint32_t *__restrict__ q_flight_next,
                      uint32_t *n_flight_next, const uint32_t iteration) {
  const uint32_t n = P.counts[C_NI];
  const unsigned lane = threadIdx.x & 31;
  Slot<ND> *slots = (Slot<ND> *)P.slots;
  uint32_t n_abs = 0, n_scat = 0, n_kill = 0;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
    const uint32_t i = base + lane;
    const bool valid = i < n;
    uint32_t slot = 0;
    bool alive = false, peel = false, scattered = false, reemitted = false, resume = false, paused = false;
    Photon<ND> p;
    Rng rng;
    double vpx = 0, vpy = 0, vpz = 0, sQ = 0, sU = 0, sV = 0;
    int dust_id = 0;
    uint64_t id = 0;
    if (valid) {
      slot = P.q_interact[i];
      load_photon<ND>(slots + slot, p, rng, M.seed, iteration);
      id = slots[slot].id;
      vpx = p.vx; vpy = p.vy; vpz = p.vz;
      sQ = p.sQ; sU = p.sU; sV = p.sV;
      resume = (p.tag & TAG_MRW_PAUSED) != 0u;
      if (resume) {
        // no interaction: the random walk this packet was in goes on where it stopped
        p.tag &= ~TAG_MRW_PAUSED;
        alive = true;
      } else
      if (p.t < 0.0) {
        // re-absorbed by a star: re-emit from its surface; always peeled (iter_final.f90:219-227)
        if (reemit_photon<ND>(M, p, rng, n_kill)) {
          alive = true;
          reemitted = true;
          peel = F.make_peeled != 0;
        }
      } else if (interact_photon<ND>(M, p, rng, n_abs, n_scat, n_kill, dust_id, scattered, M.mono_nu > 0.0) == 0) {
        alive = true;
        if (scattered) {
          uint32_t ns = (p.tag >> TAG_NSCAT_SHIFT) & TAG_NSCAT_MASK;
          if (ns < TAG_NSCAT_MASK) ++ns;
          p.tag = (p.tag & TAG_LOW_MASK) | (ns << TAG_NSCAT_SHIFT) | TAG_SCATTERED;
        } else {
          p.tag = (p.tag & ~TAG_SCATTERED & ~(3u << TAG_DUST_SHIFT)) | TAG_REPROCESSED;
        }
        p.tag = (p.tag & ~(3u << TAG_DUST_SHIFT)) | ((uint32_t)dust_id << TAG_DUST_SHIFT);  // p%dust_id (dust_interact.f90:62,68)
        peel = F.make_peeled && (scattered || !F.scattering_only);
      }
    }
    if (alive && M.use_mrw && !reemitted) {
      // grid_do_mrw_noenergy + peel-off of every random-walk step (iter_final.f90:166-185)
      bool out = false;
      int64_t step = resume ? (int64_t)p.energy0 : 0;   // (energy0 is free outside the monochromatic mode)
      for (; step < M.n_mrw_max; ++step) {
        const double R0 = distance_to_closest_wall<ND>(M, p);
        if (!(M.alpha_inv_planck[p.ic] * R0 > M.mrw_gamma)) {
          out = true;
          break;
        }
        // every step queues up to two peel-offs: when the queue of this round is nearly full the walk is interrupted
        // and resumed next round (the packet takes a flight of optical depth zero in between)
        if (*(volatile uint32_t *)F.n_jobs + F.job_margin > F.job_capacity) {
          paused = true;
          break;
        }
        // the interaction's own peel-off has to be queued before the state changes
        if (peel) {
          const uint32_t k = atomicAdd(F.n_jobs, 1u);
          if (k < F.job_capacity) fill_job<ND>((PeelJob<ND> *)F.jobs + k, p, scattered ? 1 : 0, vpx, vpy, vpz, sQ, sU, sV, dust_id + 1);
          else atomicMax(M.error_flag, ERR_JOBS);
          peel = false;
        }
        const int id = mrw_step<ND>(M, p, rng, false, R0);
        if (F.make_peeled && !F.scattering_only) {
          const uint32_t k = atomicAdd(F.n_jobs, 1u);
          if (k < F.job_capacity) fill_job<ND>((PeelJob<ND> *)F.jobs + k, p, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, id + 1);
          else atomicMax(M.error_flag, ERR_JOBS);
        }
      }
      if (paused) {
        p.energy0 = (double)step;
        p.tag |= TAG_MRW_PAUSED;
      } else if (!out) {
        ++n_kill;
        alive = false;
      }
    }
    PeelJob<ND> *J = job_append<ND>(peel, F);
    if (J) {
      if (reemitted)
        fill_job<ND>(J, p, surface_kind(p), p.nx, p.ny, p.nz, 0.0, 0.0, 0.0, 0);
      else
        fill_job<ND>(J, p, scattered ? 1 : 0, vpx, vpy, vpz, sQ, sU, sV, dust_id + 1);
    }
    if (alive) {
      p.tau_left = paused ? 0.0 : -log(1.0 - rng.next());
      store_photon<ND>(slots + slot, p, rng, id);
    }
    queue_append(alive, q_flight_next, n_flight_next, slot);
    queue_append(valid && !alive, P.q_emit, P.counts + C_NE, slot);
  }
  warp_add_scalar(M.scalars + SC_ABS, (double)n_abs);
  warp_add_scalar(M.scalars + SC_SCAT, (double)n_scat);
  warp_add_scalar(M.scalars + SC_KILLED_INT, (double)n_kill);
}

// grid_integrate_noenergy for every queued packet.  A packet on its first flight with a forced first
// interaction first has its optical depth to the grid edge measured, then draws tau from the truncated
// exponential and carries the weight (forced_interaction.f90:23-133).
template <int ND, int D>
__global__ void __launch_bounds__(FLIGHT_THREADS, FLIGHT_MIN_BLOCKS)
flight_final_kernel(const ModelDev M, Pool P, const FinalArgs F, const uint32_t *__restrict__ q_flight,
                    const uint32_t *n_flight_ptr, uint32_t *cursor, const int walls_in_smem, const uint32_t iteration) {
  extern __shared__ double s_walls[];
  const int n1 = M.n1, n2 = M.n2, n3 = M.n3;
  const double *__restrict__ W = stage_walls(M, s_walls, walls_in_smem);
  const uint32_t n_flight = *n_flight_ptr;
  Slot<ND> *slots = (Slot<ND> *)P.slots;
  CellRec *__restrict__ cells = M.cells;
  const unsigned lane = threadIdx.x & 31;
  bool active = false, exhausted = false;
  uint32_t slot = 0;
  Lane<ND> L;
  L.ic = 0;
  uint32_t n_cross = 0, n_esc = 0, n_peel_cross = 0;
  unsigned long long cross_hi = 0;
  for (;;) {
    const bool need = !active && !exhausted;
    const unsigned m_need = __ballot_sync(0xffffffffu, need);
    int fin = 0;
    if (m_need) {
      const int leader = __ffs(m_need) - 1;
      uint32_t base = 0;
      if ((int)lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(m_need));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (need) {
        const uint32_t idx = base + __popc(m_need & ((1u << lane) - 1u));
        if (idx >= n_flight) {
          exhausted = true;
        } else {
          slot = q_flight[idx];
          load_lane<ND>(slots + slot, L, W, n1 + 1, n1 + n2 + 2);
          active = true;
          if (lane_outside(L.ix, L.iy, L.iz, n1, n2, n3)) {
            fin = 1;  // emitted on the outer wall moving outwards (escaped_cell before the first step)
          } else if (L.tau < 0.0) {
            Lane<ND> E = L;
            double tau_escape = 0.0, col[ND];
            escape_march<ND, false, D>(E, W, M.rho, n1, n2, n3, tau_escape, col, n_peel_cross);
            Slot<ND> *s = slots + slot;
            Rng rng;
            rng.init(M.seed, s->id, iteration);
            rng.blk = s->rng_blk;
            rng.has_spare = s->rng_has_spare != 0;
            rng.spare = s->rng_spare;
            double tau, weight = 1.0;
            if (tau_escape > 1.e-10) {
              const double TAU_THRES = 1.e-7;
              const double one_minus_exp = tau_escape > TAU_THRES ? 1.0 - exp(-tau_escape) : tau_escape;
              if (F.algorithm == HYP_FFI_BAES16) {
                const double alpha = (1.0 - F.baes16_xi) / one_minus_exp, beta = F.baes16_xi / tau_escape;
                double tau_min = 0.0, tau_max = tau_escape;
                const double xi = rng.next();
                for (int it = 0; it < 60; ++it) {
                  tau = 0.5 * (tau_min + tau_max);
                  const double xt = tau > TAU_THRES ? alpha * (1.0 - exp(-tau)) + beta * tau : alpha * tau + beta * tau;
                  if (xt > xi) tau_max = tau; else tau_min = tau;
                }
                tau = 0.5 * (tau_min + tau_max);
                weight = 1.0 / (alpha + beta * exp(tau));
              } else {
                tau = -log(1.0 - rng.next() * one_minus_exp);
                weight = one_minus_exp;
              }
              s->energy = s->energy * weight;
            } else {
              tau = -log(1.0 - rng.next());
            }
            L.tau = tau;
            s->rng_blk = rng.blk;
            s->rng_has_spare = rng.has_spare ? 1u : 0u;
            s->rng_spare = rng.spare;
          }
        }
      }
    }
    if (__ballot_sync(0xffffffffu, active) == 0) break;
    if (active && fin == 0) {
#pragma unroll 1
      for (int g = 0; g < FLIGHT_GROUPS; ++g) {
        fin = advance_group<ND, D, false, false>(L, true, W, cells, n1, n2, n3, n_cross, M.rho);
        if (fin) break;
      }
      if (n_cross > 0x7fffff00u) {
        cross_hi += n_cross;
        n_cross = 0;
      }
    }
    if (fin == 2) store_flight_result<ND>(slots + slot, L);
    if (fin == 1 && F.binned) bin_escaped_packet<ND>(F, slots + slot, L.t, M.mono_inu);   // iter_final.f90:126-129
    queue_append(fin == 2, P.q_interact, P.counts + C_NI, slot);
    queue_append(fin == 1, P.q_emit, P.counts + C_NE, slot);
    if (fin) {
      n_esc += fin == 1 ? 1u : 0u;
      active = false;
    }
  }
  warp_add_scalar(M.scalars + SC_CROSS, (double)(cross_hi + n_cross));
  warp_add_scalar(M.scalars + SC_ESC, (double)n_esc);
  warp_add_scalar(M.scalars + SC_PEEL_CROSS, (double)n_peel_cross);
}

// ---------------------------------------------------------------------------------------------
// raytracing iteration (iter_raytracing.f90:31-141): packets that are only peeled off
// ---------------------------------------------------------------------------------------------
// energy_abs_tot(d) = sum over cells of E * rho * V (update_energy_abs_tot, grid_physics_3d.f90:605-611)
__global__ void energy_abs_tot_kernel(const ModelDev M, double *__restrict__ out) {
  const int nd = M.n_dust;
  const int64_t n = M.n_cells * nd;
  double acc[MAX_DUST] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const int id = (int)(k % nd);
    const int64_t ic = k / nd;
    const double vol = cell_volume(M, ic);
    const double v = M.specific_energy[k] * M.cells[k].rho * vol;
#pragma unroll
    for (int d = 0; d < MAX_DUST; ++d) acc[d] += d == id ? v : 0.0;
  }
#pragma unroll
  for (int d = 0; d < MAX_DUST; ++d) warp_add_scalar(out + d, acc[d]);
}

// setup_monochromatic_grid_pdfs (grid_monochromatic.f90:50-118): emission probability at the run's frequency x
// energy of every cell that holds physical quantities, per dust type; w is [n_dust][n_cells], zeroed by the caller.
// list = the cells of geo%mask_map (octree leaves, valid AMR cells) or nullptr for all cells.
__global__ void mono_weights_kernel(const ModelDev M, const int32_t *__restrict__ list, const int64_t n_list,
                                    double *__restrict__ w) {
  const int nd = M.n_dust;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n_list; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ic = list ? (int64_t)list[k] : k;
    const double vol = cell_volume(M, ic);
    for (int id = 0; id < nd; ++id) {
      const size_t q = (size_t)ic * nd + id;
      const double e = M.specific_energy[q] * M.cells[q].rho * vol;
      w[(size_t)id * M.n_cells + ic] = e > 0.0 ? mono_emit_probability(M, id, M.jnu_id[q], M.jnu_frac[q]) * e : 0.0;
    }
  }
}

__global__ void mono_divide_kernel(double *__restrict__ cdf, const int64_t n, const double total) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    cdf[k] = cdf[k] / total;
}

constexpr uint32_t ITER_MONO = 0x7f000000u;   // + 2 * inu (+ 1 for the thermal packets)
constexpr uint32_t ITER_FINAL = 0x7fffff00u, ITER_RAY_SOURCE = 0x7fffff01u, ITER_RAY_DUST = 0x7fffff02u;

// jobs [0, n_src) come from the sources, jobs [n_src, n_src + n_thermal) from random cells
template <int ND>
__global__ void raytrace_emit_kernel(const ModelDev M, PeelJob<ND> *__restrict__ jobs, uint32_t *n_jobs,
                                     const unsigned long long first_source_id, const uint32_t n_src,
                                     const double source_weight, const unsigned long long first_dust_id,
                                     const uint32_t n_thermal, const double dust_weight,
                                     const double *__restrict__ energy_abs_tot) {
  const uint32_t total = n_src + n_thermal;
  double dummy = 0.0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    Photon<ND> p;
    Rng rng;
    PeelJob<ND> J;
    bool ok = true;
    if (i < n_src) {
      rng.init(M.seed, first_source_id + i, ITER_RAY_SOURCE);
      ok = emit_photon<ND>(M, p, rng, dummy);
      p.energy = p.energy * source_weight;  // energy_total / n_photons_sources (iter_raytracing.f90:79)
      fill_job<ND>(&J, p, surface_kind(p), p.nx, p.ny, p.nz, 0.0, 0.0, 0.0, 0);
      J.emiss_type = M.sources[(p.tag & TAG_SRC_MASK) - 1].freq_type;
      if (ok && J.emiss_type == HYP_SPECTRUM_LTE) {
        // LTE map source: the spectrum is the emissivity of the dust type emit_photon picked in the cell
        const int id = (int)(p.tag >> TAG_DUST_SHIFT);
        const size_t k = (size_t)p.ic * ND + id;
        J.dust_id = id + 1;
        J.emiss_var_id = M.jnu_id[k];
        J.emiss_var_frac = M.jnu_frac[k];
      }
      if (J.kind == 0 && M.sources[J.source_id - 1].type == HYP_SOURCE_POINT) J.point_src = J.source_id;
    } else {
      // emit_from_grid (grid_physics_3d.f90:691-753)
      rng.init(M.seed, first_dust_id + (i - n_src), ITER_RAY_DUST);
      const int id = max((int)ceil(rng.next() * (double)ND), 1) - 1;
      // random_masked_cell (grid_geometry_common_3d.f90:104-115): octrees draw from their leaves
      int64_t ic;
      double n_masked = (double)M.n_cells;
      if (M.grid_type == GEO_OCT) {
        n_masked = (double)M.oct.n_leaves;
        ic = M.oct.leaves[max((int64_t)ceil(rng.next() * n_masked), (int64_t)1) - 1];
      } else if (M.grid_type == GEO_AMR) {
        n_masked = (double)M.amr.n_valid;
        ic = M.amr.valid[max((int64_t)ceil(rng.next() * n_masked), (int64_t)1) - 1];
      } else if (M.grid_type == GEO_VOR) {
        n_masked = (double)M.vor.n_valid;
        ic = M.vor.valid[max((int64_t)ceil(rng.next() * n_masked), (int64_t)1) - 1];
      } else {
        ic = max((int64_t)ceil(rng.next() * (double)M.n_cells), (int64_t)1) - 1;
      }
      random_position_cell(M, ic, rng, p.r0x, p.r0y, p.r0z);
      const size_t k = (size_t)ic * ND + id;
      const double vol = cell_volume(M, ic);
      const double etot = energy_abs_tot[id];
      p.energy = 0.0;
      if (etot > 0.0) p.energy = M.specific_energy[k] * (M.cells[k].rho * vol) * n_masked / etot;
      ok = p.energy > 0.0;
      // energy_abs_tot(dust) / n_photons_thermal * n_dust (iter_raytracing.f90:113)
      p.energy = p.energy * etot * dust_weight;
      p.nu = 0.0;
#pragma unroll
      for (int q = 0; q < ND; ++q) p.chi[q] = 0.0;
      p.tag = TAG_REPROCESSED;
      fill_job<ND>(&J, p, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, id + 1);
      J.emiss_type = 3;
      J.emiss_var_id = M.jnu_id[k];
      J.emiss_var_frac = M.jnu_frac[k];
    }
    if (ok) jobs[atomicAdd(n_jobs, 1u)] = J;
  }
}

// image_scale (image_type.f90:136-151): which = 0 values (x scale), 1 sums of squares (x scale^2)
__global__ void image_scale_kernel(double *__restrict__ a, int64_t n, double scale) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    a[k] = a[k] * scale;
}
