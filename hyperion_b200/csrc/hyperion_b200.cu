// hyperion_b200.cu -- CUDA kernels and C ABI of the B200-native photon-packet engine.
//
// Hot path: the Lucy iteration of the reference (do_lucy, src/main/iter_lucy.f90:66-237).  A pool
// of photon packets advances in rounds of three kernels (emit -> flight -> interact); the flight
// kernel (grid_integrate) is a persistent-threads loop that refills idle lanes from a queue.
//
// HBM layout: one 16-byte record {density, energy_sum} per (cell, dust) so the
// density read and the deposit RED of a crossing touch the same 32-byte sector.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "device.cuh"
#include "geometry_sph.cuh"
#include "geometry_oct.cuh"
#include "geometry_amr.cuh"
#include "geometry_vor.cuh"

using namespace hyp;

// =============================================================================================
// device-side model description
// =============================================================================================
struct CellRec {
  double rho;
  double esum;
};

enum { SC_ENERGY = 0, SC_KILLED_GEO, SC_KILLED_INT, SC_CROSS, SC_ABS, SC_SCAT, SC_ESC, SC_PHOTONS, SC_PEEL_CROSS, SC_PEELOFFS,
       SC_PEEL_CACHED, SC_COUNT };

enum { GEO_CAR = 0, GEO_SPH = 1, GEO_OCT = 2, GEO_AMR = 3, GEO_VOR = 4 };  // GEO_SPH covers both polar grids (SphGrid::kind)

struct ModelDev {
  int32_t grid_type;        // GEO_*
  SphGrid sph;              // spherical polar tables (grid_type == GEO_SPH)
  OctGrid oct;              // octree (grid_type == GEO_OCT)
  AmrGrid amr;              // block-structured AMR (grid_type == GEO_AMR)
  VorGrid vor;              // Voronoi mesh (grid_type == GEO_VOR)
  int32_t n1, n2, n3, n_dust, n_sources;
  int64_t n_cells;
  const double *w1, *w2, *w3;
  CellRec *cells;           // [n_cells][n_dust]
  double *rho;              // [n_cells][n_dust] densities alone, for the marches that deposit nothing (peel-off,
                            // imaging flights): 4 cells per 32-byte sector instead of 2
  double *specific_energy;  // [n_cells][n_dust]
  int32_t *jnu_id;          // [n_cells][n_dust]
  double *jnu_frac;         // [n_cells][n_dust]
  DustDev dust[MAX_DUST];
  const SourceDev *sources;
  const SpectrumDev *spectra;
  double min_energy[MAX_DUST];
  const double *energy_additional;  // specific_energy_type = 'additional': added after every iteration, or nullptr
  // run configuration
  uint64_t seed;
  int64_t n_inter_max;
  int32_t kill_on_absorb, kill_on_scatter, sample_evenly, enforce_energy_range;
  // modified random walk (grid_mrw_3d.f90): per-cell alpha_inv_planck and diffusion coefficient, P(y) table
  int32_t use_mrw;
  int64_t n_mrw_max;
  double mrw_gamma;
  double *alpha_inv_planck, *diff_coeff;
  const double *mrw_cdf;    // xcdf[100] | ycdf[100]
  const double *coll_xyz, *coll_cdf;  // point collections: positions [n][3] and cumulative luminosities, all sources concatenated
  const double *map_cdf;              // map sources: cumulative luminosity per cell (n_cells entries per source)
  const SpotDev *spots;               // spots of all spherical sources
  int32_t any_sphere;       // a spherical source exists: flights test for re-absorption (source.f90:206-227)
  int64_t n_reabs_max;
  // monochromatic final iteration (iter_final_mono.f90): frequency of the run (0: polychromatic), its 1-based
  // index, the energy per source packet of every source at that frequency, the kill threshold
  double mono_nu;
  int32_t mono_inu;
  const double *mono_src_w;            // [n_sources]
  const double *mono_spot_w;           // [spots of all sources]
  double mono_threshold;
  double mono_lte_w;                   // energy_total / n_photons for sources with an LTE spectrum
  const double *mono_logp[MAX_DUST];   // log10 of the emission probability per unit frequency of every emissivity state
  const double *mono_cdf;              // [n_dust][n_cells] cumulative emission probability x energy of the cells
  double mono_thermal_w[MAX_DUST];     // energy of a thermal packet of each dust type (0: the type does not emit)
  // per-cell packet counter (n_photons / last_photon_id, grid_physics_3d.f90:38-39), or nullptr
  unsigned long long *n_visits, *last_id;
  // frequency-resolved specific energy (grid_physics_3d.f90:41-56), or nullptr: deposit sums and specific energies as
  // n_spec_bins planes in device order [cell][dust]; log10 of the n_spec_bins + 1 bin edges; the fraction of every
  // emissivity state's spectrum per bin [dust][spec_jmax][bin] (setup_j_nu_bin_fractions, :325-348)
  double *spec_sums, *spec_energy;
  const double *spec_log_edges, *spec_jfrac;
  int32_t n_spec_bins, spec_jmax;
  // outputs
  double *scalars;                 // [SC_COUNT], directly after the reduced sum grid
  unsigned long long *work_counter;
  int32_t *error_flag;
};

// locate(log_nu_bin_edges, log10(nu)) (grid_propagate_3d.f90:71), 0-based; -1 outside the outer edges
__device__ __forceinline__ int spectrum_bin(const ModelDev &M, double nu) {
  const double lg = log10(nu);
  const int nb = M.n_spec_bins;
  const double *e = M.spec_log_edges;
  if (lg < __ldg(e) || lg > __ldg(e + nb)) return -1;
  int lo = 0, hi = nb;          // e[lo] <= lg, lg <= e[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (lg >= __ldg(e + mid)) lo = mid; else hi = mid;
  }
  return lo;
}

enum { ERR_NONE = 0, ERR_NOT_IN_CELL = 1, ERR_NU_RANGE = 2, ERR_SCATTER = 3, ERR_JOBS = 4, ERR_DEPOSIT = 5 };

// =============================================================================================
// photon pool (wavefront formulation of do_lucy's photon loop, src/main/iter_lucy.f90:119-209)
//
// The reference runs one photon at a time through emit -> [flight -> interact]* .  Here a pool of
// packets advances in rounds of three kernels:
//   emit_kernel      fills free slots with new packets          (emit, src/sources/source.f90:100-179)
//   flight_kernel    marches every queued packet to its next event, depositing energy
//                    (grid_integrate, src/grid/grid_propagate_3d.f90:35-234)   <- the HBM-bound part
//   interact_kernel  absorbs/re-emits or scatters                (interact, src/dust/dust_interact.f90:22-79)
// so that the crossing loop runs converged, with few registers (high occupancy) and several
// density loads in flight per lane, while the branchy table sampling runs in its own kernels.
// =============================================================================================
#ifndef SLOT_ALIGN
#define SLOT_ALIGN 32
#endif
template <int ND>
struct alignas(SLOT_ALIGN) Slot {
  // ---- hot part: what a flight needs
  double r0x, r0y, r0z;  // flight origin
  double vx, vy, vz;
  double chi[ND], kE[ND];
  // what a flight changes: for one dust type exactly the third 32-byte sector of the slot, written whole
  double tau_left;
  double t;              // out: path length from the origin at which the flight ended
  int32_t ix, iy, iz;    // in: cell whose walls bound the flight; out: cell of the event
  int32_t ic;            // 1-D id used for density / deposits (p%icell%ic of the reference)
  // ---- cold part: only the emit / interact kernels touch it
  double nu, energy;
  double sQ, sU, sV;     // Stokes (I = 1)
  double albedo[ND];
  double rng_spare;
  double energy0;        // monochromatic mode: energy at emission (packets die below threshold x energy0)
  uint64_t id;
  uint32_t rng_blk;
  uint32_t rng_has_spare;  // bit 0: the stream holds a spare number; bits 1-31: successive re-absorptions by sources
  uint32_t n_inter;
  uint32_t tag;          // final iteration: source id | scattered | reprocessed | n_scat (see imaging.cuh)
};

enum { C_NF0 = 0, C_NF1, C_NB, C_NI, C_NE, C_CURSOR, C_CURSOR_B, C_COUNT = 16 };

struct Pool {
  void *slots;
  uint32_t capacity;
  uint32_t *q_flight[2];  // slot ids with a flight pending after an interaction (double-buffered across rounds)
  uint32_t *q_beam;       // slot ids of freshly emitted packets, in emission (direction-sorted) order
  uint32_t *q_interact;   // slot ids that reached an interaction
  uint32_t *q_emit;       // free slot ids
  uint32_t *counts;       // [C_COUNT]
  unsigned long long *next_photon;
  // emission order: packets are emitted sorted by direction so that the flights running at the
  // same time cross the same cells (L2 hits, coalesced loads).  perm is a ring of two windows of
  // `window` packet offsets each; window w of the launch lives in half (w & 1).
  const uint32_t *perm;
  uint32_t window;        // power of two
};

#ifndef WAVE_DEFER_PARTIAL
#define WAVE_DEFER_PARTIAL 1
#endif

// Register view of one packet in the emit / interact kernels.
template <int ND>
struct Photon {
  double r0x, r0y, r0z;
  double vx, vy, vz;
  double t;
  double tau_left;
  double chi[ND], kE[ND], albedo[ND];
  double nu, energy;
  double energy0 = 0.0;
  double sQ, sU, sV;
  int32_t ix, iy, iz, ic;
  uint32_t n_inter;
  uint32_t n_reabs = 0;  // successive re-absorptions by sources so far (iter_lucy.f90:158-185)
  uint32_t tag;
  double nx, ny, nz;     // outward normal at the emission point of a spherical source (0 otherwise); not stored
};

template <int ND>
__device__ __forceinline__ void load_photon(const Slot<ND> *__restrict__ s, Photon<ND> &p, Rng &rng, uint64_t seed,
                                            uint32_t iteration) {
  p.r0x = s->r0x; p.r0y = s->r0y; p.r0z = s->r0z;
  p.vx = s->vx; p.vy = s->vy; p.vz = s->vz;
  p.t = s->t;
  p.tau_left = s->tau_left;
#pragma unroll
  for (int k = 0; k < ND; ++k) {
    p.chi[k] = s->chi[k];
    p.kE[k] = s->kE[k];
    p.albedo[k] = s->albedo[k];
  }
  p.nu = s->nu; p.energy = s->energy; p.energy0 = s->energy0;
  p.sQ = s->sQ; p.sU = s->sU; p.sV = s->sV;
  p.ix = s->ix; p.iy = s->iy; p.iz = s->iz; p.ic = s->ic;
  p.n_inter = s->n_inter;
  p.tag = s->tag;
  rng.init(seed, s->id, iteration);
  rng.blk = s->rng_blk;
  const uint32_t hs = s->rng_has_spare;
  rng.has_spare = (hs & 1u) != 0;
  p.n_reabs = hs >> 1;
  rng.spare = s->rng_spare;
}

template <int ND>
__device__ __forceinline__ void store_photon(Slot<ND> *__restrict__ s, const Photon<ND> &p, const Rng &rng, uint64_t id) {
  s->r0x = p.r0x; s->r0y = p.r0y; s->r0z = p.r0z;
  s->vx = p.vx; s->vy = p.vy; s->vz = p.vz;
  s->tau_left = p.tau_left;
  s->t = 0.0;
#pragma unroll
  for (int k = 0; k < ND; ++k) {
    s->chi[k] = p.chi[k];
    s->kE[k] = p.kE[k];
    s->albedo[k] = p.albedo[k];
  }
  s->ix = p.ix; s->iy = p.iy; s->iz = p.iz; s->ic = p.ic;
  s->nu = p.nu; s->energy = p.energy; s->energy0 = p.energy0;
  s->sQ = p.sQ; s->sU = p.sU; s->sV = p.sV;
  s->rng_spare = rng.spare;
  s->id = id;
  s->rng_blk = rng.blk;
  s->rng_has_spare = (rng.has_spare ? 1u : 0u) | (p.n_reabs << 1);
  s->n_inter = p.n_inter;
  s->tag = p.tag;
}

// Append `value` to a queue for every lane with pred set; must be called by the whole warp.
__device__ __forceinline__ void queue_append(bool pred, uint32_t *__restrict__ q, uint32_t *count, uint32_t value) {
  const unsigned m = __ballot_sync(0xffffffffu, pred);
  if (m == 0) return;
  const unsigned lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if ((int)lane == leader) base = atomicAdd(count, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (pred) q[base + __popc(m & ((1u << lane) - 1u))] = value;
}

// Sum a per-lane value over the warp and add it to a global scalar.
__device__ __forceinline__ void warp_add_scalar(double *dst, double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(dst, v);
}

// update_optconsts (src/dust/dust.f90:64-79)
template <int ND>
__device__ __forceinline__ bool update_optconsts(const ModelDev &M, Photon<ND> &p) {
#pragma unroll
  for (int id = 0; id < ND; ++id) {
    const DustDev &d = M.dust[id];
    if (p.nu < d.L.nu_min || p.nu > d.L.nu_max) return false;
    const double *nu = d.B + d.L.o_nu;
    int j = lower_interval(nu, d.L.n_nu, p.nu);
    const double *lognu = d.B + d.L.o_lognu;
    double l0 = __ldg(lognu + j), l1 = __ldg(lognu + j + 1);
    double frac = (log10(p.nu) - l0) / (l1 - l0);
    double chi = loglog_at(d.B + d.L.o_logchi, j, frac);
    double alb = loglog_at(d.B + d.L.o_logalb, j, frac);
    p.chi[id] = chi;
    p.albedo[id] = alb;
    p.kE[id] = chi * (1.0 - alb) * p.energy;
  }
  return true;
}

// find_cell + adjust_wall (src/grid/grid_geometry_cartesian_3d.f90:143-259) for one axis.
// Returns false if outside the grid.  i is the cell whose walls bound the flight; i_found is the
// cell find_cell reports (the reference keeps using its 1-D id until the first wall crossing).
__device__ __forceinline__ bool place_axis(const double *__restrict__ w, int n, double r, double v, int &i,
                                           int &i_found) {
  if (!(r >= __ldg(w) && r <= __ldg(w + n))) return false;
  int j = lower_interval(w, n + 1, r);  // w[j] <= r, j in [0, n-1]; r == w[n] gives n-1
  i_found = j;
  i = j;
  if (v > 0.0) {
    if (r == __ldg(w + j + 1)) i = j + 1;
  } else if (v < 0.0) {
    if (r == __ldg(w + j)) i = j - 1;
  }
  return true;
}

template <int ND>
__device__ __forceinline__ void angle_from_dir(const Photon<ND> &p, Angle &a) {
  a.cost = p.vz;
  double s2 = p.vx * p.vx + p.vy * p.vy;
  a.sint = sqrt(s2);
  if (a.sint > 0.0) {
    a.cosp = p.vx / a.sint;
    a.sinp = p.vy / a.sint;
  } else {
    a.cosp = 1.0;
    a.sinp = 0.0;
  }
}

template <int ND>
__device__ __forceinline__ void set_dir(Photon<ND> &p, const Angle &a) {
  // angle3d_to_vector3d (type_vector3d.f90:301-317)
  p.vx = a.sint * a.cosp;
  p.vy = a.sint * a.sinp;
  p.vz = a.cost;
}

// ran_mu_limb(1.5, 1) (source_type.f90:982-1084): mu from P = 1.5 mu^2 + mu, by the real root of the cubic
__device__ inline double ran_mu_limb(Rng &rng) {
  double s1 = 1.5 * (1.0 / 3.0), t1 = 1.0 * (1.0 / 2.0);
  const double norm = s1 + t1;
  s1 = s1 / norm;
  t1 = t1 / norm;
  const double xi = -rng.next();
  const double bb = t1 / s1, dd = xi / s1;
  const double alpha = 1.0 / 3.0, gamma = 1.0 / 27.0;
  const double pp = -bb * bb * alpha * alpha;
  const double q = (dd + 2.0 * bb * bb * bb * gamma) * 0.5;
  const double p3 = pp * pp * pp;
  double delta = q * q + p3;
  if (delta < 0.0) {
    const double phi = acos(-q / sqrt(fabs(p3)));
    return 2.0 * sqrt(fabs(pp)) * cos(phi * alpha) - bb * alpha;
  }
  delta = sqrt(delta);
  const double u = -q + delta, v = -q - delta;
  const double cu = u >= 0.0 ? pow(u, alpha) : -pow(fabs(u), alpha);
  const double cv = v >= 0.0 ? pow(v, alpha) : -pow(fabs(v), alpha);
  return cu + cv - bb * alpha;
}

// source_distance + find_nearest_source (source_type.f90:324-357, source.f90:206-227): path length to the
// nearest spherical source along (r, v), +inf if none is hit; `which` = its 0-based index
__device__ inline double nearest_source(const ModelDev &M, double rx, double ry, double rz, double vx, double vy,
                                        double vz, int &which) {
  double nearest = __longlong_as_double(0x7ff0000000000000LL);
  which = -1;
  if (!M.any_sphere) return nearest;
  for (int is = 0; is < M.n_sources; ++is) {
    const SourceDev &S = M.sources[is];
    if (S.type != HYP_SOURCE_SPHERE) continue;
    const double dx = rx - S.x, dy = ry - S.y, dz = rz - S.z;
    const double pB = 2.0 * (dx * vx + dy * vy + dz * vz);
    const double pC = (dx * dx + dy * dy + dz * dz) - S.radius * S.radius;
    double t1, t2;
    quad_pascal_reduced(pB, pC, t1, t2);
    const double tol = (double)1.e-8f * S.radius;
    double d = __longlong_as_double(0x7ff0000000000000LL);
    if (t1 < d && t1 > tol) d = t1;
    if (t2 < d && t2 > tol) d = t2;
    if (d < nearest) {
      nearest = d;
      which = is;
    }
  }
  return nearest;
}

__device__ inline void random_position_cell(const ModelDev &M, int64_t ic, Rng &rng, double &x, double &y, double &z);

// sample_pdf_discrete_dp on the luminosities (type_pdf.f90:313-337): first source with cdf >= xi, by bisection
__device__ __forceinline__ int pick_source(const ModelDev &M, double xi) {
  const int ns = M.n_sources;
  if (xi <= M.sources[0].cdf) return 0;
  if (xi >= M.sources[ns - 1].cdf) return ns - 1;
  int jmin = 1, jmax = ns;
  for (;;) {
    const int j = (jmax + jmin) / 2;
    if (xi > M.sources[j - 1].cdf) jmin = j; else jmax = j;
    if (jmax == jmin + 1) break;
  }
  return jmax - 1;
}
constexpr int TAG_DUST_SHIFT = 30;  // Photon::tag bits 30-31: dust type of the last interaction, 0-based (p%dust_id - 1)

// emit (src/sources/source.f90:100-179): returns false on a fatal model error
// dust_sample_emit_probability (dust_type_4elem.f90:356-377) at the frequency of the monochromatic run: the
// probabilities of the two bracketing emissivity states (tabulated by the host as log10, -inf for zero)
// interpolated in the log; zero if either is zero
__device__ __forceinline__ double mono_emit_probability(const ModelDev &M, int id, int jid, double frac) {
  const double l1 = M.mono_logp[id][jid], l2 = M.mono_logp[id][jid + 1];
  if (isinf(l1) || isinf(l2)) return 0.0;
  return pow(10.0, l1 + frac * (l2 - l1));
}

template <int ND>
__device__ bool place_emitted(const ModelDev &M, Photon<ND> &p);

template <int ND>
__device__ bool emit_photon(const ModelDev &M, Photon<ND> &p, Rng &rng, double &energy_emitted, const int reemit_src = -1,
                            const double reemit_energy = 1.0) {
  int is = 0;
  const int ns = M.n_sources;
  if (reemit_src >= 0) {
    // emit(p, reemit=.true., reemit_id, reemit_energy) (source.f90:128-140): the re-absorbing source emits
    // again; the reference still draws the source-selection number first
    if (ns > 1) rng.next();
    is = reemit_src;
  } else if (ns > 1) {
    double xi = rng.next();
    if (M.sample_evenly) {
      is = min((int)(xi * ns), ns - 1);
    } else {
      // sample_pdf_discrete_dp (type_pdf.f90:313-337): first source with cdf >= xi
      is = pick_source(M, xi);
    }
  }
  const SourceDev &S = M.sources[is];
  p.tag = (uint32_t)(is + 1);
  p.nx = p.ny = p.nz = 0.0;
  int64_t map_ic = 0;
  int ispot = -1;
  if (S.type == HYP_SOURCE_SPHERE) {
    // emit_from_sphere (source_type.f90:604-690): random point of the surface, direction from the
    // cosine law (or the limb-darkened law) about the local normal
    Angle a_coord;
    if (S.n_spots > 0) {
      // source type 3 (source_type.f90:421-427): a spot or the star by luminosity; inside a spot the
      // point is rejection-sampled (:632-637)
      const SpotDev *sp = M.spots + S.spot_off;
      const double xi = rng.next();
      int k = S.n_spots;
      if (xi <= sp[0].cdf) k = 0;
      else if (xi < sp[S.n_spots].cdf)
        for (k = 1; k < S.n_spots && xi > sp[k].cdf; ++k) {}
      ispot = k < S.n_spots ? k : -1;
      a_coord = random_sphere_angle(rng);
      if (ispot >= 0) {
        const SpotDev &q = sp[ispot];
        while (!(a_coord.sint * a_coord.cosp * q.a_sint * q.a_cosp + a_coord.sint * a_coord.sinp * q.a_sint * q.a_sinp +
                 a_coord.cost * q.a_cost > q.cost))
          a_coord = random_sphere_angle(rng);
      }
    } else {
      a_coord = random_sphere_angle(rng);
    }
    const double phi_local = 6.283185307179586476925286766559 * rng.next();
    Angle a_local;
    sincos(phi_local, &a_local.sinp, &a_local.cosp);
    a_local.cost = S.limb ? ran_mu_limb(rng) : sqrt(rng.next());
    a_local.sint = sqrt(1.0 - a_local.cost * a_local.cost);
    const Angle a = rotate_angle(a_local, a_coord);
    set_dir(p, a);
    p.nx = a_coord.sint * a_coord.cosp;
    p.ny = a_coord.sint * a_coord.sinp;
    p.nz = a_coord.cost;
    p.r0x = p.nx * S.radius + S.x;
    p.r0y = p.ny * S.radius + S.y;
    p.r0z = p.nz * S.radius + S.z;
  } else if (S.type == HYP_SOURCE_EXTERN_SPH) {
    // emit_from_extern_sph (source_type.f90:748-802): from the sphere inwards; (nx, ny, nz) is the inward
    // normal the peel-off weight refers to (source_a = -a_coord, emit_from_extern_sph_peeloff :804-820)
    const Angle a_coord = random_sphere_angle(rng);
    const double phi_local = 6.283185307179586476925286766559 * rng.next();
    Angle a_local;
    sincos(phi_local, &a_local.sinp, &a_local.cosp);
    a_local.cost = sqrt(rng.next());
    a_local.sint = sqrt(1.0 - a_local.cost * a_local.cost);
    const Angle a = rotate_angle(a_local, a_coord);
    set_dir(p, Angle{-a.cost, a.sint, -a.cosp, -a.sinp});  // minus_angle (type_angle3d.f90:443-447)
    const double ux = a_coord.sint * a_coord.cosp, uy = a_coord.sint * a_coord.sinp, uz = a_coord.cost;
    p.r0x = ux * S.radius + S.x;
    p.r0y = uy * S.radius + S.y;
    p.r0z = uz * S.radius + S.z;
    // -a_coord as a vector: sint * (-cosp), sint * (-sinp), -cost
    p.nx = a_coord.sint * -a_coord.cosp;
    p.ny = a_coord.sint * -a_coord.sinp;
    p.nz = -a_coord.cost;
  } else if (S.type == HYP_SOURCE_EXTERN_BOX) {
    // emit_from_extern_box (source_type.f90:822-904): face by area, cosine law about the inward normal
    const int face = (int)sample_discrete(S.face_cdf, 6, rng.next());
    const double phi_local = 6.283185307179586476925286766559 * rng.next();
    Angle a_local;
    sincos(phi_local, &a_local.sinp, &a_local.cosp);
    a_local.cost = sqrt(rng.next());
    a_local.sint = sqrt(1.0 - a_local.cost * a_local.cost);
    const int ax = face >> 1;  // the face is perpendicular to this axis; the other two coordinates are uniform
    double r[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (k == ax) r[k] = S.box[2 * k + (face & 1)];
      else r[k] = S.box[2 * k] + (S.box[2 * k + 1] - S.box[2 * k]) * rng.next();
    }
    p.r0x = r[0]; p.r0y = r[1]; p.r0z = r[2];
    const double sgn = (face & 1) ? -1.0 : 1.0;
    const Angle a_coord = ax == 0 ? Angle{0.0, sgn, 1.0, 0.0} : (ax == 1 ? Angle{0.0, sgn, 0.0, 1.0} : Angle{sgn, 0.0, 1.0, 0.0});
    set_dir(p, rotate_angle(a_local, a_coord));
    p.nx = a_coord.sint * a_coord.cosp;
    p.ny = a_coord.sint * a_coord.sinp;
    p.nz = a_coord.cost;
  } else if (S.type == HYP_SOURCE_PLANE_PARALLEL) {
    // emit_from_plane_parallel (source_type.f90:935-980).  Such a source is never peeled off
    // (hyperion/sources/source.py:926-929); the normal only marks the emission as non-isotropic.
    const double r = pow(rng.next(), 0.5) * S.radius;
    const double phi = 360.0 * rng.next();
    const double deg2rad = 3.14159265358979323846 / 180.0;
    const Angle a_local{cos(90.0 * deg2rad), sin(90.0 * deg2rad), cos(phi * deg2rad), sin(phi * deg2rad)};
    const Angle dir{S.dir_cost, S.dir_sint, S.dir_cosp, S.dir_sinp};
    const Angle a_final = rotate_angle(a_local, dir);
    p.r0x = a_final.sint * a_final.cosp * r + S.x;
    p.r0y = a_final.sint * a_final.sinp * r + S.y;
    p.r0z = a_final.cost * r + S.z;
    set_dir(p, dir);
    p.nx = p.vx; p.ny = p.vy; p.nz = p.vz;
  } else if (S.type == HYP_SOURCE_MAP) {
    // emit_from_map (source_type.f90:713-746): cell by luminosity, uniform position in it, isotropic
    map_ic = sample_discrete(M.map_cdf + S.map_off, M.n_cells, rng.next());
    random_position_cell(M, map_ic, rng, p.r0x, p.r0y, p.r0z);
    Angle a = random_sphere_angle(rng);
    set_dir(p, a);
  } else if (S.type == HYP_SOURCE_POINT_COLLECTION) {
    // emit_from_point_collection (source_type.f90:570-598)
    const int64_t k = S.coll_off + sample_discrete(M.coll_cdf + S.coll_off, S.coll_n, rng.next());
    p.r0x = M.coll_xyz[3 * k];
    p.r0y = M.coll_xyz[3 * k + 1];
    p.r0z = M.coll_xyz[3 * k + 2];
    Angle a = random_sphere_angle(rng);
    set_dir(p, a);
  } else {
    // emit_from_point (source_type.f90:539-564)
    p.r0x = S.x;
    p.r0y = S.y;
    p.r0z = S.z;
    Angle a = random_sphere_angle(rng);
    set_dir(p, a);
  }
  p.sQ = p.sU = p.sV = 0.0;
  p.energy = reemit_src >= 0 ? reemit_energy : 1.0;
  if (M.mono_nu > 0.0) {
    // monochromatic mode (source_emit with nu given, source_type.f90:436-474; emit, source.f90:161): the
    // frequency is fixed and the energy carries the spectrum's probability there (tabulated by the host per
    // source, times energy_total / n_photons)
    p.nu = M.mono_nu;
    if (reemit_src < 0) {
      if (ispot >= 0) {
        p.energy = M.mono_spot_w[S.spot_off + ispot];
      } else if (S.freq_type == HYP_SPECTRUM_LTE) {
        // select_dust_specific_energy_rho, then dust_sample_emit_probability at the cell's emissivity state
        const size_t base = (size_t)map_ic * ND;
        double w[ND], tot = 0.0;
#pragma unroll
        for (int k = 0; k < ND; ++k) {
          tot += M.specific_energy[base + k] * M.cells[base + k].rho;
          w[k] = tot;
        }
        const double xi = rng.next();
        int id = ND - 1;
#pragma unroll
        for (int k = ND - 2; k >= 0; --k)
          if (xi <= w[k] / tot) id = k;
        if (xi >= 1.0) id = ND - 1;
        p.energy = mono_emit_probability(M, id, M.jnu_id[base + id], M.jnu_frac[base + id]) * M.mono_lte_w;
        p.tag |= (uint32_t)id << TAG_DUST_SHIFT;
      } else {
        p.energy = M.mono_src_w[is];
      }
    }
  } else
  if (ispot >= 0) {
    // the spot's own spectrum; tables are sampled with sample_pdf_log (source_type.f90:480-486,
    // type_pdf.f90:383-400): x interpolated in the log against the cdf
    const SpotDev &q = M.spots[S.spot_off + ispot];
    if (q.freq_type == HYP_SPECTRUM_BLACKBODY) {
      p.nu = sample_planck(rng, q.temperature);
    } else {
      const SpectrumDev &sp = M.spectra[q.spectrum];
      const double *x = sp.B + sp.L.o_x, *cdf = sp.B + sp.L.o_cdf;
      const int n = sp.L.n;
      const double xi = rng.next();
      if (xi <= cdf[0]) {
        p.nu = x[0];
      } else if (xi >= cdf[n - 1]) {
        p.nu = x[n - 1];
      } else {
        const int i = lower_interval(cdf, n, xi);
        const double frac = (xi - cdf[i]) / (cdf[i + 1] - cdf[i]);
        p.nu = (x[i] == 0.0 || x[i + 1] == 0.0) ? 0.0 : pow(10.0, log10(x[i]) + frac * (log10(x[i + 1]) - log10(x[i])));
      }
    }
  } else if (S.freq_type == HYP_SPECTRUM_BLACKBODY) {
    p.nu = sample_planck(rng, S.temperature);
  } else if (S.freq_type == HYP_SPECTRUM_LTE) {
    // select_dust_specific_energy_rho (grid_physics_3d.f90:101-109; draws even for one dust type), then
    // dust_sample_j_nu at the cell's emissivity state (source_type.f90:500-505)
    const size_t base = (size_t)map_ic * ND;
    double w[ND], tot = 0.0;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      tot += M.specific_energy[base + k] * M.cells[base + k].rho;
      w[k] = tot;
    }
    const double xi = rng.next();
    int id = ND - 1;
#pragma unroll
    for (int k = ND - 2; k >= 0; --k)
      if (xi <= w[k] / tot) id = k;
    if (xi >= 1.0) id = ND - 1;
    const DustDev &d = M.dust[id];
    const int jid = M.jnu_id[base + id];
    const double frac = M.jnu_frac[base + id];
    const double x2 = rng.next();
    const int ne = d.L.n_enu;
    const double *enu = d.B + d.L.o_enu;
    double nu1, nu2;
    sample_powerlaw_pair(enu, d.B + d.L.o_ecdf + (size_t)jid * ne, d.B + d.L.o_einvb + (size_t)jid * (ne - 1),
                         d.B + d.L.o_erm1 + (size_t)jid * (ne - 1), d.B + d.L.o_ecdf + (size_t)(jid + 1) * ne,
                         d.B + d.L.o_einvb + (size_t)(jid + 1) * (ne - 1), d.B + d.L.o_erm1 + (size_t)(jid + 1) * (ne - 1),
                         ne, x2, nu1, nu2);
    const double l1 = log10(nu1);
    p.nu = pow(10.0, l1 + frac * (log10(nu2) - l1));
    p.tag |= (uint32_t)id << TAG_DUST_SHIFT;
  } else {
    const SpectrumDev &sp = M.spectra[S.spectrum];
    p.nu = sample_powerlaw(sp.B + sp.L.o_x, sp.B + sp.L.o_cdf, sp.B + sp.L.o_invb, sp.B + sp.L.o_rm1, sp.L.n,
                           rng.next());
  }
  if (reemit_src < 0) {
    if (M.sample_evenly) p.energy = p.energy * S.pdf * ns;
    energy_emitted += p.energy;
    p.energy0 = p.energy;
  }
  return place_emitted<ND>(M, p);
}

// update_optconsts + find_cell for a packet whose position, direction and frequency are set (emit, source.f90:165-171)
template <int ND>
__device__ bool place_emitted(const ModelDev &M, Photon<ND> &p) {
  if (!update_optconsts<ND>(M, p)) {
    atomicMax(M.error_flag, ERR_NU_RANGE);
    return false;
  }
  int fx, fy, fz;
  bool ok;
  if (M.grid_type == GEO_AMR) {
    int g;
    if (!amr_find_cell(M.amr, p.r0x, p.r0y, p.r0z, g, p.ix, p.iy, p.iz)) {
      atomicMax(M.error_flag, ERR_NOT_IN_CELL);
      return false;
    }
    p.ic = amr_cell_id(M.amr.grids[g], p.ix, p.iy, p.iz);
    p.n_inter = 0;
    p.n_reabs = 0;
    p.t = 0.0;
    return true;
  }
  if (M.grid_type == GEO_VOR) {
    const int cell = vor_find_cell(M.vor, p.r0x, p.r0y, p.r0z);
    if (cell < 0) {
      atomicMax(M.error_flag, ERR_NOT_IN_CELL);
      return false;
    }
    p.ix = p.iy = p.iz = 0;
    p.ic = cell;
    p.n_inter = 0;
    p.n_reabs = 0;
    p.t = 0.0;
    return true;
  }
  if (M.grid_type == GEO_OCT) {
    const int node = oct_find_cell(M.oct, p.r0x, p.r0y, p.r0z);
    if (node < 0) {
      atomicMax(M.error_flag, ERR_NOT_IN_CELL);
      return false;
    }
    p.ix = p.iy = p.iz = 0;
    p.ic = node;
    p.n_inter = 0;
    p.n_reabs = 0;
    p.t = 0.0;
    return true;
  }
  if (M.grid_type == GEO_SPH) {
    // the flight kernel applies adjust_wall when it starts the ray; the slot keeps find_cell's cell
    ok = sph_find_cell(M.sph, p.r0x, p.r0y, p.r0z, p.vx, p.vy, p.vz, fx, fy, fz);
    p.ix = fx; p.iy = fy; p.iz = fz;
  } else {
    ok = place_axis(M.w1, M.n1, p.r0x, p.vx, p.ix, fx);
    ok = place_axis(M.w2, M.n2, p.r0y, p.vy, p.iy, fy) && ok;
    ok = place_axis(M.w3, M.n3, p.r0z, p.vz, p.iz, fz) && ok;
  }
  if (!ok) {
    atomicMax(M.error_flag, ERR_NOT_IN_CELL);
    return false;
  }
  p.ic = (fz * M.n2 + fy) * M.n1 + fx;
  p.n_inter = 0;
  p.n_reabs = 0;
  p.t = 0.0;
  return true;
}

// dust_scatter (src/dust/dust_type_4elem.f90:446-566)
template <int ND>
__device__ bool scatter_photon(const ModelDev &M, const DustDev &d, Photon<ND> &p, Rng &rng) {
  Angle as = random_sphere_angle(rng);
  const double sin_2_i1 = 2.0 * as.sinp * as.cosp;
  const double cos_2_i1 = 1.0 - 2.0 * as.sinp * as.sinp;
  double c1 = 1.0;
  double c2 = cos_2_i1 * p.sQ - sin_2_i1 * p.sU;
  const double ctot = c1 + c2;
  c1 = c1 / ctot;
  c2 = c2 / ctot;
  double P1 = 1.0, P2 = 0.0, P3 = 1.0, P4 = 0.0;
  const int n_mu = d.L.n_mu;
  if (p.nu >= d.L.nu_min && p.nu <= d.L.nu_max) {
    const double *nu = d.B + d.L.o_nu;
    const double *mu = d.B + d.L.o_mu;
    const int j = lower_interval(nu, d.L.n_nu, p.nu);
    const double xi = rng.next();
    const double *C1 = d.B + d.L.o_C1 + (size_t)j * n_mu;
    const double *C2 = d.B + d.L.o_C2 + (size_t)j * n_mu;
    const bool zp2 = d.L.zero_p2 != 0;
    // bisection of dust_scatter, kept in the reference's 1-based indices (imu in [1, n_mu-1])
    int imin = 1, imax = n_mu, imu = 1;
    double cdf1 = 0.0, cdf2 = 1.0;
    bool found = false;
    for (int it = 0; it < 64; ++it) {
      imu = (imax + imin) / 2;
      if (zp2) {
        cdf1 = __ldg(C1 + imu - 1);
        cdf2 = __ldg(C1 + imu);
      } else {
        cdf1 = c1 * __ldg(C1 + imu - 1) + c2 * __ldg(C2 + imu - 1);
        cdf2 = c1 * __ldg(C1 + imu) + c2 * __ldg(C2 + imu);
      }
      if (xi > cdf2)
        imin = imu;
      else if (xi < cdf1)
        imax = imu;
      else {
        found = true;
        break;
      }
      if (imin == imax) break;
    }
    if (!found) {
      atomicMax(M.error_flag, ERR_SCATTER);
      return false;
    }
    imu -= 1;  // 0-based interval [imu, imu+1]
    const double m0 = __ldg(mu + imu), m1 = __ldg(mu + imu + 1);
    as.cost = (xi - cdf1) / (cdf2 - cdf1) * (m1 - m0) + m0;
    as.sint = sqrt(1.0 - as.cost * as.cost);
    // bilinear weights in (mu, nu)
    const int i = lower_interval(mu, n_mu, as.cost);
    const double x0 = __ldg(mu + i), x1 = __ldg(mu + i + 1);
    const double y0 = __ldg(nu + j), y1 = __ldg(nu + j + 1);
    const double norm = 1.0 / (x1 - x0) / (y1 - y0);
    const double wx0 = as.cost - x0, wx1 = x1 - as.cost, wy0 = p.nu - y0, wy1 = y1 - p.nu;
    P1 = interp_phase(d.B + d.L.o_P1, n_mu, i, j, wx0, wx1, wy0, wy1, norm);
    P2 = interp_phase(d.B + d.L.o_P2, n_mu, i, j, wx0, wx1, wy0, wy1, norm);
    P3 = interp_phase(d.B + d.L.o_P3, n_mu, i, j, wx0, wx1, wy0, wy1, norm);
    P4 = interp_phase(d.B + d.L.o_P4, n_mu, i, j, wx0, wx1, wy0, wy1, norm);
  }
  Angle ac;
  angle_from_dir(p, ac);
  Angle af = rotate_angle(as, ac);
  Stokes s{1.0, p.sQ, p.sU, p.sV};
  scatter_stokes(s, ac, as, af, P1, P2, P3, P4);
  const double norm = 1.0 / s.I;
  p.sQ = s.Q * norm;
  p.sU = s.U * norm;
  p.sV = s.V * norm;
  set_dir(p, af);
  return true;
}

// A flight that ended on a stellar surface (Slot::t < 0 carries the source): emit(p, reemit=.true., ...) from
// that source with the packet's energy (iter_lucy.f90:158-185, iter_final.f90:212-242).  After n_reabs_max
// successive re-emissions that were all re-absorbed the packet is killed, as in the reference.
template <int ND>
__device__ bool reemit_photon(const ModelDev &M, Photon<ND> &p, Rng &rng, uint32_t &n_killed_int) {
  const int src = (int)(-p.t) - 1;
  const uint32_t nre = p.n_reabs;
  if ((int64_t)nre >= M.n_reabs_max) {
    ++n_killed_int;
    return false;
  }
  const uint32_t n_inter = p.n_inter;
  double dummy = 0.0;
  if (!emit_photon<ND>(M, p, rng, dummy, src, p.energy)) return false;
  p.n_inter = n_inter;
  p.n_reabs = min(nre + 1u, 0x7ffffffeu);
  p.tag = (uint32_t)(src + 1);
  return true;
}

// interact (src/dust/dust_interact.f90:22-79).  Returns: 0 continue, 1 packet finished (killed).
// The last, partial cell of a flight that ended on the wave engine (flight_wave.cuh): the tile kernel only notices
// that the optical depth runs out inside the cell and hands the packet over with tau_left = -(optical depth left at
// the cell's entry wall); the path length to the interaction point and its deposit (grid_propagate_3d.f90:186-228)
// are made here, once per interaction, instead of in a divergent branch of the crossing loop.  The density is rounded
// to fp32 as the tile kernel stages it, so that both see the same optical depth of the cell.
template <int ND>
__device__ __forceinline__ void wave_partial_step(const ModelDev &M, Photon<ND> &p) {
  double rho[ND], chi_rho = 0.0;
#pragma unroll
  for (int k = 0; k < ND; ++k) {
    rho[k] = (double)fmaxf((float)__ldcg(&M.cells[(size_t)p.ic * ND + k].rho), 0.f);
    chi_rho = fma(p.chi[k], rho[k], chi_rho);
  }
  const double len = chi_rho > 0.0 ? -p.tau_left / chi_rho : 0.0;
#pragma unroll
  for (int k = 0; k < ND; ++k)
    if (rho[k] > 0.0) atomicAdd(&M.cells[(size_t)p.ic * ND + k].esum, len * p.kE[k]);
  p.t += len;
  p.tau_left = 0.0;
}

template <int ND>
__device__ int interact_photon(const ModelDev &M, Photon<ND> &p, Rng &rng, uint32_t &n_abs, uint32_t &n_scat,
                               uint32_t &n_killed_int, int &dust_id, bool &was_scattered, const bool force_scatter = false) {
  p.n_reabs = 0;
  // the loop guard of do_lucy (iter_lucy.f90:193-198)
  p.n_inter += 1;
  if ((int64_t)p.n_inter > M.n_inter_max) {
    ++n_killed_int;
    return 1;
  }
  // move to the interaction point
  p.r0x = p.r0x + p.t * p.vx;
  p.r0y = p.r0y + p.t * p.vy;
  p.r0z = p.r0z + p.t * p.vz;
  p.t = 0.0;
  // select_dust_chi_rho (src/grid/grid_physics_3d.f90:87-99)
  int id = 0;
  if (ND > 1) {
    double w[ND], tot = 0.0;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      tot += p.chi[k] * M.cells[(size_t)p.ic * ND + k].rho;
      w[k] = tot;
    }
    double xi = rng.next();
    id = ND - 1;
#pragma unroll
    for (int k = ND - 2; k >= 0; --k)
      if (xi <= w[k] / tot) id = k;
    if (xi >= 1.0) id = ND - 1;
  }
  const DustDev &d = M.dust[id];
  dust_id = id;
  double albedo = p.albedo[0];
#pragma unroll
  for (int k = 1; k < ND; ++k)
    if (k == id) albedo = p.albedo[k];
  // a forced scattering draws no number and weighs the packet with the albedo (dust_interact.f90:47-52,75-77)
  const double xi = force_scatter ? 0.0 : rng.next();
  bool scattered;
  if (xi > albedo) {
    // dust_emit (dust_type_4elem.f90:334-354) + dust_sample_j_nu (:379-398)
    const size_t k = (size_t)p.ic * ND + id;
    const int jid = M.jnu_id[k];
    const double frac = M.jnu_frac[k];
    const double x2 = rng.next();
    const int ne = d.L.n_enu;
    const double *enu = d.B + d.L.o_enu;
    double nu1, nu2;
    sample_powerlaw_pair(enu, d.B + d.L.o_ecdf + (size_t)jid * ne, d.B + d.L.o_einvb + (size_t)jid * (ne - 1),
                         d.B + d.L.o_erm1 + (size_t)jid * (ne - 1), d.B + d.L.o_ecdf + (size_t)(jid + 1) * ne,
                         d.B + d.L.o_einvb + (size_t)(jid + 1) * (ne - 1), d.B + d.L.o_erm1 + (size_t)(jid + 1) * (ne - 1),
                         ne, x2, nu1, nu2);
    const double l1 = log10(nu1);
    p.nu = pow(10.0, l1 + frac * (log10(nu2) - l1));
    p.sQ = p.sU = p.sV = 0.0;
    Angle a = random_sphere_angle(rng);
    set_dir(p, a);
    if (!update_optconsts<ND>(M, p)) {
      atomicMax(M.error_flag, ERR_NU_RANGE);
      return 1;
    }
    scattered = false;
    ++n_abs;
  } else {
    if (!scatter_photon<ND>(M, d, p, rng)) return 1;
    scattered = true;
    ++n_scat;
  }
  was_scattered = scattered;
  if (force_scatter) {
    p.energy = p.energy * albedo;
    // iter_final_mono.f90:334: killed without being counted
    if ((M.kill_on_scatter && scattered) || p.energy < p.energy0 * M.mono_threshold) return 1;
    return 0;
  }
  if ((M.kill_on_scatter && scattered) || (M.kill_on_absorb && !scattered)) return 1;
  return 0;
}

// distance_to_closest_wall of each geometry module (e.g. grid_geometry_cartesian_3d.f90:396-422)
template <int ND>
__device__ inline double distance_to_closest_wall(const ModelDev &M, const Photon<ND> &p) {
  const double x = p.r0x, y = p.r0y, z = p.r0z;
  double d;
  if (M.grid_type == GEO_CAR) {
    d = fmin(fmin(fmin(x - M.w1[p.ix], M.w1[p.ix + 1] - x), fmin(y - M.w2[p.iy], M.w2[p.iy + 1] - y)),
             fmin(z - M.w3[p.iz], M.w3[p.iz + 1] - z));
  } else if (M.grid_type == GEO_SPH) {
    const SphGrid &G = M.sph;
    const double *w1 = G.T + G.o_w1, *w2 = G.T + G.o_w2;
    const double rcyl = sqrt(x * x + y * y);
    double d1, d2, d3, d4;
    if (G.kind == POLAR_CYL) {
      d1 = rcyl - w1[p.ix];
      d2 = w1[p.ix + 1] - rcyl;
      d3 = z - w2[p.iy];
      d4 = w2[p.iy + 1] - z;
    } else {
      const double r = sqrt(x * x + y * y + z * z);
      d1 = r - w1[p.ix];
      d2 = w1[p.ix + 1] - r;
      if (fabs(d1) < G.T[G.o_ew1 + p.ix]) d1 = 0.0;
      if (fabs(d2) < G.T[G.o_ew1 + p.ix + 1]) d2 = 0.0;
      const double ta = G.T[G.o_wtant + p.iy], tb = G.T[G.o_wtant + p.iy + 1];
      d3 = fabs(-rcyl + ta * z) / sqrt(1.0 + ta * ta);
      d4 = fabs(-rcyl + tb * z) / sqrt(1.0 + tb * tb);
    }
    double d5 = 1.7976931348623157e308, d6 = d5;
    if (G.n3 > 1) {
      const double pa = G.T[G.o_wtanp + p.iz], pb = G.T[G.o_wtanp + p.iz + 1];
      d5 = fabs(pa * x - y) / sqrt(pa * pa + 1.0);
      d6 = fabs(pb * x - y) / sqrt(pb * pb + 1.0);
    }
    d = fmin(fmin(fmin(d1, d2), fmin(d3, d4)), fmin(d5, d6));
  } else if (M.grid_type == GEO_OCT) {
    const OctNode &N = M.oct.nodes[p.ic];
    d = fmin(fmin(fmin(x - N.x + N.dx, N.x + N.dx - x), fmin(y - N.y + N.dy, N.y + N.dy - y)),
             fmin(z - N.z + N.dz, N.z + N.dz - z));
  } else {
    const AmrGridDev &G = M.amr.grids[M.amr.cell_grid[p.ic]];
    d = fmin(fmin(fmin(x - amr_wall(G.xmin, G.xmax, p.ix, G.n1), amr_wall(G.xmin, G.xmax, p.ix + 1, G.n1) - x),
                  fmin(y - amr_wall(G.ymin, G.ymax, p.iy, G.n2), amr_wall(G.ymin, G.ymax, p.iy + 1, G.n2) - y)),
             fmin(z - amr_wall(G.zmin, G.zmax, p.iz, G.n3), amr_wall(G.zmin, G.zmax, p.iz + 1, G.n3) - z));
  }
  return d < 0.0 ? 0.0 : d;
}

__device__ __forceinline__ double mean_opacity_loglog(const DustDev &d, int64_t o_logy, double e) {
  // interp1d_loglog on the mean-opacity table (src/dust/dust.f90:81-121)
  const double *loge = d.B + d.L.o_loge;
  const double le = log10(e);
  int lo = 0, hi = d.L.n_e - 1;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(loge + mid) <= le) lo = mid; else hi = mid;
  }
  const double a = __ldg(loge + lo), b = __ldg(loge + lo + 1);
  return loglog_at(d.B + o_logy, lo, (le - a) / (b - a));
}


#include "pda.cuh"

// One step of the modified random walk in the packet's cell (grid_do_mrw / grid_do_mrw_noenergy,
// grid_mrw_3d.f90:56-149): jump to the surface of the largest sphere that fits in the cell, deposit the
// energy absorbed along the diffusive path (Lucy iteration), pick a new direction and a frequency from
// the local b_nu.  As in the reference, the opacities are NOT refreshed for the new frequency here.
template <int ND>
__device__ inline int mrw_step(const ModelDev &M, Photon<ND> &p, Rng &rng, const bool deposit, const double R0) {
  const size_t base = (size_t)p.ic * ND;
  if (deposit) {
    // sample_cumulative (:198-203): linear interpolation of x(y)
    const double *xc = M.mrw_cdf, *yc = M.mrw_cdf + 100;
    const double xi = rng.next();
    const int j = lower_interval(yc, 100, xi);
    const double y0 = __ldg(yc + j), y1 = __ldg(yc + j + 1);
    const double yv = (xi - y0) / (y1 - y0) * (__ldg(xc + j + 1) - __ldg(xc + j)) + __ldg(xc + j);
    const double q = R0 / 3.14159265358979323846;
    const double ct = -log(yv) / M.diff_coeff[p.ic] * (q * q);
#pragma unroll
    for (int id = 0; id < ND; ++id) {
      if (M.cells[base + id].rho > 0.0) {
        const double e = p.energy * ct * mean_opacity_loglog(M.dust[id], M.dust[id].L.o_logkap_planck, M.specific_energy[base + id]);
        atomicAdd(&M.cells[base + id].esum, e);
        if (M.spec_sums) {
          // deposit_specific_energy_spectrum (grid_physics_3d.f90:367-395): spread like the local emissivity
          const int iv = M.jnu_id[base + id];
          const double fr = M.jnu_frac[base + id];
          const double *f1 = M.spec_jfrac + ((size_t)id * M.spec_jmax + iv) * M.n_spec_bins, *f2 = f1 + M.n_spec_bins;
          const size_t plane = (size_t)M.n_cells * ND;
          for (int b = 0; b < M.n_spec_bins; ++b) {
            const double w = (1.0 - fr) * __ldg(f1 + b) + fr * __ldg(f2 + b);
            if (w != 0.0) atomicAdd(M.spec_sums + (size_t)b * plane + base + id, e * w);
          }
        }
      }
    }
  }
  // random_sphere_vector3d (type_vector3d.f90:323-333)
  const double mu = -1.0 + 2.0 * rng.next();
  const double phi = 6.283185307179586476925286766559 * rng.next();
  const double cut = sqrt(1.0 - mu * mu);
  double sp, cp;
  sincos(phi, &sp, &cp);
  p.r0x += cut * cp * R0;
  p.r0y += cut * sp * R0;
  p.r0z += mu * R0;
  const Angle a = random_sphere_angle(rng);
  set_dir(p, a);
  // select_dust_chi_rho with the (stale) opacities of the packet
  int id = 0;
  if (ND > 1) {
    double w[ND], tot = 0.0;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      tot += p.chi[k] * M.cells[base + k].rho;
      w[k] = tot;
    }
    const double xi = rng.next();
    id = ND - 1;
#pragma unroll
    for (int k = ND - 2; k >= 0; --k)
      if (xi <= w[k] / tot) id = k;
    if (xi >= 1.0) id = ND - 1;
  }
  // dust_sample_b_nu (dust_type_4elem.f90:400-419)
  const DustDev &d = M.dust[id];
  const int jid = M.jnu_id[base + id];
  const double frac = M.jnu_frac[base + id];
  const double x2 = rng.next();
  const int ne = d.L.n_enu;
  const double *enu = d.B + d.L.o_enu;
  const double nu1 = sample_powerlaw(enu, d.Bm + d.Lm.o_bcdf + (size_t)jid * ne, d.Bm + d.Lm.o_binvb + (size_t)jid * (ne - 1),
                                     d.Bm + d.Lm.o_brm1 + (size_t)jid * (ne - 1), ne, x2);
  const double nu2 = sample_powerlaw(enu, d.Bm + d.Lm.o_bcdf + (size_t)(jid + 1) * ne,
                                     d.Bm + d.Lm.o_binvb + (size_t)(jid + 1) * (ne - 1),
                                     d.Bm + d.Lm.o_brm1 + (size_t)(jid + 1) * (ne - 1), ne, x2);
  const double l1 = log10(nu1);
  p.nu = pow(10.0, l1 + frac * (log10(nu2) - l1));
  p.sQ = p.sU = p.sV = 0.0;
  return id;
}

// The MRW block at the top of the interaction loop (iter_lucy.f90:134-148): while the cell is optically
// thick out to its closest wall, random-walk.  Returns false if n_mrw_max steps did not get the packet out
// (the reference kills it).
template <int ND>
__device__ inline bool mrw_loop(const ModelDev &M, Photon<ND> &p, Rng &rng, uint32_t &n_killed_int) {
  for (int64_t step = 0; step < M.n_mrw_max; ++step) {
    const double R0 = distance_to_closest_wall<ND>(M, p);
    if (!(M.alpha_inv_planck[p.ic] * R0 > M.mrw_gamma)) return true;
    mrw_step<ND>(M, p, rng, true, R0);
  }
  ++n_killed_int;
  return false;
}

// cell volume (setup_grid_geometry of each geometry module)
__device__ __forceinline__ double cell_volume(const ModelDev &M, int64_t ic) {
  if (M.grid_type == GEO_SPH) return sph_volume(M.sph, ic);
  if (M.grid_type == GEO_OCT) {
    const OctNode &N = M.oct.nodes[ic];
    return N.dx * N.dy * N.dz * 8.0;  // grid_geometry_octree.f90:250-253
  }
  if (M.grid_type == GEO_AMR) return amr_volume(M.amr, ic);
  if (M.grid_type == GEO_VOR) return M.vor.volume[ic];
  const int i1 = (int)(ic % M.n1), i2 = (int)((ic / M.n1) % M.n2), i3 = (int)(ic / ((int64_t)M.n1 * M.n2));
  return ((M.w1[i1 + 1] - M.w1[i1]) * (M.w2[i2 + 1] - M.w2[i2])) * (M.w3[i3 + 1] - M.w3[i3]);
}

// random_position_cell (grid_geometry_cartesian_3d.f90:383-394, grid_geometry_spherical_3d.f90:645-677)
__device__ inline void random_position_cell(const ModelDev &M, int64_t ic, Rng &rng, double &x, double &y, double &z) {
  if (M.grid_type == GEO_AMR) {
    // grid_geometry_amr.f90:729-741
    const AmrGridDev &G = M.amr.grids[M.amr.cell_grid[ic]];
    const int k = (int)ic - G.start_id;
    const int i1 = k % G.n1, i2 = (k / G.n1) % G.n2, i3 = k / (G.n1 * G.n2);
    const double x0 = amr_wall(G.xmin, G.xmax, i1, G.n1), x1 = amr_wall(G.xmin, G.xmax, i1 + 1, G.n1);
    const double y0 = amr_wall(G.ymin, G.ymax, i2, G.n2), y1 = amr_wall(G.ymin, G.ymax, i2 + 1, G.n2);
    const double z0 = amr_wall(G.zmin, G.zmax, i3, G.n3), z1 = amr_wall(G.zmin, G.zmax, i3 + 1, G.n3);
    x = rng.next() * (x1 - x0) + x0;
    y = rng.next() * (y1 - y0) + y0;
    z = rng.next() * (z1 - z0) + z0;
    return;
  }
  if (M.grid_type == GEO_VOR) {
    // grid_geometry_voronoi.f90:285-312: rejection sampling in the cell's bounding box
    const double *bb = M.vor.bb + 6 * (size_t)ic;
    for (int i = 0; i < 1000000; ++i) {
      x = rng.next() * (bb[1] - bb[0]) + bb[0];
      y = rng.next() * (bb[3] - bb[2]) + bb[2];
      z = rng.next() * (bb[5] - bb[4]) + bb[4];
      int i0, i1;
      vor_nearest(M.vor, x, y, z, 1, i0, i1);
      if (i0 == (int)ic) return;
    }
    return;
  }
  if (M.grid_type == GEO_OCT) {
    // grid_geometry_octree.f90:396-408
    const OctNode &N = M.oct.nodes[ic];
    x = (2.0 * rng.next() - 1.0) * N.dx + N.x;
    y = (2.0 * rng.next() - 1.0) * N.dy + N.y;
    z = (2.0 * rng.next() - 1.0) * N.dz + N.z;
    return;
  }
  const int i1 = (int)(ic % M.n1), i2 = (int)((ic / M.n1) % M.n2), i3 = (int)(ic / ((int64_t)M.n1 * M.n2));
  if (M.grid_type == GEO_SPH && M.sph.kind == POLAR_CYL) {
    // grid_geometry_cylindrical_3d.f90:516-547
    const SphGrid &G = M.sph;
    const double *w1 = G.T + G.o_w1, *w2 = G.T + G.o_w2, *w3 = G.T + G.o_w3;
    double r = rng.next(), zz = rng.next(), ph = rng.next();
    const double a = w1[i1], b = w1[i1 + 1];
    r = sqrt(r * (b * b - a * a) + a * a);
    zz = zz * (w2[i2 + 1] - w2[i2]) + w2[i2];
    ph = ph * (w3[i3 + 1] - w3[i3]) + w3[i3];
    if (r <= a || r >= b) r = 0.5 * (a + b);
    if (zz <= w2[i2] || zz >= w2[i2 + 1]) zz = 0.5 * (w2[i2] + w2[i2 + 1]);
    if (ph <= w3[i3] || ph >= w3[i3 + 1]) ph = 0.5 * (w3[i3] + w3[i3 + 1]);
    x = r * cos(ph);
    y = r * sin(ph);
    z = zz;
    return;
  }
  if (M.grid_type == GEO_SPH) {
    const SphGrid &G = M.sph;
    const double *w1 = G.T + G.o_w1, *w2 = G.T + G.o_w2, *w3 = G.T + G.o_w3, *wc = G.T + G.o_wcost;
    double r = rng.next(), t = rng.next(), ph = rng.next();
    const double a = w1[i1], b = w1[i1 + 1];
    r = pow(r * (b * b * b - a * a * a) + a * a * a, 1.0 / 3.0);
    t = acos(t * (wc[i2 + 1] - wc[i2]) + wc[i2]);
    ph = ph * (w3[i3 + 1] - w3[i3]) + w3[i3];
    if (r <= a || r >= b) r = 0.5 * (a + b);
    if (t <= w2[i2] || t >= w2[i2 + 1]) t = 0.5 * (w2[i2] + w2[i2 + 1]);
    if (ph <= w3[i3] || ph >= w3[i3 + 1]) ph = 0.5 * (w3[i3] + w3[i3 + 1]);
    x = r * sin(t) * cos(ph);
    y = r * sin(t) * sin(ph);
    z = r * cos(t);
    return;
  }
  const double x0 = M.w1[i1], x1 = M.w1[i1 + 1], y0 = M.w2[i2], y1 = M.w2[i2 + 1], z0 = M.w3[i3], z1 = M.w3[i3 + 1];
  x = rng.next() * (x1 - x0) + x0;
  y = rng.next() * (y1 - y0) + y0;
  z = rng.next() * (z1 - z0) + z0;
}

// =============================================================================================
// kernels of one round
// =============================================================================================
constexpr int SERVICE_THREADS = 128;

__global__ void pool_init_kernel(Pool P, uint32_t n_slots) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += gridDim.x * blockDim.x) P.q_emit[i] = i;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int k = 0; k < C_COUNT; ++k) P.counts[k] = 0;
    P.counts[C_NE] = n_slots;
    *P.next_photon = 0ull;
  }
}

// Packets are emitted in the order of a sort key (emit_keys_kernel below).
__device__ __forceinline__ uint32_t interleave12(uint32_t a, uint32_t b) {
  // 2-D Morton code of two 12-bit integers
  auto part = [](uint32_t x) {
    x &= 0xfffu;
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
  };
  return part(a) | (part(b) << 1);
}

// Sort key of packet `id`.  mode 1: source | cube-map face | Morton code of the direction (24 bits).
// mode 2: source | face | coarse direction (2 x KEY_COARSE_BITS) | path length to the first interaction
// in units of the distance to the grid edge, 8 log bins per octave (6 bits) | finer direction bits.  The
// 32 packets a beam warp claims then share a direction bin AND die within ~10 % of each other, which
// keeps the lanes of the lockstep march busy; the price is a wider beam (more cells per warp step).
// The key replays the packet's own random numbers (emit_photon + the first tau), so it changes the
// order in which packets are marched and summed, never a packet's path.
#ifndef HYP_DEFAULT_SORT
#define HYP_DEFAULT_SORT 2
#endif
#ifndef KEY_COARSE_BITS
#define KEY_COARSE_BITS 5
#endif
template <int ND>
__global__ void emit_keys_kernel(const ModelDev M, const unsigned long long first_id, const uint32_t count,
                                 const uint32_t iteration, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals,
                                 const int mode) {
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
    Rng rng;
    rng.init(M.seed, first_id + k, iteration);
    uint32_t is = 0;
    const int ns = M.n_sources;
    if (ns > 1) {
      const double xi = rng.next();
      if (M.sample_evenly) {
        is = (uint32_t)min((int)(xi * ns), ns - 1);
      } else {
        is = (uint32_t)pick_source(M, xi);
      }
    }
    const Angle a = random_sphere_angle(rng);
    const double d[3] = {a.sint * a.cosp, a.sint * a.sinp, a.cost};
    const double ax = fabs(d[0]), ay = fabs(d[1]), az = fabs(d[2]);
    int f = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
    const double m = f == 0 ? d[0] : (f == 1 ? d[1] : d[2]);
    const double u = (f == 0 ? d[1] : d[0]) / fabs(m), v = (f == 2 ? d[1] : d[2]) / fabs(m);
    const uint32_t iu = (uint32_t)min(4095.0, (u + 1.0) * 2048.0), iv = (uint32_t)min(4095.0, (v + 1.0) * 2048.0);
    const uint32_t face = (uint32_t)(2 * f + (m < 0.0 ? 1 : 0));
    uint32_t key = ((is & 31u) << 27) | (face << 24) | interleave12(iu, iv);
    if (mode == 2 && M.grid_type == GEO_CAR && M.sources[is].type != HYP_SOURCE_SPHERE) {
      // the packet as emit_kernel will make it (same id, same stream), then its first optical depth
      Photon<ND> p;
      Rng r2;
      r2.init(M.seed, first_id + k, iteration);
      double dummy = 0.0;
      uint32_t sbin = 63;
      if (emit_photon<ND>(M, p, r2, dummy) && p.ix >= 0 && p.ix < M.n1 && p.iy >= 0 && p.iy < M.n2 && p.iz >= 0 && p.iz < M.n3) {
        const double tau = -log(1.0 - r2.next());
        double chi_rho = 0.0;
#pragma unroll
        for (int id = 0; id < ND; ++id) chi_rho += p.chi[id] * M.cells[(size_t)p.ic * ND + id].rho;
        // distance to the edge of the grid along the direction
        const double tx = p.vx > 0.0 ? (M.w1[M.n1] - p.r0x) / p.vx : (p.vx < 0.0 ? (M.w1[0] - p.r0x) / p.vx : 1e300);
        const double ty = p.vy > 0.0 ? (M.w2[M.n2] - p.r0y) / p.vy : (p.vy < 0.0 ? (M.w2[0] - p.r0y) / p.vy : 1e300);
        const double tz = p.vz > 0.0 ? (M.w3[M.n3] - p.r0z) / p.vz : (p.vz < 0.0 ? (M.w3[0] - p.r0z) / p.vz : 1e300);
        const double t_edge = fmin(tx, fmin(ty, tz));
        if (chi_rho > 0.0 && t_edge > 0.0) {
          const double frac = tau / (chi_rho * t_edge);
          sbin = (uint32_t)min(63.0, fmax(0.0, 8.0 * log2(frac) + 48.0));
        }
      }
      constexpr int CB = KEY_COARSE_BITS, FB = (18 - 2 * CB) / 2;   // coarse / fine direction bits per axis
      const uint32_t cu = iu >> (12 - CB), cv = iv >> (12 - CB);
      const uint32_t fu = (iu >> (12 - CB - FB)) & ((1u << FB) - 1u), fv = (iv >> (12 - CB - FB)) & ((1u << FB) - 1u);
      key = ((is & 31u) << 27) | (face << 24) | (interleave12(cu, cv) << (6 + 2 * FB)) | (sbin << (2 * FB)) | interleave12(fu, fv);
    }
    keys[k] = key;
    vals[k] = k;
  }
}

// Fill the free slots with new packets while ids remain.
#ifndef INTERACT_MIN_BLOCKS
#define INTERACT_MIN_BLOCKS 8   // resident 128-thread blocks per SM the register allocation aims at (64 registers)
#endif
template <int ND>
__global__ void __launch_bounds__(SERVICE_THREADS, INTERACT_MIN_BLOCKS)
emit_kernel(const ModelDev M, Pool P, const unsigned long long first_id, const unsigned long long n_photons, const uint32_t iteration) {
  const uint32_t n = P.counts[C_NE];
  const unsigned lane = threadIdx.x & 31;
  Slot<ND> *slots = (Slot<ND> *)P.slots;
  double energy_emitted = 0.0;
  uint32_t n_run = 0, n_esc = 0;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
    const uint32_t i = base + lane;
    const bool valid = i < n;
    // claim packet ids (warp-aggregated)
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    const int leader = __ffs(m) - 1;
    unsigned long long k0 = 0;
    if ((int)lane == leader) k0 = atomicAdd(P.next_photon, (unsigned long long)__popc(m));
    k0 = __shfl_sync(0xffffffffu, k0, leader);
    unsigned long long k = k0 + __popc(m & ((1u << lane) - 1u));  // position in emission order
    bool go = valid && k < n_photons;
    uint32_t slot = 0;
    if (go) {
      slot = P.q_emit[i];
      Photon<ND> p;
      Rng rng;
      unsigned long long id = 0;
      for (;;) {
        // k-th packet in emission order -> packet id (window base + sorted offset)
        id = first_id + (k & ~(unsigned long long)(P.window - 1)) + P.perm[k & (2ull * P.window - 1)];
        rng.init(M.seed, id, iteration);
        ++n_run;
        if (!emit_photon<ND>(M, p, rng, energy_emitted)) {
          go = false;  // fatal model error is flagged; the host reports it after the round
          break;
        }
        // a packet emitted on the outer wall moving outwards escapes immediately
        if (M.grid_type == GEO_CAR && (p.ix < 0 || p.ix >= M.n1 || p.iy < 0 || p.iy >= M.n2 || p.iz < 0 || p.iz >= M.n3)) {
          ++n_esc;
          k = atomicAdd(P.next_photon, 1ull);
          if (k >= n_photons) {
            go = false;
            break;
          }
          continue;
        }
        break;
      }
      if (go) {
        p.tau_left = -log(1.0 - rng.next());  // random_exp (lib_random.f90:227-236)
        store_photon<ND>(slots + slot, p, rng, id);
      }
    }
    queue_append(go, P.q_beam, P.counts + C_NB, slot);
  }
  warp_add_scalar(M.scalars + SC_ENERGY, energy_emitted);
  warp_add_scalar(M.scalars + SC_PHOTONS, (double)n_run);
  warp_add_scalar(M.scalars + SC_ESC, (double)n_esc);
}

// Interactions of every packet whose flight ended inside the grid.
template <int ND>
__global__ void __launch_bounds__(SERVICE_THREADS, INTERACT_MIN_BLOCKS)
interact_kernel(const ModelDev M, Pool P, uint32_t *__restrict__ q_flight_next, uint32_t *n_flight_next,
                const uint32_t iteration) {
  const uint32_t n = P.counts[C_NI];
  const unsigned lane = threadIdx.x & 31;
  Slot<ND> *slots = (Slot<ND> *)P.slots;
  uint32_t n_abs = 0, n_scat = 0, n_kill = 0;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
    const uint32_t i = base + lane;
    const bool valid = i < n;
    uint32_t slot = 0;
    bool alive = false;
    if (valid) {
      slot = P.q_interact[i];
      Photon<ND> p;
      Rng rng;
      load_photon<ND>(slots + slot, p, rng, M.seed, iteration);
      const uint64_t id = slots[slot].id;
      int dust_id = 0;
      bool scattered = false;
      const bool was_reabsorbed = p.t < 0.0;
      // (packets the wave engine handed over in the middle of their last cell)
      if (WAVE_DEFER_PARTIAL && !was_reabsorbed && p.tau_left < 0.0) wave_partial_step<ND>(M, p);
      bool ok = was_reabsorbed ? reemit_photon<ND>(M, p, rng, n_kill)
                               : interact_photon<ND>(M, p, rng, n_abs, n_scat, n_kill, dust_id, scattered) == 0;
      if (ok && M.use_mrw && !was_reabsorbed) ok = mrw_loop<ND>(M, p, rng, n_kill);
      if (ok) {
        p.tau_left = -log(1.0 - rng.next());
        store_photon<ND>(slots + slot, p, rng, id);
        alive = true;
      }
    }
    queue_append(alive, q_flight_next, n_flight_next, slot);
    queue_append(valid && !alive, P.q_emit, P.counts + C_NE, slot);
  }
  warp_add_scalar(M.scalars + SC_ABS, (double)n_abs);
  warp_add_scalar(M.scalars + SC_SCAT, (double)n_scat);
  warp_add_scalar(M.scalars + SC_KILLED_INT, (double)n_kill);
}

// ---------------------------------------------------------------------------------------------
// flight kernels: grid_integrate for every queued packet (persistent threads, one packet per lane)
//
// Two variants share the crossing code (advance_group):
//   flight_kernel        packets of unrelated directions (after an interaction).  Each lane refills
//                        on its own from the queue; every deposit is one RED to L2.
//   flight_beam_kernel   freshly emitted packets.  They are emitted sorted by direction, so the
//                        32 packets a warp claims together leave the source as a narrow beam and
//                        sit in the same few cells at every step: the warp marches them in
//                        lockstep, their density loads coalesce, and the deposits of all lanes in
//                        one cell are summed with shuffles before a single RED leaves the SM.
// ---------------------------------------------------------------------------------------------
constexpr int FLIGHT_THREADS = 256;
#ifndef FLIGHT_MIN_BLOCKS
#define FLIGHT_MIN_BLOCKS 3
#endif
#ifndef FLIGHT_LOOKAHEAD
#define FLIGHT_LOOKAHEAD 4   // cell crossings whose density loads are in flight together
#endif
#ifndef FLIGHT_GROUPS
#define FLIGHT_GROUPS 8      // look-ahead groups between two refill votes
#endif
#ifndef BEAM_SCAN_SPAN
#define BEAM_SCAN_SPAN 16    // longest run of equal cells summed before a RED (power of two)
#endif
#ifndef BEAM_MAX_GROUPS
#define BEAM_MAX_GROUPS 8    // distinct cells per warp step above which deposits go out lane by lane
#endif
// experiment switches (tools/sweep.sh): 0 = product path
#ifndef FLIGHT_EXPERIMENT
#define FLIGHT_EXPERIMENT 0
#endif
template <typename T>
__device__ __forceinline__ void deposit_add(T *addr, T v) {
#if FLIGHT_EXPERIMENT == 1      // no deposit at all (throughput probe)
  (void)addr; (void)v;
#else
  atomicAdd(addr, v);
#endif
}

// State of one flight in registers.
template <int ND>
struct Lane {
  double r0x, r0y, r0z;      // flight origin
  double ivx, ivy, ivz;      // 1/v, +-Inf for a ray parallel to the walls of that axis
  double t;                  // path length travelled from the origin
  double tnx, tny, tnz;      // path length at which the next x / y / z wall is reached
  double tau;                // optical depth left to the interaction
  double chi[ND], kE[ND];
  int ix, iy, iz, ic;
  int fwd;                   // bit a set: the packet moves towards +axis a
};

// Geometry of a flight that starts at r0 in cell (ix, iy, iz) with direction v.
template <int ND>
__device__ __forceinline__ void init_lane(Lane<ND> &L, double r0x, double r0y, double r0z, double vx, double vy,
                                          double vz, int ix, int iy, int iz, int ic, const double *__restrict__ W,
                                          int o2, int o3) {
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  L.r0x = r0x; L.r0y = r0y; L.r0z = r0z;
  L.t = 0.0;
  L.ix = ix; L.iy = iy; L.iz = iz; L.ic = ic;
  L.ivx = 1.0 / vx;
  L.ivy = 1.0 / vy;
  L.ivz = 1.0 / vz;
  // distance to the wall ahead on each axis; a ray parallel to an axis never reaches its walls.
  // Rounding at an interaction point can leave it a few ulp behind the wall it faces: clamp.
  L.tnx = vx != 0.0 ? fmax((W[L.ix + (vx > 0.0 ? 1 : 0)] - L.r0x) * L.ivx, 0.0) : inf;
  L.tny = vy != 0.0 ? fmax((W[o2 + L.iy + (vy > 0.0 ? 1 : 0)] - L.r0y) * L.ivy, 0.0) : inf;
  L.tnz = vz != 0.0 ? fmax((W[o3 + L.iz + (vz > 0.0 ? 1 : 0)] - L.r0z) * L.ivz, 0.0) : inf;
  L.fwd = (vx > 0.0 ? 1 : 0) | (vy > 0.0 ? 2 : 0) | (vz > 0.0 ? 4 : 0);
}

template <int ND>
__device__ __forceinline__ void load_lane(const Slot<ND> *__restrict__ s, Lane<ND> &L, const double *__restrict__ W,
                                          int o2, int o3) {
  const double2 a0 = __ldcs((const double2 *)&s->r0x);  // r0x r0y
  const double2 a1 = __ldcs((const double2 *)&s->r0z);  // r0z vx
  const double2 a2 = __ldcs((const double2 *)&s->vy);   // vy vz
  const int4 c = __ldcs((const int4 *)&s->ix);
  init_lane<ND>(L, a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, c.x, c.y, c.z, c.w, W, o2, o3);
  L.tau = __ldcs(&s->tau_left);
#pragma unroll
  for (int k = 0; k < ND; ++k) {
    L.chi[k] = __ldcs(&s->chi[k]);
    L.kE[k] = __ldcs(&s->kE[k]);
  }
}

// Deposits of a whole warp for one crossing step, summed per cell before they leave the SM.
// Must be called by all 32 lanes; `has` marks the lanes with a deposit into cell `c`.
// The packets of a beam are sorted along a space-filling curve of their direction, so lanes in
// the same cell are neighbours: a segmented suffix sum over runs of equal cell ids (5 shuffle
// steps, no loop over cells) leaves each run's total in its first lane, which issues the RED.
template <int ND>
__device__ __forceinline__ void deposit_warp(CellRec *__restrict__ cells, bool has, int c, const double (&dv)[ND]) {
  const unsigned lane = threadIdx.x & 31;
  if (__ballot_sync(0xffffffffu, has) == 0) return;
  const int key = has ? c : -1 - (int)lane;  // lanes without a deposit never join a run
  const int kp = __shfl_up_sync(0xffffffffu, key, 1);
  // runs are cut every BEAM_SCAN_SPAN lanes: log2(span) shuffle rounds instead of 5, at most 32 / span REDs more
  const bool first = (lane % BEAM_SCAN_SPAN) == 0 || kp != key;
  const unsigned heads = __ballot_sync(0xffffffffu, first);
  const unsigned above = lane == 31 ? 0u : (heads & ~((2u << lane) - 1u));
  const unsigned end = above ? (unsigned)__ffs(above) - 2u : 31u;  // last lane of my run
  double v[ND];
#pragma unroll
  for (int id = 0; id < ND; ++id) v[id] = dv[id];
#pragma unroll
  for (int o = 1; o < BEAM_SCAN_SPAN; o <<= 1) {
    const bool join = lane + o <= end;
#pragma unroll
    for (int id = 0; id < ND; ++id) {
      const double v2 = __shfl_down_sync(0xffffffffu, v[id], o);
      if (join) v[id] += v2;
    }
  }
  const bool head = has && first;
  if (head) {
#pragma unroll
    for (int id = 0; id < ND; ++id)
      if (v[id] != 0.0) deposit_add(&cells[(size_t)c * ND + id].esum, v[id]);
  }
}

// March one look-ahead group of D cell crossings (grid_propagate_3d.f90:106-232).
// Returns 0 to continue, 1 if the packet left the grid, 2 if it reached its interaction (then
// L.t / L.ix,iy,iz / L.ic describe the interaction point).  With COH set the function is called
// by the whole warp (lanes with `on` unset only take part in the shuffles).
template <int ND, int D, bool COH, bool DEP = true>
__device__ __forceinline__ int advance_group(Lane<ND> &L, const bool on, const double *__restrict__ W,
                                             CellRec *__restrict__ cells, const int n1, const int n2, const int n3,
                                             uint32_t &n_cross, const double *__restrict__ rho_only = nullptr) {
  const int o2 = n1 + 1, o3 = n1 + n2 + 2;
  // stage A: geometry of the next D crossings (independent of the density) and their loads.
  // Branch-free DDA: the axis whose wall is reached first is picked with selects.
  double tx_s[D], rho_s[D][ND];
  int ic_s[D];
  unsigned outm = 0, movedm = 0, mv = 0;
  bool dead = !on;
  double t_cur = L.t;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    ic_s[j] = L.ic;
#pragma unroll
    for (int id = 0; id < ND; ++id)
      rho_s[j][id] = DEP ? __ldcg(&cells[(size_t)L.ic * ND + id].rho) : __ldg(rho_only + (size_t)L.ic * ND + id);
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
    const bool moved = live & !out;
    t_cur = live ? t_exit : t_cur;
    tx_s[j] = t_cur;
    outm |= (live & out) ? (1u << j) : 0u;
    movedm |= moved ? (1u << j) : 0u;
    mv |= (unsigned)((bx ? 0 : (by ? 2 : 4)) + fwd) << (3 * j);
    L.ix = (moved & bx) ? i_new : L.ix;
    L.iy = (moved & by) ? i_new : L.iy;
    L.iz = (moved & !(bx | by)) ? i_new : L.iz;
    L.tnx = (moved & bx) ? tn_new : L.tnx;
    L.tny = (moved & by) ? tn_new : L.tny;
    L.tnz = (moved & !(bx | by)) ? tn_new : L.tnz;
    L.ic = moved ? (L.iz * n2 + L.iy) * n1 + L.ix : L.ic;
    dead |= out;
  }
  // stage B: optical depth and deposits, in order
  int fin = on ? 0 : 3;
  double t_prev = L.t;
  int jf = 0;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    double dv[ND];
#pragma unroll
    for (int id = 0; id < ND; ++id) dv[id] = 0.0;
    bool has = false;
    if (fin == 0) {
      const double ds = tx_s[j] - t_prev;
      double chi_rho = 0.0;
#pragma unroll
      for (int id = 0; id < ND; ++id) chi_rho += L.chi[id] * rho_s[j][id];
      const double tau_cell = chi_rho * ds;
      ++n_cross;
      double len;
      if (tau_cell < L.tau) {
        // cross the whole cell: deposit tmin * kappa * E (grid_propagate_3d.f90:148-160)
        len = ds;
        L.tau -= tau_cell;
        t_prev = tx_s[j];
        if ((outm >> j) & 1u) fin = 1;
      } else {
        // interaction inside this cell (grid_propagate_3d.f90:186-228)
        len = tau_cell > 0.0 ? ds * (L.tau / tau_cell) : 0.0;
        t_prev += len;
        fin = 2;
        jf = j;
      }
      has = true;
#pragma unroll
      for (int id = 0; id < ND; ++id) dv[id] = rho_s[j][id] > 0.0 ? len * L.kE[id] : 0.0;
    }
    if (!DEP) {
      // imaging iterations march without depositing (grid_integrate_noenergy, grid_propagate_3d.f90:237-375)
    } else if (COH) {
      deposit_warp<ND>(cells, has, ic_s[j], dv);
    } else if (has) {
#pragma unroll
      for (int id = 0; id < ND; ++id)
        if (rho_s[j][id] > 0.0) deposit_add(&cells[(size_t)ic_s[j] * ND + id].esum, dv[id]);
    }
  }
  L.t = t_prev;
  if (fin == 2) {
    // the geometry ran ahead of the interaction: step the cell indices back
#pragma unroll
    for (int j = D - 1; j >= 0; --j) {
      if (j >= jf && ((movedm >> j) & 1u)) {
        const unsigned code = (mv >> (3 * j)) & 7u;
        const int s = 2 * (int)(code & 1u) - 1;
        const unsigned ax = code >> 1;
        L.ix -= ax == 0 ? s : 0;
        L.iy -= ax == 1 ? s : 0;
        L.iz -= ax == 2 ? s : 0;
      }
    }
#pragma unroll
    for (int j = 0; j < D; ++j)
      if (j == jf) L.ic = ic_s[j];
  }
  return fin == 3 ? 0 : fin;
}

__device__ __forceinline__ const double *stage_walls(const ModelDev &M, double *s_walls, int walls_in_smem) {
  // wall table: shared memory when it fits (always for the grids of BASELINE.json), else global;
  // w1|w2|w3 are contiguous in global memory as well
  if (!walls_in_smem) return M.w1;
  for (int i = threadIdx.x; i < M.n1 + M.n2 + M.n3 + 3; i += blockDim.x) s_walls[i] = M.w1[i];
  __syncthreads();
  return s_walls;
}

template <int ND>
__device__ __forceinline__ void store_flight_result(Slot<ND> *s, const Lane<ND> &L) {
  __stcs(&s->t, L.t);
  __stcs((int4 *)&s->ix, make_int4(L.ix, L.iy, L.iz, L.ic));
}

template <int ND, int D>
__global__ void __launch_bounds__(FLIGHT_THREADS, FLIGHT_MIN_BLOCKS)
flight_kernel(const ModelDev M, Pool P, const uint32_t *__restrict__ q_flight, const uint32_t *n_flight_ptr,
              const int walls_in_smem) {
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
  uint32_t n_cross = 0, n_esc = 0;
  unsigned long long cross_hi = 0;

  for (;;) {
    // ---------------- refill idle lanes from the flight queue ----------------
    const bool need = !active && !exhausted;
    const unsigned m_need = __ballot_sync(0xffffffffu, need);
    if (m_need) {
      const int leader = __ffs(m_need) - 1;
      uint32_t base = 0;
      if ((int)lane == leader) base = atomicAdd(P.counts + C_CURSOR, (uint32_t)__popc(m_need));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (need) {
        const uint32_t idx = base + __popc(m_need & ((1u << lane) - 1u));
        if (idx >= n_flight) {
          exhausted = true;
        } else {
          slot = q_flight[idx];
          load_lane<ND>(slots + slot, L, W, n1 + 1, n1 + n2 + 2);
          L.t = __ldcs(&slots[slot].t);  // 0 for a new flight; the path length so far for one handed over by the wave engine
          active = true;
        }
      }
    }
    if (__ballot_sync(0xffffffffu, active) == 0) break;

    // ---------------- cell crossings ----------------
    int fin = 0;
    if (active) {
#pragma unroll 1
      for (int g = 0; g < FLIGHT_GROUPS; ++g) {
        fin = advance_group<ND, D, false>(L, true, W, cells, n1, n2, n3, n_cross);
        if (fin) break;
      }
      if (n_cross > 0x7fffff00u) {
        cross_hi += n_cross;
        n_cross = 0;
      }
    }

    // ---------------- hand finished packets to the next kernel ----------------
    if (fin == 2) store_flight_result<ND>(slots + slot, L);
    queue_append(fin == 2, P.q_interact, P.counts + C_NI, slot);
    queue_append(fin == 1, P.q_emit, P.counts + C_NE, slot);
    if (fin) {
      n_esc += fin == 1 ? 1u : 0u;
      active = false;
    }
  }
  warp_add_scalar(M.scalars + SC_CROSS, (double)(cross_hi + n_cross));
  warp_add_scalar(M.scalars + SC_ESC, (double)n_esc);
}

template <int ND, int D>
__global__ void __launch_bounds__(FLIGHT_THREADS, FLIGHT_MIN_BLOCKS)
flight_beam_kernel(const ModelDev M, Pool P, const uint32_t *__restrict__ q_beam, const uint32_t *n_beam_ptr,
                   const int walls_in_smem) {
  extern __shared__ double s_walls[];
  const int n1 = M.n1, n2 = M.n2, n3 = M.n3;
  const double *__restrict__ W = stage_walls(M, s_walls, walls_in_smem);
  const uint32_t n_beam = *n_beam_ptr;
  Slot<ND> *slots = (Slot<ND> *)P.slots;
  CellRec *__restrict__ cells = M.cells;
  const unsigned lane = threadIdx.x & 31;
  uint32_t n_cross = 0, n_esc = 0;
  unsigned long long cross_hi = 0;

  for (;;) {
    // the warp claims 32 consecutive packets of the emission order: one beam
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(P.counts + C_CURSOR_B, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= n_beam) break;
    const uint32_t idx = base + lane;
    uint32_t slot = 0;
    Lane<ND> L;
    L.ic = 0;
    int fin = 3;  // 3: no packet in this lane
    if (idx < n_beam) {
      slot = q_beam[idx];
      load_lane<ND>(slots + slot, L, W, n1 + 1, n1 + n2 + 2);
      fin = 0;
    }
    // lockstep march until every lane has escaped or interacted
    while (__ballot_sync(0xffffffffu, fin == 0)) {
      const int f = advance_group<ND, D, true>(L, fin == 0, W, cells, n1, n2, n3, n_cross);
      if (fin == 0) fin = f;
    }
    if (n_cross > 0x7fffff00u) {
      cross_hi += n_cross;
      n_cross = 0;
    }
    if (fin == 2) store_flight_result<ND>(slots + slot, L);
    queue_append(fin == 2, P.q_interact, P.counts + C_NI, slot);
    queue_append(fin == 1, P.q_emit, P.counts + C_NE, slot);
    n_esc += fin == 1 ? 1u : 0u;
  }
  warp_add_scalar(M.scalars + SC_CROSS, (double)(cross_hi + n_cross));
  warp_add_scalar(M.scalars + SC_ESC, (double)n_esc);
}

// =============================================================================================
// streaming kernels around the photon loop
// =============================================================================================

// grid_reset_energy (src/grid/grid_generic.f90:21-27) + precompute_jnu_var
// (src/grid/grid_physics_3d.f90:613-629, dust_jnu_var_pos_frac dust_type_4elem.f90:295-320)
__global__ void lucy_begin_kernel(ModelDev M, const int reset_sums) {
  const int nd = M.n_dust;
  const int64_t n = M.n_cells * nd;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    if (reset_sums) M.cells[k].esum = 0.0;
    const int id = (int)(k % nd);
    const DustDev &d = M.dust[id];
    const double e = M.specific_energy[k];
    int jid;
    double frac;
    if (e < d.L.jvar_min) {
      jid = 0;
      frac = 0.0;
    } else if (e > d.L.jvar_max) {
      jid = d.L.n_jnu - 2;
      frac = 1.0;
    } else {
      jid = lower_interval(d.B + d.L.o_jvar, d.L.n_jnu, e);
      const double *lj = d.B + d.L.o_logjvar;
      const double a = __ldg(lj + jid), b = __ldg(lj + jid + 1);
      frac = (log10(e) - a) / (b - a);
    }
    M.jnu_id[k] = jid;
    M.jnu_frac[k] = frac;
  }
}

// update_alpha_inv_planck (grid_physics_3d.f90:397-418) + prepare_mrw (grid_mrw_3d.f90:29-54)
__global__ void mrw_prepare_kernel(ModelDev M) {
  const int nd = M.n_dust;
  for (int64_t ic = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ic < M.n_cells; ic += (int64_t)gridDim.x * blockDim.x) {
    double alpha = 0.0;
    for (int id = 0; id < nd; ++id) {
      const double rho = M.cells[ic * nd + id].rho;
      if (rho > 0.0) alpha += rho * mean_opacity_loglog(M.dust[id], M.dust[id].L.o_logchi_invp, M.specific_energy[ic * nd + id]);
    }
    M.alpha_inv_planck[ic] = alpha;
    M.diff_coeff[ic] = 1.0 / 3.0 / alpha;
  }
}

// gather the deposit sums into the contiguous reduction buffer
__global__ void gather_sums_kernel(ModelDev M, double *__restrict__ sums) {
  const int64_t n = M.n_cells * M.n_dust;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    sums[k] = M.cells[k].esum;
}

__device__ __forceinline__ double clamp_energy(const ModelDev &M, const DustDev &d, int id, double e) {
  // check_energy_abs (src/grid/grid_physics_3d.f90:555-603)
  if (e < M.min_energy[id]) e = M.min_energy[id];
  if (M.enforce_energy_range) {
    if (e < d.L.e_min) e = d.L.e_min;
    if (e > d.L.e_max) e = d.L.e_max;
  }
  return e;
}

// update_energy_abs (grid_physics_3d.f90:500-553) + sublimate_dust (:420-498)
// mode 0: both; 1: update_energy_abs only; 2: sublimate_dust only (the PDA runs between the two, iter_lucy.f90:224-235)
__global__ void lucy_finish_kernel(ModelDev M, const double *__restrict__ sums, double scale, const int mode = 0) {
  const int nd = M.n_dust;
  const int64_t n = M.n_cells * nd;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const int id = (int)(k % nd);
    const int64_t ic = k / nd;
    const double vol = cell_volume(M, ic);
    const DustDev &d = M.dust[id];
    double e;
    const int nb = M.spec_sums ? M.n_spec_bins : 0;
    if (mode == 2) {
      e = M.specific_energy[k];
    } else {
      e = sums[k] * scale / vol;
      if (vol == 0.0) e = 0.0;
      // update_energy_abs, spectrum part (:516-524); the additional spectrum is all zeros (:143,223-225)
      for (int b = 0; b < nb; ++b) {
        double eb = M.spec_sums[(size_t)b * n + k] * scale / vol;
        if (vol == 0.0) eb = 0.0;
        M.spec_energy[(size_t)b * n + k] = eb;
      }
      if (M.energy_additional) e = e + M.energy_additional[k];   // grid_physics_3d.f90:537-545
      e = clamp_energy(M, d, id, e);
    }
    if (mode != 1 && d.L.sublimation_mode != 0 && e > d.L.sublimation_specific_energy) {
      const double es = d.L.sublimation_specific_energy;
      // the spectrum follows: reset to the minimum (:442-446) or rescaled with its shape kept (:463,479)
      for (int b = 0; b < nb; ++b) {
        double &eb = M.spec_energy[(size_t)b * n + k];
        eb = d.L.sublimation_mode == 1 ? M.min_energy[id] : eb * (es / e);
      }
      if (d.L.sublimation_mode == 1) {
        M.cells[k].rho = 0.0;
        M.rho[k] = 0.0;
        e = M.min_energy[id];
      } else if (d.L.sublimation_mode == 2) {
        const double q = mean_opacity_loglog(d, d.L.o_logchi_ross, e) / mean_opacity_loglog(d, d.L.o_logchi_ross, es);
        M.cells[k].rho = M.cells[k].rho * es / e * (q * q);
        M.rho[k] = M.cells[k].rho;
        e = es;
      } else {
        e = es;
      }
      e = clamp_energy(M, d, id, e);
    }
    M.specific_energy[k] = e;
  }
}

// initial check_energy_abs of setup_grid_physics (grid_physics_3d.f90:291)
__global__ void clamp_energy_kernel(ModelDev M) {
  const int nd = M.n_dust;
  const int64_t n = M.n_cells * nd;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const int id = (int)(k % nd);
    M.specific_energy[k] = clamp_energy(M, M.dust[id], id, M.specific_energy[k]);
  }
}

// layout conversion between the file order [n_dust][n_cells] and the device order [n_cells][n_dust]
__global__ void scatter_density_kernel(ModelDev M, const double *__restrict__ in) {
  const int nd = M.n_dust;
  const int64_t nc = M.n_cells, n = nc * nd;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const int id = (int)(k / nc);
    const int64_t ic = k % nc;
    M.cells[ic * nd + id].rho = in[k];
    M.rho[ic * nd + id] = in[k];
  }
}
__global__ void to_device_order_kernel(int nd, int64_t nc, const double *__restrict__ in, double *__restrict__ out) {
  const int64_t n = nc * nd;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const int id = (int)(k / nc);
    const int64_t ic = k % nc;
    out[ic * nd + id] = in[k];
  }
}
// which: 0 specific_energy, 1 density, 2 reduced sums
__global__ void to_file_order_kernel(ModelDev M, int which, const double *__restrict__ sums, double *__restrict__ out) {
  const int nd = M.n_dust;
  const int64_t nc = M.n_cells, n = nc * nd;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const int id = (int)(k / nc);
    const int64_t ic = k % nc;
    const int64_t s = ic * nd + id;
    out[k] = which == 0 ? M.specific_energy[s] : (which == 1 ? M.cells[s].rho : sums[s]);
  }
}

#include "march_geo.cuh"
#include "imaging.cuh"
#include "flight_geo.cuh"
#include "flight_wave.cuh"

// =============================================================================================
// host side: context + C ABI
// =============================================================================================
namespace {

thread_local std::string g_error;

struct HostDust {
  DustLayout L;
  std::vector<double> buf;
  double *dev = nullptr;
  // raw columns kept for the raytracing spectra (get_chi_nu_binned / get_j_nu_binned)
  std::vector<double> nu, chi, albedo, emiss_nu, emiss_jnu;
  int n_jnu = 0;
  // b_nu samplers for the modified random walk (built at finalize when the MRW is on)
  DustMrwLayout Lm;
  double *dev_mrw = nullptr;
};
struct HostSpectrum {
  SpectrumLayout L;
  std::vector<double> buf;
  double *dev = nullptr;
  std::vector<double> nu, fnu;  // raw table (get_spectrum_binned)
};
// one peeled group (peeled_images_setup, src/images/images_peeled.f90:272-382)
struct HostImage {
  hyp_image_conf conf;
  std::vector<double> theta, phi;
  int n_orig = 1, n_stokes = 4;
  double nu_min = 0, nu_max = 0;
  size_t n_sed = 0, n_img = 0;     // elements of one SED / image cube
  size_t o_sed = 0, o_img = 0;     // offsets into the image buffer: [val | sum of squares | count] each
  std::vector<int32_t> filt_off;   // use_filters: points of channel k are [filt_off[k], filt_off[k+1])
  std::vector<double> filt_nu, filt_tr;
};

}  // namespace

struct hyp_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;              // the beam kernel of a round runs here, next to the flight kernel
  cudaStream_t stream3 = nullptr;              // wave engine: emission next to the tile visits and the interactions
  cudaEvent_t evFork = nullptr, evJoin = nullptr, evJoin3 = nullptr;
  WaveQ wave = WaveQ();                        // wave engine (flight_wave.cuh)
  uint32_t *wave_sorted[2] = {nullptr, nullptr};  // the sorted slot lists of two consecutive rounds
  int wave_bins_alloc = 0;
  bool uniform_walls = false;                  // all three wall arrays equidistant (to 1e-10 of the spacing)
  int last_engine = 0;                         // 1: the last Lucy photon loop ran on the wave engine
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  // host model
  int grid_type = GEO_CAR;
  int polar_kind = POLAR_SPH;  // grid_type == GEO_SPH: spherical or cylindrical polar
  double *d_sph = nullptr;  // spherical polar tables
  // octree (host copies until finalize)
  std::vector<OctNode> oct_nodes;
  std::vector<int32_t> oct_children, oct_leaves;
  double oct_eps = 0.0;
  // AMR (host copies until finalize)
  std::vector<AmrGridDev> amr_grids;
  std::vector<int32_t> amr_gotos, amr_cell_grid, amr_valid;
  int amr_n_level1 = 0;
  double amr_eps = 0.0;
  AmrGridDev *d_amr_grids = nullptr;
  int32_t *d_amr_gotos = nullptr, *d_amr_cell_grid = nullptr, *d_amr_valid = nullptr;
  OctNode *d_oct_nodes = nullptr;
  int32_t *d_oct_children = nullptr, *d_oct_leaves = nullptr;
  // frequency-resolved specific energy
  std::vector<double> spec_edges;
  double *d_spec_energy = nullptr, *d_spec_tab = nullptr;   // [bins][n] ; log edges | j_nu_bin_frac
  // Voronoi mesh (host copies until finalize)
  std::vector<double> vor_sites, vor_bb, vor_volume;
  std::vector<int32_t> vor_nidx, vor_neigh, vor_valid, vor_b_start, vor_b_sites;
  double vor_box[6] = {0, 0, 0, 0, 0, 0};
  int vor_nb = 1;
  double *d_vor_f64 = nullptr;    // sites | bounding boxes | volumes
  int32_t *d_vor_i32 = nullptr;   // nidx | neigh | valid | b_start | b_sites
  int n1 = 0, n2 = 0, n3 = 0;
  int64_t n_cells = 0;
  std::vector<double> w1, w2, w3;
  std::vector<HostDust> dust;
  std::vector<hyp_source> sources;
  std::vector<HostSpectrum> spectra;
  std::vector<int> source_spectrum;
  hyp_run_conf conf;
  std::vector<double> h_density, h_energy, h_min_energy;
  bool have_density = false, have_energy = false, energy_from_caller = false;
  double energy_total = 0.0;
  // device
  double *d_w = nullptr;
  CellRec *d_cells = nullptr;
  double *d_rho = nullptr;
  double *d_energy = nullptr, *d_jfrac = nullptr, *d_sums = nullptr, *d_stage = nullptr;
  int32_t *d_jid = nullptr;
  SourceDev *d_sources = nullptr;
  SpectrumDev *d_spectra = nullptr;
  unsigned long long *d_work = nullptr;
  int32_t *d_error = nullptr;
  double *h_pinned = nullptr;  // pinned staging for grid transfers
  ModelDev M;
  bool finalized = false;
  bool sums_gathered = false;
  float kernel_ms_acc = 0.f, flight_ms_acc = 0.f;
  int64_t rounds_acc = 0, wave_rounds_acc = 0, launches_acc = 0;  // launches: this library's own kernels in the iteration
  // photon pool
  Pool pool = Pool();
  std::vector<int64_t> coll_off;              // per source: first entry of its point collection, -1 if none
  std::vector<double> coll_xyz, coll_cdf;     // point collections of all sources
  double *d_coll_xyz = nullptr, *d_coll_cdf = nullptr;
  std::vector<int64_t> map_off;               // per source: first entry of its cumulative luminosity map, -1 if none
  std::vector<double> map_cdf;                // released after the upload
  double *d_map_cdf = nullptr;
  double *d_energy_add = nullptr;             // specific_energy_type = 'additional
  std::vector<int> spot_off;                  // per source: first entry in spots
  std::vector<SpotDev> spots;
  SpotDev *d_spots = nullptr;
  uint32_t pool_cap = 0;
  uint32_t *h_counts = nullptr;  // pinned: [C_COUNT] counters + next_photon (2 words)
  // emission-order sort (direction keys)
  uint32_t sort_window = 0;
  uint32_t *d_keys_in = nullptr, *d_keys_out = nullptr, *d_vals_in = nullptr, *d_perm = nullptr;
  void *d_sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  cudaEvent_t evA = nullptr, evB = nullptr, evCtl = nullptr;
  // imaging (final / raytracing iterations)
  std::vector<HostImage> groups;
  double *d_imgbuf = nullptr;       // all cubes of all groups, then SC_COUNT scalars
  size_t imgbuf_n = 0;
  ImageDev *d_images = nullptr;
  ViewDev *d_views = nullptr;
  int n_views = 0;
  std::vector<ImageDev> h_images;
  void *d_jobs = nullptr;
  uint32_t *d_njobs = nullptr;
  uint32_t job_cap = 0;
  double *d_eabs = nullptr;
  double *d_mrw_alpha = nullptr, *d_mrw_diff = nullptr, *d_mrw_cdf = nullptr;
  double *d_src_columns = nullptr;
  std::vector<double *> ray_tables;  // device copies of the binned raytracing spectra
  std::vector<void *> filter_tables; // device copies of the filter curves
  bool images_ready = false, ray_ready = false;
  int64_t peel_launches = 0;
  // monochromatic mode (hyp_set_monochromatic)
  std::vector<double> frequencies;
  double mono_threshold = 1.e-10;
  bool mono_run = false;                      // the cubes of this final iteration were filled by hyp_final_mono_photons
  double *d_mono_src_w = nullptr, *d_mono_spot_w = nullptr, *d_mono_cdf = nullptr;
  double *d_mono_logp[MAX_DUST] = {nullptr, nullptr, nullptr, nullptr};
  void *d_scan_tmp = nullptr;
  size_t scan_tmp_bytes = 0;
  std::vector<cudaEvent_t> wave_ev;           // wave engine: start / end of the tile kernel of every round
  // packet counter and partial diffusion approximation (pda.cuh)
  unsigned long long *d_nvis = nullptr, *d_lastid = nullptr;
  double *d_pda_geo = nullptr, *d_pda_e = nullptr, *d_pda_coef = nullptr, *d_pda_counts = nullptr;
  int32_t *d_pda_list = nullptr;
  unsigned long long *d_pda_ctl = nullptr;    // [0] number of PDA cells (32 bits used), [1] largest change, [2] sum of counts
  int64_t pda_cells_last = 0, pda_sweeps_last = 0;
};

namespace {

int fail(int code, const std::string &msg) {
  g_error = msg;
  return code;
}

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return fail(HYP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));            \
  } while (0)

int grid_blocks(const hyp_ctx *c) { return c->sm_count * 8; }

int device_error_to_status(hyp_ctx *c) {
  int32_t flag = 0;
  CUDA_TRY(cudaMemcpyAsync(&flag, c->d_error, sizeof flag, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  switch (flag) {
    case ERR_NONE:
      return HYP_OK;
    case ERR_NOT_IN_CELL:
      return fail(HYP_ERR_PHYSICS,
                  "photon was not emitted inside a cell - this usually indicates that a source is not inside the grid");
    case ERR_NU_RANGE:
      return fail(HYP_ERR_PHYSICS,
                  "photon frequency is outside the range defined for the dust optical properties");
    case ERR_DEPOSIT:
      return fail(HYP_ERR_STATE, "internal error: a packet's kappa * energy exceeds the bound of the fixed-point deposits "
                                 "(set HYPERION_B200_ENGINE=rounds and report)");
    case ERR_JOBS:
      return fail(HYP_ERR_STATE, "internal error: the peel-off queue of a round overflowed");
    default:
      return fail(HYP_ERR_PHYSICS, "ERROR: in sampling mu for scattering");
  }
}

template <typename T>
void free_dev(T *&p) {
  if (p) cudaFree(p);
  p = nullptr;
}

void free_pool(hyp_ctx *c) {
  free_dev(c->pool.slots);
  free_dev(c->pool.q_flight[0]);
  free_dev(c->pool.q_flight[1]);
  free_dev(c->pool.q_beam);
  free_dev(c->pool.q_interact);
  free_dev(c->pool.q_emit);
  free_dev(c->pool.counts);
  free_dev(c->pool.next_photon);
  free_dev(c->wave.key);
  free_dev(c->wave_sorted[0]);
  free_dev(c->wave_sorted[1]);
  c->wave.sorted = nullptr;
  free_dev(c->wave.bin_count);
  free_dev(c->wave.bin_cursor);
  free_dev(c->wave.items);
  free_dev(c->wave.ctl);
  c->wave.capacity = 0;
  c->wave_bins_alloc = 0;
  c->pool_cap = 0;
  free_dev(c->d_keys_in);
  free_dev(c->d_keys_out);
  free_dev(c->d_vals_in);
  free_dev(c->d_perm);
  free_dev(c->d_sort_tmp);
  c->sort_window = 0;
}

size_t slot_bytes(int nd) {
  switch (nd) {
    case 1: return sizeof(Slot<1>);
    case 2: return sizeof(Slot<2>);
    case 3: return sizeof(Slot<3>);
    default: return sizeof(Slot<4>);
  }
}

// Packets resident in the pool at once (never more than the launch holds).  Measured on the 256^3 headline
// (profiles/r01_experiments.md): 2 M slots 91.7 ms per 2e7 packets, 8 M 87.5 ms, 16 M 84.9 ms, 20 M 84.6 ms --
// fewer, fuller rounds amortise the launch tails; with the beam and flight kernels side by side 12 M is best
// (77.3 ms): the second round then has both kinds of packets in quantity.  12 M slots are 2.4 GB of the 180 GB.
// The wave engine (flight_wave.cuh) wants twice as many: its rounds are tile visits, and fuller rounds amortise
// the staging of the tiles (12 M slots 62.8 ms per 2e7 packets, 24 M 58.9 ms, 32-48 M 58.8 ms; profiles/r02_experiments.md).
uint32_t pool_target(bool wave = false) {
  const char *e = getenv("HYPERION_B200_POOL");
  long v = e ? atol(e) : (wave ? 25165824L : 12582912L);
  if (v < 1024) v = 1024;
  if (v > (1L << 28)) v = 1L << 28;
  return (uint32_t)v;
}

int ensure_pool(hyp_ctx *c, uint32_t cap) {
  if (c->pool_cap >= cap) return HYP_OK;
  free_pool(c);
  Pool &P = c->pool;
  CUDA_TRY(cudaMalloc(&P.slots, (size_t)cap * slot_bytes(c->M.n_dust)));
  CUDA_TRY(cudaMalloc(&P.q_flight[0], (size_t)cap * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&P.q_flight[1], (size_t)cap * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&P.q_beam, (size_t)cap * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&P.q_interact, (size_t)cap * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&P.q_emit, (size_t)cap * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&P.counts, C_COUNT * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&P.next_photon, sizeof(unsigned long long)));
  P.capacity = cap;
  c->pool_cap = cap;
  // emission-order windows: a power of two >= 2 * cap so that the ring of two windows always
  // covers every id the next round can claim
  uint32_t w = 1u << 24;
  while (w < 2ull * cap) w <<= 1;
  c->sort_window = w;
  CUDA_TRY(cudaMalloc(&c->d_keys_in, (size_t)w * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&c->d_keys_out, (size_t)w * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&c->d_vals_in, (size_t)w * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&c->d_perm, 2 * (size_t)w * sizeof(uint32_t)));
  c->sort_tmp_bytes = 0;
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, c->sort_tmp_bytes, c->d_keys_in, c->d_keys_out, c->d_vals_in,
                                           c->d_perm, (int)w, 0, 32, c->stream));
  CUDA_TRY(cudaMalloc(&c->d_sort_tmp, c->sort_tmp_bytes));
  P.perm = c->d_perm;
  P.window = w;
  return HYP_OK;
}

// Emission order of window `w` of a launch: sort the packets of the window by direction key.
int prepare_window(hyp_ctx *c, int64_t first_id, int64_t n_photons, int64_t iteration, int64_t w,
                   const ModelDev *Mo = nullptr, const int force_mode = -1) {
  const int64_t win = c->sort_window;
  const int64_t count = std::min<int64_t>(win, n_photons - w * win);
  if (count <= 0) return HYP_OK;
  // HYPERION_B200_SORT: 0 emission in id order, 1 sorted by direction, 2 by direction bin and path length
  const char *e = getenv("HYPERION_B200_SORT");
  const int mode = force_mode >= 0 ? force_mode : e ? atoi(e) : HYP_DEFAULT_SORT;
  const bool sorted = mode != 0;
  uint32_t *dst = c->d_perm + (w & 1) * win;
  auto keys = c->M.n_dust == 1 ? emit_keys_kernel<1> : c->M.n_dust == 2 ? emit_keys_kernel<2>
              : c->M.n_dust == 3 ? emit_keys_kernel<3> : emit_keys_kernel<4>;
  keys<<<c->sm_count * 8, 256, 0, c->stream>>>(Mo ? *Mo : c->M, (unsigned long long)(first_id + w * win), (uint32_t)count,
                                               (uint32_t)iteration, c->d_keys_in, sorted ? c->d_vals_in : dst, mode);
  CUDA_TRY(cudaGetLastError());
  c->launches_acc += 1;
  if (sorted)
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(c->d_sort_tmp, c->sort_tmp_bytes, c->d_keys_in, c->d_keys_out,
                                             c->d_vals_in, dst, (int)count, 0, 32, c->stream));
  return HYP_OK;
}

}  // namespace

extern "C" {

const char *hyp_last_error(void) { return g_error.c_str(); }
int hyp_version(void) { return 100; }
int hyp_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(hyp_dust_tables);
    case 1: return (int)sizeof(hyp_source);
    case 2: return (int)sizeof(hyp_run_conf);
    case 3: return (int)sizeof(hyp_iter_stats);
    default: return -1;
  }
}

int hyp_ctx_create(int device_id, hyp_ctx **out) {
  if (!out) return fail(HYP_ERR_INVALID, "out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(HYP_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                  "); this engine has no CPU fallback");
  if (device_id < 0 || device_id >= n) return fail(HYP_ERR_INVALID, "device id out of range");
  CUDA_TRY(cudaSetDevice(device_id));
  std::unique_ptr<hyp_ctx> c(new hyp_ctx());
  c->device = device_id;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device_id));
  c->sm_count = prop.multiProcessorCount;
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreate(&c->ev0));
  CUDA_TRY(cudaEventCreate(&c->ev1));
  CUDA_TRY(cudaEventCreate(&c->ev2));
  CUDA_TRY(cudaEventCreate(&c->ev3));
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming));
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&c->evJoin3, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreate(&c->evA));
  CUDA_TRY(cudaEventCreateWithFlags(&c->evCtl, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreate(&c->evB));
  CUDA_TRY(cudaMallocHost(&c->h_counts, (C_COUNT + 2) * sizeof(uint32_t)));
  memset(&c->pool, 0, sizeof c->pool);
  // Cell records are touched one 32-byte sector at a time at random: do not let L2 fetch more.
  {
    const char *g = getenv("HYPERION_B200_L2_FETCH");
    size_t gran = g ? (size_t)atoi(g) : 32;
    if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    cudaGetLastError();
  }
  memset(&c->conf, 0, sizeof c->conf);
  c->conf.seed = -124902;
  c->conf.n_inter_max = 1000000;
  c->conf.n_reabs_max = 1000000;
  c->conf.enforce_energy_range = 1;
  c->conf.propagation_check_frequency = 1e-3;
  memset(&c->M, 0, sizeof c->M);
  *out = c.release();
  return HYP_OK;
}

void hyp_ctx_destroy(hyp_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  free_dev(c->d_w);
  free_dev(c->d_sph);
  free_dev(c->d_oct_nodes);
  free_dev(c->d_amr_grids);
  free_dev(c->d_amr_gotos);
  free_dev(c->d_amr_cell_grid);
  free_dev(c->d_amr_valid);
  free_dev(c->d_oct_children);
  free_dev(c->d_oct_leaves);
  free_dev(c->d_vor_f64);
  free_dev(c->d_vor_i32);
  free_dev(c->d_spec_energy);
  free_dev(c->d_spec_tab);
  free_dev(c->d_cells);
  free_dev(c->d_rho);
  free_dev(c->d_energy);
  free_dev(c->d_jfrac);
  free_dev(c->d_sums);
  free_dev(c->d_stage);
  free_dev(c->d_jid);
  free_dev(c->d_sources);
  free_dev(c->d_coll_xyz);
  free_dev(c->d_coll_cdf);
  free_dev(c->d_map_cdf);
  free_dev(c->d_energy_add);
  free_dev(c->d_spots);
  free_dev(c->d_spectra);
  free_dev(c->d_work);
  free_dev(c->d_error);
  free_dev(c->d_imgbuf);
  free_dev(c->d_images);
  free_dev(c->d_views);
  free_dev(c->d_jobs);
  free_dev(c->d_njobs);
  free_dev(c->d_eabs);
  free_dev(c->d_mrw_alpha);
  free_dev(c->d_mrw_diff);
  free_dev(c->d_mrw_cdf);
  free_dev(c->d_src_columns);
  for (auto &t : c->ray_tables) free_dev(t);
  for (auto &t : c->filter_tables) free_dev(t);
  free_pool(c);
  if (c->h_counts) cudaFreeHost(c->h_counts);
  if (c->evFork) cudaEventDestroy(c->evFork);
  if (c->evJoin) cudaEventDestroy(c->evJoin);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->evJoin3) cudaEventDestroy(c->evJoin3);
  if (c->stream3) cudaStreamDestroy(c->stream3);
  if (c->evA) cudaEventDestroy(c->evA);
  if (c->evCtl) cudaEventDestroy(c->evCtl);
  if (c->evB) cudaEventDestroy(c->evB);
  for (cudaEvent_t e : c->wave_ev) cudaEventDestroy(e);
  free_dev(c->d_nvis); free_dev(c->d_lastid); free_dev(c->d_pda_geo); free_dev(c->d_pda_e); free_dev(c->d_pda_coef);
  free_dev(c->d_pda_counts); free_dev(c->d_pda_list); free_dev(c->d_pda_ctl);
  for (auto &d : c->dust) {
    free_dev(d.dev);
    free_dev(d.dev_mrw);
  }
  for (auto &s : c->spectra) free_dev(s.dev);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->ev2) cudaEventDestroy(c->ev2);
  if (c->ev3) cudaEventDestroy(c->ev3);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

void *hyp_stream(hyp_ctx *c) { return c ? (void *)c->stream : nullptr; }

int hyp_set_grid_cartesian(hyp_ctx *c, int32_t n1, int32_t n2, int32_t n3, const double *w1, const double *w2,
                           const double *w3) {
  if (!c || !w1 || !w2 || !w3) return fail(HYP_ERR_INVALID, "NULL argument");
  if (c->finalized) return fail(HYP_ERR_STATE, "model is frozen");
  if (n1 < 1 || n2 < 1 || n3 < 1) return fail(HYP_ERR_INVALID, "grid needs at least one cell per axis");
  if ((int64_t)n1 * n2 * n3 > 2000000000LL) return fail(HYP_ERR_INVALID, "grid too large for 32-bit cell ids");
  const double *ws[3] = {w1, w2, w3};
  const int ns[3] = {n1, n2, n3};
  const char *names[3] = {"dx", "dy", "dz"};
  for (int a = 0; a < 3; ++a)
    for (int i = 0; i < ns[a]; ++i)
      if (!(ws[a][i + 1] - ws[a][i] > 0.0))
        return fail(HYP_ERR_INVALID, std::string("all ") + names[a] + " values should be greater than zero");
  c->grid_type = GEO_CAR;
  c->n1 = n1;
  c->n2 = n2;
  c->n3 = n3;
  c->n_cells = (int64_t)n1 * n2 * n3;
  c->w1.assign(w1, w1 + n1 + 1);
  c->w2.assign(w2, w2 + n2 + 1);
  c->w3.assign(w3, w3 + n3 + 1);
  // equidistant walls on all three axes (to 1e-10 of the spacing): the wave engine applies (flight_wave.cuh)
  c->uniform_walls = true;
  const std::vector<double> *wa[3] = {&c->w1, &c->w2, &c->w3};
  for (int a = 0; a < 3; ++a) {
    const std::vector<double> &w = *wa[a];
    const int n = (int)w.size() - 1;
    const double d = (w[n] - w[0]) / n;
    if (!(d > 0.0) || !std::isfinite(d)) c->uniform_walls = false;
    for (int i = 0; i < n && c->uniform_walls; ++i)
      if (!(std::fabs((w[i + 1] - w[i]) - d) <= 1e-10 * d)) c->uniform_walls = false;
  }
  return HYP_OK;
}

int hyp_set_grid_spherical(hyp_ctx *c, int32_t n1, int32_t n2, int32_t n3, const double *w1, const double *w2,
                           const double *w3) {
  if (!c || !w1 || !w2 || !w3) return fail(HYP_ERR_INVALID, "NULL argument");
  if (c->finalized) return fail(HYP_ERR_STATE, "model is frozen");
  if (n1 < 1 || n2 < 1 || n3 < 1) return fail(HYP_ERR_INVALID, "grid needs at least one cell per axis");
  if ((int64_t)n1 * n2 * n3 > 2000000000LL) return fail(HYP_ERR_INVALID, "grid too large for 32-bit cell ids");
  const double pi = SPH_PI;
  for (int i = 0; i <= n1; ++i)
    if (w1[i] < 0.) return fail(HYP_ERR_INVALID, "r walls should be positive");
  for (int i = 0; i <= n2; ++i)
    if (w2[i] < 0. || w2[i] > pi) return fail(HYP_ERR_INVALID, "theta walls should be between 0 and pi");
  for (int i = 0; i <= n3; ++i)
    if (w3[i] < 0. || w3[i] > pi + pi) return fail(HYP_ERR_INVALID, "phi walls should be between 0 and 2*pi");
  const double *ws[3] = {w1, w2, w3};
  const int ns[3] = {n1, n2, n3};
  const char *names[3] = {"dr", "dt", "dphi"};
  for (int a = 0; a < 3; ++a)
    for (int i = 0; i < ns[a]; ++i)
      if (!(ws[a][i + 1] - ws[a][i] > 0.0))
        return fail(HYP_ERR_INVALID, std::string("all ") + names[a] + " values should be greater than zero");
  c->grid_type = GEO_SPH;
  c->polar_kind = POLAR_SPH;
  c->n1 = n1;
  c->n2 = n2;
  c->n3 = n3;
  c->n_cells = (int64_t)n1 * n2 * n3;
  c->w1.assign(w1, w1 + n1 + 1);
  c->w2.assign(w2, w2 + n2 + 1);
  c->w3.assign(w3, w3 + n3 + 1);
  return HYP_OK;
}

int hyp_set_grid_cylindrical(hyp_ctx *c, int32_t n1, int32_t n2, int32_t n3, const double *w1, const double *w2,
                             const double *w3) {
  if (!c || !w1 || !w2 || !w3) return fail(HYP_ERR_INVALID, "NULL argument");
  if (c->finalized) return fail(HYP_ERR_STATE, "model is frozen");
  if (n1 < 1 || n2 < 1 || n3 < 1) return fail(HYP_ERR_INVALID, "grid needs at least one cell per axis");
  if ((int64_t)n1 * n2 * n3 > 2000000000LL) return fail(HYP_ERR_INVALID, "grid too large for 32-bit cell ids");
  for (int i = 0; i <= n1; ++i)
    if (w1[i] < 0.) return fail(HYP_ERR_INVALID, "w walls should be positive");
  for (int i = 0; i <= n3; ++i)
    if (w3[i] < 0. || w3[i] > SPH_TWOPI) return fail(HYP_ERR_INVALID, "phi walls should be between 0 and 2*pi");
  const double *ws[3] = {w1, w2, w3};
  const int ns[3] = {n1, n2, n3};
  const char *names[3] = {"dw", "dz", "dphi"};
  for (int a = 0; a < 3; ++a)
    for (int i = 0; i < ns[a]; ++i)
      if (!(ws[a][i + 1] - ws[a][i] > 0.0))
        return fail(HYP_ERR_INVALID, std::string("all ") + names[a] + " values should be greater than zero");
  c->grid_type = GEO_SPH;
  c->polar_kind = POLAR_CYL;
  c->n1 = n1;
  c->n2 = n2;
  c->n3 = n3;
  c->n_cells = (int64_t)n1 * n2 * n3;
  c->w1.assign(w1, w1 + n1 + 1);
  c->w2.assign(w2, w2 + n2 + 1);
  c->w3.assign(w3, w3 + n3 + 1);
  return HYP_OK;
}

int hyp_set_grid_octree(hyp_ctx *c, int32_t n_cells, const int32_t *refined, double x, double y, double z, double dx,
                        double dy, double dz) {
  if (!c || !refined) return fail(HYP_ERR_INVALID, "NULL argument");
  if (c->finalized) return fail(HYP_ERR_STATE, "model is frozen");
  if (n_cells < 1) return fail(HYP_ERR_INVALID, "octree needs at least one cell");
  if (!(dx > 0.0 && dy > 0.0 && dz > 0.0)) return fail(HYP_ERR_INVALID, "all volumes should be greater than zero");
  // octree_setup_indiv (grid_geometry_octree.f90:148-187), iteratively: nodes in depth-first order
  std::vector<OctNode> nodes(n_cells);
  std::vector<int32_t> parent(n_cells, -1), parent_sub(n_cells, 0), children;
  for (auto &n : nodes) {
    n.first_child = -1;
    n.pad = 0;
    for (int k = 0; k < 6; ++k) n.nb[k] = -1;
  }
  nodes[0].x = x; nodes[0].y = y; nodes[0].z = z;
  nodes[0].dx = dx; nodes[0].dy = dy; nodes[0].dz = dz;
  int n_filled = 1;
  std::vector<std::pair<int, int>> stack;
  auto refine = [&](int id) {
    nodes[id].first_child = (int32_t)(children.size() / 8);
    children.resize(children.size() + 8, -1);
    stack.push_back({id, 0});
  };
  if (refined[0] == 1) refine(0);
  while (!stack.empty()) {
    const int par = stack.back().first, k = stack.back().second;
    if (k == 8) {
      stack.pop_back();
      continue;
    }
    stack.back().second = k + 1;
    if (n_filled >= n_cells) return fail(HYP_ERR_INVALID, "refined array is not self-consistent");
    const int child = n_filled++;
    children[(size_t)nodes[par].first_child * 8 + k] = child;
    const int sx = (k & 1) ? 1 : -1, sy = (k & 2) ? 1 : -1, sz = (k & 4) ? 1 : -1;
    nodes[child].x = nodes[par].x + sx * nodes[par].dx / 2.0;
    nodes[child].y = nodes[par].y + sy * nodes[par].dy / 2.0;
    nodes[child].z = nodes[par].z + sz * nodes[par].dz / 2.0;
    nodes[child].dx = nodes[par].dx / 2.0;
    nodes[child].dy = nodes[par].dy / 2.0;
    nodes[child].dz = nodes[par].dz / 2.0;
    parent[child] = par;
    parent_sub[child] = k;
    if (refined[child] == 1) refine(child);
  }
  if (n_filled != n_cells) return fail(HYP_ERR_INVALID, "refined array is not self-consistent");
  // neighbour links: nodes are in depth-first order, so a parent's links exist before its children's.
  // Behind wall w of child k of parent p lies a sibling if the child sits on the near side of that
  // axis, else the matching child (same k with the axis bit flipped) of p's neighbour -- or that
  // neighbour itself when it is a leaf (a coarser cell).
  for (int id = 1; id < n_cells; ++id) {
    const int par = parent[id], k = parent_sub[id];
    for (int w = 0; w < 6; ++w) {
      const int axis = w / 2, bit = 1 << axis, up = w & 1;
      const bool high = (k & bit) != 0;
      if (high != (up != 0)) {
        nodes[id].nb[w] = children[(size_t)nodes[par].first_child * 8 + (k ^ bit)];
      } else {
        const int n = nodes[par].nb[w];
        if (n < 0 || nodes[n].first_child < 0)
          nodes[id].nb[w] = n;
        else
          nodes[id].nb[w] = children[(size_t)nodes[n].first_child * 8 + (k ^ bit)];
      }
    }
  }
  c->oct_leaves.clear();
  for (int id = 0; id < n_cells; ++id)
    if (refined[id] == 0) c->oct_leaves.push_back(id);
  c->oct_nodes.swap(nodes);
  c->oct_children.swap(children);
  const double m = std::max(dx, std::max(dy, dz));
  c->oct_eps = 3.0 * (std::nextafter(m, std::numeric_limits<double>::infinity()) - m);
  c->grid_type = GEO_OCT;
  c->n1 = n_cells;
  c->n2 = c->n3 = 1;
  c->n_cells = n_cells;
  c->w1.assign(2, 0.0);
  c->w2.assign(2, 0.0);
  c->w3.assign(2, 0.0);
  return HYP_OK;
}

int hyp_set_grid_amr(hyp_ctx *c, int32_t n_levels, const int32_t *n_grids, const int32_t *dims, const double *bounds) {
  if (!c || !n_grids || !dims || !bounds) return fail(HYP_ERR_INVALID, "NULL argument");
  if (c->finalized) return fail(HYP_ERR_STATE, "model is frozen");
  if (n_levels < 1 || n_grids[0] < 1) return fail(HYP_ERR_INVALID, "AMR grid needs at least one level-1 grid");
  // read_grid + setup_grid_geometry (grid_geometry_amr.f90:111-507), grids of all levels in one flat list
  std::vector<AmrGridDev> G;
  std::vector<int> level_of;
  std::vector<std::array<double, 3>> width;
  int64_t n_cells = 0, goto_total = 0;
  double min_width = std::numeric_limits<double>::max();
  for (int il = 0, k = 0; il < n_levels; ++il)
    for (int ig = 0; ig < n_grids[il]; ++ig, ++k) {
      AmrGridDev g;
      g.n1 = dims[3 * k]; g.n2 = dims[3 * k + 1]; g.n3 = dims[3 * k + 2];
      if (g.n1 < 1 || g.n2 < 1 || g.n3 < 1) return fail(HYP_ERR_INVALID, "grid needs at least one cell per axis");
      g.xmin = bounds[6 * k]; g.xmax = bounds[6 * k + 1];
      g.ymin = bounds[6 * k + 2]; g.ymax = bounds[6 * k + 3];
      g.zmin = bounds[6 * k + 4]; g.zmax = bounds[6 * k + 5];
      if (!(g.xmax > g.xmin && g.ymax > g.ymin && g.zmax > g.zmin))
        return fail(HYP_ERR_INVALID, "all volumes should be greater than zero");
      g.start_id = (int32_t)n_cells;
      g.goto_off = goto_total;
      n_cells += (int64_t)g.n1 * g.n2 * g.n3;
      goto_total += (int64_t)(g.n1 + 2) * (g.n2 + 2) * (g.n3 + 2);
      if (n_cells > 2000000000LL) return fail(HYP_ERR_INVALID, "grid too large for 32-bit cell ids");
      std::array<double, 3> w = {(g.xmax - g.xmin) / g.n1, (g.ymax - g.ymin) / g.n2, (g.zmax - g.zmin) / g.n3};
      for (double v : w) min_width = std::min(min_width, v);
      G.push_back(g);
      level_of.push_back(il);
      width.push_back(w);
    }
  // The consistency checks of setup_grid_geometry (grid_geometry_amr.f90:239-314), with its messages
  // (hyperion/model/tests/test_amr_checks.py greps the log for them).
  {
    const char *axis = "xyz";
    auto lo = [&](size_t k, int d) { return d == 0 ? G[k].xmin : (d == 1 ? G[k].ymin : G[k].zmin); };
    auto aligned = [](double x1, double x2, double dx) {  // :184-191
      double r = std::fmod(std::fabs(x1 - x2), dx);
      if (r > 0.5 * dx) r = dx - r;
      return std::fabs(r / dx) < 1.e-8;
    };
    std::vector<size_t> first_of;  // first grid of every level
    for (size_t k = 0; k < G.size(); ++k)
      if ((size_t)level_of[k] == first_of.size()) first_of.push_back(k);
    char msg[256];
    // grids of a level share their cell widths and sit on a common lattice
    for (size_t k = 0; k < G.size(); ++k) {
      const size_t ref = first_of[level_of[k]];
      if (k == ref) continue;
      const int igrid = (int)(k - ref) + 1, ilevel = level_of[k] + 1;
      for (int d = 0; d < 3; ++d)
        if (std::fabs(width[k][d] - width[ref][d]) > 1.e-10 * width[k][d]) {
          snprintf(msg, sizeof msg, "Grids 1 and %d in level %d have differing cell widths in the %c direction (%11.4E and %11.4E respectively)",
                   igrid, ilevel, axis[d], width[ref][d], width[k][d]);
          return fail(HYP_ERR_INVALID, msg);
        }
      for (int d = 0; d < 3; ++d)
        if (!aligned(lo(k, d), lo(ref, d), width[ref][d])) {
          snprintf(msg, sizeof msg, "Grids 1 and %d in level %d have edges that are not separated by an integer number of cells in the %c direction",
                   igrid, ilevel, axis[d]);
          return fail(HYP_ERR_INVALID, msg);
        }
    }
    // integer refinement factors between consecutive levels
    for (size_t il = 0; il + 1 < first_of.size(); ++il)
      for (int d = 0; d < 3; ++d) {
        const double ref = width[first_of[il]][d] / width[first_of[il + 1]][d];
        if (std::fabs(ref - std::nearbyint(ref)) > 1.e-10) {
          snprintf(msg, sizeof msg, "Refinement factor in the %c direction between level %d and level %d is not an integer (%.3f)",
                   axis[d], (int)il + 1, (int)il + 2, ref);
          return fail(HYP_ERR_INVALID, msg);
        }
      }
    // grids line up with the cells of the parent level
    for (size_t k = 0; k < G.size(); ++k) {
      if (level_of[k] == 0) continue;
      const size_t ref = first_of[level_of[k] - 1];
      for (int d = 0; d < 3; ++d)
        if (!aligned(lo(k, d), lo(ref, d), width[ref][d])) {
          snprintf(msg, sizeof msg, "Grid %d in level %d is not aligned with cells in level %d in the %c direction",
                   (int)(k - first_of[level_of[k]]) + 1, level_of[k] + 1, level_of[k], axis[d]);
          return fail(HYP_ERR_INVALID, msg);
        }
    }
  }
  std::vector<int32_t> go((size_t)goto_total, 0);
  auto gidx = [&](const AmrGridDev &g, int i1, int i2, int i3) {
    return (size_t)g.goto_off + i1 + (size_t)(g.n1 + 2) * (i2 + (size_t)(g.n2 + 2) * i3);
  };
  auto wall = [](double a, double b, int i, int n) { return (b - a) * (double)i / (double)n + a; };
  auto centre = [&](const AmrGridDev &g, int axis, int i) {  // 1-based cell index
    if (axis == 0) return 0.5 * (wall(g.xmin, g.xmax, i - 1, g.n1) + wall(g.xmin, g.xmax, i, g.n1));
    if (axis == 1) return 0.5 * (wall(g.ymin, g.ymax, i - 1, g.n2) + wall(g.ymin, g.ymax, i, g.n2));
    return 0.5 * (wall(g.zmin, g.zmax, i - 1, g.n3) + wall(g.zmin, g.zmax, i, g.n3));
  };
  auto inside = [](const AmrGridDev &g, double x, double y, double z) {
    return !(x < g.xmin || x > g.xmax || y < g.ymin || y > g.ymax || z < g.zmin || z > g.zmax);
  };
  // cells covered by a grid of the next finer level point to it (:355-381)
  for (size_t a = 0; a < G.size(); ++a)
    for (size_t b = 0; b < G.size(); ++b) {
      if (level_of[b] != level_of[a] + 1) continue;
      const AmrGridDev &g1 = G[a], &g2 = G[b];
      if (g1.xmax < g2.xmin || g1.xmin > g2.xmax || g1.ymax < g2.ymin || g1.ymin > g2.ymax || g1.zmax < g2.zmin ||
          g1.zmin > g2.zmax)
        continue;
      for (int i3 = 1; i3 <= g1.n3; ++i3)
        for (int i2 = 1; i2 <= g1.n2; ++i2)
          for (int i1 = 1; i1 <= g1.n1; ++i1)
            if (inside(g2, centre(g1, 0, i1), centre(g1, 1, i2), centre(g1, 2, i3))) go[gidx(g1, i1, i2, i3)] = (int32_t)b + 1;
    }
  // ghost layer: the grid of the same or a coarser level (finest first) one half cell outside each face (:383-487)
  for (size_t a = 0; a < G.size(); ++a) {
    const AmrGridDev &g1 = G[a];
    for (int lev = level_of[a]; lev >= 0; --lev)
      for (size_t b = 0; b < G.size(); ++b) {
        if (level_of[b] != lev || a == b) continue;
        const AmrGridDev &g2 = G[b];
        const std::array<double, 3> &w2 = width[b];
        if (g1.xmax < g2.xmin - w2[0] * 0.5 || g1.xmin > g2.xmax + w2[0] * 0.5 || g1.ymax < g2.ymin - w2[1] * 0.5 ||
            g1.ymin > g2.ymax + w2[1] * 0.5 || g1.zmax < g2.zmin - w2[2] * 0.5 || g1.zmin > g2.zmax + w2[2] * 0.5)
          continue;
        auto mark = [&](int i1, int i2, int i3, double x, double y, double z) {
          int32_t &e = go[gidx(g1, i1, i2, i3)];
          if (e == 0 && inside(g2, x, y, z)) e = (int32_t)b + 1;
        };
        const std::array<double, 3> &w1 = width[a];
        for (int i3 = 1; i3 <= g1.n3; ++i3)
          for (int i2 = 1; i2 <= g1.n2; ++i2) {
            mark(0, i2, i3, g1.xmin - w1[0] * 0.5, centre(g1, 1, i2), centre(g1, 2, i3));
            mark(g1.n1 + 1, i2, i3, g1.xmax + w1[0] * 0.5, centre(g1, 1, i2), centre(g1, 2, i3));
          }
        for (int i3 = 1; i3 <= g1.n3; ++i3)
          for (int i1 = 1; i1 <= g1.n1; ++i1) {
            mark(i1, 0, i3, centre(g1, 0, i1), g1.ymin - w1[1] * 0.5, centre(g1, 2, i3));
            mark(i1, g1.n2 + 1, i3, centre(g1, 0, i1), g1.ymax + w1[1] * 0.5, centre(g1, 2, i3));
          }
        for (int i2 = 1; i2 <= g1.n2; ++i2)
          for (int i1 = 1; i1 <= g1.n1; ++i1) {
            mark(i1, i2, 0, centre(g1, 0, i1), centre(g1, 1, i2), g1.zmin - w1[2] * 0.5);
            mark(i1, i2, g1.n3 + 1, centre(g1, 0, i1), centre(g1, 1, i2), g1.zmax + w1[2] * 0.5);
          }
      }
  }
  // cell -> grid map and the list of valid (uncovered) cells (:489-506)
  std::vector<int32_t> cell_grid((size_t)n_cells), valid;
  for (size_t a = 0; a < G.size(); ++a) {
    const AmrGridDev &g = G[a];
    for (int i3 = 1; i3 <= g.n3; ++i3)
      for (int i2 = 1; i2 <= g.n2; ++i2)
        for (int i1 = 1; i1 <= g.n1; ++i1) {
          const int32_t ic = g.start_id + ((i3 - 1) * g.n2 + (i2 - 1)) * g.n1 + (i1 - 1);
          cell_grid[ic] = (int32_t)a;
          if (go[gidx(g, i1, i2, i3)] == 0) valid.push_back(ic);
        }
  }
  c->amr_grids.swap(G);
  c->amr_gotos.swap(go);
  c->amr_cell_grid.swap(cell_grid);
  c->amr_valid.swap(valid);
  c->amr_n_level1 = n_grids[0];
  c->amr_eps = min_width / 2.0;
  c->grid_type = GEO_AMR;
  c->n1 = (int)n_cells;
  c->n2 = c->n3 = 1;
  c->n_cells = n_cells;
  c->w1.assign(2, 0.0);
  c->w2.assign(2, 0.0);
  c->w3.assign(2, 0.0);
  return HYP_OK;
}

int hyp_set_grid_voronoi(hyp_ctx *c, int32_t n_cells, const double *coords, const double *bb_min, const double *bb_max,
                         const double *volume, const int32_t *sparse_idx, const int32_t *sparse_neighs, const double *box) {
  if (!c || !coords || !bb_min || !bb_max || !volume || !sparse_idx || !sparse_neighs || !box)
    return fail(HYP_ERR_INVALID, "NULL argument");
  if (c->finalized) return fail(HYP_ERR_STATE, "model is frozen");
  if (n_cells < 1) return fail(HYP_ERR_INVALID, "Voronoi grid needs at least one cell");
  for (int a = 0; a < 3; ++a)
    if (!(box[2 * a + 1] > box[2 * a])) return fail(HYP_ERR_INVALID, "Voronoi grid: empty bounding box");
  if (sparse_idx[0] != 0) return fail(HYP_ERR_INVALID, "sparse_idx should start at 0");
  for (int i = 0; i < n_cells; ++i)
    if (sparse_idx[i + 1] < sparse_idx[i]) return fail(HYP_ERR_INVALID, "sparse_idx should not decrease");
  for (int q = 0; q < sparse_idx[n_cells]; ++q)
    if (sparse_neighs[q] < -6 || sparse_neighs[q] >= n_cells) return fail(HYP_ERR_INVALID, "sparse_neighs out of range");
  // setup_grid_geometry (grid_geometry_voronoi.f90:92-187)
  c->vor_sites.assign(coords, coords + 3 * (size_t)n_cells);
  c->vor_bb.resize(6 * (size_t)n_cells);
  c->vor_volume.resize(n_cells);
  c->vor_valid.clear();
  for (int i = 0; i < n_cells; ++i) {
    for (int a = 0; a < 3; ++a) {
      c->vor_bb[6 * (size_t)i + 2 * a] = bb_min[3 * (size_t)i + a];
      c->vor_bb[6 * (size_t)i + 2 * a + 1] = bb_max[3 * (size_t)i + a];
    }
    if (volume[i] > 0.0) c->vor_valid.push_back(i);          // geo%mask = geo%volume > 0
    c->vor_volume[i] = volume[i] < 0.0 ? 0.0 : volume[i];
  }
  c->vor_nidx.assign(sparse_idx, sparse_idx + n_cells + 1);
  c->vor_neigh.assign(sparse_neighs, sparse_neighs + sparse_idx[n_cells]);
  for (int a = 0; a < 6; ++a) c->vor_box[a] = box[a];
  // nearest-site search: sites bucketed on a uniform grid over the box, about two sites per bucket
  const int per_axis = std::max(1, (int)std::cbrt((double)n_cells / 2.0));
  c->vor_nb = per_axis;
  const int nb = per_axis * per_axis * per_axis;
  std::vector<int> bucket(n_cells);
  c->vor_b_start.assign(nb + 1, 0);
  for (int i = 0; i < n_cells; ++i) {
    int b[3];
    for (int a = 0; a < 3; ++a) {
      const double w = (box[2 * a + 1] - box[2 * a]) / per_axis;
      b[a] = std::min(std::max((int)((coords[3 * (size_t)i + a] - box[2 * a]) / w), 0), per_axis - 1);
    }
    bucket[i] = (b[2] * per_axis + b[1]) * per_axis + b[0];
    c->vor_b_start[bucket[i] + 1]++;
  }
  for (int k = 0; k < nb; ++k) c->vor_b_start[k + 1] += c->vor_b_start[k];
  c->vor_b_sites.assign(n_cells, 0);
  std::vector<int> fill(c->vor_b_start.begin(), c->vor_b_start.end() - 1);
  for (int i = 0; i < n_cells; ++i) c->vor_b_sites[fill[bucket[i]]++] = i;
  c->grid_type = GEO_VOR;
  c->n1 = n_cells;
  c->n2 = c->n3 = 1;
  c->n_cells = n_cells;
  c->w1.assign(2, 0.0);
  c->w2.assign(2, 0.0);
  c->w3.assign(2, 0.0);
  return HYP_OK;
}

int hyp_add_dust(hyp_ctx *c, const hyp_dust_tables *t) {
  if (!c || !t) return fail(HYP_ERR_INVALID, "NULL argument");
  if (c->finalized) return fail(HYP_ERR_STATE, "model is frozen");
  if ((int)c->dust.size() >= MAX_DUST) return fail(HYP_ERR_INVALID, "too many dust types (max 4)");
  try {
    HostDust d;
    build_dust(*t, d.L, d.buf);
    d.nu.assign(t->nu, t->nu + t->n_nu);
    d.chi.assign(t->chi, t->chi + t->n_nu);
    d.albedo.assign(t->albedo, t->albedo + t->n_nu);
    d.emiss_nu.assign(t->emiss_nu, t->emiss_nu + t->n_emiss_nu);
    d.emiss_jnu.assign(t->emiss_jnu, t->emiss_jnu + (size_t)t->n_emiss_nu * t->n_jnu);
    d.n_jnu = t->n_jnu;
    c->dust.push_back(std::move(d));
  } catch (std::exception &e) {
    return fail(HYP_ERR_INVALID, e.what());
  }
  return HYP_OK;
}

int hyp_add_source(hyp_ctx *c, const hyp_source *s) {
  if (!c || !s) return fail(HYP_ERR_INVALID, "NULL argument");
  if (c->finalized) return fail(HYP_ERR_STATE, "model is frozen");
  if ((int)c->sources.size() >= MAX_SOURCES) return fail(HYP_ERR_INVALID, "too many sources");
  if (s->type != HYP_SOURCE_POINT && s->type != HYP_SOURCE_SPHERE && s->type != HYP_SOURCE_EXTERN_SPH &&
      s->type != HYP_SOURCE_EXTERN_BOX && s->type != HYP_SOURCE_PLANE_PARALLEL && s->type != HYP_SOURCE_POINT_COLLECTION &&
      s->type != HYP_SOURCE_MAP)
    return fail(HYP_ERR_INVALID, "unknown type in source list");
  if (s->type == HYP_SOURCE_MAP) {
    // grid_load_pdf_map (grid_geometry_common_3d.f90:47-63)
    if (c->n_cells == 0) return fail(HYP_ERR_STATE, "set the grid before a map source");
    if (!s->map || s->n_map != c->n_cells) return fail(HYP_ERR_INVALID, "luminosity map should have one entry per cell");
    double norm = 0.0;
    for (int64_t i = 0; i < s->n_map; ++i) {
      if (!(s->map[i] >= 0.0)) return fail(HYP_ERR_INVALID, "luminosity map should be positive");
      norm = norm + s->map[i];
    }
    if (!(norm > 0.0)) return fail(HYP_ERR_INVALID, "[normalize_pdf_discrete] all PDF elements are zero");
  }
  if (s->spectrum_type == HYP_SPECTRUM_LTE && s->type != HYP_SOURCE_MAP) {
    static const char *who[] = {"", "Point source", "Spherical source", "Spot", "", "External spherical source",
                                "External box source", "Plane parallel", "Point source collection"};
    return fail(HYP_ERR_INVALID, std::string(who[s->type]) + " cannot have LTE spectrum");
  }
  if ((s->type == HYP_SOURCE_SPHERE || s->type == HYP_SOURCE_EXTERN_SPH || s->type == HYP_SOURCE_PLANE_PARALLEL) && !(s->radius > 0.0))
    return fail(HYP_ERR_INVALID, "source radius should be positive");
  if (s->type == HYP_SOURCE_PLANE_PARALLEL && s->peeloff) return fail(HYP_ERR_INVALID, "Cannot peeloff plane parallel source");
  if (s->type == HYP_SOURCE_EXTERN_BOX && !(s->box[1] > s->box[0] && s->box[3] > s->box[2] && s->box[5] > s->box[4]))
    return fail(HYP_ERR_INVALID, "external box source: bounds should be increasing");
  if (s->type == HYP_SOURCE_POINT_COLLECTION && (s->n_points < 1 || !s->points_xyz || !s->points_lum))
    return fail(HYP_ERR_INVALID, "point collection is empty");
  if (s->type == HYP_SOURCE_POINT_COLLECTION) {
    double norm = 0.0;
    for (int64_t i = 0; i < s->n_points; ++i) norm = norm + s->points_lum[i];
    if (!(norm > 0.0)) return fail(HYP_ERR_INVALID, "[normalize_pdf_discrete] all PDF elements are zero");
  }
  if (!(s->luminosity >= 0.0)) return fail(HYP_ERR_INVALID, "source luminosity should be positive");
  int spec = -1;
  if (s->spectrum_type == HYP_SPECTRUM_TABLE) {
    try {
      HostSpectrum sp;
      build_spectrum(s->spec_nu, s->spec_fnu, s->n_spec, sp.L, sp.buf);
      sp.nu.assign(s->spec_nu, s->spec_nu + s->n_spec);
      sp.fnu.assign(s->spec_fnu, s->spec_fnu + s->n_spec);
      spec = (int)c->spectra.size();
      c->spectra.push_back(std::move(sp));
    } catch (std::exception &e) {
      return fail(HYP_ERR_INVALID, e.what());
    }
  } else if (s->spectrum_type != HYP_SPECTRUM_BLACKBODY && s->spectrum_type != HYP_SPECTRUM_LTE) {
    return fail(HYP_ERR_INVALID, "unknown spectrum specifier");
  }
  if (s->n_spots > 0 && s->type != HYP_SOURCE_SPHERE) return fail(HYP_ERR_INVALID, "only spherical sources can have spots");
  if (s->n_spots > 0 && !s->spots) return fail(HYP_ERR_INVALID, "NULL argument");
  hyp_source copy = *s;
  copy.spots = nullptr;
  {
    // source_read (source_type.f90:150-188): spot_pdf = (spot luminosities..., star), luminosity += spots
    c->spot_off.push_back((int)c->spots.size());
    if (s->n_spots > 0) {
      std::vector<double> cdf((size_t)s->n_spots + 1);
      for (int i = 0; i < s->n_spots; ++i) {
        if (!(s->spots[i].luminosity >= 0.0)) return fail(HYP_ERR_INVALID, "source luminosity should be positive");
        cdf[i] = s->spots[i].luminosity;
        copy.luminosity = copy.luminosity + s->spots[i].luminosity;
      }
      cdf[s->n_spots] = s->luminosity;
      for (int i = 1; i <= s->n_spots; ++i) cdf[i] = cdf[i - 1] + cdf[i];
      const double norm = cdf[s->n_spots];
      if (!(norm > 0.0)) return fail(HYP_ERR_INVALID, "[find_cdf_discrete] all PDF elements are zero");
      const double deg = 3.14159265358979323846 / 180.0;
      for (int i = 0; i <= s->n_spots; ++i) {
        SpotDev q;
        memset(&q, 0, sizeof q);
        q.cdf = cdf[i] / norm;
        q.spectrum = -1;
        if (i < s->n_spots) {
          const hyp_spot &h = s->spots[i];
          q.a_cost = cos(h.longitude * deg); q.a_sint = sin(h.longitude * deg);   // angle3d_deg(lon, lat) (:176)
          q.a_cosp = cos(h.latitude * deg);  q.a_sinp = sin(h.latitude * deg);
          q.cost = cos(h.radius * deg);
          q.freq_type = h.spectrum_type;
          q.temperature = h.temperature;
          if (h.spectrum_type == HYP_SPECTRUM_TABLE) {
            try {
              HostSpectrum sp;
              build_spectrum(h.spec_nu, h.spec_fnu, h.n_spec, sp.L, sp.buf);
              sp.nu.assign(h.spec_nu, h.spec_nu + h.n_spec);
              sp.fnu.assign(h.spec_fnu, h.spec_fnu + h.n_spec);
              q.spectrum = (int)c->spectra.size();
              c->spectra.push_back(std::move(sp));
            } catch (std::exception &e) {
              return fail(HYP_ERR_INVALID, e.what());
            }
          } else if (h.spectrum_type != HYP_SPECTRUM_BLACKBODY) {
            return fail(HYP_ERR_INVALID, "Spot cannot have LTE spectrum");
          }
        }
        c->spots.push_back(q);
      }
    }
  }
  copy.spec_nu = copy.spec_fnu = nullptr;
  copy.points_xyz = copy.points_lum = nullptr;
  copy.map = nullptr;
  if (s->type == HYP_SOURCE_MAP) {
    // set_pdf_discrete (type_pdf.f90:222-231) on the map
    std::vector<double> cdf(s->map, s->map + s->n_map);
    double norm = 0.0;
    for (double v : cdf) norm = norm + v;
    for (double &v : cdf) v = v / norm;
    for (int64_t i = 1; i < s->n_map; ++i) cdf[i] = cdf[i - 1] + cdf[i];
    const double last = cdf[s->n_map - 1];
    for (double &v : cdf) v = v / last;
    c->map_off.push_back((int64_t)c->map_cdf.size());
    c->map_cdf.insert(c->map_cdf.end(), cdf.begin(), cdf.end());
  } else {
    c->map_off.push_back(-1);
  }
  if (s->type == HYP_SOURCE_POINT_COLLECTION) {
    // set_pdf_discrete (type_pdf.f90:222-231): normalise, accumulate, normalise the cdf
    std::vector<double> cdf(s->points_lum, s->points_lum + s->n_points);
    double norm = 0.0;
    for (double v : cdf) norm = norm + v;
    for (double &v : cdf) v = v / norm;
    for (int64_t i = 1; i < s->n_points; ++i) cdf[i] = cdf[i - 1] + cdf[i];
    const double last = cdf[s->n_points - 1];
    for (double &v : cdf) v = v / last;
    c->coll_off.push_back((int64_t)c->coll_cdf.size());
    c->coll_cdf.insert(c->coll_cdf.end(), cdf.begin(), cdf.end());
    c->coll_xyz.insert(c->coll_xyz.end(), s->points_xyz, s->points_xyz + 3 * s->n_points);
    copy.luminosity = norm;  // s%luminosity = sum(luminosity_collection) (source_type.f90:268)
  } else {
    c->coll_off.push_back(-1);
  }
  c->sources.push_back(copy);
  c->source_spectrum.push_back(spec);
  return HYP_OK;
}

int hyp_set_run_conf(hyp_ctx *c, const hyp_run_conf *conf) {
  if (!c || !conf) return fail(HYP_ERR_INVALID, "NULL argument");
  if (c->finalized && conf->use_mrw && !c->M.use_mrw)
    return fail(HYP_ERR_STATE, "the modified random walk has to be enabled before hyp_finalize_setup");
  c->conf = *conf;
  if (c->finalized) {
    c->M.seed = (uint64_t)conf->seed;
    c->M.n_inter_max = conf->n_inter_max;
    c->M.n_reabs_max = conf->n_reabs_max;
    c->M.use_mrw = conf->use_mrw;
    c->M.mrw_gamma = conf->mrw_gamma;
    c->M.n_mrw_max = conf->n_mrw_max;
    c->M.kill_on_absorb = conf->kill_on_absorb;
    c->M.kill_on_scatter = conf->kill_on_scatter;
    c->M.sample_evenly = conf->sample_sources_evenly;
    c->M.enforce_energy_range = conf->enforce_energy_range;
  }
  return HYP_OK;
}

static int upload_density(hyp_ctx *c, const double *density) {
  const size_t n = (size_t)c->n_cells * c->dust.size();
  // cudaMemcpyDefault: after hyp_finalize_setup the caller may pass a host or a device pointer
  CUDA_TRY(cudaMemcpyAsync(c->d_stage, density, n * sizeof(double), cudaMemcpyDefault, c->stream));
  scatter_density_kernel<<<grid_blocks(c), 256, 0, c->stream>>>(c->M, c->d_stage);
  CUDA_TRY(cudaGetLastError());
  return HYP_OK;
}

int hyp_set_density(hyp_ctx *c, int32_t n_dust, const double *density) {
  if (!c || !density) return fail(HYP_ERR_INVALID, "NULL argument");
  if (n_dust != (int)c->dust.size()) return fail(HYP_ERR_INVALID, "density array has wrong number of dust types");
  if (c->n_cells == 0) return fail(HYP_ERR_STATE, "set the grid before the density");
  if (c->finalized) {
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = upload_density(c, density);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return HYP_OK;
  }
  c->h_density.assign(density, density + (size_t)n_dust * c->n_cells);
  c->have_density = true;
  return HYP_OK;
}

static int upload_energy(hyp_ctx *c) {
  const size_t n = (size_t)c->n_cells * c->dust.size();
  CUDA_TRY(cudaMemcpyAsync(c->d_stage, c->h_energy.data(), n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  to_device_order_kernel<<<grid_blocks(c), 256, 0, c->stream>>>((int)c->dust.size(), c->n_cells, c->d_stage, c->d_energy);
  if (c->conf.specific_energy_additional && c->energy_from_caller) {
    // grid_physics_3d.f90:213-235: keep the given array as the extra heating term and start from the minimum
    if (!c->d_energy_add) CUDA_TRY(cudaMalloc(&c->d_energy_add, n * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(c->d_energy_add, c->d_energy, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    c->M.energy_additional = c->d_energy_add;
    std::vector<double> start(n);
    const size_t nd = c->dust.size();
    for (size_t id = 0; id < nd; ++id)
      std::fill(start.begin() + id * c->n_cells, start.begin() + (id + 1) * c->n_cells, c->h_min_energy[id]);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpy(c->d_stage, start.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    to_device_order_kernel<<<grid_blocks(c), 256, 0, c->stream>>>((int)nd, c->n_cells, c->d_stage, c->d_energy);
  }
  clamp_energy_kernel<<<grid_blocks(c), 256, 0, c->stream>>>(c->M);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return HYP_OK;
}

int hyp_set_specific_energy(hyp_ctx *c, const double *se, const double *min_e) {
  if (!c) return fail(HYP_ERR_INVALID, "NULL argument");
  if (c->n_cells == 0 || c->dust.empty()) return fail(HYP_ERR_STATE, "set the grid and dust first");
  const size_t nd = c->dust.size(), n = (size_t)c->n_cells * nd;
  c->h_min_energy.assign(nd, 0.0);
  if (min_e)
    for (size_t i = 0; i < nd; ++i) c->h_min_energy[i] = min_e[i];
  c->h_energy.resize(n);
  if (se) {
    std::copy(se, se + n, c->h_energy.begin());
  } else {
    for (size_t id = 0; id < nd; ++id)
      std::fill(c->h_energy.begin() + id * c->n_cells, c->h_energy.begin() + (id + 1) * c->n_cells, c->h_min_energy[id]);
  }
  c->have_energy = true;
  c->energy_from_caller = se != nullptr;
  if (c->finalized) {
    CUDA_TRY(cudaSetDevice(c->device));
    for (size_t i = 0; i < nd; ++i) c->M.min_energy[i] = c->h_min_energy[i];
    return upload_energy(c);
  }
  return HYP_OK;
}

int hyp_finalize_setup(hyp_ctx *c) {
  if (!c) return fail(HYP_ERR_INVALID, "NULL argument");
  if (c->finalized) return fail(HYP_ERR_STATE, "already finalized");
  if (c->n_cells == 0) return fail(HYP_ERR_STATE, "no grid");
  if (c->dust.empty()) return fail(HYP_ERR_STATE, "no dust");
  if (!c->have_density) return fail(HYP_ERR_STATE, "no density");
  if (c->sources.empty()) return fail(HYP_ERR_INVALID, "no sources set up - need sources for initial iteration(s)");
  if (!c->have_energy) {
    int rc = hyp_set_specific_energy(c, nullptr, nullptr);
    if (rc) return rc;
  }
  if (c->conf.specific_energy_additional && !c->energy_from_caller)
    return fail(HYP_ERR_INVALID, "cannot specify specific_energy_type since specific_energy was not given");
  if (c->grid_type == GEO_VOR && c->conf.use_mrw)   // distance_to_closest_wall, grid_geometry_voronoi.f90:314-320
    return fail(HYP_ERR_INVALID, "not implemented for Voronoi grid");
  CUDA_TRY(cudaSetDevice(c->device));
  const int nd = (int)c->dust.size();
  const size_t n = (size_t)c->n_cells * nd;
  ModelDev &M = c->M;
  M.n1 = c->n1;
  M.n2 = c->n2;
  M.n3 = c->n3;
  M.n_dust = nd;
  M.n_cells = c->n_cells;
  M.n_sources = (int)c->sources.size();
  // walls
  const size_t nw = c->w1.size() + c->w2.size() + c->w3.size();
  CUDA_TRY(cudaMalloc(&c->d_w, nw * sizeof(double)));
  CUDA_TRY(cudaMemcpy(c->d_w, c->w1.data(), c->w1.size() * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(c->d_w + c->w1.size(), c->w2.data(), c->w2.size() * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(c->d_w + c->w1.size() + c->w2.size(), c->w3.data(), c->w3.size() * sizeof(double),
                      cudaMemcpyHostToDevice));
  M.w1 = c->d_w;
  M.w2 = c->d_w + c->w1.size();
  M.w3 = M.w2 + c->w2.size();
  M.grid_type = c->grid_type;
  if (c->grid_type == GEO_AMR) {
    AmrGrid &A = M.amr;
    auto up = [&](auto *&dst, const auto &src) -> int {
      CUDA_TRY(cudaMalloc(&dst, std::max<size_t>(src.size(), 1) * sizeof(src[0])));
      CUDA_TRY(cudaMemcpy(dst, src.data(), src.size() * sizeof(src[0]), cudaMemcpyHostToDevice));
      return HYP_OK;
    };
    int rc = up(c->d_amr_grids, c->amr_grids);
    if (!rc) rc = up(c->d_amr_gotos, c->amr_gotos);
    if (!rc) rc = up(c->d_amr_cell_grid, c->amr_cell_grid);
    if (!rc) rc = up(c->d_amr_valid, c->amr_valid);
    if (rc) return rc;
    A.grids = c->d_amr_grids;
    A.gotos = c->d_amr_gotos;
    A.cell_grid = c->d_amr_cell_grid;
    A.valid = c->d_amr_valid;
    A.n_grids = (int32_t)c->amr_grids.size();
    A.n_level1 = c->amr_n_level1;
    A.n_cells = (int32_t)c->n_cells;
    A.n_valid = (int32_t)c->amr_valid.size();
    A.eps = c->amr_eps;
    // cells covered by a finer grid hold no dust (grid_physics_3d.f90:156-164)
    std::vector<char> ok((size_t)c->n_cells, 0);
    for (int32_t ic : c->amr_valid) ok[ic] = 1;
    const size_t nc = (size_t)c->n_cells;
    for (size_t id = 0; id < c->dust.size(); ++id)
      for (size_t ic = 0; ic < nc; ++ic)
        if (!ok[ic]) {
          c->h_density[id * nc + ic] = 0.0;
          if (c->energy_from_caller) c->h_energy[id * nc + ic] = 0.0;
        }
  }
  if (c->grid_type == GEO_VOR) {
    VorGrid &V = M.vor;
    const size_t n = (size_t)c->n_cells;
    std::vector<double> f;
    f.insert(f.end(), c->vor_sites.begin(), c->vor_sites.end());
    f.insert(f.end(), c->vor_bb.begin(), c->vor_bb.end());
    f.insert(f.end(), c->vor_volume.begin(), c->vor_volume.end());
    CUDA_TRY(cudaMalloc(&c->d_vor_f64, f.size() * sizeof(double)));
    CUDA_TRY(cudaMemcpy(c->d_vor_f64, f.data(), f.size() * sizeof(double), cudaMemcpyHostToDevice));
    V.sites = c->d_vor_f64;
    V.bb = c->d_vor_f64 + 3 * n;
    V.volume = c->d_vor_f64 + 9 * n;
    std::vector<int32_t> I;
    size_t off[5];
    const std::vector<int32_t> *parts[5] = {&c->vor_nidx, &c->vor_neigh, &c->vor_valid, &c->vor_b_start, &c->vor_b_sites};
    for (int k = 0; k < 5; ++k) {
      off[k] = I.size();
      I.insert(I.end(), parts[k]->begin(), parts[k]->end());
    }
    CUDA_TRY(cudaMalloc(&c->d_vor_i32, std::max<size_t>(I.size(), 1) * sizeof(int32_t)));
    CUDA_TRY(cudaMemcpy(c->d_vor_i32, I.data(), I.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    V.nidx = c->d_vor_i32 + off[0];
    V.neigh = c->d_vor_i32 + off[1];
    V.valid = c->d_vor_i32 + off[2];
    V.b_start = c->d_vor_i32 + off[3];
    V.b_sites = c->d_vor_i32 + off[4];
    for (int a = 0; a < 6; ++a) V.box[a] = c->vor_box[a];
    for (int a = 0; a < 3; ++a) {
      V.nb[a] = c->vor_nb;
      V.bw[a] = (c->vor_box[2 * a + 1] - c->vor_box[2 * a]) / c->vor_nb;
    }
    V.n_cells = (int32_t)n;
    V.n_valid = (int32_t)c->vor_valid.size();
    // cells without volume hold no dust (geo%mask, grid_physics_3d.f90:156-164)
    std::vector<char> ok(n, 0);
    for (int32_t ic : c->vor_valid) ok[ic] = 1;
    for (size_t id = 0; id < c->dust.size(); ++id)
      for (size_t ic = 0; ic < n; ++ic)
        if (!ok[ic]) {
          c->h_density[id * n + ic] = 0.0;
          if (c->energy_from_caller) c->h_energy[id * n + ic] = 0.0;
        }
  }
  if (c->grid_type == GEO_OCT) {
    OctGrid &G = M.oct;
    CUDA_TRY(cudaMalloc(&c->d_oct_nodes, c->oct_nodes.size() * sizeof(OctNode)));
    CUDA_TRY(cudaMemcpy(c->d_oct_nodes, c->oct_nodes.data(), c->oct_nodes.size() * sizeof(OctNode), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&c->d_oct_children, std::max<size_t>(c->oct_children.size(), 8) * sizeof(int32_t)));
    CUDA_TRY(cudaMemcpy(c->d_oct_children, c->oct_children.data(), c->oct_children.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&c->d_oct_leaves, c->oct_leaves.size() * sizeof(int32_t)));
    CUDA_TRY(cudaMemcpy(c->d_oct_leaves, c->oct_leaves.data(), c->oct_leaves.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    G.nodes = c->d_oct_nodes;
    G.children = c->d_oct_children;
    G.leaves = c->d_oct_leaves;
    G.n_nodes = (int32_t)c->oct_nodes.size();
    G.n_leaves = (int32_t)c->oct_leaves.size();
    G.eps = c->oct_eps;
    // refined nodes hold no dust (setup_grid_physics applies the mask, grid_physics_3d.f90:156-164)
    const size_t nc = (size_t)c->n_cells;
    for (size_t id = 0; id < c->dust.size(); ++id)
      for (size_t ic = 0; ic < nc; ++ic)
        if (c->oct_nodes[ic].first_child >= 0) {
          c->h_density[id * nc + ic] = 0.0;
          if (c->energy_from_caller) c->h_energy[id * nc + ic] = 0.0;
        }
  }
  if (c->grid_type == GEO_SPH) {
    // derived wall quantities of setup_grid_geometry (grid_geometry_spherical_3d.f90:137-201)
    const int n1 = c->n1, n2 = c->n2, n3 = c->n3;
    SphGrid &G = M.sph;
    G.kind = c->polar_kind;
    const bool cyl = c->polar_kind == POLAR_CYL;
    G.n1 = n1; G.n2 = n2; G.n3 = n3;
    int off = 0;
    auto take = [&](int n) { int o = off; off += n; return o; };
    G.o_w1 = take(n1 + 1); G.o_wr2 = take(n1 + 1); G.o_ew1 = take(n1 + 1);
    G.o_w2 = take(n2 + 1); G.o_wtant = take(n2 + 1); G.o_wtant2 = take(n2 + 1); G.o_wcost = take(n2 + 1);
    G.o_w3 = take(n3 + 1); G.o_wtanp = take(n3 + 1); G.o_wcosp = take(n3 + 1); G.o_wsinp = take(n3 + 1);
    G.o_dr3 = take(n1); G.o_dcost = take(n2); G.o_dphi = take(n3);
    G.o_ew2 = take(n2 + 1);
    std::vector<double> T(off);
    for (int i = 0; i <= n1; ++i) {
      const double w = c->w1[i];
      T[G.o_w1 + i] = w;
      T[G.o_wr2 + i] = w * w;
      const double aw = std::fabs(w);
      T[G.o_ew1 + i] = 3 * (w == 0.0 ? std::numeric_limits<double>::min()
                                     : std::nextafter(aw, std::numeric_limits<double>::infinity()) - aw);
    }
    G.midplane = -1;
    bool any = false;
    for (int i = 0; i <= n2; ++i) {
      const double w = c->w2[i];
      T[G.o_w2 + i] = w;
      T[G.o_wtant + i] = std::tan(w);
      T[G.o_wtant2 + i] = std::tan(w) * std::tan(w);
      T[G.o_wcost + i] = std::cos(w);
      if (std::fabs(w - SPH_PI / 2.0) < (double)1.e-6f) any = true;
      const double aw = std::fabs(w);
      T[G.o_ew2 + i] = 3 * (w == 0.0 ? std::numeric_limits<double>::min()
                                     : std::nextafter(aw, std::numeric_limits<double>::infinity()) - aw);
    }
    if (any && !cyl) {
      int best = 0;
      for (int i = 1; i <= n2; ++i)
        if (std::fabs(c->w2[i] - SPH_PI / 2.0) < std::fabs(c->w2[best] - SPH_PI / 2.0)) best = i;
      G.midplane = best;
    }
    for (int i = 0; i <= n3; ++i) {
      const double w = c->w3[i];
      T[G.o_w3 + i] = w;
      T[G.o_wtanp + i] = std::tan(w);
      T[G.o_wcosp + i] = std::cos(w);
      T[G.o_wsinp + i] = std::sin(w);
    }
    for (int i = 0; i < n1; ++i)
      T[G.o_dr3 + i] = cyl ? c->w1[i + 1] * c->w1[i + 1] - c->w1[i] * c->w1[i]
                           : c->w1[i + 1] * c->w1[i + 1] * c->w1[i + 1] - c->w1[i] * c->w1[i] * c->w1[i];
    for (int i = 0; i < n2; ++i) T[G.o_dcost + i] = cyl ? c->w2[i + 1] - c->w2[i] : std::cos(c->w2[i]) - std::cos(c->w2[i + 1]);
    for (int i = 0; i < n3; ++i) T[G.o_dphi + i] = c->w3[i + 1] - c->w3[i];
    for (int i2 = 0; i2 < n2; ++i2)
      if (T[G.o_dcost + i2] == 0.0) return fail(HYP_ERR_INVALID, "all volumes should be greater than zero");
    CUDA_TRY(cudaMalloc(&c->d_sph, T.size() * sizeof(double)));
    CUDA_TRY(cudaMemcpy(c->d_sph, T.data(), T.size() * sizeof(double), cudaMemcpyHostToDevice));
    G.T = c->d_sph;
  }
  // dust tables
  for (int id = 0; id < nd; ++id) {
    HostDust &d = c->dust[id];
    CUDA_TRY(cudaMalloc(&d.dev, d.buf.size() * sizeof(double)));
    CUDA_TRY(cudaMemcpy(d.dev, d.buf.data(), d.buf.size() * sizeof(double), cudaMemcpyHostToDevice));
    M.dust[id].L = d.L;
    M.dust[id].B = d.dev;
    M.dust[id].Bm = nullptr;
    if (c->conf.use_mrw) {
      std::vector<double> mb;
      try {
        build_dust_mrw(d.nu.data(), d.chi.data(), d.albedo.data(), (int)d.nu.size(), d.emiss_nu.data(), d.emiss_jnu.data(),
                       (int)d.emiss_nu.size(), d.n_jnu, d.Lm, mb);
      } catch (std::exception &e) {
        return fail(HYP_ERR_INVALID, e.what());
      }
      CUDA_TRY(cudaMalloc(&d.dev_mrw, mb.size() * sizeof(double)));
      CUDA_TRY(cudaMemcpy(d.dev_mrw, mb.data(), mb.size() * sizeof(double), cudaMemcpyHostToDevice));
      M.dust[id].Lm = d.Lm;
      M.dust[id].Bm = d.dev_mrw;
    }
    M.min_energy[id] = c->h_min_energy[id];
  }
  // spectra + sources
  std::vector<SpectrumDev> hs(c->spectra.size());
  for (size_t i = 0; i < c->spectra.size(); ++i) {
    HostSpectrum &s = c->spectra[i];
    CUDA_TRY(cudaMalloc(&s.dev, s.buf.size() * sizeof(double)));
    CUDA_TRY(cudaMemcpy(s.dev, s.buf.data(), s.buf.size() * sizeof(double), cudaMemcpyHostToDevice));
    hs[i].L = s.L;
    hs[i].B = s.dev;
  }
  if (!hs.empty()) {
    CUDA_TRY(cudaMalloc(&c->d_spectra, hs.size() * sizeof(SpectrumDev)));
    CUDA_TRY(cudaMemcpy(c->d_spectra, hs.data(), hs.size() * sizeof(SpectrumDev), cudaMemcpyHostToDevice));
  }
  M.spectra = c->d_spectra;
  // luminosity PDF: set_pdf_discrete (type_pdf.f90:222-231)
  double ltot = 0.0;
  for (auto &s : c->sources) ltot += s.luminosity;
  if (!(ltot > 0.0)) return fail(HYP_ERR_INVALID, "[normalize_pdf_discrete] all PDF elements are zero");
  c->energy_total = ltot;
  std::vector<SourceDev> sd(c->sources.size());
  double cum = 0.0;
  for (size_t i = 0; i < sd.size(); ++i) {
    const hyp_source &s = c->sources[i];
    sd[i].type = s.type;
    sd[i].freq_type = s.spectrum_type;
    sd[i].x = s.x;
    sd[i].y = s.y;
    sd[i].z = s.z;
    sd[i].radius = s.radius;
    sd[i].temperature = s.temperature;
    sd[i].limb = s.limb_darkening;
    sd[i].peeloff = s.peeloff;
    sd[i].pad = 0;
    sd[i].spectrum = c->source_spectrum[i];
    for (int k = 0; k < 6; ++k) sd[i].box[k] = s.box[k];
    if (s.type == HYP_SOURCE_EXTERN_BOX) {
      // set_pdf(s%face, (/dy*dz, dy*dz, dz*dx, dz*dx, dx*dy, dx*dy/)) (source_type.f90:229)
      const double dx = s.box[1] - s.box[0], dy = s.box[3] - s.box[2], dz = s.box[5] - s.box[4];
      double a[6] = {dy * dz, dy * dz, dz * dx, dz * dx, dx * dy, dx * dy};
      double norm = 0.0;
      for (double v : a) norm = norm + v;
      for (double &v : a) v = v / norm;
      for (int k = 1; k < 6; ++k) a[k] = a[k - 1] + a[k];
      for (int k = 0; k < 6; ++k) sd[i].face_cdf[k] = a[k] / a[5];
    } else {
      for (int k = 0; k < 6; ++k) sd[i].face_cdf[k] = 1.0;
    }
    {
      // angle3d_deg (type_angle3d.f90:127-134)
      const double deg2rad = 3.14159265358979323846 / 180.0;
      sd[i].dir_cost = cos(s.theta * deg2rad);
      sd[i].dir_sint = sin(s.theta * deg2rad);
      sd[i].dir_cosp = cos(s.phi * deg2rad);
      sd[i].dir_sinp = sin(s.phi * deg2rad);
    }
    sd[i].map_off = c->map_off[i];
    sd[i].spot_off = c->spot_off[i];
    sd[i].n_spots = s.type == HYP_SOURCE_SPHERE ? s.n_spots : 0;
    sd[i].coll_off = c->coll_off[i];
    sd[i].coll_n = s.type == HYP_SOURCE_POINT_COLLECTION ? s.n_points : 0;
    sd[i].pdf = s.luminosity / ltot;
    cum += sd[i].pdf;
    sd[i].cdf = cum;
  }
  for (auto &s : sd) s.cdf /= cum;
  if (!c->coll_cdf.empty()) {
    CUDA_TRY(cudaMalloc(&c->d_coll_xyz, c->coll_xyz.size() * sizeof(double)));
    CUDA_TRY(cudaMalloc(&c->d_coll_cdf, c->coll_cdf.size() * sizeof(double)));
    CUDA_TRY(cudaMemcpy(c->d_coll_xyz, c->coll_xyz.data(), c->coll_xyz.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->d_coll_cdf, c->coll_cdf.data(), c->coll_cdf.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  M.coll_xyz = c->d_coll_xyz;
  M.coll_cdf = c->d_coll_cdf;
  if (!c->map_cdf.empty()) {
    CUDA_TRY(cudaMalloc(&c->d_map_cdf, c->map_cdf.size() * sizeof(double)));
    CUDA_TRY(cudaMemcpy(c->d_map_cdf, c->map_cdf.data(), c->map_cdf.size() * sizeof(double), cudaMemcpyHostToDevice));
    std::vector<double>().swap(c->map_cdf);
  }
  M.map_cdf = c->d_map_cdf;
  if (!c->spots.empty()) {
    CUDA_TRY(cudaMalloc(&c->d_spots, c->spots.size() * sizeof(SpotDev)));
    CUDA_TRY(cudaMemcpy(c->d_spots, c->spots.data(), c->spots.size() * sizeof(SpotDev), cudaMemcpyHostToDevice));
  }
  M.spots = c->d_spots;
  CUDA_TRY(cudaMalloc(&c->d_sources, sd.size() * sizeof(SourceDev)));
  CUDA_TRY(cudaMemcpy(c->d_sources, sd.data(), sd.size() * sizeof(SourceDev), cudaMemcpyHostToDevice));
  M.sources = c->d_sources;
  // grids
  CUDA_TRY(cudaMalloc(&c->d_cells, n * sizeof(CellRec)));
  CUDA_TRY(cudaMalloc(&c->d_rho, n * sizeof(double)));
  CUDA_TRY(cudaMalloc(&c->d_energy, n * sizeof(double)));
  CUDA_TRY(cudaMalloc(&c->d_jfrac, n * sizeof(double)));
  CUDA_TRY(cudaMalloc(&c->d_jid, n * sizeof(int32_t)));
  // [sums | scalars | n_photons as doubles]: one buffer, one collective
  const size_t n_spec = c->spec_edges.empty() ? 0 : c->spec_edges.size() - 1;
  CUDA_TRY(cudaMalloc(&c->d_sums, (n + SC_COUNT + (size_t)c->n_cells + n_spec * n) * sizeof(double)));
  CUDA_TRY(cudaMalloc(&c->d_stage, n * sizeof(double)));
  CUDA_TRY(cudaMalloc(&c->d_work, sizeof(unsigned long long)));
  CUDA_TRY(cudaMalloc(&c->d_error, sizeof(int32_t)));
  CUDA_TRY(cudaMemset(c->d_cells, 0, n * sizeof(CellRec)));
  CUDA_TRY(cudaMemset(c->d_sums, 0, (n + SC_COUNT + (size_t)c->n_cells + n_spec * n) * sizeof(double)));
  M.spec_sums = M.spec_energy = nullptr;
  M.spec_log_edges = M.spec_jfrac = nullptr;
  M.n_spec_bins = M.spec_jmax = 0;
  if (n_spec) {
    // setup_grid_physics (grid_physics_3d.f90:124-143,199-207,229-233,249-252,269-284)
    M.n_spec_bins = (int32_t)n_spec;
    M.spec_sums = c->d_sums + n + SC_COUNT + (size_t)c->n_cells;
    CUDA_TRY(cudaMalloc(&c->d_spec_energy, n_spec * n * sizeof(double)));
    std::vector<double> e0(n_spec * n, 0.0);
    if (!c->energy_from_caller || c->conf.specific_energy_additional)
      for (size_t b = 0; b < n_spec; ++b)
        for (size_t k = 0; k < n; ++k) e0[b * n + k] = c->h_min_energy[k % nd];
    CUDA_TRY(cudaMemcpy(c->d_spec_energy, e0.data(), e0.size() * sizeof(double), cudaMemcpyHostToDevice));
    M.spec_energy = c->d_spec_energy;
    int jmax = 0;
    for (const HostDust &d : c->dust) jmax = std::max(jmax, d.n_jnu);
    M.spec_jmax = jmax;
    std::vector<double> tab(n_spec + 1 + (size_t)nd * jmax * n_spec, 0.0);
    for (size_t i = 0; i <= n_spec; ++i) tab[i] = std::log10(c->spec_edges[i]);
    // get_j_nu_bin_fractions (dust_type_4elem.f90:752-778)
    for (int id = 0; id < nd; ++id) {
      const HostDust &D = c->dust[id];
      const int ne = (int)D.emiss_nu.size();
      std::vector<double> col(ne);
      for (int iv = 0; iv < D.n_jnu; ++iv) {
        for (int k = 0; k < ne; ++k) col[k] = D.emiss_jnu[(size_t)k * D.n_jnu + iv];
        const double norm = detail::integral_loglog_all(D.emiss_nu.data(), col.data(), ne);
        double *frac = &tab[n_spec + 1 + ((size_t)id * jmax + iv) * n_spec];
        for (size_t b = 0; b < n_spec; ++b) {
          frac[b] = detail::integral_loglog_range(D.emiss_nu.data(), col.data(), ne, c->spec_edges[b], c->spec_edges[b + 1]);
          if (norm > 0.0) frac[b] /= norm;
        }
      }
    }
    CUDA_TRY(cudaMalloc(&c->d_spec_tab, tab.size() * sizeof(double)));
    CUDA_TRY(cudaMemcpy(c->d_spec_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    M.spec_log_edges = c->d_spec_tab;
    M.spec_jfrac = c->d_spec_tab + n_spec + 1;
  }
  CUDA_TRY(cudaMemset(c->d_error, 0, sizeof(int32_t)));
  M.cells = c->d_cells;
  M.rho = c->d_rho;
  M.specific_energy = c->d_energy;
  M.jnu_id = c->d_jid;
  M.jnu_frac = c->d_jfrac;
  M.scalars = c->d_sums + n;
  M.work_counter = c->d_work;
  M.error_flag = c->d_error;
  M.seed = (uint64_t)c->conf.seed;
  M.n_inter_max = c->conf.n_inter_max;
  M.n_reabs_max = c->conf.n_reabs_max;
  M.use_mrw = c->conf.use_mrw;
  M.mrw_gamma = c->conf.mrw_gamma;
  M.n_mrw_max = c->conf.n_mrw_max;
  if (c->conf.use_mrw) {
    double cdf[200];
    build_mrw_cumulative(cdf, cdf + 100);
    CUDA_TRY(cudaMalloc(&c->d_mrw_cdf, sizeof cdf));
    CUDA_TRY(cudaMemcpy(c->d_mrw_cdf, cdf, sizeof cdf, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&c->d_mrw_alpha, (size_t)c->n_cells * sizeof(double)));
    CUDA_TRY(cudaMalloc(&c->d_mrw_diff, (size_t)c->n_cells * sizeof(double)));
    M.mrw_cdf = c->d_mrw_cdf;
    M.alpha_inv_planck = c->d_mrw_alpha;
    M.diff_coeff = c->d_mrw_diff;
  }
  M.any_sphere = 0;
  for (auto &src : c->sources)
    if (src.type == HYP_SOURCE_SPHERE) M.any_sphere = 1;
  M.kill_on_absorb = c->conf.kill_on_absorb;
  M.kill_on_scatter = c->conf.kill_on_scatter;
  M.sample_evenly = c->conf.sample_sources_evenly;
  M.enforce_energy_range = c->conf.enforce_energy_range;
  int rc = upload_density(c, c->h_density.data());
  if (rc) return rc;
  rc = upload_energy(c);
  if (rc) return rc;
  c->h_density.clear();
  c->h_density.shrink_to_fit();
  c->h_energy.clear();
  c->h_energy.shrink_to_fit();
  c->finalized = true;
  return HYP_OK;
}

int hyp_lucy_begin(hyp_ctx *c) {
  if (!c || !c->finalized) return fail(HYP_ERR_STATE, "hyp_finalize_setup has not been called");
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t n = (size_t)c->n_cells * c->dust.size();
  CUDA_TRY(cudaEventRecord(c->ev2, c->stream));
  lucy_begin_kernel<<<grid_blocks(c), 256, 0, c->stream>>>(c->M, 1);
  CUDA_TRY(cudaGetLastError());
  c->launches_acc = 1;
  if (c->M.use_mrw) {
    mrw_prepare_kernel<<<grid_blocks(c), 256, 0, c->stream>>>(c->M);
    CUDA_TRY(cudaGetLastError());
    c->launches_acc += 1;
  }
  CUDA_TRY(cudaMemsetAsync(c->d_sums + n, 0, SC_COUNT * sizeof(double), c->stream));
  if (c->M.spec_sums)   // grid_reset_energy (grid_generic.f90:26)
    CUDA_TRY(cudaMemsetAsync(c->M.spec_sums, 0, (size_t)c->M.n_spec_bins * n * sizeof(double), c->stream));
  CUDA_TRY(cudaMemsetAsync(c->d_error, 0, sizeof(int32_t), c->stream));
  // grid_reset_energy (grid_generic.f90:18-25): the packet counter exists with the PDA or the n_photons output
  // (grid_physics_3d.f90:308-317)
  if (c->conf.use_pda || c->conf.count_photons) {
    if (c->conf.use_pda && (c->grid_type == GEO_OCT || c->grid_type == GEO_AMR || c->grid_type == GEO_VOR))
      return fail(HYP_ERR_INVALID, "PDA is not available for this grid type");   // grid_pda_disabled.f90
    if (!c->d_nvis) {
      CUDA_TRY(cudaMalloc(&c->d_nvis, (size_t)c->n_cells * sizeof(unsigned long long)));
      CUDA_TRY(cudaMalloc(&c->d_lastid, (size_t)c->n_cells * sizeof(unsigned long long)));
    }
    CUDA_TRY(cudaMemsetAsync(c->d_nvis, 0, (size_t)c->n_cells * sizeof(unsigned long long), c->stream));
    CUDA_TRY(cudaMemsetAsync(c->d_lastid, 0, (size_t)c->n_cells * sizeof(unsigned long long), c->stream));
    c->M.n_visits = c->d_nvis;
    c->M.last_id = c->d_lastid;
  } else {
    c->M.n_visits = c->M.last_id = nullptr;
  }
  c->sums_gathered = false;
  c->kernel_ms_acc = 0.f;
  c->flight_ms_acc = 0.f;
  c->rounds_acc = 0;
  c->wave_rounds_acc = 0;
  return HYP_OK;
}

}  // extern "C"

// The rounds of the packet pool for a fixed number of dust types.  With `handoff` set the pool already
// holds the packets (queues filled by wave_handoff_kernel, every id claimed) and the rounds only finish them.
template <int ND>
static int run_rounds(hyp_ctx *c, int64_t first_id, int64_t n_photons, int64_t iteration, const bool handoff = false,
                      const uint32_t handoff_flights = 0) {
  const uint32_t cap = handoff ? c->pool_cap : (uint32_t)std::min<int64_t>(pool_target(), n_photons);
  int rc = ensure_pool(c, cap);
  if (rc) return rc;
  Pool &P = c->pool;
  // the three wall arrays go to shared memory when they fit next to the resident blocks
  size_t wall_bytes = (size_t)(c->n1 + c->n2 + c->n3 + 3) * sizeof(double);
  int walls_smem = 1;
  if (wall_bytes > 40 * 1024) {
    wall_bytes = 0;
    walls_smem = 0;
  }
  auto flight = flight_kernel<ND, FLIGHT_LOOKAHEAD>;
  auto beam = flight_beam_kernel<ND, FLIGHT_LOOKAHEAD>;
  CUDA_TRY(cudaFuncSetAttribute(flight, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wall_bytes));
  CUDA_TRY(cudaFuncSetAttribute(beam, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wall_bytes));
  int per_sm = 0, per_sm_beam = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, flight, FLIGHT_THREADS, wall_bytes));
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_beam, beam, FLIGHT_THREADS, wall_bytes));
  if (per_sm < 1) per_sm = 1;
  if (per_sm_beam < 1) per_sm_beam = 1;
  const int flight_blocks_max = per_sm * c->sm_count;
  const int beam_blocks_max = per_sm_beam * c->sm_count;
  const int service_blocks_max = c->sm_count * 8;
  cudaStream_t st = c->stream;

  int overlap_b = 2, overlap_f = 2;   // measured best with a 12 M pool (profiles/r01_experiments.md); "0" = one after the other
  if (const char *e = getenv("HYPERION_B200_OVERLAP")) {
    if (sscanf(e, "%d,%d", &overlap_b, &overlap_f) != 2 || overlap_b < 1 || overlap_f < 1) overlap_b = overlap_f = 0;
  }

  if (!handoff) {
    pool_init_kernel<<<c->sm_count, 256, 0, st>>>(P, cap);
    c->launches_acc += 1;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(c->ev0, st));
  }
  int cur = 0;
  uint32_t n_emit = handoff ? 0 : cap, n_flight_prev = handoff ? handoff_flights : 0;
  unsigned long long claimed = handoff ? (unsigned long long)n_photons : 0;
  int64_t windows_ready = 0;
  for (int64_t round = 0;; ++round) {
    uint32_t *nF = P.counts + C_NF0 + cur, *nF_next = P.counts + C_NF0 + (1 - cur);
    while (!handoff && windows_ready * (int64_t)c->sort_window < n_photons &&
           (int64_t)claimed + 2 * (int64_t)cap > windows_ready * (int64_t)c->sort_window) {
      rc = prepare_window(c, first_id, n_photons, iteration, windows_ready);
      if (rc) return rc;
      ++windows_ready;
    }
    // 1. new packets into the free slots -> beam queue
    int64_t n_new = 0;
    if (claimed < (unsigned long long)n_photons && n_emit > 0) {
      n_new = std::min<int64_t>(n_emit, n_photons - (int64_t)claimed);
      int blocks = (int)std::min<int64_t>(((int64_t)n_emit + SERVICE_THREADS - 1) / SERVICE_THREADS, service_blocks_max);
      emit_kernel<ND><<<blocks, SERVICE_THREADS, 0, st>>>(c->M, P, (unsigned long long)first_id,
                                                          (unsigned long long)n_photons, (uint32_t)iteration);
      CUDA_TRY(cudaGetLastError());
      c->launches_acc += 1;
    }
    if (handoff && round == 0) {
      // the interactions pending at the hand-over are already in q_interact: keep its count
      CUDA_TRY(cudaMemsetAsync(P.counts + C_NE, 0, 3 * sizeof(uint32_t), st));  // C_NE, C_CURSOR, C_CURSOR_B
    } else {
      CUDA_TRY(cudaMemsetAsync(P.counts + C_NI, 0, 4 * sizeof(uint32_t), st));  // C_NI, C_NE, C_CURSOR, C_CURSOR_B
    }
    // 2. flights: the beams of new packets, then the packets that come out of an interaction
    CUDA_TRY(cudaEventRecord(c->evA, st));
    if (c->grid_type != GEO_CAR || c->M.any_sphere || c->M.n_visits || c->M.spec_sums) {
      const FinalArgs none = FinalArgs();
      auto geo_flight = c->grid_type == GEO_OCT   ? flight_geo_kernel<GEO_OCT, ND, true, false>
                        : c->grid_type == GEO_AMR ? flight_geo_kernel<GEO_AMR, ND, true, false>
                        : c->grid_type == GEO_VOR ? flight_geo_kernel<GEO_VOR, ND, true, false>
                        : c->grid_type == GEO_CAR ? flight_geo_kernel<GEO_CAR, ND, true, false>
                                                  : flight_geo_kernel<GEO_SPH, ND, true, false>;
      const int sph_blocks_max = c->sm_count * 12;
      if (n_new > 0) {
        int blocks = (int)std::min<int64_t>((n_new + SPH_FLIGHT_THREADS - 1) / SPH_FLIGHT_THREADS, sph_blocks_max);
        geo_flight<<<blocks, SPH_FLIGHT_THREADS, 0, st>>>(c->M, P, none, P.q_beam, P.counts + C_NB,
                                                                                 P.counts + C_CURSOR_B, (uint32_t)iteration);
        c->launches_acc += 1;
      }
      if (n_flight_prev > 0) {
        int blocks = (int)std::min<int64_t>(((int64_t)n_flight_prev + SPH_FLIGHT_THREADS - 1) / SPH_FLIGHT_THREADS, sph_blocks_max);
        geo_flight<<<blocks, SPH_FLIGHT_THREADS, 0, st>>>(c->M, P, none, P.q_flight[cur], nF,
                                                                                 P.counts + C_CURSOR, (uint32_t)iteration);
        c->launches_acc += 1;
      }
      CUDA_TRY(cudaGetLastError());
    } else {
    // The beam kernel (new packets: issue-bound, its cells live in L2) and the flight kernel (packets that
    // left an interaction: bound by scattered REDs) want different resources and touch different packets:
    // they run side by side on two streams, b and f blocks per SM each (HYPERION_B200_OVERLAP="b,f").
    const bool side_by_side = overlap_b > 0 && n_new > 0 && n_flight_prev > 0;
    if (side_by_side) {
      CUDA_TRY(cudaEventRecord(c->evFork, st));
      CUDA_TRY(cudaStreamWaitEvent(c->stream2, c->evFork, 0));
      int blocks = (int)std::min<int64_t>((n_new + FLIGHT_THREADS - 1) / FLIGHT_THREADS, (int64_t)overlap_b * c->sm_count);
      beam<<<blocks, FLIGHT_THREADS, wall_bytes, c->stream2>>>(c->M, P, P.q_beam, P.counts + C_NB, walls_smem);
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaEventRecord(c->evJoin, c->stream2));
      c->launches_acc += 1;
    } else if (n_new > 0) {
      int blocks = (int)std::min<int64_t>((n_new + FLIGHT_THREADS - 1) / FLIGHT_THREADS, beam_blocks_max);
      beam<<<blocks, FLIGHT_THREADS, wall_bytes, st>>>(c->M, P, P.q_beam, P.counts + C_NB, walls_smem);
      CUDA_TRY(cudaGetLastError());
      c->launches_acc += 1;
    }
    if (n_flight_prev > 0) {
      int blocks = (int)std::min<int64_t>(((int64_t)n_flight_prev + FLIGHT_THREADS - 1) / FLIGHT_THREADS,
                                          side_by_side ? (int64_t)overlap_f * c->sm_count : (int64_t)flight_blocks_max);
      flight<<<blocks, FLIGHT_THREADS, wall_bytes, st>>>(c->M, P, P.q_flight[cur], nF, walls_smem);
      CUDA_TRY(cudaGetLastError());
      c->launches_acc += 1;
    }
    if (side_by_side) CUDA_TRY(cudaStreamWaitEvent(st, c->evJoin, 0));
    }
    CUDA_TRY(cudaEventRecord(c->evB, st));
    CUDA_TRY(cudaMemsetAsync(nF_next, 0, sizeof(uint32_t), st));
    CUDA_TRY(cudaMemsetAsync(P.counts + C_NB, 0, sizeof(uint32_t), st));
    // 3. interactions -> next round's flight queue; killed packets free their slot
    interact_kernel<ND><<<service_blocks_max, SERVICE_THREADS, 0, st>>>(c->M, P, P.q_flight[1 - cur], nF_next,
                                                                        (uint32_t)iteration);
    CUDA_TRY(cudaGetLastError());
    c->launches_acc += 1;
    CUDA_TRY(cudaMemcpyAsync(c->h_counts, P.counts, C_COUNT * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(c->h_counts + C_COUNT, P.next_photon, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->evA, c->evB) == cudaSuccess) c->flight_ms_acc += ms;
    c->rounds_acc += 1;
    cur = 1 - cur;
    if (!handoff) memcpy(&claimed, c->h_counts + C_COUNT, sizeof claimed);
    n_emit = c->h_counts[C_NE];
    n_flight_prev = c->h_counts[C_NF0 + cur];
    const bool ids_left = claimed < (unsigned long long)n_photons;
    if (n_flight_prev == 0 && (!ids_left || n_emit == 0)) break;
  }
  CUDA_TRY(cudaEventRecord(c->ev1, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->kernel_ms_acc += ms;
  return HYP_OK;
}

// ---------------------------------------------------------------------------------------------------
// Wave engine (flight_wave.cuh): host side
// ---------------------------------------------------------------------------------------------------
namespace {

// HYPERION_B200_ENGINE = "rounds" keeps the direct kernels for every model; default: the wave engine wherever
// it applies (Cartesian grid with equidistant walls, no spherical source, tiles that fit shared memory).
bool wave_wanted() {
  const char *e = getenv("HYPERION_B200_ENGINE");
  return !(e && strcmp(e, "rounds") == 0);
}

// Blocks of the tile kernel per SM: one block with 2 x 112 KB (densities, sums), or two blocks with 2 x 56 KB each
// (one dust type only; HYPERION_B200_WAVE_CTAS=2).  The offsets are template constants of wave_tile_kernel.
// bytes of the density half (= of the sum half) of a tile; one dust type: 30^3 cells of 4 bytes rounded up to 128,
// which leaves 16 KB of the SM's shared memory to the service blocks that run beside a tile block
// (two and more dust types: blocks of 768 / 512 threads, the largest halves that still leave room for the walls and
// the 12 bytes per thread of packet positions: 24^3 x 8, 21^3 x 12 and 19^3 x 16 bytes fit as before)
constexpr uint32_t WAVE_SUM_OFF_ND2 = 110592, WAVE_SUM_OFF_ND34 = 111616, WAVE_SUM_OFF_2 = 53248, WAVE_SUM_OFF_ND1 = 108032;
#ifndef WAVE_TILE_THREADS_DEFAULT
#define WAVE_TILE_THREADS_DEFAULT 1024
#endif
constexpr uint32_t wave_sum_off(int nd, int ctas) {
  return ctas == 2 ? WAVE_SUM_OFF_2 : (nd == 1 ? WAVE_SUM_OFF_ND1 : (nd == 2 ? WAVE_SUM_OFF_ND2 : WAVE_SUM_OFF_ND34));
}

int wave_ctas(int nd) {
  const char *e = getenv("HYPERION_B200_WAVE_CTAS");
  return (nd == 1 && e && atoi(e) == 2) ? 2 : 1;
}

size_t wave_smem_bytes(int tx, int ty, int tz, int ctas, int nd, int threads) {
  const int tw = std::max(tx, std::max(ty, tz)) + 1;
  // densities | sums | walls of the tile | positions of the three packets every lane holds
  return 2 * (size_t)wave_sum_off(nd, ctas) + (size_t)3 * tw * sizeof(double) + (size_t)3 * threads * sizeof(uint32_t);
}

// Tile shape: the largest cube whose densities fit their half of the shared memory
// (HYPERION_B200_TILE="tx,ty,tz" overrides).
bool wave_plan(hyp_ctx *c, int nd) {
  WaveQ &W = c->wave;
  if (!c->uniform_walls || c->grid_type != GEO_CAR || c->M.any_sphere) return false;
  const int ctas = wave_ctas(nd);
  const size_t cells_max = (size_t)wave_sum_off(nd, ctas) / (4 * (size_t)nd);
  auto fits = [&](int tx, int ty, int tz) {
    return (size_t)(tx + 2) * (ty + 2) * (tz + 2) <= cells_max && std::max(tx, std::max(ty, tz)) <= 62;
  };
  int tx = 0, ty = 0, tz = 0;
  if (const char *e = getenv("HYPERION_B200_TILE")) {
    if (sscanf(e, "%d,%d,%d", &tx, &ty, &tz) != 3 || tx < 1 || ty < 1 || tz < 1) tx = ty = tz = 0;
  }
  if (tx == 0) {
    // the largest cube that fits (longer visits beat evenly cut axes: 28^3 tiles 58.2 ms per step of the 256^3
    // headline, 26^3 = ten equal parts per axis 60.6 ms; profiles/r02_experiments.md)
    for (int t0 = 62; t0 >= 2; --t0) {
      const int cx = std::min(t0, c->n1), cy = std::min(t0, c->n2), cz = std::min(t0, c->n3);
      if (fits(cx, cy, cz)) {
        tx = cx; ty = cy; tz = cz;
        break;
      }
    }
  }
  if (tx == 0) return false;
  tx = std::min(tx, c->n1); ty = std::min(ty, c->n2); tz = std::min(tz, c->n3);
  if (!fits(tx, ty, tz)) return false;
  W.tx = tx; W.ty = ty; W.tz = tz;
  W.ntx = (c->n1 + tx - 1) / tx;
  W.nty = (c->n2 + ty - 1) / ty;
  W.ntz = (c->n3 + tz - 1) / tz;
  const int64_t nt = (int64_t)W.ntx * W.nty * W.ntz;
  if (nt + 2 > (int64_t)WAVE_MAX_BINS) return false;
  W.n_tiles = (int)nt;
  W.dx = (c->w1[c->n1] - c->w1[0]) / c->n1;
  W.dy = (c->w2[c->n2] - c->w2[0]) / c->n2;
  W.dz = (c->w3[c->n3] - c->w3[0]) / c->n3;
  // fixed-point scale of the deposits: the largest len * kappa * E any packet can produce maps to 2^17
  W.diag = std::sqrt(W.dx * W.dx + W.dy * W.dy + W.dz * W.dz);
  double e_max = 1.0;
  if (c->conf.sample_sources_evenly && c->energy_total > 0.0)
    for (const hyp_source &sr : c->sources) e_max = std::max(e_max, sr.luminosity / c->energy_total * (double)c->sources.size());
  for (int id = 0; id < nd; ++id) {
    const HostDust &d = c->dust[id];
    double kmax = 0.0;
    for (size_t i = 0; i < d.chi.size(); ++i) kmax = std::max(kmax, d.chi[i] * (1.0 - d.albedo[i]));
    const double bound = W.diag * kmax * e_max * 1.001;
    if (!(bound > 0.0) || !std::isfinite(bound)) return false;
    W.dep_scale[id] = (double)WAVE_DEP_MAX / bound;
    W.dep_inv[id] = bound / (double)WAVE_DEP_MAX;
  }
  return true;
}

int ensure_wave(hyp_ctx *c) {
  WaveQ &W = c->wave;
  const uint32_t cap = c->pool_cap;
  const int nb = W.n_tiles + 2;
  if (W.capacity >= cap && c->wave_bins_alloc >= nb) {
    W.capacity = cap;
    return HYP_OK;
  }
  free_dev(W.key); free_dev(W.key_pos); free_dev(c->wave_sorted[0]); free_dev(c->wave_sorted[1]); free_dev(W.bin_count); free_dev(W.bin_cursor); free_dev(W.items); free_dev(W.ctl);
  CUDA_TRY(cudaMalloc(&W.key, (size_t)cap * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&W.key_pos, (size_t)cap * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&c->wave_sorted[0], (size_t)cap * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&c->wave_sorted[1], (size_t)cap * sizeof(uint32_t)));
  W.sorted = c->wave_sorted[0];
  CUDA_TRY(cudaMalloc(&W.bin_count, (size_t)nb * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&W.bin_cursor, (size_t)nb * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&W.items, ((size_t)nb + cap / 256 + 2) * sizeof(uint4)));
  CUDA_TRY(cudaMalloc(&W.ctl, WC_COUNT * sizeof(uint32_t)));
  W.capacity = cap;
  c->wave_bins_alloc = nb;
  return HYP_OK;
}

}  // namespace

template <int ND>
static int run_wave(hyp_ctx *c, int64_t first_id, int64_t n_photons, int64_t iteration) {
  const uint32_t cap = (uint32_t)std::min<int64_t>(pool_target(true), n_photons);
  int rc = ensure_pool(c, cap);
  if (rc) return rc;
  rc = ensure_wave(c);
  if (rc) return rc;
  Pool &P = c->pool;
  WaveQ &W = c->wave;
  W.capacity = cap;
  {
    long ch = getenv("HYPERION_B200_WAVE_CHUNK") ? atol(getenv("HYPERION_B200_WAVE_CHUNK")) : 16384L;
    ch = std::max(256L, std::min<long>(ch, (long)WAVE_CHUNK_MAX));
    W.chunk = (uint32_t)ch;
  }
  // below this many packets in flight the tiles are mostly empty: the direct kernels finish the iteration
  const uint32_t tail_min = getenv("HYPERION_B200_WAVE_TAIL") ? (uint32_t)atol(getenv("HYPERION_B200_WAVE_TAIL")) : 1000000u;
  W.refill = getenv("HYPERION_B200_WAVE_REFILL") ? std::max(1, std::min(32, atoi(getenv("HYPERION_B200_WAVE_REFILL")))) : 12;
  {
    // HYPERION_B200_WAVE_EMIT: new packets per round at most.  Spreading the emission over several rounds lets it
    // run next to tile visits, but measured slower (75.5 vs 72.8 ms per step) than filling the pool at once:
    // the first rounds then hold fewer packets.
    const long q = getenv("HYPERION_B200_WAVE_EMIT") ? atol(getenv("HYPERION_B200_WAVE_EMIT")) : (long)cap;
    W.emit_max = (uint32_t)std::max(1024L, std::min<long>(q, (long)cap));
  }
  W.iteration = (uint32_t)iteration;
  const int ctas = wave_ctas(ND);
  W.queue = getenv("HYPERION_B200_WAVE_QUEUE") ? atoi(getenv("HYPERION_B200_WAVE_QUEUE")) : 1;
  constexpr int WT = ND == 1 ? 1024 : (ND == 2 ? 768 : 512);
  void (*tile)(const ModelDev, Pool, const WaveQ) = wave_tile_kernel<ND, WT, 1, wave_sum_off(ND, 1)>;
  int tile_threads = WT;
  if constexpr (ND == 1) {
    // HYPERION_B200_WAVE_THREADS: threads of a tile block, all with the 64 registers of a 1024-thread block; below
    // 1024 the interaction (and emission) blocks of the round find registers on the same SM
    const int want = getenv("HYPERION_B200_WAVE_THREADS") ? atoi(getenv("HYPERION_B200_WAVE_THREADS")) : WAVE_TILE_THREADS_DEFAULT;
    if (ctas == 2) {
      tile = wave_tile_kernel<ND, 512, 2, WAVE_SUM_OFF_2>;
      tile_threads = 512;
    } else if (want == 1024 && W.tx == 28 && W.ty == 28 && W.tz == 28 && !getenv("HYPERION_B200_WAVE_NOCUBE")) {
      tile = wave_tile_kernel<ND, 1024, 1, WAVE_SUM_OFF_ND1, 1024, 28>;   // the default tile of large grids
    } else if (want == 896) {
      tile = wave_tile_kernel<ND, 896, 1, WAVE_SUM_OFF_ND1, 1024>;
      tile_threads = 896;
    } else if (want == 768) {
      tile = wave_tile_kernel<ND, 768, 1, WAVE_SUM_OFF_ND1, 1024>;
      tile_threads = 768;
    } else if (want == 640) {
      tile = wave_tile_kernel<ND, 640, 1, WAVE_SUM_OFF_ND1, 1024>;
      tile_threads = 640;
    }
  }
  // HYPERION_B200_WAVE_ORDER: 1 (default) the tile kernel is launched before the interactions and the emission of
  // the round, so that its blocks take the SMs first and the service blocks fill what is left (and the SMs of tile
  // blocks that have run out of items); 0 the other way round
  const bool tile_first = getenv("HYPERION_B200_WAVE_ORDER") ? atoi(getenv("HYPERION_B200_WAVE_ORDER")) != 0 : true;
  const size_t tile_smem = wave_smem_bytes(W.tx, W.ty, W.tz, ctas, ND, tile_threads);
  CUDA_TRY(cudaFuncSetAttribute(tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem));
  const size_t sort_smem = 2 * (size_t)(W.n_tiles + 2) * sizeof(uint32_t);
  CUDA_TRY(cudaFuncSetAttribute(wave_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem));
  CUDA_TRY(cudaFuncSetAttribute(wave_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem));
  cudaStream_t st = c->stream, s2 = c->stream2, s3 = c->stream3;
  const bool dbg = getenv("HYPERION_B200_TIMING") != nullptr;
  // HYPERION_B200_TIMING=2: every kernel of a round runs alone (a device synchronisation after each) and the summed
  // times per kernel are printed: the step without any overlap, with warm caches and at the clocks of a real run
  const bool serial = dbg && atoi(getenv("HYPERION_B200_TIMING")) == 2;
  double t_acc[5] = {0, 0, 0, 0, 0};   // sort, tile, interact, emit, host gaps
  auto lap = [&](int which, cudaStream_t s, bool begin) {
    if (!serial) return;
    if (begin) {
      cudaDeviceSynchronize();
      cudaEventRecord(c->evA, s);
    } else {
      cudaEventRecord(c->evB, s);
      cudaEventSynchronize(c->evB);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, c->evA, c->evB);
      t_acc[which] += ms;
    }
  };
  // HYPERION_B200_WAVE_SERVICE: blocks per SM of the interaction / emission kernels of a round that has tile visits
  const int service_per_sm = getenv("HYPERION_B200_WAVE_SERVICE") ? std::max(1, atoi(getenv("HYPERION_B200_WAVE_SERVICE"))) : 16;
  const int service_blocks = c->sm_count * service_per_sm;

  wave_init_kernel<<<c->sm_count * 4, 256, 0, st>>>(W, P);
  CUDA_TRY(cudaGetLastError());
  c->launches_acc += 1;
  CUDA_TRY(cudaEventRecord(c->ev0, st));
  uint32_t *h = c->h_counts;
  bool handoff = false, compact = false;
  const bool speculative = getenv("HYPERION_B200_WAVE_SPEC") ? atoi(getenv("HYPERION_B200_WAVE_SPEC")) != 0 : false;
  bool stop_next = false, spec_emit = true;
  int n_tile_rounds = 0;
  uint32_t handoff_flights = 0, n_busy_prev = 0;
  for (int64_t round = 0;; ++round) {
    // slots to sort: all of them while packets are still being emitted, afterwards the busy slots of the
    // previous round (a slot freed after the last emission stays free)
    const uint32_t *src = compact ? c->wave_sorted[(round + 1) & 1] : nullptr;
    const uint32_t n_src = compact ? n_busy_prev : cap;
    W.sorted = c->wave_sorted[round & 1];
    const int sort_blocks = (int)std::max<uint32_t>(1u, (n_src + WAVE_SORT_SEG - 1) / WAVE_SORT_SEG);
    lap(0, st, true);
    wave_hist_kernel<<<sort_blocks, WAVE_SORT_THREADS, sort_smem / 2, st>>>(W, src, n_src);
    wave_scan_kernel<<<1, 1024, 0, st>>>(W, P, (unsigned long long)n_photons);
    wave_scatter_kernel<<<sort_blocks, WAVE_SORT_THREADS, sort_smem, st>>>(W, src, n_src);
    lap(0, st, false);
    CUDA_TRY(cudaGetLastError());
    c->launches_acc += 3;
    CUDA_TRY(cudaMemcpyAsync(h, W.ctl, WC_COUNT * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(c->evCtl, st));
    // The kernels of a round read their counts from device memory, so they CAN be launched before the host has seen
    // the counts of this round's sort (HYPERION_B200_WAVE_SPEC=1): the GPU then does not wait for the host between
    // the sort and the tile kernel, and what the host decides from the counts (end of the loop, hand-off to the
    // direct kernels, which slots the next sort covers) takes effect one round later.  Measured: 55.8 against
    // 55.4 ms per step (one more round, full-size grids for empty kernels), so the default waits for the counts.
    uint32_t n_flight = 0, n_interact = 0, n_free = 0;
    bool ids_left = true;
    auto read_counts = [&]() -> int {
      CUDA_TRY(cudaEventSynchronize(c->evCtl));
      n_flight = h[WC_N_FLIGHT]; n_interact = h[WC_N_INTERACT]; n_free = h[WC_N_FREE];
      const unsigned long long claimed = (unsigned long long)h[WC_CLAIMED_LO] | ((unsigned long long)h[WC_CLAIMED_HI] << 32);
      ids_left = claimed < (unsigned long long)n_photons;
      if (dbg)
        fprintf(stderr, "[wave %lld] flights %u in %u items, interactions %u, free %u, claimed %llu\n", (long long)round,
                n_flight, h[WC_NITEMS], n_interact, n_free, claimed);
      return HYP_OK;
    };
    // known = the host has this round's counts: empty kernels are skipped and the grids sized to the work
    auto launch_round = [&](const bool known) -> int {
      const bool want_tile = !known || n_flight > 0, want_interact = !known || n_interact > 0;
      const bool want_emit = known ? (n_free > 0 && ids_left) : spec_emit;
      CUDA_TRY(cudaEventRecord(c->evFork, st));
      auto launch_tile = [&]() -> int {
        if (!want_tile) return HYP_OK;
        // the tile kernel's own time: one pair of events per round, read when the photon loop has ended
        while (c->wave_ev.size() < 2 * (size_t)(n_tile_rounds + 1)) {
          cudaEvent_t e;
          CUDA_TRY(cudaEventCreate(&e));
          c->wave_ev.push_back(e);
        }
        const uint32_t grid_max = (uint32_t)(c->sm_count * ctas);
        lap(1, st, true);
        CUDA_TRY(cudaEventRecord(c->wave_ev[2 * n_tile_rounds], st));
        tile<<<(int)(known ? std::min<uint32_t>(h[WC_NITEMS], grid_max) : grid_max), tile_threads, tile_smem, st>>>(c->M, P, W);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(c->wave_ev[2 * n_tile_rounds + 1], st));
        lap(1, st, false);
        ++n_tile_rounds;
        c->launches_acc += 1;
        return HYP_OK;
      };
      if (tile_first) {
        rc = launch_tile();
        if (rc) return rc;
      }
      if (want_interact) {
        CUDA_TRY(cudaStreamWaitEvent(s2, c->evFork, 0));
        const int blocks = known ? (int)std::min<int64_t>(((int64_t)n_interact + SERVICE_THREADS - 1) / SERVICE_THREADS, service_blocks)
                                 : service_blocks;
        lap(2, s2, true);
        wave_interact_kernel<ND><<<blocks, SERVICE_THREADS, 0, s2>>>(c->M, P, W, (uint32_t)iteration);
        lap(2, s2, false);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(c->evJoin, s2));
        c->launches_acc += 1;
      }
      if (want_emit) {
        CUDA_TRY(cudaStreamWaitEvent(s3, c->evFork, 0));
        const int64_t n_emit = known ? std::min<int64_t>(n_free, W.emit_max) : (int64_t)W.emit_max;
        const int blocks = (int)std::min<int64_t>((n_emit + SERVICE_THREADS - 1) / SERVICE_THREADS, service_blocks);
        lap(3, s3, true);
        wave_emit_kernel<ND><<<blocks, SERVICE_THREADS, 0, s3>>>(c->M, P, W, (unsigned long long)first_id,
                                                               (unsigned long long)n_photons, (uint32_t)iteration);
        lap(3, s3, false);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(c->evJoin3, s3));
        c->launches_acc += 1;
      }
      if (!tile_first) {
        rc = launch_tile();
        if (rc) return rc;
      }
      if (want_interact) CUDA_TRY(cudaStreamWaitEvent(st, c->evJoin, 0));
      if (want_emit) CUDA_TRY(cudaStreamWaitEvent(st, c->evJoin3, 0));
      return HYP_OK;
    };
    bool launched = false;
    if (speculative && !stop_next) {
      rc = launch_round(false);
      if (rc) return rc;
      launched = true;
    }
    rc = read_counts();
    if (rc) return rc;
    if (n_flight == 0 && n_interact == 0 && !ids_left) break;   // (kernels launched ahead found nothing to do)
    // this round emits nothing if every id was claimed before its sort: the next sort only needs its busy slots
    compact = !ids_left;
    spec_emit = ids_left;
    n_busy_prev = n_flight + n_interact;
    const bool tail = !ids_left && n_flight + n_interact < tail_min;
    if (tail && !launched) {
      wave_handoff_kernel<<<c->sm_count, 256, 0, st>>>(P, W);
      CUDA_TRY(cudaGetLastError());
      c->launches_acc += 1;
      handoff = true;
      handoff_flights = n_flight;
      break;
    }
    stop_next = tail;   // launched ahead: this round runs on the wave engine, the hand-off follows the next sort
    c->rounds_acc += 1;
    c->wave_rounds_acc += 1;
    if (!launched) {
      rc = launch_round(true);
      if (rc) return rc;
    }
  }
  auto tile_times = [&]() -> int {
    if (n_tile_rounds == 0) return HYP_OK;
    CUDA_TRY(cudaEventSynchronize(c->wave_ev[2 * n_tile_rounds - 1]));
    for (int r = 0; r < n_tile_rounds; ++r) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, c->wave_ev[2 * r], c->wave_ev[2 * r + 1]) == cudaSuccess) c->flight_ms_acc += ms;
      if (dbg) fprintf(stderr, "[wave tile round %d] %.3f ms\n", r, ms);
    }
    return HYP_OK;
  };
  rc = tile_times();
  if (rc) return rc;
  if (serial)
    fprintf(stderr, "[wave serial] sort %.3f ms, tile %.3f ms, interact %.3f ms, emit %.3f ms (each kernel alone)\n", t_acc[0],
            t_acc[1], t_acc[2], t_acc[3]);
  if (handoff) return run_rounds<ND>(c, first_id, n_photons, iteration, true, handoff_flights);
  CUDA_TRY(cudaEventRecord(c->ev1, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->kernel_ms_acc += ms;
  return HYP_OK;
}

extern "C" {

int hyp_lucy_photons(hyp_ctx *c, int64_t first_id, int64_t n_photons, int64_t iteration) {
  if (!c || !c->finalized) return fail(HYP_ERR_STATE, "hyp_finalize_setup has not been called");
  if (n_photons < 0 || first_id < 0) return fail(HYP_ERR_INVALID, "negative photon count");
  if (n_photons == 0) return HYP_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  c->sums_gathered = false;
  // the packet counter lives in the generic march: neither the tiles nor the beams keep it
  const bool wave = !c->M.n_visits && !c->M.spec_sums && wave_wanted() && wave_plan(c, c->M.n_dust);
  c->last_engine = wave ? 1 : 0;
  switch (c->M.n_dust) {
    case 1: return wave ? run_wave<1>(c, first_id, n_photons, iteration) : run_rounds<1>(c, first_id, n_photons, iteration);
    case 2: return wave ? run_wave<2>(c, first_id, n_photons, iteration) : run_rounds<2>(c, first_id, n_photons, iteration);
    case 3: return wave ? run_wave<3>(c, first_id, n_photons, iteration) : run_rounds<3>(c, first_id, n_photons, iteration);
    case 4: return wave ? run_wave<4>(c, first_id, n_photons, iteration) : run_rounds<4>(c, first_id, n_photons, iteration);
    default: return fail(HYP_ERR_INVALID, "unsupported number of dust types");
  }
}

}  // extern "C"

namespace {

// cell_width / geometrical_factor of the regular grids as per-axis tables (grid_geometry_*_3d.f90 cell_width,
// grid_pda_*_3d.f90 geometrical_factor); layout: 9 width tables (direction-major, axis-minor), then 4 wall tables
int pda_geometry(hyp_ctx *c, PdaGeo &G) {
  const int n1 = c->n1, n2 = c->n2, n3 = c->n3;
  const int ns[3] = {n1, n2, n3};
  const std::vector<double> *ws[3] = {&c->w1, &c->w2, &c->w3};
  std::vector<double> buf;
  size_t off_w[3][3], off_f[4];
  auto push = [&](const std::vector<double> &v) { const size_t o = buf.size(); buf.insert(buf.end(), v.begin(), v.end()); return o; };
  auto ones = [&](int n) { return std::vector<double>(n, 1.0); };
  auto diff = [&](int a) { std::vector<double> v(ns[a]); for (int i = 0; i < ns[a]; ++i) v[i] = (*ws[a])[i + 1] - (*ws[a])[i]; return v; };
  auto mid_log = [&]() {   // geo%r / geo%w (grid_geometry_spherical_3d.f90:139-143, _cylindrical_3d.f90:130-134)
    std::vector<double> v(n1);
    for (int i = 0; i < n1; ++i)
      v[i] = c->w1[i] == 0.0 ? c->w1[i + 1] / 2.0 : std::pow(10.0, (std::log10(c->w1[i]) + std::log10(c->w1[i + 1])) / 2.0);
    return v;
  };
  const bool car = c->grid_type == GEO_CAR, sph = c->grid_type == GEO_SPH && c->polar_kind == POLAR_SPH;
  for (int d = 0; d < 3; ++d)
    for (int a = 0; a < 3; ++a) {
      std::vector<double> v = ones(ns[a]);
      if (car) {
        if (a == d) v = diff(a);
      } else if (sph) {
        if (d == 0 && a == 0) v = diff(0);
        if (d == 1 && a == 0) v = mid_log();
        if (d == 1 && a == 1) v = diff(1);
        if (d == 2 && a == 0) v = mid_log();
        if (d == 2 && a == 1) for (int i = 0; i < n2; ++i) v[i] = std::sin((c->w2[i] + c->w2[i + 1]) / 2.0);
        if (d == 2 && a == 2) v = diff(2);
      } else {
        if (d == 0 && a == 0) v = diff(0);
        if (d == 1 && a == 1) v = diff(1);
        if (d == 2 && a == 0) v = mid_log();
        if (d == 2 && a == 2) v = diff(2);
      }
      off_w[d][a] = push(v);
    }
  for (int w = 0; w < 4; ++w) {
    std::vector<double> v = ones(w < 2 ? n1 : n2);
    if (!car && w < 2)
      for (int i = 0; i < n1; ++i) {
        const double a = c->w1[i], b = c->w1[i + 1], x = w == 0 ? a : b;
        v[i] = sph ? 4. * x * x / ((a + b) * (a + b)) : 2. * x / (a + b);
      }
    if (sph && w >= 2)
      for (int i = 0; i < n2; ++i) {
        const double sa = std::sin(c->w2[i]), sb = std::sin(c->w2[i + 1]);
        v[i] = 2. * (w == 2 ? sa : sb) / (sa + sb);
      }
    off_f[w] = push(v);
  }
  free_dev(c->d_pda_geo);
  CUDA_TRY(cudaMalloc(&c->d_pda_geo, buf.size() * sizeof(double)));
  CUDA_TRY(cudaMemcpy(c->d_pda_geo, buf.data(), buf.size() * sizeof(double), cudaMemcpyHostToDevice));
  for (int d = 0; d < 3; ++d)
    for (int a = 0; a < 3; ++a) G.W[d][a] = c->d_pda_geo + off_w[d][a];
  for (int w = 0; w < 4; ++w) G.F[w] = c->d_pda_geo + off_f[w];
  G.n1 = n1; G.n2 = n2; G.n3 = n3;
  G.periodic3 = car ? 0 : 1;
  G.n_dim = (!car && n3 == 1) ? 2 : 3;
  return HYP_OK;
}

// solve_pda (grid_pda_3d.f90:105-169) on the current specific energy.  counts: device array of n_photons as
// doubles; n_pda_out: number of PDA cells.
int solve_pda_device(hyp_ctx *c, const double *counts, int64_t *n_pda_out) {
  if (c->grid_type == GEO_OCT || c->grid_type == GEO_AMR || c->grid_type == GEO_VOR)
    return fail(HYP_ERR_INVALID, "PDA is not available for this grid type");
  cudaStream_t st = c->stream;
  const size_t nc = (size_t)c->n_cells;
  PdaDev P;
  memset(&P, 0, sizeof P);
  int rc = pda_geometry(c, P.G);
  if (rc) return rc;
  if (!c->d_pda_ctl) CUDA_TRY(cudaMalloc(&c->d_pda_ctl, 4 * sizeof(unsigned long long)));
  if (!c->d_pda_list) CUDA_TRY(cudaMalloc(&c->d_pda_list, nc * sizeof(int32_t)));
  if (!c->d_pda_e) CUDA_TRY(cudaMalloc(&c->d_pda_e, 2 * nc * sizeof(double)));
  CUDA_TRY(cudaMemsetAsync(c->d_pda_ctl, 0, 4 * sizeof(unsigned long long), st));
  P.counts = counts;
  P.list = c->d_pda_list;
  P.n_list = (uint32_t *)c->d_pda_ctl;
  P.maxdiff = c->d_pda_ctl + 1;
  P.e_a = c->d_pda_e;
  P.e_b = c->d_pda_e + nc;
  // mean_n_photons = sum(n_photons) / size(n_photons); threshold max(30, ceiling(0.005 mean)) (:124-129)
  pda_sum_counts_kernel<<<grid_blocks(c), 256, 0, st>>>(counts, (int64_t)nc, (double *)(c->d_pda_ctl + 2));
  double total = 0.0;
  CUDA_TRY(cudaMemcpyAsync(&total, c->d_pda_ctl + 2, sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  P.limit = std::max(30.0, std::ceil(0.005 * (total / (double)nc)));
  pda_mark_kernel<<<grid_blocks(c), 256, 0, st>>>(c->M, P);
  uint32_t n_pda = 0;
  CUDA_TRY(cudaMemcpyAsync(&n_pda, c->d_pda_ctl, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  c->launches_acc += 2;
  c->pda_cells_last = n_pda;
  c->pda_sweeps_last = 0;
  if (n_pda_out) *n_pda_out = n_pda;
  if (n_pda == 0) return HYP_OK;   // " [pda] not necessary for this iteration"
  free_dev(c->d_pda_coef);
  CUDA_TRY(cudaMalloc(&c->d_pda_coef, (size_t)n_pda * 6 * sizeof(double)));
  P.coef = c->d_pda_coef;
  const double tolerance = n_pda < 10000 ? 1.e-5 : 1.e-4;   // tolerance_exact / tolerance_iter (:34-35,138-144)
  const int blocks = (int)std::min<int64_t>(((int64_t)n_pda + 255) / 256, (int64_t)grid_blocks(c));
  const int batch = 64;
  for (int outer = 0; outer < 1000; ++outer) {
    pda_emean_kernel<<<grid_blocks(c), 256, 0, st>>>(c->M, P);
    pda_coef_kernel<<<blocks, 256, 0, st>>>(c->M, P);
    c->launches_acc += 2;
    const double *cur = P.e_a;
    for (int64_t sweep = 0; sweep < 4000000; sweep += batch) {
      CUDA_TRY(cudaMemsetAsync(P.maxdiff, 0, sizeof(unsigned long long), st));
      for (int k = 0; k < batch; ++k) {
        double *dst = cur == P.e_a ? P.e_b : P.e_a;
        pda_sweep_kernel<<<blocks, 256, 0, st>>>(P, cur, dst, k == batch - 1 ? 1 : 0);
        cur = dst;
      }
      CUDA_TRY(cudaGetLastError());
      c->launches_acc += batch;
      c->pda_sweeps_last += batch;
      double worst = 0.0;
      CUDA_TRY(cudaMemcpyAsync(&worst, P.maxdiff, sizeof(double), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      if (worst < 1.e-9) break;
    }
    CUDA_TRY(cudaMemsetAsync(P.maxdiff, 0, sizeof(unsigned long long), st));
    pda_update_energy_kernel<<<blocks, 256, 0, st>>>(c->M, P, cur);
    CUDA_TRY(cudaGetLastError());
    c->launches_acc += 1;
    double maxdiff = 0.0;
    CUDA_TRY(cudaMemcpyAsync(&maxdiff, P.maxdiff, sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (maxdiff < tolerance) break;
  }
  // update_energy_abs_tot + check_energy_abs (:164-166)
  clamp_energy_kernel<<<grid_blocks(c), 256, 0, st>>>(c->M);
  CUDA_TRY(cudaGetLastError());
  c->launches_acc += 1;
  return HYP_OK;
}

}  // namespace

extern "C" {

static int gather_sums(hyp_ctx *c) {
  if (!c->sums_gathered) {
    gather_sums_kernel<<<grid_blocks(c), 256, 0, c->stream>>>(c->M, c->d_sums);
    CUDA_TRY(cudaGetLastError());
    c->launches_acc += 1;
    if (c->M.n_visits) {
      pda_counts_to_double_kernel<<<grid_blocks(c), 256, 0, c->stream>>>(
          c->M.n_visits, c->n_cells, c->d_sums + (size_t)c->n_cells * c->dust.size() + SC_COUNT);
      CUDA_TRY(cudaGetLastError());
      c->launches_acc += 1;
    }
    c->sums_gathered = true;
  }
  return HYP_OK;
}

int hyp_lucy_device_buffers(hyp_ctx *c, void **sum_and_scalars, int64_t *n_values) {
  if (!c || !c->finalized) return fail(HYP_ERR_STATE, "hyp_finalize_setup has not been called");
  CUDA_TRY(cudaSetDevice(c->device));
  int rc = gather_sums(c);
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (sum_and_scalars) *sum_and_scalars = c->d_sums;
  // [sums | scalars | n_photons | spectrum sums]: the tail only as far as it is in use
  if (n_values)
    *n_values = c->n_cells * (int64_t)c->dust.size() + SC_COUNT +
                (c->M.spec_sums ? c->n_cells + (int64_t)c->M.n_spec_bins * c->n_cells * (int64_t)c->dust.size()
                                : (c->M.n_visits ? c->n_cells : 0));
  return HYP_OK;
}

int hyp_lucy_finish(hyp_ctx *c, hyp_iter_stats *st) {
  if (!c || !c->finalized) return fail(HYP_ERR_STATE, "hyp_finalize_setup has not been called");
  CUDA_TRY(cudaSetDevice(c->device));
  int rc = gather_sums(c);
  if (rc) return rc;
  const size_t n = (size_t)c->n_cells * c->dust.size();
  double sc[SC_COUNT];
  CUDA_TRY(cudaMemcpyAsync(sc, c->d_sums + n, sizeof sc, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  rc = device_error_to_status(c);
  if (rc) return rc;
  if (!(sc[SC_ENERGY] > 0.0)) return fail(HYP_ERR_STATE, "no photons were emitted in this iteration");
  // update_energy_abs(energy_total / energy_current)  (iter_lucy.f90:224)
  const double scale = c->energy_total / sc[SC_ENERGY];
  lucy_finish_kernel<<<grid_blocks(c), 256, 0, c->stream>>>(c->M, c->d_sums, scale, c->conf.use_pda ? 1 : 0);
  CUDA_TRY(cudaGetLastError());
  c->launches_acc += 1;
  if (c->conf.use_pda) {
    // iter_lucy.f90:227: the PDA between update_energy_abs and sublimate_dust, with the packet counts of all processes
    rc = solve_pda_device(c, c->d_sums + n + SC_COUNT, nullptr);
    if (rc) return rc;
    lucy_finish_kernel<<<grid_blocks(c), 256, 0, c->stream>>>(c->M, c->d_sums, scale, 2);
    CUDA_TRY(cudaGetLastError());
    c->launches_acc += 1;
  }
  CUDA_TRY(cudaEventRecord(c->ev3, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (st) {
    memset(st, 0, sizeof *st);
    st->energy_emitted = sc[SC_ENERGY];
    st->n_photons = (int64_t)sc[SC_PHOTONS];
    st->killed_geo = (int64_t)sc[SC_KILLED_GEO];
    st->killed_int = (int64_t)sc[SC_KILLED_INT];
    st->n_crossings = (int64_t)sc[SC_CROSS];
    st->n_absorptions = (int64_t)sc[SC_ABS];
    st->n_scatterings = (int64_t)sc[SC_SCAT];
    st->n_escaped = (int64_t)sc[SC_ESC];
    float total = 0.f;
    st->kernel_ms = c->kernel_ms_acc;
    st->flight_ms = c->flight_ms_acc;
    st->n_rounds = c->rounds_acc;
    st->n_wave_rounds = c->wave_rounds_acc;
    st->n_launches = c->launches_acc;
    if (cudaEventElapsedTime(&total, c->ev2, c->ev3) == cudaSuccess) st->epilogue_ms = total - c->kernel_ms_acc;
  }
  return HYP_OK;
}

int hyp_run_lucy_iteration(hyp_ctx *c, int64_t n_photons, int64_t iteration, hyp_iter_stats *st) {
  int rc = hyp_lucy_begin(c);
  if (rc) return rc;
  rc = hyp_lucy_photons(c, 0, n_photons, iteration);
  if (rc) return rc;
  return hyp_lucy_finish(c, st);
}

static int get_grid(hyp_ctx *c, int which, double *out) {
  if (!c || !c->finalized) return fail(HYP_ERR_STATE, "hyp_finalize_setup has not been called");
  if (!out) return fail(HYP_ERR_INVALID, "NULL argument");
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t n = (size_t)c->n_cells * c->dust.size();
  if (which == 2) {
    int rc = gather_sums(c);
    if (rc) return rc;
  }
  to_file_order_kernel<<<grid_blocks(c), 256, 0, c->stream>>>(c->M, which, c->d_sums, c->d_stage);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(out, c->d_stage, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return HYP_OK;
}

int hyp_get_specific_energy(hyp_ctx *c, double *out) { return get_grid(c, 0, out); }
int hyp_get_density(hyp_ctx *c, double *out) { return get_grid(c, 1, out); }
int hyp_get_energy_sum(hyp_ctx *c, double *out) { return get_grid(c, 2, out); }

int hyp_set_specific_energy_spectrum_bins(hyp_ctx *c, int32_t n_edges, const double *edges) {
  if (!c || !edges) return fail(HYP_ERR_INVALID, "NULL argument");
  if (c->finalized) return fail(HYP_ERR_STATE, "model is frozen");
  if (n_edges < 2) return fail(HYP_ERR_INVALID, "specific_energy_spectrum_bin_edges should have at least two values");
  for (int i = 0; i < n_edges; ++i)
    if (!(edges[i] > 0.0)) return fail(HYP_ERR_INVALID, "specific_energy_spectrum_bin_edges should be positive");
  for (int i = 1; i < n_edges; ++i)
    if (edges[i] <= edges[i - 1])   // grid_physics_3d.f90:126-128
      return fail(HYP_ERR_INVALID, "specific_energy_spectrum_bin_edges should be strictly increasing");
  c->spec_edges.assign(edges, edges + n_edges);
  return HYP_OK;
}

int hyp_get_specific_energy_spectrum(hyp_ctx *c, double *out) {
  if (!c || !c->finalized) return fail(HYP_ERR_STATE, "hyp_finalize_setup has not been called");
  if (!out) return fail(HYP_ERR_INVALID, "NULL argument");
  if (!c->M.spec_energy) return fail(HYP_ERR_STATE, "specific_energy_spectrum array is not allocated");   // grid_generic.f90:86
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t n = (size_t)c->n_cells * c->dust.size();
  for (int b = 0; b < c->M.n_spec_bins; ++b) {
    to_file_order_kernel<<<grid_blocks(c), 256, 0, c->stream>>>(c->M, 2, c->M.spec_energy + (size_t)b * n, c->d_stage);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out + (size_t)b * n, c->d_stage, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  return HYP_OK;
}

int hyp_get_n_photons(hyp_ctx *c, int64_t *out) {
  if (!c || !c->finalized || !out) return fail(HYP_ERR_STATE, "hyp_finalize_setup has not been called");
  if (!c->d_nvis) return fail(HYP_ERR_STATE, "n_photons array is not allocated");   // output_grid, grid_generic.f90:44
  CUDA_TRY(cudaSetDevice(c->device));
  // the counts of all processes once the host has reduced the buffer (they travel behind the scalars)
  const size_t n = (size_t)c->n_cells * c->dust.size(), nc = (size_t)c->n_cells;
  std::vector<double> h(nc);
  CUDA_TRY(cudaMemcpyAsync(h.data(), c->d_sums + n + SC_COUNT, nc * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  for (size_t i = 0; i < nc; ++i) out[i] = (int64_t)h[i];
  return HYP_OK;
}

int hyp_solve_pda(hyp_ctx *c, const int64_t *n_photons, int64_t *n_pda_cells) {
  if (!c || !c->finalized) return fail(HYP_ERR_STATE, "hyp_finalize_setup has not been called");
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t n = (size_t)c->n_cells * c->dust.size(), nc = (size_t)c->n_cells;
  const double *counts = c->d_sums + n + SC_COUNT;
  if (n_photons) {
    std::vector<double> h(nc);
    for (size_t i = 0; i < nc; ++i) h[i] = (double)n_photons[i];
    if (!c->d_pda_counts) CUDA_TRY(cudaMalloc(&c->d_pda_counts, nc * sizeof(double)));
    CUDA_TRY(cudaMemcpy(c->d_pda_counts, h.data(), nc * sizeof(double), cudaMemcpyHostToDevice));
    counts = c->d_pda_counts;
  } else if (!c->d_nvis) {
    return fail(HYP_ERR_STATE, "n_photons array is not allocated");
  }
  return solve_pda_device(c, counts, n_pda_cells);
}

}  // extern "C"

// =============================================================================================
// final (imaging) and raytracing iterations: host side
// =============================================================================================
namespace {

int n_orig_of(const hyp_image_conf &g, int n_sources, int n_dust) {
  switch (g.track_origin) {
    case HYP_TRACK_SCATTERINGS: return 4 + 2 * g.track_n_scat;
    case HYP_TRACK_DETAILED: return 2 * (n_sources + n_dust);
    case HYP_TRACK_BASIC: return 4;
    default: return 1;
  }
}

// Allocate the image cubes of every group in ONE buffer (so that one collective reduces them) and
// build the device descriptors.  Cubes start at zero and accumulate over calls, as in the reference.
int ensure_images(hyp_ctx *c) {
  if (c->images_ready) return HYP_OK;
  const int ns = (int)c->sources.size(), nd = (int)c->dust.size();
  size_t off = 0;
  int n_views = 0;
  for (auto &g : c->groups) {
    const hyp_image_conf &k = g.conf;
    g.n_orig = n_orig_of(k, ns, nd);
    g.n_stokes = k.compute_stokes ? 4 : 1;
    const double c_cgs = 2.99792458e10, micron = (double)1.e-4f;  // single-precision literal in image_type.f90:262-263
    g.nu_min = c_cgs / (k.wav_max * micron);
    g.nu_max = c_cgs / (k.wav_min * micron);
    const size_t outer = (size_t)k.n_view * g.n_orig * g.n_stokes;
    g.n_sed = k.compute_sed ? (size_t)k.n_wav * k.n_ap * outer : 0;
    g.n_img = k.compute_image ? (size_t)k.n_wav * k.n_x * k.n_y * outer : 0;
    const int copies = k.uncertainties ? 3 : 1;
    g.o_sed = off;
    off += g.n_sed * copies;
    g.o_img = off;
    off += g.n_img * copies;
    if (!k.binned) n_views += k.n_view;
  }
  c->imgbuf_n = off;
  CUDA_TRY(cudaMalloc(&c->d_imgbuf, (off + SC_COUNT) * sizeof(double)));
  CUDA_TRY(cudaMemset(c->d_imgbuf, 0, (off + SC_COUNT) * sizeof(double)));
  CUDA_TRY(cudaMalloc(&c->d_eabs, MAX_DUST * sizeof(double)));
  std::vector<ViewDev> views;
  c->h_images.assign(c->groups.size(), ImageDev());
  const double deg = 3.14159265358979323846264338327950288419 / 180.0;
  for (size_t ig = 0; ig < c->groups.size(); ++ig) {
    const HostImage &g = c->groups[ig];
    const hyp_image_conf &k = g.conf;
    ImageDev &d = c->h_images[ig];
    memset(&d, 0, sizeof d);
    d.n_view = k.n_view; d.n_nu = k.n_wav; d.n_x = k.n_x; d.n_y = k.n_y; d.n_ap = k.n_ap;
    d.n_orig = g.n_orig; d.n_stokes = g.n_stokes;
    d.compute_image = k.compute_image; d.compute_sed = k.compute_sed;
    d.track_origin = k.track_origin; d.track_n_scat = k.track_n_scat;
    d.uncertainties = k.uncertainties; d.ignore_optical_depth = k.ignore_optical_depth;
    d.inside_observer = k.inside_observer;
    d.inu_min = k.inu_min > 0 ? k.inu_min : 0;
    d.use_filters = k.use_filters;
    if (k.use_filters) {
      int32_t *d_off = nullptr;
      double *d_nu = nullptr, *d_tr = nullptr;
      CUDA_TRY(cudaMalloc(&d_off, g.filt_off.size() * sizeof(int32_t)));
      CUDA_TRY(cudaMalloc(&d_nu, g.filt_nu.size() * sizeof(double)));
      CUDA_TRY(cudaMalloc(&d_tr, g.filt_tr.size() * sizeof(double)));
      CUDA_TRY(cudaMemcpy(d_off, g.filt_off.data(), g.filt_off.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
      CUDA_TRY(cudaMemcpy(d_nu, g.filt_nu.data(), g.filt_nu.size() * sizeof(double), cudaMemcpyHostToDevice));
      CUDA_TRY(cudaMemcpy(d_tr, g.filt_tr.data(), g.filt_tr.size() * sizeof(double), cudaMemcpyHostToDevice));
      c->filter_tables.push_back(d_off);
      c->filter_tables.push_back(d_nu);
      c->filter_tables.push_back(d_tr);
      d.filt_off = d_off; d.filt_nu = d_nu; d.filt_tr = d_tr;
    }
    d.n_sources = ns; d.n_dust = nd;
    d.x_min = k.x_min; d.x_max = k.x_max; d.y_min = k.y_min; d.y_max = k.y_max;
    d.ap_min = k.ap_min; d.ap_max = k.ap_max;
    if (k.compute_sed) {
      d.log10_ap_min = std::log10(k.ap_min);
      d.log10_ap_max = std::log10(k.ap_max);
    }
    d.log10_nu_min = std::log10(g.nu_min);
    d.log10_nu_max = std::log10(g.nu_max);
    d.d_min = k.d_min; d.d_max = k.d_max;
    d.rpx = k.peeloff_x; d.rpy = k.peeloff_y; d.rpz = k.peeloff_z;
    double *b = c->d_imgbuf;
    if (g.n_sed) {
      d.sed = b + g.o_sed;
      if (k.uncertainties) { d.sed2 = d.sed + g.n_sed; d.sedn = d.sed2 + g.n_sed; }
    }
    if (g.n_img) {
      d.img = b + g.o_img;
      if (k.uncertainties) { d.img2 = d.img + g.n_img; d.imgn = d.img2 + g.n_img; }
    }
    // a binned group has no viewing directions: packets choose their own (images_binned.f90:57-77)
    for (int iv = 0; iv < (k.binned ? 0 : k.n_view); ++iv) {
      ViewDev v;
      v.group = (int)ig;
      v.view = iv;
      // angle3d_deg (type_angle3d.f90:127-147)
      v.a.cost = std::cos(g.theta[iv] * deg);
      v.a.sint = std::sin(g.theta[iv] * deg);
      v.a.cosp = std::cos(g.phi[iv] * deg);
      v.a.sinp = std::sin(g.phi[iv] * deg);
      views.push_back(v);
    }
  }
  c->n_views = n_views;
  if (!c->groups.empty()) {
    CUDA_TRY(cudaMalloc(&c->d_images, c->h_images.size() * sizeof(ImageDev)));
    CUDA_TRY(cudaMemcpy(c->d_images, c->h_images.data(), c->h_images.size() * sizeof(ImageDev), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&c->d_views, (views.size() + 1) * sizeof(ViewDev)));
    CUDA_TRY(cudaMemcpy(c->d_views, views.data(), views.size() * sizeof(ViewDev), cudaMemcpyHostToDevice));
  }
  c->images_ready = true;
  return HYP_OK;
}

// Spectra of every source and dust type on every group's frequency grid (the lazily filled caches of
// images_peeled.f90:423-530, computed up front here).
int ensure_ray_tables(hyp_ctx *c) {
  if (c->ray_ready) return HYP_OK;
  const int ns = (int)c->sources.size(), nd = (int)c->dust.size();
  auto upload = [&](const std::vector<double> &h, const double **dst) -> int {
    double *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, std::max<size_t>(h.size(), 1) * sizeof(double)));
    CUDA_TRY(cudaMemcpy(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    c->ray_tables.push_back(d);
    *dst = d;
    return HYP_OK;
  };
  for (size_t ig = 0; ig < c->groups.size(); ++ig) {
    ImageDev &d = c->h_images[ig];
    const int n_nu = d.n_nu;
    const double l0 = d.log10_nu_min, l1 = d.log10_nu_max;
    std::vector<double> spec((size_t)ns * n_nu), chi((size_t)nd * n_nu);
    std::vector<double> bnu, bfnu;
    if (d.inu_min > 0) {
      // exact frequencies: get_spectrum_interp (source_type.f90:1098-1116), get_j_nu_interp / get_chi_nu_interp
      // (dust_type_4elem.f90:708-720,780-791)
      const double *nu = c->frequencies.data() + (d.inu_min - 1);
      for (int is = 0; is < ns; ++is) {
        const int sp = c->source_spectrum[is];
        if (c->sources[is].spectrum_type == HYP_SPECTRUM_LTE) continue;
        for (int i = 0; i < n_nu; ++i)
          spec[(size_t)is * n_nu + i] = sp >= 0 ? pdf_loglog_at(c->spectra[sp].nu.data(), c->spectra[sp].fnu.data(),
                                                                 (int)c->spectra[sp].nu.size(), nu[i])
                                                : normalized_B_nu(nu[i], c->sources[is].temperature);
      }
      int rc = upload(spec, &d.src_spec);
      if (rc) return rc;
      for (int id = 0; id < nd; ++id) {
        const HostDust &D = c->dust[id];
        for (int i = 0; i < n_nu; ++i)
          chi[(size_t)id * n_nu + i] = interp_loglog_fill0(D.nu.data(), D.chi.data(), (int)D.nu.size(), nu[i]);
        const int ne = (int)D.emiss_nu.size();
        std::vector<double> logj((size_t)D.n_jnu * n_nu), col(ne);
        for (int st = 0; st < D.n_jnu; ++st) {
          for (int k = 0; k < ne; ++k) col[k] = D.emiss_jnu[(size_t)k * D.n_jnu + st];
          const double tot = detail::integral_loglog_all(D.emiss_nu.data(), col.data(), ne);
          for (int i = 0; i < n_nu; ++i)
            logj[(size_t)st * n_nu + i] = std::log10(interp_loglog_fill0(D.emiss_nu.data(), col.data(), ne, nu[i]) / tot);
        }
        rc = upload(logj, &d.dust_logj[id]);
        if (rc) return rc;
      }
      rc = upload(chi, &d.dust_chi);
      if (rc) return rc;
      continue;
    }
    for (int is = 0; is < ns; ++is) {
      const int sp = c->source_spectrum[is];
      if (sp >= 0) {
        const HostSpectrum &S = c->spectra[sp];
        binned_fraction(S.nu.data(), S.fnu.data(), (int)S.nu.size(), l0, l1, n_nu, &spec[(size_t)is * n_nu]);
      } else if (c->sources[is].spectrum_type == HYP_SPECTRUM_LTE) {
        // emiss_type 3: peeled with the dust emissivity of the emitting cell (images_peeled.f90:223-224)
      } else {
        blackbody_table(c->sources[is].temperature, bnu, bfnu);
        binned_fraction(bnu.data(), bfnu.data(), (int)bnu.size(), l0, l1, n_nu, &spec[(size_t)is * n_nu]);
      }
    }
    int rc = upload(spec, &d.src_spec);
    if (rc) return rc;
    for (int id = 0; id < nd; ++id) {
      const HostDust &D = c->dust[id];
      binned_chi(D.nu.data(), D.chi.data(), (int)D.nu.size(), l0, l1, n_nu, &chi[(size_t)id * n_nu]);
      const int ne = (int)D.emiss_nu.size();
      std::vector<double> logj((size_t)D.n_jnu * n_nu), col(ne);
      for (int s = 0; s < D.n_jnu; ++s) {
        for (int k = 0; k < ne; ++k) col[k] = D.emiss_jnu[(size_t)k * D.n_jnu + s];
        double *o = &logj[(size_t)s * n_nu];
        binned_fraction(D.emiss_nu.data(), col.data(), ne, l0, l1, n_nu, o);
        for (int i = 0; i < n_nu; ++i) o[i] = std::log10(o[i]);
      }
      rc = upload(logj, &d.dust_logj[id]);
      if (rc) return rc;
    }
    rc = upload(chi, &d.dust_chi);
    if (rc) return rc;
  }
  if (!c->groups.empty())
    CUDA_TRY(cudaMemcpy(c->d_images, c->h_images.data(), c->h_images.size() * sizeof(ImageDev), cudaMemcpyHostToDevice));
  c->ray_ready = true;
  return HYP_OK;
}

size_t job_bytes(int nd) {
  switch (nd) {
    case 1: return sizeof(PeelJob<1>);
    case 2: return sizeof(PeelJob<2>);
    case 3: return sizeof(PeelJob<3>);
    default: return sizeof(PeelJob<4>);
  }
}

int ensure_jobs(hyp_ctx *c, uint32_t cap) {
  if (c->job_cap >= cap) return HYP_OK;
  free_dev(c->d_jobs);
  free_dev(c->d_njobs);
  CUDA_TRY(cudaMalloc(&c->d_jobs, (size_t)cap * job_bytes(c->M.n_dust)));
  // [0] number of queued jobs, [2..3] the peel kernel's 64-bit work cursor
  CUDA_TRY(cudaMalloc(&c->d_njobs, 4 * sizeof(uint32_t)));
  CUDA_TRY(cudaMemset(c->d_njobs, 0, 4 * sizeof(uint32_t)));
  c->job_cap = cap;
  return HYP_OK;
}

struct WallSmem {
  size_t bytes;
  int on;
};
WallSmem wall_smem(const hyp_ctx *c) {
  size_t b = (size_t)(c->n1 + c->n2 + c->n3 + 3) * sizeof(double);
  if (b > 40 * 1024) return {0, 0};
  return {b, 1};
}

template <int ND, bool POLY>
int launch_peel(hyp_ctx *c, const ModelDev &M, uint32_t n_jobs_max) {
  if (c->n_views == 0 || n_jobs_max == 0) return HYP_OK;
  WallSmem ws = wall_smem(c);
  auto k = peel_kernel<ND, POLY, GEO_CAR>;
  if (c->grid_type == GEO_SPH) {
    k = peel_kernel<ND, POLY, GEO_SPH>;
    ws = WallSmem{0, 0};
  } else if (c->grid_type == GEO_OCT) {
    k = peel_kernel<ND, POLY, GEO_OCT>;
    ws = WallSmem{0, 0};
  } else if (c->grid_type == GEO_AMR) {
    k = peel_kernel<ND, POLY, GEO_AMR>;
    ws = WallSmem{0, 0};
  } else if (c->grid_type == GEO_VOR) {
    k = peel_kernel<ND, POLY, GEO_VOR>;
    ws = WallSmem{0, 0};
  }
  CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ws.bytes));
  ImagingDev I{c->d_images, c->d_views, (int)c->groups.size(), c->n_views, c->d_src_columns};
  const int64_t work = (int64_t)n_jobs_max * c->n_views;
  int per_sm = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, PEEL_THREADS, ws.bytes));
  if (per_sm < 1) per_sm = 1;
  const int blocks = (int)std::min<int64_t>((work + PEEL_THREADS - 1) / PEEL_THREADS, (int64_t)c->sm_count * per_sm);
  CUDA_TRY(cudaMemsetAsync(c->d_njobs + 2, 0, 2 * sizeof(uint32_t), c->stream));
  k<<<blocks, PEEL_THREADS, ws.bytes, c->stream>>>(M, I, (const PeelJob<ND> *)c->d_jobs, c->d_njobs,
                                                  (unsigned long long *)(c->d_njobs + 2), ws.on);
  CUDA_TRY(cudaGetLastError());
  c->launches_acc += 1;
  return HYP_OK;
}

// (re)compute the point-source column cache; the density may have changed since the last call
template <int ND>
int update_source_columns(hyp_ctx *c) {
  if (c->n_views == 0) return HYP_OK;
  const size_t n = (size_t)c->sources.size() * c->n_views * (ND + 1);
  if (!c->d_src_columns) CUDA_TRY(cudaMalloc(&c->d_src_columns, std::max<size_t>(n, 1) * sizeof(double)));
  ImagingDev I{c->d_images, c->d_views, (int)c->groups.size(), c->n_views, c->d_src_columns};
  auto k = source_columns_kernel<ND, GEO_CAR>;
  if (c->grid_type == GEO_SPH) k = source_columns_kernel<ND, GEO_SPH>;
  else if (c->grid_type == GEO_OCT) k = source_columns_kernel<ND, GEO_OCT>;
  else if (c->grid_type == GEO_AMR) k = source_columns_kernel<ND, GEO_AMR>;
  else if (c->grid_type == GEO_VOR) k = source_columns_kernel<ND, GEO_VOR>;
  k<<<1, 128, 0, c->stream>>>(c->M, I, c->d_src_columns);
  CUDA_TRY(cudaGetLastError());
  c->launches_acc += 1;
  return HYP_OK;
}

int update_source_columns_nd(hyp_ctx *c) {
  switch (c->M.n_dust) {
    case 1: return update_source_columns<1>(c);
    case 2: return update_source_columns<2>(c);
    case 3: return update_source_columns<3>(c);
    default: return update_source_columns<4>(c);
  }
}

ModelDev imaging_model(hyp_ctx *c) {
  ModelDev M = c->M;
  M.scalars = c->d_imgbuf + c->imgbuf_n;
  return M;
}

// rounds of the packet pool for the imaging iteration (propagate, iter_final.f90:147-273)
template <int ND>
int run_final_rounds(hyp_ctx *c, int64_t first_id, int64_t n_photons, int scattering_only, const ModelDev *Mo = nullptr,
                     const uint32_t iteration = ITER_FINAL, const int thermal = 0) {
  const uint32_t cap = (uint32_t)std::min<int64_t>(pool_target(), n_photons);
  int rc = ensure_pool(c, cap);
  if (rc) return rc;
  // peel-off queue of a round: one job per emission and per interaction, plus two per random-walk step; a walk
  // that finds the queue nearly full is interrupted and resumed next round (interact_final_kernel), so the size
  // only sets how many steps fit in a round.  HYPERION_B200_JOBS: capacity in units of the packet pool (>= 4).
  const int job_mult = c->M.use_mrw ? std::max(4, getenv("HYPERION_B200_JOBS") ? atoi(getenv("HYPERION_B200_JOBS")) : 8) : 2;
  rc = ensure_jobs(c, (uint32_t)job_mult * c->pool_cap);
  if (rc) return rc;
  Pool &P = c->pool;
  const WallSmem ws = wall_smem(c);
  auto flight = flight_final_kernel<ND, FLIGHT_LOOKAHEAD>;
  CUDA_TRY(cudaFuncSetAttribute(flight, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ws.bytes));
  int per_sm = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, flight, FLIGHT_THREADS, ws.bytes));
  if (per_sm < 1) per_sm = 1;
  const int flight_blocks_max = per_sm * c->sm_count;
  const int service_blocks_max = c->sm_count * 8;
  cudaStream_t st = c->stream;
  const ModelDev M = Mo ? *Mo : imaging_model(c);
  FinalArgs F;
  F.thermal = thermal;
  F.jobs = c->d_jobs;
  F.n_jobs = c->d_njobs;
  F.job_capacity = (uint32_t)job_mult * c->pool_cap;
  F.job_margin = 2u * std::min<uint32_t>((uint32_t)(c->sm_count * 8 * SERVICE_THREADS), c->pool_cap);
  F.scattering_only = scattering_only;
  F.forced = c->conf.forced_first_interaction;
  F.algorithm = c->conf.forced_first_interaction_algorithm;
  F.baes16_xi = c->conf.baes16_xi;
  F.make_peeled = c->n_views > 0;
  F.binned = nullptr;
  F.n_theta = F.n_phi = 0;
  for (size_t ig = 0; ig < c->groups.size(); ++ig)
    if (c->groups[ig].conf.binned) {
      F.binned = c->d_images + ig;
      F.n_theta = c->groups[ig].conf.n_theta;
      F.n_phi = c->groups[ig].conf.n_phi;
    }

  pool_init_kernel<<<c->sm_count, 256, 0, st>>>(P, cap);
  c->launches_acc += 1;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemsetAsync(c->d_njobs, 0, sizeof(uint32_t), st));
  CUDA_TRY(cudaEventRecord(c->ev0, st));
  int cur = 0;
  uint32_t n_emit = cap, n_flight_prev = 0;
  unsigned long long claimed = 0;
  int64_t windows_ready = 0;
  for (int64_t round = 0;; ++round) {
    uint32_t *nF = P.counts + C_NF0 + cur, *nF_next = P.counts + C_NF0 + (1 - cur);
    while (windows_ready * (int64_t)c->sort_window < n_photons &&
           (int64_t)claimed + 2 * (int64_t)cap > windows_ready * (int64_t)c->sort_window) {
      rc = prepare_window(c, first_id, n_photons, iteration, windows_ready, &M, thermal ? 0 : -1);
      if (rc) return rc;
      ++windows_ready;
    }
    int64_t n_new = 0;
    if (claimed < (unsigned long long)n_photons && n_emit > 0) {
      n_new = std::min<int64_t>(n_emit, n_photons - (int64_t)claimed);
      int blocks = (int)std::min<int64_t>(((int64_t)n_emit + SERVICE_THREADS - 1) / SERVICE_THREADS, service_blocks_max);
      emit_final_kernel<ND><<<blocks, SERVICE_THREADS, 0, st>>>(M, P, F, (unsigned long long)first_id,
                                                                (unsigned long long)n_photons, iteration);
      CUDA_TRY(cudaGetLastError());
      c->launches_acc += 1;
    }
    CUDA_TRY(cudaMemsetAsync(P.counts + C_NI, 0, 4 * sizeof(uint32_t), st));  // C_NI, C_NE, C_CURSOR, C_CURSOR_B
    CUDA_TRY(cudaEventRecord(c->evA, st));
    if (c->grid_type != GEO_CAR || c->M.any_sphere) {
      const int sph_blocks_max = c->sm_count * 12;
      auto geo_flight = c->grid_type == GEO_OCT   ? flight_geo_kernel<GEO_OCT, ND, false, true>
                        : c->grid_type == GEO_AMR ? flight_geo_kernel<GEO_AMR, ND, false, true>
                        : c->grid_type == GEO_VOR ? flight_geo_kernel<GEO_VOR, ND, false, true>
                        : c->grid_type == GEO_CAR ? flight_geo_kernel<GEO_CAR, ND, false, true>
                                                  : flight_geo_kernel<GEO_SPH, ND, false, true>;
      if (n_new > 0) {
        int blocks = (int)std::min<int64_t>((n_new + SPH_FLIGHT_THREADS - 1) / SPH_FLIGHT_THREADS, sph_blocks_max);
        geo_flight<<<blocks, SPH_FLIGHT_THREADS, 0, st>>>(M, P, F, P.q_beam, P.counts + C_NB,
                                                                                 P.counts + C_CURSOR_B, iteration);
        c->launches_acc += 1;
      }
      if (n_flight_prev > 0) {
        int blocks = (int)std::min<int64_t>(((int64_t)n_flight_prev + SPH_FLIGHT_THREADS - 1) / SPH_FLIGHT_THREADS, sph_blocks_max);
        geo_flight<<<blocks, SPH_FLIGHT_THREADS, 0, st>>>(M, P, F, P.q_flight[cur], nF,
                                                                                 P.counts + C_CURSOR, iteration);
        c->launches_acc += 1;
      }
      CUDA_TRY(cudaGetLastError());
    } else {
    if (n_new > 0) {
      int blocks = (int)std::min<int64_t>((n_new + FLIGHT_THREADS - 1) / FLIGHT_THREADS, flight_blocks_max);
      flight<<<blocks, FLIGHT_THREADS, ws.bytes, st>>>(M, P, F, P.q_beam, P.counts + C_NB, P.counts + C_CURSOR_B, ws.on, iteration);
      CUDA_TRY(cudaGetLastError());
      c->launches_acc += 1;
    }
    if (n_flight_prev > 0) {
      int blocks = (int)std::min<int64_t>(((int64_t)n_flight_prev + FLIGHT_THREADS - 1) / FLIGHT_THREADS, flight_blocks_max);
      flight<<<blocks, FLIGHT_THREADS, ws.bytes, st>>>(M, P, F, P.q_flight[cur], nF, P.counts + C_CURSOR, ws.on, iteration);
      CUDA_TRY(cudaGetLastError());
      c->launches_acc += 1;
    }
    }
    CUDA_TRY(cudaEventRecord(c->evB, st));
    CUDA_TRY(cudaMemsetAsync(nF_next, 0, sizeof(uint32_t), st));
    CUDA_TRY(cudaMemsetAsync(P.counts + C_NB, 0, sizeof(uint32_t), st));
    interact_final_kernel<ND><<<service_blocks_max, SERVICE_THREADS, 0, st>>>(M, P, F, P.q_flight[1 - cur], nF_next, iteration);
    CUDA_TRY(cudaGetLastError());
    c->launches_acc += 1;
    // peel-offs of this round's emissions and interactions
    rc = launch_peel<ND, false>(c, M, c->M.use_mrw ? c->job_cap : (uint32_t)std::min<int64_t>((int64_t)c->job_cap, n_new + (int64_t)cap));
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(c->d_njobs, 0, sizeof(uint32_t), st));
    CUDA_TRY(cudaMemcpyAsync(c->h_counts, P.counts, C_COUNT * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(c->h_counts + C_COUNT, P.next_photon, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->evA, c->evB) == cudaSuccess) c->flight_ms_acc += ms;
    c->rounds_acc += 1;
    cur = 1 - cur;
    memcpy(&claimed, c->h_counts + C_COUNT, sizeof claimed);
    n_emit = c->h_counts[C_NE];
    n_flight_prev = c->h_counts[C_NF0 + cur];
    const bool ids_left = claimed < (unsigned long long)n_photons;
    if (n_flight_prev == 0 && (!ids_left || n_emit == 0)) break;
  }
  CUDA_TRY(cudaEventRecord(c->ev1, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->kernel_ms_acc += ms;
  return HYP_OK;
}

template <int ND>
int run_raytracing(hyp_ctx *c, int64_t first_source_id, int64_t n_sources, int64_t n_total_sources,
                   int64_t first_dust_id, int64_t n_thermal, int64_t n_total_dust) {
  const ModelDev M = imaging_model(c);
  cudaStream_t st = c->stream;
  // do_raytracing starts with precompute_jnu_var (iter_raytracing.f90:60)
  lucy_begin_kernel<<<grid_blocks(c), 256, 0, st>>>(c->M, 0);
  CUDA_TRY(cudaMemsetAsync(c->d_eabs, 0, MAX_DUST * sizeof(double), st));
  energy_abs_tot_kernel<<<grid_blocks(c), 256, 0, st>>>(c->M, c->d_eabs);
  CUDA_TRY(cudaGetLastError());
  c->launches_acc += 2;
  const uint32_t chunk = 1u << 20;
  int rc = ensure_jobs(c, chunk);
  if (rc) return rc;
  const double source_weight = n_total_sources > 0 ? c->energy_total / (double)n_total_sources : 0.0;
  const double dust_weight = n_total_dust > 0 ? (double)ND / (double)n_total_dust : 0.0;
  int64_t done_s = 0, done_d = 0;
  while (done_s < n_sources || done_d < n_thermal) {
    const uint32_t ns = (uint32_t)std::min<int64_t>(chunk, n_sources - done_s);
    const uint32_t nt = (uint32_t)std::min<int64_t>(chunk - ns, n_thermal - done_d);
    CUDA_TRY(cudaMemsetAsync(c->d_njobs, 0, sizeof(uint32_t), st));
    const int blocks = (int)std::min<int64_t>(((int64_t)ns + nt + 255) / 256, (int64_t)c->sm_count * 8);
    raytrace_emit_kernel<ND><<<blocks, 256, 0, st>>>(M, (PeelJob<ND> *)c->d_jobs, c->d_njobs,
                                                     (unsigned long long)(first_source_id + done_s), ns, source_weight,
                                                     (unsigned long long)(first_dust_id + done_d), nt, dust_weight,
                                                     c->d_eabs);
    CUDA_TRY(cudaGetLastError());
    c->launches_acc += 1;
    rc = launch_peel<ND, true>(c, M, ns + nt);
    if (rc) return rc;
    done_s += ns;
    done_d += nt;
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  return HYP_OK;
}

int read_image_scalars(hyp_ctx *c, double *sc) {
  CUDA_TRY(cudaMemcpyAsync(sc, c->d_imgbuf + c->imgbuf_n, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return HYP_OK;
}

void fill_image_stats(hyp_ctx *c, const double *sc, hyp_iter_stats *st) {
  if (!st) return;
  memset(st, 0, sizeof *st);
  st->energy_emitted = sc[SC_ENERGY];
  st->n_photons = (int64_t)sc[SC_PHOTONS];
  st->killed_geo = (int64_t)sc[SC_KILLED_GEO];
  st->killed_int = (int64_t)sc[SC_KILLED_INT];
  st->n_crossings = (int64_t)sc[SC_CROSS];
  st->n_absorptions = (int64_t)sc[SC_ABS];
  st->n_scatterings = (int64_t)sc[SC_SCAT];
  st->n_escaped = (int64_t)sc[SC_ESC];
  st->n_peel_crossings = (int64_t)sc[SC_PEEL_CROSS];
  st->n_peeloffs = (int64_t)sc[SC_PEELOFFS];
  st->n_peel_cached = (int64_t)sc[SC_PEEL_CACHED];
  st->kernel_ms = c->kernel_ms_acc;
  st->flight_ms = c->flight_ms_acc;
  st->n_rounds = c->rounds_acc;
  st->n_launches = c->launches_acc;
}

}  // namespace

extern "C" {

int hyp_add_peeled_group(hyp_ctx *c, const hyp_image_conf *g) {
  if (!c || !g) return fail(HYP_ERR_INVALID, "NULL argument");
  if (c->images_ready) return fail(HYP_ERR_STATE, "image groups are frozen once an imaging iteration has started");
  if (g->binned) {
    // setup_final_iteration (setup_rt.f90:318-331), binned_images_setup (images_binned.f90:41-55)
    for (auto &o : c->groups)
      if (o.conf.binned) return fail(HYP_ERR_INVALID, "can't have more than one binned image group");
    if (c->conf.forced_first_interaction)
      return fail(HYP_ERR_INVALID, "can't use binned images with forced first interaction");
    if (g->n_theta < 1 || g->n_phi < 1 || g->n_view != g->n_theta * g->n_phi)
      return fail(HYP_ERR_INVALID, "binned images: n_view should be n_theta * n_phi");
  } else if (!(g->n_view > 0) || !g->theta || !g->phi) {
    return fail(HYP_ERR_INVALID, "n_view should be a positive integer");
  }
  if (g->n_wav < 1) return fail(HYP_ERR_INVALID, "n_nu should be >= 1");
  if (g->inu_min > 0) {
    // image_setup (image_type.f90:243-258)
    if (g->use_filters) return fail(HYP_ERR_INVALID, "cannot use filters in monochromatic mode");
    const int nf = (int)c->frequencies.size();
    if (nf == 0) return fail(HYP_ERR_STATE, "hyp_set_monochromatic has not been called");
    if (g->inu_min < 1 || g->inu_min > nf) return fail(HYP_ERR_INVALID, "inu_min value is out of range");
    if (g->inu_max < 1 || g->inu_max > nf) return fail(HYP_ERR_INVALID, "inu_max value is out of range");
    if (g->n_wav != g->inu_max - g->inu_min + 1) return fail(HYP_ERR_INVALID, "n_nu should match length of frequencies array");
  }
  if (g->io_bytes != 4 && g->io_bytes != 8) return fail(HYP_ERR_INVALID, "unexpected value of io_bytes (should be 4 or 8)");
  if (g->track_origin < HYP_TRACK_NO || g->track_origin > HYP_TRACK_SCATTERINGS)
    return fail(HYP_ERR_INVALID, "unknown track_origin flag");
  if (g->track_origin == HYP_TRACK_SCATTERINGS && (g->track_n_scat < 0 || g->track_n_scat > 1000))
    return fail(HYP_ERR_INVALID, "track_n_scat should be between 0 and 1000 (the packet tag counts scatterings up to 1023)");
  if (g->compute_image && (g->n_x < 1 || g->n_y < 1)) return fail(HYP_ERR_INVALID, "image needs at least one pixel");
  if (g->compute_sed && g->n_ap < 1) return fail(HYP_ERR_INVALID, "SED needs at least one aperture");
  HostImage h;
  h.conf = *g;
  if (!g->binned) {
    h.theta.assign(g->theta, g->theta + g->n_view);
    h.phi.assign(g->phi, g->phi + g->n_view);
  }
  if (g->use_filters) {
    if (!g->filt_n || !g->filt_nu || !g->filt_tr) return fail(HYP_ERR_INVALID, "filter tables are missing");
    h.filt_off.push_back(0);
    for (int i = 0; i < g->n_wav; ++i) {
      if (g->filt_n[i] < 2) return fail(HYP_ERR_INVALID, "a filter needs at least two points");
      h.filt_off.push_back(h.filt_off.back() + g->filt_n[i]);
    }
    h.filt_nu.assign(g->filt_nu, g->filt_nu + h.filt_off.back());
    h.filt_tr.assign(g->filt_tr, g->filt_tr + h.filt_off.back());
    h.conf.wav_min = h.conf.wav_max = 1.0;   // unused with filters
  }
  if (g->inu_min > 0) h.conf.wav_min = h.conf.wav_max = 1.0;   // unused at exact frequencies
  h.conf.filt_n = nullptr;
  h.conf.filt_nu = h.conf.filt_tr = h.conf.filt_nu0 = nullptr;
  h.conf.theta = h.conf.phi = nullptr;
  c->groups.push_back(std::move(h));
  return HYP_OK;
}

int hyp_final_begin(hyp_ctx *c) {
  if (!c || !c->finalized) return fail(HYP_ERR_STATE, "hyp_finalize_setup has not been called");
  CUDA_TRY(cudaSetDevice(c->device));
  int rc = ensure_images(c);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(c->ev2, c->stream));
  // precompute_jnu_var (iter_final.f90:99); the deposit grid is left alone
  lucy_begin_kernel<<<grid_blocks(c), 256, 0, c->stream>>>(c->M, 0);
  CUDA_TRY(cudaGetLastError());
  c->launches_acc = 1;
  if (c->M.use_mrw) {
    mrw_prepare_kernel<<<grid_blocks(c), 256, 0, c->stream>>>(c->M);
    CUDA_TRY(cudaGetLastError());
    c->launches_acc += 1;
  }
  CUDA_TRY(cudaMemsetAsync(c->d_imgbuf + c->imgbuf_n, 0, SC_COUNT * sizeof(double), c->stream));
  CUDA_TRY(cudaMemsetAsync(c->d_error, 0, sizeof(int32_t), c->stream));
  rc = update_source_columns_nd(c);
  if (rc) return rc;
  c->kernel_ms_acc = 0.f;
  c->flight_ms_acc = 0.f;
  c->rounds_acc = 0;
  return HYP_OK;
}

int hyp_final_photons(hyp_ctx *c, int64_t first_id, int64_t n_photons, int32_t peeloff_scattering_only) {
  if (!c || !c->finalized) return fail(HYP_ERR_STATE, "hyp_finalize_setup has not been called");
  if (!c->images_ready) return fail(HYP_ERR_STATE, "hyp_final_begin has not been called");
  if (n_photons < 0 || first_id < 0) return fail(HYP_ERR_INVALID, "negative photon count");
  if (n_photons == 0) return HYP_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  switch (c->M.n_dust) {
    case 1: return run_final_rounds<1>(c, first_id, n_photons, peeloff_scattering_only);
    case 2: return run_final_rounds<2>(c, first_id, n_photons, peeloff_scattering_only);
    case 3: return run_final_rounds<3>(c, first_id, n_photons, peeloff_scattering_only);
    case 4: return run_final_rounds<4>(c, first_id, n_photons, peeloff_scattering_only);
    default: return fail(HYP_ERR_INVALID, "unsupported number of dust types");
  }
}

int hyp_set_monochromatic(hyp_ctx *c, int32_t n_nu, const double *frequencies, double energy_threshold) {
  if (!c) return fail(HYP_ERR_INVALID, "NULL argument");
  if (n_nu < 1 || !frequencies) return fail(HYP_ERR_INVALID, "monochromatic mode needs at least one frequency");
  if (!c->groups.empty()) return fail(HYP_ERR_STATE, "hyp_set_monochromatic has to be called before hyp_add_peeled_group");
  c->frequencies.assign(frequencies, frequencies + n_nu);
  c->mono_threshold = energy_threshold;
  return HYP_OK;
}

}  // extern "C"

namespace {

template <typename T>
int upload_vec(const std::vector<T> &h, T *&d) {
  free_dev(d);
  CUDA_TRY(cudaMalloc(&d, std::max<size_t>(h.size(), 1) * sizeof(T)));
  CUDA_TRY(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return HYP_OK;
}

template <int ND>
int run_final_mono(hyp_ctx *c, int32_t inu, int64_t first_source_id, int64_t n_sources, int64_t n_total_sources,
                   int64_t first_dust_id, int64_t n_thermal, int64_t n_total_thermal, int32_t scattering_only) {
  const double nu = c->frequencies[inu - 1];
  const int ns = (int)c->sources.size();
  cudaStream_t st = c->stream;
  ModelDev M = imaging_model(c);
  M.mono_nu = nu;
  M.mono_inu = inu;
  M.mono_threshold = c->mono_threshold;
  M.use_mrw = 0;   // propagate of iter_final_mono.f90:231-341 has no random walk
  int rc;
  if (n_sources > 0 && ns > 0) {
    // source_emit with nu given (source_type.f90:436-474) x energy_total (source.f90:161) / n_photons
    // (iter_final_mono.f90:118)
    const double wgt = c->energy_total / (double)n_total_sources;
    std::vector<double> w(ns, 0.0), ws(std::max<size_t>(c->spots.size(), 1), 0.0);
    for (int is = 0; is < ns; ++is) {
      const int sp = c->source_spectrum[is];
      if (c->sources[is].spectrum_type == HYP_SPECTRUM_LTE) continue;
      const double prob = sp >= 0 ? pdf_loglog_at(c->spectra[sp].nu.data(), c->spectra[sp].fnu.data(),
                                                  (int)c->spectra[sp].nu.size(), nu)
                                  : normalized_B_nu(nu, c->sources[is].temperature);
      w[is] = prob * wgt;
    }
    for (size_t k = 0; k < c->spots.size(); ++k) {
      const SpotDev &q = c->spots[k];
      if (q.freq_type != HYP_SPECTRUM_BLACKBODY && q.spectrum < 0) continue;   // the star's own entry after a source's spots
      const double prob = q.freq_type == HYP_SPECTRUM_BLACKBODY
                              ? normalized_B_nu(nu, q.temperature)
                              : pdf_loglog_at(c->spectra[q.spectrum].nu.data(), c->spectra[q.spectrum].fnu.data(),
                                              (int)c->spectra[q.spectrum].nu.size(), nu);
      ws[k] = prob * wgt;
    }
    rc = upload_vec(w, c->d_mono_src_w);
    if (rc) return rc;
    rc = upload_vec(ws, c->d_mono_spot_w);
    if (rc) return rc;
    M.mono_src_w = c->d_mono_src_w;
    M.mono_spot_w = c->d_mono_spot_w;
    M.mono_lte_w = wgt;
  }
  // dust_sample_emit_probability's two factors: interpolate_pdf of every emissivity state at nu, as log10
  for (int id = 0; id < ND; ++id) {
    const HostDust &D = c->dust[id];
    const int ne = (int)D.emiss_nu.size();
    std::vector<double> lp(D.n_jnu), col(ne);
    for (int s = 0; s < D.n_jnu; ++s) {
      for (int k = 0; k < ne; ++k) col[k] = D.emiss_jnu[(size_t)k * D.n_jnu + s];
      lp[s] = std::log10(pdf_loglog_at(D.emiss_nu.data(), col.data(), ne, nu));
    }
    rc = upload_vec(lp, c->d_mono_logp[id]);
    if (rc) return rc;
    M.mono_logp[id] = c->d_mono_logp[id];
  }
  if (n_sources > 0 && ns > 0) {
    rc = run_final_rounds<ND>(c, first_source_id, n_sources, scattering_only, &M, ITER_MONO + 2u * (uint32_t)inu, 0);
    if (rc) return rc;
  }
  if (n_thermal > 0) {
    // setup_monochromatic_grid_pdfs (grid_monochromatic.f90:50-118)
    const size_t nc = (size_t)c->n_cells;
    if (!c->d_mono_cdf) CUDA_TRY(cudaMalloc(&c->d_mono_cdf, nc * ND * sizeof(double)));
    CUDA_TRY(cudaMemsetAsync(c->d_mono_cdf, 0, nc * ND * sizeof(double), st));
    const int32_t *list = c->grid_type == GEO_OCT   ? c->d_oct_leaves
                          : c->grid_type == GEO_AMR ? c->d_amr_valid
                          : c->grid_type == GEO_VOR ? c->M.vor.valid : nullptr;
    const int64_t n_list = c->grid_type == GEO_OCT   ? (int64_t)c->M.oct.n_leaves
                           : c->grid_type == GEO_AMR ? (int64_t)c->M.amr.n_valid
                           : c->grid_type == GEO_VOR ? (int64_t)c->M.vor.n_valid : (int64_t)nc;
    mono_weights_kernel<<<grid_blocks(c), 256, 0, st>>>(M, list, n_list, c->d_mono_cdf);
    CUDA_TRY(cudaGetLastError());
    c->launches_acc += 1;
    size_t need = 0;
    CUDA_TRY(cub::DeviceScan::InclusiveSum(nullptr, need, c->d_mono_cdf, c->d_mono_cdf, (int)nc, st));
    if (need > c->scan_tmp_bytes) {
      free_dev(c->d_scan_tmp);
      CUDA_TRY(cudaMalloc(&c->d_scan_tmp, need));
      c->scan_tmp_bytes = need;
    }
    bool any = false;
    for (int id = 0; id < ND; ++id) {
      double *cdf = c->d_mono_cdf + (size_t)id * nc;
      CUDA_TRY(cub::DeviceScan::InclusiveSum(c->d_scan_tmp, need, cdf, cdf, (int)nc, st));
      double total = 0.0;
      CUDA_TRY(cudaMemcpyAsync(&total, cdf + nc - 1, sizeof(double), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      M.mono_thermal_w[id] = 0.0;
      if (total > 0.0) {
        mono_divide_kernel<<<grid_blocks(c), 256, 0, st>>>(cdf, (int64_t)nc, total);
        CUDA_TRY(cudaGetLastError());
        // mean_prob x energy_abs_tot = sum over cells of prob x E rho V; x n_dust / n_photons (iter_final_mono.f90:187)
        M.mono_thermal_w[id] = total * (double)ND / (double)n_total_thermal;
        any = true;
      }
      c->launches_acc += 2;
    }
    M.mono_cdf = c->d_mono_cdf;
    if (any) {
      rc = run_final_rounds<ND>(c, first_dust_id, n_thermal, scattering_only, &M, ITER_MONO + 2u * (uint32_t)inu + 1u, 1);
      if (rc) return rc;
    }
  }
  return HYP_OK;
}

}  // namespace

extern "C" {

int hyp_final_mono_photons(hyp_ctx *c, int32_t inu, int64_t first_source_id, int64_t n_sources, int64_t n_total_sources,
                           int64_t first_dust_id, int64_t n_dust, int64_t n_total_dust, int32_t peeloff_scattering_only) {
  if (!c || !c->finalized) return fail(HYP_ERR_STATE, "hyp_finalize_setup has not been called");
  if (!c->images_ready) return fail(HYP_ERR_STATE, "hyp_final_begin has not been called");
  if (c->frequencies.empty()) return fail(HYP_ERR_STATE, "hyp_set_monochromatic has not been called");
  if (inu < 1 || inu > (int)c->frequencies.size()) return fail(HYP_ERR_INVALID, "incorrect inu");
  if (n_sources < 0 || n_dust < 0 || first_source_id < 0 || first_dust_id < 0)
    return fail(HYP_ERR_INVALID, "negative photon count");
  if ((n_sources > 0 && n_total_sources < n_sources) || (n_dust > 0 && n_total_dust < n_dust))
    return fail(HYP_ERR_INVALID, "the job's packet totals are smaller than this call's share");
  CUDA_TRY(cudaSetDevice(c->device));
  c->mono_run = true;
  switch (c->M.n_dust) {
    case 1: return run_final_mono<1>(c, inu, first_source_id, n_sources, n_total_sources, first_dust_id, n_dust, n_total_dust, peeloff_scattering_only);
    case 2: return run_final_mono<2>(c, inu, first_source_id, n_sources, n_total_sources, first_dust_id, n_dust, n_total_dust, peeloff_scattering_only);
    case 3: return run_final_mono<3>(c, inu, first_source_id, n_sources, n_total_sources, first_dust_id, n_dust, n_total_dust, peeloff_scattering_only);
    case 4: return run_final_mono<4>(c, inu, first_source_id, n_sources, n_total_sources, first_dust_id, n_dust, n_total_dust, peeloff_scattering_only);
    default: return fail(HYP_ERR_INVALID, "unsupported number of dust types");
  }
}

int hyp_image_device_buffers(hyp_ctx *c, void **buffer, int64_t *n_values) {
  if (!c || !c->finalized) return fail(HYP_ERR_STATE, "hyp_finalize_setup has not been called");
  CUDA_TRY(cudaSetDevice(c->device));
  int rc = ensure_images(c);
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (buffer) *buffer = c->d_imgbuf;
  if (n_values) *n_values = (int64_t)c->imgbuf_n + SC_COUNT;
  return HYP_OK;
}

int hyp_final_finish(hyp_ctx *c, hyp_iter_stats *st) {
  if (!c || !c->images_ready) return fail(HYP_ERR_STATE, "hyp_final_begin has not been called");
  CUDA_TRY(cudaSetDevice(c->device));
  double sc[SC_COUNT];
  int rc = read_image_scalars(c, sc);
  if (rc) return rc;
  rc = device_error_to_status(c);
  if (rc) return rc;
  // do_final_mono scales every packet itself (iter_final_mono.f90:118,187)
  const bool mono = c->mono_run;
  c->mono_run = false;
  if (!mono && !(sc[SC_ENERGY] > 0.0)) return fail(HYP_ERR_STATE, "no photons were emitted in this iteration");
  // peeled_images_adjust_scale(energy_total / energy_current) (iter_final.f90:140-143)
  const double scale0 = mono ? 1.0 : c->energy_total / sc[SC_ENERGY];
  for (auto &g : c->groups) {
    if (mono) break;
    // binned_images_adjust_scale multiplies by the number of direction bins as well (images_binned.f90:35-39)
    const double scale = g.conf.binned ? scale0 * (double)g.conf.n_theta * (double)g.conf.n_phi : scale0;
    const size_t ns[2] = {g.n_sed, g.n_img}, os[2] = {g.o_sed, g.o_img};
    for (int w = 0; w < 2; ++w) {
      if (!ns[w]) continue;
      image_scale_kernel<<<grid_blocks(c), 256, 0, c->stream>>>(c->d_imgbuf + os[w], (int64_t)ns[w], scale);
      if (g.conf.uncertainties)
        image_scale_kernel<<<grid_blocks(c), 256, 0, c->stream>>>(c->d_imgbuf + os[w] + ns[w], (int64_t)ns[w], scale * scale);
      c->launches_acc += g.conf.uncertainties ? 2 : 1;
    }
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(c->ev3, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  fill_image_stats(c, sc, st);
  if (st) {
    float total = 0.f;
    if (cudaEventElapsedTime(&total, c->ev2, c->ev3) == cudaSuccess) st->epilogue_ms = total - c->kernel_ms_acc;
  }
  return HYP_OK;
}

int hyp_raytracing_photons(hyp_ctx *c, int64_t first_source_id, int64_t n_sources, int64_t n_total_sources,
                           int64_t first_dust_id, int64_t n_dust, int64_t n_total_dust, hyp_iter_stats *st) {
  if (!c || !c->finalized) return fail(HYP_ERR_STATE, "hyp_finalize_setup has not been called");
  if (n_sources < 0 || n_dust < 0 || first_source_id < 0 || first_dust_id < 0)
    return fail(HYP_ERR_INVALID, "negative photon count");
  CUDA_TRY(cudaSetDevice(c->device));
  for (auto &g : c->groups)
    if (g.conf.use_filters && !g.conf.binned)
      return fail(HYP_ERR_INVALID, "filter convolution cannot be used with raytracing");   // image_type.f90:541
  int rc = ensure_images(c);
  if (rc) return rc;
  rc = ensure_ray_tables(c);
  if (rc) return rc;
  CUDA_TRY(cudaMemsetAsync(c->d_imgbuf + c->imgbuf_n, 0, SC_COUNT * sizeof(double), c->stream));
  CUDA_TRY(cudaMemsetAsync(c->d_error, 0, sizeof(int32_t), c->stream));
  c->launches_acc = 0;
  c->kernel_ms_acc = c->flight_ms_acc = 0.f;
  c->rounds_acc = 0;
  CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
  rc = update_source_columns_nd(c);
  if (rc) return rc;
  switch (c->M.n_dust) {
    case 1: rc = run_raytracing<1>(c, first_source_id, n_sources, n_total_sources, first_dust_id, n_dust, n_total_dust); break;
    case 2: rc = run_raytracing<2>(c, first_source_id, n_sources, n_total_sources, first_dust_id, n_dust, n_total_dust); break;
    case 3: rc = run_raytracing<3>(c, first_source_id, n_sources, n_total_sources, first_dust_id, n_dust, n_total_dust); break;
    case 4: rc = run_raytracing<4>(c, first_source_id, n_sources, n_total_sources, first_dust_id, n_dust, n_total_dust); break;
    default: return fail(HYP_ERR_INVALID, "unsupported number of dust types");
  }
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->kernel_ms_acc = ms;
  double sc[SC_COUNT];
  rc = read_image_scalars(c, sc);
  if (rc) return rc;
  rc = device_error_to_status(c);
  if (rc) return rc;
  fill_image_stats(c, sc, st);
  return HYP_OK;
}

int hyp_image_shape(hyp_ctx *c, int32_t group, int32_t which, int64_t dims[6], int32_t *ndim) {
  if (!c || !dims || !ndim) return fail(HYP_ERR_INVALID, "NULL argument");
  if (group < 0 || group >= (int)c->groups.size()) return fail(HYP_ERR_INVALID, "no such image group");
  const HostImage &g = c->groups[group];
  const int n_orig = n_orig_of(g.conf, (int)c->sources.size(), (int)c->dust.size());
  const int n_stokes = g.conf.compute_stokes ? 4 : 1;
  if (which == 0) {
    if (!g.conf.compute_sed) return fail(HYP_ERR_INVALID, "group has no SED");
    const int64_t d[5] = {n_stokes, n_orig, g.conf.n_view, g.conf.n_ap, g.conf.n_wav};
    for (int i = 0; i < 5; ++i) dims[i] = d[i];
    *ndim = 5;
  } else {
    if (!g.conf.compute_image) return fail(HYP_ERR_INVALID, "group has no image");
    const int64_t d[6] = {n_stokes, n_orig, g.conf.n_view, g.conf.n_y, g.conf.n_x, g.conf.n_wav};
    for (int i = 0; i < 6; ++i) dims[i] = d[i];
    *ndim = 6;
  }
  return HYP_OK;
}

// image_write (image_type.f90:608-788): divide by the relative bin width, make the apertures cumulative
static int get_cube(hyp_ctx *c, int32_t group, bool sed, double *out, double *unc) {
  if (!c || !out) return fail(HYP_ERR_INVALID, "NULL argument");
  if (group < 0 || group >= (int)c->groups.size()) return fail(HYP_ERR_INVALID, "no such image group");
  CUDA_TRY(cudaSetDevice(c->device));
  int rc = ensure_images(c);
  if (rc) return rc;
  const HostImage &g = c->groups[group];
  const size_t n = sed ? g.n_sed : g.n_img, off = sed ? g.o_sed : g.o_img;
  if (!n) return fail(HYP_ERR_INVALID, sed ? "group has no SED" : "group has no image");
  CUDA_TRY(cudaMemcpyAsync(out, c->d_imgbuf + off, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  const bool have_unc = unc && g.conf.uncertainties;
  if (have_unc) CUDA_TRY(cudaMemcpyAsync(unc, c->d_imgbuf + off + n, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  const int n_nu = g.conf.n_wav;
  // with filters the flux stays in F_nu dnu: the filter carries the normalisation (image_type.f90:649-657)
  if (g.conf.inu_min > 0) {
    // image_type.f90:679-682,737-740: nu F_nu at the exact frequencies
    const double *nu = c->frequencies.data() + (g.conf.inu_min - 1);
    for (size_t i = 0; i < n; ++i) out[i] = out[i] * nu[i % (size_t)n_nu];
    if (have_unc)
      for (size_t i = 0; i < n; ++i) unc[i] = std::sqrt(unc[i]) * nu[i % (size_t)n_nu];
  } else {
    const double dnunorm = g.conf.use_filters ? 1.0 : std::pow(g.nu_max / g.nu_min, +0.5 / (double)n_nu) - std::pow(g.nu_max / g.nu_min, -0.5 / (double)n_nu);
    for (size_t i = 0; i < n; ++i) out[i] = out[i] / dnunorm;
    if (have_unc)
      for (size_t i = 0; i < n; ++i) unc[i] = std::sqrt(unc[i]) / dnunorm;
  }
  if (sed) {
    const size_t n_ap = g.conf.n_ap, outer = n / ((size_t)n_nu * n_ap);
    for (size_t o = 0; o < outer; ++o)
      for (size_t ia = 1; ia < n_ap; ++ia)
        for (size_t inu = 0; inu < (size_t)n_nu; ++inu) {
          const size_t k = inu + n_nu * (ia + n_ap * o), km = inu + n_nu * (ia - 1 + n_ap * o);
          out[k] = out[km] + out[k];
          if (have_unc) unc[k] = std::sqrt(unc[km] * unc[km] + unc[k] * unc[k]);
        }
  }
  return HYP_OK;
}

int hyp_get_sed(hyp_ctx *c, int32_t group, double *sed, double *unc) { return get_cube(c, group, true, sed, unc); }
int hyp_get_image(hyp_ctx *c, int32_t group, double *image, double *unc) { return get_cube(c, group, false, image, unc); }

}  // extern "C"
