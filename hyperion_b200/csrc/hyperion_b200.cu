// hyperion_b200.cu -- CUDA kernels and C ABI of the B200-native photon-packet engine.
//
// Hot path: the Lucy iteration of the reference (do_lucy, src/main/iter_lucy.f90:66-237) as a
// persistent-threads photon loop.  Each lane owns one packet and runs a three-state machine
//   EMIT     emit (src/sources/source.f90:100-179)                  -> FLIGHT
//   FLIGHT   grid_integrate (src/grid/grid_propagate_3d.f90:35-234)  -> INTERACT | EMIT (escaped/killed)
//   INTERACT interact (src/dust/dust_interact.f90:22-79)             -> FLIGHT | EMIT (killed)
// The warp keeps stepping cell crossings while most lanes are in FLIGHT and only services the
// divergent EMIT / INTERACT work once enough lanes are waiting, so the crossing loop (>90 % of
// the work) runs converged.
//
// HBM layout: one 16-byte record {density, energy_sum} per (cell, dust) so the
// density read and the deposit RED of a crossing touch the same 32-byte sector.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "device.cuh"

using namespace hyp;

// =============================================================================================
// device-side model description
// =============================================================================================
struct CellRec {
  double rho;
  double esum;
};

enum { SC_ENERGY = 0, SC_KILLED_GEO, SC_KILLED_INT, SC_CROSS, SC_ABS, SC_SCAT, SC_ESC, SC_PHOTONS, SC_COUNT };

struct ModelDev {
  int32_t n1, n2, n3, n_dust, n_sources;
  int64_t n_cells;
  const double *w1, *w2, *w3;
  CellRec *cells;           // [n_cells][n_dust]
  double *specific_energy;  // [n_cells][n_dust]
  int32_t *jnu_id;          // [n_cells][n_dust]
  double *jnu_frac;         // [n_cells][n_dust]
  DustDev dust[MAX_DUST];
  const SourceDev *sources;
  const SpectrumDev *spectra;
  double min_energy[MAX_DUST];
  // run configuration
  uint64_t seed;
  int64_t n_inter_max;
  int32_t kill_on_absorb, kill_on_scatter, sample_evenly, enforce_energy_range;
  // outputs
  double *scalars;                 // [SC_COUNT], directly after the reduced sum grid
  unsigned long long *work_counter;
  int32_t *error_flag;
};

enum { ERR_NONE = 0, ERR_NOT_IN_CELL = 1, ERR_NU_RANGE = 2, ERR_SCATTER = 3 };

// =============================================================================================
// photon state
// =============================================================================================
enum : int { ST_EMIT = 0, ST_FLIGHT = 1, ST_INTERACT = 2, ST_DONE = 3 };

template <int ND>
struct Photon {
  // flight origin, direction, path length travelled from the origin
  double r0x, r0y, r0z;
  double vx, vy, vz;
  double ivx, ivy, ivz;  // 1/v, +-Inf for a ray parallel to the walls of that axis
  double t;
  // distance (from the flight origin) at which the next x / y / z wall is reached
  double tnx, tny, tnz;
  double tau_left;
  // optical constants at the current frequency, per dust type
  double chi[ND], kE[ND], albedo[ND];
  double nu, energy;
  double sQ, sU, sV;  // Stokes (I = 1)
  int32_t ix, iy, iz;
  int32_t ic;         // 1-D cell id used for density / deposits (p%icell%ic of the reference)
  int64_t n_inter;
};

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

// Set up the wall-distance table for a new straight flight from (r0, v) in cell (ix,iy,iz).
// W holds the three wall arrays back to back: w1 at 0, w2 at o2, w3 at o3.
template <int ND>
__device__ __forceinline__ void start_flight(const double *__restrict__ W, int o2, int o3, Photon<ND> &p, Rng &rng) {
  // random_exp (lib_random.f90:227-236)
  p.tau_left = -log(1.0 - rng.next());
  p.t = 0.0;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  p.ivx = 1.0 / p.vx;
  p.ivy = 1.0 / p.vy;
  p.ivz = 1.0 / p.vz;
  // distance to the wall ahead on each axis; a ray parallel to an axis never reaches its walls
  p.tnx = p.vx != 0.0 ? (W[p.ix + (p.vx > 0.0 ? 1 : 0)] - p.r0x) * p.ivx : inf;
  p.tny = p.vy != 0.0 ? (W[o2 + p.iy + (p.vy > 0.0 ? 1 : 0)] - p.r0y) * p.ivy : inf;
  p.tnz = p.vz != 0.0 ? (W[o3 + p.iz + (p.vz > 0.0 ? 1 : 0)] - p.r0z) * p.ivz : inf;
  // rounding at an interaction point can leave it a few ulp behind the wall it faces
  p.tnx = fmax(p.tnx, 0.0);
  p.tny = fmax(p.tny, 0.0);
  p.tnz = fmax(p.tnz, 0.0);
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

// emit (src/sources/source.f90:100-179): returns false on a fatal model error
template <int ND>
__device__ bool emit_photon(const ModelDev &M, Photon<ND> &p, Rng &rng, double &energy_emitted) {
  int is = 0;
  const int ns = M.n_sources;
  if (ns > 1) {
    double xi = rng.next();
    if (M.sample_evenly) {
      is = min((int)(xi * ns), ns - 1);
    } else {
      // sample_pdf_discrete_dp (type_pdf.f90:313-337): first source with cdf >= xi
      if (xi >= M.sources[ns - 1].cdf) {
        is = ns - 1;
      } else {
        is = 0;
        while (is < ns - 1 && xi > M.sources[is].cdf) ++is;
      }
    }
  }
  const SourceDev &S = M.sources[is];
  // emit_from_point (source_type.f90:539-564)
  p.r0x = S.x;
  p.r0y = S.y;
  p.r0z = S.z;
  Angle a = random_sphere_angle(rng);
  set_dir(p, a);
  p.sQ = p.sU = p.sV = 0.0;
  p.energy = 1.0;
  if (S.freq_type == HYP_SPECTRUM_BLACKBODY) {
    p.nu = sample_planck(rng, S.temperature);
  } else {
    const SpectrumDev &sp = M.spectra[S.spectrum];
    p.nu = sample_powerlaw(sp.B + sp.L.o_x, sp.B + sp.L.o_cdf, sp.B + sp.L.o_invb, sp.B + sp.L.o_rm1, sp.L.n,
                           rng.next());
  }
  if (M.sample_evenly) p.energy = p.energy * S.pdf * ns;
  energy_emitted += p.energy;
  if (!update_optconsts<ND>(M, p)) {
    atomicMax(M.error_flag, ERR_NU_RANGE);
    return false;
  }
  int fx, fy, fz;
  bool ok = place_axis(M.w1, M.n1, p.r0x, p.vx, p.ix, fx);
  ok = place_axis(M.w2, M.n2, p.r0y, p.vy, p.iy, fy) && ok;
  ok = place_axis(M.w3, M.n3, p.r0z, p.vz, p.iz, fz) && ok;
  if (!ok) {
    atomicMax(M.error_flag, ERR_NOT_IN_CELL);
    return false;
  }
  p.ic = (fz * M.n2 + fy) * M.n1 + fx;
  p.n_inter = 0;
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

// interact (src/dust/dust_interact.f90:22-79).  Returns: 0 continue, 1 packet finished (killed).
template <int ND>
__device__ int interact_photon(const ModelDev &M, Photon<ND> &p, Rng &rng, uint32_t &n_abs, uint32_t &n_scat,
                               uint32_t &n_killed_int) {
  // the loop guard of do_lucy (iter_lucy.f90:193-198)
  p.n_inter += 1;
  if (p.n_inter > M.n_inter_max) {
    ++n_killed_int;
    return 1;
  }
  // move to the interaction point
  p.r0x = p.r0x + p.t * p.vx;
  p.r0y = p.r0y + p.t * p.vy;
  p.r0z = p.r0z + p.t * p.vz;
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
  const double albedo = p.albedo[id];
  const double xi = rng.next();
  bool scattered;
  if (xi > albedo) {
    // dust_emit (dust_type_4elem.f90:334-354) + dust_sample_j_nu (:379-398)
    const size_t k = (size_t)p.ic * ND + id;
    const int jid = M.jnu_id[k];
    const double frac = M.jnu_frac[k];
    const double x2 = rng.next();
    const int ne = d.L.n_enu;
    const double *enu = d.B + d.L.o_enu;
    const double nu1 = sample_powerlaw(enu, d.B + d.L.o_ecdf + (size_t)jid * ne, d.B + d.L.o_einvb + (size_t)jid * (ne - 1),
                                       d.B + d.L.o_erm1 + (size_t)jid * (ne - 1), ne, x2);
    const double nu2 = sample_powerlaw(enu, d.B + d.L.o_ecdf + (size_t)(jid + 1) * ne,
                                       d.B + d.L.o_einvb + (size_t)(jid + 1) * (ne - 1),
                                       d.B + d.L.o_erm1 + (size_t)(jid + 1) * (ne - 1), ne, x2);
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
  if ((M.kill_on_scatter && scattered) || (M.kill_on_absorb && !scattered)) return 1;
  return 0;
}

// =============================================================================================
// the photon kernel
// =============================================================================================
constexpr int LUCY_THREADS = 256;
#ifndef LUCY_MIN_BLOCKS
#define LUCY_MIN_BLOCKS 3
#endif
#ifndef LUCY_STEPS_PER_ROUND
#define LUCY_STEPS_PER_ROUND 32   // crossings a lane may take before the warp re-votes
#endif
#ifndef LUCY_SERVICE_THRESHOLD
#define LUCY_SERVICE_THRESHOLD 8  // waiting lanes that trigger an EMIT/INTERACT service pass
#endif

template <int ND>
__global__ void __launch_bounds__(LUCY_THREADS, LUCY_MIN_BLOCKS)
lucy_photon_kernel(const ModelDev M, const unsigned long long first_id, const unsigned long long n_photons,
                   const uint32_t iteration, const int walls_in_smem) {
  extern __shared__ double s_walls[];
  const int n1 = M.n1, n2 = M.n2, n3 = M.n3;
  const int o2 = n1 + 1, o3 = n1 + n2 + 2;
  // wall table: shared memory when it fits (always for the grids of BASELINE.json), else global
  const double *__restrict__ W = M.w1;  // w1|w2|w3 are contiguous in global memory as well
  if (walls_in_smem) {
    for (int i = threadIdx.x; i < n1 + n2 + n3 + 3; i += blockDim.x) s_walls[i] = M.w1[i];
    __syncthreads();
    W = s_walls;
  }

  Photon<ND> p;
  Rng rng;
  int state = ST_EMIT;
  double energy_emitted = 0.0;
  uint32_t n_cross = 0, n_abs = 0, n_scat = 0, n_esc = 0, n_kill_int = 0, n_run = 0;
  unsigned long long cross_hi = 0;
  const unsigned lane = threadIdx.x & 31;
  CellRec *__restrict__ cells = M.cells;

  for (;;) {
    const unsigned m_flight = __ballot_sync(0xffffffffu, state == ST_FLIGHT);
    const unsigned m_wait = __ballot_sync(0xffffffffu, state == ST_EMIT || state == ST_INTERACT);
    if (m_flight == 0 && m_wait == 0) break;

    if (m_wait != 0 && (__popc(m_wait) >= LUCY_SERVICE_THRESHOLD || m_flight == 0)) {
      // ---------------- service pass: divergent, rare ----------------
      if (state == ST_INTERACT) {
        int fin = interact_photon<ND>(M, p, rng, n_abs, n_scat, n_kill_int);
        if (fin) {
          state = ST_EMIT;
        } else {
          start_flight<ND>(W, o2, o3, p, rng);
          state = ST_FLIGHT;
        }
      }
      // claim packet ids for every lane that needs one (warp-aggregated)
      const unsigned m_emit = __ballot_sync(0xffffffffu, state == ST_EMIT);
      if (m_emit) {
        const int leader = __ffs(m_emit) - 1;
        unsigned long long base = 0;
        if ((int)lane == leader) base = atomicAdd(M.work_counter, (unsigned long long)__popc(m_emit));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (state == ST_EMIT) {
          const unsigned long long k = base + __popc(m_emit & ((1u << lane) - 1u));
          if (k >= n_photons) {
            state = ST_DONE;
          } else {
            rng.init(M.seed, first_id + k, iteration);
            ++n_run;
            if (emit_photon<ND>(M, p, rng, energy_emitted)) {
              // a packet emitted on the outer wall moving outwards escapes immediately
              if (p.ix < 0 || p.ix >= n1 || p.iy < 0 || p.iy >= n2 || p.iz < 0 || p.iz >= n3) {
                ++n_esc;
                state = ST_EMIT;
              } else {
                start_flight<ND>(W, o2, o3, p, rng);
                state = ST_FLIGHT;
              }
            } else {
              state = ST_EMIT;  // fatal model error is flagged; drain the remaining ids quickly
            }
          }
        }
      }
      continue;
    }

    // ---------------- flight pass: cell crossings (grid_integrate) ----------------
    // Branch-free DDA step: the axis whose wall is reached first is picked with selects, so all
    // lanes of the warp execute the same instruction stream whatever their direction.
    if (state == ST_FLIGHT) {
#pragma unroll 1
      for (int step = 0; step < LUCY_STEPS_PER_ROUND; ++step) {
        CellRec *rec = cells + (size_t)p.ic * ND;
        double rho[ND];
#pragma unroll
        for (int id = 0; id < ND; ++id) rho[id] = rec[id].rho;

        const bool bx = (p.tnx <= p.tny) & (p.tnx <= p.tnz);
        const bool by = (!bx) & (p.tny <= p.tnz);
        const double t_exit = bx ? p.tnx : (by ? p.tny : p.tnz);
        const double ds = t_exit - p.t;
        // geometry of the step (does not depend on the density): next cell and its far wall
        const double v_ax = bx ? p.vx : (by ? p.vy : p.vz);
        const int fwd = v_ax > 0.0 ? 1 : 0;
        const int i_new = (bx ? p.ix : (by ? p.iy : p.iz)) + 2 * fwd - 1;
        const int n_ax = bx ? n1 : (by ? n2 : n3);
        const bool out = (unsigned)i_new >= (unsigned)n_ax;
        const int woff = bx ? 0 : (by ? o2 : o3);
        const double wall = W[woff + (out ? 0 : i_new + fwd)];
        const double tn_new = (wall - (bx ? p.r0x : (by ? p.r0y : p.r0z))) * (bx ? p.ivx : (by ? p.ivy : p.ivz));

        double chi_rho = 0.0;
#pragma unroll
        for (int id = 0; id < ND; ++id) chi_rho += p.chi[id] * rho[id];
        const double tau_cell = chi_rho * ds;
        ++n_cross;
        if (tau_cell < p.tau_left) {
          // cross the whole cell: deposit tmin * kappa * E (grid_propagate_3d.f90:148-160)
#pragma unroll
          for (int id = 0; id < ND; ++id)
            if (rho[id] > 0.0) atomicAdd(&rec[id].esum, ds * p.kE[id]);
          p.tau_left -= tau_cell;
          p.t = t_exit;
          if (out) {
            ++n_esc;
            state = ST_EMIT;
            break;
          }
          p.ix = bx ? i_new : p.ix;
          p.iy = by ? i_new : p.iy;
          p.iz = (bx | by) ? p.iz : i_new;
          p.tnx = bx ? tn_new : p.tnx;
          p.tny = by ? tn_new : p.tny;
          p.tnz = (bx | by) ? p.tnz : tn_new;
          p.ic = (p.iz * n2 + p.iy) * n1 + p.ix;
        } else {
          // interaction inside this cell (grid_propagate_3d.f90:186-228)
          const double tact = tau_cell > 0.0 ? ds * (p.tau_left / tau_cell) : 0.0;
#pragma unroll
          for (int id = 0; id < ND; ++id)
            if (rho[id] > 0.0) atomicAdd(&rec[id].esum, tact * p.kE[id]);
          p.t += tact;
          state = ST_INTERACT;
          break;
        }
      }
      if (n_cross > 0x7fffff00u) {
        cross_hi += n_cross;
        n_cross = 0;
      }
    }
  }

  // ---------------- reduce the per-lane counters ----------------
  unsigned long long cross = cross_hi + n_cross;
  double vals[SC_COUNT];
  vals[SC_ENERGY] = energy_emitted;
  vals[SC_KILLED_GEO] = 0.0;
  vals[SC_KILLED_INT] = (double)n_kill_int;
  vals[SC_CROSS] = (double)cross;
  vals[SC_ABS] = (double)n_abs;
  vals[SC_SCAT] = (double)n_scat;
  vals[SC_ESC] = (double)n_esc;
  vals[SC_PHOTONS] = (double)n_run;
#pragma unroll
  for (int q = 0; q < SC_COUNT; ++q) {
    double v = vals[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && v != 0.0) atomicAdd(M.scalars + q, v);
  }
}

// =============================================================================================
// streaming kernels around the photon loop
// =============================================================================================

// grid_reset_energy (src/grid/grid_generic.f90:21-27) + precompute_jnu_var
// (src/grid/grid_physics_3d.f90:613-629, dust_jnu_var_pos_frac dust_type_4elem.f90:295-320)
__global__ void lucy_begin_kernel(ModelDev M) {
  const int nd = M.n_dust;
  const int64_t n = M.n_cells * nd;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    M.cells[k].esum = 0.0;
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

// gather the deposit sums into the contiguous reduction buffer
__global__ void gather_sums_kernel(ModelDev M, double *__restrict__ sums) {
  const int64_t n = M.n_cells * M.n_dust;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    sums[k] = M.cells[k].esum;
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
__global__ void lucy_finish_kernel(ModelDev M, const double *__restrict__ sums, double scale) {
  const int nd = M.n_dust;
  const int64_t n = M.n_cells * nd;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const int id = (int)(k % nd);
    const int64_t ic = k / nd;
    const int i1 = (int)(ic % M.n1), i2 = (int)((ic / M.n1) % M.n2), i3 = (int)(ic / ((int64_t)M.n1 * M.n2));
    const double vol = ((M.w1[i1 + 1] - M.w1[i1]) * (M.w2[i2 + 1] - M.w2[i2])) * (M.w3[i3 + 1] - M.w3[i3]);
    const DustDev &d = M.dust[id];
    double e = sums[k] * scale / vol;
    if (vol == 0.0) e = 0.0;
    e = clamp_energy(M, d, id, e);
    if (d.L.sublimation_mode != 0 && e > d.L.sublimation_specific_energy) {
      const double es = d.L.sublimation_specific_energy;
      if (d.L.sublimation_mode == 1) {
        M.cells[k].rho = 0.0;
        e = M.min_energy[id];
      } else if (d.L.sublimation_mode == 2) {
        const double q = mean_opacity_loglog(d, d.L.o_logchi_ross, e) / mean_opacity_loglog(d, d.L.o_logchi_ross, es);
        M.cells[k].rho = M.cells[k].rho * es / e * (q * q);
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

// =============================================================================================
// host side: context + C ABI
// =============================================================================================
namespace {

thread_local std::string g_error;

struct HostDust {
  DustLayout L;
  std::vector<double> buf;
  double *dev = nullptr;
};
struct HostSpectrum {
  SpectrumLayout L;
  std::vector<double> buf;
  double *dev = nullptr;
};

}  // namespace

struct hyp_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  // host model
  int n1 = 0, n2 = 0, n3 = 0;
  int64_t n_cells = 0;
  std::vector<double> w1, w2, w3;
  std::vector<HostDust> dust;
  std::vector<hyp_source> sources;
  std::vector<HostSpectrum> spectra;
  std::vector<int> source_spectrum;
  hyp_run_conf conf;
  std::vector<double> h_density, h_energy, h_min_energy;
  bool have_density = false, have_energy = false;
  double energy_total = 0.0;
  // device
  double *d_w = nullptr;
  CellRec *d_cells = nullptr;
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
  float kernel_ms_acc = 0.f;
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
    default:
      return fail(HYP_ERR_PHYSICS, "ERROR: in sampling mu for scattering");
  }
}

template <typename T>
void free_dev(T *&p) {
  if (p) cudaFree(p);
  p = nullptr;
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
  free_dev(c->d_cells);
  free_dev(c->d_energy);
  free_dev(c->d_jfrac);
  free_dev(c->d_sums);
  free_dev(c->d_stage);
  free_dev(c->d_jid);
  free_dev(c->d_sources);
  free_dev(c->d_spectra);
  free_dev(c->d_work);
  free_dev(c->d_error);
  for (auto &d : c->dust) free_dev(d.dev);
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
  c->n1 = n1;
  c->n2 = n2;
  c->n3 = n3;
  c->n_cells = (int64_t)n1 * n2 * n3;
  c->w1.assign(w1, w1 + n1 + 1);
  c->w2.assign(w2, w2 + n2 + 1);
  c->w3.assign(w3, w3 + n3 + 1);
  return HYP_OK;
}

int hyp_add_dust(hyp_ctx *c, const hyp_dust_tables *t) {
  if (!c || !t) return fail(HYP_ERR_INVALID, "NULL argument");
  if (c->finalized) return fail(HYP_ERR_STATE, "model is frozen");
  if ((int)c->dust.size() >= MAX_DUST) return fail(HYP_ERR_INVALID, "too many dust types (max 4)");
  try {
    HostDust d;
    build_dust(*t, d.L, d.buf);
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
  if (s->type != HYP_SOURCE_POINT) return fail(HYP_ERR_INVALID, "only point sources are implemented on the device");
  if (!(s->luminosity >= 0.0)) return fail(HYP_ERR_INVALID, "source luminosity should be positive");
  int spec = -1;
  if (s->spectrum_type == HYP_SPECTRUM_TABLE) {
    try {
      HostSpectrum sp;
      build_spectrum(s->spec_nu, s->spec_fnu, s->n_spec, sp.L, sp.buf);
      spec = (int)c->spectra.size();
      c->spectra.push_back(std::move(sp));
    } catch (std::exception &e) {
      return fail(HYP_ERR_INVALID, e.what());
    }
  } else if (s->spectrum_type != HYP_SPECTRUM_BLACKBODY) {
    return fail(HYP_ERR_INVALID, "unknown spectrum specifier");
  }
  hyp_source copy = *s;
  copy.spec_nu = copy.spec_fnu = nullptr;
  c->sources.push_back(copy);
  c->source_spectrum.push_back(spec);
  return HYP_OK;
}

int hyp_set_run_conf(hyp_ctx *c, const hyp_run_conf *conf) {
  if (!c || !conf) return fail(HYP_ERR_INVALID, "NULL argument");
  if (conf->use_mrw) return fail(HYP_ERR_INVALID, "the modified random walk is not implemented on the device yet");
  c->conf = *conf;
  if (c->finalized) {
    c->M.seed = (uint64_t)conf->seed;
    c->M.n_inter_max = conf->n_inter_max;
    c->M.kill_on_absorb = conf->kill_on_absorb;
    c->M.kill_on_scatter = conf->kill_on_scatter;
    c->M.sample_evenly = conf->sample_sources_evenly;
    c->M.enforce_energy_range = conf->enforce_energy_range;
  }
  return HYP_OK;
}

static int upload_density(hyp_ctx *c, const double *density) {
  const size_t n = (size_t)c->n_cells * c->dust.size();
  CUDA_TRY(cudaMemcpyAsync(c->d_stage, density, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
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
  // dust tables
  for (int id = 0; id < nd; ++id) {
    HostDust &d = c->dust[id];
    CUDA_TRY(cudaMalloc(&d.dev, d.buf.size() * sizeof(double)));
    CUDA_TRY(cudaMemcpy(d.dev, d.buf.data(), d.buf.size() * sizeof(double), cudaMemcpyHostToDevice));
    M.dust[id].L = d.L;
    M.dust[id].B = d.dev;
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
    sd[i].spectrum = c->source_spectrum[i];
    sd[i].pdf = s.luminosity / ltot;
    cum += sd[i].pdf;
    sd[i].cdf = cum;
  }
  for (auto &s : sd) s.cdf /= cum;
  CUDA_TRY(cudaMalloc(&c->d_sources, sd.size() * sizeof(SourceDev)));
  CUDA_TRY(cudaMemcpy(c->d_sources, sd.data(), sd.size() * sizeof(SourceDev), cudaMemcpyHostToDevice));
  M.sources = c->d_sources;
  // grids
  CUDA_TRY(cudaMalloc(&c->d_cells, n * sizeof(CellRec)));
  CUDA_TRY(cudaMalloc(&c->d_energy, n * sizeof(double)));
  CUDA_TRY(cudaMalloc(&c->d_jfrac, n * sizeof(double)));
  CUDA_TRY(cudaMalloc(&c->d_jid, n * sizeof(int32_t)));
  CUDA_TRY(cudaMalloc(&c->d_sums, (n + SC_COUNT) * sizeof(double)));
  CUDA_TRY(cudaMalloc(&c->d_stage, n * sizeof(double)));
  CUDA_TRY(cudaMalloc(&c->d_work, sizeof(unsigned long long)));
  CUDA_TRY(cudaMalloc(&c->d_error, sizeof(int32_t)));
  CUDA_TRY(cudaMemset(c->d_cells, 0, n * sizeof(CellRec)));
  CUDA_TRY(cudaMemset(c->d_sums, 0, (n + SC_COUNT) * sizeof(double)));
  CUDA_TRY(cudaMemset(c->d_error, 0, sizeof(int32_t)));
  M.cells = c->d_cells;
  M.specific_energy = c->d_energy;
  M.jnu_id = c->d_jid;
  M.jnu_frac = c->d_jfrac;
  M.scalars = c->d_sums + n;
  M.work_counter = c->d_work;
  M.error_flag = c->d_error;
  M.seed = (uint64_t)c->conf.seed;
  M.n_inter_max = c->conf.n_inter_max;
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
  lucy_begin_kernel<<<grid_blocks(c), 256, 0, c->stream>>>(c->M);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemsetAsync(c->d_sums + n, 0, SC_COUNT * sizeof(double), c->stream));
  CUDA_TRY(cudaMemsetAsync(c->d_error, 0, sizeof(int32_t), c->stream));
  c->sums_gathered = false;
  c->kernel_ms_acc = 0.f;
  return HYP_OK;
}

int hyp_lucy_photons(hyp_ctx *c, int64_t first_id, int64_t n_photons, int64_t iteration) {
  if (!c || !c->finalized) return fail(HYP_ERR_STATE, "hyp_finalize_setup has not been called");
  if (n_photons < 0 || first_id < 0) return fail(HYP_ERR_INVALID, "negative photon count");
  if (n_photons == 0) return HYP_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemsetAsync(c->d_work, 0, sizeof(unsigned long long), c->stream));
  int per_sm = 0;
  const int nd = c->M.n_dust;
  // the three wall arrays go to shared memory when they fit next to 3+ resident blocks
  size_t wall_bytes = (size_t)(c->n1 + c->n2 + c->n3 + 3) * sizeof(double);
  int walls_smem = 1;
  if (wall_bytes > 48 * 1024) {
    wall_bytes = 0;
    walls_smem = 0;
  }
#define LAUNCH(ND)                                                                                              \
  do {                                                                                                          \
    CUDA_TRY(cudaFuncSetAttribute(lucy_photon_kernel<ND>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                  (int)wall_bytes));                                                            \
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lucy_photon_kernel<ND>, LUCY_THREADS,       \
                                                           wall_bytes));                                        \
    int64_t blocks = (int64_t)per_sm * c->sm_count;                                                             \
    int64_t need = (n_photons + LUCY_THREADS - 1) / LUCY_THREADS;                                               \
    if (blocks > need) blocks = need;                                                                           \
    if (blocks < 1) blocks = 1;                                                                                 \
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));                                                               \
    lucy_photon_kernel<ND><<<(int)blocks, LUCY_THREADS, wall_bytes, c->stream>>>(                               \
        c->M, (unsigned long long)first_id, (unsigned long long)n_photons, (uint32_t)iteration, walls_smem);    \
    CUDA_TRY(cudaGetLastError());                                                                               \
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));                                                               \
  } while (0)
  switch (nd) {
    case 1: LAUNCH(1); break;
    case 2: LAUNCH(2); break;
    case 3: LAUNCH(3); break;
    case 4: LAUNCH(4); break;
    default: return fail(HYP_ERR_INVALID, "unsupported number of dust types");
  }
#undef LAUNCH
  c->sums_gathered = false;
  return HYP_OK;
}

static int gather_sums(hyp_ctx *c) {
  if (!c->sums_gathered) {
    gather_sums_kernel<<<grid_blocks(c), 256, 0, c->stream>>>(c->M, c->d_sums);
    CUDA_TRY(cudaGetLastError());
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
  if (n_values) *n_values = c->n_cells * (int64_t)c->dust.size() + SC_COUNT;
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
  lucy_finish_kernel<<<grid_blocks(c), 256, 0, c->stream>>>(c->M, c->d_sums, scale);
  CUDA_TRY(cudaGetLastError());
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
    float ms = 0.f, total = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) st->kernel_ms = ms;
    if (cudaEventElapsedTime(&total, c->ev2, c->ev3) == cudaSuccess) st->epilogue_ms = total - ms;
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

}  // extern "C"
