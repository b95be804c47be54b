// march_geo.cuh -- the marches of grid_propagate_3d.f90 for the geometries that are walked one wall search
// at a time (spherical / cylindrical polar grids, octrees).  Each geometry provides the same small
// interface (Geo<GEO>), the marches and kernels are written once.
// Included by hyperion_b200.cu after ModelDev / CellRec are defined and before imaging.cuh.
#pragma once

enum { MARCH_ESCAPED = 1, MARCH_INTERACT = 2, MARCH_KILLED = 4, MARCH_REABSORBED = 8 };

template <int GEO>
struct Geo;

// Cartesian grids, one wall search per crossing (grid_geometry_cartesian_3d.f90:424-521).  The Lucy and
// imaging iterations of Cartesian models normally run through the specialised look-ahead kernels of
// hyperion_b200.cu; this plain form serves models with spherical sources, whose flights must also be
// tested against the stellar surfaces at every step.
struct CarRay {
  double r0x, r0y, r0z, vx, vy, vz, ivx, ivy, ivz;
  double t;
  int ix, iy, iz, ic;
};

template <>
struct Geo<GEO_CAR> {
  using Ray = CarRay;
  struct Cross {
    int wall;
  };
  static __device__ __forceinline__ bool find_cell(const ModelDev &M, double rx, double ry, double rz, double vx,
                                                   double vy, double vz, int &ix, int &iy, int &iz, int &ic) {
    int fx, fy, fz;
    bool ok = place_axis(M.w1, M.n1, rx, vx, ix, fx);
    ok = place_axis(M.w2, M.n2, ry, vy, iy, fy) && ok;
    ok = place_axis(M.w3, M.n3, rz, vz, iz, fz) && ok;
    if (!ok) return false;
    ic = (fz * M.n2 + fy) * M.n1 + fx;
    return true;
  }
  static __device__ __forceinline__ void start(const ModelDev &M, Ray &R, double rx, double ry, double rz, double vx,
                                               double vy, double vz, int ix, int iy, int iz, int ic) {
    R.r0x = rx; R.r0y = ry; R.r0z = rz;
    R.vx = vx; R.vy = vy; R.vz = vz;
    R.ivx = 1.0 / vx; R.ivy = 1.0 / vy; R.ivz = 1.0 / vz;
    R.t = 0.0;
    R.ix = ix; R.iy = iy; R.iz = iz; R.ic = ic;
  }
  static __device__ __forceinline__ bool escaped(const ModelDev &M, const Ray &R) {
    return (unsigned)R.ix >= (unsigned)M.n1 || (unsigned)R.iy >= (unsigned)M.n2 || (unsigned)R.iz >= (unsigned)M.n3;
  }
  static __device__ __forceinline__ bool find_wall(const ModelDev &M, const Ray &R, double &dt, Cross &c) {
    const double huge = 1.7976931348623157e308;
    const bool px = R.vx > 0.0, py = R.vy > 0.0, pz = R.vz > 0.0;
    const double tx = R.vx != 0.0 ? (__ldg(M.w1 + R.ix + (px ? 1 : 0)) - R.r0x) * R.ivx - R.t : huge;
    const double ty = R.vy != 0.0 ? (__ldg(M.w2 + R.iy + (py ? 1 : 0)) - R.r0y) * R.ivy - R.t : huge;
    const double tz = R.vz != 0.0 ? (__ldg(M.w3 + R.iz + (pz ? 1 : 0)) - R.r0z) * R.ivz - R.t : huge;
    if (tx <= ty && tx <= tz) { c.wall = px ? 1 : 0; dt = tx; }
    else if (ty <= tz) { c.wall = py ? 3 : 2; dt = ty; }
    else { c.wall = pz ? 5 : 4; dt = tz; }
    if (dt < 0.0) dt = 0.0;
    return dt < huge;
  }
  static __device__ __forceinline__ void step(const ModelDev &M, Ray &R, const Cross &c) {
    const int s = (c.wall & 1) ? 1 : -1, axis = c.wall >> 1;
    if (axis == 0) R.ix += s; else if (axis == 1) R.iy += s; else R.iz += s;
    R.ic = (R.iz * M.n2 + R.iy) * M.n1 + R.ix;
  }
  static __device__ __forceinline__ void stop_inside(Ray &R) {}
  static __device__ __forceinline__ void store(const Ray &R, int &ix, int &iy, int &iz, int &ic) {
    ix = R.ix; iy = R.iy; iz = R.iz; ic = R.ic;
  }
};

// spherical and cylindrical polar grids (geometry_sph.cuh)
template <>
struct Geo<GEO_SPH> {
  using Ray = SphRay;
  struct Cross {
    int d1, d2, d3;
  };
  // find_cell: the slot / job keeps (ix, iy, iz) as find_cell reports them; adjust_wall runs in start()
  static __device__ __forceinline__ bool find_cell(const ModelDev &M, double rx, double ry, double rz, double vx,
                                                   double vy, double vz, int &ix, int &iy, int &iz, int &ic) {
    if (!sph_find_cell(M.sph, rx, ry, rz, vx, vy, vz, ix, iy, iz)) return false;
    ic = (iz * M.sph.n2 + iy) * M.sph.n1 + ix;
    return true;
  }
  static __device__ __forceinline__ void start(const ModelDev &M, Ray &R, double rx, double ry, double rz, double vx,
                                               double vy, double vz, int ix, int iy, int iz, int ic) {
    sph_start(M.sph, R, rx, ry, rz, vx, vy, vz, ix, iy, iz);
  }
  static __device__ __forceinline__ bool escaped(const ModelDev &M, const Ray &R) { return sph_escaped(M.sph, R); }
  static __device__ __forceinline__ bool find_wall(const ModelDev &M, const Ray &R, double &dt, Cross &c) {
    return sph_find_wall(M.sph, R, dt, c.d1, c.d2, c.d3);
  }
  static __device__ __forceinline__ void step(const ModelDev &M, Ray &R, const Cross &c) { sph_step(M.sph, R, c.d1, c.d2, c.d3); }
  static __device__ __forceinline__ void stop_inside(Ray &R) { R.ow1 = R.ow2 = R.ow3 = 0; }
  static __device__ __forceinline__ void store(const Ray &R, int &ix, int &iy, int &iz, int &ic) {
    ix = R.i1; iy = R.i2; iz = R.i3; ic = R.ic;
  }
};

// octrees (geometry_oct.cuh)
template <>
struct Geo<GEO_OCT> {
  using Ray = OctRay;
  struct Cross {
    int nb;
  };
  static __device__ __forceinline__ bool find_cell(const ModelDev &M, double rx, double ry, double rz, double vx,
                                                   double vy, double vz, int &ix, int &iy, int &iz, int &ic) {
    ix = iy = iz = 0;
    ic = oct_find_cell(M.oct, rx, ry, rz);
    return ic >= 0;
  }
  static __device__ __forceinline__ void start(const ModelDev &M, Ray &R, double rx, double ry, double rz, double vx,
                                               double vy, double vz, int ix, int iy, int iz, int ic) {
    oct_start(R, rx, ry, rz, vx, vy, vz, ic);
  }
  static __device__ __forceinline__ bool escaped(const ModelDev &M, const Ray &R) { return oct_escaped(M.oct, R); }
  static __device__ __forceinline__ bool find_wall(const ModelDev &M, const Ray &R, double &dt, Cross &c) {
#if OCT_LDG
    const OctNode N = oct_load(M.oct.nodes + R.ic);
#else
    const OctNode N = M.oct.nodes[R.ic];
#endif
    int wall;
    if (!oct_find_wall(M.oct, R, N, dt, wall)) return false;
    c.nb = N.nb[wall];
    return true;
  }
  static __device__ __forceinline__ void step(const ModelDev &M, Ray &R, const Cross &c) {
    if (c.nb < 0) {
      R.ic = M.oct.n_nodes;
      return;
    }
    R.ic = oct_descend(M.oct, c.nb, R.r0x + R.t * R.vx, R.r0y + R.t * R.vy, R.r0z + R.t * R.vz);
  }
  static __device__ __forceinline__ void stop_inside(Ray &R) {}
  static __device__ __forceinline__ void store(const Ray &R, int &ix, int &iy, int &iz, int &ic) {
    ix = iy = iz = 0;
    ic = R.ic;
  }
};

// Voronoi meshes (geometry_vor.cuh)
template <>
struct Geo<GEO_VOR> {
  using Ray = VorRay;
  struct Cross {
    int next;
  };
  static __device__ __forceinline__ bool find_cell(const ModelDev &M, double rx, double ry, double rz, double vx,
                                                   double vy, double vz, int &ix, int &iy, int &iz, int &ic) {
    ix = iy = iz = 0;
    ic = vor_find_cell(M.vor, rx, ry, rz);
    return ic >= 0;
  }
  static __device__ __forceinline__ void start(const ModelDev &M, Ray &R, double rx, double ry, double rz, double vx,
                                               double vy, double vz, int ix, int iy, int iz, int ic) {
    R.r0x = rx; R.r0y = ry; R.r0z = rz;
    R.vx = vx; R.vy = vy; R.vz = vz;
    R.t = 0.0;
    R.ic = ic;
  }
  static __device__ __forceinline__ bool escaped(const ModelDev &M, const Ray &R) { return R.ic == M.vor.n_cells; }
  static __device__ __forceinline__ bool find_wall(const ModelDev &M, const Ray &R, double &dt, Cross &c) {
    return vor_find_wall(M.vor, R, dt, c.next);
  }
  static __device__ __forceinline__ void step(const ModelDev &M, Ray &R, const Cross &c) {
    R.ic = c.next;
  }
  static __device__ __forceinline__ void stop_inside(Ray &R) {}
  static __device__ __forceinline__ void store(const Ray &R, int &ix, int &iy, int &iz, int &ic) {
    ix = iy = iz = 0;
    ic = R.ic;
  }
};

// block-structured AMR (geometry_amr.cuh)
template <>
struct Geo<GEO_AMR> {
  using Ray = AmrRay;
  struct Cross {
    int wall;
  };
  static __device__ __forceinline__ bool find_cell(const ModelDev &M, double rx, double ry, double rz, double vx,
                                                   double vy, double vz, int &ix, int &iy, int &iz, int &ic) {
    int g;
    if (!amr_find_cell(M.amr, rx, ry, rz, g, ix, iy, iz)) return false;
    ic = amr_cell_id(M.amr.grids[g], ix, iy, iz);
    return true;
  }
  static __device__ __forceinline__ void start(const ModelDev &M, Ray &R, double rx, double ry, double rz, double vx,
                                               double vy, double vz, int ix, int iy, int iz, int ic) {
    amr_start(M.amr, R, rx, ry, rz, vx, vy, vz, ix, iy, iz, ic);
  }
  static __device__ __forceinline__ bool escaped(const ModelDev &M, const Ray &R) { return R.ic == -2; }
  static __device__ __forceinline__ bool find_wall(const ModelDev &M, const Ray &R, double &dt, Cross &c) {
    if (R.ic < 0) return false;  // invalid_cell: the re-location after a grid change failed
    amr_find_wall(M.amr.grids[R.g], R, dt, c.wall);
    return true;
  }
  static __device__ __forceinline__ void step(const ModelDev &M, Ray &R, const Cross &c) { amr_step(M.amr, R, c.wall); }
  static __device__ __forceinline__ void stop_inside(Ray &R) {}
  static __device__ __forceinline__ void store(const Ray &R, int &ix, int &iy, int &iz, int &ic) {
    ix = R.i1; iy = R.i2; iz = R.i3; ic = R.ic;
  }
};

// grid_integrate / grid_integrate_noenergy (grid_propagate_3d.f90:35-375) for one packet: walk cell by
// cell until tau_left is used up (MARCH_INTERACT, R.t = path length to the event, R.ic its cell), the
// packet leaves the grid (MARCH_ESCAPED) or no wall is found (MARCH_KILLED).  DEP: deposit
// path length x kappa x energy in every crossed cell.
// t_source: path length at which the flight hits a stellar surface (+inf: never); crossing it ends the march
// with MARCH_REABSORBED before the segment is deposited, as in the reference (grid_propagate_3d.f90:140-146).
template <int GEO, int ND, bool DEP>
__device__ inline int geo_march(const ModelDev &M, typename Geo<GEO>::Ray &R, double &tau_left, const double (&chi)[ND],
                                const double (&kE)[ND], CellRec *__restrict__ cells, uint32_t &n_cross,
                                const double t_source, const unsigned long long pid = 0ull,
                                const bool count_start = true, int max_steps = 0x7fffffff,
                                double *__restrict__ spec = nullptr) {
  using G = Geo<GEO>;
  if (G::escaped(M, R)) return MARCH_ESCAPED;
  // n_photons (grid_propagate_3d.f90:90-95,175-180): a packet counts once per cell for as long as no other packet
  // has entered the cell since; pid = packet id + 1, 0 when no counter is kept.  The reference runs one packet at
  // a time, so its counter is the number of DISTINCT packets per cell; here all packets are in flight together and
  // another packet's visit can separate two visits of the same packet.  A flight that starts where the packet has
  // just interacted is therefore not counted at all (the packet was counted when it entered that cell): what
  // remains are packets that come back to a cell after leaving it while another packet passed through.
  const bool counting = DEP && pid != 0ull && M.n_visits != nullptr;
  if (counting && count_start && R.ic >= 0 && atomicExch(M.last_id + R.ic, pid) != pid) atomicAdd(M.n_visits + R.ic, 1ull);
  // max_steps: crossings after which the march returns 0 (still in flight; call again with count_start = false)
  for (;; --max_steps) {
    if (max_steps <= 0) return 0;
    double dt;
    typename G::Cross cr;
    double rho[ND];
    const int ic = R.ic;
    if (ic < 0) return MARCH_KILLED;  // AMR: the cell behind a grid boundary could not be located
#pragma unroll
    for (int id = 0; id < ND; ++id) rho[id] = DEP ? __ldcg(&cells[(size_t)ic * ND + id].rho) : __ldg(M.rho + (size_t)ic * ND + id);
    if (!G::find_wall(M, R, dt, cr)) return MARCH_KILLED;
    double chi_rho = 0.0;
#pragma unroll
    for (int id = 0; id < ND; ++id) chi_rho += chi[id] * rho[id];
    const double tau_cell = chi_rho * dt;
    ++n_cross;
    if (tau_cell < tau_left) {
      if (R.t + dt > t_source) return MARCH_REABSORBED;
      if (DEP) {
#pragma unroll
        for (int id = 0; id < ND; ++id)
          if (rho[id] > 0.0) atomicAdd(&cells[(size_t)ic * ND + id].esum, dt * kE[id]);
        // the packet's frequency bin of specific_energy_sum_spectrum (grid_propagate_3d.f90:155-158)
        if (spec) {
#pragma unroll
          for (int id = 0; id < ND; ++id)
            if (rho[id] > 0.0) atomicAdd(spec + (size_t)ic * ND + id, dt * kE[id]);
        }
      }
      tau_left -= tau_cell;
      R.t += dt;
      G::step(M, R, cr);
      if (G::escaped(M, R)) return MARCH_ESCAPED;
      if (counting && R.ic >= 0 && atomicExch(M.last_id + R.ic, pid) != pid) atomicAdd(M.n_visits + R.ic, 1ull);
    } else {
      const double len = tau_cell > 0.0 ? dt * (tau_left / tau_cell) : 0.0;   // (a flight of optical depth zero ends at once)
      if (R.t + len > t_source) return MARCH_REABSORBED;
      if (DEP) {
#pragma unroll
        for (int id = 0; id < ND; ++id)
          if (rho[id] > 0.0) atomicAdd(&cells[(size_t)ic * ND + id].esum, len * kE[id]);
        if (spec) {   // grid_propagate_3d.f90:217-225
#pragma unroll
          for (int id = 0; id < ND; ++id)
            if (rho[id] > 0.0) atomicAdd(spec + (size_t)ic * ND + id, len * kE[id]);
        }
      }
      R.t += len;
      tau_left = 0.0;
      G::stop_inside(R);
      return MARCH_INTERACT;
    }
  }
}

// grid_escape_tau / grid_escape_column_density (grid_propagate_3d.f90:377-582) with tmax = huge.
// At most max_steps crossings per call.  Returns 1 once the ray has left the grid, 0 if it is still inside,
// -1 if the packet had to be killed (no wall found).
template <int GEO, int ND, bool COLUMN>
__device__ inline int geo_escape(const ModelDev &M, typename Geo<GEO>::Ray &R, const double (&chi)[ND],
                                 const double *__restrict__ rho_only, double &tau, double (&col)[ND], uint32_t &n_cross,
                                 const int max_steps = 0x7fffffff, const double tmax = 1.7976931348623157e308) {
  using G = Geo<GEO>;
  for (int step = 0; step < max_steps; ++step) {
    if (G::escaped(M, R)) return 1;
    double dt;
    typename G::Cross cr;
    double rho[ND];
    if (R.ic < 0) return -1;
#pragma unroll
    for (int id = 0; id < ND; ++id) rho[id] = __ldg(rho_only + (size_t)R.ic * ND + id);
    if (!G::find_wall(M, R, dt, cr)) return -1;
    ++n_cross;
    if (R.t + dt > tmax) {
      // the ray ends inside this cell (inside observers, grid_propagate_3d.f90:440-443)
      dt = tmax - R.t;
#pragma unroll
      for (int id = 0; id < ND; ++id) {
        if (COLUMN)
          col[id] = col[id] + rho[id] * dt;
        else
          tau = tau + chi[id] * rho[id] * dt;
      }
      R.t = tmax;
      return 1;
    }
#pragma unroll
    for (int id = 0; id < ND; ++id) {
      if (COLUMN)
        col[id] = col[id] + rho[id] * dt;
      else
        tau = tau + chi[id] * rho[id] * dt;
    }
    R.t += dt;
    G::step(M, R, cr);
  }
  return G::escaped(M, R) ? 1 : 0;
}
