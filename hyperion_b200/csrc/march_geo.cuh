// march_geo.cuh -- the marches of grid_propagate_3d.f90 for the geometries that are walked one wall search
// at a time (spherical / cylindrical polar grids, octrees).  Each geometry provides the same small
// interface (Geo<GEO>), the marches and kernels are written once.
// Included by hyperion_b200.cu after ModelDev / CellRec are defined and before imaging.cuh.
#pragma once

enum { MARCH_ESCAPED = 1, MARCH_INTERACT = 2, MARCH_KILLED = 4 };

template <int GEO>
struct Geo;

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
    const OctNode N = M.oct.nodes[R.ic];
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
template <int GEO, int ND, bool DEP>
__device__ inline int geo_march(const ModelDev &M, typename Geo<GEO>::Ray &R, double &tau_left, const double (&chi)[ND],
                                const double (&kE)[ND], CellRec *__restrict__ cells, uint32_t &n_cross) {
  using G = Geo<GEO>;
  if (G::escaped(M, R)) return MARCH_ESCAPED;
  for (;;) {
    double dt;
    typename G::Cross cr;
    double rho[ND];
    const int ic = R.ic;
    if (ic < 0) return MARCH_KILLED;  // AMR: the cell behind a grid boundary could not be located
#pragma unroll
    for (int id = 0; id < ND; ++id) rho[id] = __ldcg(&cells[(size_t)ic * ND + id].rho);
    if (!G::find_wall(M, R, dt, cr)) return MARCH_KILLED;
    double chi_rho = 0.0;
#pragma unroll
    for (int id = 0; id < ND; ++id) chi_rho += chi[id] * rho[id];
    const double tau_cell = chi_rho * dt;
    ++n_cross;
    if (tau_cell < tau_left) {
      if (DEP) {
#pragma unroll
        for (int id = 0; id < ND; ++id)
          if (rho[id] > 0.0) atomicAdd(&cells[(size_t)ic * ND + id].esum, dt * kE[id]);
      }
      tau_left -= tau_cell;
      R.t += dt;
      G::step(M, R, cr);
      if (G::escaped(M, R)) return MARCH_ESCAPED;
    } else {
      const double len = dt * (tau_left / tau_cell);
      if (DEP) {
#pragma unroll
        for (int id = 0; id < ND; ++id)
          if (rho[id] > 0.0) atomicAdd(&cells[(size_t)ic * ND + id].esum, len * kE[id]);
      }
      R.t += len;
      tau_left = 0.0;
      G::stop_inside(R);
      return MARCH_INTERACT;
    }
  }
}

// grid_escape_tau / grid_escape_column_density (grid_propagate_3d.f90:377-582) with tmax = huge.
// Returns false if the packet had to be killed (no wall found).
template <int GEO, int ND, bool COLUMN>
__device__ inline bool geo_escape(const ModelDev &M, typename Geo<GEO>::Ray &R, const double (&chi)[ND],
                                  const CellRec *__restrict__ cells, double &tau, double (&col)[ND], uint32_t &n_cross) {
  using G = Geo<GEO>;
  while (!G::escaped(M, R)) {
    double dt;
    typename G::Cross cr;
    double rho[ND];
    if (R.ic < 0) return false;
#pragma unroll
    for (int id = 0; id < ND; ++id) rho[id] = __ldg(&cells[(size_t)R.ic * ND + id].rho);
    if (!G::find_wall(M, R, dt, cr)) return false;
    ++n_cross;
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
  return true;
}
