// march_sph.cuh -- the marches of grid_propagate_3d.f90 on a spherical polar grid.
// Included by hyperion_b200.cu after CellRec is defined and before imaging.cuh.
#pragma once

enum { MARCH_ESCAPED = 1, MARCH_INTERACT = 2, MARCH_KILLED = 4 };

// grid_integrate / grid_integrate_noenergy (grid_propagate_3d.f90:35-375) for one packet: walk cell by
// cell until tau_left is used up (MARCH_INTERACT, R.t = path length to the event, R.ic its cell), the
// packet leaves the grid (MARCH_ESCAPED) or no wall is found (MARCH_KILLED).  DEP: deposit
// path length x kappa x energy in every crossed cell.
template <int ND, bool DEP>
__device__ inline int sph_march(const SphGrid &G, SphRay &R, double &tau_left, const double (&chi)[ND],
                                const double (&kE)[ND], CellRec *__restrict__ cells, uint32_t &n_cross) {
  if (sph_escaped(G, R)) return MARCH_ESCAPED;
  for (;;) {
    double dt;
    int d1, d2, d3;
    double rho[ND];
#pragma unroll
    for (int id = 0; id < ND; ++id) rho[id] = __ldcg(&cells[(size_t)R.ic * ND + id].rho);
    if (!sph_find_wall(G, R, dt, d1, d2, d3)) return MARCH_KILLED;
    double chi_rho = 0.0;
#pragma unroll
    for (int id = 0; id < ND; ++id) chi_rho += chi[id] * rho[id];
    const double tau_cell = chi_rho * dt;
    ++n_cross;
    if (tau_cell < tau_left) {
      if (DEP) {
#pragma unroll
        for (int id = 0; id < ND; ++id)
          if (rho[id] > 0.0) atomicAdd(&cells[(size_t)R.ic * ND + id].esum, dt * kE[id]);
      }
      tau_left -= tau_cell;
      R.t += dt;
      sph_step(G, R, d1, d2, d3);
      if (sph_escaped(G, R)) return MARCH_ESCAPED;
    } else {
      const double len = dt * (tau_left / tau_cell);
      if (DEP) {
#pragma unroll
        for (int id = 0; id < ND; ++id)
          if (rho[id] > 0.0) atomicAdd(&cells[(size_t)R.ic * ND + id].esum, len * kE[id]);
      }
      R.t += len;
      tau_left = 0.0;
      R.ow1 = R.ow2 = R.ow3 = 0;
      return MARCH_INTERACT;
    }
  }
}

// grid_escape_tau / grid_escape_column_density (grid_propagate_3d.f90:377-582) with tmax = huge.
// Returns false if the packet had to be killed (no wall found).
template <int ND, bool COLUMN>
__device__ inline bool sph_escape(const SphGrid &G, SphRay &R, const double (&chi)[ND], const CellRec *__restrict__ cells,
                                  double &tau, double (&col)[ND], uint32_t &n_cross) {
  while (!sph_escaped(G, R)) {
    double dt;
    int d1, d2, d3;
    double rho[ND];
#pragma unroll
    for (int id = 0; id < ND; ++id) rho[id] = __ldg(&cells[(size_t)R.ic * ND + id].rho);
    if (!sph_find_wall(G, R, dt, d1, d2, d3)) return false;
    ++n_cross;
#pragma unroll
    for (int id = 0; id < ND; ++id) {
      if (COLUMN)
        col[id] = col[id] + rho[id] * dt;
      else
        tau = tau + chi[id] * rho[id] * dt;
    }
    R.t += dt;
    sph_step(G, R, d1, d2, d3);
  }
  return true;
}
