// flight_geo.cuh -- flight kernel for the geometries of march_geo.cuh (Lucy and imaging iterations).
// Included by hyperion_b200.cu after imaging.cuh (it uses FinalArgs).
#pragma once

constexpr int SPH_FLIGHT_THREADS = 128;
// Resident blocks per SM the register allocation of the spherical / cylindrical march is made for: its find_wall is
// 470 instructions of fp64 arithmetic per crossing and the kernel is bound by instruction issue at the 4 blocks (16
// warps) that 110 registers allow; 6 / 7 / 8 blocks (80 / 72 / 64 registers, a few spills) measure 16 / 21 / 22 % faster on
// the c3 disk.  The
// other geometries wait for dependent loads and lose 3-4 % with the same bound, so they keep their registers.
#ifndef GEO_TREE_MIN_BLOCKS
#define GEO_TREE_MIN_BLOCKS 1
#endif
#ifndef GEO_SPH_MIN_BLOCKS
#define GEO_SPH_MIN_BLOCKS 8
#endif

// grid_integrate (DEP) / grid_integrate_noenergy for every queued packet; persistent threads, each lane
// refills from the queue on its own.  FINAL: a packet on its first flight with a forced first interaction
// measures its optical depth to the grid edge first (iter_final.f90:191-209).
template <int GEO, int ND, bool DEP, bool FINAL>
__global__ void __launch_bounds__(SPH_FLIGHT_THREADS, GEO == GEO_SPH ? GEO_SPH_MIN_BLOCKS : (GEO == GEO_CAR ? 1 : GEO_TREE_MIN_BLOCKS))
flight_geo_kernel(const ModelDev M, Pool P, const FinalArgs F, const uint32_t *__restrict__ q_flight,
                  const uint32_t *n_flight_ptr, uint32_t *cursor, const uint32_t iteration) {
  using G = Geo<GEO>;
  const uint32_t n_flight = *n_flight_ptr;
  Slot<ND> *slots = (Slot<ND> *)P.slots;
  CellRec *__restrict__ cells = M.cells;
  const unsigned lane = threadIdx.x & 31;
  uint32_t n_cross = 0, n_esc = 0, n_peel_cross = 0, n_killed = 0;
  for (;;) {
    // the warp claims 32 queue entries at a time
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(cursor, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= n_flight) break;
    const uint32_t idx = base + lane;
    int fin = 0;
    uint32_t slot = 0;
    if (idx < n_flight) {
      slot = q_flight[idx];
      Slot<ND> *s = slots + slot;
      typename G::Ray R;
      G::start(M, R, s->r0x, s->r0y, s->r0z, s->vx, s->vy, s->vz, s->ix, s->iy, s->iz, s->ic);
      double tau = s->tau_left;
      double chi[ND], kE[ND];
#pragma unroll
      for (int k = 0; k < ND; ++k) {
        chi[k] = s->chi[k];
        kE[k] = s->kE[k];
      }
      // find_nearest_source (source.f90:206-227): where, if anywhere, this flight hits a stellar surface
      int src_hit = -1;
      const double t_source = nearest_source(M, s->r0x, s->r0y, s->r0z, s->vx, s->vy, s->vz, src_hit);
      if (FINAL && tau < 0.0 && !G::escaped(M, R)) {
        typename G::Ray E = R;
        double tau_escape = 0.0, col[ND];
        // grid_escape_tau gives up (killed) when a source lies on the way out (grid_propagate_3d.f90:410-415)
        const bool ok = src_hit < 0 && geo_escape<GEO, ND, false>(M, E, chi, M.rho, tau_escape, col, n_peel_cross) > 0;
        if (!ok && src_hit < 0) ++n_killed;  // grid_escape_tau killed its copy; the packet itself goes on unforced
        Rng rng;
        rng.init(M.seed, s->id, iteration);
        rng.blk = s->rng_blk;
        rng.has_spare = s->rng_has_spare != 0;
        rng.spare = s->rng_spare;
        if (ok && tau_escape > 1.e-10) {
          const double TAU_THRES = 1.e-7;
          const double one_minus_exp = tau_escape > TAU_THRES ? 1.0 - exp(-tau_escape) : tau_escape;
          double weight;
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
        s->rng_blk = rng.blk;
        s->rng_has_spare = rng.has_spare ? 1u : 0u;
        s->rng_spare = rng.spare;
      } else if (FINAL && tau < 0.0) {
        tau = 1.0;  // escaped before the first step: the value is never used
      }
      double *spec = nullptr;
      if (DEP && M.spec_sums) {
        const int b = spectrum_bin(M, s->nu);
        if (b >= 0) spec = M.spec_sums + (size_t)b * (size_t)M.n_cells * ND;
      }
      fin = geo_march<GEO, ND, DEP>(M, R, tau, chi, kE, cells, n_cross, t_source, (DEP && M.n_visits) ? s->id + 1ull : 0ull,
                                    // first flight of a packet, or of its re-emission by a star that absorbed it
                                    s->n_inter == 0u || (s->rng_has_spare >> 1) != 0u, 0x7fffffff, spec);
      if (fin == MARCH_REABSORBED) {
        // hand the packet to the interact kernel, which re-emits it from that source: t < 0 carries the id
        s->t = -(double)(src_hit + 1);
        fin = MARCH_INTERACT | 16;
      }
      if (fin == MARCH_INTERACT) {
        s->t = R.t;
        int ix, iy, iz, ic;
        G::store(R, ix, iy, iz, ic);
        s->ix = ix; s->iy = iy; s->iz = iz; s->ic = ic;
      }
      if (FINAL && fin == MARCH_ESCAPED && F.binned) bin_escaped_packet<ND>(F, s, R.t, M.mono_inu);   // iter_final.f90:126-129
      n_esc += fin == MARCH_ESCAPED ? 1u : 0u;
      n_killed += fin == MARCH_KILLED ? 1u : 0u;
    }
    queue_append((fin & MARCH_INTERACT) != 0, P.q_interact, P.counts + C_NI, slot);
    queue_append(fin == MARCH_ESCAPED || fin == MARCH_KILLED, P.q_emit, P.counts + C_NE, slot);
  }
  warp_add_scalar(M.scalars + SC_CROSS, (double)n_cross);
  warp_add_scalar(M.scalars + SC_ESC, (double)n_esc);
  warp_add_scalar(M.scalars + SC_KILLED_GEO, (double)n_killed);
  if (FINAL) warp_add_scalar(M.scalars + SC_PEEL_CROSS, (double)n_peel_cross);
}
