// Tile-staged flights on Cartesian grids: grid_integrate (src/grid/grid_propagate_3d.f90:35-234) with the
// density read from, and the specific_energy_sum deposits accumulated in, SHARED MEMORY.
//
// Why: a flight that deposits straight into HBM-resident cell records issues one scattered RED.ADD.F64
// and one scattered density load per crossing; measured (profiles/r01_experiments.md) that path saturates
// the SM -> L2 fabric at 40-65 G crossings/s whether or not the records are L2-resident.  Here the grid is
// cut into tiles of 16 x 16 x 16 cells (fewer for several dust types, so that density + sums of a tile fit
// in 64 KB).  Once per round the pending flights are bucketed by the tile they sit in (histogram ->
// scan -> scatter, three small kernels); a block then takes one (tile, chunk of <= 4096 packets) work item,
// copies the tile's densities into shared memory with coalesced loads, marches its packets until they
// interact, leave the grid or step out of the tile (they are then "parked": path length, cell and optical
// depth left go back to the slot and the packet joins the list of the next round), and finally adds the
// tile's sums to the HBM grid with one coalesced RED per touched cell.  Per crossing the kernel touches
// no global memory at all.
//
// A parked flight resumes with the same origin, direction and walls, so the crossing sequence, every
// path segment and every deposit are bit-identical to the untiled march (flight_kernel); only the
// order in which the deposits are summed differs.
#pragma once

#ifndef TILE_THREADS_N
#define TILE_THREADS_N 384
#endif
#ifndef TILE_MIN_BLOCKS_N
#define TILE_MIN_BLOCKS_N 2
#endif
constexpr int TILE_THREADS = TILE_THREADS_N;
constexpr int TILE_MIN_BLOCKS = TILE_MIN_BLOCKS_N;
constexpr uint32_t TILE_CHUNK = 4096;   // packets per work item
#ifndef TILE_REFILL_IDLE
#define TILE_REFILL_IDLE 8              // idle lanes of a warp that trigger a refill
#endif

// ---- bucketing of the pending flights by tile -------------------------------------------------
__global__ void tile_hist_kernel(Pool P, const int buf) {
  const uint32_t n = P.counts[C_NP0 + buf];
  const unsigned lane = threadIdx.x & 31;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
    const uint32_t i = base + lane;
    const uint32_t t = i < n ? P.tile.park_tile[buf][i] : 0xffffff00u + lane;
    const unsigned m = __match_any_sync(0xffffffffu, t);
    if (i < n && (int)lane == __ffs(m) - 1) atomicAdd(P.tile.tile_count + t, (uint32_t)__popc(m));
  }
}

// One block: offsets of every tile's packets in the sorted list and the (tile, chunk) work items.
__global__ void __launch_bounds__(1024) tile_scan_kernel(Pool P) {
  typedef cub::BlockScan<uint32_t, 1024> Scan;
  __shared__ typename Scan::TempStorage tmp_a, tmp_b;
  uint32_t carry_off = 0, carry_items = 0;
  const int nt = P.tile.n_tiles;
  for (int base = 0; base < nt; base += 1024) {
    const int t = base + (int)threadIdx.x;
    const uint32_t cnt = t < nt ? P.tile.tile_count[t] : 0u;
    const uint32_t nch = (cnt + TILE_CHUNK - 1) / TILE_CHUNK;
    uint32_t off, ioff, tot_off, tot_items;
    Scan(tmp_a).ExclusiveSum(cnt, off, tot_off);
    Scan(tmp_b).ExclusiveSum(nch, ioff, tot_items);
    off += carry_off;
    ioff += carry_items;
    if (t < nt) {
      P.tile.tile_cursor[t] = off;
      P.tile.tile_count[t] = 0;  // ready for the next round's histogram
      for (uint32_t k = 0; k < nch; ++k)
        P.tile.items[ioff + k] = make_uint4((uint32_t)t, off + k * TILE_CHUNK, min(TILE_CHUNK, cnt - k * TILE_CHUNK), 0u);
    }
    carry_off += tot_off;
    carry_items += tot_items;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    P.counts[C_NITEMS] = carry_items;
    P.counts[C_ITEM_CURSOR] = 0;
  }
}

__global__ void tile_scatter_kernel(Pool P, const int buf) {
  const uint32_t n = P.counts[C_NP0 + buf];
  const unsigned lane = threadIdx.x & 31;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
    const uint32_t i = base + lane;
    const uint32_t t = i < n ? P.tile.park_tile[buf][i] : 0xffffff00u + lane;
    const unsigned m = __match_any_sync(0xffffffffu, t);
    const int leader = __ffs(m) - 1;
    uint32_t pos = 0;
    if (i < n && (int)lane == leader) pos = atomicAdd(P.tile.tile_cursor + t, (uint32_t)__popc(m));
    pos = __shfl_sync(0xffffffffu, pos, leader);
    if (i < n) P.tile.sorted[pos + __popc(m & ((1u << lane) - 1u))] = P.tile.park_slot[buf][i];
  }
}

// ---- the march ---------------------------------------------------------------------------------
// State of one flight inside a tile.  Cell indices are local to the tile.
template <int ND>
struct TileLane {
  double r0x, r0y, r0z, ivx, ivy, ivz;
  double t, tnx, tny, tnz, tau;
  double chi[ND], kE[ND];
  int lx, ly, lz;
  int first_ic;  // >= 0: the 1-D cell id find_cell gave a packet placed on a wall; it overrides (ix, iy, iz)
                 // for the first segment (grid_geometry_cartesian_3d.f90:184-232) and is served from HBM
};

template <int ND>
__global__ void __launch_bounds__(TILE_THREADS, TILE_MIN_BLOCKS)
flight_tile_kernel(const ModelDev M, Pool P, const int park_buf) {
  using TD = TileDims<ND>;
  extern __shared__ double t_smem[];
  double *__restrict__ s_rho = t_smem;
  double *__restrict__ s_esum = t_smem + TD::CELLS * ND;
  constexpr int TW = TD::WALLS;                        // wall slots per axis
  double *__restrict__ s_w = s_esum + TD::CELLS * ND;  // [3][TW] walls of the tile
  constexpr int REC16 = (80 + 16 * ND) / 16;            // hot part of a Slot<ND> in 16-byte units
  uint4 *__restrict__ s_rec = (uint4 *)(s_w + 3 * TW + 1);   // [TILE_THREADS][REC16], 16-byte aligned
  constexpr uint32_t NO_PACKET = 0xffffffffu;
  static_assert(offsetof(Slot<ND>, ix) == 64 + 16 * ND, "hot part of Slot is contiguous");
  __shared__ uint32_t s_item, s_next;
  const int n1 = M.n1, n2 = M.n2, n3 = M.n3;
  const uint32_t n_items = P.counts[C_NITEMS];
  Slot<ND> *slots = (Slot<ND> *)P.slots;
  CellRec *__restrict__ cells = M.cells;
  const unsigned lane = threadIdx.x & 31;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  uint32_t n_cross = 0, n_esc = 0;
  unsigned long long cross_hi = 0;

  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(P.counts + C_ITEM_CURSOR, 1u);
    __syncthreads();
    const uint32_t item = s_item;
    if (item >= n_items) break;
    const uint4 it = P.tile.items[item];
    const int tx = (int)it.x % P.tile.ntx, ty = ((int)it.x / P.tile.ntx) % P.tile.nty, tz = (int)it.x / (P.tile.ntx * P.tile.nty);
    const int x0 = tx * TD::X, y0 = ty * TD::Y, z0 = tz * TD::Z;
    // ---------------- stage the tile: densities in, sums zeroed ----------------
    for (int c = threadIdx.x; c < TD::CELLS; c += TILE_THREADS) {
      const int gx = x0 + (c % TD::X), gy = y0 + (c / TD::X) % TD::Y, gz = z0 + c / (TD::X * TD::Y);
      const bool inside = gx < n1 && gy < n2 && gz < n3;
      const size_t g = ((size_t)((size_t)gz * n2 + gy) * n1 + gx) * ND;
#pragma unroll
      for (int id = 0; id < ND; ++id) {
        s_rho[c * ND + id] = inside ? __ldcg(&cells[g + id].rho) : 0.0;
        s_esum[c * ND + id] = 0.0;
      }
    }
    for (int k = threadIdx.x; k < 3 * TW; k += TILE_THREADS) {
      const int a = k / TW, j = k % TW;
      const int o = a == 0 ? 0 : (a == 1 ? n1 + 1 : n1 + n2 + 2);
      const int na = a == 0 ? n1 : (a == 1 ? n2 : n3);
      const int i0 = a == 0 ? x0 : (a == 1 ? y0 : z0);
      s_w[k] = M.w1[o + min(i0 + j, na)];
    }
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();

    // ---------------- march the packets of the work item ----------------
    // Every lane owns one record in shared memory into which the hot part of its NEXT packet is
    // copied with cp.async while the lane still marches the current one, so that the HBM latency of
    // the (randomly placed) slots is not exposed at every hand-over.
    bool exhausted = false;  // warp-uniform: the chunk has no unclaimed packet left
    int fin = 3;  // 0 in flight, 1 left the grid, 2 interaction, 3 no packet, 4 stepped out of the tile
    uint32_t slot = 0, pend = NO_PACKET;
    TileLane<ND> L;
    L.lx = L.ly = L.lz = 0;
    L.first_ic = -1;
    for (;;) {
      // -------- hand over finished packets, start the prefetched ones (all conditions warp-uniform) --------
      const int n_active = __popc(__ballot_sync(0xffffffffu, fin == 0));
      const int n_ready = __popc(__ballot_sync(0xffffffffu, fin != 0 && pend != NO_PACKET));
      if (n_active == 0 || n_ready >= TILE_REFILL_IDLE || (!exhausted && n_ready == 0 && n_active < 32)) {
        const int gx = x0 + L.lx, gy = y0 + L.ly, gz = z0 + L.lz;
        if (fin == 2 || fin == 4) {
          const int ic = (fin == 2 && L.first_ic >= 0) ? L.first_ic : (gz * n2 + gy) * n1 + gx;
          __stcs(&slots[slot].t, L.t);
          __stcs((int4 *)&slots[slot].ix, make_int4(gx, gy, gz, ic));
          if (fin == 4) __stcs(&slots[slot].tau_left, L.tau);
        }
        queue_append(fin == 2, P.q_interact, P.counts + C_NI, slot);
        queue_append(fin == 1, P.q_emit, P.counts + C_NE, slot);
        park_append(P, park_buf, fin == 4, slot, tile_of_cell<ND>(P.tile, gx, gy, gz));
        n_esc += fin == 1 ? 1u : 0u;
        if (fin != 0) fin = 3;
        asm volatile("cp.async.wait_all;" ::: "memory");
        if (fin == 3 && pend != NO_PACKET) {
          // start the prefetched packet: record layout = first REC_BYTES of Slot<ND>
          const double *r = (const double *)(s_rec + (size_t)threadIdx.x * REC16);
          const int4 cc = *(const int4 *)(r + 8 + 2 * ND);
          slot = pend;
          pend = NO_PACKET;
          L.r0x = r[0]; L.r0y = r[1]; L.r0z = r[2];
          const double vx = r[3], vy = r[4], vz = r[5];
          L.tau = r[6];
          L.t = r[7];
#pragma unroll
          for (int k = 0; k < ND; ++k) {
            L.chi[k] = r[8 + k];
            L.kE[k] = r[8 + ND + k];
          }
          L.lx = cc.x - x0; L.ly = cc.y - y0; L.lz = cc.z - z0;
          L.first_ic = cc.w != (cc.z * n2 + cc.y) * n1 + cc.x ? cc.w : -1;
          L.ivx = 1.0 / vx; L.ivy = 1.0 / vy; L.ivz = 1.0 / vz;
          fin = 0;
          if ((unsigned)L.lx >= (unsigned)TD::X || (unsigned)L.ly >= (unsigned)TD::Y || (unsigned)L.lz >= (unsigned)TD::Z) {
            // cannot happen for a packet bucketed by its own cell; never index shared memory with it
            atomicCAS(M.error_flag, ERR_NONE, ERR_NOT_IN_CELL);
            L.lx = L.ly = L.lz = 0;
            fin = 3;
          } else {
            // distance to the wall ahead on each axis (init_lane of the untiled kernels)
            L.tnx = vx != 0.0 ? fmax((s_w[L.lx + (vx > 0.0 ? 1 : 0)] - L.r0x) * L.ivx, 0.0) : inf;
            L.tny = vy != 0.0 ? fmax((s_w[TW + L.ly + (vy > 0.0 ? 1 : 0)] - L.r0y) * L.ivy, 0.0) : inf;
            L.tnz = vz != 0.0 ? fmax((s_w[2 * TW + L.lz + (vz > 0.0 ? 1 : 0)] - L.r0z) * L.ivz, 0.0) : inf;
          }
        }
        if (!exhausted) {
          // claim the next packet of every lane without one in its record and start the copy
          const unsigned m_need = __ballot_sync(0xffffffffu, pend == NO_PACKET);
          bool failed = false;
          if (m_need) {
            const int leader = __ffs(m_need) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(&s_next, (uint32_t)__popc(m_need));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (pend == NO_PACKET) {
              const uint32_t idx = base + __popc(m_need & ((1u << lane) - 1u));
              if (idx >= it.z) {
                failed = true;
              } else {
                pend = P.tile.sorted[it.y + idx];
                const char *src = (const char *)(slots + pend);
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_rec + (size_t)threadIdx.x * REC16);
#pragma unroll
                for (int k = 0; k < REC16; ++k)
                  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * k), "l"(src + 16 * k) : "memory");
              }
            }
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
          exhausted = __any_sync(0xffffffffu, failed);
        }
        if (__ballot_sync(0xffffffffu, fin == 0) == 0) {
          if (__ballot_sync(0xffffffffu, pend != NO_PACKET) == 0) break;
          continue;
        }
      }

      // -------- one cell crossing (grid_propagate_3d.f90:106-232) --------
      if (fin == 0) {
        const int c = (L.lz * TD::Y + L.ly) * TD::X + L.lx;
        double rho[ND];
        const bool from_hbm = L.first_ic >= 0;
#pragma unroll
        for (int id = 0; id < ND; ++id)
          rho[id] = from_hbm ? __ldcg(&cells[(size_t)L.first_ic * ND + id].rho) : s_rho[c * ND + id];
        const bool bx = (L.tnx <= L.tny) & (L.tnx <= L.tnz);
        const bool by = (!bx) & (L.tny <= L.tnz);
        const double t_exit = bx ? L.tnx : (by ? L.tny : L.tnz);
        const double iv_ax = bx ? L.ivx : (by ? L.ivy : L.ivz);
        const int fwd = iv_ax > 0.0 ? 1 : 0;
        const double ds = t_exit - L.t;
        double chi_rho = 0.0;
#pragma unroll
        for (int id = 0; id < ND; ++id) chi_rho += L.chi[id] * rho[id];
        const double tau_cell = chi_rho * ds;
        ++n_cross;
        double len;
        if (tau_cell < L.tau) {
          // cross the whole cell: deposit tmin * kappa * E (grid_propagate_3d.f90:148-160)
          len = ds;
          L.tau -= tau_cell;
          L.t = t_exit;
          const int l_old = bx ? L.lx : (by ? L.ly : L.lz);
          const int l_new = l_old + 2 * fwd - 1;
          const int g_new = (bx ? x0 : (by ? y0 : z0)) + l_new;
          const int n_ax = bx ? n1 : (by ? n2 : n3);
          const int t_ax = bx ? TD::X : (by ? TD::Y : TD::Z);
          if ((unsigned)g_new >= (unsigned)n_ax) {
            fin = 1;
          } else {
            L.lx = bx ? l_new : L.lx;
            L.ly = by ? l_new : L.ly;
            L.lz = (bx | by) ? L.lz : l_new;
            if ((unsigned)l_new >= (unsigned)t_ax) {
              fin = 4;
            } else {
              const double wall = s_w[(bx ? 0 : (by ? TW : 2 * TW)) + l_new + fwd];
              const double tn_new = (wall - (bx ? L.r0x : (by ? L.r0y : L.r0z))) * iv_ax;
              L.tnx = bx ? tn_new : L.tnx;
              L.tny = by ? tn_new : L.tny;
              L.tnz = (bx | by) ? L.tnz : tn_new;
            }
          }
        } else {
          // interaction inside this cell (grid_propagate_3d.f90:186-228)
          len = tau_cell > 0.0 ? ds * (L.tau / tau_cell) : 0.0;
          L.t += len;
          fin = 2;
        }
#pragma unroll
        for (int id = 0; id < ND; ++id) {
          if (rho[id] > 0.0) {
            const double dv = len * L.kE[id];
            if (from_hbm) atomicAdd(&cells[(size_t)L.first_ic * ND + id].esum, dv);
            else atomicAdd(&s_esum[c * ND + id], dv);
          }
        }
        if (fin != 2) L.first_ic = -1;
      }
    }
    if (n_cross > 0x7fffff00u) {
      cross_hi += n_cross;
      n_cross = 0;
    }
    __syncthreads();

    // ---------------- add the tile's sums to the grid ----------------
    for (int c = threadIdx.x; c < TD::CELLS; c += TILE_THREADS) {
      const int gx = x0 + (c % TD::X), gy = y0 + (c / TD::X) % TD::Y, gz = z0 + c / (TD::X * TD::Y);
      const size_t g = ((size_t)((size_t)gz * n2 + gy) * n1 + gx) * ND;
#pragma unroll
      for (int id = 0; id < ND; ++id) {
        const double v = s_esum[c * ND + id];
        if (v != 0.0) atomicAdd(&cells[g + id].esum, v);
      }
    }
    __syncthreads();
  }
  warp_add_scalar(M.scalars + SC_CROSS, (double)(cross_hi + n_cross));
  warp_add_scalar(M.scalars + SC_ESC, (double)n_esc);
}
