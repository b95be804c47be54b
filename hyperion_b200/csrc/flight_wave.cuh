// Wave engine: the Lucy photon loop (do_lucy, src/main/iter_lucy.f90:119-209) on uniformly spaced Cartesian
// grids as a wavefront of TILE VISITS, with grid_integrate (src/grid/grid_propagate_3d.f90:35-234) reading
// the densities from, and accumulating the specific_energy_sum deposits in, SHARED MEMORY.
//
// Why (profiles/r01_experiments.md): a flight that deposits straight into the HBM-resident cell records
// issues one scattered density load and one scattered RED per crossing; that path saturates at 60-85 G
// crossings/s wherever the records live.  Here the grid is cut into tiles of about 26^3 cells whose
// densities (fp32) and sums (32-bit fixed point, native ATOMS.ADD) fill the 227 KB of shared memory of an
// SM, and a crossing touches no global memory at all.
//
// State machine.  Every slot of the packet pool carries a KEY in `key[slot]`: the tile its packet sits in
// (a flight is pending), WK_INTERACT (the flight reached its interaction) or WK_FREE.  A round is
//   wave_hist / wave_scan / wave_scatter   counting sort of the slot ids by key -> `sorted`, work items
//   wave_tile_kernel      one visit of every pending flight to its tile: march until the packet interacts,
//                         leaves the grid or steps into the next tile; writes the slot's mutable sector
//                         (tau, t, cell) and its next key
//   wave_interact_kernel  interact (src/dust/dust_interact.f90:22-79) for the WK_INTERACT bucket
//   wave_emit_kernel      emit (src/sources/source.f90:100-179) into the WK_FREE bucket while ids remain
// The three march kernels of a round touch disjoint slots and run side by side on three streams.  When few
// packets are left the remaining flights are finished by the direct kernels (flight_kernel / interact_kernel).
//
// Geometry.  Inside a visit the wall-crossing path lengths advance by the constant dt = dx |1/v| of the
// uniformly spaced axis (the wave engine is only used when all three wall arrays are equidistant to 1e-10,
// checked on the host); at the start of every visit they are recomputed from the wall table, so rounding does
// not accumulate over more than one tile.  The tile is stored with a one-cell halo whose density is a
// negative sentinel: a packet that steps out of the tile reads it and stops (-1: next tile, -2: outside the
// grid), no index test per crossing.
//
// Deposits.  len * kappa * E is scaled so that the largest possible single deposit is 2^17 and rounded to an
// integer with ERROR DIFFUSION along the packet's path: the remainder is carried to the packet's next
// crossing, and every visit starts from a remainder drawn uniformly in [0, 1) (a hash of the packet's
// optical depth left, path length and the iteration), so the sum a visit deposits is floor(exact + u): unbiased for deposits of any size, also
// far below one unit.  A work item holds at most 2^14 packets and a packet crosses a cell at most once per
// visit, so the 32-bit sums cannot overflow.  After the item the sums are converted back and added to the
// fp64 grid with one RED per touched cell.
#pragma once

constexpr uint32_t WAVE_MAX_BINS = 11264;       // tiles + 2; the histogram kernels keep one counter per bin in shared memory
constexpr int WAVE_SORT_THREADS = 512;
constexpr int WAVE_SORT_SEG = 8192;             // slots per block of the counting sort
constexpr float WAVE_DEP_MAX = 131072.0f;       // 2^17: fixed-point value of the largest possible deposit
constexpr uint32_t WAVE_CHUNK_MAX = 16384;      // packets per work item (2^14 * 2^17 < 2^32)
#ifndef WAVE_UNROLL
#define WAVE_UNROLL 4                           // crossings between two hand-over votes
#endif
// throughput probes (tools/gpu_r02f.sh): 0 = product path; 1 no deposit; 2 density 1.0f instead of the shared-memory
// load (halo test kept on the loaded word only every crossing's address...); 3 both
#ifndef WAVE_EXPERIMENT
#define WAVE_EXPERIMENT 0
#endif

enum { WC_NITEMS = 0, WC_ITEM_CURSOR, WC_N_FLIGHT, WC_N_INTERACT, WC_INTERACT_START, WC_N_FREE, WC_FREE_START,
       WC_CLAIMED_LO, WC_CLAIMED_HI, WC_N_EMIT, WC_BY_SLOT, WC_COUNT = 16 };

struct WaveQ {
  uint32_t *key;         // [capacity] state of every slot
  uint32_t *key_pos;     // [capacity] the same keys indexed by the slot's POSITION in `sorted`: once emission has
                         // ended the next sort reads them (and the list itself) as two streams instead of gathering
                         // key[slot] through the list, one 32-byte sector per 4-byte key
                         // (ctl[WC_BY_SLOT], set by the scan kernel: 1 while packet ids are left, i.e. while the next
                         // sort still covers all slots and reads key[slot])
  uint32_t *sorted;      // [capacity] slot ids ordered by key (this round's list; the host alternates two buffers)
  uint32_t *bin_count;   // [n_tiles + 2]
  uint32_t *bin_cursor;  // [n_tiles + 2]
  uint4 *items;          // work items {tile, first index in sorted, packets, -}
  uint32_t *ctl;         // [WC_COUNT]
  uint32_t capacity;
  int tx, ty, tz;        // cells of a tile
  int ntx, nty, ntz, n_tiles;
  uint32_t chunk;        // packets per work item
  int refill;            // finished lanes of a warp that trigger a hand-over
  int queue;             // 1: lanes keep two packets queued behind the one they march (ids and records requested early)
  uint32_t emit_max;     // packets emitted per round at most (emission then overlaps the tile visits of later rounds)
  uint32_t iteration;
  double dx, dy, dz;     // cell widths
  double diag;           // longest path through a cell
  double dep_scale[MAX_DUST], dep_inv[MAX_DUST];
};

__device__ __forceinline__ uint32_t wave_tile_of(const WaveQ &W, int ix, int iy, int iz) {
  return (uint32_t)(((iz / W.tz) * W.nty + iy / W.ty) * W.ntx + ix / W.tx);
}

__global__ void wave_init_kernel(WaveQ W, Pool P) {
  const uint32_t k_free = (uint32_t)W.n_tiles + 1u;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < W.capacity; i += gridDim.x * blockDim.x) W.key[i] = k_free;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (uint32_t)W.n_tiles + 2u; i += gridDim.x * blockDim.x)
    W.bin_count[i] = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int k = 0; k < WC_COUNT; ++k) W.ctl[k] = 0;
    for (int k = 0; k < C_COUNT; ++k) P.counts[k] = 0;
    *P.next_photon = 0ull;
  }
}

// ---- counting sort of the slot ids by key ------------------------------------------------------
// The slots to sort are `src[0 .. n_src)`, or all slots 0 .. n_src-1 when src is null.  Once every packet id
// has been claimed a free slot stays free, and the host passes the previous round's list of busy slots.
__global__ void __launch_bounds__(WAVE_SORT_THREADS)
wave_hist_kernel(WaveQ W, const uint32_t *__restrict__ src, const uint32_t n_src) {
  extern __shared__ uint32_t s_cnt[];
  const uint32_t nb = (uint32_t)W.n_tiles + 2u;
  for (uint32_t k = threadIdx.x; k < nb; k += blockDim.x) s_cnt[k] = 0;
  __syncthreads();
  const uint32_t lo = blockIdx.x * (uint32_t)WAVE_SORT_SEG, hi = min(lo + (uint32_t)WAVE_SORT_SEG, n_src);
  const uint32_t *__restrict__ keys = src ? W.key_pos : W.key;
  for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&s_cnt[min(keys[i], nb - 1u)], 1u);
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < nb; k += blockDim.x) {
    const uint32_t v = s_cnt[k];
    if (v) atomicAdd(W.bin_count + k, v);
  }
}

// One block: start of every bin in `sorted`, the work items (full chunks first so that the last blocks to
// finish hold small items), the control words of the round.
__global__ void __launch_bounds__(1024) wave_scan_kernel(WaveQ W, Pool P, const unsigned long long n_photons) {
  typedef cub::BlockScan<uint32_t, 1024> Scan;
  __shared__ typename Scan::TempStorage tmp_a, tmp_b, tmp_c;
  __shared__ uint32_t s_full;
  const int nt = W.n_tiles, nb = nt + 2;
  const uint32_t chunk = W.chunk;
  // total number of full chunks
  uint32_t mine = 0;
  for (int t = threadIdx.x; t < nt; t += 1024) mine += W.bin_count[t] / chunk;
  uint32_t tot;
  Scan(tmp_a).ExclusiveSum(mine, mine, tot);
  if (threadIdx.x == 0) s_full = tot;
  __syncthreads();
  const uint32_t n_full = s_full;
  uint32_t carry_off = 0, carry_full = 0, carry_part = 0;
  for (int base = 0; base < nb; base += 1024) {
    const int t = base + (int)threadIdx.x;
    const uint32_t cnt = t < nb ? W.bin_count[t] : 0u;
    const bool is_tile = t < nt;
    const uint32_t nfull = is_tile ? cnt / chunk : 0u, npart = (is_tile && cnt % chunk) ? 1u : 0u;
    uint32_t off, ifull, ipart, tot_off, tot_full, tot_part;
    Scan(tmp_a).ExclusiveSum(cnt, off, tot_off);
    Scan(tmp_b).ExclusiveSum(nfull, ifull, tot_full);
    Scan(tmp_c).ExclusiveSum(npart, ipart, tot_part);
    off += carry_off;
    ifull += carry_full;
    ipart += carry_part;
    if (t < nb) {
      W.bin_cursor[t] = off;
      W.bin_count[t] = 0;  // ready for the next round's histogram
      for (uint32_t k = 0; k < nfull; ++k) W.items[ifull + k] = make_uint4((uint32_t)t, off + k * chunk, chunk, 0u);
      if (npart) W.items[n_full + ipart] = make_uint4((uint32_t)t, off + nfull * chunk, cnt - nfull * chunk, 0u);
      if (t == nt) {
        W.ctl[WC_N_FLIGHT] = off;
        W.ctl[WC_N_INTERACT] = cnt;
        W.ctl[WC_INTERACT_START] = off;
      }
      if (t == nt + 1) {
        W.ctl[WC_N_FREE] = cnt;
        W.ctl[WC_FREE_START] = off;
      }
    }
    carry_off += tot_off;
    carry_full += tot_full;
    carry_part += tot_part;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    W.ctl[WC_NITEMS] = n_full + carry_part;
    W.ctl[WC_ITEM_CURSOR] = 0;
    const unsigned long long claimed = *P.next_photon;
    W.ctl[WC_CLAIMED_LO] = (uint32_t)claimed;
    W.ctl[WC_CLAIMED_HI] = (uint32_t)(claimed >> 32);
    // the ids of this round's emission are handed out here, in one piece: free slot i of the round emits packet
    // claimed + i (one atomic per warp on one address was 13 % of the emission kernel's stalls)
    const unsigned long long left = n_photons > claimed ? n_photons - claimed : 0ull;
    const uint32_t n_emit = (uint32_t)min((unsigned long long)min(W.ctl[WC_N_FREE], W.emit_max), left);
    W.ctl[WC_N_EMIT] = n_emit;
    W.ctl[WC_BY_SLOT] = left > 0ull ? 1u : 0u;
    *P.next_photon = claimed + n_emit;
  }
}

__global__ void __launch_bounds__(WAVE_SORT_THREADS)
wave_scatter_kernel(WaveQ W, const uint32_t *__restrict__ src, const uint32_t n_src) {
  extern __shared__ uint32_t s_cnt[];   // [nb] counts, then [nb] bases
  const uint32_t nb = (uint32_t)W.n_tiles + 2u;
  uint32_t *s_base = s_cnt + nb;
  for (uint32_t k = threadIdx.x; k < nb; k += blockDim.x) s_cnt[k] = 0;
  __syncthreads();
  const uint32_t lo = blockIdx.x * (uint32_t)WAVE_SORT_SEG, hi = min(lo + (uint32_t)WAVE_SORT_SEG, n_src);
  // one pass over the keys: the shared-memory atomic that counts a slot also returns its rank among the slots
  // of its bin in this block; key and rank wait in registers until the block has claimed its ranges
  constexpr int PER = WAVE_SORT_SEG / WAVE_SORT_THREADS;
  uint32_t slot_r[PER], key_r[PER], rank_r[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const uint32_t i = lo + threadIdx.x + (uint32_t)j * WAVE_SORT_THREADS;
    key_r[j] = 0xffffffffu;
    if (i < hi) {
      slot_r[j] = src ? src[i] : i;
      key_r[j] = min(src ? W.key_pos[i] : W.key[i], nb - 1u);
      rank_r[j] = atomicAdd(&s_cnt[key_r[j]], 1u);
    }
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < nb; k += blockDim.x) {
    const uint32_t v = s_cnt[k];
    s_base[k] = v ? atomicAdd(W.bin_cursor + k, v) : 0u;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < PER; ++j)
    if (key_r[j] != 0xffffffffu) W.sorted[s_base[key_r[j]] + rank_r[j]] = slot_r[j];
}

#ifndef WAVE_SIGNED_STRIDES
#define WAVE_SIGNED_STRIDES 1
#endif
#ifndef WAVE_PRED_TN
#define WAVE_PRED_TN 0
#endif
#ifndef WAVE_TRIM
#define WAVE_TRIM 1
#endif
#ifndef WAVE_MAGIC_FLOOR
#define WAVE_MAGIC_FLOOR 0
#endif

// ---- the march ---------------------------------------------------------------------------------
// State of one flight inside a tile.  c is the shared-window ADDRESS of the packet's cell in the haloed density
// array (base + cell index * 4 * ND); the sums lie SUM_OFF bytes behind the densities.
template <int ND>
struct WaveLane {
  double tnx, tny, tnz;     // path length at which the next x / y / z wall is reached
  double dtx, dty, dtz;     // path length between two walls of an axis (1e300 for a ray parallel to them)
  double t, tau;
  double chi[ND];
  float kEs[ND];            // kappa * E in fixed-point units per length
  float resid[ND];          // error diffusion: what rounding has left over so far, in [0, 1)
  uint32_t c, cd;           // cd: cell whose density and sum the next crossing uses (differs from c only for the
                            // first segment of a packet placed on a wall, grid_geometry_cartesian_3d.f90:184-232)
  // (the direction signs of the packet live in the three top bits of the lane's crossing counter, WAVE_NEG_SHIFT: as
  // a member of their own the compiler spilled them and every crossing began with a load from local memory)
#if WAVE_SIGNED_STRIDES
  int ssx, ssy, ssz;        // bytes to the next cell along the packet's direction on each axis (signed)
#endif
};
constexpr uint32_t WAVE_NEG_SHIFT = 29, WAVE_CROSS_MASK = (1u << WAVE_NEG_SHIFT) - 1u;

// 1 / x to the last ulp or two for normal x != 0 (MUFU.RCP64H + two Newton steps); only used for wall distances
__device__ __forceinline__ double wave_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

constexpr double WAVE_FAR = 1e300;   // "never": wall distance of a ray parallel to the walls of an axis

// One cell crossing (grid_propagate_3d.f90:106-232).  fin: 0 in flight, 1 left the grid, 2 interaction,
// 4 stepped into the next tile.  sx, sy, sz: bytes between neighbouring cells along x, y, z (block-uniform).
template <int ND, uint32_t SUM_OFF>
__device__ __forceinline__ void wave_cross(WaveLane<ND> &L, int &fin, uint32_t &n_cross, const int sx, const int sy,
                                           const int sz) {
  uint32_t rho[ND];   // fp32 bit patterns
  const uint32_t a_rho = L.cd;
#pragma unroll
  for (int id = 0; id < ND; ++id) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(rho[id]) : "r"(a_rho + 4u * id));
  if ((int)rho[0] < 0) {
    // halo: -1.f next tile, -2.f outside the grid
    fin = rho[0] == 0xc0000000u ? 1 : 4;
    return;
  }
  const bool bx = (L.tnx <= L.tny) & (L.tnx <= L.tnz);
  const bool by = (!bx) & (L.tny <= L.tnz);
  const double t_exit = bx ? L.tnx : (by ? L.tny : L.tnz);
  const double ds = t_exit - L.t;
  double chi_rho = 0.0;
#pragma unroll
  for (int id = 0; id < ND; ++id) {
    // float -> double of a non-negative normal number by re-biasing the exponent (0 becomes 2^-127: harmless);
    // integer instructions instead of one more trip through the narrow conversion pipe
#if WAVE_EXPERIMENT == 7
    const double rd = (double)__uint_as_float(rho[id]);
#else
    const double rd = __hiloint2double((int)((rho[id] >> 3) + 0x38000000u), (int)(rho[id] << 29));
#endif
    chi_rho = fma(L.chi[id], rd, chi_rho);
  }
  const double tau_cell = chi_rho * ds;
  ++n_cross;
  double len;
  if (tau_cell < L.tau) {
    // cross the whole cell: deposit tmin * kappa * E (grid_propagate_3d.f90:148-160)
    len = ds;
    L.tau -= tau_cell;
    L.t = t_exit;
#if WAVE_PRED_TN == 2
    // tn += dt on the axis that was crossed: one PREDICATED add per axis (written in PTX: from C the compiler makes
    // an unconditional add and two selects of every axis)
    asm("{ .reg .pred p; setp.ne.s32 p, %1, 0; @p add.f64 %0, %0, %2; }" : "+d"(L.tnx) : "r"((int)bx), "d"(L.dtx));
    asm("{ .reg .pred p; setp.ne.s32 p, %1, 0; @p add.f64 %0, %0, %2; }" : "+d"(L.tny) : "r"((int)by), "d"(L.dty));
    asm("{ .reg .pred p; setp.eq.s32 p, %1, 0; @p add.f64 %0, %0, %2; }" : "+d"(L.tnz) : "r"((int)(bx | by)), "d"(L.dtz));
#elif WAVE_PRED_TN
    // tn += dt on the axis that was crossed: one predicated add per axis
    if (bx) L.tnx += L.dtx;
    if (by) L.tny += L.dty;
    if (!(bx | by)) L.tnz += L.dtz;
#else
    // tn += dt on the axis that was crossed, as a multiply-add with a 0/1 factor (one select per axis)
    const double mx = __hiloint2double(bx ? 0x3ff00000 : 0, 0), my = __hiloint2double(by ? 0x3ff00000 : 0, 0),
                 mz = __hiloint2double((bx | by) ? 0 : 0x3ff00000, 0);
    L.tnx = fma(mx, L.dtx, L.tnx);
    L.tny = fma(my, L.dty, L.tny);
    L.tnz = fma(mz, L.dtz, L.tnz);
#endif
#if WAVE_SIGNED_STRIDES
    L.c += (uint32_t)(bx ? L.ssx : (by ? L.ssy : L.ssz));
#else
    const int mag = bx ? sx : (by ? sy : sz);
    const uint32_t bit = bx ? (1u << WAVE_NEG_SHIFT) : (by ? (2u << WAVE_NEG_SHIFT) : (4u << WAVE_NEG_SHIFT));
    L.c = (n_cross & bit) ? L.c - (uint32_t)mag : L.c + (uint32_t)mag;
#endif
    L.cd = L.c;
  } else {
    // interaction inside this cell (grid_propagate_3d.f90:186-228); cd keeps the cell the packet interacted in.
    // One lane of a warp ends its flight in most steps while the others wait, so this branch is kept to one move:
    // the path length to the interaction point and its deposit are left to the interaction kernel
    // (wave_partial_step), which gets the optical depth left at the entry wall with a minus sign.
#if WAVE_DEFER_PARTIAL
    len = 0.0;
#else
    // the quotient uses the short reciprocal (MUFU + two Newton steps) instead of the full division sequence
    len = tau_cell > 0.0 ? ds * (L.tau * wave_rcp(tau_cell)) : 0.0;
    len = fmin(len, ds);
    L.t += len;
#endif
    fin = 2;
  }
  const float lenf = (float)len;
#pragma unroll
  for (int id = 0; id < ND; ++id) {
    // no deposit where the density is zero (grid_propagate_3d.f90:150): branch-free, a zero path length
    const float x = fmaf(rho[id] != 0u ? lenf : 0.f, L.kEs[id], L.resid[id]);
#if WAVE_MAGIC_FLOOR
    // floor(x) for 0 <= x < 2^23 without the conversion pipe: x + 2^23 rounded DOWN holds floor(x) in its mantissa
    const float xf = __fadd_rd(x, 8388608.0f);
    const uint32_t q = __float_as_uint(xf) - 0x4B000000u;
    L.resid[id] = x - (xf - 8388608.0f);
#else
    const uint32_t q = __float2uint_rd(x);
    L.resid[id] = x - (float)q;
#endif
#if WAVE_EXPERIMENT == 1
    if (q == 0xffffffffu)
#endif
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a_rho + (SUM_OFF + 4u * id)), "r"(q) : "memory");
  }
}

// uniform in [0, 1): 24 bits of a 32-bit mix (the start value of a visit's rounding remainder)
__device__ __forceinline__ float wave_unit_hash(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t h = a * 0x9E3779B1u ^ b * 0x85EBCA77u ^ c * 0xC2B2AE3Du;
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}

// Shared memory of a block: [densities SUM_OFF bytes][sums SUM_OFF bytes][walls of the tile 3 x TW doubles].
// SUM_OFF is a compile-time constant so that the sum of a cell is addressed as [cell + immediate].
// BOUND: the block size the register allocation is made for (1024 -> 64 registers).  Launched with THREADS < BOUND
// the block leaves registers for the interaction / emission blocks of the round, which then run on the same SMs.
// CUBE: edge of a cubic tile known at compile time (the strides between cells are then immediates of the crossing
// loop instead of constant-bank loads and a multiply per crossing); 0: any shape, from WaveQ.
template <int ND, int THREADS, int MINB, uint32_t SUM_OFF, int BOUND = THREADS, int CUBE = 0>
__global__ void __launch_bounds__(BOUND, MINB)
wave_tile_kernel(const ModelDev M, Pool P, const WaveQ W) {
  extern __shared__ __align__(16) unsigned char w_smem[];
  const int TX = CUBE ? CUBE : W.tx, TY = CUBE ? CUBE : W.ty, TZ = CUBE ? CUBE : W.tz;
  const int TXh = TX + 2, TYh = TY + 2, TZh = TZ + 2;
  float *__restrict__ s_rho = (float *)w_smem;                       // [n_h][ND], halo = sentinel
  uint32_t *__restrict__ s_sum = (uint32_t *)(w_smem + SUM_OFF);     // [n_h][ND]
  const int TW = max(TX, max(TY, TZ)) + 1;
  double *__restrict__ s_w = (double *)(w_smem + 2 * SUM_OFF);       // [3][TW] walls of the tile
  // positions in `sorted` of the three packets a lane holds (marching, record requested, id requested)
  uint32_t *__restrict__ s_pos = (uint32_t *)(s_w + 3 * TW) + threadIdx.x;
  constexpr int PQ = THREADS;   // stride between the three entries of a lane
  __shared__ uint32_t s_item, s_next;
  __shared__ unsigned long long s_cross, s_esc;   // work counters of the block
  const int n1 = M.n1, n2 = M.n2, n3 = M.n3;
  const uint32_t n_items = W.ctl[WC_NITEMS];
  const bool by_slot = W.ctl[WC_BY_SLOT] != 0u;
  Slot<ND> *slots = (Slot<ND> *)P.slots;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NWARPS = THREADS / 32;
  if (threadIdx.x == 0) s_cross = s_esc = 0ull;
  const uint32_t k_interact = (uint32_t)W.n_tiles, k_free = (uint32_t)W.n_tiles + 1u;
  const float inv_row = 1.0f / (float)TXh, inv_slab = 1.0f / (float)(TXh * TYh);
  const uint32_t rho_base = (uint32_t)__cvta_generic_to_shared(s_rho);
  constexpr int CB = 4 * ND;   // bytes of one cell in the density array
  const int sx = CB, sy = CB * TXh, sz = CB * TXh * TYh;

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) {
      s_item = atomicAdd(W.ctl + WC_ITEM_CURSOR, 1u);
      s_next = 0;
    }
    __syncthreads();
    const uint32_t item = s_item;
    if (item >= n_items) break;
    const uint4 it = W.items[item];
    const int tix = (int)it.x % W.ntx, tiy = ((int)it.x / W.ntx) % W.nty, tiz = (int)it.x / (W.ntx * W.nty);
    const int x0 = tix * TX, y0 = tiy * TY, z0 = tiz * TZ;
    // ---------------- stage the tile: densities in (fp32), halo sentinels, sums zeroed ----------------
    // four rows per warp at a time, so that their density loads are in flight together
    for (int row0 = warp; row0 < TYh * TZh; row0 += 4 * NWARPS) {
      float v[4][ND];
      bool act[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int row = row0 + j * NWARPS;
        const int hz = row / TYh, hy = row - hz * TYh;
        const int gy = y0 + hy - 1, gz = z0 + hz - 1, hx = lane, gx = x0 + hx - 1;
        act[j] = row < TYh * TZh && hx < TXh;
        const bool in_grid = (unsigned)gy < (unsigned)n2 && (unsigned)gz < (unsigned)n3 && (unsigned)gx < (unsigned)n1;
        const bool in_tile = hy >= 1 && hy <= TY && hz >= 1 && hz <= TZ && hx >= 1 && hx <= TX;
        const size_t g = ((size_t)((size_t)gz * n2 + gy) * n1 + gx) * ND;
#pragma unroll
        for (int id = 0; id < ND; ++id) {
          v[j][id] = in_grid ? -1.f : -2.f;
          if (act[j] && in_grid && in_tile) v[j][id] = fmaxf((float)__ldg(M.rho + g + id), 0.f);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (act[j]) {
          const int c = (row0 + j * NWARPS) * TXh + lane;
#pragma unroll
          for (int id = 0; id < ND; ++id) {
            s_rho[c * ND + id] = v[j][id];
            s_sum[c * ND + id] = 0u;
          }
        }
      }
    }
    if (TXh > 32) {
      // columns 32 .. TXh-1 of wide tiles
      for (int row = warp; row < TYh * TZh; row += NWARPS) {
        const int hz = row / TYh, hy = row - hz * TYh;
        const int gy = y0 + hy - 1, gz = z0 + hz - 1;
        for (int hx = 32 + lane; hx < TXh; hx += 32) {
          const int gx = x0 + hx - 1;
          const bool in_grid = (unsigned)gy < (unsigned)n2 && (unsigned)gz < (unsigned)n3 && (unsigned)gx < (unsigned)n1;
          const bool in_tile = hy >= 1 && hy <= TY && hz >= 1 && hz <= TZ && hx >= 1 && hx <= TX;
          const int c = row * TXh + hx;
          const size_t g = ((size_t)((size_t)gz * n2 + gy) * n1 + gx) * ND;
#pragma unroll
          for (int id = 0; id < ND; ++id) {
            float vv = in_grid ? -1.f : -2.f;
            if (in_grid && in_tile) vv = fmaxf((float)__ldg(M.rho + g + id), 0.f);
            s_rho[c * ND + id] = vv;
            s_sum[c * ND + id] = 0u;
          }
        }
      }
    }
    for (int k = threadIdx.x; k < 3 * TW; k += THREADS) {
      const int a = k / TW, j = k - a * TW;
      const int o = a == 0 ? 0 : (a == 1 ? n1 + 1 : n1 + n2 + 2);
      const int na = a == 0 ? n1 : (a == 1 ? n2 : n3);
      const int i0 = a == 0 ? x0 : (a == 1 ? y0 : z0);
      s_w[k] = M.w1[o + min(i0 + j, na)];
    }
    __syncthreads();

    // ---------------- march the packets of the work item ----------------
    // Every lane keeps two packets queued behind the one it marches: `n2slot` (its slot id is being loaded
    // from the sorted list) and `nslot` (id known, record requested into L2 with a prefetch).  Ids and records
    // are requested one hand-over ahead of their use, so a hand-over does not wait for HBM twice.
    constexpr uint32_t NONE = 0xffffffffu;
    bool exhausted = false;  // warp-uniform: the item has no unclaimed packet left
    int fin = 3;             // 3: no packet in this lane
    uint32_t slot = 0, nslot = NONE, n2slot = NONE;
    // whether nslot / n2slot hold a packet: kept apart from the values, so that no decision of the hand-over waits
    // for the load of an id from the sorted list (it used to: 4 % of the stall samples on the test n2slot != NONE)
    bool has1 = false, has2 = false;
    uint32_t n_cross = 0;
    WaveLane<ND> L;
    L.c = L.cd = rho_base;
    for (;;) {
      const unsigned m_act = __ballot_sync(0xffffffffu, fin == 0);
      const unsigned m_wait = __ballot_sync(0xffffffffu, fin == 1 || fin == 2 || fin == 4 || (fin == 3 && has1));
      if (m_act == 0 || __popc(m_wait) >= W.refill) {
        // -------- hand over the finished packets --------
        {
          const unsigned m_esc = __ballot_sync(0xffffffffu, fin == 1);
          if (m_esc && lane == 0) atomicAdd(&s_esc, (unsigned long long)__popc(m_esc));
        }
        if (fin == 1 || fin == 2 || fin == 4) {
          // cell of the packet from its index in the haloed tile
          const int ci = (int)(L.c - rho_base) / CB;
          // (with a compile-time tile the divisions are by constants: multiply-high and shift)
          const int hz = (CUBE && WAVE_TRIM) ? ci / ((CUBE + 2) * (CUBE + 2)) : (int)(((float)ci + 0.5f) * inv_slab);
          const int rem = ci - hz * TXh * TYh;
          const int hy = (CUBE && WAVE_TRIM) ? rem / (CUBE + 2) : (int)(((float)rem + 0.5f) * inv_row);
          const int hx = rem - hy * TXh;
          const int gx = x0 + hx - 1, gy = y0 + hy - 1, gz = z0 + hz - 1;
          int ic = (gz * n2 + gy) * n1 + gx;
          uint32_t nk;
          if (fin == 2) {
            if (L.cd != L.c) {
              const int di = (int)(L.cd - rho_base) / CB;
              const int dz = (int)(((float)di + 0.5f) * inv_slab);
              const int drem = di - dz * TXh * TYh;
              const int dy = (int)(((float)drem + 0.5f) * inv_row);
              const int dx = drem - dy * TXh;
              ic = ((z0 + dz - 1) * n2 + (y0 + dy - 1)) * n1 + (x0 + dx - 1);
            }
            nk = k_interact;
          } else if (fin == 1) {
            nk = k_free;
          } else {
            const int d = (hx == 0 ? -1 : (hx == TXh - 1 ? 1 : 0)) + W.ntx * (hy == 0 ? -1 : (hy == TYh - 1 ? 1 : 0)) +
                          W.ntx * W.nty * (hz == 0 ? -1 : (hz == TZh - 1 ? 1 : 0));
            nk = it.x + (uint32_t)d;
          }
          Slot<ND> *s = slots + slot;
          __stcs((double2 *)&s->tau_left, make_double2((WAVE_DEFER_PARTIAL && fin == 2) ? -L.tau : L.tau, L.t));
          __stcs((int4 *)&s->ix, make_int4(gx, gy, gz, ic));
          W.key_pos[s_pos[0]] = nk;
          if (by_slot) W.key[slot] = nk;
          fin = 3;
        }
        // While the item has plenty of packets left every lane keeps its queue full; towards the end a lane only
        // claims a packet when it has none, and walks it through the queue at once (three passes), so that the
        // last packets are spread over all lanes.
        const bool plenty = W.queue && *(volatile uint32_t *)&s_next + 2u * THREADS <= it.z;
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
          // -------- start the queued packets --------
          int first_far = -1;   // >= 0: find_cell's cell of a packet placed on a wall, when it lies in ANOTHER tile
          if (fin == 3 && has1) {
            slot = nslot;
            has1 = false;
            s_pos[0] = s_pos[PQ];
            const Slot<ND> *s = slots + slot;
              // ld.global.cs: streaming loads that still go through L1, so that the six 16-byte loads of a record
            // become one or two line fills instead of six L2 requests (measured: 75.5 -> 66.7 ms per step against ld.cg)
#define WAVE_LD __ldcs
            const double2 a0 = WAVE_LD((const double2 *)&s->r0x);  // r0x r0y
            const double2 a1 = WAVE_LD((const double2 *)&s->r0z);  // r0z vx
            const double2 a2 = WAVE_LD((const double2 *)&s->vy);   // vy vz
            const double2 a3 = WAVE_LD((const double2 *)&s->tau_left);  // tau t
            const int4 cc = WAVE_LD((const int4 *)&s->ix);
            bool bad = false;
  #pragma unroll
            for (int k = 0; k < ND; ++k) {
              L.chi[k] = __ldcg(&s->chi[k]);
              L.kEs[k] = __ldcg(&s->kE[k]) * W.dep_scale[k];
              // fixed-point bound of the deposits (see wave_plan): kappa * E above the table maximum cannot happen
              bad |= !(L.kEs[k] * W.diag <= (double)WAVE_DEP_MAX);
              // seeded by what the packet itself carries (optical depth left, path length), not by its slot: the
              // rounding of a visit is then the same whichever slot, pool size or GPU count the run uses
              L.resid[k] = wave_unit_hash((uint32_t)__double2loint(a3.x) ^ (uint32_t)__double2hiint(a3.x),
                                          (uint32_t)__double2loint(a3.y) ^ (uint32_t)__double2hiint(a3.y),
                                          W.iteration * 4u + (uint32_t)k);
            }
            L.tau = a3.x;
            L.t = a3.y;
            const int lx = cc.x - x0, ly = cc.y - y0, lz = cc.z - z0;
            if (bad || (unsigned)lx >= (unsigned)TX || (unsigned)ly >= (unsigned)TY || (unsigned)lz >= (unsigned)TZ) {
              // cannot happen for a packet bucketed by its own cell; never index shared memory with it
              atomicCAS(M.error_flag, ERR_NONE, bad ? ERR_DEPOSIT : ERR_NOT_IN_CELL);
              W.key_pos[s_pos[0]] = k_free;
              W.key[slot] = k_free;
            } else {
              const double vx = a1.y, vy = a2.x, vz = a2.y;
              const double ivx = wave_rcp(vx), ivy = wave_rcp(vy), ivz = wave_rcp(vz);
              // distance to the wall ahead on each axis; a ray parallel to an axis never reaches its walls.
              // Rounding can leave a resumed packet a few ulp past a wall it faces: clamp to its path length.
#if WAVE_TRIM
              // (an infinite distance -- a direction component below 1e-300 -- is as good as WAVE_FAR here: the
              // increments below stay finite, so no 0 x inf can arise in the crossing's multiply-add)
              L.tnx = vx != 0.0 ? fmax((s_w[lx + (vx > 0.0 ? 1 : 0)] - a0.x) * ivx, L.t) : WAVE_FAR;
              L.tny = vy != 0.0 ? fmax((s_w[TW + ly + (vy > 0.0 ? 1 : 0)] - a0.y) * ivy, L.t) : WAVE_FAR;
              L.tnz = vz != 0.0 ? fmax((s_w[2 * TW + lz + (vz > 0.0 ? 1 : 0)] - a1.x) * ivz, L.t) : WAVE_FAR;
#else
              L.tnx = vx != 0.0 ? fmin(fmax((s_w[lx + (vx > 0.0 ? 1 : 0)] - a0.x) * ivx, L.t), WAVE_FAR) : WAVE_FAR;
              L.tny = vy != 0.0 ? fmin(fmax((s_w[TW + ly + (vy > 0.0 ? 1 : 0)] - a0.y) * ivy, L.t), WAVE_FAR) : WAVE_FAR;
              L.tnz = vz != 0.0 ? fmin(fmax((s_w[2 * TW + lz + (vz > 0.0 ? 1 : 0)] - a1.x) * ivz, L.t), WAVE_FAR) : WAVE_FAR;
#endif
              L.dtx = vx != 0.0 ? fmin(W.dx * fabs(ivx), WAVE_FAR) : WAVE_FAR;
              L.dty = vy != 0.0 ? fmin(W.dy * fabs(ivy), WAVE_FAR) : WAVE_FAR;
              L.dtz = vz != 0.0 ? fmin(W.dz * fabs(ivz), WAVE_FAR) : WAVE_FAR;
              n_cross = (n_cross & WAVE_CROSS_MASK) |
                        (((vx > 0.0 ? 0u : 1u) | (vy > 0.0 ? 0u : 2u) | (vz > 0.0 ? 0u : 4u)) << WAVE_NEG_SHIFT);
#if WAVE_SIGNED_STRIDES
              L.ssx = vx > 0.0 ? sx : -sx;
              L.ssy = vy > 0.0 ? sy : -sy;
              L.ssz = vz > 0.0 ? sz : -sz;
#endif
              L.c = rho_base + (uint32_t)((((lz + 1) * TYh + (ly + 1)) * TXh + lx + 1) * CB);
              L.cd = L.c;
              fin = 0;
              if (cc.w != (cc.z * n2 + cc.y) * n1 + cc.x) {
                // first segment of a packet placed on a wall: density and deposit of find_cell's cell
                const int fz = cc.w / (n1 * n2), frem = cc.w - fz * n1 * n2;
                const int fy = frem / n1, fx = frem - fy * n1;
                const int qx = fx - x0, qy = fy - y0, qz = fz - z0;
                if ((unsigned)qx < (unsigned)TX && (unsigned)qy < (unsigned)TY && (unsigned)qz < (unsigned)TZ)
                  L.cd = rho_base + (uint32_t)((((qz + 1) * TYh + (qy + 1)) * TXh + qx + 1) * CB);
                else
                  first_far = cc.w;
              }
            }
          }
          if (__any_sync(0xffffffffu, first_far >= 0)) {
            // find_cell's cell belongs to another tile: that one crossing goes through global memory.  All
            // packets of a point source on a tile boundary take this path, so the deposits of the lanes that
            // share a cell are summed before ONE RED per cell leaves the warp.
            const bool far = first_far >= 0;
            double dep[ND];
  #pragma unroll
            for (int id = 0; id < ND; ++id) dep[id] = 0.0;
            if (far) {
              double rho[ND], chi_rho = 0.0;
  #pragma unroll
              for (int id = 0; id < ND; ++id) {
                rho[id] = __ldg(M.rho + (size_t)first_far * ND + id);
                chi_rho += L.chi[id] * rho[id];
              }
              const bool bx = (L.tnx <= L.tny) & (L.tnx <= L.tnz);
              const bool by = (!bx) & (L.tny <= L.tnz);
              const double t_exit = bx ? L.tnx : (by ? L.tny : L.tnz);
              const double ds = t_exit - L.t;
              const double tau_cell = chi_rho * ds;
              ++n_cross;
              double len;
              if (tau_cell < L.tau) {
                len = ds;
                L.tau -= tau_cell;
                L.t = t_exit;
                if (bx) L.tnx += L.dtx;
                if (by) L.tny += L.dty;
                if (!(bx | by)) L.tnz += L.dtz;
                const int mag = bx ? sx : (by ? sy : sz);
                L.c = ((n_cross >> WAVE_NEG_SHIFT) & (bx ? 1u : (by ? 2u : 4u))) ? L.c - (uint32_t)mag : L.c + (uint32_t)mag;
                L.cd = L.c;
              } else {
                len = tau_cell > 0.0 ? ds * (L.tau / tau_cell) : 0.0;
                L.t += len;
                // the interaction keeps the cell indices and find_cell's 1-D id the slot already holds
                Slot<ND> *sw = slots + slot;
                __stcs((double2 *)&sw->tau_left, make_double2(L.tau, L.t));
                W.key_pos[s_pos[0]] = k_interact;
                W.key[slot] = k_interact;
                fin = 3;
              }
  #pragma unroll
              for (int id = 0; id < ND; ++id) dep[id] = rho[id] > 0.0 ? len * ((double)L.kEs[id] * W.dep_inv[id]) : 0.0;
            }
            unsigned todo = __ballot_sync(0xffffffffu, far);
            while (todo) {
              const int head = __ffs(todo) - 1;
              const int cell = __shfl_sync(0xffffffffu, first_far, head);
              const bool mine = far && first_far == cell;
  #pragma unroll
              for (int id = 0; id < ND; ++id) {
                double v = mine ? dep[id] : 0.0;
  #pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if ((int)lane == head && v != 0.0) atomicAdd(&M.cells[(size_t)cell * ND + id].esum, v);
              }
              todo &= ~__ballot_sync(0xffffffffu, mine);
            }
          }
          // -------- move the queue up: request the record of the packet whose id has arrived --------
          if (!has1 && has2) {
            nslot = n2slot;
            has1 = true;
            has2 = false;
            s_pos[PQ] = s_pos[2 * PQ];
            const char *rec = (const char *)(slots + nslot);
#if WAVE_EXPERIMENT != 6
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rec));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + 95));
#endif
          }
          // -------- claim the packets after those --------
          if (!exhausted) {
            const bool need = plenty ? !has2 : (fin == 3 && !has1 && !has2);
            const unsigned m_need = __ballot_sync(0xffffffffu, need);
            bool failed = false;
            if (m_need) {
              const int leader = __ffs(m_need) - 1;
              uint32_t base = 0;
              if ((int)lane == leader) base = atomicAdd(&s_next, (uint32_t)__popc(m_need));
              base = __shfl_sync(0xffffffffu, base, leader);
              if (need) {
                const uint32_t idx = base + __popc(m_need & ((1u << lane) - 1u));
                if (idx >= it.z) failed = true;
                else {
                  n2slot = __ldcs(W.sorted + it.y + idx);
                  has2 = true;
                  s_pos[2 * PQ] = it.y + idx;
                }
              }
            }
            exhausted = __any_sync(0xffffffffu, failed);
          }
          if (__ballot_sync(0xffffffffu, fin == 3 && (has1 || has2)) == 0) break;
        }
        if (__ballot_sync(0xffffffffu, fin == 0) == 0) {
          if (exhausted && __ballot_sync(0xffffffffu, has1 || has2) == 0) break;
          continue;
        }
      }
      // -------- cell crossings --------
#pragma unroll
      for (int u = 0; u < WAVE_UNROLL; ++u) {
        if (fin == 0) wave_cross<ND, SUM_OFF>(L, fin, n_cross, sx, sy, sz);
      }
    }
    {
      // crossings of the item (a lane makes far fewer than 2^29 per item)
      n_cross &= WAVE_CROSS_MASK;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) n_cross += __shfl_xor_sync(0xffffffffu, n_cross, o);
      if (lane == 0 && n_cross) atomicAdd(&s_cross, (unsigned long long)n_cross);
    }
    __syncthreads();

    // ---------------- add the tile's sums to the grid ----------------
    for (int row = warp; row < TY * TZ; row += NWARPS) {
      const int lz = row / TY, ly = row - lz * TY;
      const int gy = y0 + ly, gz = z0 + lz;
      if (gy >= n2 || gz >= n3) continue;
      for (int lx = lane; lx < TX; lx += 32) {
        const int gx = x0 + lx;
        if (gx >= n1) continue;
        const int c = ((lz + 1) * TYh + (ly + 1)) * TXh + lx + 1;
        const size_t g = ((size_t)((size_t)gz * n2 + gy) * n1 + gx) * ND;
#pragma unroll
        for (int id = 0; id < ND; ++id) {
          const uint32_t v = s_sum[c * ND + id];
          if (v) atomicAdd(&M.cells[g + id].esum, (double)v * W.dep_inv[id]);
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_cross) atomicAdd(M.scalars + SC_CROSS, (double)s_cross);
    if (s_esc) atomicAdd(M.scalars + SC_ESC, (double)s_esc);
  }
}

// ---- interactions and emission of a round --------------------------------------------------------
template <int ND>
__global__ void __launch_bounds__(SERVICE_THREADS, INTERACT_MIN_BLOCKS)
wave_interact_kernel(const ModelDev M, Pool P, const WaveQ W, const uint32_t iteration) {
  const uint32_t n = W.ctl[WC_N_INTERACT], start = W.ctl[WC_INTERACT_START];
  const bool by_slot = W.ctl[WC_BY_SLOT] != 0u;
  Slot<ND> *slots = (Slot<ND> *)P.slots;
  uint32_t n_abs = 0, n_scat = 0, n_kill = 0;
  const uint32_t k_free = (uint32_t)W.n_tiles + 1u;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t slot = W.sorted[start + i];
    Photon<ND> p;
    Rng rng;
    load_photon<ND>(slots + slot, p, rng, M.seed, iteration);
    const uint64_t id = slots[slot].id;
    int dust_id = 0;
    bool scattered = false;
    if (WAVE_DEFER_PARTIAL && p.tau_left < 0.0) wave_partial_step<ND>(M, p);
    bool ok = interact_photon<ND>(M, p, rng, n_abs, n_scat, n_kill, dust_id, scattered) == 0;
    if (ok && M.use_mrw) ok = mrw_loop<ND>(M, p, rng, n_kill);
    uint32_t nk = k_free;
    if (ok) {
      p.tau_left = -log(1.0 - rng.next());
      store_photon<ND>(slots + slot, p, rng, id);
      nk = wave_tile_of(W, min(max(p.ix, 0), M.n1 - 1), min(max(p.iy, 0), M.n2 - 1), min(max(p.iz, 0), M.n3 - 1));
    }
    W.key_pos[start + i] = nk;
    if (by_slot) W.key[slot] = nk;
  }
  warp_add_scalar(M.scalars + SC_ABS, (double)n_abs);
  warp_add_scalar(M.scalars + SC_SCAT, (double)n_scat);
  warp_add_scalar(M.scalars + SC_KILLED_INT, (double)n_kill);
}

template <int ND>
__global__ void __launch_bounds__(SERVICE_THREADS, INTERACT_MIN_BLOCKS)
wave_emit_kernel(const ModelDev M, Pool P, const WaveQ W, const unsigned long long first_id,
                 const unsigned long long n_photons, const uint32_t iteration) {
  const uint32_t n = W.ctl[WC_N_EMIT], start = W.ctl[WC_FREE_START];
  const unsigned long long k_base = (unsigned long long)W.ctl[WC_CLAIMED_LO] | ((unsigned long long)W.ctl[WC_CLAIMED_HI] << 32);
  const unsigned lane = threadIdx.x & 31;
  Slot<ND> *slots = (Slot<ND> *)P.slots;
  double energy_emitted = 0.0;
  uint32_t n_run = 0, n_esc = 0;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
    const uint32_t i = base + lane;
    const bool valid = i < n;
    // packet ids: the scan kernel has reserved [k_base, k_base + n) for this round; packets are emitted in id order
    unsigned long long k = k_base + i;
    if (!valid || k >= n_photons) continue;
    const uint32_t slot = W.sorted[start + i];
    Photon<ND> p;
    Rng rng;
    bool go = true;
    unsigned long long id = 0;
    for (;;) {
      id = first_id + k;
      rng.init(M.seed, id, iteration);
      ++n_run;
      if (!emit_photon<ND>(M, p, rng, energy_emitted)) {
        go = false;  // fatal model error is flagged; the host reports it after the round
        break;
      }
      // a packet emitted on the outer wall moving outwards escapes immediately
      if (p.ix < 0 || p.ix >= M.n1 || p.iy < 0 || p.iy >= M.n2 || p.iz < 0 || p.iz >= M.n3) {
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
      W.key[slot] = wave_tile_of(W, p.ix, p.iy, p.iz);
    }
  }
  warp_add_scalar(M.scalars + SC_ENERGY, energy_emitted);
  warp_add_scalar(M.scalars + SC_PHOTONS, (double)n_run);
  warp_add_scalar(M.scalars + SC_ESC, (double)n_esc);
}

// Hand the packets that are still in the pool to the direct kernels: flights -> q_flight[0], pending
// interactions -> q_interact.
__global__ void wave_handoff_kernel(Pool P, const WaveQ W) {
  const uint32_t nf = W.ctl[WC_N_FLIGHT], ni = W.ctl[WC_N_INTERACT], is = W.ctl[WC_INTERACT_START];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += gridDim.x * blockDim.x) P.q_flight[0][i] = W.sorted[i];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ni; i += gridDim.x * blockDim.x) P.q_interact[i] = W.sorted[is + i];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    P.counts[C_NF0] = nf;
    P.counts[C_NF0 + 1] = 0;
    P.counts[C_NI] = ni;
    P.counts[C_NE] = 0;
    P.counts[C_NB] = 0;
    P.counts[C_CURSOR] = 0;
    P.counts[C_CURSOR_B] = 0;
  }
}
