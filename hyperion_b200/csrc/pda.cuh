// pda.cuh -- the partial diffusion approximation on the device (solve_pda, src/grid/grid_pda_3d.f90:105-325;
// geometry factors grid_pda_{cartesian,spherical,cylindrical}_3d.f90) and the per-cell packet counter it needs
// (n_photons / last_photon_id, src/grid/grid_propagate_3d.f90:90-95,175-180).
//
// The reference solves the linear system  sum_w c_w (e_mean(next_w) - e_mean(cell)) = 0  over the badly sampled
// cells with a dense Gauss elimination (< 10000 cells) or Gauss-Seidel sweeps in cell order stopped at a relative
// change of 1e-4 per sweep, and repeats it until the specific energy (on which the Rosseland opacities in c_w
// depend) changes by less than 1e-5 / 1e-4.  Both inner solvers are sequential.  Here the same system is relaxed
// with Jacobi sweeps (one thread per PDA cell, two e_mean buffers) down to a relative change of 1e-9 per sweep, i.e.
// to the solution the exact solver returns; the outer loop and its tolerances are the reference's.
#pragma once

// Regular grids (Cartesian, spherical polar, cylindrical polar): cell_width(cell, d) = W[d][0][i1] * W[d][1][i2] *
// W[d][2][i3] and geometrical_factor(wall, cell) = F[wall][i1 or i2]  (tables made by the host, pda_geometry()).
struct PdaGeo {
  const double *W[3][3];
  const double *F[4];        // walls 1, 2 by i1; walls 3, 4 by i2; walls 5, 6 have factor 1
  int32_t n1, n2, n3, n_dim, periodic3;
};

struct PdaDev {
  PdaGeo G;
  const double *counts;      // [n_cells] n_photons, summed over the processes
  double limit;              // max(30, ceiling(0.005 x mean n_photons))
  int32_t *list;             // ids of the PDA cells
  uint32_t *n_list;
  double *e_a, *e_b;         // [n_cells] e_mean, two buffers of the Jacobi sweeps
  double *coef;              // [n_list][6]
  unsigned long long *maxdiff;  // bits of the largest relative change (non-negative doubles order like integers)
};

__device__ __forceinline__ double pda_width(const PdaGeo &G, int d, int i1, int i2, int i3) {
  return __ldg(G.W[d][0] + i1) * __ldg(G.W[d][1] + i2) * __ldg(G.W[d][2] + i3);
}

// neighbour across wall w (0..5) of cell (i1, i2, i3); the polar grids wrap in phi (next_cell_int)
__device__ __forceinline__ void pda_next(const PdaGeo &G, int w, int &i1, int &i2, int &i3) {
  switch (w) {
    case 0: --i1; break;
    case 1: ++i1; break;
    case 2: --i2; break;
    case 3: ++i2; break;
    case 4:
      --i3;
      if (G.periodic3 && i3 < 0) i3 = G.n3 - 1;
      break;
    default:
      ++i3;
      if (G.periodic3 && i3 == G.n3) i3 = 0;
  }
}

__global__ void pda_sum_counts_kernel(const double *__restrict__ counts, int64_t n, double *out) {
  double acc = 0.0;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) acc += counts[k];
  warp_add_scalar(out, acc);
}

// do_pda = n_photons < limit .and. sum(density) > 0, minus the cells on the edge of the grid (check_allowed_pda)
__global__ void pda_mark_kernel(const ModelDev M, const PdaDev P) {
  const PdaGeo &G = P.G;
  const int nd = M.n_dust;
  for (int64_t ic = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ic < M.n_cells; ic += (int64_t)gridDim.x * blockDim.x) {
    const int i1 = (int)(ic % G.n1), i2 = (int)((ic / G.n1) % G.n2), i3 = (int)(ic / ((int64_t)G.n1 * G.n2));
    double rs = 0.0;
    for (int id = 0; id < nd; ++id) rs += M.cells[(size_t)ic * nd + id].rho;
    bool on = P.counts[ic] < P.limit && rs > 0.0;
    if (i1 == 0 || i1 == G.n1 - 1 || i2 == 0 || i2 == G.n2 - 1) on = false;
    if (!G.periodic3 && (i3 == 0 || i3 == G.n3 - 1)) on = false;
    if (on) P.list[atomicAdd(P.n_list, 1u)] = (int32_t)ic;
  }
}

// update_e_mean (grid_pda_3d.f90:92-103) for every cell, into both buffers
__global__ void pda_emean_kernel(const ModelDev M, const PdaDev P) {
  const int nd = M.n_dust;
  for (int64_t ic = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ic < M.n_cells; ic += (int64_t)gridDim.x * blockDim.x) {
    double rs = 0.0, e = 0.0;
    for (int id = 0; id < nd; ++id) {
      const size_t k = (size_t)ic * nd + id;
      const double rho = M.cells[k].rho, s = M.specific_energy[k];
      rs += rho;
      if (rho > 0.0) e += rho * s / mean_opacity_loglog(M.dust[id], M.dust[id].L.o_logkap_planck, s);
    }
    e = rs > 0.0 ? e / rs : 0.0;
    P.e_a[ic] = e;
    P.e_b[ic] = e;
  }
}

// dtau_rosseland (grid_pda_3d.f90:171-181)
__device__ inline double pda_dtau(const ModelDev &M, const PdaGeo &G, int i1, int i2, int i3, int d) {
  const int64_t ic = ((int64_t)i3 * G.n2 + i2) * G.n1 + i1;
  const double w = pda_width(G, d, i1, i2, i3);
  double t = 0.0;
  for (int id = 0; id < M.n_dust; ++id) {
    const size_t k = (size_t)ic * M.n_dust + id;
    const double rho = M.cells[k].rho;
    if (rho > 0.0) t += rho * mean_opacity_loglog(M.dust[id], M.dust[id].L.o_logchi_ross, M.specific_energy[k]) * w;
  }
  return t;
}

// the coefficients of the walls of every PDA cell (grid_pda_3d.f90:209-226)
__global__ void pda_coef_kernel(const ModelDev M, const PdaDev P) {
  const PdaGeo &G = P.G;
  const uint32_t n = *P.n_list;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
    const int64_t ic = P.list[q];
    const int i1 = (int)(ic % G.n1), i2 = (int)((ic / G.n1) % G.n2), i3 = (int)(ic / ((int64_t)G.n1 * G.n2));
    for (int w = 0; w < 6; ++w) {
      double c = 0.0;
      if (w < 2 * G.n_dim) {
        const int d = w >> 1;
        int j1 = i1, j2 = i2, j3 = i3;
        pda_next(G, w, j1, j2, j3);
        double dtau_sum = pda_dtau(M, G, i1, i2, i3, d) + pda_dtau(M, G, j1, j2, j3, d);
        if (dtau_sum < 1e-100) dtau_sum = 1e-100;
        c = 1.0 / dtau_sum / pda_width(G, d, i1, i2, i3);
        if (w < 4) c = c * __ldg(G.F[w] + (w < 2 ? i1 : i2));
      }
      P.coef[(size_t)q * 6 + w] = c;
    }
  }
}

// one Jacobi sweep: e_new = sum_w c_w e(next_w) / sum_w c_w  (grid_pda_3d.f90:283-310, with the neighbours'
// values of the previous sweep)
__global__ void pda_sweep_kernel(const PdaDev P, const double *__restrict__ src, double *__restrict__ dst, const int track) {
  const PdaGeo &G = P.G;
  const uint32_t n = *P.n_list;
  double worst = 0.0;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
    const int64_t ic = P.list[q];
    const int i1 = (int)(ic % G.n1), i2 = (int)((ic / G.n1) % G.n2), i3 = (int)(ic / ((int64_t)G.n1 * G.n2));
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 2 * G.n_dim; ++w) {
      int j1 = i1, j2 = i2, j3 = i3;
      pda_next(G, w, j1, j2, j3);
      const double c = P.coef[(size_t)q * 6 + w];
      a += c;
      b += c * src[((int64_t)j3 * G.n2 + j2) * G.n1 + j1];
    }
    const double e_old = src[ic], e_new = b / a;
    dst[ic] = e_new;
    if (track) worst = fmax(worst, e_old != 0.0 ? fabs(e_new - e_old) / fabs(e_old) : (e_new != 0.0 ? 1.0 : 0.0));
  }
  if (track) {
    for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    if ((threadIdx.x & 31) == 0 && worst > 0.0) atomicMax(P.maxdiff, (unsigned long long)__double_as_longlong(worst));
  }
}

// update_specific_energy (grid_pda_3d.f90:52-90) for the PDA cells; the largest relative change goes to maxdiff
__global__ void pda_update_energy_kernel(const ModelDev M, const PdaDev P, const double *__restrict__ e_mean) {
  const uint32_t n = *P.n_list;
  const int nd = M.n_dust;
  double worst = 0.0;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
    const int64_t ic = P.list[q];
    const double em = e_mean[ic];
    for (int id = 0; id < nd; ++id) {
      const DustDev &d = M.dust[id];
      const size_t k = (size_t)ic * nd + id;
      const double s_old = M.specific_energy[k];
      double s = s_old;
      const double smin = d.L.e_min, smax = d.L.e_max;
      if (em < smin / mean_opacity_loglog(d, d.L.o_logkap_planck, smin)) {
        s = smin;
      } else if (em > smax / mean_opacity_loglog(d, d.L.o_logkap_planck, smax)) {
        s = smax;
      } else {
        for (int it = 0; it < 10000; ++it) {
          const double s_prev = s;
          s = em * mean_opacity_loglog(d, d.L.o_logkap_planck, s);
          if (fmax(s / s_prev, s_prev / s) - 1.0 < 1.e-5) break;
        }
      }
      M.specific_energy[k] = s;
      // the spectrum keeps its shape (grid_pda_3d.f90:63-67)
      if (M.spec_sums && s_old > 0.0)
        for (int b = 0; b < M.n_spec_bins; ++b) M.spec_energy[((size_t)b * M.n_cells * nd) + k] *= s / s_old;
      worst = fmax(worst, fabs(s - s_old) / s_old);
    }
  }
  for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
  if ((threadIdx.x & 31) == 0 && worst > 0.0) atomicMax(P.maxdiff, (unsigned long long)__double_as_longlong(worst));
}

// n_photons as doubles behind the scalars of the reduction buffer (one collective sums grid, scalars and counts)
__global__ void pda_counts_to_double_kernel(const unsigned long long *__restrict__ n_visits, int64_t n, double *__restrict__ out) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    out[k] = (double)n_visits[k];
}
