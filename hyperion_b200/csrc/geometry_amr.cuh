// geometry_amr.cuh -- block-structured AMR traversal on the device.
//
// Restates src/grid/grid_geometry_amr.f90: find_cell_position / find_position_in_grid (:521-573, with the
// edge-tolerant ipos2 :510-519), find_wall (:775-871), next_cell (:599-675: step inside the grid, or follow
// the goto link of the neighbouring (possibly ghost) cell and re-locate the crossing point, nudged by
// eps = half the smallest cell width, in the target grid, descending through covered cells).
// Grids of all levels live in one flat table; a goto entry is the flat index of the target grid + 1
// (0 = none), the recursion of the reference becomes a loop.
#pragma once

namespace hyp {

struct AmrGridDev {
  int32_t n1, n2, n3, start_id;  // start_id: 0-based id of the grid's first cell
  double xmin, xmax, ymin, ymax, zmin, zmax;
  int64_t goto_off;              // offset of this grid's (n1+2)(n2+2)(n3+2) goto block, first index fastest
};

struct AmrGrid {
  const AmrGridDev *grids;
  const int32_t *gotos;
  const int32_t *cell_grid;      // [n_cells] flat grid index of every cell
  const int32_t *valid;          // ids of the cells not covered by a finer grid (geo%mask_map)
  int32_t n_grids, n_level1, n_cells, n_valid;
  double eps;                    // min cell width / 2 (:349)
};

struct AmrRay {
  double r0x, r0y, r0z, vx, vy, vz, ivx, ivy, ivz;
  double t;
  int g, i1, i2, i3, ic;  // flat grid index, 0-based cell in the grid, global cell id; ic = -2 outside, -1 invalid
};

// linspace_dp (fortranlib/src/lib_array.f90:284-302): wall i of n cells
__device__ __forceinline__ double amr_wall(double a, double b, int i, int n) { return (b - a) * (double)i / (double)n + a; }

// ipos_dp + ipos2: 1-based bin, 0 / n+1 outside
__device__ __forceinline__ int amr_ipos2(double xmin, double xmax, double x, int nbin) {
  int i;
  if (x < xmin) i = 0;
  else if (x > xmax) i = nbin + 1;
  else if (x < xmax) i = (int)((x - xmin) / (xmax - xmin) * (double)nbin) + 1;
  else i = nbin;
  const double eps = (xmax - xmin) * 1.e-10;
  if (i == 0 && fabs(x - xmin) < eps) i = 1;
  if (i == nbin + 1 && fabs(x - xmax) < eps) i = nbin;
  return i;
}

// find_position_in_grid: follow goto links from grid g down to the cell that holds (x, y, z)
__device__ inline bool amr_locate(const AmrGrid &A, int g, double x, double y, double z, int &gout, int &i1, int &i2, int &i3) {
  for (;;) {
    const AmrGridDev &G = A.grids[g];
    const int j1 = amr_ipos2(G.xmin, G.xmax, x, G.n1), j2 = amr_ipos2(G.ymin, G.ymax, y, G.n2),
              j3 = amr_ipos2(G.zmin, G.zmax, z, G.n3);
    const int go = __ldg(A.gotos + G.goto_off + j1 + (int64_t)(G.n1 + 2) * (j2 + (int64_t)(G.n2 + 2) * j3));
    if (go == 0) {
      if (j1 < 1 || j1 > G.n1 || j2 < 1 || j2 > G.n2 || j3 < 1 || j3 > G.n3) return false;
      gout = g;
      i1 = j1 - 1; i2 = j2 - 1; i3 = j3 - 1;
      return true;
    }
    g = go - 1;
  }
}

// find_cell_position: level-1 grid that contains the point, then refine
__device__ inline bool amr_find_cell(const AmrGrid &A, double x, double y, double z, int &g, int &i1, int &i2, int &i3) {
  for (int k = 0; k < A.n_level1; ++k) {
    const AmrGridDev &G = A.grids[k];
    if (x < G.xmin || x > G.xmax || y < G.ymin || y > G.ymax || z < G.zmin || z > G.zmax) continue;
    return amr_locate(A, k, x, y, z, g, i1, i2, i3);
  }
  return false;
}

__device__ __forceinline__ int amr_cell_id(const AmrGridDev &G, int i1, int i2, int i3) {
  return G.start_id + (i3 * G.n2 + i2) * G.n1 + i1;
}

__device__ __forceinline__ void amr_start(const AmrGrid &A, AmrRay &R, double rx, double ry, double rz, double vx, double vy,
                                          double vz, int i1, int i2, int i3, int ic) {
  R.r0x = rx; R.r0y = ry; R.r0z = rz;
  R.vx = vx; R.vy = vy; R.vz = vz;
  R.ivx = 1.0 / vx; R.ivy = 1.0 / vy; R.ivz = 1.0 / vz;
  R.t = 0.0;
  R.g = __ldg(A.cell_grid + ic);
  R.i1 = i1; R.i2 = i2; R.i3 = i3;
  R.ic = ic;
}

// find_wall: path length to the wall the ray leaves its cell through and that wall (0..5)
__device__ __forceinline__ void amr_find_wall(const AmrGridDev &G, const AmrRay &R, double &dt, int &wall) {
  const double huge = 1.7976931348623157e308;
  const bool px = R.vx > 0.0, py = R.vy > 0.0, pz = R.vz > 0.0;
  const double tx = R.vx != 0.0 ? (amr_wall(G.xmin, G.xmax, R.i1 + (px ? 1 : 0), G.n1) - R.r0x) * R.ivx - R.t : huge;
  const double ty = R.vy != 0.0 ? (amr_wall(G.ymin, G.ymax, R.i2 + (py ? 1 : 0), G.n2) - R.r0y) * R.ivy - R.t : huge;
  const double tz = R.vz != 0.0 ? (amr_wall(G.zmin, G.zmax, R.i3 + (pz ? 1 : 0), G.n3) - R.r0z) * R.ivz - R.t : huge;
  if (tx < tz) {
    if (tx < ty) { wall = px ? 1 : 0; dt = tx; } else { wall = py ? 3 : 2; dt = ty; }
  } else {
    if (tz < ty) { wall = pz ? 5 : 4; dt = tz; } else { wall = py ? 3 : 2; dt = ty; }
  }
  // the reference recomputes from the moved position and aborts on a negative distance; here the distance
  // is a difference of absolute path lengths and can only be negative by rounding
  if (dt < 0.0) dt = 0.0;
}

// next_cell: R.t has been advanced to the wall
__device__ inline void amr_step(const AmrGrid &A, AmrRay &R, int wall) {
  const AmrGridDev &G = A.grids[R.g];
  int j1 = R.i1 + 1, j2 = R.i2 + 1, j3 = R.i3 + 1;  // 1-based, ghost layer at 0 and n+1
  const int axis = wall >> 1, s = (wall & 1) ? 1 : -1;
  if (axis == 0) j1 += s; else if (axis == 1) j2 += s; else j3 += s;
  const int go = __ldg(A.gotos + G.goto_off + j1 + (int64_t)(G.n1 + 2) * (j2 + (int64_t)(G.n2 + 2) * j3));
  if (go == 0) {
    if (j1 == 0 || j1 == G.n1 + 1 || j2 == 0 || j2 == G.n2 + 1 || j3 == 0 || j3 == G.n3 + 1) {
      R.ic = -2;  // outside_cell
      return;
    }
    R.i1 = j1 - 1; R.i2 = j2 - 1; R.i3 = j3 - 1;
    R.ic = amr_cell_id(G, R.i1, R.i2, R.i3);
    return;
  }
  double x = R.r0x + R.t * R.vx, y = R.r0y + R.t * R.vy, z = R.r0z + R.t * R.vz;
  if (axis == 0) x += s * A.eps; else if (axis == 1) y += s * A.eps; else z += s * A.eps;
  int g, i1, i2, i3;
  if (!amr_locate(A, go - 1, x, y, z, g, i1, i2, i3)) {
    R.ic = -1;  // invalid_cell
    return;
  }
  R.g = g;
  R.i1 = i1; R.i2 = i2; R.i3 = i3;
  R.ic = amr_cell_id(A.grids[g], i1, i2, i3);
}

__device__ __forceinline__ double amr_volume(const AmrGrid &A, int64_t ic) {
  const AmrGridDev &G = A.grids[__ldg(A.cell_grid + ic)];
  return (G.xmax - G.xmin) / (double)G.n1 * ((G.ymax - G.ymin) / (double)G.n2) * ((G.zmax - G.zmin) / (double)G.n3);
}

}  // namespace hyp
