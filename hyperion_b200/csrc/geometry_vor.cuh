// geometry_vor.cuh -- Voronoi meshes (src/grid/grid_geometry_voronoi.f90, type_grid_voronoi.f90).
//
// A cell is the set of points nearer to its site than to any other; the file carries the sites, the neighbour
// lists (CSR; >= 0 a cell, -1 .. -6 the walls xmin, xmax, ymin, ymax, zmin, zmax of the box), bounding boxes and
// volumes (computed by the Python front end with voro++).  A packet leaves its cell through the nearest of the planes
// that bisect the segments to the neighbouring sites (find_wall, :322-402); the cell it is in is the nearest site
// (find_cell, :195-228 -- a kd-tree in the reference, a uniform grid of buckets here: any exact nearest-neighbour
// search returns the same site).
#pragma once

struct VorGrid {
  const double *sites;          // [n_cells][3]
  const double *bb;             // [n_cells][6] xmin, xmax, ymin, ymax, zmin, zmax of the cell's bounding box
  const double *volume;         // [n_cells], negative volumes set to 0
  const int32_t *nidx, *neigh;  // neighbour lists, the file's numbering
  const int32_t *valid;         // cells with volume > 0 (geo%mask_map)
  const int32_t *b_start, *b_sites;  // buckets of the nearest-site search
  double box[6];
  double bw[3];                 // bucket widths
  int32_t nb[3];                // buckets per axis
  int32_t n_cells, n_valid;
};

struct VorRay {
  double r0x, r0y, r0z, vx, vy, vz;
  double t;
  int ic;        // cell (n_cells = outside)
};

// the k <= 2 nearest sites of (x, y, z), 0-based; shells of buckets around the point's bucket until no unvisited
// bucket can hold a closer site (kdtree2_n_nearest)
__device__ inline void vor_nearest(const VorGrid &G, double x, double y, double z, int k, int &i0, int &i1) {
  double best0 = 1.7976931348623157e308, best1 = 1.7976931348623157e308;
  i0 = i1 = 0;
  const int bx = min(max((int)((x - G.box[0]) / G.bw[0]), 0), G.nb[0] - 1);
  const int by = min(max((int)((y - G.box[2]) / G.bw[1]), 0), G.nb[1] - 1);
  const int bz = min(max((int)((z - G.box[4]) / G.bw[2]), 0), G.nb[2] - 1);
  const double wmin = fmin(G.bw[0], fmin(G.bw[1], G.bw[2]));
  const int rmax = max(G.nb[0], max(G.nb[1], G.nb[2]));
  for (int r = 0; r <= rmax; ++r) {
    if (r > 0) {
      const double reach = (double)(r - 1) * wmin;
      if ((k == 1 ? best0 : best1) <= reach * reach) break;
    }
    for (int k3 = bz - r; k3 <= bz + r; ++k3) {
      if (k3 < 0 || k3 >= G.nb[2]) continue;
      for (int k2 = by - r; k2 <= by + r; ++k2) {
        if (k2 < 0 || k2 >= G.nb[1]) continue;
        const bool shell23 = abs(k3 - bz) == r || abs(k2 - by) == r;
        // inside the (k2, k3) shell only the two end buckets along x belong to shell r
        const int step = (shell23 || r == 0) ? 1 : 2 * r;
        for (int k1 = bx - r; k1 <= bx + r; k1 += step) {
          if (k1 < 0 || k1 >= G.nb[0]) continue;
          const int cell = (k3 * G.nb[1] + k2) * G.nb[0] + k1;
          for (int q = __ldg(G.b_start + cell); q < __ldg(G.b_start + cell + 1); ++q) {
            const int i = __ldg(G.b_sites + q);
            const double dx = __ldg(G.sites + 3 * (size_t)i) - x, dy = __ldg(G.sites + 3 * (size_t)i + 1) - y,
                         dz = __ldg(G.sites + 3 * (size_t)i + 2) - z;
            const double d2 = dx * dx + dy * dy + dz * dz;
            if (d2 < best0) {
              best1 = best0; i1 = i0;
              best0 = d2; i0 = i;
            } else if (d2 < best1) {
              best1 = d2; i1 = i;
            }
          }
        }
      }
    }
  }
}

// find_cell (:195-228): -1 outside the box
__device__ inline int vor_find_cell(const VorGrid &G, double x, double y, double z) {
  if (x < G.box[0] || x > G.box[1] || y < G.box[2] || y > G.box[3] || z < G.box[4] || z > G.box[5]) return -1;
  int i0, i1;
  vor_nearest(G, x, y, z, 1, i0, i1);
  return i0;
}

// find_wall (:322-402) at the packet's position r0 + t v: path length to the nearest wall and the cell behind it.
// The reference takes the smallest positive t over all bisecting planes; here only the planes the packet moves
// TOWARDS are candidates (n . v > 0: a point inside its cell can only leave through those), and a path length that
// rounding made negative counts as zero.  The wall the packet has just come through has n . v < 0 and drops out by
// itself, and every crossing moves to a site farther along v, so a march cannot cycle at a vertex of the mesh.
__device__ inline bool vor_find_wall(const VorGrid &G, const VorRay &R, double &dt, int &next) {
  const double x = R.r0x + R.t * R.vx, y = R.r0y + R.t * R.vy, z = R.r0z + R.t * R.vz;
  const double sx = __ldg(G.sites + 3 * (size_t)R.ic), sy = __ldg(G.sites + 3 * (size_t)R.ic + 1),
               sz = __ldg(G.sites + 3 * (size_t)R.ic + 2);
  double tmin = 1.7976931348623157e308;
  int imin = -1;
  for (int q = __ldg(G.nidx + R.ic), q1 = __ldg(G.nidx + R.ic + 1); q < q1; ++q) {
    const int nb = __ldg(G.neigh + q);
    double t;
    int id;
    if (nb < 0) {
      // the walls of the box, only when the packet moves towards them
      const int a = (-nb - 1) >> 1, hi = (-nb - 1) & 1;
      const double v = a == 0 ? R.vx : (a == 1 ? R.vy : R.vz), r = a == 0 ? x : (a == 1 ? y : z);
      if (hi ? !(v > 0.0) : !(v < 0.0)) continue;
      t = (G.box[2 * a + hi] - r) / v;
      id = G.n_cells;
    } else {
      const double ox = __ldg(G.sites + 3 * (size_t)nb), oy = __ldg(G.sites + 3 * (size_t)nb + 1),
                   oz = __ldg(G.sites + 3 * (size_t)nb + 2);
      const double nx = ox - sx, ny = oy - sy, nz = oz - sz;
      const double den = nx * R.vx + ny * R.vy + nz * R.vz;
      if (!(den > 0.0)) continue;
      const double mx = 0.5 * (ox + sx), my = 0.5 * (oy + sy), mz = 0.5 * (oz + sz);
      t = (nx * (mx - x) + ny * (my - y) + nz * (mz - z)) / den;
      id = nb;
    }
    t = fmax(t, 0.0);
    if (t < tmin) {
      tmin = t;
      imin = id;
    }
  }
  if (imin < 0) return false;
  dt = tmin;
  next = imin;
  return true;
}
