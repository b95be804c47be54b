// geometry_oct.cuh -- octree traversal on the device.
//
// Restates src/grid/grid_geometry_octree.f90: locate_cell (:135-146), find_cell (:277-297), find_wall
// (:438-537), next_cell (:328-367), random_position_cell (:396-408).  The reference finds the cell
// behind a wall by walking UP the tree until a sibling covers the far side and then DOWN again by
// position.  Here the upward walk is done once on the host: every node stores, for each of its six
// faces, the node of the same or a coarser level that covers the region behind that face (-1 outside the
// grid), so a crossing costs one link load plus the downward walk through refined nodes, with no
// recursion.
#pragma once

#ifndef OCT_LDG
#define OCT_LDG 1
#endif
#ifndef OCT_PREFETCH
#define OCT_PREFETCH 1
#endif

namespace hyp {

struct OctNode {
  double x, y, z, dx, dy, dz;  // centre and HALF-widths (grid_geometry_octree.f90:74,252)
  int32_t nb[6];               // neighbour behind wall -x, +x, -y, +y, -z, +z
  int32_t first_child;         // index into the children table (8 entries), -1 for a leaf
  int32_t pad;
};

struct OctGrid {
  const OctNode *nodes;
  const int32_t *children;     // [n_refined][8], x fastest (subcell order of :41-49)
  const int32_t *leaves;       // ids of the leaf nodes (geo%mask_map)
  int32_t n_nodes, n_leaves;
  double eps;                  // 3 * spacing(largest root half-width) (:262)
};

struct OctRay {
  double r0x, r0y, r0z, vx, vy, vz, ivx, ivy, ivz;
  double t;
  int ic;  // node id, n_nodes = outside
};

// a node through the read-only path, as five 16-byte loads
__device__ __forceinline__ OctNode oct_load(const OctNode *p) {
  union {
    OctNode n;
    int4 q[5];
  } u;
  static_assert(sizeof(OctNode) == 80, "OctNode is five 16-byte words");
#pragma unroll
  for (int k = 0; k < 5; ++k) u.q[k] = __ldg((const int4 *)p + k);
  return u.n;
}

// locate_cell: descend from `node` to the leaf that contains (x, y, z)
__device__ __forceinline__ int oct_descend(const OctGrid &G, int node, double x, double y, double z) {
  for (;;) {
    const OctNode *N = G.nodes + node;
#if OCT_LDG
#if OCT_PREFETCH
    // the march reads the whole node next; an 80-byte node straddles two 128-byte lines more often than not, so
    // ask for the line of its first word now, beside the load of its last word
    asm volatile("prefetch.global.L1 [%0];" ::"l"(N));
#endif
    const int4 tail = __ldg((const int4 *)N + 4);   // nb[4], nb[5], first_child, pad
    const int fc = tail.z;
    if (fc < 0) return node;
    const double2 xy = __ldg((const double2 *)N);
    const double nz = __ldg(&N->z);
    const int sub = (x < xy.x ? 0 : 1) + (y < xy.y ? 0 : 2) + (z < nz ? 0 : 4);
#else
    const int fc = N->first_child;
    if (fc < 0) return node;
    const int sub = (x < N->x ? 0 : 1) + (y < N->y ? 0 : 2) + (z < N->z ? 0 : 4);
#endif
    node = __ldg(G.children + (size_t)fc * 8 + sub);
  }
}

// find_cell: -1 if outside the grid
__device__ __forceinline__ int oct_find_cell(const OctGrid &G, double x, double y, double z) {
  const OctNode *R = G.nodes;
  if (x < R->x - R->dx || x > R->x + R->dx) return -1;
  if (y < R->y - R->dy || y > R->y + R->dy) return -1;
  if (z < R->z - R->dz || z > R->z + R->dz) return -1;
  return oct_descend(G, 0, x, y, z);
}

__device__ __forceinline__ void oct_start(OctRay &R, double rx, double ry, double rz, double vx, double vy, double vz, int ic) {
  R.r0x = rx; R.r0y = ry; R.r0z = rz;
  R.vx = vx; R.vy = vy; R.vz = vz;
  R.ivx = 1.0 / vx; R.ivy = 1.0 / vy; R.ivz = 1.0 / vz;
  R.t = 0.0;
  R.ic = ic;
}

__device__ __forceinline__ bool oct_escaped(const OctGrid &G, const OctRay &R) { return R.ic >= G.n_nodes; }

// find_wall: path length from the current position to the wall the ray leaves the cell through, and
// that wall (0..5).  Returns false for the reference's "negative t" failure.
__device__ __forceinline__ bool oct_find_wall(const OctGrid &G, const OctRay &R, const OctNode &N, double &dt, int &wall) {
  const double huge = 1.7976931348623157e308;
  const bool px = R.vx > 0.0, py = R.vy > 0.0, pz = R.vz > 0.0;
  const double tx = R.vx != 0.0 ? (N.x + (px ? N.dx : -N.dx) - R.r0x) * R.ivx - R.t : huge;
  const double ty = R.vy != 0.0 ? (N.y + (py ? N.dy : -N.dy) - R.r0y) * R.ivy - R.t : huge;
  const double tz = R.vz != 0.0 ? (N.z + (pz ? N.dz : -N.dz) - R.r0z) * R.ivz - R.t : huge;
  if (tx < tz) {
    if (tx < ty) { wall = px ? 1 : 0; dt = tx; } else { wall = py ? 3 : 2; dt = ty; }
  } else {
    if (tz < ty) { wall = pz ? 5 : 4; dt = tz; } else { wall = py ? 3 : 2; dt = ty; }
  }
  if (dt < 0.0) {
    if (dt > -10.0 * G.eps) dt = 0.0; else return false;
  }
  return true;
}

// cross `wall` of node N at path length R.t: the neighbour link, then down to the leaf at the crossing point
__device__ __forceinline__ void oct_step(const OctGrid &G, OctRay &R, const OctNode &N, int wall) {
  const int nb = N.nb[wall];
  if (nb < 0) {
    R.ic = G.n_nodes;
    return;
  }
  R.ic = oct_descend(G, nb, R.r0x + R.t * R.vx, R.r0y + R.t * R.vy, R.r0z + R.t * R.vz);
}

}  // namespace hyp
