// Π assembly: row frames -> fused (point generation + LBVH location + pull-back + tabulation)
// -> segment-structured COO -> CSR.
//
// Reference stages replaced (file:line under /root/reference/src/fenicsx_ii/):
//   row_frames_kernel      utils.py:11-33 (push-forward, tangents), quadrature.py:35-69 (Rodrigues),
//                          restriction_operators.py:189-205 / 301-315 (radius, weights, scales)
//   pi_locate_tabulate     restriction_operators.py:191-199 (averaging points),
//                          interpolation.py:66-68 (determine_point_ownership),
//                          interpolation_utils.py:36-38 (pull_back), :59-74 (basis tabulation),
//                          interpolation.py:234-238 (value = (phi*w)/scale)
//   rows_count / rows_fill interpolation.py:140-163 (SparsityPattern insert + finalize),
//                          :205-250 (first-visit-wins masking, ADD_VALUES in l order, assemble)
//
// This translation unit is compiled with -fmad=false so the fp64 geometry follows the same
// operation order as the CPU oracle (no FMA contraction).
#include <cub/cub.cuh>
#include <vector>

#include "common.cuh"

namespace fxii {

#ifndef FXII_KSTACK
#define FXII_KSTACK 96
#endif
// Traversal stack: the 4-wide walk pushes at most three siblings per wide level; the binary LBVH is at most 95 levels
// deep (63-bit Morton key + 32-bit index tie-break), i.e. 48 wide levels.  Real meshes need 30-45 entries (config 5:
// binary depth 27).  A walk that would need more reports it (PiCounters::stack_overflow -> FXII_E_UNSUPPORTED) instead
// of dropping the subtree.  (-DFXII_KSTACK=n builds a variant with a small stack for the test that trips the check.)
constexpr int kStack = FXII_KSTACK;
constexpr double kCollisionD2 = 10.0 * 2.220446049250313e-16;  // dolfinx: d^2 < 10*eps

// frame layout (16 doubles = 128 B per point row):
//  [0..2] x0, [3] R, [4..12] Rot row-major, [13] weight multiplier, [14] scale, [15] unused
constexpr int kFrame = 16;

// Frame of point row r: the K interpolation point on its Γ segment and, for the averaging operators,
// the rotation taking the reference circle/disk into the plane normal to the segment.
__device__ __forceinline__ void compute_frame(const double* __restrict__ gx, const int32_t* __restrict__ gcells,
                                              const int32_t* __restrict__ cells_k, const double* __restrict__ xref, int nip,
                                              const double* __restrict__ radius, int radius_per_row, int kind, int64_t r,
                                              double* f) {
  const double PI = 3.141592653589793;
  const int64_t i = r / nip;
  const int j = (int)(r - i * nip);
  const int c = cells_k[i];
  const int g0 = gcells[2 * (int64_t)c], g1 = gcells[2 * (int64_t)c + 1];
  double v0[3], v1[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    v0[d] = gx[3 * (int64_t)g0 + d];
    v1[d] = gx[3 * (int64_t)g1 + d];
  }
  const double X = xref[j];
  const double omx = 1.0 - X;
#pragma unroll
  for (int d = 0; d < 3; ++d) f[d] = v0[d] * omx + v1[d] * X;  // utils.py:18-20
  if (kind == FXII_OP_TRACE) {
    f[3] = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) f[4 + k] = 0.0;
    f[13] = 1.0;
    f[14] = 1.0;
    f[15] = 0.0;
    return;
  }
  // tangent (utils.py:31-32) then the normalisation of rotation_matrix (quadrature.py:43)
  double t[3] = {v0[0] - v1[0], v0[1] - v1[1], v0[2] - v1[2]};
  double nrm = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
#pragma unroll
  for (int d = 0; d < 3; ++d) t[d] = t[d] / nrm;
  nrm = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
  double n[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) n[d] = t[d] / nrm;
  // axis = ez x n ; ez when the cross product vanishes (np.isclose(.,0): <= 1e-8)
  double a[3] = {-n[1], n[0], 0.0};
  double an = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  if (an <= 1.0e-8) {
    a[0] = 0.0;
    a[1] = 0.0;
    a[2] = 1.0;
  }
  an = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
#pragma unroll
  for (int d = 0; d < 3; ++d) a[d] = a[d] / an;
  const double ct = n[2];
  const double st = sqrt(1.0 - ct * ct);
  const double omc = 1.0 - ct;
  const double sm[9] = {0.0, -a[2], a[1], a[2], 0.0, -a[0], -a[1], a[0], 0.0};
#pragma unroll
  for (int p = 0; p < 3; ++p)
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double diag = (p == q) ? ct : 0.0;
      f[4 + 3 * p + q] = (diag + sm[3 * p + q] * st) + (a[p] * a[q]) * omc;
    }
  // radius_per_row: 0 one radius, 1 one per point row, 2 one per processed Γ cell (row r belongs to cell slot i = r / nip)
  const double R = radius_per_row == 1 ? radius[r] : (radius_per_row == 2 ? radius[i] : radius[0]);
  f[3] = R;
  if (kind == FXII_OP_CIRCLE) {
    const double scale = 2.0 * R * PI;  // restriction_operators.py:203
    f[13] = scale;
    f[14] = scale;
  } else {  // DISK
    f[13] = R * R;         // :312
    f[14] = PI * (R * R);  // :311
  }
  f[15] = 0.0;
}

__global__ void __launch_bounds__(128)
row_frames_kernel(const double* __restrict__ gx, const int32_t* __restrict__ gcells, const int32_t* __restrict__ cells_k,
                  const double* __restrict__ xref, int nip, const double* __restrict__ radius, int radius_per_row,
                  int kind, int64_t n_point_rows, double* __restrict__ frames) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_point_rows; r += (int64_t)gridDim.x * blockDim.x) {
    double f[kFrame];
    compute_frame(gx, gcells, cells_k, xref, nip, radius, radius_per_row, kind, r, f);
#pragma unroll
    for (int k = 0; k < kFrame; ++k) frames[kFrame * r + k] = f[k];
  }
}

// ---- geometry helpers --------------------------------------------------------------------------
__device__ __forceinline__ double dot3(const double* a, const double* b) {
  return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}

// squared distance point-triangle (Ericson, Real-Time Collision Detection 5.1.5); mirrors
// oracle/pi_oracle.py:_closest_point_triangle_d2
__device__ __noinline__ double tri_dist2(const double* p, const double* a, const double* b, const double* c) {
  double ab[3], ac[3], ap[3], bp[3], cp[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    ab[d] = b[d] - a[d];
    ac[d] = c[d] - a[d];
    ap[d] = p[d] - a[d];
    bp[d] = p[d] - b[d];
    cp[d] = p[d] - c[d];
  }
  const double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
  const double d3 = dot3(ab, bp), d4 = dot3(ac, bp);
  const double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
  const double vc = d1 * d4 - d3 * d2;
  const double vb = d5 * d2 - d1 * d6;
  const double va = d3 * d6 - d5 * d4;
  double q[3];
  if (d1 <= 0 && d2 <= 0) {
    q[0] = a[0]; q[1] = a[1]; q[2] = a[2];
  } else if (d3 >= 0 && d4 <= d3) {
    q[0] = b[0]; q[1] = b[1]; q[2] = b[2];
  } else if (vc <= 0 && d1 >= 0 && d3 <= 0) {
    const double v = d1 / (d1 - d3);
#pragma unroll
    for (int d = 0; d < 3; ++d) q[d] = a[d] + v * ab[d];
  } else if (d6 >= 0 && d5 <= d6) {
    q[0] = c[0]; q[1] = c[1]; q[2] = c[2];
  } else if (vb <= 0 && d2 >= 0 && d6 <= 0) {
    const double w = d2 / (d2 - d6);
#pragma unroll
    for (int d = 0; d < 3; ++d) q[d] = a[d] + w * ac[d];
  } else if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
#pragma unroll
    for (int d = 0; d < 3; ++d) q[d] = b[d] + w * (c[d] - b[d]);
  } else {
    const double denom = 1.0 / (va + vb + vc);
    const double v = vb * denom, w = vc * denom;
#pragma unroll
    for (int d = 0; d < 3; ++d) q[d] = (a[d] + ab[d] * v) + ac[d] * w;
  }
  double e[3] = {p[0] - q[0], p[1] - q[1], p[2] - q[2]};
  return dot3(e, e);
}

struct CellGeo {
  double v0[3];
  double K[9];
};

__device__ __forceinline__ void load_cellgeo(const double* __restrict__ cellgeo, int cell, CellGeo& g) {
  const double2* src = reinterpret_cast<const double2*>(cellgeo + 12 * (int64_t)cell);
  double2 a0 = __ldg(src), a1 = __ldg(src + 1), a2 = __ldg(src + 2), a3 = __ldg(src + 3), a4 = __ldg(src + 4),
          a5 = __ldg(src + 5);
  g.v0[0] = a0.x; g.v0[1] = a0.y; g.v0[2] = a1.x;
  g.K[0] = a1.y; g.K[1] = a2.x; g.K[2] = a2.y;
  g.K[3] = a3.x; g.K[4] = a3.y; g.K[5] = a4.x;
  g.K[6] = a4.y; g.K[7] = a5.x; g.K[8] = a5.y;
}

// X = K (x - v0): interpolation_utils.py:36-38 for affine cells
__device__ __forceinline__ void pull_back(const CellGeo& g, const double* p, double* X) {
  const double d0 = p[0] - g.v0[0], d1 = p[1] - g.v0[1], d2 = p[2] - g.v0[2];
  X[0] = (g.K[0] * d0 + g.K[1] * d1) + g.K[2] * d2;
  X[1] = (g.K[3] * d0 + g.K[4] * d1) + g.K[5] * d2;
  X[2] = (g.K[6] * d0 + g.K[7] * d1) + g.K[8] * d2;
}

// Slow path of the leaf test (some barycentric coordinate negative): exact fp64 padded-box test
// (dolfinx point_in_bbox, slack 1e-14*extent) and exact point-tetrahedron distance.
// Returns -1 when the cell is not a candidate; otherwise d2 (>= 0).
// (point by value, result by value: a pointer or reference here would pin the caller's point and search state
// in local memory for the whole traversal)
__device__ __noinline__ double cell_distance2(const double* __restrict__ x, const int4* __restrict__ cell_nodes, int cell,
                                              double p0, double p1, double p2, double padding) {
  const double p[3] = {p0, p1, p2};
  const int4 nd = __ldg(&cell_nodes[cell]);
  const int id[4] = {nd.x, nd.y, nd.z, nd.w};
  double v[4][3];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int d = 0; d < 3; ++d) v[k][d] = __ldg(&x[3 * (int64_t)id[k] + d]);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double lo = fmin(fmin(v[0][d], v[1][d]), fmin(v[2][d], v[3][d])) - padding;
    const double hi = fmax(fmax(v[0][d], v[1][d]), fmax(v[2][d], v[3][d])) + padding;
    const double eps = 1e-14 * (hi - lo);
    if (!(p[d] >= lo - eps && p[d] <= hi + eps)) return -1.0;
  }
  double best = tri_dist2(p, v[1], v[2], v[3]);
  best = fmin(best, tri_dist2(p, v[0], v[2], v[3]));
  best = fmin(best, tri_dist2(p, v[0], v[1], v[3]));
  best = fmin(best, tri_dist2(p, v[0], v[1], v[2]));
  return best;
}

// ---- hexahedra (Q1 geometry, vertices v = i + 2j + 4k; mirrors oracle/pi_oracle.py: q1_tabulate, pull_back_hex,
// point_hex_distance2) -----------------------------------------------------------------------------------------------
struct HexX {
  double X0, X1, X2;
  int conv;
};
constexpr double kNewtonTol2 = 1.0e-12;  // |dX| < 1e-6: dolfinx CoordinateElement::pull_back_nonaffine defaults [3P-recall]
constexpr int kNewtonMaxIt = 15;

// Newton pull-back into the hexahedron whose eight vertices are hv[24] (interpolation_utils.py:36-38 for non-affine
// cells): X = 0; X += J^{-1} (x - x(X)) until |dX| < tol.  The vertices are re-read from L1 in every iteration
// instead of being held in registers (24 doubles); point and result travel by value.
__device__ __noinline__ HexX hex_pull_back(const double* __restrict__ hv, double p0, double p1, double p2) {
  double X[3] = {0.0, 0.0, 0.0};
  int conv = 0;
  for (int it = 0; it < kNewtonMaxIt; ++it) {
    const double fx[2] = {1.0 - X[0], X[0]}, fy[2] = {1.0 - X[1], X[1]}, fz[2] = {1.0 - X[2], X[2]};
    double xk[3] = {0.0, 0.0, 0.0};
    double J[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const double a = fx[v & 1], b = fy[(v >> 1) & 1], c = fz[(v >> 2) & 1];
      const double sa = (v & 1) ? 1.0 : -1.0, sb = ((v >> 1) & 1) ? 1.0 : -1.0, sc = ((v >> 2) & 1) ? 1.0 : -1.0;
      const double ph = (a * b) * c;
      const double g0 = (sa * b) * c, g1 = (a * sb) * c, g2 = (a * b) * sc;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double vd = __ldg(&hv[3 * v + d]);
        xk[d] += ph * vd;
        J[d][0] += vd * g0;
        J[d][1] += vd * g1;
        J[d][2] += vd * g2;
      }
    }
    const double r0 = p0 - xk[0], r1 = p1 - xk[1], r2 = p2 - xk[2];
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    const double det = (J[0][0] * c00 + J[0][1] * c01) + J[0][2] * c02;
    if (!(fabs(det) > 1e-300)) break;  // singular map: not converged
    const double dX0 = ((c00 * r0 + (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * r1) + (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * r2) / det;
    const double dX1 = ((c01 * r0 + (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * r1) + (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * r2) / det;
    const double dX2 = ((c02 * r0 + (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * r1) + (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * r2) / det;
    X[0] += dX0;
    X[1] += dX1;
    X[2] += dX2;
    if ((dX0 * dX0 + dX1 * dX1) + dX2 * dX2 < kNewtonTol2) {
      conv = 1;
      break;
    }
  }
  return HexX{X[0], X[1], X[2], conv};
}

// Squared distance of a point to a hexahedron in the sense of dolfinx's GJK test on the convex hull of the eight
// vertices, for cells with planar faces: 0 when the Newton pull-back converges into the closed reference cube (the
// trilinear image lies inside the hull), else the minimum over the twelve boundary triangles.  Returns -1 when the
// point is outside the padded box (dolfinx point_in_bbox, slack 1e-14*extent): not a candidate.
__device__ __noinline__ double hex_distance2(const double* __restrict__ hv, double p0, double p1, double p2, double padding) {
  const double p[3] = {p0, p1, p2};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    double lo = __ldg(&hv[d]), hi = lo;
#pragma unroll
    for (int v = 1; v < 8; ++v) {
      const double c = __ldg(&hv[3 * v + d]);
      lo = fmin(lo, c);
      hi = fmax(hi, c);
    }
    lo -= padding;
    hi += padding;
    const double eps = 1e-14 * (hi - lo);
    if (!(p[d] >= lo - eps && p[d] <= hi + eps)) return -1.0;
  }
  const HexX r = hex_pull_back(hv, p0, p1, p2);
  if (r.conv && r.X0 >= 0.0 && r.X0 <= 1.0 && r.X1 >= 0.0 && r.X1 <= 1.0 && r.X2 >= 0.0 && r.X2 <= 1.0) return 0.0;
  // faces x=0, x=1, y=0, y=1, z=0, z=1, each as two triangles along a fixed diagonal
  const int face[6][4] = {{0, 2, 6, 4}, {1, 3, 7, 5}, {0, 1, 5, 4}, {2, 3, 7, 6}, {0, 1, 3, 2}, {4, 5, 7, 6}};
  double best = 1e300;
  for (int f = 0; f < 6; ++f) {
    double q[4][3];
    for (int k = 0; k < 4; ++k)
      for (int d = 0; d < 3; ++d) q[k][d] = __ldg(&hv[3 * face[f][k] + d]);
    best = fmin(best, tri_dist2(p, q[0], q[1], q[2]));
    best = fmin(best, tri_dist2(p, q[0], q[2], q[3]));
  }
  return best;
}

struct LocateResult {
  int cell;      // lowest colliding cell index, or nearest-cell fallback, or -1
  int ncoll;     // size of the colliding set
  bool fallback;
};

struct LocateState {
  int best = 0x7fffffff;  // lowest colliding cell so far
  int ncoll = 0;
  int fb_cell = -1;       // nearest non-colliding candidate (fallback)
  double fb_d2 = 1e300;
  bool done = false;      // the point is strictly inside a cell: the colliding set is final
};

// Leaf test of one candidate cell (dolfinx compute_first_colliding_cell semantics).
__device__ __forceinline__ void test_cell(const double* __restrict__ cellgeo, const double* __restrict__ x,
                                          const int4* __restrict__ cell_nodes, double padding, double tau2, int cell,
                                          const double* p, LocateState& st) {
  CellGeo g;
  load_cellgeo(cellgeo, cell, g);
  double X[3];
  pull_back(g, p, X);
  const double l0 = ((1.0 - X[0]) - X[1]) - X[2];
  // squared gradient norms of the barycentric coordinates (grad lambda_1..3 = rows of K,
  // grad lambda_0 = -(row0+row1+row2)): |lambda_i| / |grad lambda_i| is the distance to face plane i
  const double s0 = (g.K[0] + g.K[3]) + g.K[6], s1 = (g.K[1] + g.K[4]) + g.K[7], s2 = (g.K[2] + g.K[5]) + g.K[8];
  const double t0 = tau2 * ((s0 * s0 + s1 * s1) + s2 * s2);
  const double t1 = tau2 * ((g.K[0] * g.K[0] + g.K[1] * g.K[1]) + g.K[2] * g.K[2]);
  const double t2 = tau2 * ((g.K[3] * g.K[3] + g.K[4] * g.K[4]) + g.K[5] * g.K[5]);
  const double t3 = tau2 * ((g.K[6] * g.K[6] + g.K[7] * g.K[7]) + g.K[8] * g.K[8]);
  if (l0 >= 0 && X[0] >= 0 && X[1] >= 0 && X[2] >= 0) {
    st.best = min(st.best, cell);
    ++st.ncoll;
    // Strictly inside by more than tau from every face plane: cells of a mesh have disjoint
    // interiors, so every other cell is farther than tau from the point and can neither collide
    // nor win the tie-break — the search is complete (colliding set = {cell}).
    st.done = st.ncoll == 1 && l0 * l0 > t0 && X[0] * X[0] > t1 && X[1] * X[1] > t2 && X[2] * X[2] > t3;
    return;
  }
  // Quick reject without touching the vertices: the distance to the tetrahedron is at least the
  // distance to the plane of any face the point is outside of.  Cells farther than
  // tau = max(collision distance, padding) can neither collide nor serve as fallback.
  bool far = (l0 < 0) && (l0 * l0 > t0);
  far = far || ((X[0] < 0) && (X[0] * X[0] > t1));
  far = far || ((X[1] < 0) && (X[1] * X[1] > t2));
  far = far || ((X[2] < 0) && (X[2] * X[2] > t3));
  const double d2 = far ? -1.0 : cell_distance2(x, cell_nodes, cell, p[0], p[1], p[2], padding);
  if (d2 >= 0.0) {
    if (d2 < kCollisionD2) {
      st.best = min(st.best, cell);
      ++st.ncoll;
    } else if (d2 < st.fb_d2 || (d2 == st.fb_d2 && cell < st.fb_cell)) {
      st.fb_d2 = d2;
      st.fb_cell = cell;
    }
  }
}

// Leaf test of one candidate hexahedron: same bookkeeping as test_cell (no early exit: `done` stays false).
__device__ __forceinline__ void test_cell_hex(const double* __restrict__ hexgeo, double padding, int cell, const double* p,
                                              LocateState& st) {
  const double d2 = hex_distance2(hexgeo + 24 * (int64_t)cell, p[0], p[1], p[2], padding);
  if (d2 >= 0.0) {
    if (d2 < kCollisionD2) {
      st.best = min(st.best, cell);
      ++st.ncoll;
    } else if (d2 < st.fb_d2 || (d2 == st.fb_d2 && cell < st.fb_cell)) {
      st.fb_d2 = d2;
      st.fb_cell = cell;
    }
  }
}

__device__ __forceinline__ LocateResult finish_locate(const LocateState& st, double padding) {
  LocateResult r;
  r.ncoll = st.ncoll;
  r.fallback = false;
  if (st.best != 0x7fffffff) {
    r.cell = st.best;
  } else if (st.fb_cell >= 0 && st.fb_d2 <= padding * padding) {
    r.cell = st.fb_cell;
    r.fallback = true;
  } else {
    r.cell = -1;
  }
  return r;
}

// The point on the scene grid of the quantised boxes, each coordinate replicated in both halfwords
// for the two-slot SIMD compares.  q() is the same monotone map the builder applied to the box
// corners, so a point inside a (fp32, outward-rounded) box is always inside its quantised box.
struct QPoint {
  uint32_t x2, y2, z2;  // coordinate | 0x8000 in both halfwords
};
__device__ __forceinline__ QPoint quantise_point(const SceneGrid& g, const double* p) {
  QPoint q;
  uint32_t v[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) v[a] = (uint32_t)fmin(fmax(floor((p[a] - g.origin[a]) * g.scale[a]), 0.0), 32767.0);
  q.x2 = (v[0] | (v[0] << 16)) | 0x80008000u;
  q.y2 = (v[1] | (v[1] << 16)) | 0x80008000u;
  q.z2 = (v[2] | (v[2] << 16)) | 0x80008000u;
  return q;
}

// one 64 B wide node: hit flags of its four slots
struct WideHit {
  int4 child;
  uint32_t m01, m23;  // bit 15 / bit 31: slots (0,1) and (2,3)
};
// 15-bit coordinates, two per word.  For halfwords a, b < 0x8000: bit 15 of ((b | 0x8000) - a) is set
// iff a <= b, and the subtraction never borrows across halfwords — so one 32-bit SUB compares two slots.
// The point carries the 0x8000 bits for `lo <= q`, the stored hi carries them for `q <= hi`.
__device__ __forceinline__ WideHit load_test_node(const WideNode* __restrict__ nodes, int node, const QPoint& q) {
  const uint4* wp = reinterpret_cast<const uint4*>(nodes + node);
  // a = lo_x01 lo_x23 lo_y01 lo_y23 ; b = lo_z01 lo_z23 hi_x01 hi_x23 ; c = hi_y01 hi_y23 hi_z01 hi_z23
  const uint4 a = __ldg(wp), b = __ldg(wp + 1), c = __ldg(wp + 2);
  WideHit h;
  h.child = __ldg(reinterpret_cast<const int4*>(wp + 3));
  const uint32_t qx = q.x2 & 0x7fff7fffu, qy = q.y2 & 0x7fff7fffu, qz = q.z2 & 0x7fff7fffu;
  h.m01 = (q.x2 - a.x) & (q.y2 - a.z) & (q.z2 - b.x) & (b.z - qx) & (c.x - qy) & (c.z - qz) & 0x80008000u;
  h.m23 = (q.x2 - a.y) & (q.y2 - a.w) & (q.z2 - b.y) & (b.w - qx) & (c.y - qy) & (c.w - qz) & 0x80008000u;
  return h;
}

__device__ __forceinline__ void prefetch_l1(const void* ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }

// Per-thread traversal of the 4-wide tree.  Every leaf whose box contains the point is tested and the
// LOWEST colliding cell index wins, independent of tree shape (north_star tie-break); the only early
// exit is the provably final one: a point strictly inside a cell (by more than the collision distance)
// cannot collide with any other cell of a mesh.
template <bool HEX>
__device__ LocateResult locate_point(const WideNode* __restrict__ nodes, const SceneGrid& grid,
                                     const double* __restrict__ cellgeo, const double* __restrict__ x,
                                     const int4* __restrict__ cell_nodes, double padding, const double* p,
                                     unsigned long long* __restrict__ stack_overflow) {
  const QPoint q = quantise_point(grid, p);
  // 1.001: cells within 0.05% of the threshold still take the exact path
  const double tau2 = fmax(kCollisionD2, padding * padding) * 1.001;
  int stack[kStack];
  int sp = 0;
  int node = 0;
  LocateState st;
  while (true) {
    const WideHit h = load_test_node(nodes, node, q);
    const int kids[4] = {h.child.x, h.child.y, h.child.z, h.child.w};
    const bool hit[4] = {(h.m01 & 0xffffu) != 0, (h.m01 >> 16) != 0, (h.m23 & 0xffffu) != 0, (h.m23 >> 16) != 0};
    // Memory-level parallelism: the hit internal children are requested before anything is used (the ones that go
    // on the stack are otherwise cold misses when they are popped).  Prefetching the candidate cells' records as well
    // was measured slightly slower on config 5 (row kernel 9.59 ms with both, 9.39 ms with internal children only,
    // 9.54 ms with none, 10.08 ms with cells only): the records are used a few instructions later anyway.
#ifndef FXII_WALK_PREFETCH
#define FXII_WALK_PREFETCH 1  // bit 0: hit internal children, bit 1: candidate cells
#endif
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!hit[k]) continue;
      if (kids[k] >= 0) {
        if (FXII_WALK_PREFETCH & 1) prefetch_l1(nodes + kids[k]);
      } else if (FXII_WALK_PREFETCH & 2) {
        const double* g = cellgeo + (HEX ? 24 : 12) * (int64_t)(~kids[k]);
        prefetch_l1(g);
        prefetch_l1(g + 11);
        if (HEX) prefetch_l1(g + 23);
      }
    }
    int next = -1;
    // (Round 2, tried and rejected: sorting out the hit children first and testing the candidate cells of the node in ONE
    // loop body — a warp runs up to four inlined copies of the leaf test below with 8-14 of its 32 lanes active in each,
    // because which slots are leaves differs between lanes.  The converged loop spilled three times as much (592 vs 212
    // bytes at the 64-register cap) and lost the overlap of the four copies' loads: 11.42 vs 9.41 ms on config 5.)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!hit[k]) continue;
      if (kids[k] >= 0) {
        if (next < 0) next = kids[k];
        else if (sp < kStack) stack[sp++] = kids[k];
        else *stack_overflow = 1ULL;  // a dropped subtree could hide the owning cell: reported, never silent
      } else if (!st.done) {
        if (HEX) test_cell_hex(cellgeo, padding, ~kids[k], p, st);
        else test_cell(cellgeo, x, cell_nodes, padding, tau2, ~kids[k], p, st);
      }
    }
    if (st.done) break;
    if (next >= 0) node = next;
    else if (sp > 0) node = stack[--sp];
    else break;
  }
  return finish_locate(st, padding);
}

// Warp-cooperative (packet) traversal: the warp walks the UNION of its lanes' paths with one
// warp-uniform stack; every node and candidate cell is fetched once per warp at a uniform address
// and lanes vote with __ballot_sync on the children to descend into.  Results are identical to
// locate_point.  All 32 lanes must call this (inactive lanes pass active=false).  Measured slower
// than the per-thread walk on every BASELINE config (see packet_mode()).
template <bool HEX>
__device__ LocateResult locate_packet(const WideNode* __restrict__ nodes, const SceneGrid& grid,
                                      const double* __restrict__ cellgeo, const double* __restrict__ x,
                                      const int4* __restrict__ cell_nodes, double padding, const double* p, bool active,
                                      unsigned long long* __restrict__ stack_overflow) {
  const QPoint q = quantise_point(grid, p);
  const double tau2 = fmax(kCollisionD2, padding * padding) * 1.001;
  int stack[kStack];  // warp-uniform contents
  int sp = 0;
  int node = 0;
  LocateState st;
  while (true) {
    const WideHit h = load_test_node(nodes, node, q);  // uniform address: broadcast
    const int kids[4] = {h.child.x, h.child.y, h.child.z, h.child.w};
    const bool live = active && !st.done;  // finished lanes stop voting
    const bool hit[4] = {live && (h.m01 & 0xffffu) != 0, live && (h.m01 >> 16) != 0, live && (h.m23 & 0xffffu) != 0,
                         live && (h.m23 >> 16) != 0};
    int next = -1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (__ballot_sync(0xffffffffu, hit[k]) == 0u) continue;  // no lane needs this child
      if (kids[k] >= 0) {
        if (next < 0) next = kids[k];
        else if (sp < kStack) stack[sp++] = kids[k];
        else *stack_overflow = 1ULL;
      } else if (hit[k] && !st.done) {
        if (HEX) test_cell_hex(cellgeo, padding, ~kids[k], p, st);
        else test_cell(cellgeo, x, cell_nodes, padding, tau2, ~kids[k], p, st);
      }
    }
    if (next >= 0) node = next;
    else if (sp > 0) node = stack[--sp];
    else break;
  }
  return finish_locate(st, padding);
}

struct PiCounters {
  unsigned long long not_found, ties, fallback;
  unsigned long long stack_overflow;  // traversals that ran out of stack (a tree deeper than kStack allows): the build fails
};

// The 3D averaging point p = r*nq + l with its weight W and scale S
// (restriction_operators.py:191-205 / 301-315; PointwiseTrace :60-68; generic operators: uploaded).
__device__ __forceinline__ void make_point(int kind, const double* __restrict__ frames, const double* s_tab,
                                           const double* __restrict__ pts_in, const double* __restrict__ w_in,
                                           const double* __restrict__ s_in, int64_t r, int l, int64_t pidx, double* p,
                                           double& W, double& S) {
  if (kind == FXII_OP_POINTS) {
    p[0] = pts_in[3 * pidx];
    p[1] = pts_in[3 * pidx + 1];
    p[2] = pts_in[3 * pidx + 2];
    W = w_in[pidx];
    S = s_in[r];
    return;
  }
  const double* f = frames;  // the frame of point row r
  if (kind == FXII_OP_TRACE) {
    p[0] = f[0]; p[1] = f[1]; p[2] = f[2];
    W = 1.0;
    S = 1.0;
    return;
  }
  const double xp0 = s_tab[4 * l], xp1 = s_tab[4 * l + 1], xp2 = s_tab[4 * l + 2];
  const double R = f[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double applied = (f[4 + 3 * d] * xp0 + f[5 + 3 * d] * xp1) + f[6 + 3 * d] * xp2;
    p[d] = f[d] + applied * R;  // restriction_operators.py:191-199
  }
  W = s_tab[4 * l + 3] * f[13];
  S = f[14];
}

// ---- P3 (Lagrange degree 3 on tetrahedra, basix's default gll_warped variant) ---------------------------------------
// phi_i = sum_k c_p3[i][k] * m_k, m_k the 20 barycentric monomials l0^a l1^b l2^c l3^d (a+b+c+d = 3, a outermost, then
// b, then c).  The nodal basis is dual to point evaluation at: the vertices; on every edge (2,3),(1,3),(1,2),(0,3),
// (0,2),(0,1) the two Gauss-Lobatto-Legendre abscissae (1 -+ 1/sqrt 5)/2 measured from the edge's first vertex; the
// centroids of the faces (1,2,3),(0,2,3),(0,1,3),(0,1,2).  Lagrange elements need no dof transformation at tabulation
// time (they are permutations dolfinx applies to the dofmap), so the reference basis is used as is.  The table is a
// constant of the element: computed once on the host (Gaussian elimination of the 20x20 collocation matrix).
__constant__ double c_p3[20 * 20];

static void p3_exponents(int (*ex)[4]) {
  int k = 0;
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4 - a; ++b)
      for (int c = 0; c < 4 - a - b; ++c) {
        ex[k][0] = a; ex[k][1] = b; ex[k][2] = c; ex[k][3] = 3 - a - b - c;
        ++k;
      }
}

static void ensure_p3_table() {
  static int ready_dev = -1;
  int dev = 0;
  FXII_CUDA(cudaGetDevice(&dev));
  if (ready_dev == dev) return;
  static const int edges[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}};
  static const int faces[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};
  double node[20][4] = {};
  int n = 0;
  for (int v = 0; v < 4; ++v) node[n++][v] = 1.0;
  const double t[2] = {0.5 - 0.5 / sqrt(5.0), 0.5 + 0.5 / sqrt(5.0)};
  for (int e = 0; e < 6; ++e)
    for (int k = 0; k < 2; ++k) {
      node[n][edges[e][0]] = 1.0 - t[k];
      node[n][edges[e][1]] = t[k];
      ++n;
    }
  for (int f = 0; f < 4; ++f) {
    for (int k = 0; k < 3; ++k) node[n][faces[f][k]] = 1.0 / 3.0;
    ++n;
  }
  int ex[20][4];
  p3_exponents(ex);
  // A[j][k] = m_k(node_j); solve A X = I: column i of X holds the monomial coefficients of phi_i
  double A[20][40];
  for (int j = 0; j < 20; ++j) {
    for (int k = 0; k < 20; ++k) {
      double m = 1.0;
      for (int d = 0; d < 4; ++d)
        for (int q = 0; q < ex[k][d]; ++q) m *= node[j][d];
      A[j][k] = m;
    }
    for (int k = 0; k < 20; ++k) A[j][20 + k] = (j == k) ? 1.0 : 0.0;
  }
  for (int c = 0; c < 20; ++c) {  // Gauss-Jordan with partial pivoting (condition number ~40)
    int piv = c;
    for (int r = c + 1; r < 20; ++r)
      if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
    FXII_CHECK(fabs(A[piv][c]) > 1e-12, "P3 collocation matrix is singular");
    if (piv != c)
      for (int k = 0; k < 40; ++k) std::swap(A[piv][k], A[c][k]);
    const double inv = 1.0 / A[c][c];
    for (int k = 0; k < 40; ++k) A[c][k] *= inv;
    for (int r = 0; r < 20; ++r) {
      if (r == c) continue;
      const double f = A[r][c];
      if (f != 0.0)
        for (int k = 0; k < 40; ++k) A[r][k] -= f * A[c][k];
    }
  }
  double coef[400];
  for (int i = 0; i < 20; ++i)
    for (int k = 0; k < 20; ++k) coef[20 * i + k] = A[k][20 + i];
  FXII_CUDA(cudaMemcpyToSymbol(c_p3, coef, sizeof(coef)));
  ready_dev = dev;
}

// basis values of the owning cell at the point, scaled: (phi*W)/S  (interpolation.py:234-238)
template <int ND>
__device__ __forceinline__ void tabulate_vals(const double* __restrict__ cellgeo, int cell, const double* p, double W,
                                              double S, double* vals) {
  if (ND == 8) {  // Q1 on a hexahedron: Newton pull-back, tensor-product basis in vertex order v = i + 2j + 4k
    const HexX r = hex_pull_back(cellgeo + 24 * (int64_t)cell, p[0], p[1], p[2]);
    const double fx[2] = {1.0 - r.X0, r.X0}, fy[2] = {1.0 - r.X1, r.X1}, fz[2] = {1.0 - r.X2, r.X2};
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const double phi = (fx[v & 1] * fy[(v >> 1) & 1]) * fz[(v >> 2) & 1];
      vals[v % ND] = (phi * W) / S;
    }
    return;
  }
  CellGeo g;
  load_cellgeo(cellgeo, cell, g);
  double X[3];
  pull_back(g, p, X);
  const double lam[4] = {((1.0 - X[0]) - X[1]) - X[2], X[0], X[1], X[2]};
  if (ND == 20) {  // P3: table-driven (c_p3), monomials in the order of p3_exponents
    double pw[4][4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      pw[d][0] = 1.0;
      pw[d][1] = lam[d];
      pw[d][2] = lam[d] * lam[d];
      pw[d][3] = pw[d][2] * lam[d];
    }
    double m[20];
    int k = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4 - a; ++b)
#pragma unroll
        for (int c = 0; c < 4 - a - b; ++c) m[k++] = ((pw[0][a] * pw[1][b]) * pw[2][c]) * pw[3][3 - a - b - c];
    for (int i = 0; i < 20; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int q = 0; q < 20; ++q) acc += c_p3[20 * i + q] * m[q];
      vals[i % ND] = (acc * W) / S;
    }
    return;
  }
  double phi[ND];
  if (ND == 4) {
#pragma unroll
    for (int d = 0; d < 4; ++d) phi[d] = lam[d];
  } else {  // P2, basix order: vertices, then edges (2,3),(1,3),(1,2),(0,3),(0,2),(0,1)
#pragma unroll
    for (int d = 0; d < 4; ++d) phi[d] = lam[d] * (2.0 * lam[d] - 1.0);
    phi[4 % ND] = 4.0 * lam[2] * lam[3];
    phi[5 % ND] = 4.0 * lam[1] * lam[3];
    phi[6 % ND] = 4.0 * lam[1] * lam[2];
    phi[7 % ND] = 4.0 * lam[0] * lam[3];
    phi[8 % ND] = 4.0 * lam[0] * lam[2];
    phi[9 % ND] = 4.0 * lam[0] * lam[1];
  }
#pragma unroll
  for (int d = 0; d < ND; ++d) vals[d] = (phi[d] * W) / S;
}

template <int ND>
__device__ __forceinline__ void load_cols(const int32_t* __restrict__ dofmap, int cell, int32_t* cols) {
  const int32_t* dm = dofmap + (int64_t)ND * cell;
  if (ND == 4) {
    const int4 d4 = __ldg(reinterpret_cast<const int4*>(dm));
    cols[0] = d4.x; cols[1] = d4.y; cols[2] = d4.z; cols[3 % ND] = d4.w;
  } else {
#pragma unroll
    for (int d = 0; d < ND; d += 2) {
      const int2 d2 = __ldg(reinterpret_cast<const int2*>(dm + d));
      cols[d] = d2.x;
      cols[(d + 1) % ND] = d2.y;
    }
  }
}

// One thread per 3D point p = r*nq + l (interpolation.py:57).  ND = dofs per V cell.
template <int ND>
__global__ void __launch_bounds__(128, 8)
pi_locate_tabulate_kernel(const WideNode* __restrict__ nodes, const SceneGrid grid, const double* __restrict__ cellgeo,
                          const double* __restrict__ x, const int4* __restrict__ cell_nodes, double padding,
                          const int32_t* __restrict__ dofmap, const double* __restrict__ frames,
                          const double* __restrict__ table, const double* __restrict__ wq,
                          const double* __restrict__ pts_in, const double* __restrict__ w_in, const double* __restrict__ s_in,
                          int kind, int nq, int packet, int64_t n_points, int32_t* __restrict__ out_cell,
                          uint8_t* __restrict__ out_ncoll, double* __restrict__ out_points,
                          int32_t* __restrict__ coo_cols, double* __restrict__ coo_vals, PiCounters* counters) {
  extern __shared__ double s_tab[];  // [nq*4]: xp (3) + w
  if (kind == FXII_OP_CIRCLE || kind == FXII_OP_DISK) {
    for (int k = threadIdx.x; k < nq; k += blockDim.x) {
      s_tab[4 * k + 0] = table[3 * k + 0];
      s_tab[4 * k + 1] = table[3 * k + 1];
      s_tab[4 * k + 2] = table[3 * k + 2];
      s_tab[4 * k + 3] = wq[k];
    }
    __syncthreads();
  }
  for (int64_t base = blockIdx.x * (int64_t)blockDim.x; base < n_points; base += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pidx = base + threadIdx.x;  // the loop is block-uniform: the packet traversal votes
    const bool active = pidx < n_points;
    const int64_t r = active ? pidx / nq : 0;
    const int l = active ? (int)(pidx - r * nq) : 0;
    double p[3] = {0.0, 0.0, 0.0}, W = 1.0, S = 1.0;
    if (active) make_point(kind, frames + kFrame * r, s_tab, pts_in, w_in, s_in, r, l, pidx, p, W, S);
    LocateResult res{-1, 0, false};
    if (packet == 1) res = locate_packet<ND == 8>(nodes, grid, cellgeo, x, cell_nodes, padding, p, active, &counters->stack_overflow);
    else if (active) res = locate_point<ND == 8>(nodes, grid, cellgeo, x, cell_nodes, padding, p, &counters->stack_overflow);
    if (!active) continue;
    out_cell[pidx] = res.cell;
    out_ncoll[pidx] = (uint8_t)min(res.ncoll, 255);
    if (out_points) {
      out_points[3 * pidx] = p[0];
      out_points[3 * pidx + 1] = p[1];
      out_points[3 * pidx + 2] = p[2];
    }
    if (res.ncoll > 1) atomicAdd(&counters->ties, 1ULL);
    if (res.fallback) atomicAdd(&counters->fallback, 1ULL);
    int32_t cols[ND];
    double vals[ND];
    if (res.cell < 0) {
      atomicAdd(&counters->not_found, 1ULL);
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        cols[d] = 0;
        vals[d] = 0.0;
      }
    } else {
      tabulate_vals<ND>(cellgeo, res.cell, p, W, S, vals);
      load_cols<ND>(dofmap, res.cell, cols);
    }
    int32_t* oc = coo_cols + (int64_t)ND * pidx;
    double* ov = coo_vals + (int64_t)ND * pidx;
    if (ND == 4) {
      *reinterpret_cast<int4*>(oc) = make_int4(cols[0], cols[1], cols[2], cols[3 % ND]);
    } else {
#pragma unroll
      for (int d = 0; d < ND; d += 2) *reinterpret_cast<int2*>(oc + d) = make_int2(cols[d], cols[(d + 1) % ND]);
    }
#pragma unroll
    for (int d = 0; d < ND; d += 2) *reinterpret_cast<double2*>(ov + d) = make_double2(vals[d], vals[(d + 1) % ND]);
  }
}

// =================================================================================================
// COO -> CSR.  The COO is segment structured: point row r owns the SEG = nq*nd consecutive
// entries [r*SEG, (r+1)*SEG).  Matrix row rho collects the segments of all point rows with
// row_dofs[r] == rho, in ascending r; only the FIRST of them contributes values
// (first-visit-wins, interpolation.py:208-217,242), later ones contribute explicit zeros to the
// pattern.  Each row is sorted by column in shared memory with (col, position) keys — a stable
// sort — and duplicates are summed left to right, i.e. in l order (ADD_VALUES).
// =================================================================================================

__device__ __forceinline__ uint32_t next_pow2(uint32_t v) {
  return v <= 1 ? 1u : 1u << (32 - __clz(v - 1));
}

// Shared body: loads the n keys of matrix row `rho` into keys[] (shared memory, or a per-block buffer in global memory
// for rows beyond the shared-memory budget), padded to the next power of two, and sorts them.
template <bool BLOCK>
__device__ void load_sort_row(const int64_t* __restrict__ seg_ptr, const int32_t* __restrict__ seg_rows,
                              const int32_t* __restrict__ coo_cols, int seg, int64_t rho, uint32_t n, uint64_t* keys,
                              int tid, int nthreads) {
  const int64_t s0 = seg_ptr[rho];
  const uint32_t cap = n <= 2 ? 2u : next_pow2(n);  // the row's own sort size, not the launch's maximum
  for (uint32_t e = tid; e < cap; e += nthreads) {
    uint64_t key = ~0ULL;
    if (e < n) {
      const uint32_t si = e / seg;
      const uint32_t off = e - si * seg;
      const int64_t r = seg_rows[s0 + si];
      const uint32_t col = (uint32_t)coo_cols[r * seg + off];
      key = ((uint64_t)col << 32) | e;
    }
    keys[e] = key;
  }
  group_sync<BLOCK>();
  bitonic_sort_u64<BLOCK>(keys, cap, tid, nthreads);
}

// entries of matrix row rho; rows outside (n_lo, n_hi] belong to the other launch (shared- vs global-memory keys)
__device__ __forceinline__ bool row_in_class(const int64_t* __restrict__ seg_ptr, int seg, int64_t rho, int64_t n_lo, int64_t n_hi,
                                             uint32_t& n) {
  const int64_t cnt = (seg_ptr[rho + 1] - seg_ptr[rho]) * seg;
  n = (uint32_t)cnt;
  return cnt > n_lo && cnt <= n_hi;
}

// pass 1: nnz per matrix row.  BLOCK=false: one warp per row; BLOCK=true: one block per row.
template <bool BLOCK>
__global__ void rows_count_kernel(const int64_t* __restrict__ seg_ptr, const int32_t* __restrict__ seg_rows,
                                  const int32_t* __restrict__ coo_cols, int seg, int64_t n_rows, uint32_t cap,
                                  uint64_t* __restrict__ g_keys, int64_t n_lo, int64_t n_hi, int64_t* __restrict__ row_nnz) {
  extern __shared__ uint64_t s_keys[];
  const int nthreads = BLOCK ? blockDim.x : 32;
  const int tid = BLOCK ? threadIdx.x : (threadIdx.x & 31);
  const int groups_per_block = BLOCK ? 1 : blockDim.x / 32;
  const int gid = BLOCK ? 0 : threadIdx.x / 32;
  uint64_t* keys = g_keys ? g_keys + (size_t)blockIdx.x * cap : s_keys + (size_t)gid * cap;
  __shared__ int s_count;
  for (int64_t rho = (int64_t)blockIdx.x * groups_per_block + gid; rho < n_rows;
       rho += (int64_t)gridDim.x * groups_per_block) {
    uint32_t n;
    if (!row_in_class(seg_ptr, seg, rho, n_lo, n_hi, n)) continue;  // uniform over the group
    load_sort_row<BLOCK>(seg_ptr, seg_rows, coo_cols, seg, rho, n, keys, tid, nthreads);
    int local = 0;
    for (uint32_t e = tid; e < n; e += nthreads) {
      const uint32_t col = (uint32_t)(keys[e] >> 32);
      local += (e == 0 || (uint32_t)(keys[e - 1] >> 32) != col) ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if (BLOCK) {
      if (threadIdx.x == 0) s_count = 0;
      __syncthreads();
      if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_count, local);
      __syncthreads();
      if (threadIdx.x == 0) row_nnz[rho] = s_count;
      __syncthreads();
    } else {
      if (tid == 0) row_nnz[rho] = local;
      __syncwarp();
    }
  }
}

// pass 2: sorted unique columns and summed values.
template <bool BLOCK>
__global__ void rows_fill_kernel(const int64_t* __restrict__ seg_ptr, const int32_t* __restrict__ seg_rows,
                                 const int32_t* __restrict__ coo_cols, const double* __restrict__ coo_vals, int seg,
                                 int64_t n_rows, uint32_t cap, uint64_t* __restrict__ g_keys, int64_t n_lo, int64_t n_hi,
                                 const int64_t* __restrict__ indptr, int32_t* __restrict__ indices, double* __restrict__ values) {
  extern __shared__ uint64_t s_keys[];
  const int nthreads = BLOCK ? blockDim.x : 32;
  const int tid = BLOCK ? threadIdx.x : (threadIdx.x & 31);
  const int groups_per_block = BLOCK ? 1 : blockDim.x / 32;
  const int gid = BLOCK ? 0 : threadIdx.x / 32;
  uint64_t* keys = g_keys ? g_keys + (size_t)blockIdx.x * cap : s_keys + (size_t)gid * cap;
  __shared__ int s_warp_heads[32];
  for (int64_t rho = (int64_t)blockIdx.x * groups_per_block + gid; rho < n_rows;
       rho += (int64_t)gridDim.x * groups_per_block) {
    uint32_t n;
    if (!row_in_class(seg_ptr, seg, rho, n_lo, n_hi, n)) continue;  // uniform over the group
    load_sort_row<BLOCK>(seg_ptr, seg_rows, coo_cols, seg, rho, n, keys, tid, nthreads);
    const int64_t s0 = seg_ptr[rho];
    const int64_t out0 = indptr[rho];
    // heads are processed in chunks of nthreads sorted positions; running head count kept uniform
    int base = 0;
    for (uint32_t e0 = 0; e0 < n; e0 += nthreads) {
      const uint32_t e = e0 + tid;
      bool head = false;
      uint32_t col = 0;
      if (e < n) {
        col = (uint32_t)(keys[e] >> 32);
        head = (e == 0) || ((uint32_t)(keys[e - 1] >> 32) != col);
      }
      const unsigned ball = __ballot_sync(0xffffffffu, head);
      int pos = __popc(ball & ((1u << (threadIdx.x & 31)) - 1));
      int chunk_total;
      if (BLOCK) {
        const int wid = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) s_warp_heads[wid] = __popc(ball);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
          const int c = s_warp_heads[w];
          if (w < wid) before += c;
          total += c;
        }
        pos += before;
        chunk_total = total;
        __syncthreads();
      } else {
        chunk_total = __popc(ball);
      }
      if (head) {
        double sum = 0.0;
        for (uint32_t k = e; k < n && (uint32_t)(keys[k] >> 32) == col; ++k) {
          const uint32_t src = (uint32_t)(keys[k] & 0xffffffffu);
          const uint32_t si = src / seg;
          const uint32_t off = src - si * seg;
          // only the first (lowest r) segment of a row carries values: first-visit-wins
          const double v = si == 0 ? coo_vals[(int64_t)seg_rows[s0] * seg + off] : 0.0;
          sum += v;
        }
        indices[out0 + base + pos] = (int32_t)col;
        values[out0 + base + pos] = sum;
      }
      base += chunk_total;
    }
    group_sync<BLOCK>();
  }
}

// seg_count[rho] += 1 for every point row
__global__ void __launch_bounds__(256)
seg_hist_kernel(const int32_t* __restrict__ row_dofs, int64_t n_point_rows, int64_t n_rows_total,
                unsigned long long* __restrict__ seg_count, int* __restrict__ bad) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_point_rows; r += (int64_t)gridDim.x * blockDim.x) {
    const int rho = row_dofs[r];
    if (rho < 0 || rho >= n_rows_total) {
      *bad = 1;
      continue;
    }
    atomicAdd(&seg_count[rho], 1ULL);
  }
}

__global__ void __launch_bounds__(256) iota_kernel(int32_t* __restrict__ a, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a[i] = (int32_t)i;
}

void device_iota(int32_t* a, int64_t n, cudaStream_t s) {
  if (n <= 0) return;
  { iota_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(a, n); FXII_LAUNCHED(1); }
  FXII_CUDA(cudaGetLastError());
}

__global__ void __launch_bounds__(256)
max_u64_kernel(const unsigned long long* __restrict__ a, int64_t n, unsigned long long* __restrict__ out) {
  unsigned long long m = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = max(m, a[i]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// traversal flavour: per-thread walk by default.  FXII_TRAVERSAL=packet selects the warp-cooperative
// packet traversal, which was measured 2-3x SLOWER on every BASELINE config (cfg2 0.107 vs 0.038 ms,
// cfg5 38.7 vs 17.1 ms, cfg4 253 vs 120 ms): walking the union of 32 paths serialises their memory
// latencies, while 32 independent walks overlap them.  Kept for workloads with identical paths.
static int packet_mode() {
  static const int mode = [] {
    const char* e = getenv("FXII_TRAVERSAL");
    if (e && std::string(e) == "packet") return 1;
    return 0;
  }();
  return mode;
}

static float elapsed_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

// Segment-structured COO -> CSR.  coo_cols/coo_vals: [n_point_rows*seg].
void coo_segments_to_csr(const int32_t* row_dofs, int64_t n_point_rows, int64_t n_rows_total, int seg,
                         const int32_t* coo_cols, const double* coo_vals, int64_t n_cols, cudaStream_t s,
                         fxii_csr* out, float* ms_count, float* ms_fill) {
  cudaEvent_t ke[4];
  for (auto& e : ke) FXII_CUDA(cudaEventCreate(&e));
  out->n_rows = n_rows_total;
  out->n_cols = n_cols;
  DevBuf<unsigned long long> seg_count;  // [n_rows_total + 1], last slot stays 0
  seg_count.alloc(n_rows_total + 1, s);
  FXII_CUDA(cudaMemsetAsync(seg_count.p, 0, sizeof(unsigned long long) * (size_t)(n_rows_total + 1), s));
  DevBuf<int> bad;
  bad.alloc(2, s);
  FXII_CUDA(cudaMemsetAsync(bad.p, 0, 2 * sizeof(int), s));
  DevBuf<int64_t> seg_ptr;
  seg_ptr.alloc(n_rows_total + 1, s);
  DevBuf<int32_t> seg_rows, iota, keys_sorted;
  seg_rows.alloc(n_point_rows, s);
  DevBuf<unsigned long long> max_seg;
  max_seg.alloc(1, s);
  FXII_CUDA(cudaMemsetAsync(max_seg.p, 0, sizeof(unsigned long long), s));
  if (n_point_rows > 0) {
    { seg_hist_kernel<<<grid_for(n_point_rows, 256, 8), 256, 0, s>>>(row_dofs, n_point_rows, n_rows_total, seg_count.p, bad.p); FXII_LAUNCHED(1); }
    { max_u64_kernel<<<grid_for(n_rows_total, 256, 8), 256, 0, s>>>(seg_count.p, n_rows_total, max_seg.p); FXII_LAUNCHED(1); }
  }
  {
    size_t tb = 0;
    FXII_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, (const long long*)seg_count.p, (long long*)seg_ptr.p,
                                            (int64_t)(n_rows_total + 1), s));
    DevBuf<uint8_t> tmp;
    tmp.alloc((int64_t)tb, s);
    FXII_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, (const long long*)seg_count.p, (long long*)seg_ptr.p,
                                            (int64_t)(n_rows_total + 1), s));
  }
  if (n_point_rows > 0) {
    // stable sort of point rows by matrix row: ascending r inside each row
    iota.alloc(n_point_rows, s);
    keys_sorted.alloc(n_point_rows, s);
    { iota_kernel<<<grid_for(n_point_rows, 256, 8), 256, 0, s>>>(iota.p, n_point_rows); FXII_LAUNCHED(1); }
    int end_bit = 1;
    while (end_bit < 31 && ((int64_t)1 << end_bit) < n_rows_total) ++end_bit;
    size_t tb = 0;
    FXII_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, row_dofs, keys_sorted.p, iota.p, seg_rows.p, n_point_rows, 0,
                                              end_bit, s));
    DevBuf<uint8_t> tmp;
    tmp.alloc((int64_t)tb, s);
    FXII_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, row_dofs, keys_sorted.p, iota.p, seg_rows.p, n_point_rows, 0,
                                              end_bit, s));
  }
  unsigned long long h_max_seg = 0;
  int h_bad[2] = {0, 0};
  FXII_CUDA(cudaMemcpyAsync(&h_max_seg, max_seg.p, sizeof(h_max_seg), cudaMemcpyDeviceToHost, s));
  FXII_CUDA(cudaMemcpyAsync(h_bad, bad.p, sizeof(h_bad), cudaMemcpyDeviceToHost, s));
  FXII_CUDA(cudaStreamSynchronize(s));
  FXII_CHECK(h_bad[0] == 0, "row_dofs out of range [0, n_rows_total)");
  const uint64_t max_entries = h_max_seg * (uint64_t)seg;
  // Rows of up to kSmemRowEntries COO entries are sorted in shared memory.  Longer rows (a K dof shared by hundreds of
  // point rows) take the same kernels with their keys in a per-block buffer in GLOBAL memory — slower, never truncated.
  constexpr uint64_t kSmemRowEntries = 16384;
  FXII_CHECK(max_entries < ((uint64_t)1 << 31), "a matrix row collects 2^31 or more COO entries");
  uint32_t cap = 1;
  while (cap < std::min(max_entries, kSmemRowEntries)) cap <<= 1;
  if (cap < 2) cap = 2;
  const int64_t n_hi_smem = (int64_t)kSmemRowEntries;
  uint32_t g_cap = 0;
  int g_grid = 0;
  DevBuf<uint64_t> g_keys;
  if (max_entries > kSmemRowEntries) {
    g_cap = 1;
    while (g_cap < max_entries) g_cap <<= 1;
    // one buffer per block; at most 2 GiB of keys in total
    g_grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)num_sms(), ((int64_t)1 << 28) / (int64_t)g_cap));
    g_keys.alloc((int64_t)g_grid * g_cap, s);
  }

  DevBuf<int64_t> row_nnz;
  row_nnz.alloc(n_rows_total + 1, s);
  FXII_CUDA(cudaMemsetAsync(row_nnz.p, 0, sizeof(int64_t) * (size_t)(n_rows_total + 1), s));
  out->indptr.alloc(n_rows_total + 1, s);
  const bool block_mode = cap > 64;
  int threads, grid;
  size_t smem;
  if (block_mode) {
    threads = cap >= 512 ? 256 : 128;
    smem = (size_t)cap * sizeof(uint64_t);
    grid = (int)std::min<int64_t>(n_rows_total, (int64_t)num_sms() * 8);
    if (smem > 48 * 1024) {
      FXII_CUDA(cudaFuncSetAttribute(rows_count_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      FXII_CUDA(cudaFuncSetAttribute(rows_fill_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
  } else {
    threads = 256;
    smem = (size_t)cap * sizeof(uint64_t) * (threads / 32);
    grid = (int)std::min<int64_t>((n_rows_total + 7) / 8, (int64_t)num_sms() * 8);
  }
  if (grid < 1) grid = 1;
  FXII_CUDA(cudaEventRecord(ke[0], s));
  if (n_rows_total > 0) {
    if (block_mode)
      { rows_count_kernel<true><<<grid, threads, smem, s>>>(seg_ptr.p, seg_rows.p, coo_cols, seg, n_rows_total, cap, nullptr, -1, n_hi_smem, row_nnz.p); FXII_LAUNCHED(1); }
    else
      { rows_count_kernel<false><<<grid, threads, smem, s>>>(seg_ptr.p, seg_rows.p, coo_cols, seg, n_rows_total, cap, nullptr, -1, n_hi_smem, row_nnz.p); FXII_LAUNCHED(1); }
    if (g_cap)
      { rows_count_kernel<true><<<g_grid, 256, 0, s>>>(seg_ptr.p, seg_rows.p, coo_cols, seg, n_rows_total, g_cap, g_keys.p, n_hi_smem, INT64_MAX, row_nnz.p); FXII_LAUNCHED(1); }
    FXII_CUDA(cudaGetLastError());
  }
  FXII_CUDA(cudaEventRecord(ke[1], s));
  {
    size_t tb = 0;
    FXII_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, (const long long*)row_nnz.p, (long long*)out->indptr.p,
                                            (int64_t)(n_rows_total + 1), s));
    DevBuf<uint8_t> tmp;
    tmp.alloc((int64_t)tb, s);
    FXII_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, (const long long*)row_nnz.p, (long long*)out->indptr.p,
                                            (int64_t)(n_rows_total + 1), s));
  }
  int64_t nnz = 0;
  FXII_CUDA(cudaMemcpyAsync(&nnz, out->indptr.p + n_rows_total, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  FXII_CUDA(cudaStreamSynchronize(s));
  out->nnz = nnz;
  out->indices.alloc(nnz, s);
  out->values.alloc(nnz, s);
  FXII_CUDA(cudaEventRecord(ke[2], s));
  if (n_rows_total > 0 && nnz > 0) {
    if (block_mode)
      { rows_fill_kernel<true><<<grid, threads, smem, s>>>(seg_ptr.p, seg_rows.p, coo_cols, coo_vals, seg, n_rows_total, cap, nullptr, -1,
                                                         n_hi_smem, out->indptr.p, out->indices.p, out->values.p); FXII_LAUNCHED(1); }
    else
      { rows_fill_kernel<false><<<grid, threads, smem, s>>>(seg_ptr.p, seg_rows.p, coo_cols, coo_vals, seg, n_rows_total, cap, nullptr, -1,
                                                          n_hi_smem, out->indptr.p, out->indices.p, out->values.p); FXII_LAUNCHED(1); }
    if (g_cap)
      { rows_fill_kernel<true><<<g_grid, 256, 0, s>>>(seg_ptr.p, seg_rows.p, coo_cols, coo_vals, seg, n_rows_total, g_cap, g_keys.p,
                                                     n_hi_smem, INT64_MAX, out->indptr.p, out->indices.p, out->values.p); FXII_LAUNCHED(1); }
    FXII_CUDA(cudaGetLastError());
  }
  FXII_CUDA(cudaEventRecord(ke[3], s));
  FXII_CUDA(cudaEventSynchronize(ke[3]));
  if (ms_count) *ms_count = elapsed_ms(ke[0], ke[1]);
  if (ms_fill) *ms_fill = elapsed_ms(ke[2], ke[3]);
  for (auto& e : ke) cudaEventDestroy(e);
}

// =================================================================================================
// FUSED PATH.  When every matrix row is produced by exactly one point row (cell-local K: quadrature
// or DG elements — all BASELINE configs), the block that located the points of a row also reduces it:
// the COO never reaches HBM.  A TEAM of threads owns one point row (team = pow2 >= nq inside a warp,
// or whole warps for nq > 32):
//   A  locate + tabulate: every thread its point                     -> s_cell, s_val
//   B  cell-level dedupe: the first point of every distinct cell sums the values of all points of
//      that cell in l order (points of one cell share the column set, so nd entries collapse at once)
//   C  leaders emit nd (column, value) entries                       -> s_ekey / s_eval
//   D  team-local bitonic sort of the (column, slot) keys — stable
//   E  duplicates (dofs shared by neighbouring cells) are summed left to right; sorted unique
//      entries go to the row's slice of a staging buffer, the row length to row_cnt
// then one scan and a compaction kernel produce the CSR.  Summation order differs from the
// unfused path only in association (per cell first); it is fixed, so results are reproducible.
// =================================================================================================
// Barrier of one team.  Multi-warp teams use named barriers with COMPILE-TIME ids: a register id makes
// ptxas reserve all 16 hardware barriers for the CTA, which capped residency at 4 CTAs per SM
// (measured: 24 % achieved occupancy instead of 50 %).  At most 4 multi-warp teams fit a block.
__device__ __forceinline__ void team_sync(int team, int team_id) {
  if (team <= 32) {
    __syncwarp();
    return;
  }
  switch (team_id) {
    case 0: asm volatile("bar.sync 1, %0;" ::"r"(team) : "memory"); break;
    case 1: asm volatile("bar.sync 2, %0;" ::"r"(team) : "memory"); break;
    case 2: asm volatile("bar.sync 3, %0;" ::"r"(team) : "memory"); break;
    default: asm volatile("bar.sync 4, %0;" ::"r"(team) : "memory"); break;
  }
}

// exclusive prefix of a flag over the threads of a team (+ team total); two team barriers for
// multi-warp teams
__device__ __forceinline__ int team_prefix(bool flag, int team, int team_id, int l, int* s_wcnt, int& total) {
  const unsigned ball = __ballot_sync(0xffffffffu, flag);
  const int lane = threadIdx.x & 31;
  if (team <= 32) {
    const unsigned tmask = (team == 32 ? 0xffffffffu : ((1u << team) - 1u)) << (lane & ~(team - 1));
    total = __popc(ball & tmask);
    return __popc(ball & tmask & ((1u << lane) - 1u));
  }
  const int warp = threadIdx.x >> 5;
  const int wpt = team >> 5;               // warps per team
  const int w0 = team_id * wpt;            // first warp of the team
  if (lane == 0) s_wcnt[warp] = __popc(ball);
  team_sync(team, team_id);
  int before = 0, tot = 0;
  for (int w = 0; w < wpt; ++w) {
    const int c = s_wcnt[w0 + w];
    if (w0 + w < warp) before += c;
    tot += c;
  }
  team_sync(team, team_id);
  total = tot;
  return before + __popc(ball & ((1u << lane) - 1u));
}

// what a build reports back to the host (one device-to-host copy)
struct FusedResult {
  int64_t nnz;
  PiCounters counters;
  int conflict;
  int pad;
  unsigned long long ns_rows, ns_csr;  // single-kernel path: device time of the two phases (%globaltimer)
};

// Extra arguments of the single-kernel (cooperative) variant: the inputs of the row frames, which it
// computes itself, and the outputs of the scan + compaction phase that follows a grid-wide barrier.
struct CoopArgs {
  const double* gx;
  const int32_t* gcells;
  const int32_t* cells_k;
  const double* xref;
  const double* radius;
  int nip, radius_per_row;
  unsigned int* barrier;  // arrival counter, zeroed with the workspace
  unsigned long long* ticket;  // next unit of point rows, zeroed with the workspace
  int ticket_batch;            // units handed out per atomic
  int64_t* cta_sum;       // [gridDim.x] entries per slice of matrix rows, zeroed with the workspace
  int64_t* indptr;
  int32_t* indices;
  double* values;
  int64_t capacity;       // entries indices/values can hold; 0: count only
  FusedResult* result;
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Barrier over all CTAs of a cooperative launch (co-residency guaranteed by the launch API).
// `target` = gridDim.x * (number of barriers passed so far + 1); the counter only grows.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v < target) __nanosleep(40);
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

// sum over the block (every thread gets it); s_red: 32 slots
__device__ __forceinline__ int64_t block_sum_i64(int64_t v, int64_t* s_red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  int64_t t = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
  __syncthreads();
  return t;
}

// frame + point in one out-of-line call: its ~30 live doubles stay out of the traversal's register budget.
// Everything goes in and out BY VALUE: a pointer or reference argument would pin the caller's point (and a copy of
// the kernel's argument record) in local memory for the whole walk.
struct PointWS {
  double p0, p1, p2, W, S;
};
__device__ __noinline__ PointWS make_point_with_frame(const double* __restrict__ gx, const int32_t* __restrict__ gcells,
                                                      const int32_t* __restrict__ cells_k, const double* __restrict__ xref,
                                                      int nip, const double* __restrict__ radius, int radius_per_row, int kind,
                                                      const double* s_tab, const double* __restrict__ pts_in,
                                                      const double* __restrict__ w_in, const double* __restrict__ s_in,
                                                      int64_t r, int l, int64_t pidx) {
  double fr[kFrame];
  if (kind != FXII_OP_POINTS) compute_frame(gx, gcells, cells_k, xref, nip, radius, radius_per_row, kind, r, fr);
  double p[3], W, S;
  make_point(kind, fr, s_tab, pts_in, w_in, s_in, r, l, pidx, p, W, S);
  return PointWS{p[0], p[1], p[2], W, S};
}

// Bitonic sort of the 4*team keys of a sub-warp team held four per lane (element e = 4*l + r): the
// steps that pair elements of different lanes exchange them with shuffles, the others swap registers.
// Replaces ~120 conflicted 64-bit shared-memory accesses per lane (team of 8) by 48 shuffles.
__device__ __forceinline__ void cswap(uint64_t& a, uint64_t& b, bool up) {
  const bool sw = (a > b) == up;
  const uint64_t lo = sw ? b : a, hi = sw ? a : b;
  a = lo;
  b = hi;
}
__device__ __forceinline__ void team_sort_regs4(uint64_t (&key)[4], int team, int l) {
  const uint32_t S = 4u * (uint32_t)team;
  for (uint32_t k = 2; k <= S; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      if (j >= 4) {
        const uint32_t lm = j >> 2;                 // partner lane = l ^ lm (inside the team)
        const bool up = ((4u * (uint32_t)l) & k) == 0;  // k >= 8 here: the direction depends on the lane only
        const bool lower = ((uint32_t)l & lm) == 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const uint64_t other = __shfl_xor_sync(0xffffffffu, key[r], (int)lm);
          const bool take_min = lower == up;
          const bool other_smaller = other < key[r];
          key[r] = (take_min == other_smaller) ? other : key[r];
        }
      } else if (j == 2) {
        // pairs (0,2), (1,3); direction: k == 4 -> by lane parity, k >= 8 -> by lane
        const bool up = ((4u * (uint32_t)l) & k) == 0;  // k >= 4: bits of r (0..3) never reach k
        cswap(key[0], key[2], up);
        cswap(key[1], key[3], up);
      } else {  // j == 1: pairs (0,1), (2,3)
        if (k == 2) {  // direction alternates inside the lane: e & 2
          cswap(key[0], key[1], true);
          cswap(key[2], key[3], false);
        } else {
          const bool up = ((4u * (uint32_t)l) & k) == 0;
          cswap(key[0], key[1], up);
          cswap(key[2], key[3], up);
        }
      }
    }
  }
}

#ifndef FXII_FUSED_MINBLOCKS
#define FXII_FUSED_MINBLOCKS 4
#endif
// COOP: launched cooperatively with one wave of CTAs.  The kernel then also computes the row frames
// (every lane of a team recomputes its row's frame in registers: SIMT makes that as cheap as one lane
// doing it) and, after a grid-wide barrier, scans the row lengths and copies the staged rows into the
// CSR — memset + ONE kernel per build instead of memset + frames + rows + 2 scan kernels + compaction.
template <int ND, int MODE>  // 0: rows only (frames, scan and compaction are separate kernels); 2: cooperative single kernel
__global__ void __launch_bounds__(256, FXII_FUSED_MINBLOCKS)
pi_fused_kernel(const WideNode* __restrict__ nodes, const SceneGrid grid, const double* __restrict__ cellgeo,
                const double* __restrict__ x,
                const int4* __restrict__ cell_nodes, double padding, const int32_t* __restrict__ dofmap,
                const double* __restrict__ frames, const double* __restrict__ table, const double* __restrict__ wq,
                const double* __restrict__ pts_in, const double* __restrict__ w_in, const double* __restrict__ s_in,
                int kind, int nq, int packet, int team, uint32_t scap, int64_t n_point_rows, int64_t n_rows_total,
                const int32_t* __restrict__ row_dofs,
                int32_t* __restrict__ out_cell, uint8_t* __restrict__ out_ncoll, double* __restrict__ out_points,
                int32_t* __restrict__ stage_col, double* __restrict__ stage_val, int64_t* __restrict__ row_cnt,
                int32_t* __restrict__ src_row, int* __restrict__ conflict, PiCounters* counters, const CoopArgs ca) {
  constexpr bool COOP = MODE == 2;
  constexpr bool FRAMES = MODE >= 1;
  unsigned long long t_begin = 0;
  if (COOP && blockIdx.x == 0 && threadIdx.x == 0) t_begin = global_timer_ns();
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int B = blockDim.x;
  const int rpb = B / team;  // point rows per block
  // layout: s_tab f64[nq*4] | s_val f64[B*ND] | s_eval f64[B*ND] | s_ekey u64[rpb*scap] | s_cell i32[B] | s_wcnt i32[B/32]
  double* s_tab = reinterpret_cast<double*>(s_raw);
  double* s_val = s_tab + 4 * (size_t)nq;
  double* s_eval = s_val + (size_t)B * ND;
  uint64_t* s_ekey = reinterpret_cast<uint64_t*>(s_eval + (size_t)B * ND + 2 * (size_t)rpb);
  int32_t* s_cell = reinterpret_cast<int32_t*>(s_ekey + (size_t)rpb * (scap + 2));
  int* s_wcnt = s_cell + B;
  if (kind == FXII_OP_CIRCLE || kind == FXII_OP_DISK) {
    for (int k = threadIdx.x; k < nq; k += B) {
      s_tab[4 * k + 0] = table[3 * k + 0];
      s_tab[4 * k + 1] = table[3 * k + 1];
      s_tab[4 * k + 2] = table[3 * k + 2];
      s_tab[4 * k + 3] = wq[k];
    }
  }
  __syncthreads();
  const int tid = threadIdx.x;
  const int team_id = tid / team;
  const int l = tid - team_id * team;
  const int t0 = team_id * team;           // first thread of the team
  // team regions are padded by two entries: power-of-two strides would put the teams of a warp on the same banks
  uint64_t* ekey = s_ekey + (size_t)team_id * (scap + 2);
  double* eval = s_eval + (size_t)team_id * ((size_t)team * ND + 2);
  const int seg = nq * ND;
  const bool in_team = team_id < rpb;      // threads beyond the last whole team idle (B % team != 0 never happens)
  const int64_t n_groups = (n_point_rows + rpb - 1) / rpb;
  // COOP: rows are handed out dynamically, one ticket per warp (sub-warp teams: 32/team rows) or per
  // multi-warp team — the single wave of CTAs has no block scheduler behind it to even out slow rows
  __shared__ long long s_ticket[8];
  const int unit_rows = team >= 32 ? 1 : 32 / team;
  const int team_in_unit = team >= 32 ? 0 : (tid & 31) / team;
  // COOP: matrix rows are cut into gridDim.x contiguous slices for the scan phase; the entry count of
  // every slice is accumulated here, as rows finish, so that phase needs no second grid barrier
  const int64_t slice_rows = COOP ? max((int64_t)1, (n_rows_total + gridDim.x - 1) / gridDim.x) : 1;
  int64_t rb = blockIdx.x;
  long long tk_next = 0, tk_end = 0;  // COOP: the batch of tickets this warp / team currently owns
  while (true) {
    int64_t r;
    if (COOP) {
      // one atomic buys `ticket_batch` consecutive units: a single counter serialises at a few ns per atomic, which
      // one-point rows (a unit is ~10 us of work) would otherwise be bound by
      if (tk_next == tk_end) {
        long long tk = 0;
        if (team <= 32) {
          if ((tid & 31) == 0) tk = (long long)atomicAdd(ca.ticket, (unsigned long long)ca.ticket_batch);
          tk = __shfl_sync(0xffffffffu, tk, 0);
        } else {
          if (l == 0) s_ticket[team_id & 7] = (long long)atomicAdd(ca.ticket, (unsigned long long)ca.ticket_batch);
          team_sync(team, team_id);
          tk = s_ticket[team_id & 7];
        }
        tk_next = tk;
        tk_end = tk + ca.ticket_batch;
      }
      const long long tk = tk_next++;
      if (tk * unit_rows >= n_point_rows) break;
      r = tk * unit_rows + team_in_unit;
    } else {
      if (rb >= n_groups) break;
      r = rb * rpb + team_id;
      rb += gridDim.x;
    }
    const bool active = in_team && r < n_point_rows && l < nq;
    // ---- A: locate + tabulate ------------------------------------------------------------------
    int cell = -1;
    double vals[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) vals[d] = 0.0;
    {
      const int64_t pidx = active ? r * nq + l : 0;
      double p[3] = {0.0, 0.0, 0.0}, W = 1.0, S = 1.0;
      if (active) {
        if (FRAMES) {
          const PointWS pw = make_point_with_frame(ca.gx, ca.gcells, ca.cells_k, ca.xref, ca.nip, ca.radius, ca.radius_per_row,
                                                   kind, s_tab, pts_in, w_in, s_in, r, l, pidx);
          p[0] = pw.p0, p[1] = pw.p1, p[2] = pw.p2, W = pw.W, S = pw.S;
        } else {
          make_point(kind, frames + kFrame * r, s_tab, pts_in, w_in, s_in, r, l, pidx, p, W, S);
        }
      }
      // (the packet traversal is only kept in the unfused kernel: a second inlined walk costs this kernel registers —
      // 12.2 -> 11.0 ms on the 1M-segment tree together with passing the point by value to the slow-path distance.
      // A batched walk that first collects candidate cells and tests them in a second loop measured 12.2 ms.)
      LocateResult res{-1, 0, false};
      if (active) res = locate_point<ND == 8>(nodes, grid, cellgeo, x, cell_nodes, padding, p, &counters->stack_overflow);
      if (active) {
        out_cell[pidx] = res.cell;
        out_ncoll[pidx] = (uint8_t)min(res.ncoll, 255);
        if (out_points) {
          out_points[3 * pidx] = p[0];
          out_points[3 * pidx + 1] = p[1];
          out_points[3 * pidx + 2] = p[2];
        }
        if (res.ncoll > 1) atomicAdd(&counters->ties, 1ULL);
        if (res.fallback) atomicAdd(&counters->fallback, 1ULL);
        cell = res.cell;
        if (cell < 0) atomicAdd(&counters->not_found, 1ULL);
        else tabulate_vals<ND>(cellgeo, cell, p, W, S, vals);
      }
    }
    if (team == 1) {
      // one point per row (PointwiseTrace, mapped points): the row is the cell's nd (column, value)
      // pairs — distinct columns — sorted by an odd-even transposition network in registers
      if (in_team && r < n_point_rows) {
        int32_t cols[ND];
        int nrow = 0;
        if (cell >= 0) {
          load_cols<ND>(dofmap, cell, cols);
#pragma unroll
          for (int pass = 0; pass < ND; ++pass) {
#pragma unroll
            for (int i = pass & 1; i + 1 < ND; i += 2) {
              if (cols[i] > cols[i + 1]) {
                const int32_t tc = cols[i]; cols[i] = cols[i + 1]; cols[i + 1] = tc;
                const double tv = vals[i]; vals[i] = vals[i + 1]; vals[i + 1] = tv;
              }
            }
          }
          const int64_t o = r * (int64_t)seg;
#pragma unroll
          for (int d = 0; d < ND; ++d) {
            stage_col[o + d] = cols[d];
            stage_val[o + d] = vals[d];
          }
          nrow = ND;
        }
        const int rho = row_dofs[r];
        if (rho < 0 || rho >= n_rows_total) *conflict = 2;
        else if (atomicCAS(&src_row[rho], 0, (int32_t)r + 1) != 0) *conflict = 1;
        else {
          row_cnt[rho] = nrow;
          if (COOP && nrow) atomicAdd((unsigned long long*)&ca.cta_sum[rho / slice_rows], (unsigned long long)nrow);
        }
      }
      continue;
    }
    s_cell[tid] = cell;
#pragma unroll
    for (int d = 0; d < ND; ++d) s_val[(size_t)d * B + tid] = vals[d];  // [d][thread]: conflict-free
    team_sync(team, team_id);
    // ---- B: first point of every distinct cell sums that cell's points in l order ----------------
    bool leader = cell >= 0;
    if (team <= 32) {
      // sub-warp teams: one MATCH gives the lanes of the team that found the same cell; the lowest is the leader and
      // adds the others in ascending lane (= l) order — the same sums as the scan below, without its O(nq) loops
      const int lane = tid & 31;
      const unsigned tmask = (team == 32 ? 0xffffffffu : ((1u << team) - 1u)) << (lane & ~(team - 1));
      const unsigned same = __match_any_sync(0xffffffffu, cell) & tmask;
      leader = leader && (lane == __ffs(same) - 1);
      if (leader) {
        unsigned rest = same & ~(1u << lane);
        const int w0 = tid - lane;  // first thread of this warp
        while (rest) {
          const int j = __ffs(rest) - 1;
          rest &= rest - 1;
#pragma unroll
          for (int d = 0; d < ND; ++d) vals[d] += s_val[(size_t)d * B + w0 + j];
        }
      }
    } else {
      for (int j = 0; j < l && leader; ++j) leader = s_cell[t0 + j] != cell;
      if (leader) {
        for (int j = l + 1; j < nq; ++j) {
          if (s_cell[t0 + j] == cell) {
#pragma unroll
            for (int d = 0; d < ND; ++d) vals[d] += s_val[(size_t)d * B + t0 + j];
          }
        }
      }
    }
    // ---- C: leaders emit nd entries ---------------------------------------------------------------
    int n_cells;
    const int u = team_prefix(leader, team, team_id, l, s_wcnt, n_cells);
    const uint32_t M = (uint32_t)n_cells * ND;
    const bool reg_sort = ND == 4 && team <= 32;  // sub-warp P1 teams sort in registers (team >= 2 here)
    uint32_t S = 2;
    if (reg_sort) {
      S = 4u * (uint32_t)team;
    } else {
      while (S < M) S <<= 1;
      if (team < 32) {  // teams sharing a warp iterate in lockstep: use the warp-wide maximum
        for (int o = 16; o > 0; o >>= 1) S = max(S, __shfl_xor_sync(0xffffffffu, S, o));
      }
    }
    for (uint32_t e = M + l; e < S; e += team) ekey[e] = ~0ULL;
    if (leader) {
      int32_t cols[ND];
      load_cols<ND>(dofmap, cell, cols);
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        const uint32_t idx = (uint32_t)u * ND + d;
        ekey[idx] = ((uint64_t)(uint32_t)cols[d] << 32) | idx;
        eval[idx] = vals[d];
      }
    }
    team_sync(team, team_id);
    // ---- D: team-local bitonic sort -----------------------------------------------------------------
    if (reg_sort) {
      uint64_t key[4];
      const ulonglong2 k01 = *reinterpret_cast<const ulonglong2*>(ekey + 4 * l);
      const ulonglong2 k23 = *reinterpret_cast<const ulonglong2*>(ekey + 4 * l + 2);
      key[0] = k01.x, key[1] = k01.y, key[2] = k23.x, key[3] = k23.y;
      team_sort_regs4(key, team, l);
      *reinterpret_cast<ulonglong2*>(ekey + 4 * l) = make_ulonglong2(key[0], key[1]);
      *reinterpret_cast<ulonglong2*>(ekey + 4 * l + 2) = make_ulonglong2(key[2], key[3]);
      team_sync(team, team_id);
    } else
    for (uint32_t k = 2; k <= S; k <<= 1) {
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        for (uint32_t t = l; t < S / 2; t += team) {
          const uint32_t i = 2 * t - (t & (j - 1));
          const uint32_t ixj = i | j;
          const bool up = (i & k) == 0;
          const uint64_t a = ekey[i], b = ekey[ixj];
          if ((a > b) == up) {
            ekey[i] = b;
            ekey[ixj] = a;
          }
        }
        team_sync(team, team_id);
      }
    }
    // ---- E: heads, left-to-right sums, staging ------------------------------------------------------
    uint32_t Mloop = M;
    if (team < 32) {
      for (int o = 16; o > 0; o >>= 1) Mloop = max(Mloop, __shfl_xor_sync(0xffffffffu, Mloop, o));
    }
    int base = 0;
    const int64_t out0 = r * (int64_t)seg;
    for (uint32_t e0 = 0; e0 < Mloop; e0 += team) {
      const uint32_t e = e0 + l;
      bool head = false;
      uint32_t col = 0;
      if (e < M) {
        col = (uint32_t)(ekey[e] >> 32);
        head = (e == 0) || ((uint32_t)(ekey[e - 1] >> 32) != col);
      }
      int chunk_total;
      const int pos = team_prefix(head, team, team_id, l, s_wcnt, chunk_total);
      if (head) {
        double sum = 0.0;
        for (uint32_t k = e; k < M && (uint32_t)(ekey[k] >> 32) == col; ++k) sum += eval[(uint32_t)(ekey[k] & 0xffffffffu)];
        stage_col[out0 + base + pos] = (int32_t)col;
        stage_val[out0 + base + pos] = sum;
      }
      base += chunk_total;
    }
    if (in_team && r < n_point_rows && l == 0) {
      const int rho = row_dofs[r];
      if (rho < 0 || rho >= n_rows_total) *conflict = 2;  // bad row index: reported by the general path
      else if (atomicCAS(&src_row[rho], 0, (int32_t)r + 1) != 0) *conflict = 1;  // row produced twice: not cell-local
      else {
        row_cnt[rho] = base;
        if (COOP && base) atomicAdd((unsigned long long*)&ca.cta_sum[rho / slice_rows], (unsigned long long)base);
      }
    }
    team_sync(team, team_id);  // s_cell / s_val / ekey are rewritten by the next row group
  }
  if (!COOP) return;
  // ================= scan + compaction (cooperative launch only) ==================================
  grid_barrier(ca.barrier, gridDim.x);
  unsigned long long t_rows = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) t_rows = global_timer_ns();
  const int conf = __ldcg(conflict);
  if (conf != 0) {  // K dofs shared between point rows: the host reruns the general path
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      ca.result->conflict = conf;
      ca.result->counters.not_found = __ldcg(&counters->not_found);
      ca.result->counters.ties = __ldcg(&counters->ties);
      ca.result->counters.fallback = __ldcg(&counters->fallback);
      ca.result->counters.stack_overflow = __ldcg(&counters->stack_overflow);
    }
    return;
  }
  // shared memory is reused: s_red i64[32] | s_rel i32[B] | s_src i32[B] | s_w i32[32]
  int64_t* s_red = reinterpret_cast<int64_t*>(s_raw);
  int32_t* s_rel = reinterpret_cast<int32_t*>(s_red + 32);
  int32_t* s_src = s_rel + B;
  int32_t* s_w = s_src + B;
  const int lane = tid & 31, warp = tid >> 5, n_warps = B >> 5;
  int32_t* __restrict__ out_idx = ca.indices;
  double* __restrict__ out_val = ca.values;
  // contiguous slice of matrix rows per CTA; its base offset = the slice totals before it
  const int64_t r0 = min(n_rows_total, (int64_t)blockIdx.x * slice_rows), r1 = min(n_rows_total, r0 + slice_rows);
  int64_t before = 0, total = 0;
  for (int i = tid; i < (int)gridDim.x; i += B) {
    const int64_t v = __ldcg(&ca.cta_sum[i]);
    total += v;
    if (i < (int)blockIdx.x) before += v;
  }
  before = block_sum_i64(before, s_red);
  total = block_sum_i64(total, s_red);
  const bool fits = total <= ca.capacity;  // otherwise the host compacts into exact-size arrays
  int64_t run = before;
  for (int64_t tile0 = r0; tile0 < r1; tile0 += B) {
    const int64_t rho = tile0 + tid;
    const int c = rho < r1 ? (int)__ldcg(&row_cnt[rho]) : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int wbefore = 0, tile_total = 0;
    for (int w = 0; w < n_warps; ++w) {
      const int v = s_w[w];
      if (w < warp) wbefore += v;
      tile_total += v;
    }
    const int rel = wbefore + incl - c;  // offset of this row inside the tile's output range
    if (rho < r1) ca.indptr[rho] = run + rel;
    s_rel[tid] = rel;
    s_src[tid] = rho < r1 ? __ldcg(&src_row[rho]) - 1 : -1;
    __syncthreads();
    if (fits) {
      // the tile's entries are copied flat: consecutive threads write consecutive CSR slots; the source
      // row is the last one whose offset is <= the slot (empty rows share the offset of their successor).
      // Four independent entries per thread keep enough loads in flight.
      for (int e0 = tid; e0 < tile_total; e0 += 4 * B) {
        int32_t cv[4];
        double vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u * B;
          if (e < tile_total) {
            int lo = 0, hi = B - 1;
            while (lo < hi) {
              const int mid = (lo + hi + 1) >> 1;
              if (s_rel[mid] <= e) lo = mid;
              else hi = mid - 1;
            }
            const int64_t src = (int64_t)s_src[lo] * seg + (e - s_rel[lo]);
            cv[u] = __ldcg(&stage_col[src]);
            vv[u] = __ldcg(&stage_val[src]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u * B;
          if (e < tile_total) {
            out_idx[run + e] = cv[u];
            out_val[run + e] = vv[u];
          }
        }
      }
    }
    __syncthreads();
    run += tile_total;
  }
  if (blockIdx.x == 0 && tid == 0) {
    ca.indptr[n_rows_total] = total;
    ca.result->nnz = total;
    ca.result->conflict = 0;
    ca.result->counters.not_found = __ldcg(&counters->not_found);
    ca.result->counters.ties = __ldcg(&counters->ties);
    ca.result->counters.fallback = __ldcg(&counters->fallback);
    ca.result->counters.stack_overflow = __ldcg(&counters->stack_overflow);
    ca.result->ns_rows = t_rows - t_begin;
    ca.result->ns_csr = global_timer_ns() - t_rows;
  }
}

// staging -> CSR: a group of G lanes copies one matrix row.  Nothing is written when the matrix does
// not fit `capacity` entries (the host then repeats the copy into exact-size arrays).
__global__ void __launch_bounds__(256)
compact_rows_kernel(const int32_t* __restrict__ src_row, const int64_t* __restrict__ indptr, const int32_t* __restrict__ stage_col,
                    const double* __restrict__ stage_val, int seg, int64_t n_rows, int G, int64_t capacity,
                    int32_t* __restrict__ indices, double* __restrict__ values) {
  if (indptr[n_rows] > capacity) return;
  const int64_t gid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / G;
  const int lg = threadIdx.x % G;
  const int64_t n_groups = ((int64_t)gridDim.x * blockDim.x) / G;
  for (int64_t rho = gid; rho < n_rows; rho += n_groups) {
    const int32_t r = src_row[rho] - 1;  // 0 = no point row produced this matrix row
    if (r < 0) continue;
    const int64_t o0 = indptr[rho];
    const int n = (int)(indptr[rho + 1] - o0);
    const int64_t s0 = (int64_t)r * seg;
    for (int k = lg; k < n; k += G) {
      indices[o0 + k] = stage_col[s0 + k];
      values[o0 + k] = stage_val[s0 + k];
    }
  }
}

__global__ void publish_result_kernel(FusedResult* __restrict__ host_result, const int64_t* __restrict__ nnz,
                                      const int* __restrict__ conflict, const PiCounters* __restrict__ counters) {
  if (threadIdx.x == 0) {
    host_result->nnz = *nnz;
    host_result->conflict = *conflict;
    host_result->counters = *counters;
    __threadfence_system();
  }
}

// ---- host side of the fused path -------------------------------------------------------------
struct FusedPlan {
  int team, B, rpb, grid;
  uint32_t scap;
  size_t smem;
  size_t off_counters, off_conflict, off_src, ws_bytes;  // workspace layout
  size_t off_barrier, off_result, off_cta;               // single-kernel path
  int max_grid;
  size_t scan_tmp_bytes;
};

// false: the fused path does not apply (nq > 256, too much shared memory, empty)
static bool make_fused_plan(const fxii_space* sp, const fxii_rows* rows, FusedPlan& p) {
  const int nd = sp->nd, nq = rows->nq;
  const int64_t Nr = rows->n_point_rows, E = Nr * nq * nd, n_rows = rows->n_rows_total;
  if (nq > 256 || Nr == 0 || E >= ((int64_t)1 << 40)) return false;
  p.team = 1;
  if (nq <= 32) {
    while (p.team < nq) p.team <<= 1;
  } else {
    p.team = 32 * ((nq + 31) / 32);
  }
  // block size: whole teams; 128 threads unless a team needs more
  p.B = p.team <= 128 ? 128 : 256;
  if (p.team > 32 && p.B % p.team != 0) p.B = p.team * (256 / p.team > 0 ? 256 / p.team : 1);
  if (p.B > 256 || p.B % 32 != 0) return false;
  p.rpb = p.B / p.team;
  p.scap = 2;
  while (p.scap < (uint32_t)p.team * nd) p.scap <<= 1;
  p.smem = sizeof(double) * (4 * (size_t)nq + 2 * (size_t)p.B * nd + 2 * (size_t)p.rpb) +
           sizeof(uint64_t) * (size_t)p.rpb * (p.scap + 2) + sizeof(int32_t) * (size_t)(p.B + p.B / 32 + 1);
  if (p.smem > 200 * 1024) return false;
  const int64_t n_groups = (Nr + p.rpb - 1) / p.rpb;
  // grid-stride over row groups: 2048 resident threads per SM at most, two waves of blocks so that
  // the tail of slow row groups is shared out
  p.grid = (int)std::min<int64_t>(n_groups, (int64_t)num_sms() * (2048 / p.B) * 2);
  // workspace, cleared by ONE memset: row_cnt i64[n_rows+1] | counters | conflict | src_row i32[n_rows]
  p.off_counters = sizeof(int64_t) * (size_t)(n_rows + 1);
  p.off_conflict = p.off_counters + sizeof(PiCounters);
  p.off_src = p.off_conflict + 8;
  p.ws_bytes = p.off_src + sizeof(int32_t) * (size_t)n_rows;
  // single-kernel path: barrier counter | result | per-CTA sums (one wave: at most 16 CTAs of 128 threads per SM)
  p.max_grid = num_sms() * 16;
  p.off_barrier = (p.ws_bytes + 15) & ~(size_t)15;
  p.off_result = p.off_barrier + 16;
  p.off_cta = p.off_result + ((sizeof(FusedResult) + 15) & ~(size_t)15);
  p.ws_bytes = p.off_cta + sizeof(int64_t) * (size_t)p.max_grid;
  p.scan_tmp_bytes = 0;
  FXII_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, p.scan_tmp_bytes, (const long long*)nullptr, (long long*)nullptr,
                                          (int64_t)(n_rows + 1), (cudaStream_t)0));
  return true;
}

static void launch_frames(fxii_rows* rows, cudaStream_t s) {
  const int64_t Nr = rows->n_point_rows;
  if (rows->kind == FXII_OP_POINTS || Nr == 0) return;
  { row_frames_kernel<<<grid_for(Nr, 128, 16), 128, 0, s>>>(rows->gx.p, rows->gcells.p, rows->cells_k.p, rows->xref.p, rows->nip,
                                                          rows->radius.p, rows->radius_per_row, rows->kind, Nr, rows->frames.p); FXII_LAUNCHED(1); }
  FXII_CUDA(cudaGetLastError());
}

static void launch_compact(const fxii_space* sp, const fxii_rows* rows, const uint8_t* ws, const FusedPlan& p, const int64_t* indptr,
                           int64_t capacity, int32_t* indices, double* values, cudaStream_t s) {
  if (capacity <= 0) return;
  const int64_t n_rows = rows->n_rows_total;
  int G = 4;
  const int64_t avg = capacity / std::max<int64_t>(1, rows->n_point_rows);
  while (G < 32 && G < avg) G <<= 1;
  { compact_rows_kernel<<<grid_for(n_rows * G, 256, 8), 256, 0, s>>>(reinterpret_cast<const int32_t*>(ws + p.off_src), indptr,
                                                                  rows->coo_cols.p, rows->coo_vals.p, rows->nq * sp->nd, n_rows, G,
                                                                  capacity, indices, values); FXII_LAUNCHED(1); }
  FXII_CUDA(cudaGetLastError());
}

// Everything of one fused build, enqueued on `s` without any host synchronisation (so it can be
// captured into a CUDA graph): clear workspace | frames | locate+reduce | scan | compact | results.
// ev[0..3] bracket the three stages.  capacity == 0 skips the compaction (nnz still unknown).
static void enqueue_fused(const fxii_space* sp, fxii_rows* rows, const FusedPlan& p, cudaStream_t s, uint8_t* ws, uint8_t* scan_tmp,
                          int64_t* indptr, int32_t* indices, double* values, int64_t capacity, FusedResult* host_result,
                          cudaEvent_t* ev) {
  const fxii_mesh* m = sp->mesh;
  const int64_t n_rows = rows->n_rows_total;
  // inside a stream capture the stage timers become external event-record nodes of the graph
  cudaStreamCaptureStatus cap_status = cudaStreamCaptureStatusNone;
  FXII_CUDA(cudaStreamIsCapturing(s, &cap_status));
  const unsigned ev_flags = cap_status == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault;
  FXII_CUDA(cudaMemsetAsync(ws, 0, p.ws_bytes, s));
  FXII_CUDA(cudaEventRecordWithFlags(ev[0], s, ev_flags));
  launch_frames(rows, s);
  FXII_CUDA(cudaEventRecordWithFlags(ev[1], s, ev_flags));
  int64_t* row_cnt = reinterpret_cast<int64_t*>(ws);
  PiCounters* counters_dev = reinterpret_cast<PiCounters*>(ws + p.off_counters);
  int* conflict = reinterpret_cast<int*>(ws + p.off_conflict);
  int32_t* src_row = reinterpret_cast<int32_t*>(ws + p.off_src);
  auto launch = [&](auto kernel) {
    const int4* cn = reinterpret_cast<const int4*>(m->cell_nodes.p);
    kernel<<<p.grid, p.B, p.smem, s>>>(m->wnodes.p, m->grid, m->cellgeo.p, m->x.p, cn, m->padding, sp->dofmap.p, rows->frames.p,
                                       rows->table.p, rows->w.p, rows->pts_in.p, rows->w_in.p, rows->s_in.p, rows->kind, rows->nq,
                                       (rows->kind != FXII_OP_POINTS) ? packet_mode() : 0, p.team, p.scap, rows->n_point_rows,
                                       n_rows, rows->row_dofs.p, rows->cell.p, rows->ncoll.p, rows->points.p, rows->coo_cols.p,
                                       rows->coo_vals.p, row_cnt, src_row, conflict, counters_dev, CoopArgs{});
    FXII_LAUNCHED(1);
    FXII_CUDA(cudaGetLastError());
  };
  if (sp->nd == 4) launch(pi_fused_kernel<4, 0>);
  else if (sp->nd == 8) launch(pi_fused_kernel<8, 0>);
  else launch(pi_fused_kernel<10, 0>);
  FXII_CUDA(cudaEventRecordWithFlags(ev[2], s, ev_flags));
  size_t tb = p.scan_tmp_bytes;
  FXII_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp, tb, (const long long*)row_cnt, (long long*)indptr, (int64_t)(n_rows + 1), s));
  launch_compact(sp, rows, ws, p, indptr, capacity, indices, values, s);
  // the result record goes into the pinned (device-mapped) slot by a store from a kernel, not by the copy engine: a
  // device-to-host copy of a few bytes queues behind any CSR download still in flight on another stream
  { publish_result_kernel<<<1, 32, 0, s>>>(host_result, indptr + n_rows, conflict, counters_dev); FXII_LAUNCHED(1); }
  FXII_CUDA(cudaGetLastError());
  FXII_CUDA(cudaEventRecordWithFlags(ev[3], s, ev_flags));
}

static void set_fused_attributes(const fxii_space* sp, const FusedPlan& p) {
  // the attributes stick to the kernels: only set them again when the launch shape changes
  static int last_nd = -1, last_B = -1, last_dev = -1;
  static size_t last_smem = 0;
  int dev = 0;
  FXII_CUDA(cudaGetDevice(&dev));
  if (last_nd == sp->nd && last_B == p.B && last_smem == p.smem && last_dev == dev) return;
  last_nd = sp->nd, last_B = p.B, last_smem = p.smem, last_dev = dev;
  auto set = [&](auto kernel) {
    if (p.smem > 48 * 1024) FXII_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    // shared memory for all blocks the register file can hold (1024 threads at 64 registers), the rest stays L1
    int carve = (int)((p.smem + 1024) * (1024 / p.B) * 100 / (228 * 1024)) + 1;
    if (carve > 100) carve = 100;
    FXII_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
  };
  if (sp->nd == 4) {
    set(pi_fused_kernel<4, 0>);
    set(pi_fused_kernel<4, 2>);
  } else if (sp->nd == 8) {
    set(pi_fused_kernel<8, 0>);
    set(pi_fused_kernel<8, 2>);
  } else {
    set(pi_fused_kernel<10, 0>);
    set(pi_fused_kernel<10, 2>);
  }
}

// The single cooperative kernel wins while a build is launch-bound (config 2: 0.096 ms against 0.13-0.18 ms
// for memset + 5 kernels, graph-replayed) and loses once the row phase runs for milliseconds: computing
// the frames per team costs ~7 % there (13.7 against 12.2 + 0.6 ms on the 1M-segment tree) and the
// per-CTA scan + copy runs at half the rate of the dedicated kernels.  Measured break-even: ~8 M COO
// entries; the default switch-over is 4 M.  FXII_COOP=0 disables it, FXII_COOP_MAX_ENTRIES moves it.
static bool coop_enabled(int64_t E) {
  static const int64_t max_entries = [] {
    const char* e = getenv("FXII_COOP");
    if (e && e[0] == '0') return (int64_t)-1;
    const char* m = getenv("FXII_COOP_MAX_ENTRIES");
    return m ? (int64_t)atoll(m) : ((int64_t)1 << 22);
  }();
  return E <= max_entries;
}

// CTAs of the fused kernel that fit one SM (cached per kernel variant and launch shape)
template <typename K>
static int fused_blocks_per_sm(K kernel, int B, size_t smem) {
  struct Key { const void* k; int B; size_t smem; int dev; int n; };
  static std::vector<Key> cache;
  int dev = 0;
  FXII_CUDA(cudaGetDevice(&dev));
  for (const auto& c : cache)
    if (c.k == (const void*)kernel && c.B == B && c.smem == smem && c.dev == dev) return c.n;
  int n = 0;
  FXII_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, B, smem));
  cache.push_back({(const void*)kernel, B, smem, dev, n});
  return n;
}

// Single-kernel build: workspace memset + ONE cooperative launch (frames, locate, row reduction, scan,
// compaction), one 56-byte result copy, one host synchronisation.  Without an nnz hint (first build on
// a handle) the kernel only counts and the compaction runs as a second launch into exact-size arrays.
// Returns false when K is not cell-local (the caller falls back to the COO path).
// FXII_HOSTPROF=1: host wall clock of the phases of a cooperative build, averaged and printed every 100 builds
struct HostProf {
  bool on;
  double t[6] = {0, 0, 0, 0, 0, 0};
  int n = 0;
  double last = 0;
  HostProf() {
    const char* e = getenv("FXII_HOSTPROF");
    on = e && e[0] == '1';
  }
  void start() {
    if (on) last = Trace::now();
  }
  void lap(int i) {
    if (!on) return;
    const double now = Trace::now();
    t[i] += now - last;
    last = now;
  }
  void done() {
    if (!on || ++n < 100) return;
    fprintf(stderr, "[fxii hostprof] per build, us: prep %.1f | memset+event %.1f | coop launch %.1f | event+copies %.1f | sync %.1f | post %.1f\n",
            10 * t[0], 10 * t[1], 10 * t[2], 10 * t[3], 10 * t[4], 10 * t[5]);
    for (auto& v : t) v = 0;
    n = 0;
  }
};
static HostProf g_hostprof;

static bool pi_assemble_coop(const fxii_space* sp, fxii_rows* rows, const FusedPlan& p, cudaStream_t s, fxii_csr* out,
                             PiCounters* h_counters, float* ms_frames, float* ms_locate, float* ms_csr, bool* launched) {
  *launched = false;
  g_hostprof.start();
  const fxii_mesh* m = sp->mesh;
  const int64_t n_rows = rows->n_rows_total;
  FusedResult* hr = static_cast<FusedResult*>(rows->host_result);
  cudaEvent_t* ev = rows->ev;
  const int64_t n_groups = (rows->n_point_rows + p.rpb - 1) / p.rpb;
  const int per_sm = sp->nd == 4   ? fused_blocks_per_sm(pi_fused_kernel<4, 2>, p.B, p.smem)
                     : sp->nd == 8 ? fused_blocks_per_sm(pi_fused_kernel<8, 2>, p.B, p.smem)
                                   : fused_blocks_per_sm(pi_fused_kernel<10, 2>, p.B, p.smem);
  if (per_sm <= 0) return false;
  const int grid = (int)std::min<int64_t>(std::min<int64_t>(n_groups, (int64_t)per_sm * num_sms()), p.max_grid);
  const int64_t hint = rows->nnz_hint;
  out->n_rows = n_rows;
  out->n_cols = sp->n_dofs;
  out->indptr.alloc(n_rows + 1, s);
  if (hint > 0) {
    out->indices.alloc(hint, s);
    out->values.alloc(hint, s);
  }
  uint8_t* ws = rows->ws.p;
  CoopArgs ca;
  ca.gx = rows->gx.p;
  ca.gcells = rows->gcells.p;
  ca.cells_k = rows->cells_k.p;
  ca.xref = rows->xref.p;
  ca.radius = rows->radius.p;
  ca.nip = rows->nip;
  ca.radius_per_row = rows->radius_per_row;
  ca.barrier = reinterpret_cast<unsigned int*>(ws + p.off_barrier);
  ca.ticket = reinterpret_cast<unsigned long long*>(ws + p.off_barrier + 8);
  {
    // units (warps of sub-warp teams, or multi-warp teams) and how many of them the grid runs at a time
    const int64_t units = p.team >= 32 ? rows->n_point_rows : (rows->n_point_rows + 32 / p.team - 1) / (32 / p.team);
    const int64_t slots = (int64_t)grid * (p.team >= 32 ? p.rpb : p.B / 32);
    // only one-point rows are cheap enough for the counter to matter (1.2 M trace rows: 0.48 -> 0.35 ms); for
    // averaging operators coarser tickets just lengthen the tail (measured: 0.84 -> 1.07 ms with batches of 16)
    ca.ticket_batch = p.team == 1 ? (int)std::max<int64_t>(1, std::min<int64_t>(8, units / (2 * std::max<int64_t>(slots, 1)))) : 1;
  }
  ca.cta_sum = reinterpret_cast<int64_t*>(ws + p.off_cta);
  ca.indptr = out->indptr.p;
  ca.indices = hint > 0 ? out->indices.p : nullptr;
  ca.values = hint > 0 ? out->values.p : nullptr;
  ca.capacity = hint > 0 ? hint : 0;
  // the result record is written straight into the handle's pinned host slot (unified addressing): no copy node
  memset(hr, 0, sizeof(FusedResult));
  ca.result = hr;
  // kernel arguments (by address, cudaLaunchCooperativeKernel)
  const WideNode* a_nodes = m->wnodes.p;
  SceneGrid a_grid = m->grid;
  const double* a_cellgeo = m->cellgeo.p;
  const double* a_x = m->x.p;
  const int4* a_cn = reinterpret_cast<const int4*>(m->cell_nodes.p);
  double a_pad = m->padding;
  const int32_t* a_dofmap = sp->dofmap.p;
  const double* a_frames = nullptr;
  const double* a_table = rows->table.p;
  const double* a_wq = rows->w.p;
  const double* a_pts = rows->pts_in.p;
  const double* a_win = rows->w_in.p;
  const double* a_sin = rows->s_in.p;
  int a_kind = rows->kind, a_nq = rows->nq, a_packet = (rows->kind != FXII_OP_POINTS) ? packet_mode() : 0, a_team = p.team;
  uint32_t a_scap = p.scap;
  int64_t a_npr = rows->n_point_rows, a_nrows = n_rows;
  const int32_t* a_rowdofs = rows->row_dofs.p;
  int32_t* a_cell = rows->cell.p;
  uint8_t* a_ncoll = rows->ncoll.p;
  double* a_points = rows->points.p;
  int32_t* a_scol = rows->coo_cols.p;
  double* a_sval = rows->coo_vals.p;
  int64_t* a_rowcnt = reinterpret_cast<int64_t*>(ws);
  int32_t* a_src = reinterpret_cast<int32_t*>(ws + p.off_src);
  int* a_conf = reinterpret_cast<int*>(ws + p.off_conflict);
  PiCounters* a_counters = reinterpret_cast<PiCounters*>(ws + p.off_counters);
  void* args[] = {&a_nodes, &a_grid, &a_cellgeo, &a_x, &a_cn, &a_pad, &a_dofmap, &a_frames, &a_table, &a_wq, &a_pts, &a_win,
                  &a_sin, &a_kind, &a_nq, &a_packet, &a_team, &a_scap, &a_npr, &a_nrows, &a_rowdofs, &a_cell, &a_ncoll,
                  &a_points, &a_scol, &a_sval, &a_rowcnt, &a_src, &a_conf, &a_counters, &ca};
  g_hostprof.lap(0);
  FXII_CUDA(cudaMemsetAsync(ws, 0, p.ws_bytes, s));
  FXII_CUDA(cudaEventRecord(ev[1], s));
  g_hostprof.lap(1);
  const void* fn = sp->nd == 4   ? (const void*)pi_fused_kernel<4, 2>
                   : sp->nd == 8 ? (const void*)pi_fused_kernel<8, 2>
                                 : (const void*)pi_fused_kernel<10, 2>;
  const cudaError_t le = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(p.B), args, p.smem, s);
  if (le != cudaSuccess) {  // e.g. a shared device that cannot hold the wave: use the multi-kernel path
    cudaGetLastError();
    return false;
  }
  *launched = true;
  FXII_LAUNCHED(1);
  g_hostprof.lap(2);
  FXII_CUDA(cudaEventRecord(ev[2], s));
  // host destination: the download is speculated on the hint and rides behind the kernel, before the one sync
  int64_t speculated = -1;
  if (rows->sink && hint > 0 && rows->sink->capacity >= hint) {
    HostSink* k = rows->sink;
    FXII_CUDA(cudaMemcpyAsync(k->indptr, out->indptr.p, sizeof(int64_t) * (size_t)(n_rows + 1), cudaMemcpyDeviceToHost, s));
    FXII_CUDA(cudaMemcpyAsync(k->indices, out->indices.p, sizeof(int32_t) * (size_t)hint, cudaMemcpyDeviceToHost, s));
    FXII_CUDA(cudaMemcpyAsync(k->values, out->values.p, sizeof(double) * (size_t)hint, cudaMemcpyDeviceToHost, s));
    speculated = hint;
  }
  g_hostprof.lap(3);
  FXII_CUDA(cudaStreamSynchronize(s));
  g_hostprof.lap(4);
  *h_counters = hr->counters;
  if (hr->conflict) return false;
  out->nnz = hr->nnz;
  if (rows->sink && speculated >= hr->nnz) rows->sink->done = true;  // the kernel compacted (nnz <= hint) and all of it was copied
  float ms_extra = 0.f;
  if (hint <= 0 || hr->nnz > hint) {
    out->indices.alloc(hr->nnz, s);
    out->values.alloc(hr->nnz, s);
    FXII_CUDA(cudaEventRecord(ev[0], s));
    launch_compact(sp, rows, ws, p, out->indptr.p, hr->nnz, out->indices.p, out->values.p, s);
    FXII_CUDA(cudaEventRecord(ev[3], s));
    FXII_CUDA(cudaStreamSynchronize(s));
    ms_extra = elapsed_ms(ev[0], ev[3]);
  }
  rows->nnz_hint = hr->nnz;
  rows->n_builds += 1;
  // stage split of the single launch: event duration, apportioned by the in-kernel timers
  const float ms_kernel = elapsed_ms(ev[1], ev[2]);
  const double ns_all = (double)hr->ns_rows + (double)hr->ns_csr;
  const float frac_csr = ns_all > 0 ? (float)((double)hr->ns_csr / ns_all) : 0.f;
  *ms_frames = 0.f;
  *ms_csr = ms_kernel * frac_csr + ms_extra;
  *ms_locate = ms_kernel - ms_kernel * frac_csr;
  g_hostprof.lap(5);
  g_hostprof.done();
  return true;
}

static void destroy_graph(fxii_rows* rows) {
  if (rows->graph_exec) cudaGraphExecDestroy(rows->graph_exec);
  rows->graph_exec = nullptr;
}

static bool graphs_enabled() {
  static const bool on = [] {
    const char* e = getenv("FXII_GRAPH");
    return !(e && e[0] == '0');
  }();
  return on;
}

// Fused build.  Returns false when the fused path cannot be used (caller falls back to the COO path).
// First build on a handle: the nnz is read back before the CSR arrays are allocated (two host
// synchronisations).  Later builds know the nnz of the previous one (`nnz_hint`): a single
// synchronisation, and for small and medium workloads the whole sequence is replayed as ONE CUDA
// graph launch into buffers owned by the handle (launch-bound regime: config 2 runs ~9 launches of
// 5-50 us each), followed by device-to-device copies into the caller's exact-size CSR.
static bool pi_assemble_fused(const fxii_space* sp, fxii_rows* rows, cudaStream_t s, fxii_csr* out, PiCounters* h_counters,
                              float* ms_frames, float* ms_locate, float* ms_csr) {
  FusedPlan p;
  if (!make_fused_plan(sp, rows, p)) return false;
  const int64_t n_rows = rows->n_rows_total;
  const int64_t E = rows->n_point_rows * rows->nq * sp->nd;
  if (!rows->ev[0]) {
    for (auto& e : rows->ev) e = event_acquire();
  }
  static_assert(sizeof(FusedResult) <= 64, "FusedResult must fit a pinned slot");
  if (!rows->host_result) rows->host_result = pinned_slot_acquire();
  FusedResult* hr = static_cast<FusedResult*>(rows->host_result);
  cudaEvent_t* ev = rows->ev;
  set_fused_attributes(sp, p);
  bool moved = rows->ws.ensure((int64_t)p.ws_bytes, s);
  moved |= rows->scan_tmp.ensure((int64_t)p.scan_tmp_bytes + 16, s);
  out->n_rows = n_rows;
  out->n_cols = sp->n_dofs;
  const int64_t hint = rows->nnz_hint;
  rows->last_coop = false;
  // one-point rows (PointwiseTrace, mapped points) stay launch/latency bound four times longer
  if (coop_enabled(p.team == 1 ? E / 4 : E) && rows->mode != 2) {
    bool launched = false;
    rows->last_graph = false;
    const bool ok = pi_assemble_coop(sp, rows, p, s, out, h_counters, ms_frames, ms_locate, ms_csr, &launched);
    if (ok) {
      rows->last_coop = true;
      return true;
    }
    if (launched) return false;  // not cell-local: the COO path takes over
  }
  auto times = [&] {
    *ms_frames = elapsed_ms(ev[0], ev[1]);
    *ms_locate = elapsed_ms(ev[1], ev[2]);
    *ms_csr = elapsed_ms(ev[2], ev[3]);
  };
  // ---- graph replay -------------------------------------------------------------------------------
  // (a graph only pays off when the handle is reused: capture + instantiation cost ~0.2 ms)
  if (hint > 0 && rows->n_builds >= 1 && graphs_enabled() && E <= ((int64_t)1 << 26)) {
    moved |= rows->g_indptr.ensure(n_rows + 1, s);
    moved |= rows->g_indices.ensure(hint, s);
    moved |= rows->g_values.ensure(hint, s);
    if (moved || rows->buffers_moved || rows->graph_space != sp || rows->graph_stream != s || rows->graph_cap != hint) destroy_graph(rows);
    rows->buffers_moved = false;
    if (!rows->graph_exec) {
      cudaGraph_t graph = nullptr;
      // capture on a private stream: the caller's stream may be the legacy default stream, which cannot capture
      if (!rows->capture_stream) FXII_CUDA(cudaStreamCreateWithFlags(&rows->capture_stream, cudaStreamNonBlocking));
      cudaStream_t cs = rows->capture_stream;
      FXII_CUDA(cudaStreamSynchronize(s));
      FXII_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
      const int64_t launches_before = g_launches;
      try {
        enqueue_fused(sp, rows, p, cs, rows->ws.p, rows->scan_tmp.p, rows->g_indptr.p, rows->g_indices.p, rows->g_values.p, hint, hr, ev);
      } catch (...) {
        cudaStreamEndCapture(cs, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
      }
      FXII_CUDA(cudaStreamEndCapture(cs, &graph));
      rows->graph_kernels = (int)(g_launches - launches_before);  // kernels of this library inside the graph
      g_launches = launches_before;                                // captured, not launched
      const cudaError_t e = cudaGraphInstantiate(&rows->graph_exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) {
        cudaGetLastError();
        rows->graph_exec = nullptr;
      }
      rows->graph_space = sp;
      rows->graph_stream = s;
      rows->graph_cap = hint;
    }
    if (rows->graph_exec) {
      FXII_CUDA(cudaGraphLaunch(rows->graph_exec, s));
      FXII_LAUNCHED(rows->graph_kernels);
      FXII_CUDA(cudaStreamSynchronize(s));
      if (!hr->conflict && hr->nnz <= hint) {
        out->nnz = hr->nnz;
        out->indptr.alloc(n_rows + 1, s);
        out->indices.alloc(hr->nnz, s);
        out->values.alloc(hr->nnz, s);
        FXII_CUDA(cudaMemcpyAsync(out->indptr.p, rows->g_indptr.p, sizeof(int64_t) * (size_t)(n_rows + 1), cudaMemcpyDeviceToDevice, s));
        if (hr->nnz) {
          FXII_CUDA(cudaMemcpyAsync(out->indices.p, rows->g_indices.p, sizeof(int32_t) * (size_t)hr->nnz, cudaMemcpyDeviceToDevice, s));
          FXII_CUDA(cudaMemcpyAsync(out->values.p, rows->g_values.p, sizeof(double) * (size_t)hr->nnz, cudaMemcpyDeviceToDevice, s));
        }
        *h_counters = hr->counters;
        rows->nnz_hint = hr->nnz;
        rows->n_builds += 1;
        rows->last_graph = true;
        times();
        return true;
      }
      destroy_graph(rows);  // the matrix outgrew the hint (or the inputs changed): rebuild directly
    }
  }
  // ---- direct enqueue -----------------------------------------------------------------------------
  rows->last_graph = false;
  out->indptr.alloc(n_rows + 1, s);
  if (hint > 0) {  // speculative: the CSR arrays exist before the kernels run, no host round trip
    out->indices.alloc(hint, s);
    out->values.alloc(hint, s);
  }
  enqueue_fused(sp, rows, p, s, rows->ws.p, rows->scan_tmp.p, out->indptr.p, out->indices.p, out->values.p, hint, hr, ev);
  FXII_CUDA(cudaStreamSynchronize(s));
  *h_counters = hr->counters;
  if (hr->conflict) return false;  // K dofs shared between point rows: the general path handles first-visit-wins
  out->nnz = hr->nnz;
  if (hint <= 0 || hr->nnz > hint) {
    out->indices.alloc(hr->nnz, s);
    out->values.alloc(hr->nnz, s);
    launch_compact(sp, rows, rows->ws.p, p, out->indptr.p, hr->nnz, out->indices.p, out->values.p, s);
    FXII_CUDA(cudaEventRecord(ev[3], s));
    FXII_CUDA(cudaStreamSynchronize(s));
  }
  rows->nnz_hint = hr->nnz;
  rows->n_builds += 1;
  times();
  return true;
}

// -------------------------------------------------------------------------------------------------
void pi_assemble(const fxii_space* sp, fxii_rows* rows, cudaStream_t s, fxii_csr* out, fxii_pi_stats* stats,
                 bool keep_points) {
  const fxii_mesh* m = sp->mesh;
  const int nd = sp->nd;
  const int nq = rows->nq;
  const int64_t Nr = rows->n_point_rows;
  const int64_t Np = Nr * nq;
  const int64_t E = Np * nd;
  rows->nd = nd;
  // stage timers of the unfused path: created on first use, owned by the handle
  cudaEvent_t* ev = rows->ev_coo;
  // per-build buffers keep their addresses while their sizes do not change (a captured graph refers to them)
  bool moved = false;
  if (rows->kind != FXII_OP_POINTS) moved |= rows->frames.ensure(kFrame * Nr, s);
  moved |= rows->cell.ensure(Np, s);
  moved |= rows->ncoll.ensure(Np, s);
  if (keep_points) moved |= rows->points.ensure(3 * Np, s);
  else if (rows->points.p) {
    rows->points.release();
    moved = true;
  }
  moved |= rows->coo_cols.ensure(E, s);  // COO (unfused) or per-row staging (fused)
  moved |= rows->coo_vals.ensure(E, s);
  rows->buffers_moved |= moved;
  DevBuf<PiCounters> counters;
  PiCounters h;
  memset(&h, 0, sizeof(h));
  const size_t tab_smem = (rows->kind == FXII_OP_CIRCLE || rows->kind == FXII_OP_DISK) ? sizeof(double) * 4 * (size_t)nq : 0;
  FXII_CHECK(tab_smem <= 48 * 1024, "operator table too large for shared memory (nq > 1536)");
  float ms_count = 0.f, ms_fill = 0.f, ms_fr_f = 0.f, ms_loc_f = 0.f, ms_csr_f = 0.f;
  bool fused = false;
  // a handle whose rows were found not to be cell-local (continuous K) goes straight to the COO path from then on
  if (nd == 20) ensure_p3_table();
  // (P3 takes the general path: its 20-entry point records do not fit the fused kernel's per-team sort budget)
  if (rows->mode != 1 && !rows->not_cell_local && nd != 20) {
    fused = pi_assemble_fused(sp, rows, s, out, &h, &ms_fr_f, &ms_loc_f, &ms_csr_f);
    if (!fused && rows->host_result && static_cast<FusedResult*>(rows->host_result)->conflict == 1) rows->not_cell_local = true;
  }
  rows->last_fused = fused;
  if (!fused) {
    if (!ev[0]) {
      for (int i = 0; i < 4; ++i) ev[i] = event_acquire();
    }
    counters.alloc(1, s);
    FXII_CUDA(cudaMemsetAsync(counters.p, 0, sizeof(PiCounters), s));
    FXII_CUDA(cudaEventRecord(ev[0], s));
    launch_frames(rows, s);
    FXII_CUDA(cudaEventRecord(ev[1], s));
    if (Np > 0) {
      const int threads = 128;
      const int grid = grid_for(Np, threads, 8);
      const size_t smem = tab_smem;
      const int4* cn = reinterpret_cast<const int4*>(m->cell_nodes.p);
      if (nd == 4)
        { pi_locate_tabulate_kernel<4><<<grid, threads, smem, s>>>(
            m->wnodes.p, m->grid, m->cellgeo.p, m->x.p, cn, m->padding, sp->dofmap.p, rows->frames.p, rows->table.p, rows->w.p,
            rows->pts_in.p, rows->w_in.p, rows->s_in.p, rows->kind, nq, (rows->kind != FXII_OP_POINTS) ? packet_mode() : 0, Np, rows->cell.p, rows->ncoll.p, rows->points.p,
            rows->coo_cols.p, rows->coo_vals.p, counters.p); FXII_LAUNCHED(1); }
      else if (nd == 8)
        { pi_locate_tabulate_kernel<8><<<grid, threads, smem, s>>>(
            m->wnodes.p, m->grid, m->cellgeo.p, m->x.p, cn, m->padding, sp->dofmap.p, rows->frames.p, rows->table.p, rows->w.p,
            rows->pts_in.p, rows->w_in.p, rows->s_in.p, rows->kind, nq, (rows->kind != FXII_OP_POINTS) ? packet_mode() : 0, Np, rows->cell.p, rows->ncoll.p, rows->points.p,
            rows->coo_cols.p, rows->coo_vals.p, counters.p); FXII_LAUNCHED(1); }
      else if (nd == 20)
        { pi_locate_tabulate_kernel<20><<<grid, threads, smem, s>>>(
            m->wnodes.p, m->grid, m->cellgeo.p, m->x.p, cn, m->padding, sp->dofmap.p, rows->frames.p, rows->table.p, rows->w.p,
            rows->pts_in.p, rows->w_in.p, rows->s_in.p, rows->kind, nq, (rows->kind != FXII_OP_POINTS) ? packet_mode() : 0, Np, rows->cell.p, rows->ncoll.p, rows->points.p,
            rows->coo_cols.p, rows->coo_vals.p, counters.p); FXII_LAUNCHED(1); }
      else
        { pi_locate_tabulate_kernel<10><<<grid, threads, smem, s>>>(
            m->wnodes.p, m->grid, m->cellgeo.p, m->x.p, cn, m->padding, sp->dofmap.p, rows->frames.p, rows->table.p, rows->w.p,
            rows->pts_in.p, rows->w_in.p, rows->s_in.p, rows->kind, nq, (rows->kind != FXII_OP_POINTS) ? packet_mode() : 0, Np, rows->cell.p, rows->ncoll.p, rows->points.p,
            rows->coo_cols.p, rows->coo_vals.p, counters.p); FXII_LAUNCHED(1); }
      FXII_CUDA(cudaGetLastError());
    }
    FXII_CUDA(cudaEventRecord(ev[2], s));
    coo_segments_to_csr(rows->row_dofs.p, Nr, rows->n_rows_total, nq * nd, rows->coo_cols.p, rows->coo_vals.p, sp->n_dofs, s,
                        out, &ms_count, &ms_fill);
    FXII_CUDA(cudaEventRecord(ev[3], s));
    FXII_CUDA(cudaMemcpyAsync(&h, counters.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    FXII_CUDA(cudaStreamSynchronize(s));
  }
  if (stats) {
    stats->n_points = Np;
    stats->n_entries = E;
    stats->n_not_found = (int64_t)h.not_found;
    stats->n_ties = (int64_t)h.ties;
    stats->n_fallback = (int64_t)h.fallback;
    stats->ms_frames = fused ? ms_fr_f : elapsed_ms(ev[0], ev[1]);
    stats->ms_locate = fused ? ms_loc_f : elapsed_ms(ev[1], ev[2]);
    stats->ms_csr = fused ? ms_csr_f : elapsed_ms(ev[2], ev[3]);
    stats->ms_csr_count = ms_count;
    stats->ms_csr_fill = ms_fill;
    stats->fused = fused ? (rows->last_coop ? 3 : (rows->last_graph ? 2 : 1)) : 0;
  }
  if (h.stack_overflow > 0)
    throw Error(FXII_E_UNSUPPORTED, "the traversal stack (" + std::to_string(kStack) + " entries) overflowed: the LBVH of this mesh is too deep "
                                    "(degenerate cell distribution); ownership would be unreliable, nothing is returned");
  if (h.not_found > 0)
    throw Error(FXII_E_NOT_FOUND, std::to_string(h.not_found) + " of " + std::to_string(Np) +
                                      " points lie in no cell of the 3D mesh (interpolation.py:78)");
}

}  // namespace fxii
