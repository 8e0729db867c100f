// Host policy of dlio::OdomNode's per-scan loop over the device path (BASELINE configs 4 and 5), in
// C++ behind the C ABI: what ngicp/odom.py:OdomLoop restates in Python, with the same arithmetic so that both make the same
// decisions scan by scan. Reference src/dlio/src/dlio/odom.cc:
//   callbackPointCloud :737-837, preprocessPoints / deskewPointcloud :528-706 (device: ngicp_scan_ingest / ngicp_scan_deskew),
//   computeSpaciousness / computeDensity / setAdaptiveParams :1398-1436, :1600-1626, getNextPose :984-1018,
//   propagateGICP :1230-1246, updateKeyframes :1517-1598, pushSubmapIndices / buildSubmap / buildKeyframesAndSubmap :1628-1780,
//   computeConvexHull / computeConcaveHull :1438-1515 (PCL ConvexHull / ConcaveHull over qhull; planar keyframe sets — the
//   usual case — are handled here with a 2-D hull and a 2-D Delaunay alpha shape, spatial sets through a caller-supplied
//   callback, e.g. scipy's qhull, so that no decision is approximated).
// Out of scope, therefore inputs: IMU integration (the caller supplies the prior pose of every unique time stamp) and the
// geometric observer. Everything numeric runs in the kernels behind ngicp_*; this file only decides.
#include <algorithm>
#if defined(__SSE2__)
#include <emmintrin.h>
#include <xmmintrin.h>
#endif
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <new>
#include <queue>
#include <string>
#include <vector>

#include "../../include/ngicp_b200.h"
#include "internal.h"

namespace {

struct Keyframe {
  float p[3];
  double q[4];               // (w, x, y, z)
  ngicp_keyframe* kf = nullptr;
  float T_corr[16];          // row-major
};

// 4x4 row-major fp32 product, accumulation in index order without contraction (numpy float32 matmul of 4x4 operands)
void matmul4(const float* A, const float* B, float* C) {
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      float acc = 0.f;
      for (int k = 0; k < 4; k++) {
        volatile float prod = A[4 * r + k] * B[4 * k + c];
        acc = acc + prod;
      }
      C[4 * r + c] = acc;
    }
}
void to_colmajor(const float* rm, float* cm) {
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) cm[4 * c + r] = rm[4 * r + c];
}

// (w, x, y, z) of a rotation matrix, normalised (propagateGICP, odom.cc:1234-1245)
void quat_from_rot(const float* T, double q[4]) {
  double R[3][3];
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[r][c] = (double)T[4 * r + c];
  const double t = R[0][0] + R[1][1] + R[2][2];
  if (t > 0) {
    const double s = std::sqrt(t + 1.0) * 2;
    q[0] = 0.25 * s; q[1] = (R[2][1] - R[1][2]) / s; q[2] = (R[0][2] - R[2][0]) / s; q[3] = (R[1][0] - R[0][1]) / s;
  } else {
    int i = 0;
    if (R[1][1] > R[i][i]) i = 1;
    if (R[2][2] > R[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    const double s = std::sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0) * 2;
    q[0] = q[1] = q[2] = q[3] = 0.0;
    q[1 + i] = 0.25 * s;
    q[0] = (R[k][j] - R[j][k]) / s;
    q[1 + j] = (R[j][i] + R[i][j]) / s;
    q[1 + k] = (R[k][i] + R[i][k]) / s;
  }
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int a = 0; a < 4; a++) q[a] /= n;
}
// angle of q * r^-1 with the sign fix of updateKeyframes (odom.cc:1560-1572)
double quat_angle_deg(const double q[4], const double r_in[4]) {
  double r[4] = {r_in[0], r_in[1], r_in[2], r_in[3]};
  if (q[0] * r[0] + q[1] * r[1] + q[2] * r[2] + q[3] * r[3] < 0) for (double& v : r) v = -v;
  const double rr = r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3];
  const double ri[4] = {r[0] / rr, -r[1] / rr, -r[2] / rr, -r[3] / rr};
  const double w = q[0] * ri[0] - (q[1] * ri[1] + q[2] * ri[2] + q[3] * ri[3]);
  const double v[3] = {q[0] * ri[1] + ri[0] * q[1] + (q[2] * ri[3] - q[3] * ri[2]), q[0] * ri[2] + ri[0] * q[2] + (q[3] * ri[1] - q[1] * ri[3]),
                       q[0] * ri[3] + ri[0] * q[3] + (q[1] * ri[2] - q[2] * ri[1])};
  return 2.0 * std::atan2(std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), w) * (180.0 / 3.14159265358979323846);
}

// pushSubmapIndices (odom.cc:1628-1652): every frame whose distance is <= the k-th smallest (ties all kept)
void push_submap_indices(const std::vector<float>& dists, int k, const std::vector<int>& frames, std::vector<int>& out) {
  if (dists.empty()) return;
  std::priority_queue<float> heap;
  for (float d : dists) {
    if ((int)heap.size() >= k && heap.top() > d) { heap.pop(); heap.push(d); }
    else if ((int)heap.size() < k) heap.push(d);
  }
  const float kth = heap.top();
  for (size_t i = 0; i < dists.size(); i++) if (dists[i] <= kth) out.push_back(frames[i]);
}

// eigenvalues (ascending) and eigenvectors (columns) of a symmetric 3x3 by cyclic Jacobi
void eigh3(const double A[3][3], double w[3], double V[3][3]) {
  double a[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { a[i][j] = A[i][j]; V[i][j] = i == j ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 32; sweep++) {
    const double off = std::fabs(a[0][1]) + std::fabs(a[0][2]) + std::fabs(a[1][2]);
    if (off == 0.0 || off <= 1e-300) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int r = 0; r < 3; r++) { const double x = a[r][p], y = a[r][q]; a[r][p] = c * x - s * y; a[r][q] = s * x + c * y; }
        for (int r = 0; r < 3; r++) { const double x = a[p][r], y = a[q][r]; a[p][r] = c * x - s * y; a[q][r] = s * x + c * y; }
        for (int r = 0; r < 3; r++) { const double x = V[r][p], y = V[r][q]; V[r][p] = c * x - s * y; V[r][q] = s * x + c * y; }
      }
  }
  int idx[3] = {0, 1, 2};
  std::sort(idx, idx + 3, [&](int x, int y) { return a[x][x] < a[y][y]; });
  double Vs[3][3];
  for (int c = 0; c < 3; c++) { w[c] = a[idx[c]][idx[c]]; for (int r = 0; r < 3; r++) Vs[r][c] = V[r][idx[c]]; }
  std::memcpy(V, Vs, sizeof Vs);
}

// PCL's calculateInputDimension: 2 when the smallest covariance eigenvalue is < 1e-3 of the largest, else 3; for the planar
// case the coordinate dropped is the one the plane normal is most aligned with
int hull_dimension(const std::vector<double>& P, int n, int& drop) {
  double mean[3] = {0, 0, 0};
  for (int i = 0; i < n; i++) for (int a = 0; a < 3; a++) mean[a] += P[3 * i + a];
  for (double& m : mean) m /= n;
  double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < n; i++)
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) C[a][b] += (P[3 * i + a] - mean[a]) * (P[3 * i + b] - mean[b]);
  const double den = std::max(n - 1, 1);
  for (auto& row : C) for (double& v : row) v /= den;
  double w[3], V[3][3];
  eigh3(C, w, V);
  drop = -1;
  if (w[2] <= 0 || w[0] / w[2] < 1e-3) {
    drop = 0;
    for (int a = 1; a < 3; a++) if (std::fabs(V[a][0]) > std::fabs(V[drop][0])) drop = a;
    return 2;
  }
  return 3;
}

struct P2 { double x, y; };
double cross(const P2& o, const P2& a, const P2& b) { return (a.x - o.x) * (b.y - o.y) - (a.y - o.y) * (b.x - o.x); }

// vertices of the 2-D convex hull (qhull's extreme points: no interior, no mid-edge points); collinear sets: the two extremes
void convex_hull_2d(const std::vector<P2>& Q, std::vector<int>& out) {
  const int n = (int)Q.size();
  std::vector<int> idx(n);
  for (int i = 0; i < n; i++) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return Q[a].x < Q[b].x || (Q[a].x == Q[b].x && Q[a].y < Q[b].y); });
  std::vector<int> h(2 * n);
  int k = 0;
  for (int i = 0; i < n; i++) {
    while (k >= 2 && cross(Q[h[k - 2]], Q[h[k - 1]], Q[idx[i]]) <= 0) k--;
    h[k++] = idx[i];
  }
  for (int i = n - 2, t = k + 1; i >= 0; i--) {
    while (k >= t && cross(Q[h[k - 2]], Q[h[k - 1]], Q[idx[i]]) <= 0) k--;
    h[k++] = idx[i];
  }
  h.resize(std::max(k - 1, 0));
  if (h.size() < 3) { h.clear(); if (n > 0) { h.push_back(idx.front()); if (n > 1) h.push_back(idx.back()); } }
  std::sort(h.begin(), h.end());
  h.erase(std::unique(h.begin(), h.end()), h.end());
  out = h;
}

// 2-D Delaunay triangulation (Bowyer-Watson over a super-triangle); triangles as index triples
struct Tri { int v[3]; };
bool in_circumcircle(const P2& a, const P2& b, const P2& c, const P2& p) {
  const double ax = a.x - p.x, ay = a.y - p.y, bx = b.x - p.x, by = b.y - p.y, cx = c.x - p.x, cy = c.y - p.y;
  const double det = (ax * ax + ay * ay) * (bx * cy - cx * by) - (bx * bx + by * by) * (ax * cy - cx * ay) + (cx * cx + cy * cy) * (ax * by - bx * ay);
  const double orient = (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x);
  return orient > 0 ? det > 0 : det < 0;
}
void delaunay_2d(const std::vector<P2>& Q, std::vector<Tri>& tris) {
  const int n = (int)Q.size();
  double lo[2] = {DBL_MAX, DBL_MAX}, hi[2] = {-DBL_MAX, -DBL_MAX};
  for (const P2& p : Q) { lo[0] = std::min(lo[0], p.x); lo[1] = std::min(lo[1], p.y); hi[0] = std::max(hi[0], p.x); hi[1] = std::max(hi[1], p.y); }
  const double d = std::max(hi[0] - lo[0], hi[1] - lo[1]) + 1.0, mx = 0.5 * (lo[0] + hi[0]), my = 0.5 * (lo[1] + hi[1]);
  std::vector<P2> pts = Q;
  pts.push_back({mx - 1e4 * d, my - 1e4 * d});
  pts.push_back({mx + 1e4 * d, my - 1e4 * d});
  pts.push_back({mx, my + 1e4 * d});
  tris.clear();
  tris.push_back({{n, n + 1, n + 2}});
  for (int i = 0; i < n; i++) {
    std::vector<std::pair<int, int>> edges;
    std::vector<Tri> keep;
    for (const Tri& t : tris) {
      if (in_circumcircle(pts[t.v[0]], pts[t.v[1]], pts[t.v[2]], pts[i])) {
        for (int e = 0; e < 3; e++) edges.push_back({t.v[e], t.v[(e + 1) % 3]});
      } else keep.push_back(t);
    }
    // boundary of the cavity = edges that appear once
    for (size_t a = 0; a < edges.size(); a++) {
      bool shared = false;
      for (size_t b = 0; b < edges.size(); b++)
        if (a != b && ((edges[a].first == edges[b].first && edges[a].second == edges[b].second) ||
                       (edges[a].first == edges[b].second && edges[a].second == edges[b].first))) { shared = true; break; }
      if (!shared) keep.push_back({{edges[a].first, edges[a].second, i}});
    }
    tris.swap(keep);
  }
  std::vector<Tri> real;
  for (const Tri& t : tris) if (t.v[0] < n && t.v[1] < n && t.v[2] < n) real.push_back(t);
  tris.swap(real);
}

// vertices of the 2-D alpha shape: Delaunay triangles with circumradius <= alpha, boundary = edges of exactly one kept
// triangle (pcl::ConcaveHull with setAlpha, odom.cc:1496-1512)
// A Delaunay simplex with its circumradius: the alpha shape for ANY alpha is read off a list of these (the triangulation
// depends on the points only, so the loop keeps the list between scans and re-filters it when only alpha moved).
struct Simplex { int v[4]; int nv; double radius; };

// boundary vertices of the alpha shape: simplices with circumradius <= alpha are kept, the shape's boundary is made of the
// faces that belong to exactly one kept simplex (pcl::ConcaveHull with setAlpha, odom.cc:1496-1512)
void alpha_boundary(const std::vector<Simplex>& S, double alpha, std::vector<int>& out) {
  out.clear();
  struct Face { int v[3]; bool operator<(const Face& o) const { return v[0] != o.v[0] ? v[0] < o.v[0] : (v[1] != o.v[1] ? v[1] < o.v[1] : v[2] < o.v[2]); }
                bool operator==(const Face& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; } };
  std::vector<Face> faces;
  int fv = 0;
  for (const Simplex& t : S) {
    if (t.radius > alpha) continue;
    fv = t.nv - 1;
    for (int skip = 0; skip < t.nv; skip++) {
      Face f = {{-1, -1, -1}};
      int k = 0;
      for (int q = 0; q < t.nv; q++) if (q != skip) f.v[k++] = t.v[q];
      std::sort(f.v, f.v + fv);
      faces.push_back(f);
    }
  }
  std::sort(faces.begin(), faces.end());
  for (size_t i = 0; i < faces.size();) {
    size_t j = i;
    while (j < faces.size() && faces[j] == faces[i]) j++;
    if (j - i == 1) for (int k = 0; k < fv; k++) out.push_back(faces[i].v[k]);
    i = j;
  }
  std::sort(out.begin(), out.end());
  out.erase(std::unique(out.begin(), out.end()), out.end());
}

void alpha_simplices_2d(const std::vector<P2>& Q, std::vector<Simplex>& S) {
  S.clear();
  if (Q.size() < 3) return;
  std::vector<Tri> tris;
  delaunay_2d(Q, tris);
  for (const Tri& t : tris) {
    const P2 &v0 = Q[t.v[0]], &v1 = Q[t.v[1]], &v2 = Q[t.v[2]];
    // circumcentre: 2 (V_i - V_0) . c = |V_i|^2 - |V_0|^2
    const double a11 = 2 * (v1.x - v0.x), a12 = 2 * (v1.y - v0.y), a21 = 2 * (v2.x - v0.x), a22 = 2 * (v2.y - v0.y);
    const double r1 = v1.x * v1.x + v1.y * v1.y - (v0.x * v0.x + v0.y * v0.y), r2 = v2.x * v2.x + v2.y * v2.y - (v0.x * v0.x + v0.y * v0.y);
    const double det = a11 * a22 - a12 * a21;
    if (!(std::fabs(det) > 1e-300)) continue;
    const double cx = (r1 * a22 - a12 * r2) / det, cy = (a11 * r2 - r1 * a21) / det;
    S.push_back({{t.v[0], t.v[1], t.v[2], -1}, 3, std::sqrt((cx - v0.x) * (cx - v0.x) + (cy - v0.y) * (cy - v0.y))});
  }
}
void concave_hull_2d(const std::vector<P2>& Q, double alpha, std::vector<int>& out) {
  std::vector<Simplex> S;
  alpha_simplices_2d(Q, S);
  alpha_boundary(S, alpha, out);
}

// ---- spatial keyframe sets (a sensor that also moves vertically): 3-D convex hull and 3-D alpha shape ----
struct P3 { double x, y, z; };
inline P3 sub(const P3& a, const P3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline P3 cross3(const P3& a, const P3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double dot3(const P3& a, const P3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// signed volume x 6 of (a, b, c, d): > 0 when d is on the side of the plane (a, b, c) its normal (b-a) x (c-a) points to
inline double orient3(const P3& a, const P3& b, const P3& c, const P3& d) { return dot3(cross3(sub(b, a), sub(c, a)), sub(d, a)); }

// Extreme points of the 3-D convex hull by incremental insertion (faces keep outward normals; a point that sees no face
// beyond a relative tolerance is inside or on the hull and is not a vertex — what qhull reports). false: degenerate input.
bool convex_hull_3d(const std::vector<P3>& Q, std::vector<int>& out) {
  out.clear();
  const int n = (int)Q.size();
  if (n < 4) return false;
  double ext = 0;
  for (const P3& p : Q) ext = std::max(ext, std::max(std::fabs(p.x - Q[0].x), std::max(std::fabs(p.y - Q[0].y), std::fabs(p.z - Q[0].z))));
  const double eps = 1e-10 * ext * ext * ext + 1e-300;
  // initial tetrahedron: two far points, the point farthest from their line, the point farthest from that plane
  int i0 = 0, i1 = 0, i2 = -1, i3 = -1;
  for (int i = 1; i < n; i++) if (Q[i].x < Q[i0].x || (Q[i].x == Q[i0].x && (Q[i].y < Q[i0].y || (Q[i].y == Q[i0].y && Q[i].z < Q[i0].z)))) i0 = i;
  double best = -1;
  for (int i = 0; i < n; i++) { const P3 d = sub(Q[i], Q[i0]); const double v = dot3(d, d); if (v > best) { best = v; i1 = i; } }
  best = -1;
  for (int i = 0; i < n; i++) { const P3 c = cross3(sub(Q[i1], Q[i0]), sub(Q[i], Q[i0])); const double v = dot3(c, c); if (v > best) { best = v; i2 = i; } }
  best = -1;
  for (int i = 0; i < n; i++) { const double v = std::fabs(orient3(Q[i0], Q[i1], Q[i2], Q[i])); if (v > best) { best = v; i3 = i; } }
  if (i2 < 0 || i3 < 0 || best <= eps) return false;
  struct Face { int a, b, c; bool alive; };
  std::vector<Face> F;
  auto add_face = [&](int a, int b, int c, int inside) {      // oriented so that `inside` is behind it
    if (orient3(Q[a], Q[b], Q[c], Q[inside]) > 0) std::swap(b, c);
    F.push_back({a, b, c, true});
  };
  add_face(i0, i1, i2, i3); add_face(i0, i1, i3, i2); add_face(i0, i2, i3, i1); add_face(i1, i2, i3, i0);
  const P3 centre = {(Q[i0].x + Q[i1].x + Q[i2].x + Q[i3].x) / 4, (Q[i0].y + Q[i1].y + Q[i2].y + Q[i3].y) / 4, (Q[i0].z + Q[i1].z + Q[i2].z + Q[i3].z) / 4};
  for (int p = 0; p < n; p++) {
    if (p == i0 || p == i1 || p == i2 || p == i3) continue;
    std::vector<int> vis;
    for (int f = 0; f < (int)F.size(); f++)
      if (F[f].alive && orient3(Q[F[f].a], Q[F[f].b], Q[F[f].c], Q[p]) > eps) vis.push_back(f);
    if (vis.empty()) continue;
    // horizon = directed edges of visible faces whose reverse is not an edge of a visible face
    std::vector<std::pair<int, int>> edges;
    for (int f : vis) { edges.push_back({F[f].a, F[f].b}); edges.push_back({F[f].b, F[f].c}); edges.push_back({F[f].c, F[f].a}); F[f].alive = false; }
    for (size_t e = 0; e < edges.size(); e++) {
      bool shared = false;
      for (size_t g = 0; g < edges.size(); g++) if (edges[g].first == edges[e].second && edges[g].second == edges[e].first) { shared = true; break; }
      if (!shared) {
        int a = edges[e].first, b = edges[e].second, c = p;
        // keep the interior reference point behind the new face
        const double o = dot3(cross3(sub(Q[b], Q[a]), sub(Q[c], Q[a])), sub(centre, Q[a]));
        if (o > 0) std::swap(a, b);
        F.push_back({a, b, c, true});
      }
    }
  }
  for (const Face& f : F) if (f.alive) { out.push_back(f.a); out.push_back(f.b); out.push_back(f.c); }
  std::sort(out.begin(), out.end());
  out.erase(std::unique(out.begin(), out.end()), out.end());
  return true;
}

// 3-D Delaunay (Bowyer-Watson over a super-tetrahedron; keyframe counts are in the hundreds), tetrahedra as index quads
struct Tet { int v[4]; };
bool in_circumsphere(const P3& a, const P3& b, const P3& c, const P3& d, const P3& p) {
  const P3 A = sub(a, p), B = sub(b, p), C = sub(c, p), D = sub(d, p);
  const double a2 = dot3(A, A), b2 = dot3(B, B), c2 = dot3(C, C), d2 = dot3(D, D);
  // 4x4 determinant | A a2 ; B b2 ; C c2 ; D d2 |, expanded along the last column
  const double det = -a2 * dot3(B, cross3(C, D)) + b2 * dot3(A, cross3(C, D)) - c2 * dot3(A, cross3(B, D)) + d2 * dot3(A, cross3(B, C));
  const double o = orient3(a, b, c, d);
  return o > 0 ? det < 0 : det > 0;
}
bool delaunay_3d(const std::vector<P3>& Q, std::vector<Tet>& tets) {
  const int n = (int)Q.size();
  double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (const P3& p : Q) { lo[0] = std::min(lo[0], p.x); lo[1] = std::min(lo[1], p.y); lo[2] = std::min(lo[2], p.z); hi[0] = std::max(hi[0], p.x); hi[1] = std::max(hi[1], p.y); hi[2] = std::max(hi[2], p.z); }
  const double d = std::max(hi[0] - lo[0], std::max(hi[1] - lo[1], hi[2] - lo[2])) + 1.0;
  const P3 m = {0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2])};
  std::vector<P3> pts = Q;
  const double S = 1e3 * d;
  pts.push_back({m.x - S, m.y - S, m.z - S});
  pts.push_back({m.x + S, m.y - S, m.z - S});
  pts.push_back({m.x, m.y + S, m.z - S});
  pts.push_back({m.x, m.y, m.z + S});
  tets.clear();
  tets.push_back({{n, n + 1, n + 2, n + 3}});
  struct Tri3 { int a, b, c; };
  for (int i = 0; i < n; i++) {
    std::vector<Tri3> faces;
    std::vector<Tet> keep;
    for (const Tet& t : tets) {
      if (in_circumsphere(pts[t.v[0]], pts[t.v[1]], pts[t.v[2]], pts[t.v[3]], pts[i])) {
        for (int s = 0; s < 4; s++) {
          int f[3], k = 0;
          for (int q = 0; q < 4; q++) if (q != s) f[k++] = t.v[q];
          std::sort(f, f + 3);
          faces.push_back({f[0], f[1], f[2]});
        }
      } else keep.push_back(t);
    }
    if (faces.empty()) return false;                       // the point is in no circumsphere: numerical trouble
    for (size_t a = 0; a < faces.size(); a++) {
      int cnt = 0;
      for (size_t b = 0; b < faces.size(); b++) if (faces[a].a == faces[b].a && faces[a].b == faces[b].b && faces[a].c == faces[b].c) cnt++;
      if (cnt == 1) {
        if (std::fabs(orient3(pts[faces[a].a], pts[faces[a].b], pts[faces[a].c], pts[i])) <= 0.0) return false;   // flat tetrahedron
        keep.push_back({{faces[a].a, faces[a].b, faces[a].c, i}});
      }
    }
    tets.swap(keep);
  }
  std::vector<Tet> real;
  for (const Tet& t : tets) if (t.v[0] < n && t.v[1] < n && t.v[2] < n && t.v[3] < n) real.push_back(t);
  tets.swap(real);
  return true;
}

// the 3-D Delaunay tetrahedra with their circumradii (false: numerical trouble, see delaunay_3d)
bool alpha_simplices_3d(const std::vector<P3>& Q, std::vector<Simplex>& S) {
  S.clear();
  if (Q.size() < 4) return true;
  std::vector<Tet> tets;
  if (!delaunay_3d(Q, tets)) return false;
  for (const Tet& t : tets) {
    const P3 &v0 = Q[t.v[0]];
    double A[3][3], r[3];
    for (int i = 0; i < 3; i++) {
      const P3& vi = Q[t.v[i + 1]];
      A[i][0] = 2 * (vi.x - v0.x); A[i][1] = 2 * (vi.y - v0.y); A[i][2] = 2 * (vi.z - v0.z);
      r[i] = dot3(vi, vi) - dot3(v0, v0);
    }
    const double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) + A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
    if (!(std::fabs(det) > 1e-300)) continue;
    const double cx = (r[0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (r[1] * A[2][2] - A[1][2] * r[2]) + A[0][2] * (r[1] * A[2][1] - A[1][1] * r[2])) / det;
    const double cy = (A[0][0] * (r[1] * A[2][2] - A[1][2] * r[2]) - r[0] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) + A[0][2] * (A[1][0] * r[2] - r[1] * A[2][0])) / det;
    const double cz = (A[0][0] * (A[1][1] * r[2] - r[1] * A[2][1]) - A[0][1] * (A[1][0] * r[2] - r[1] * A[2][0]) + r[0] * (A[1][0] * A[2][1] - A[1][1] * A[2][0])) / det;
    const P3 dc = {cx - v0.x, cy - v0.y, cz - v0.z};
    S.push_back({{t.v[0], t.v[1], t.v[2], t.v[3]}, 4, std::sqrt(dot3(dc, dc))});
  }
  return true;
}
// vertices of the 3-D alpha shape (the same rule as concave_hull_2d one dimension up)
bool concave_hull_3d(const std::vector<P3>& Q, double alpha, std::vector<int>& out) {
  std::vector<Simplex> S;
  if (!alpha_simplices_3d(Q, S)) { out.clear(); return false; }
  alpha_boundary(S, alpha, out);
  return true;
}

}  // namespace

struct ngicp_odom {
  ngicp_handle* h = nullptr;
  ngicp_odom_params p;
  float T[16], T_prior[16], T_corr[16];     // row-major
  float lidar_p[3] = {0, 0, 0};
  double lidar_q[4] = {1, 0, 0, 0};
  std::vector<Keyframe> keyframes;
  int num_processed = 0;
  std::vector<int> convex, concave, submap_curr, submap_prev;
  bool submap_changed = true;
  double keyframe_thresh_dist, concave_alpha;
  bool have_median = false, have_density = false, first_opt_done = false;
  float median_prev = 0.f, density_prev = 0.f;
  double spaciousness = 0.0, density = 0.0;
  float source_density = 0.f;
  // scan in flight between begin and finish
  std::vector<float> ranges;
  size_t n_ranges = 0;
  float median_curr = 0.f;
  bool median_ready = false;
  size_t n_unique = 0, n_kept = 0;
  void* pack = nullptr;       // page-locked: the scan reduced to what the device needs, (x, y, z, stamp) per record
  size_t pack_cap = 0;
  ngicp_hull_fn convex_cb = nullptr, concave_cb = nullptr;
  void* cb_user = nullptr;
  // hulls of the first hull_n keyframes (their positions never change): the convex hull and the Delaunay simplices are
  // rebuilt only when a keyframe has been added; the alpha shape is re-read from the simplices every scan (alpha adapts)
  int hull_n = -1, hull_dim = 0;
  bool hull_ok = false;
  std::vector<int> hull_convex;
  std::vector<Simplex> hull_simplices;
  std::string err;
  double prof[NGICP_ODOM_STAGES] = {0};     // host wall clock per stage, seconds, summed over scans
  long prof_scans = 0;
};

namespace {

struct StageClock {
  ngicp_odom* o;
  std::chrono::steady_clock::time_point t;
  explicit StageClock(ngicp_odom* o_) : o(o_), t(std::chrono::steady_clock::now()) {}
  void lap(int stage) {
    const auto n = std::chrono::steady_clock::now();
    o->prof[stage] += std::chrono::duration<double>(n - t).count();
    t = n;
  }
};

void identity16(float* T) { std::memset(T, 0, 16 * sizeof(float)); T[0] = T[5] = T[10] = T[15] = 1.f; }

void propagate(ngicp_odom* o) {
  o->lidar_p[0] = o->T[3]; o->lidar_p[1] = o->T[7]; o->lidar_p[2] = o->T[11];
  quat_from_rot(o->T, o->lidar_q);
}

int hull_indices(ngicp_odom* o, bool concave, std::vector<int>& out) {
  const int n = o->num_processed;
  if (o->hull_n != n) {
    std::vector<double> P(3 * (size_t)n);
    for (int i = 0; i < n; i++) for (int a = 0; a < 3; a++) P[3 * i + a] = (double)o->keyframes[i].p[a];
    int drop;
    o->hull_dim = hull_dimension(P, n, drop);
    o->hull_n = n;
    o->hull_ok = true;
    o->hull_convex.clear();
    o->hull_simplices.clear();
    if (o->hull_dim == 3) {
      if (!o->convex_cb || !o->concave_cb) {      // no caller-supplied hulls (e.g. PCL's): the native 3-D convex hull / alpha shape
        std::vector<P3> Q3(n);
        for (int i = 0; i < n; i++) Q3[i] = {P[3 * i], P[3 * i + 1], P[3 * i + 2]};
        o->hull_ok = convex_hull_3d(Q3, o->hull_convex) && alpha_simplices_3d(Q3, o->hull_simplices);
      }
    } else {
      std::vector<P2> Q(n);
      const int ax = drop == 0 ? 1 : 0, ay = drop == 2 ? 1 : 2;
      for (int i = 0; i < n; i++) Q[i] = {P[3 * i + ax], P[3 * i + ay]};
      convex_hull_2d(Q, o->hull_convex);
      alpha_simplices_2d(Q, o->hull_simplices);
    }
  }
  if (o->hull_dim == 3 && o->convex_cb && o->concave_cb) {
    std::vector<double> P(3 * (size_t)n);
    for (int i = 0; i < n; i++) for (int a = 0; a < 3; a++) P[3 * i + a] = (double)o->keyframes[i].p[a];
    std::vector<int> buf(n);
    const int m = (concave ? o->concave_cb : o->convex_cb)(P.data(), n, concave ? o->concave_alpha : 0.0, buf.data(), o->cb_user);
    if (m < 0) { o->err = "odom loop: hull callback failed"; return NGICP_ERR_INVALID; }
    out.assign(buf.begin(), buf.begin() + m);
    std::sort(out.begin(), out.end());
    return NGICP_OK;
  }
  if (!o->hull_ok) { o->err = "odom loop: degenerate spatial keyframe set (install hull callbacks)"; return NGICP_ERR_UNSUPPORTED; }
  if (concave) alpha_boundary(o->hull_simplices, o->concave_alpha, out);
  else out = o->hull_convex;
  return NGICP_OK;
}

// buildSubmap (odom.cc:1654-1742)
int build_submap(ngicp_odom* o) {
  const int n = o->num_processed;
  std::vector<float> ds(n);
  for (int i = 0; i < n; i++) {
    float acc = 0.f;
    for (int a = 0; a < 3; a++) { const float d = o->lidar_p[a] - o->keyframes[i].p[a]; volatile float sq = d * d; acc = acc + sq; }
    ds[i] = std::sqrt(acc);
  }
  std::vector<int> all(n), cur;
  for (int i = 0; i < n; i++) all[i] = i;
  push_submap_indices(ds, o->p.submap_knn, all, cur);
  if (n >= 4) if (int rc = hull_indices(o, false, o->convex)) return rc;
  {
    std::vector<float> d2;
    for (int c : o->convex) d2.push_back(ds[c]);
    push_submap_indices(d2, o->p.submap_kcv, o->convex, cur);
  }
  if (n >= 5) if (int rc = hull_indices(o, true, o->concave)) return rc;
  {
    std::vector<float> d2;
    for (int c : o->concave) d2.push_back(ds[c]);
    push_submap_indices(d2, o->p.submap_kcc, o->concave, cur);
  }
  std::sort(cur.begin(), cur.end());
  cur.erase(std::unique(cur.begin(), cur.end()), cur.end());
  o->submap_curr = cur;
  if (o->submap_curr != o->submap_prev) {
    o->submap_changed = true;
    std::vector<ngicp_keyframe*> kfs;
    for (int k : cur) kfs.push_back(o->keyframes[k].kf);
    if (int rc = ngicp_submap_assemble(o->h, kfs.data(), (int)kfs.size())) return rc;   // odom.cc:1719-1738
    o->submap_prev = cur;
  }
  return NGICP_OK;
}

int build_keyframes_and_submap(ngicp_odom* o) {
  for (int i = o->num_processed; i < (int)o->keyframes.size(); i++) {
    float cm[16];
    to_colmajor(o->keyframes[i].T_corr, cm);
    if (int rc = ngicp_keyframe_transform(o->h, o->keyframes[i].kf, cm)) return rc;         // odom.cc:1757-1762
    o->num_processed++;
  }
  return build_submap(o);
}

int push_keyframe(ngicp_odom* o) {
  Keyframe k;
  std::memcpy(k.p, o->lidar_p, sizeof k.p);
  std::memcpy(k.q, o->lidar_q, sizeof k.q);
  std::memcpy(k.T_corr, o->T_corr, sizeof k.T_corr);
  if (int rc = ngicp_keyframe_capture(o->h, &k.kf)) return rc;
  o->keyframes.push_back(k);
  return NGICP_OK;
}

// updateKeyframes (odom.cc:1517-1598)
int update_keyframes(ngicp_odom* o, bool* is_new) {
  const int n = (int)o->keyframes.size();
  int closest = 0, num_nearby = 0;
  float best = FLT_MAX;
  const float near_thr = (float)(o->keyframe_thresh_dist * 1.5);
  for (int i = 0; i < n; i++) {
    float acc = 0.f;
    for (int a = 0; a < 3; a++) { const float d = o->lidar_p[a] - o->keyframes[i].p[a]; volatile float sq = d * d; acc = acc + sq; }
    const float d = std::sqrt(acc);
    if (d <= near_thr) num_nearby++;
    if (d < best) { best = d; closest = i; }
  }
  const double dd = (double)best;
  const double theta = quat_angle_deg(o->lidar_q, o->keyframes[closest].q);
  bool nk = dd > o->keyframe_thresh_dist || std::fabs(theta) > (double)o->p.keyframe_thresh_rot;
  if (dd <= o->keyframe_thresh_dist) nk = false;
  if (dd <= o->keyframe_thresh_dist && std::fabs(theta) > (double)o->p.keyframe_thresh_rot && num_nearby <= 1) nk = true;
  *is_new = nk;
  return nk ? push_keyframe(o) : NGICP_OK;
}

}  // namespace

extern "C" {

int ngicp_hull_planar(const double* xyz, int n, int concave, double alpha, int* out_indices) {
  if (!xyz || n <= 0 || !out_indices) return -1;
  std::vector<double> P(xyz, xyz + 3 * (size_t)n);
  int drop;
  if (hull_dimension(P, n, drop) == 3) return -3;
  std::vector<P2> Q(n);
  const int ax = drop == 0 ? 1 : 0, ay = drop == 2 ? 1 : 2;
  for (int i = 0; i < n; i++) Q[i] = {P[3 * i + ax], P[3 * i + ay]};
  std::vector<int> out;
  if (concave) concave_hull_2d(Q, alpha, out);
  else convex_hull_2d(Q, out);
  std::copy(out.begin(), out.end(), out_indices);
  return (int)out.size();
}

int ngicp_hull_spatial(const double* xyz, int n, int concave, double alpha, int* out_indices) {
  if (!xyz || n <= 0 || !out_indices) return -1;
  std::vector<P3> Q(n);
  for (int i = 0; i < n; i++) Q[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
  std::vector<int> out;
  if (!(concave ? concave_hull_3d(Q, alpha, out) : convex_hull_3d(Q, out))) return -2;
  std::copy(out.begin(), out.end(), out_indices);
  return (int)out.size();
}

void ngicp_odom_default_params(ngicp_odom_params* p) {
  if (!p) return;
  p->crop_size = 1.0f; p->voxel_res = 0.25f; p->keyframe_thresh_dist = 1.0f; p->keyframe_thresh_rot = 45.0f;      // cfg/params.yaml:43-49
  p->submap_knn = 10; p->submap_kcv = 10; p->submap_kcc = 10;                                                     // :53-55
  p->gicp_min_num_points = 64; p->gicp_max_corr_dist = 0.5f;                                                      // :57-59
  p->adaptive = 1;                                                                                                // dlio.yaml:17
  p->time_offset_bytes = 20; p->time_type = 0;                                                                    // dlio::Point `t`, uint32 ns (odom.cc:603-611)
}

int ngicp_odom_create(ngicp_handle* h, const ngicp_odom_params* p, ngicp_odom** out) {
  if (!h || !out) return NGICP_ERR_INVALID;
  ngicp_odom* o = new (std::nothrow) ngicp_odom;
  if (!o) return NGICP_ERR_INVALID;
  o->h = h;
  if (p) o->p = *p; else ngicp_odom_default_params(&o->p);
  identity16(o->T); identity16(o->T_prior); identity16(o->T_corr);
  o->keyframe_thresh_dist = o->p.keyframe_thresh_dist;
  o->concave_alpha = o->p.keyframe_thresh_dist;
  *out = o;
  return NGICP_OK;
}

int ngicp_odom_destroy(ngicp_odom* o) {
  if (!o) return NGICP_OK;
  for (Keyframe& k : o->keyframes) if (k.kf) ngicp_keyframe_release(o->h, k.kf);
  if (o->pack) cudaFreeHost(o->pack);
  delete o;
  return NGICP_OK;
}

const char* ngicp_odom_last_error(const ngicp_odom* o) { return o ? (o->err.empty() ? ngicp_last_error(o->h) : o->err.c_str()) : ""; }

int ngicp_odom_set_hull_callbacks(ngicp_odom* o, ngicp_hull_fn convex, ngicp_hull_fn concave, void* user) {
  if (!o) return NGICP_ERR_INVALID;
  o->convex_cb = convex; o->concave_cb = concave; o->cb_user = user;
  o->hull_n = -1;       // the cached hulls were made by the other route
  return NGICP_OK;
}

int ngicp_odom_get_profile(ngicp_odom* o, double seconds[NGICP_ODOM_STAGES], long* scans, int reset) {
  if (!o) return NGICP_ERR_INVALID;
  if (seconds) std::memcpy(seconds, o->prof, sizeof o->prof);
  if (scans) *scans = o->prof_scans;
  if (reset) { std::memset(o->prof, 0, sizeof o->prof); o->prof_scans = 0; }
  return NGICP_OK;
}

int ngicp_odom_set_pose(ngicp_odom* o, const float T_colmajor[16]) {
  if (!o || !T_colmajor) return NGICP_ERR_INVALID;
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) o->T[4 * r + c] = T_colmajor[4 * c + r];
  propagate(o);
  return NGICP_OK;
}

int ngicp_odom_scan_begin(ngicp_odom* o, const void* records, size_t n, size_t stride_bytes, double* unique_stamps, size_t* n_unique, size_t* n_kept) {
  if (!o || !records || !n_unique || n == 0) return NGICP_ERR_INVALID;
  o->err.clear();
  const int tt = o->p.time_type;
  const size_t toff = (size_t)o->p.time_offset_bytes, tsz = tt == 2 ? 8 : 4;
  if (tt < 0 || tt > 2 || stride_bytes < 12 || toff % tsz || toff + tsz > stride_bytes) { o->err = "odom loop: bad record stride or time-stamp offset"; return NGICP_ERR_INVALID; }
  StageClock clk(o);
  // One pass over the caller's records: (x, y, z, stamp) into a page-locked buffer — half (Ouster, Velodyne) or three quarters
  // (Hesai) of a 32-byte dlio::Point, and a DMA instead of a staged pageable copy — and the planar ranges of original_scan
  // (the cloud after removeNaN + CropBox, odom.cc:490-526) that computeSpaciousness takes the median of (:1398-1418).
  const size_t pstride = tt == 2 ? 24 : 16, ptoff = tt == 2 ? 16 : 12;
  if (o->pack_cap < n * pstride) {
    if (int rc = ngicp::select_device(reinterpret_cast<ngicp::Handle*>(o->h))) return rc;
    if (o->pack) cudaFreeHost(o->pack);
    o->pack = nullptr; o->pack_cap = 0;
    const size_t want = n * pstride + n * pstride / 4;
    if (cudaHostAlloc(&o->pack, want, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); o->err = "odom loop: page-locked staging buffer"; return NGICP_ERR_CUDA; }
    o->pack_cap = want;
  }
  const float cs = o->p.crop_size;
  if (o->ranges.size() < n + 4) o->ranges.resize(n + 4);
  float* rg = o->ranges.data();
  size_t nr = 0;
  const char* src = static_cast<const char*>(records);
  char* dst = static_cast<char*>(o->pack);
  size_t i = 0;
  uint32_t tmax = 0;
#if defined(__SSE2__)
  if (tt != 2 && stride_bytes >= 16) {
    // four records at a time: the packed record is one streaming 16-byte store, the crop test and the range are 4 wide
    const __m128 absmask = _mm_castsi128_ps(_mm_set1_epi32(0x7fffffff)), inf = _mm_set1_ps(INFINITY), vcs = _mm_set1_ps(cs);
    const __m128i lane3 = _mm_set_epi32(-1, 0, 0, 0);
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
    for (; i + 4 <= n; i += 4) {
      __m128 a[4];
      for (int u = 0; u < 4; u++) {
        const char* rec = src + (i + u) * stride_bytes;
        _mm_prefetch(rec + 2048, _MM_HINT_NTA);     // across page boundaries, where the hardware prefetcher stops
        uint32_t t;
        std::memcpy(&t, rec + toff, 4);
        tmax = std::max(tmax, t);
        a[u] = _mm_loadu_ps(reinterpret_cast<const float*>(rec));
        const __m128i packed = _mm_or_si128(_mm_andnot_si128(lane3, _mm_castps_si128(a[u])), _mm_and_si128(lane3, _mm_set1_epi32((int)t)));
        if (aligned) _mm_stream_si128(reinterpret_cast<__m128i*>(dst + (i + u) * 16), packed);
        else _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + (i + u) * 16), packed);
      }
      _MM_TRANSPOSE4_PS(a[0], a[1], a[2], a[3]);      // a[0] = x of the four records, a[1] = y, a[2] = z
      const __m128 ax = _mm_and_ps(a[0], absmask), ay = _mm_and_ps(a[1], absmask), az = _mm_and_ps(a[2], absmask);
      const __m128 m = _mm_max_ps(_mm_max_ps(ax, ay), az);
      const __m128 finite = _mm_and_ps(_mm_and_ps(_mm_cmplt_ps(ax, inf), _mm_cmplt_ps(ay, inf)), _mm_cmplt_ps(az, inf));   // false for NaN
      const int keep = _mm_movemask_ps(_mm_and_ps(finite, _mm_cmpge_ps(m, vcs)));
      float r4[4];
      _mm_storeu_ps(r4, _mm_sqrt_ps(_mm_add_ps(_mm_mul_ps(a[0], a[0]), _mm_mul_ps(a[1], a[1]))));
      for (int u = 0; u < 4; u++) { rg[nr] = r4[u]; nr += (keep >> u) & 1; }
    }
    if (aligned) _mm_sfence();
  }
#endif
  for (; i < n; i++) {
    const char* rec = src + i * stride_bytes;
    float q[3];
    std::memcpy(q, rec, 12);
    char* d = dst + i * pstride;
    std::memcpy(d, q, 12);
    if (tt == 2) { const uint32_t z = 0; std::memcpy(d + 12, &z, 4); std::memcpy(d + 16, rec + toff, 8); }
    else { uint32_t t; std::memcpy(&t, rec + toff, 4); tmax = std::max(tmax, t); std::memcpy(d + 12, &t, 4); }
    const float ax = std::fabs(q[0]), ay = std::fabs(q[1]), az = std::fabs(q[2]);
    const bool finite = ax < INFINITY && ay < INFINITY && az < INFINITY;     // false for NaN
    const float xx = q[0] * q[0], yy = q[1] * q[1];     // no contraction: host code is built without FMA
    rg[nr] = std::sqrt(xx + yy);
    nr += (finite && std::max(std::max(ax, ay), az) >= cs) ? 1 : 0;
  }
  o->n_ranges = nr;
  o->median_ready = false;
  clk.lap(1);
  // the median is taken while the device crops and sorts the scan (the hook runs inside the ingest's wait)
  ngicp::Handle* hh = reinterpret_cast<ngicp::Handle*>(o->h);
  if (tt == 0) { int bits = 0; while (bits < 32 && (tmax >> bits)) bits++; hh->ingest_stamp_bits = bits; }
  hh->overlap_arg = o;
  hh->overlap_fn = [](void* arg) {
    ngicp_odom* od = static_cast<ngicp_odom*>(arg);
    if (od->n_ranges) {
      const size_t mid = od->n_ranges / 2;
      std::nth_element(od->ranges.begin(), od->ranges.begin() + mid, od->ranges.begin() + od->n_ranges);
      od->median_curr = od->ranges[mid];
    }
    od->median_ready = true;
  };
  const float mn[3] = {-cs, -cs, -cs}, mx[3] = {cs, cs, cs};
  size_t nu = 0, nk = 0;
  const int rc = ngicp_scan_ingest(o->h, o->pack, n, pstride, ptoff, tt, mn, mx, 1, unique_stamps, &nu, &nk);
  hh->overlap_fn = nullptr;
  if (rc) return rc;
  clk.lap(0);
  o->n_unique = nu; o->n_kept = nk;
  *n_unique = nu;
  if (n_kept) *n_kept = nk;
  return NGICP_OK;
}

int ngicp_odom_scan_finish(ngicp_odom* o, const float* frames16, size_t n_frames, ngicp_odom_result* res, int* submap_ids, int submap_cap) {
  if (!o || !res) return NGICP_ERR_INVALID;
  std::memset(res, 0, sizeof *res);
  res->valid = 0;
  if (o->n_kept == 0) return NGICP_OK;
  const bool first = o->keyframes.empty();
  float cm[16];
  const float* frames = frames16;
  size_t nf = n_frames;
  if (first || !frames16) {
    to_colmajor(o->T, cm);
    frames = cm; nf = 1;
    std::memcpy(o->T_prior, o->T, sizeof o->T);
  } else {
    if (n_frames != o->n_unique && n_frames != 1) { o->err = "odom loop: one prior frame per unique time stamp"; return NGICP_ERR_INVALID; }
    const float* mid = frames16 + 16 * (n_frames == 1 ? 0 : o->n_unique / 2);
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) o->T_prior[4 * r + c] = mid[4 * c + r];
  }
  const float leaf[3] = {o->p.voxel_res, o->p.voxel_res, o->p.voxel_res};
  size_t n_src = 0;
  StageClock clk(o);
  if (int rc = ngicp_scan_deskew(o->h, frames, nf, o->p.voxel_res > 0 ? leaf : nullptr, NGICP_SOURCE, nullptr, &n_src)) return rc;
  clk.lap(2);
  res->n_points = (int)n_src;
  if ((int)n_src <= o->p.gicp_min_num_points) return NGICP_OK;          // "Low number of points in the cloud!" (odom.cc:764-767)

  // metrics and adaptive parameters, in the reference's order (odom.cc:769-779): the density is still the previous scan's
  if (o->n_ranges) {
    if (!o->median_ready) {
      const size_t mid = o->n_ranges / 2;
      std::nth_element(o->ranges.begin(), o->ranges.begin() + mid, o->ranges.begin() + o->n_ranges);
      o->median_curr = o->ranges[mid];
      o->median_ready = true;
    }
    const float median_curr = o->median_curr;
    if (!o->have_median) { o->median_prev = median_curr; o->have_median = true; }
    volatile float a = 0.95f * o->median_prev, b = 0.05f * median_curr;
    const float lpf = a + b;
    o->median_prev = lpf;
    o->spaciousness = (double)lpf;
  }
  {
    const float d = o->first_opt_done ? o->source_density : 0.f;
    if (!o->have_density) { o->density_prev = d; o->have_density = true; }
    volatile float a = 0.95f * o->density_prev, b = 0.05f * d;
    const float lpf = a + b;
    o->density_prev = lpf;
    o->density = (double)lpf;
  }
  if (o->p.adaptive) {                                                   // setAdaptiveParams, odom.cc:1600-1626
    const double sp = std::min(std::max(o->spaciousness, 0.5), 5.0);
    o->keyframe_thresh_dist = sp;
    const double mcd = (double)o->p.gicp_max_corr_dist;
    double den = std::min(std::max(o->density, 0.5 * mcd), 2.0 * mcd);
    if (sp < 5.0) den = 0.5 * mcd;
    if (sp > 5.0) den = 2.0 * mcd;
    ngicp_params prm;
    ngicp_get_params(o->h, &prm);
    prm.max_corr_dist = den;
    if (int rc = ngicp_set_params(o->h, &prm)) return rc;
    o->concave_alpha = o->keyframe_thresh_dist;
  }
  clk.lap(3);
  if (int rc = ngicp_compute_covariances(o->h, NGICP_SOURCE, &o->source_density)) return rc;
  clk.lap(4);

  bool new_kf = false;
  int conv = 1, iters = 0;
  bool changed = true;
  if (first) {                                                           // initializeInputTarget, odom.cc:708-718
    if (int rc = push_keyframe(o)) return rc;
    clk.lap(6);
    if (int rc = build_keyframes_and_submap(o)) return rc;
    clk.lap(7);
    new_kf = true;
  } else {
    changed = o->submap_changed;
    o->submap_changed = false;                                           // getNextPose, odom.cc:984-1018
    float Tc[16];
    if (int rc = ngicp_align(o->h, nullptr, Tc, &iters, &conv, nullptr, nullptr)) { if (rc != NGICP_ERR_LM_NOT_CONVERGED) return rc; }
    clk.lap(5);
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) o->T_corr[4 * r + c] = Tc[4 * c + r];
    float Tn[16];
    matmul4(o->T_corr, o->T_prior, Tn);
    std::memcpy(o->T, Tn, sizeof Tn);
    propagate(o);
    if (int rc = update_keyframes(o, &new_kf)) return rc;
    clk.lap(6);
    if (int rc = build_keyframes_and_submap(o)) return rc;
    clk.lap(7);
    o->first_opt_done = true;
  }
  o->prof_scans++;
  res->valid = 1;
  to_colmajor(o->T, res->T);
  to_colmajor(o->T_corr, res->T_corr);
  res->converged = conv; res->iterations = iters; res->new_keyframe = new_kf ? 1 : 0; res->submap_changed = changed ? 1 : 0;
  res->n_keyframes = (int)o->keyframes.size();
  res->n_submap = (int)o->submap_curr.size();
  if (submap_ids) for (int i = 0; i < res->n_submap && i < submap_cap; i++) submap_ids[i] = o->submap_curr[i];
  return NGICP_OK;
}

}  // extern "C"
