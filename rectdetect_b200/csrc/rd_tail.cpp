// rd_tail.cpp - the host tail of the rectangle detector: quad assembly and pose estimation from the line segments
// and the (segment x region) vote table (executeCPUTask, oclrect.c:1049-1226, and its helpers oclrect.c:385-1045).
//
// The reference walks three full-size read-backs (segment list, region map, vote table).  Here the device has
// already looked up everything the tail will ask for (rd_rect.cu : k_tail_gather), so the input is the segment list
// plus, per segment, 15 (region id, vote entry) samples.  rd_rect_tail() offers the reference's view (full arrays)
// by doing that gather on the host first.  All arithmetic is IEEE double in the reference's order of operations
// (compiled with -ffp-contract=off); the two qsort calls are stable sorts.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <functional>
#include <vector>
#include "../../include/rectdetect_b200.h"

struct rd_tail_sample { int32_t segid; int32_t vote[5]; };
typedef void (*rd_parallel_for_t)(int n, const std::function<void(int)> &fn);   // as in rd_common.cuh
#define RD_TAIL_NSAMPLE 15

namespace {

struct P2 { double x, y; };
struct P3 { double x, y, z; };
struct P4 { double v[4]; };
struct Edge { P2 a, b; };                                   // ls_t, oclrect.c:385-387

inline P2 operator+(P2 p, P2 q) { return {p.x + q.x, p.y + q.y}; }
inline P2 operator-(P2 p, P2 q) { return {p.x - q.x, p.y - q.y}; }
inline P2 operator*(P2 p, double s) { return {p.x * s, p.y * s}; }
inline double dot(P2 p, P2 q) { double s = 0; s += p.x * q.x; s += p.y * q.y; return s; }
inline double norm2(P2 p) { return dot(p, p); }
inline double dist2(P2 p, P2 q) { return norm2(p - q); }
inline P2 unit(P2 p) { return p * (1.0 / (sqrt(norm2(p)) + 1e-20)); }                 // normalize2, vec234.h

inline P3 operator+(P3 p, P3 q) { return {p.x + q.x, p.y + q.y, p.z + q.z}; }
inline P3 operator-(P3 p, P3 q) { return {p.x - q.x, p.y - q.y, p.z - q.z}; }
inline P3 operator*(P3 p, double s) { return {p.x * s, p.y * s, p.z * s}; }
inline double dot(P3 p, P3 q) { double s = 0; s += p.x * q.x; s += p.y * q.y; s += p.z * q.z; return s; }
inline double norm2(P3 p) { return dot(p, p); }
inline double dist2(P3 p, P3 q) { return norm2(p - q); }
inline P3 unit(P3 p) { return p * (1.0 / (sqrt(norm2(p)) + 1e-20)); }
inline P3 cross(P3 v, P3 w) { return {v.y * w.z - v.z * w.y, v.z * w.x - v.x * w.z, v.x * w.y - v.y * w.x}; }

inline P4 add4(P4 p, P4 q) { P4 r; for (int i = 0; i < 4; i++) r.v[i] = p.v[i] + q.v[i]; return r; }
inline P4 sub4(P4 p, P4 q) { P4 r; for (int i = 0; i < 4; i++) r.v[i] = p.v[i] - q.v[i]; return r; }
inline P4 mul4(P4 p, double s) { P4 r; for (int i = 0; i < 4; i++) r.v[i] = p.v[i] * s; return r; }
inline double dot4(P4 p, P4 q) { double s = 0; for (int i = 0; i < 4; i++) s += p.v[i] * q.v[i]; return s; }
inline P4 unit4(P4 p) { return mul4(p, 1.0 / (sqrt(dot4(p, p)) + 1e-20)); }

inline double sq(double x) { return x * x; }
inline float edgeLen2f(const Edge &e) { return (float)dist2(e.a, e.b); }             // lsSquLen returns float (oclrect.c:390)

// foot of the perpendicular from p on the line through v, w (oclrect.c:400-406), and on the segment (:408-416)
inline P2 footOnLine(P2 v, P2 w, P2 p) {
  const double l2 = dist2(v, w);
  if (l2 == 0.0) return v;
  const double t = ((p.x - v.x) * (w.x - v.x) + (p.y - v.y) * (w.y - v.y)) / l2;
  return {v.x + t * (w.x - v.x), v.y + t * (w.y - v.y)};
}
inline P2 footOnSegment(P2 v, P2 w, P2 p) {
  const double l2 = dist2(v, w);
  if (l2 == 0.0) return v;
  const double t = ((p.x - v.x) * (w.x - v.x) + (p.y - v.y) * (w.y - v.y)) / l2;
  if (t < 0) return v;
  else if (t > 1.0) return w;
  return {v.x + t * (w.x - v.x), v.y + t * (w.y - v.y)};
}
// oclrect.c:418-425
inline P2 lineIntersection(const Edge &u, const Edge &v) {
  const double d = (v.b.x - v.a.x) * (u.b.y - u.a.y) - (v.b.y - v.a.y) * (u.b.x - u.a.x);
  if (fabs(d) < 1e-4) return {NAN, NAN};
  const double n = (v.a.y - u.a.y) * (u.b.x - u.a.x) - (v.a.x - u.a.x) * (u.b.y - u.a.y);
  const double q = n / d;
  return {v.a.x + q * (v.b.x - v.a.x), v.a.y + q * (v.b.y - v.a.y)};
}

// ------------------------------------------------------------------ pose estimation (oclrect.c:427-634)
// The four image corners define four rays; the unknowns are the depths along them.  The objective scores how far
// the scaled points are from a planar rectangle (two variants that normalise a different side to length 1).
struct Pose { const P3 *ray; int mode; };
const double H = 1e-6;

double objective(P4 d, const Pose &ps) {                                               // oclrect.c:441-477
  const int m = ps.mode;
  P3 q[4];
  for (int i = 0; i < 4; i++) q[i] = ps.ray[i] * d.v[i];
  double score = 0;
  const double l01 = dist2(q[0], q[1]), l12 = dist2(q[1], q[2]), l23 = dist2(q[2], q[3]);
  const double l03 = dist2(q[0], q[3]), l02 = dist2(q[0], q[2]), l13 = dist2(q[1], q[3]);
  score += sq((m ? l23 : l03) - 1);
  score += sq((m ? l01 : l12) - 1);
  const double comp = 1.0 / (m ? l12 : l01);
  score += norm2(((m ? q[0] : q[2]) - q[1]) + ((m ? q[2] : q[0]) - q[3]));
  score += comp * norm2((q[1] - (m ? q[2] : q[0])) + (q[3] - (m ? q[0] : q[2])));
  score += sq(l01 + l12 - l02);
  score += sq(l03 + l23 - l02);
  score += sq(l01 + l03 - l13);
  score += sq(l12 + l23 - l13);
  const P3 n013 = cross(q[1] - q[0], q[3] - q[0]);
  score += comp * sq(dot(n013, q[2]) - dot(n013, q[0])) / dot(n013, n013);
  const P3 n102 = cross(q[0] - q[1], q[2] - q[1]);
  score += comp * sq(dot(n102, q[3]) - dot(n102, q[1])) / dot(n102, n102);
  return score;
}

// value, first and second directional derivative by central differences (oclrect.c:479-490)
inline void directional(P4 x, P4 dir, const Pose &ps, double &f0, double &d1, double &d2) {
  f0 = objective(x, ps);
  const double fp = objective(add4(x, mul4(dir, H)), ps);
  const double fm = objective(add4(x, mul4(dir, -H)), ps);
  d1 = (fp - fm) * (1.0 / (2 * H));
  d2 = (fp + fm - 2 * f0) * (1.0 / (H * H));
}
// gradient and diagonal of the Hessian (oclrect.c:492-512)
inline void gradDiag(P4 x, const Pose &ps, P4 &g, P4 &h) {
  const double fx = objective(x, ps);
  for (int i = 0; i < 4; i++) {
    P4 e;
    for (int j = 0; j < 4; j++) { e.v[j] = 0; if (j == i) e.v[j] = H; }
    const double fm = objective(sub4(x, e), ps);
    const double fp = objective(add4(x, e), ps);
    g.v[i] = (fp - fm) / (2 * H);
    h.v[i] = (fm - 2 * fx + fp) / (H * H);
  }
}
// Newton steps along dir with step halving (oclrect.c:514-536)
inline P4 lineSearch(P4 x, P4 dir, int iters, const Pose &ps) {
  dir = unit4(dir);
  double sc = 1.0;
  for (int i = 0; i < iters; i++) {
    double f0, d1, d2;
    directional(x, dir, ps, f0, d1, d2);
    if (d2 * d2 < 1e-10) d2 = 1;
    const double delta = fabs(d1 / d2);
    if (delta < 1e-10) return x;
    const P4 cand = add4(x, mul4(dir, delta * sc));
    const double f1 = objective(cand, ps);
    if (f0 < f1) { sc *= 0.5; continue; }
    x = cand;
  }
  return x;
}
// Jacobi preconditioner: r / diag when every diagonal entry is positive (oclrect.c:538-555)
inline P4 precondition(P4 diag, P4 r) {
  for (int i = 0; i < 4; i++) if (diag.v[i] <= 0) return r;
  P4 a;
  for (int i = 0; i < 4; i++) { a.v[i] = 1.0 / diag.v[i]; a.v[i] *= r.v[i]; }
  return a;
}
// preconditioned nonlinear CG, Polak-Ribiere with restart every 10 steps (oclrect.c:557-588)
P4 conjugateGradient(P4 x, int outer, int inner, const Pose &ps) {
  int k = 0;
  P4 g, h;
  gradDiag(x, ps, g, h);
  P4 r = mul4(g, -1);
  P4 s = precondition(h, r), d = s;
  double deltaNew = dot4(r, d);
  for (int i = 0; i < outer; i++) {
    x = lineSearch(x, d, inner, ps);
    gradDiag(x, ps, g, h);
    r = mul4(g, -1);
    const double deltaOld = deltaNew;
    const double deltaMid = dot4(r, s);
    s = precondition(h, r);
    deltaNew = dot4(r, s);
    const double beta = (deltaNew - deltaMid) / deltaOld;
    if (k == 10 || beta <= 0 || deltaOld == 0) { d = s; k = 0; }
    else d = add4(s, mul4(d, beta));
    k++;
  }
  return x;
}

// oclrect.c:590-634 : corners are the start points of the four edges, rotated so that the edge facing up comes first
void estimatePose(const Edge *e, P2 centre, int iw, int ih, double tanAOV, rect_t *out) {
  int first = 0;
  double mn = 1e+100;
  for (int i = 0; i < 4; i++) {
    P2 v = unit(e[i].b - e[i].a);
    v = {-v.y, v.x};
    if (dot(e[i].a - centre, v) < 0) v = v * -1;
    if (v.y < mn) { mn = v.y; first = i; }
  }
  P3 ray[4];
  for (int i = 0; i < 4; i++) {
    const P2 c = e[(i + first) & 3].a;
    ray[i] = unit(P3{(c.x - (iw / 2)), (-(c.y - ih / 2)), iw / 2 / tanAOV});
  }
  const double d01 = 1.0 / sqrt(dist2(ray[0], ray[1])), d23 = 1.0 / sqrt(dist2(ray[2], ray[3]));
  const Pose p1 = {ray, 1};
  const P4 x0 = conjugateGradient(P4{{d01, d01, d23, d23}}, 12, 10, p1);
  const double val0 = objective(x0, p1);
  const double d12 = 1.0 / sqrt(dist2(ray[1], ray[2])), d03 = 1.0 / sqrt(dist2(ray[0], ray[3]));
  const Pose p0 = {ray, 0};
  const P4 x1 = conjugateGradient(P4{{d03, d12, d12, d03}}, 12, 10, p0);
  const double val1 = objective(x1, p0);

  out->value = val0 < val1 ? val0 : val1;
  P4 x = val0 < val1 ? x0 : x1;
  if (x.v[0] < 0) x = mul4(x, -1);
  for (int i = 0; i < 4; i++) {
    const P3 c = ray[i] * x.v[i];
    out->c3[i].a[0] = c.x; out->c3[i].a[1] = c.y; out->c3[i].a[2] = c.z;
    out->c2[i].a[0] = e[(i + first) & 3].a.x;
    out->c2[i].a[1] = e[(i + first) & 3].a.y;
  }
}

// oclrect.c:636-656
int looksLikeAScreen(const rect_t &r) {
  if (r.value > 0.05) return 0;
  P3 c3[4]; P2 c2[4];
  for (int i = 0; i < 4; i++) { c3[i] = {r.c3[i].a[0], r.c3[i].a[1], r.c3[i].a[2]}; c2[i] = {r.c2[i].a[0], r.c2[i].a[1]}; }
  if (c3[0].z < 0 || c3[1].z < 0 || c3[2].z < 0 || c3[3].z < 0) return 0;
  const double asp = sqrt(dist2(c3[0], c3[1])) / sqrt(dist2(c3[1], c3[2]));
  if (asp < 1.0 / 12 || 12 < asp) return 0;
  double maxs = 0, mins = 1e+100;
  for (int i = 0; i < 4; i++) {
    const double s0 = dist2(c2[(i + 2) % 4], footOnSegment(c2[i], c2[(i + 1) % 4], c2[(i + 2) % 4]));
    const double s1 = dist2(c2[(i + 3) % 4], footOnSegment(c2[i], c2[(i + 1) % 4], c2[(i + 3) % 4]));
    maxs = fmax(maxs, fmax(s0, s1));
    mins = fmin(mins, fmax(s0, s1));
  }
  if (maxs / mins > 100) return 0;
  return 1;
}

// ------------------------------------------------------------------ quick hull (oclrect.c:658-734)
void hullSide(std::vector<P2> &hull, const std::vector<P2> &pts, P2 left, P2 right) {
  int far = -1;
  double d = 0;
  for (int i = 0; i < (int)pts.size(); i++) {
    const double e = dist2(footOnLine(left, right, pts[i]), pts[i]);
    if (far < 0 || e > d) { far = i; d = e; }
  }
  if (d < 0.01 || far < 0) return;
  const P2 pf = pts[far];
  const P2 nr = {pf.y - right.y, right.x - pf.x}, nl = {left.y - pf.y, pf.x - left.x};
  std::vector<P2> sr, sl;
  for (int i = 0; i < (int)pts.size(); i++) {
    if (i == far) continue;
    if (dot(pts[i] - pf, nr) > 0) sr.push_back(pts[i]);
    if (dot(pts[i] - pf, nl) > 0) sl.push_back(pts[i]);
  }
  hullSide(hull, sr, pf, right);
  hull.push_back(pf);
  hullSide(hull, sl, left, pf);
}
std::vector<P2> convexHull(const std::vector<P2> &pts) {
  std::vector<P2> hull;
  if (pts.empty()) return hull;
  P2 right = pts[0], left = pts[0];
  for (const P2 &p : pts) {
    if (p.x > right.x) right = p;
    if (p.x < left.x) left = p;
  }
  const P2 up = {left.y - right.y, right.x - left.x};
  std::vector<P2> top, bot;
  for (const P2 &p : pts) {
    if (p.x == left.x && p.y == left.y) continue;
    if (p.x == right.x && p.y == right.y) continue;
    if (dot(p - left, up) > 0) top.push_back(p); else bot.push_back(p);
  }
  hull.push_back(right);
  hullSide(hull, top, left, right);
  hull.push_back(left);
  hullSide(hull, bot, right, left);
  return hull;
}

// ------------------------------------------------------------------ Cohen-Sutherland clip (oclrect.c:744-802)
inline int outcode(double x, double y, double xmin, double ymin, double xmax, double ymax) {
  int c = 0;
  if (x < xmin) c |= 1;
  if (x > xmax) c |= 2;
  if (y < ymin) c |= 4;
  if (y > ymax) c |= 8;
  return c;
}
bool clipToBox(double &x0, double &y0, double &x1, double &y1, double xmin, double ymin, double xmax, double ymax) {
  int c0 = outcode(x0, y0, xmin, ymin, xmax, ymax), c1 = outcode(x1, y1, xmin, ymin, xmax, ymax);
  for (;;) {
    if ((c0 | c1) == 0) return true;
    if ((c0 & c1) != 0) return false;
    double x = 0, y = 0;
    const int co = c0 != 0 ? c0 : c1;
    if (co & 8) { x = x0 + (x1 - x0) * (ymax - y0) / (y1 - y0); y = ymax; }
    else if (co & 4) { x = x0 + (x1 - x0) * (ymin - y0) / (y1 - y0); y = ymin; }
    else if (co & 2) { y = y0 + (y1 - y0) * (xmax - x0) / (x1 - x0); x = xmax; }
    else if (co & 1) { y = y0 + (y1 - y0) * (xmin - x0) / (x1 - x0); x = xmin; }
    if (co == c0) { x0 = x; y0 = y; c0 = outcode(x0, y0, xmin, ymin, xmax, ymax); }
    else { x1 = x; y1 = y; c1 = outcode(x1, y1, xmin, ymin, xmax, ymax); }
  }
}

// ------------------------------------------------------------------ edge-list filters (oclrect.c:806-1045)
void sortByLength(std::vector<Edge> &es) {
  std::stable_sort(es.begin(), es.end(), [](const Edge &p, const Edge &q) { return edgeLen2f(p) < edgeLen2f(q); });
}
double outwardAngle(const Edge &e, P2 centre) {                                        // oclrect.c:829-834
  P2 v = e.a - e.b;
  v = {v.y, -v.x};
  if (dot(v, e.a - centre) < 0) v = v * -1;
  return atan2(v.x, v.y);
}
void sortByAngle(std::vector<Edge> &es, P2 centre) {
  std::stable_sort(es.begin(), es.end(), [centre](const Edge &p, const Edge &q) { return outwardAngle(p, centre) < outwardAngle(q, centre); });
}
P2 weightedCentre(const std::vector<Edge> &es) {                                       // gv, oclrect.c:864-877
  P2 g = {0, 0};
  double total = 0;
  for (const Edge &e : es) {
    const double len = sqrt(dist2(e.a, e.b));
    g = g + (e.a + e.b) * len;
    total += len;
  }
  return g * (0.5 / total);
}
double totalLength(const std::vector<Edge> &es) {                                      // oclrect.c:879-884 (sqrt of the float-rounded square)
  double s = 0;
  for (const Edge &e : es) s += sqrt((double)edgeLen2f(e));
  return s;
}
bool nearlyTriangle(const std::vector<Edge> &es, double ratio) {                       // oclrect.c:886-895
  const int n = (int)es.size();
  for (int i = 0; i < n; i++) {
    const Edge &e0 = es[i], &e1 = es[(i + 1) % n];
    const double d0 = dist2(e0.b, footOnLine(e0.a, e1.b, e0.b));
    const double d1 = dist2(e0.a, e1.b);
    if (d0 / d1 < ratio) return true;
  }
  return false;
}
bool convex(const std::vector<Edge> &es) {                                             // oclrect.c:897-922
  const int n = (int)es.size();
  auto turn = [](const Edge &p, const Edge &q) { return (p.b.x - p.a.x) * (q.b.y - q.a.y) - (p.b.y - p.a.y) * (q.b.x - q.a.x) > 0; };
  const bool sign = turn(es[0], es[1]);
  for (int i = 1; i < n; i++) if (sign != turn(es[i], es[(i + 1) % n])) return false;
  return true;
}
void dropShort(std::vector<Edge> &es, float ratio) {                                   // removeShortLS, oclrect.c:926-943
  if (es.size() <= 4) return;
  sortByLength(es);
  const float longest = edgeLen2f(es.back());
  while (es.size() > 4) {
    const float shortest = edgeLen2f(es[0]);
    if (shortest / longest > ratio * ratio) break;
    es.erase(es.begin());
  }
}
// keep, per hull edge, the longest segment lying on it (oclrect.c:945-992)
std::vector<Edge> keepOuter(std::vector<Edge> es) {
  std::vector<P2> pts;
  for (const Edge &e : es) { pts.push_back(e.a); pts.push_back(e.b); }
  const std::vector<P2> hull = convexHull(pts);
  std::vector<Edge> kept;
  const int hs = (int)hull.size();
  for (int i = 0; i < hs; i++) {
    const P2 q0 = hull[i], q1 = hull[(i + 1) % hs];
    const P2 mid = (q0 + q1) * 0.5, dir = unit(q0 - q1);
    int taken = -1;
    sortByLength(es);
    for (int j = (int)es.size() - 1; j >= 0; j--) {
      const Edge e = es[j];
      if (dist2(mid, footOnSegment(e.a, e.b, mid)) < 1) { kept.push_back(e); taken = j; break; }
      if (fabs(dot(dir, unit(e.a - e.b))) > 0.95 && dist2(mid, footOnSegment(e.a, e.b, mid)) / dist2(q0, q1) < 0.01) { kept.push_back(e); taken = j; break; }
    }
    if (taken != -1) es.erase(es.begin() + taken);
  }
  return kept;
}
std::vector<Edge> keepLongest(std::vector<Edge> es, int n) {                           // oclrect.c:994-1009
  if ((int)es.size() <= n) return es;
  sortByLength(es);
  std::vector<Edge> r;
  for (int j = (int)es.size() - 1; j >= 0 && (int)r.size() < n; j--) r.push_back(es[j]);
  return r;
}
bool toCorners(std::vector<Edge> &es) {                                                // findCorners, oclrect.c:1011-1045
  const int n = (int)es.size();
  std::vector<P2> c(n);
  for (int i = 0; i < n; i++) {
    c[i] = lineIntersection(es[i], es[(i + 1) % n]);
    if (isnan(c[i].x)) return false;
  }
  for (int i = 0; i < n; i++) { es[i].a = c[i]; es[i].b = c[(i + 1) % n]; }
  return true;
}

// the candidate test shared by both loops of executeCPUTask (oclrect.c:1134-1160, 1190-1216)
bool tryQuad(std::vector<Edge> es, uint32_t status, int iw, int ih, double tanAOV, rect_t &r) {
  dropShort(es, 0.05f);
  es = keepOuter(es);
  const double len0 = totalLength(es);
  es = keepLongest(es, 4);
  sortByAngle(es, weightedCentre(es));
  if (!toCorners(es)) return false;
  const double len1 = totalLength(es);
  if (nearlyTriangle(es, 0.001) || es.size() < 4 || len1 / len0 > 2 || !convex(es)) return false;
  memset(&r, 0, sizeof(r));
  estimatePose(es.data(), weightedCentre(es), iw, ih, tanAOV, &r);
  r.status = status;
  if (looksLikeAScreen(r)) r.status |= 1;
  return true;
}

inline int bucketOf(uint64_t key) { return (int)((key ^ (key >> 10) ^ (key >> 20) ^ (key >> 30)) & 1023); }   // helper.c:129-131

}  // namespace

// executeCPUTask on the compact record.  ls: n+1 entries (entry 0 = header); samples: (n+1) x 15.
// The candidates are independent (each ends in a pose estimation, which is where the time goes); `pfor`, when given, runs
// them in parallel - the results are collected in candidate order either way, so the list does not depend on it.
rect_t *rd_tail_compact(const linesegment_t *ls, const rd_tail_sample *samples, int iw, int ih, double tanAOV, rd_parallel_for_t pfor) {
  const int n = *(const int32_t *)ls;
  struct Candidate { std::vector<Edge> es; uint32_t status; };
  std::vector<Candidate> cands;

  // (i) regions each live segment touches, grouped by region in ArrayMap iteration order (bucket, then first insertion)
  struct Region { int segid; std::vector<int> lsids; std::vector<const int32_t *> votes; };
  std::vector<std::vector<Region>> buckets(1024);
  for (int i = 1; i <= n; i++) {
    if (ls[i].polyid == 0) continue;
    for (int k = 0; k < RD_TAIL_NSAMPLE; k++) {
      const rd_tail_sample &sm = samples[(size_t)i * RD_TAIL_NSAMPLE + k];
      if (sm.segid <= 0) continue;
      std::vector<Region> &b = buckets[bucketOf((uint64_t)sm.segid)];
      Region *r = NULL;
      for (Region &c : b) if (c.segid == sm.segid) { r = &c; break; }
      if (!r) { b.push_back(Region()); r = &b.back(); r->segid = sm.segid; }
      if (std::find(r->lsids.begin(), r->lsids.end(), i) == r->lsids.end()) { r->lsids.push_back(i); r->votes.push_back(sm.vote); }
    }
  }

  // (ii) one candidate per region with at least 4 segments: clip each segment to where it touches the region
  for (int bk = 0; bk < 1024; bk++)
    for (const Region &r : buckets[bk]) {
      if (r.lsids.size() < 4) continue;
      std::vector<Edge> es;
      for (size_t j = 0; j < r.lsids.size(); j++) {
        const int id = r.lsids[j];
        const int32_t *v = r.votes[j];
        if (v[0] != id) {                                     // slot owned by another segment (hash collision): unclipped
          if (v[0] != 0) es.push_back(Edge{{ls[id].x0, ls[id].y0}, {ls[id].x1, ls[id].y1}});
          continue;
        }
        double x0 = ls[id].x0, y0 = ls[id].y0, x1 = ls[id].x1, y1 = ls[id].y1;
        if (!clipToBox(x0, y0, x1, y1, iw - v[1], ih - v[3], v[2], v[4])) continue;
        es.push_back(Edge{{x0, y0}, {x1, y1}});
      }
      cands.push_back(Candidate{std::move(es), 0u});
    }

  // (iii) one candidate per polyline chain, from its segments longer than 32 px
  for (int i = 1; i <= n; i++) {
    if (ls[i].polyid == 0 || ls[i].leftPtr > 0) continue;
    std::vector<Edge> es;
    for (int j = i; j > 0; j = ls[j].rightPtr) {
      const P2 a = {ls[j].x0, ls[j].y0}, b = {ls[j].x1, ls[j].y1};
      if (dist2(a, b) > 32.0 * 32.0) es.push_back(Edge{a, b});
    }
    cands.push_back(Candidate{std::move(es), 2u});
  }

  std::vector<rect_t> rects(cands.size());
  std::vector<char> ok(cands.size(), 0);
  auto one = [&](int i) { ok[i] = tryQuad(cands[i].es, cands[i].status, iw, ih, tanAOV, rects[i]) ? 1 : 0; };
  if (pfor && cands.size() > 1) pfor((int)cands.size(), one);
  else for (size_t i = 0; i < cands.size(); i++) one((int)i);
  size_t nfound = 0;
  for (size_t i = 0; i < cands.size(); i++) nfound += ok[i];
  rect_t *out = (rect_t *)calloc(nfound + 1, sizeof(rect_t));
  size_t k = 1;
  for (size_t i = 0; i < cands.size(); i++) if (ok[i]) out[k++] = rects[i];
  out[0].nItems = (int)nfound + 1;
  return out;
}

// the sampling step of executeCPUTask (oclrect.c:1066-1098) on full-size arrays; mirrors k_tail_gather
void rd_tail_gather_host(const linesegment_t *ls, const int32_t *segidMap, const int32_t *votes, int iw, int ih, rd_tail_sample *out) {
  const int n = *(const int32_t *)ls;
  const unsigned nentry = (unsigned)(iw * ih * 4 / 5);
  memset(out, 0, sizeof(rd_tail_sample) * (size_t)(n + 1) * RD_TAIL_NSAMPLE);
  for (int i = 1; i <= n; i++) {
    if (ls[i].polyid == 0) continue;
    const P2 s = {rint(ls[i].x0), rint(ls[i].y0)}, e = {rint(ls[i].x1), rint(ls[i].y1)};
    const P2 d = unit(e - s), nrm = {-d.y, d.x};
    int k = 0;
    for (int j = 0; j < 3; j++)
      for (int off = -2; off <= 2; off++, k++) {
        const P2 p = s + (e - s) * ((j + 0.5) / 3);
        const P2 c = p + nrm * off;
        const int x = (int)(c.x + 0.5), y = (int)(c.y + 0.5);
        if (x < 0 || x >= iw || y < 0 || y >= ih) continue;
        const int segid = segidMap[x + y * iw];
        if (segid <= 0) continue;
        const int hash = (int)((((uint32_t)i * (uint32_t)segid) & 0x7fffffff) % nentry);
        rd_tail_sample &sm = out[(size_t)i * RD_TAIL_NSAMPLE + k];
        sm.segid = segid;
        for (int q = 0; q < 5; q++) sm.vote[q] = votes[(size_t)hash * 5 + q];
      }
  }
}

extern "C" rect_t *rd_rect_tail(const linesegment_t *ls, const int32_t *segid, const int32_t *votes, int iw, int ih, double tanAOV) {
  const int n = *(const int32_t *)ls;
  std::vector<rd_tail_sample> sm((size_t)(n + 1) * RD_TAIL_NSAMPLE);
  rd_tail_gather_host(ls, segid, votes, iw, ih, sm.data());
  return rd_tail_compact(ls, sm.data(), iw, ih, tanAOV, NULL);
}
