// ora_tail.cpp - CPU oracle, host tail: restatement of executeCPUTask and its helpers
// (oclrect.c:385-1226, vec234.h, egbuf.h, helper.c:124-267).  TEST INFRASTRUCTURE ONLY (see rd_oracle.h).
// PINNED: checked bit for bit (order and every field) against the reference's own executeCPUTask compiled from
// /root/reference (oracle/_ref/librd_ref_tail.so, tests/test_ref_tail.py) and against reference-generated fixtures.
// Everything is IEEE double in the order the reference writes it.  CANONICAL (Q20): the qsort calls are
// replaced by stable sorts (glibc's qsort is a stable merge sort for these sizes).
#include <vector>
#include <algorithm>
#include "ora_internal.h"

namespace ora {

struct V2 { double a[2]; };
struct V3 { double a[3]; };
struct V4 { double a[4]; };
struct Seg { V2 e0, e1; };                      // ls_t, oclrect.c:385-387

static inline V2 v2(double x, double y) { V2 v; v.a[0] = x; v.a[1] = y; return v; }
static inline V3 v3(double x, double y, double z) { V3 v; v.a[0] = x; v.a[1] = y; v.a[2] = z; return v; }
static inline V4 v4(double x, double y, double z, double w) { V4 v; v.a[0] = x; v.a[1] = y; v.a[2] = z; v.a[3] = w; return v; }

// vec234.h : sums start at 0 and accumulate in index order
#define VEC_OPS(T, N)                                                                                       \
  static inline T plus(T a, T b) { T v; for (int i = 0; i < N; i++) v.a[i] = a.a[i] + b.a[i]; return v; }    \
  static inline T minus(T a, T b) { T v; for (int i = 0; i < N; i++) v.a[i] = a.a[i] - b.a[i]; return v; }   \
  static inline double vdot(T a, T b) { double s = 0; for (int i = 0; i < N; i++) s += a.a[i] * b.a[i]; return s; } \
  static inline T scale(T a, double d) { T v; for (int i = 0; i < N; i++) v.a[i] = a.a[i] * d; return v; }   \
  static inline double lengthSqu(T v) { double s = 0; for (int i = 0; i < N; i++) s += v.a[i] * v.a[i]; return s; } \
  static inline T normalize(T v) { double s = 0; for (int i = 0; i < N; i++) s += v.a[i] * v.a[i]; return scale(v, 1.0 / (sqrt(s) + 1e-20)); } \
  static inline double distanceSqu(T v, T w) { return lengthSqu(minus(v, w)); }                              \
  static inline double distance(T v, T w) { return sqrt(distanceSqu(v, w)); }
VEC_OPS(V2, 2)
VEC_OPS(V3, 3)
VEC_OPS(V4, 4)
static inline V2 midpoint(V2 p0, V2 p1) { return scale(plus(p0, p1), 0.5); }

static inline double squ(double x) { return x * x; }
static inline float lsSquLen(const Seg &s) { return (float)distanceSqu(s.e0, s.e1); }   // oclrect.c:390 (float return: Q20)

static inline V3 cross3(V3 v, V3 w) {
  return v3(v.a[1] * w.a[2] - v.a[2] * w.a[1], v.a[2] * w.a[0] - v.a[0] * w.a[2], v.a[0] * w.a[1] - v.a[1] * w.a[0]);
}

// oclrect.c:400-416
static inline V2 closestPoint2(V2 v, V2 w, V2 p) {
  double l2 = distanceSqu(v, w);
  if (l2 == 0.0) return v2(v.a[0], v.a[1]);
  double t = ((p.a[0] - v.a[0]) * (w.a[0] - v.a[0]) + (p.a[1] - v.a[1]) * (w.a[1] - v.a[1])) / l2;
  return v2(v.a[0] + t * (w.a[0] - v.a[0]), v.a[1] + t * (w.a[1] - v.a[1]));
}

static inline V2 closestPointLS2(V2 v, V2 w, V2 p) {
  double l2 = distanceSqu(v, w);
  if (l2 == 0.0) return v2(v.a[0], v.a[1]);
  double t = ((p.a[0] - v.a[0]) * (w.a[0] - v.a[0]) + (p.a[1] - v.a[1]) * (w.a[1] - v.a[1])) / l2;
  if (t < 0) return v2(v.a[0], v.a[1]);
  else if (t > 1.0) return v2(w.a[0], w.a[1]);
  return v2(v.a[0] + t * (w.a[0] - v.a[0]), v.a[1] + t * (w.a[1] - v.a[1]));
}

// oclrect.c:418-425
static inline V2 intersection2(Seg u, Seg v) {
  double d = (v.e1.a[0] - v.e0.a[0]) * (u.e1.a[1] - u.e0.a[1]) - (v.e1.a[1] - v.e0.a[1]) * (u.e1.a[0] - u.e0.a[0]);
  if (fabs(d) < 1e-4) return v2(NAN, NAN);
  double n = (v.e0.a[1] - u.e0.a[1]) * (u.e1.a[0] - u.e0.a[0]) - (v.e0.a[0] - u.e0.a[0]) * (u.e1.a[1] - u.e0.a[1]);
  double q = n / d;
  return v2(v.e0.a[0] + q * (v.e1.a[0] - v.e0.a[0]), v.e0.a[1] + q * (v.e1.a[1] - v.e0.a[1]));
}

// ---------------- pose estimator, oclrect.c:427-634 ----------------
#define POSE_EPS (1e-6)
struct PoseArg { const V3 *points; int mode; };

// oclrect.c:441-477
static double value(V4 v, const PoseArg *arg) {
  const V3 *points = arg->points;
  const int mode = arg->mode;
  V3 q[4];
  for (int i = 0; i < 4; i++) q[i] = scale(points[i], v.a[i]);
  double score = 0;
  double l01 = distanceSqu(q[0], q[1]);
  double l12 = distanceSqu(q[1], q[2]);
  double l23 = distanceSqu(q[2], q[3]);
  double l03 = distanceSqu(q[0], q[3]);
  double l02 = distanceSqu(q[0], q[2]);
  double l13 = distanceSqu(q[1], q[3]);
  double comp = 1.0;
  score += squ((mode ? l23 : l03) - 1);
  score += squ((mode ? l01 : l12) - 1);
  comp = 1.0 / (mode ? l12 : l01);
  score += lengthSqu(plus(minus(mode ? q[0] : q[2], q[1]), minus(mode ? q[2] : q[0], q[3])));
  score += comp * lengthSqu(plus(minus(q[1], mode ? q[2] : q[0]), minus(q[3], mode ? q[0] : q[2])));
  score += squ(l01 + l12 - l02);
  score += squ(l03 + l23 - l02);
  score += squ(l01 + l03 - l13);
  score += squ(l12 + l23 - l13);
  V3 n013 = cross3(minus(q[1], q[0]), minus(q[3], q[0]));
  score += comp * squ(vdot(n013, q[2]) - vdot(n013, q[0])) / vdot(n013, n013);
  V3 n102 = cross3(minus(q[0], q[1]), minus(q[2], q[1]));
  score += comp * squ(vdot(n102, q[3]) - vdot(n102, q[1])) / vdot(n102, n102);
  return score;
}

// oclrect.c:479-490
static inline V3 gradient(V4 v, V4 dir, const PoseArg *arg) {
  double h = POSE_EPS;
  double f0 = value(v, arg);
  double fp = value(plus(v, scale(dir, h)), arg);
  double fm = value(plus(v, scale(dir, -h)), arg);
  V3 ret;
  ret.a[0] = f0;
  ret.a[1] = (fp - fm) * (1.0 / (2 * h));
  ret.a[2] = (fp + fm - 2 * f0) * (1.0 / (h * h));
  return ret;
}

// oclrect.c:492-512
static inline void gradient2(V4 v, const PoseArg *arg, V4 &a, V4 &a2) {
  V4 d;
  double fx = value(v, arg);
  for (int i = 0; i < 4; i++) {
    for (int j = 0; j < 4; j++) { d.a[j] = 0; if (j == i) d.a[j] = POSE_EPS; }
    double fxmh = value(minus(v, d), arg);
    double fxph = value(plus(v, d), arg);
    a.a[i] = (fxph - fxmh) / (2 * POSE_EPS);
    a2.a[i] = (fxmh - 2 * fx + fxph) / (POSE_EPS * POSE_EPS);
  }
}

// oclrect.c:514-536
static inline V4 lineSearch(V4 iv, V4 dir, int nIter2, const PoseArg *arg) {
  dir = normalize(dir);
  V3 gd;
  double sc = 1.0;   // initScale
  for (int i = 0; i < nIter2; i++) {
    gd = gradient(iv, dir, arg);
    double ep = gd.a[0];
    if (gd.a[2] * gd.a[2] < 1e-10) gd.a[2] = 1;
    double delta = fabs(gd.a[1] / gd.a[2]);
    if (delta < 1e-10) return iv;
    V4 v = plus(iv, scale(dir, delta * sc));
    double e1 = value(v, arg);
    if (ep < e1) { sc *= 0.5; continue; }
    iv = v;
  }
  return iv;
}

// oclrect.c:538-555
static inline V4 inversedot(V4 m, V4 r) {
  V4 a;
  int isAllPositive = 1;
  for (int i = 0; i < 4; i++) if (m.a[i] <= 0) isAllPositive = 0;
  if (isAllPositive) {
    for (int i = 0; i < 4; i++) { a.a[i] = 1.0 / m.a[i]; a.a[i] *= r.a[i]; }
  } else return r;
  return a;
}

// oclrect.c:557-588
static V4 cgexecute(V4 iv, int loopCnt, int nIter2, const PoseArg *arg) {
  int i = 0, k = 0;
  V4 x = iv, g, m;
  gradient2(x, arg, g, m);
  V4 r = scale(g, -1);
  V4 s = inversedot(m, r), d = s;
  double deltanew = vdot(r, d);
  while (i < loopCnt) {
    x = lineSearch(x, d, nIter2, arg);
    gradient2(x, arg, g, m);
    r = scale(g, -1);
    double deltaold = deltanew;
    double deltamid = vdot(r, s);
    s = inversedot(m, r);
    deltanew = vdot(r, s);
    double beta = (deltanew - deltamid) / deltaold;
    if (k == 10 || beta <= 0 || deltaold == 0) { d = s; k = 0; }
    else d = plus(s, scale(d, beta));
    k++; i++;
  }
  return x;
}

// oclrect.c:590-634
static void poseEstimation(const Seg *als, V2 gvv, int iw, int ih, double tanAOV, ora_rect_t *ret) {
  V3 p[4];
  int tl = 0;
  double mn = 1e+100;
  for (int i = 0; i < 4; i++) {
    V2 v = normalize(minus(als[i].e1, als[i].e0));
    v = v2(-v.a[1], v.a[0]);
    if (vdot(minus(als[i].e0, gvv), v) < 0) v = scale(v, -1);
    if (v.a[1] < mn) { mn = v.a[1]; tl = i; }
  }
  for (int i = 0; i < 4; i++)
    p[i] = normalize(v3((als[(i + tl) & 3].e0.a[0] - (iw / 2)), (-(als[(i + tl) & 3].e0.a[1] - ih / 2)), iw / 2 / tanAOV));

  double d01 = 1.0 / distance(p[0], p[1]);
  double d23 = 1.0 / distance(p[2], p[3]);
  PoseArg arg0 = {p, 1};
  V4 x0 = cgexecute(v4(d01, d01, d23, d23), 12, 10, &arg0);
  double val0 = value(x0, &arg0);

  double d12 = 1.0 / distance(p[1], p[2]);
  double d03 = 1.0 / distance(p[0], p[3]);
  PoseArg arg1 = {p, 0};
  V4 x1 = cgexecute(v4(d03, d12, d12, d03), 12, 10, &arg1);
  double val1 = value(x1, &arg1);

  ret->value = val0 < val1 ? val0 : val1;
  V4 x = val0 < val1 ? x0 : x1;
  if (x.a[0] < 0) x = scale(x, -1);
  for (int i = 0; i < 4; i++) {
    V3 c = scale(p[i], x.a[i]);
    ret->c3[i][0] = c.a[0]; ret->c3[i][1] = c.a[1]; ret->c3[i][2] = c.a[2];
    ret->c2[i][0] = als[(i + tl) & 3].e0.a[0];
    ret->c2[i][1] = als[(i + tl) & 3].e0.a[1];
  }
}

// oclrect.c:636-656
static int looksLikeAScreen(const ora_rect_t &r) {
  if (r.value > 0.05) return 0;
  if (r.c3[0][2] < 0 || r.c3[1][2] < 0 || r.c3[2][2] < 0 || r.c3[3][2] < 0) return 0;
  V3 c3[4]; V2 c2[4];
  for (int i = 0; i < 4; i++) { c3[i] = v3(r.c3[i][0], r.c3[i][1], r.c3[i][2]); c2[i] = v2(r.c2[i][0], r.c2[i][1]); }
  double asp = distance(c3[0], c3[1]) / distance(c3[1], c3[2]);
  if (asp < 1.0 / 12 || 12 < asp) return 0;
  double maxs = 0, mins = 1e+100;
  for (int i = 0; i < 4; i++) {
    double s0 = distanceSqu(c2[(i + 2) % 4], closestPointLS2(c2[i], c2[(i + 1) % 4], c2[(i + 2) % 4]));
    double s1 = distanceSqu(c2[(i + 3) % 4], closestPointLS2(c2[i], c2[(i + 1) % 4], c2[(i + 3) % 4]));
    maxs = fmax(maxs, fmax(s0, s1));
    mins = fmin(mins, fmax(s0, s1));
  }
  if (maxs / mins > 100) return 0;
  return 1;
}

// ---------------- quick hull, oclrect.c:658-734 ----------------
// The reference stores points by value in EGBufs and identifies "the farthest point" by address within the
// current subset; the same is done here with indices into the subset vector.
static void findHull2(std::vector<V2> &hull, const std::vector<V2> &s, V2 vLeft, V2 vRight) {
  int far = -1;
  double d = 0;
  for (int i = 0; i < (int)s.size(); i++) {
    double e = distanceSqu(closestPoint2(vLeft, vRight, s[i]), s[i]);
    if (far < 0 || e > d) { far = i; d = e; }
  }
  if (d < 0.01 || far < 0) return;
  const V2 pf = s[far];
  V2 vTopRight = v2(pf.a[1] - vRight.a[1], vRight.a[0] - pf.a[0]);
  V2 vTopLeft = v2(vLeft.a[1] - pf.a[1], pf.a[0] - vLeft.a[0]);
  std::vector<V2> sTopRight, sTopLeft;
  for (int i = 0; i < (int)s.size(); i++) {
    if (i == far) continue;
    if (vdot(minus(s[i], pf), vTopRight) > 0) sTopRight.push_back(s[i]);
    if (vdot(minus(s[i], pf), vTopLeft) > 0) sTopLeft.push_back(s[i]);
  }
  findHull2(hull, sTopRight, pf, vRight);
  hull.push_back(pf);
  findHull2(hull, sTopLeft, vLeft, pf);
}

static std::vector<V2> quickHull2(const std::vector<V2> &s) {
  std::vector<V2> hull;
  if (s.empty()) return hull;
  V2 vRight = s[0], vLeft = s[0];
  for (size_t i = 0; i < s.size(); i++) {
    if (s[i].a[0] > vRight.a[0]) vRight = s[i];
    if (s[i].a[0] < vLeft.a[0]) vLeft = s[i];
  }
  V2 vTop = v2(vLeft.a[1] - vRight.a[1], vRight.a[0] - vLeft.a[0]);
  std::vector<V2> sTop, sBot;
  for (size_t i = 0; i < s.size(); i++) {
    const V2 &p = s[i];
    if (p.a[0] == vLeft.a[0] && p.a[1] == vLeft.a[1]) continue;
    if (p.a[0] == vRight.a[0] && p.a[1] == vRight.a[1]) continue;
    if (vdot(minus(p, vLeft), vTop) > 0) sTop.push_back(p); else sBot.push_back(p);
  }
  hull.push_back(vRight);
  findHull2(hull, sTop, vLeft, vRight);
  hull.push_back(vLeft);
  findHull2(hull, sBot, vRight, vLeft);
  return hull;
}

// ---------------- Cohen-Sutherland, oclrect.c:744-802 ----------------
static inline int outCode(double x, double y, double xmin, double ymin, double xmax, double ymax) {
  int code = 0;
  if (x < xmin) code |= 1;
  if (x > xmax) code |= 2;
  if (y < ymin) code |= 4;
  if (y > ymax) code |= 8;
  return code;
}

static V4 clipLineWithRect(double x0, double y0, double x1, double y1, double xmin, double ymin, double xmax, double ymax) {
  int outcode0 = outCode(x0, y0, xmin, ymin, xmax, ymax);
  int outcode1 = outCode(x1, y1, xmin, ymin, xmax, ymax);
  int accept = 0;
  for (;;) {
    if ((outcode0 | outcode1) == 0) { accept = 1; break; }
    else if ((outcode0 & outcode1) != 0) break;
    else {
      double x = 0, y = 0;
      int outcodeOut = outcode0 != 0 ? outcode0 : outcode1;
      if ((outcodeOut & 8) != 0) { x = x0 + (x1 - x0) * (ymax - y0) / (y1 - y0); y = ymax; }
      else if ((outcodeOut & 4) != 0) { x = x0 + (x1 - x0) * (ymin - y0) / (y1 - y0); y = ymin; }
      else if ((outcodeOut & 2) != 0) { y = y0 + (y1 - y0) * (xmax - x0) / (x1 - x0); x = xmax; }
      else if ((outcodeOut & 1) != 0) { y = y0 + (y1 - y0) * (xmin - x0) / (x1 - x0); x = xmin; }
      if (outcodeOut == outcode0) { x0 = x; y0 = y; outcode0 = outCode(x0, y0, xmin, ymin, xmax, ymax); }
      else { x1 = x; y1 = y; outcode1 = outCode(x1, y1, xmin, ymin, xmax, ymax); }
    }
  }
  if (accept) return v4(x0, y0, x1, y1);
  return v4(NAN, NAN, NAN, NAN);
}

// ---------------- list helpers, oclrect.c:806-1045 ----------------
static void sortByLength(std::vector<Seg> &als) {
  std::stable_sort(als.begin(), als.end(), [](const Seg &a, const Seg &b) { return lsSquLen(a) < lsSquLen(b); });
}

static double segAngle(const Seg &s, V2 gvv) {      // oclrect.c:829-834
  V2 v = minus(s.e0, s.e1);
  v = v2(v.a[1], -v.a[0]);
  if (vdot(v, minus(s.e0, gvv)) < 0) v = scale(v, -1);
  return atan2(v.a[0], v.a[1]);
}

static void sortByAngle(std::vector<Seg> &als, V2 gvv) {
  std::stable_sort(als.begin(), als.end(), [gvv](const Seg &a, const Seg &b) { return segAngle(a, gvv) < segAngle(b, gvv); });
}

static V2 gv(const std::vector<Seg> &als) {        // oclrect.c:864-877
  V2 g = v2(0, 0);
  double lenSum = 0;
  for (size_t i = 0; i < als.size(); i++) {
    double len = distance(als[i].e0, als[i].e1);
    g = plus(g, scale(plus(als[i].e0, als[i].e1), len));
    lenSum += len;
  }
  return scale(g, 0.5 / lenSum);
}

static double sumLength(const std::vector<Seg> *als) {   // oclrect.c:879-884
  if (als == NULL) return 0;
  double ret = 0;
  for (size_t i = 0; i < als->size(); i++) ret += sqrt((double)lsSquLen((*als)[i]));   // C: sqrt(double) of the float-rounded square
  return ret;
}

static int closeToTriangle(const std::vector<Seg> &als, double ratio) {   // oclrect.c:886-895
  const int n = (int)als.size();
  for (int i = 0; i < n; i++) {
    const Seg &ls0 = als[i], &ls1 = als[(i + 1) % n];
    double d0 = distanceSqu(ls0.e1, closestPoint2(ls0.e0, ls1.e1, ls0.e1));
    double d1 = distanceSqu(ls0.e0, ls1.e1);
    if (d0 / d1 < ratio) return 1;
  }
  return 0;
}

static int isConvex(const std::vector<Seg> &als) {        // oclrect.c:897-922
  const Seg &a0 = als[0], &a1 = als[1];
  double px0 = a0.e1.a[0] - a0.e0.a[0], py0 = a0.e1.a[1] - a0.e0.a[1];
  double px1 = a1.e1.a[0] - a1.e0.a[0], py1 = a1.e1.a[1] - a1.e0.a[1];
  int sign = px0 * py1 - py0 * px1 > 0;
  const int as = (int)als.size();
  for (int i = 1; i < as; i++) {
    const Seg &l0 = als[i], &l1 = als[(i + 1) % as];
    double qx0 = l0.e1.a[0] - l0.e0.a[0], qy0 = l0.e1.a[1] - l0.e0.a[1];
    double qx1 = l1.e1.a[0] - l1.e0.a[0], qy1 = l1.e1.a[1] - l1.e0.a[1];
    if (sign != (qx0 * qy1 - qy0 * qx1 > 0)) return 0;
  }
  return 1;
}

static void removeShortLS(std::vector<Seg> &als, float ratio) {   // oclrect.c:926-943
  if (als.size() <= 4) return;
  sortByLength(als);
  float longestSquLen = lsSquLen(als.back());
  for (;;) {
    if (als.size() <= 4) break;
    float shortestSquLen = lsSquLen(als[0]);
    if (shortestSquLen / longestSquLen > ratio * ratio) break;
    als.erase(als.begin());
  }
}

static std::vector<Seg> pickExternalLS(std::vector<Seg> als) {     // oclrect.c:945-992
  std::vector<V2> plist;
  for (size_t i = 0; i < als.size(); i++) { plist.push_back(als[i].e0); plist.push_back(als[i].e1); }
  std::vector<V2> q = quickHull2(plist);
  std::vector<Seg> als2;
  const double DTHRE0 = 1, ATHRE1 = 0.95, DTHRE1 = 0.01;
  const int qs = (int)q.size();
  for (int i = 0; i < qs; i++) {
    V2 q0 = q[i], q1 = q[(i + 1) % qs];
    V2 m = midpoint(q0, q1), nq01 = normalize(minus(q0, q1));
    int lastAdded = -1;
    sortByLength(als);
    for (int j = (int)als.size() - 1; j >= 0; j--) {
      Seg e = als[j];
      if (distanceSqu(m, closestPointLS2(e.e0, e.e1, m)) < DTHRE0) { als2.push_back(e); lastAdded = j; break; }
      if (fabs(vdot(nq01, normalize(minus(e.e0, e.e1)))) > ATHRE1 &&
          distanceSqu(m, closestPointLS2(e.e0, e.e1, m)) / distanceSqu(q0, q1) < DTHRE1) { als2.push_back(e); lastAdded = j; break; }
    }
    if (lastAdded != -1) als.erase(als.begin() + lastAdded);
  }
  return als2;
}

static std::vector<Seg> pickLongestLS(std::vector<Seg> als, int n) {   // oclrect.c:994-1009
  if ((int)als.size() <= n) return als;
  sortByLength(als);
  std::vector<Seg> ret;
  for (int j = (int)als.size() - 1; j >= 0; j--) {
    ret.push_back(als[j]);
    if ((int)ret.size() == n) break;
  }
  return ret;
}

// oclrect.c:1011-1045 ; returns false for the reference's NULL
static bool findCorners(std::vector<Seg> &als) {
  const int n = (int)als.size();
  std::vector<V2> c(n);
  for (int i = 0; i < n; i++) {
    c[i] = intersection2(als[i], als[(i + 1) % n]);
    if (isnan(c[i].a[0])) return false;
  }
  std::vector<Seg> ret(n);
  for (int i = 0; i < n; i++) { ret[i].e0 = c[i]; ret[i].e1 = c[(i + 1) % n]; }
  als.swap(ret);
  return true;
}

// the common part of the two candidate loops, oclrect.c:1134-1160 and 1190-1216
static void tryCandidate(std::vector<Seg> als, uint32_t baseStatus, int iw, int ih, double tanAOV, std::vector<ora_rect_t> &out) {
  removeShortLS(als, 0.05f);
  als = pickExternalLS(als);
  double len0 = sumLength(&als);
  als = pickLongestLS(als, 4);
  sortByAngle(als, gv(als));
  bool ok = findCorners(als);
  double len1 = ok ? sumLength(&als) : 0;
  if (!ok || closeToTriangle(als, 0.001) || als.size() < 4 || len1 / len0 > 2 || !isConvex(als)) return;
  ora_rect_t rect;
  memset(&rect, 0, sizeof(rect));
  poseEstimation(als.data(), gv(als), iw, ih, tanAOV, &rect);
  rect.status = baseStatus;
  if (looksLikeAScreen(rect)) rect.status |= 1;
  out.push_back(rect);
}

// ArrayMap bucket of a key, helper.c:129-131
static inline int bucketOf(uint64_t key) { return (int)((key ^ (key >> 10) ^ (key >> 20) ^ (key >> 30)) & 1023); }

}  // namespace ora

using namespace ora;

extern "C" {

// executeCPUTask, oclrect.c:1049-1226
ora_rect_t *ora_tail(const ora_ls_t *ls, const int32_t *segidMap, const int32_t *votes, int iw, int ih, double tanAOV) {
  const int n = *(const int32_t *)ls;
  std::vector<ora_rect_t> ret;

  // (i) every live segment samples the segid map at 3 points x 5 normal offsets (oclrect.c:1066-1098).
  // The reference keeps an ArrayMap<segid, list<lsid>>; iteration order = bucket order, then insertion order.
  struct Entry { int segid; std::vector<int> lsids; };
  std::vector<std::vector<Entry>> buckets(1024);
  for (int i = 1; i <= n; i++) {
    if (ls[i].polyid == 0) continue;
    double x0 = rint(ls[i].x0), y0 = rint(ls[i].y0), x1 = rint(ls[i].x1), y1 = rint(ls[i].y1);
    const int N = 3, DIST = 2;
    V2 d = normalize(minus(v2(x1, y1), v2(x0, y0)));
    V2 vd = v2(-d.a[1], d.a[0]);
    for (int j = 0; j < N; j++)
      for (int dist = -DIST; dist <= DIST; dist++) {
        V2 p = plus(v2(x0, y0), scale(minus(v2(x1, y1), v2(x0, y0)), (j + 0.5) / N));
        V2 c = plus(p, scale(vd, dist));
        int x = (int)(c.a[0] + 0.5), y = (int)(c.a[1] + 0.5);
        if (x < 0 || x >= iw || y < 0 || y >= ih) continue;
        int segid = segidMap[x + y * iw];
        if (segid <= 0) continue;
        std::vector<Entry> &b = buckets[bucketOf((uint64_t)segid)];
        Entry *e = NULL;
        for (size_t k = 0; k < b.size(); k++) if (b[k].segid == segid) { e = &b[k]; break; }
        if (!e) { b.push_back(Entry()); e = &b.back(); e->segid = segid; }
        if (std::find(e->lsids.begin(), e->lsids.end(), i) == e->lsids.end()) e->lsids.push_back(i);
      }
  }

  // (ii) per region with >= 4 segments (oclrect.c:1103-1161)
  const unsigned nentry = (unsigned)(iw * ih * 4 / 5);
  for (int bk = 0; bk < 1024; bk++)
    for (size_t ei = 0; ei < buckets[bk].size(); ei++) {
      const Entry &e = buckets[bk][ei];
      if (e.lsids.size() < 4) continue;
      std::vector<Seg> als;
      for (size_t j = 0; j < e.lsids.size(); j++) {
        const int lsid = e.lsids[j];
        const int hash = (int)((((uint32_t)lsid * (uint32_t)e.segid) & 0x7fffffff) % nentry);
        const int32_t *v = &votes[(size_t)hash * 5];
        if (v[0] != lsid) {
          if (v[0] != 0) {
            Seg s = {v2(ls[lsid].x0, ls[lsid].y0), v2(ls[lsid].x1, ls[lsid].y1)};
            als.push_back(s);
          }
          continue;
        }
        V4 cl = clipLineWithRect(ls[lsid].x0, ls[lsid].y0, ls[lsid].x1, ls[lsid].y1, iw - v[1], ih - v[3], v[2], v[4]);
        if (isnan(cl.a[0])) continue;
        Seg s = {v2(cl.a[0], cl.a[1]), v2(cl.a[2], cl.a[3])};
        als.push_back(s);
      }
      tryCandidate(als, 0, iw, ih, tanAOV, ret);
    }

  // (iii) per polyline chain (oclrect.c:1175-1217)
  for (int i = 1; i <= n; i++) {
    if (ls[i].polyid == 0) continue;
    if (ls[i].leftPtr > 0) continue;
    std::vector<Seg> als;
    for (int j = i; j > 0; j = ls[j].rightPtr) {
      const double LSTHRE = 32;
      V2 e0 = v2(ls[j].x0, ls[j].y0), e1 = v2(ls[j].x1, ls[j].y1);
      if (distanceSqu(e0, e1) > LSTHRE * LSTHRE) { Seg s = {e0, e1}; als.push_back(s); }
    }
    tryCandidate(als, 2, iw, ih, tanAOV, ret);
  }

  ora_rect_t *out = (ora_rect_t *)calloc(ret.size() + 1, sizeof(ora_rect_t));
  for (size_t i = 0; i < ret.size(); i++) out[i + 1] = ret[i];
  out[0].nItems = (int)ret.size() + 1;
  return out;
}

void ora_clip_line(double x0, double y0, double x1, double y1, double xmin, double ymin, double xmax, double ymax, double out[4]) {
  V4 r = clipLineWithRect(x0, y0, x1, y1, xmin, ymin, xmax, ymax);
  for (int i = 0; i < 4; i++) out[i] = r.a[i];
}

void ora_intersection2(const double u[4], const double v[4], double out[2]) {
  Seg su = {v2(u[0], u[1]), v2(u[2], u[3])}, sv = {v2(v[0], v[1]), v2(v[2], v[3])};
  V2 r = intersection2(su, sv);
  out[0] = r.a[0]; out[1] = r.a[1];
}

void ora_pose(const double corners[4][2], int iw, int ih, double tanAOV, ora_rect_t *out) {
  Seg als[4];
  for (int i = 0; i < 4; i++) {
    als[i].e0 = v2(corners[i][0], corners[i][1]);
    als[i].e1 = v2(corners[(i + 1) % 4][0], corners[(i + 1) % 4][1]);
  }
  std::vector<Seg> v(als, als + 4);
  memset(out, 0, sizeof(*out));
  poseEstimation(als, gv(v), iw, ih, tanAOV, out);
  out->status = looksLikeAScreen(*out) ? 1 : 0;
}

}  // extern "C"
