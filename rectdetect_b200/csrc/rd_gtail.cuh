// rd_gtail.cuh - executeCPUTask (oclrect.c:1049-1226 and its helpers :385-1045) restated for the GPU: quad assembly and pose
// estimation from the line segments, the region map and the (segment x region) vote table, all in IEEE double in the reference's
// order of operations (no FMA contraction: -fmad=false on the device, -ffp-contract=off on the host).
//
// The building blocks below are shared by the CUDA kernels (rd_gtail.cu) and by a host replay (tests/emu_gtail.cpp, where the 32
// lanes of a warp run as fibers and the warp collectives are rendezvous points), so the logic is checked against the reference's
// own executeCPUTask (oracle/_ref) and rd_tail.cpp without a GPU.  W is the warp interface:
//     int lane;  unsigned ballot(bool);  int shfl(int v, int src);  double shfl(double v, int src);  void sync();
// Every collective is called by all 32 lanes (warp-uniform control flow around them).
#ifndef RD_GTAIL_CUH
#define RD_GTAIL_CUH
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define GT_FN __device__ __forceinline__
#define GT_FN_NOINLINE __device__ __noinline__
#else
#define GT_FN static inline
#define GT_FN_NOINLINE static
#endif

struct GtP2 { double x, y; };
struct GtP3 { double x, y, z; };
struct GtP4 { double v[4]; };
struct GtEdge { GtP2 a, b; };                                   // ls_t, oclrect.c:385-387
struct GtLS { float x0, y0, x1, y1; int32_t startIndex, endIndex, leftPtr, rightPtr, startCount, endCount, maxDist, polyid, npix, level; };   // linesegment_t

// ---- vec234.h in the reference's order of operations (rd_tail.cpp holds the same restatement for the host)
GT_FN GtP2 gt_add(GtP2 p, GtP2 q) { return {p.x + q.x, p.y + q.y}; }
GT_FN GtP2 gt_sub(GtP2 p, GtP2 q) { return {p.x - q.x, p.y - q.y}; }
GT_FN GtP2 gt_mul(GtP2 p, double s) { return {p.x * s, p.y * s}; }
GT_FN double gt_dot(GtP2 p, GtP2 q) { double s = 0; s += p.x * q.x; s += p.y * q.y; return s; }
GT_FN double gt_dist2(GtP2 p, GtP2 q) { const GtP2 d = gt_sub(p, q); return gt_dot(d, d); }
GT_FN GtP2 gt_unit(GtP2 p) { return gt_mul(p, 1.0 / (sqrt(gt_dot(p, p)) + 1e-20)); }
GT_FN GtP3 gt_add(GtP3 p, GtP3 q) { return {p.x + q.x, p.y + q.y, p.z + q.z}; }
GT_FN GtP3 gt_sub(GtP3 p, GtP3 q) { return {p.x - q.x, p.y - q.y, p.z - q.z}; }
GT_FN GtP3 gt_mul(GtP3 p, double s) { return {p.x * s, p.y * s, p.z * s}; }
GT_FN double gt_dot(GtP3 p, GtP3 q) { double s = 0; s += p.x * q.x; s += p.y * q.y; s += p.z * q.z; return s; }
GT_FN double gt_dist2(GtP3 p, GtP3 q) { const GtP3 d = gt_sub(p, q); return gt_dot(d, d); }
GT_FN GtP3 gt_unit(GtP3 p) { return gt_mul(p, 1.0 / (sqrt(gt_dot(p, p)) + 1e-20)); }
GT_FN GtP3 gt_cross(GtP3 v, GtP3 w) { return {v.y * w.z - v.z * w.y, v.z * w.x - v.x * w.z, v.x * w.y - v.y * w.x}; }
GT_FN GtP4 gt_add4(GtP4 p, GtP4 q) { GtP4 r; for (int i = 0; i < 4; i++) r.v[i] = p.v[i] + q.v[i]; return r; }
GT_FN GtP4 gt_sub4(GtP4 p, GtP4 q) { GtP4 r; for (int i = 0; i < 4; i++) r.v[i] = p.v[i] - q.v[i]; return r; }
GT_FN GtP4 gt_mul4(GtP4 p, double s) { GtP4 r; for (int i = 0; i < 4; i++) r.v[i] = p.v[i] * s; return r; }
GT_FN double gt_dot4(GtP4 p, GtP4 q) { double s = 0; for (int i = 0; i < 4; i++) s += p.v[i] * q.v[i]; return s; }
GT_FN GtP4 gt_unit4(GtP4 p) { return gt_mul4(p, 1.0 / (sqrt(gt_dot4(p, p)) + 1e-20)); }
GT_FN double gt_sq(double x) { return x * x; }
GT_FN float gt_len2f(const GtEdge &e) { return (float)gt_dist2(e.a, e.b); }            // lsSquLen returns float (oclrect.c:390)

// foot of the perpendicular from p on the line through v, w (oclrect.c:400-406), and on the segment (:408-416)
GT_FN GtP2 gt_footOnLine(GtP2 v, GtP2 w, GtP2 p) {
  const double l2 = gt_dist2(v, w);
  if (l2 == 0.0) return v;
  const double t = ((p.x - v.x) * (w.x - v.x) + (p.y - v.y) * (w.y - v.y)) / l2;
  return {v.x + t * (w.x - v.x), v.y + t * (w.y - v.y)};
}
GT_FN GtP2 gt_footOnSegment(GtP2 v, GtP2 w, GtP2 p) {
  const double l2 = gt_dist2(v, w);
  if (l2 == 0.0) return v;
  const double t = ((p.x - v.x) * (w.x - v.x) + (p.y - v.y) * (w.y - v.y)) / l2;
  if (t < 0) return v;
  else if (t > 1.0) return w;
  return {v.x + t * (w.x - v.x), v.y + t * (w.y - v.y)};
}
GT_FN GtP2 gt_lineIntersection(const GtEdge &u, const GtEdge &v) {                       // oclrect.c:418-425
  const double d = (v.b.x - v.a.x) * (u.b.y - u.a.y) - (v.b.y - v.a.y) * (u.b.x - u.a.x);
  if (fabs(d) < 1e-4) return {NAN, NAN};
  const double n = (v.a.y - u.a.y) * (u.b.x - u.a.x) - (v.a.x - u.a.x) * (u.b.y - u.a.y);
  const double q = n / d;
  return {v.a.x + q * (v.b.x - v.a.x), v.a.y + q * (v.b.y - v.a.y)};
}

// ---- Cohen-Sutherland clip (oclrect.c:744-802)
GT_FN int gt_outcode(double x, double y, double xmin, double ymin, double xmax, double ymax) {
  int c = 0;
  if (x < xmin) c |= 1;
  if (x > xmax) c |= 2;
  if (y < ymin) c |= 4;
  if (y > ymax) c |= 8;
  return c;
}
GT_FN bool gt_clipToBox(double &x0, double &y0, double &x1, double &y1, double xmin, double ymin, double xmax, double ymax) {
  int c0 = gt_outcode(x0, y0, xmin, ymin, xmax, ymax), c1 = gt_outcode(x1, y1, xmin, ymin, xmax, ymax);
  for (;;) {
    if ((c0 | c1) == 0) return true;
    if ((c0 & c1) != 0) return false;
    double x = 0, y = 0;
    const int co = c0 != 0 ? c0 : c1;
    if (co & 8) { x = x0 + (x1 - x0) * (ymax - y0) / (y1 - y0); y = ymax; }
    else if (co & 4) { x = x0 + (x1 - x0) * (ymin - y0) / (y1 - y0); y = ymin; }
    else if (co & 2) { y = y0 + (y1 - y0) * (xmax - x0) / (x1 - x0); x = xmax; }
    else if (co & 1) { y = y0 + (y1 - y0) * (xmin - x0) / (x1 - x0); x = xmin; }
    if (co == c0) { x0 = x; y0 = y; c0 = gt_outcode(x0, y0, xmin, ymin, xmax, ymax); }
    else { x1 = x; y1 = y; c1 = gt_outcode(x1, y1, xmin, ymin, xmax, ymax); }
  }
}

// ---- the sampling step (oclrect.c:1066-1098): the k-th of the 15 sample points of segment e -> pixel index, or -1 outside the frame
GT_FN int gt_sample_pixel(const GtLS &e, int k, int iw, int ih) {
  const GtP2 s = {rint((double)e.x0), rint((double)e.y0)}, t = {rint((double)e.x1), rint((double)e.y1)};
  const GtP2 d = gt_unit(gt_sub(t, s)), nrm = {-d.y, d.x};
  const int j = k / 5, off = k % 5 - 2;
  const GtP2 p = gt_add(s, gt_mul(gt_sub(t, s), (j + 0.5) / 3));
  const GtP2 c = gt_add(p, gt_mul(nrm, off));
  const int x = (int)(c.x + 0.5), y = (int)(c.y + 0.5);
  if (x < 0 || x >= iw || y < 0 || y >= ih) return -1;
  return x + y * iw;
}
GT_FN int gt_vote_slot(int lsid, int segid, int nentry) { return (int)((((uint32_t)lsid * (uint32_t)segid) & 0x7fffffffu) % (uint32_t)nentry); }
GT_FN int gt_bucketOf(uint64_t key) { return (int)((key ^ (key >> 10) ^ (key >> 20) ^ (key >> 30)) & 1023); }   // helper.c:129-131

// the edge a (segment, region) pair contributes to the region's candidate (oclrect.c:1108-1127): 0 = none
GT_FN int gt_region_edge(const GtLS &l, int id, const int *v, int iw, int ih, GtEdge &out) {
  if (v[0] != id) {                                             // slot owned by another segment (hash collision): unclipped
    if (v[0] == 0) return 0;
    out = GtEdge{{(double)l.x0, (double)l.y0}, {(double)l.x1, (double)l.y1}};
    return 1;
  }
  double x0 = l.x0, y0 = l.y0, x1 = l.x1, y1 = l.y1;
  if (!gt_clipToBox(x0, y0, x1, y1, iw - v[1], ih - v[3], v[2], v[4])) return 0;
  out = GtEdge{{x0, y0}, {x1, y1}};
  return 1;
}

// ================================================================ candidate -> quadrilateral (oclrect.c:1129-1139, helpers :806-1045)
// Work storage of one candidate with up to m edges (all in global scratch, carved by gt_cand_scratch):
struct GtWork {
  GtEdge *e0, *e1, *kept;        // edges as built / sorted by length / kept by pickExternalLS     [m] each
  float *key;                    // squared lengths (float, as lsSquLen)                            [m]
  int *alive;                    // e1[j] not yet taken                                              [m]
  GtP2 *hull;                    // hull vertices                                                    [2m + 2]
  struct Frame { GtP2 left, right, pf; int far, stage; } *stack;                                  // [2m + 2]
};
GT_FN size_t gt_work_bytes(int m) {
  return (size_t)m * (3 * sizeof(GtEdge) + 8) + (size_t)(2 * m + 2) * (sizeof(GtP2) + sizeof(GtWork::Frame));
}
GT_FN GtWork gt_work_carve(unsigned char *p, int m) {
  GtWork w;
  w.e0 = (GtEdge *)p; p += (size_t)m * sizeof(GtEdge);
  w.e1 = (GtEdge *)p; p += (size_t)m * sizeof(GtEdge);
  w.kept = (GtEdge *)p; p += (size_t)m * sizeof(GtEdge);
  w.hull = (GtP2 *)p; p += (size_t)(2 * m + 2) * sizeof(GtP2);
  w.stack = (GtWork::Frame *)p; p += (size_t)(2 * m + 2) * sizeof(GtWork::Frame);
  w.key = (float *)p; p += (size_t)m * sizeof(float);
  w.alive = (int *)p;
  return w;
}

#ifndef GT_DBG3
#define GT_DBG3(...)
#endif
#ifndef GT_DBG2
#define GT_DBG2(...)
#endif
#ifndef GT_DBG
#define GT_DBG(k)
#endif
struct GtQuad { GtP2 c[4]; GtP2 centre; int valid; int status; };      // the four corners in angle order; centre = gv() of the cornered edges

// stable sort of src[0..n) by key ascending into dst (rank sort: rank = #{smaller key} + #{equal key, smaller index})
template <class W>
GT_FN void gt_sort_edges(W &w, GtEdge *dst, float *dkey, const GtEdge *src, int n) {
  for (int i = w.lane; i < n; i += 32) {
    const float ki = gt_len2f(src[i]);
    int r = 0;
    for (int j = 0; j < n; j++) { const float kj = gt_len2f(src[j]); r += (kj < ki || (kj == ki && j < i)) ? 1 : 0; }
    dst[r] = src[i];
    dkey[r] = ki;
  }
  w.sync();
}

// pts[i] of pickExternalLS: the end points of es in order a0 b0 a1 b1 ...
GT_FN GtP2 gt_pt(const GtEdge *es, int i) { return (i & 1) ? es[i >> 1].b : es[i >> 1].a; }

// is point i a member of the subset the frame at `depth` works on?  (top = 1: the upper chain, 0: the lower one)
GT_FN bool gt_hull_member(const GtEdge *es, int i, const GtWork::Frame *st, int depth, int top, GtP2 L, GtP2 R, GtP2 up) {
  const GtP2 p = gt_pt(es, i);
  if (p.x == L.x && p.y == L.y) return false;
  if (p.x == R.x && p.y == R.y) return false;
  const bool isTop = gt_dot(gt_sub(p, L), up) > 0;
  if (isTop != (top != 0)) return false;
  for (int k = 0; k < depth; k++) {
    const GtWork::Frame &f = st[k];
    if (i == f.far) return false;
    const GtP2 n = f.stage == 1 ? GtP2{f.pf.y - f.right.y, f.right.x - f.pf.x} : GtP2{f.left.y - f.pf.y, f.pf.x - f.left.x};
    if (!(gt_dot(gt_sub(p, f.pf), n) > 0)) return false;
  }
  return true;
}

// quick hull of the 2n end points (oclrect.c:658-734), recursion unrolled on an explicit stack; membership of a point in the
// subset of a call is re-derived from the predicates of the calls above it, so no point lists are stored.  Returns the hull size.
template <class W>
GT_FN int gt_convex_hull(W &w, const GtEdge *es, int n, GtWork &wk) {
  const int np = 2 * n;
  if (np == 0) return 0;
  GtP2 R = gt_pt(es, 0), L = R;
  for (int i = 0; i < np; i++) {                                  // serial scan order matters for ties: every lane does the same scan
    const GtP2 p = gt_pt(es, i);
    if (p.x > R.x) R = p;
    if (p.x < L.x) L = p;
  }
  const GtP2 up = {L.y - R.y, R.x - L.x};
  int hs = 0;
  for (int top = 1; top >= 0; top--) {
    if (w.lane == 0) wk.hull[hs] = top ? R : L;
    hs++;
    int depth = 0;
    if (w.lane == 0) { wk.stack[0].left = top ? L : R; wk.stack[0].right = top ? R : L; wk.stack[0].stage = 0; wk.stack[0].far = -1; }
    w.sync();
    while (depth >= 0) {
      GtWork::Frame &f = wk.stack[depth];
      const int stage = f.stage;                                  // uniform
      w.sync();                                                   // every lane has read the frame before lane 0 moves it on
      if (stage == 0) {
        // farthest member from the line (left, right): first arg-max in index order
        const GtP2 fl = f.left, fr = f.right;
        double bd = -1.0; int bi = 0x7fffffff;
        for (int i = w.lane; i < np; i += 32) {
          if (!gt_hull_member(es, i, wk.stack, depth, top, L, R, up)) continue;
          const GtP2 p = gt_pt(es, i);
          const double e = gt_dist2(gt_footOnLine(fl, fr, p), p);
          if (bi == 0x7fffffff || e > bd) { bd = e; bi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
          const double od = w.shfl(bd, w.lane ^ o);
          const int oi = w.shfl(bi, w.lane ^ o);
          if (oi != 0x7fffffff && (bi == 0x7fffffff || od > bd || (od == bd && oi < bi))) { bd = od; bi = oi; }
        }
        if (bi == 0x7fffffff || bd < 0.01) { depth--; w.sync(); continue; }
        if (w.lane == 0) {
          f.pf = gt_pt(es, bi); f.far = bi; f.stage = 1;
          GtWork::Frame &c = wk.stack[depth + 1];
          c.left = f.pf; c.right = f.right; c.stage = 0; c.far = -1;
        }
        depth++;
        w.sync();
      } else if (stage == 1) {
        if (w.lane == 0) {
          wk.hull[hs] = f.pf;
          f.stage = 2;
          GtWork::Frame &c = wk.stack[depth + 1];
          c.left = f.left; c.right = f.pf; c.stage = 0; c.far = -1;
        }
        hs++;
        depth++;
        w.sync();
      } else {
        depth--;
        w.sync();
      }
    }
  }
  return hs;
}

// removeShortLS + pickExternalLS + sumLength + pickLongestLS + sortByAngle + findCorners + the acceptance tests of
// executeCPUTask (oclrect.c:1129-1139) on the ne edges in wk.e0.  Every lane returns the same GtQuad.
template <class W>
GT_FN GtQuad gt_try_quad(W &w, GtWork &wk, int ne, int status) {
  GtQuad q;
  q.valid = 0; q.status = status;
  for (int i = 0; i < 4; i++) q.c[i] = GtP2{0, 0};
  q.centre = GtP2{0, 0};
  if (ne < 4) { GT_DBG(1); return q; }                                           // pickExternalLS never adds edges; fewer than four never pass
  // removeShortLS(0.05): only lists longer than four are sorted and trimmed (oclrect.c:926-943)
  gt_sort_edges(w, wk.e1, wk.key, wk.e0, ne);
  const GtEdge *pts_src = wk.e0;                                  // the hull sees the edges in the order removeShortLS leaves them
  int first = 0, n = ne;
  if (ne > 4) {
    const float longest = wk.key[ne - 1], r2 = 0.05f * 0.05f;
    int drop = 0;
    for (int i = w.lane; i < ne; i += 32) drop += !(wk.key[i] / longest > r2) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) drop += w.shfl(drop, w.lane ^ o);
    if (drop > ne - 4) drop = ne - 4;
    first = drop; n = ne - drop;
    pts_src = wk.e1 + first;
  }
  GtEdge *es = wk.e1 + first;                                     // sorted by length, n entries
  float *key = wk.key + first;
  for (int i = w.lane; i < n; i += 32) wk.alive[i] = 1;
  w.sync();
  const int hs = gt_convex_hull(w, pts_src, n, wk);
  GT_DBG2("hs %d n %d first %d\n", hs, n, first);
  // pickExternalLS (oclrect.c:945-992): per hull edge the longest segment lying on it
  int nk = 0;
  for (int i = 0; i < hs; i++) {
    const GtP2 q0 = wk.hull[i], q1 = wk.hull[(i + 1) % hs];
    const GtP2 mid = gt_mul(gt_add(q0, q1), 0.5), dir = gt_unit(gt_sub(q0, q1));
    const double hl2 = gt_dist2(q0, q1);
    int taken = -1;
    for (int base = n - 1; base >= 0 && taken < 0; base -= 32) {
      const int j = base - w.lane;
      bool hit = false;
      if (j >= 0 && wk.alive[j]) {
        const GtEdge e = es[j];
        const double d = gt_dist2(mid, gt_footOnSegment(e.a, e.b, mid));
        hit = d < 1 || (fabs(gt_dot(dir, gt_unit(gt_sub(e.a, e.b)))) > 0.95 && d / hl2 < 0.01);
      }
      const unsigned b = w.ballot(hit);
      if (b) { int l = 0; while (!((b >> l) & 1u)) l++; taken = base - l; }
    }
    if (taken >= 0) {
      if (w.lane == 0) { wk.kept[nk] = es[taken]; wk.alive[taken] = 0; }
      nk++;
    }
    w.sync();
  }
  (void)key;
  if (nk < 4) { GT_DBG(2); return q; }
  // the rest works on at most nk <= hs edges and is serial in the reference's order (sums, sorts of four)
  double len0 = 0;
  for (int i = 0; i < nk; i++) len0 += sqrt((double)gt_len2f(wk.kept[i]));
  GtEdge e4[4];
  if (nk > 4) {                                                   // pickLongestLS(4): stable sort by length, the last four from the back
    gt_sort_edges(w, wk.e0, wk.key, wk.kept, nk);
    for (int i = 0; i < 4; i++) e4[i] = wk.e0[nk - 1 - i];
  } else {
    for (int i = 0; i < 4; i++) e4[i] = wk.kept[i];
  }
  // gv (oclrect.c:864-877) and sortByAngle (:829-862, stable)
  GtP2 g = {0, 0};
  double total = 0;
  for (int i = 0; i < 4; i++) {
    const double len = sqrt(gt_dist2(e4[i].a, e4[i].b));
    g = gt_add(g, gt_mul(gt_add(e4[i].a, e4[i].b), len));
    total += len;
  }
  const GtP2 centre = gt_mul(g, 0.5 / total);
  double ang[4];
  for (int i = 0; i < 4; i++) {
    GtP2 v = gt_sub(e4[i].a, e4[i].b);
    v = GtP2{v.y, -v.x};
    if (gt_dot(v, gt_sub(e4[i].a, centre)) < 0) v = gt_mul(v, -1);
    ang[i] = atan2(v.x, v.y);
  }
  for (int i = 1; i < 4; i++) {                                   // insertion sort = stable
    const GtEdge e = e4[i]; const double a = ang[i];
    int j = i - 1;
    while (j >= 0 && a < ang[j]) { e4[j + 1] = e4[j]; ang[j + 1] = ang[j]; j--; }
    e4[j + 1] = e; ang[j + 1] = a;
  }
  // findCorners (oclrect.c:1011-1045)
  GtP2 c[4];
  for (int i = 0; i < 4; i++) {
    c[i] = gt_lineIntersection(e4[i], e4[(i + 1) & 3]);
    if (isnan(c[i].x)) { GT_DBG(3); return q; }
  }
  for (int i = 0; i < 4; i++) { e4[i].a = c[i]; e4[i].b = c[(i + 1) & 3]; }
  double len1 = 0;
  for (int i = 0; i < 4; i++) len1 += sqrt((double)gt_len2f(e4[i]));
  for (int i = 0; i < 4; i++) {                                   // closeToTriangle(0.001), oclrect.c:886-895
    const GtEdge &a = e4[i], &b = e4[(i + 1) & 3];
    const double d0 = gt_dist2(a.b, gt_footOnLine(a.a, b.b, a.b));
    const double d1 = gt_dist2(a.a, b.b);
    if (d0 / d1 < 0.001) { GT_DBG(4); return q; }
  }
  if (len1 / len0 > 2) { GT_DBG(5); return q; }
  {                                                               // isConvex, oclrect.c:897-922
    bool sign = false;
    for (int i = 0; i < 4; i++) {
      const GtEdge &a = e4[i], &b = e4[(i + 1) & 3];
      const bool t = (a.b.x - a.a.x) * (b.b.y - b.a.y) - (a.b.y - a.a.y) * (b.b.x - b.a.x) > 0;
      if (i == 0) sign = t;
      else if (t != sign) { GT_DBG(6); return q; }
    }
  }
  g = GtP2{0, 0}; total = 0;
  for (int i = 0; i < 4; i++) {
    const double len = sqrt(gt_dist2(e4[i].a, e4[i].b));
    g = gt_add(g, gt_mul(gt_add(e4[i].a, e4[i].b), len));
    total += len;
  }
  q.centre = gt_mul(g, 0.5 / total);
  for (int i = 0; i < 4; i++) q.c[i] = c[i];
  q.valid = 1;
  { GT_DBG(7); return q; }
}

// ================================================================ pose estimation (oclrect.c:427-634)
struct GtPose { GtP3 ray[4]; int mode; };
#define GT_H 1e-6
GT_FN_NOINLINE double gt_objective(GtP4 d, const GtPose &ps) {                         // oclrect.c:441-477
  const int m = ps.mode;
  GtP3 q[4];
  for (int i = 0; i < 4; i++) q[i] = gt_mul(ps.ray[i], d.v[i]);
  double score = 0;
  const double l01 = gt_dist2(q[0], q[1]), l12 = gt_dist2(q[1], q[2]), l23 = gt_dist2(q[2], q[3]);
  const double l03 = gt_dist2(q[0], q[3]), l02 = gt_dist2(q[0], q[2]), l13 = gt_dist2(q[1], q[3]);
  score += gt_sq((m ? l23 : l03) - 1);
  score += gt_sq((m ? l01 : l12) - 1);
  const double comp = 1.0 / (m ? l12 : l01);
  {
    const GtP3 s = gt_add(gt_sub(m ? q[0] : q[2], q[1]), gt_sub(m ? q[2] : q[0], q[3]));
    score += gt_dot(s, s);
  }
  {
    const GtP3 s = gt_add(gt_sub(q[1], m ? q[2] : q[0]), gt_sub(q[3], m ? q[0] : q[2]));
    score += comp * gt_dot(s, s);
  }
  score += gt_sq(l01 + l12 - l02);
  score += gt_sq(l03 + l23 - l02);
  score += gt_sq(l01 + l03 - l13);
  score += gt_sq(l12 + l23 - l13);
  const GtP3 n013 = gt_cross(gt_sub(q[1], q[0]), gt_sub(q[3], q[0]));
  score += comp * gt_sq(gt_dot(n013, q[2]) - gt_dot(n013, q[0])) / gt_dot(n013, n013);
  const GtP3 n102 = gt_cross(gt_sub(q[0], q[1]), gt_sub(q[2], q[1]));
  score += comp * gt_sq(gt_dot(n102, q[3]) - gt_dot(n102, q[1])) / gt_dot(n102, n102);
  return score;
}
GT_FN void gt_gradDiag(GtP4 x, const GtPose &ps, GtP4 &g, GtP4 &h) {                   // oclrect.c:492-512
  const double fx = gt_objective(x, ps);
  for (int i = 0; i < 4; i++) {
    GtP4 e;
    for (int j = 0; j < 4; j++) { e.v[j] = 0; if (j == i) e.v[j] = GT_H; }
    const double fm = gt_objective(gt_sub4(x, e), ps);
    const double fp = gt_objective(gt_add4(x, e), ps);
    g.v[i] = (fp - fm) / (2 * GT_H);
    h.v[i] = (fm - 2 * fx + fp) / (GT_H * GT_H);
  }
}
GT_FN GtP4 gt_lineSearch(GtP4 x, GtP4 dir, int iters, const GtPose &ps) {               // oclrect.c:514-536
  dir = gt_unit4(dir);
  double sc = 1.0;
  for (int i = 0; i < iters; i++) {
    const double f0 = gt_objective(x, ps);
    const double fp = gt_objective(gt_add4(x, gt_mul4(dir, GT_H)), ps);
    const double fm = gt_objective(gt_add4(x, gt_mul4(dir, -GT_H)), ps);
    const double d1 = (fp - fm) * (1.0 / (2 * GT_H));
    double d2 = (fp + fm - 2 * f0) * (1.0 / (GT_H * GT_H));
    if (d2 * d2 < 1e-10) d2 = 1;
    const double delta = fabs(d1 / d2);
    if (delta < 1e-10) return x;
    const GtP4 cand = gt_add4(x, gt_mul4(dir, delta * sc));
    const double f1 = gt_objective(cand, ps);
    if (f0 < f1) { sc *= 0.5; continue; }
    x = cand;
  }
  return x;
}
GT_FN GtP4 gt_precondition(GtP4 diag, GtP4 r) {                                        // oclrect.c:538-555
  for (int i = 0; i < 4; i++) if (diag.v[i] <= 0) return r;
  GtP4 a;
  for (int i = 0; i < 4; i++) { a.v[i] = 1.0 / diag.v[i]; a.v[i] *= r.v[i]; }
  return a;
}
GT_FN GtP4 gt_conjugateGradient(GtP4 x, int outer, int inner, const GtPose &ps) {       // oclrect.c:557-588
  int k = 0;
  GtP4 g, h;
  gt_gradDiag(x, ps, g, h);
  GtP4 r = gt_mul4(g, -1);
  GtP4 s = gt_precondition(h, r), d = s;
  double deltaNew = gt_dot4(r, d);
  for (int i = 0; i < outer; i++) {
    x = gt_lineSearch(x, d, inner, ps);
    gt_gradDiag(x, ps, g, h);
    r = gt_mul4(g, -1);
    const double deltaOld = deltaNew;
    const double deltaMid = gt_dot4(r, s);
    s = gt_precondition(h, r);
    deltaNew = gt_dot4(r, s);
    const double beta = (deltaNew - deltaMid) / deltaOld;
    if (k == 10 || beta <= 0 || deltaOld == 0) { d = s; k = 0; }
    else d = gt_add4(s, gt_mul4(d, beta));
    k++;
  }
  return x;
}
// which corner comes first, and the four rays (oclrect.c:590-606)
GT_FN int gt_pose_setup(const GtQuad &q, int iw, int ih, double tanAOV, GtP3 *ray) {
  int first = 0;
  double mn = 1e+100;
  for (int i = 0; i < 4; i++) {
    GtP2 v = gt_unit(gt_sub(q.c[(i + 1) & 3], q.c[i]));
    v = GtP2{-v.y, v.x};
    if (gt_dot(gt_sub(q.c[i], q.centre), v) < 0) v = gt_mul(v, -1);
    if (v.y < mn) { mn = v.y; first = i; }
  }
  for (int i = 0; i < 4; i++) {
    const GtP2 c = q.c[(i + first) & 3];
    ray[i] = gt_unit(GtP3{(double)(c.x - (iw / 2)), (double)(-(c.y - ih / 2)), (double)(iw / 2 / tanAOV)});
  }
  return first;
}
struct GtPoseOut { GtP4 x; double val; };
// one of the two runs of poseEstimation (mode 1 first in the reference, oclrect.c:608-616)
GT_FN GtPoseOut gt_pose_run(const GtP3 *ray, int mode) {
  GtPose ps;
  for (int i = 0; i < 4; i++) ps.ray[i] = ray[i];
  ps.mode = mode;
  GtP4 x0;
  if (mode) {
    const double d01 = 1.0 / sqrt(gt_dist2(ray[0], ray[1])), d23 = 1.0 / sqrt(gt_dist2(ray[2], ray[3]));
    x0 = GtP4{{d01, d01, d23, d23}};
  } else {
    const double d12 = 1.0 / sqrt(gt_dist2(ray[1], ray[2])), d03 = 1.0 / sqrt(gt_dist2(ray[0], ray[3]));
    x0 = GtP4{{d03, d12, d12, d03}};
  }
  GtPoseOut o;
  o.x = gt_conjugateGradient(x0, 12, 10, ps);
  o.val = gt_objective(o.x, ps);
  return o;
}
struct GtRect { double c2[4][2]; double c3[4][3]; double value; uint32_t status; uint32_t pad; };   // rect_t, oclrect.h:5-15 (176 bytes)
// the rest of poseEstimation (oclrect.c:618-634) + looksLikeAScreen (:636-656)
GT_FN void gt_pose_finish(const GtQuad &q, int first, const GtP3 *ray, const GtPoseOut &m1, const GtPoseOut &m0, GtRect &out) {
  const double val0 = m1.val, val1 = m0.val;
  out.value = val0 < val1 ? val0 : val1;
  GtP4 x = val0 < val1 ? m1.x : m0.x;
  if (x.v[0] < 0) x = gt_mul4(x, -1);
  GtP3 c3[4]; GtP2 c2[4];
  for (int i = 0; i < 4; i++) {
    c3[i] = gt_mul(ray[i], x.v[i]);
    c2[i] = q.c[(i + first) & 3];
    out.c3[i][0] = c3[i].x; out.c3[i][1] = c3[i].y; out.c3[i][2] = c3[i].z;
    out.c2[i][0] = c2[i].x; out.c2[i][1] = c2[i].y;
  }
  out.status = (uint32_t)q.status;
  out.pad = 0;
  bool screen = !(out.value > 0.05);
  if (screen && (c3[0].z < 0 || c3[1].z < 0 || c3[2].z < 0 || c3[3].z < 0)) screen = false;
  if (screen) {
    const double asp = sqrt(gt_dist2(c3[0], c3[1])) / sqrt(gt_dist2(c3[1], c3[2]));
    if (asp < 1.0 / 12 || 12 < asp) screen = false;
  }
  if (screen) {
    double maxs = 0, mins = 1e+100;
    for (int i = 0; i < 4; i++) {
      const double s0 = gt_dist2(c2[(i + 2) % 4], gt_footOnSegment(c2[i], c2[(i + 1) % 4], c2[(i + 2) % 4]));
      const double s1 = gt_dist2(c2[(i + 3) % 4], gt_footOnSegment(c2[i], c2[(i + 1) % 4], c2[(i + 3) % 4]));
      maxs = fmax(maxs, fmax(s0, s1));
      mins = fmin(mins, fmax(s0, s1));
    }
    if (maxs / mins > 100) screen = false;
  }
  if (screen) out.status |= 1;
}


// ================================================================ grouping: segments -> regions -> candidates
// Per-frame storage, carved by gt_layout() (a pure function of the segment count n, so every kernel derives the same pointers).  table: two ints per region id (pixel index of the region's root), zero between frames:
// [2*segid] = number of distinct segments touching the region (later -(region index + 1)), [2*segid + 1] = 0x7fffffff - first sample.
struct GtHdr { int n, npairs, nreg, nmember, nchain, ncand, nvalid, nrect, err, pad[7]; };
struct GtPair { int segid, ls, f; };
struct GtRegion { int segid, cnt, f, off, fill; };
struct GtChain { int head, m; };
struct GtCand { int type, idx, m, pad; unsigned long long off; };        // type 0: region, 1: polyline chain; off: byte offset of its work storage
struct GtLayout {
  GtHdr *hdr; GtPair *pairs; GtRegion *regions; GtChain *chains; int *members; GtCand *cands; GtQuad *quads; GtPoseOut *pose; int *vlist;
  unsigned char *work; size_t workBytes; int ok, okPersist;
};
#define GT_ERR_SCRATCH 1        // the frame needs more tail scratch than the arena holds
#define GT_ERR_RECTS 2          // more rectangles than the read-back record holds
GT_FN size_t gt_align16(size_t v) { return (v + 15) & ~(size_t)15; }
// scratch: volatile work space (dead once kt_quad is through); persist: what the pose / finish kernels read (header, quadrilaterals,
// pose results) - it lives in the per-page read-back record, so those two kernels can be re-run at poll time for another tanAOV
GT_FN GtLayout gt_layout(unsigned char *scratch, size_t S, unsigned char *persist, size_t PS, int n) {
  GtLayout L;
  const size_t np = (size_t)15 * n + 16, nr = np / 4 + 1, nc = nr + n + 1;
  size_t o = 0;
  L.hdr = (GtHdr *)(persist + o); o += gt_align16(sizeof(GtHdr));
  L.quads = (GtQuad *)(persist + o); o += gt_align16(nc * sizeof(GtQuad));
  L.pose = (GtPoseOut *)(persist + o); o += gt_align16(2 * nc * sizeof(GtPoseOut));
  L.vlist = (int *)(persist + o); o += gt_align16(nc * sizeof(int));
  L.ok = o <= PS;
  o = 0;
  L.pairs = (GtPair *)(scratch + o); o += gt_align16(np * sizeof(GtPair));
  L.regions = (GtRegion *)(scratch + o); o += gt_align16(nr * sizeof(GtRegion));
  L.chains = (GtChain *)(scratch + o); o += gt_align16((size_t)(n + 1) * sizeof(GtChain));
  L.members = (int *)(scratch + o); o += gt_align16(np * sizeof(int));
  L.cands = (GtCand *)(scratch + o); o += gt_align16(nc * sizeof(GtCand));
  L.work = scratch + o;
  L.okPersist = L.ok;                                             // all the pose / finish kernels need
  L.ok = L.ok && o < S;
  L.workBytes = L.ok ? S - o : 0;
  return L;
}

#ifdef __CUDACC__
#define GT_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define GT_ATOMIC_MAX(p, v) atomicMax((p), (v))
#else
static inline int gt_host_fetch_add(int *p, int v) { const int o = *p; *p = o + v; return o; }
static inline int gt_host_fetch_max(int *p, int v) { const int o = *p; if (v > o) *p = v; return o; }
#define GT_ATOMIC_ADD(p, v) gt_host_fetch_add((p), (v))
#define GT_ATOMIC_MAX(p, v) gt_host_fetch_max((p), (v))
#endif

// segment i: its (region, segment) pairs (oclrect.c:1066-1098) and, if it heads a polyline chain, the chain (oclrect.c:1169-1183)
GT_FN void gt_item_samples(int i, const GtLS *ls, const int *segidMap, int *table, const GtLayout &L, int n, int iw, int ih) {
  const GtLS e = ls[i];
  if (e.polyid == 0) return;
  int seg[15];
  for (int k = 0; k < 15; k++) {
    const int p = gt_sample_pixel(e, k, iw, ih);
    seg[k] = p < 0 ? 0 : segidMap[p];
  }
  for (int k = 0; k < 15; k++) {
    if (seg[k] <= 0) continue;
    bool seen = false;
    for (int j = 0; j < k; j++) seen |= seg[j] == seg[k];
    if (seen) continue;
    const int f = i * 15 + k;
    const int p = GT_ATOMIC_ADD(&L.hdr->npairs, 1);
    L.pairs[p] = GtPair{seg[k], i, f};
    GT_ATOMIC_ADD(&table[2 * (size_t)seg[k]], 1);
    GT_ATOMIC_MAX(&table[2 * (size_t)seg[k] + 1], 0x7fffffff - f);
  }
  if (e.leftPtr > 0) return;
  int m = 0, steps = 0;
  for (int j = i; j > 0 && steps <= n; j = ls[j].rightPtr, steps++) {
    const GtLS q = ls[j];
    if (gt_dist2(GtP2{(double)q.x0, (double)q.y0}, GtP2{(double)q.x1, (double)q.y1}) > 32.0 * 32.0) m++;
  }
  const int c = GT_ATOMIC_ADD(&L.hdr->nchain, 1);
  L.chains[c] = GtChain{i, m};
}
// pair p: the pair that saw its region first opens the region if at least four segments touch it
GT_FN void gt_item_regions(int p, int *table, const GtLayout &L) {
  const GtPair pr = L.pairs[p];
  if (0x7fffffff - table[2 * (size_t)pr.segid + 1] != pr.f) return;
  const int cnt = table[2 * (size_t)pr.segid];
  if (cnt < 4) return;
  const int r = GT_ATOMIC_ADD(&L.hdr->nreg, 1);
  const int off = GT_ATOMIC_ADD(&L.hdr->nmember, cnt);
  L.regions[r] = GtRegion{pr.segid, cnt, pr.f, off, 0};
  table[2 * (size_t)pr.segid] = -(r + 1);
}
GT_FN void gt_item_members(int p, const int *table, const GtLayout &L) {
  const GtPair pr = L.pairs[p];
  const int v = table[2 * (size_t)pr.segid];
  if (v >= 0) return;
  GtRegion &rg = L.regions[-v - 1];
  const int slot = GT_ATOMIC_ADD(&rg.fill, 1);
  L.members[rg.off + slot] = pr.ls;
}
// candidate order of executeCPUTask: regions in ArrayMap iteration order (bucket of the region id, then first insertion,
// helper.c:124-190), then the chains by ascending head
GT_FN unsigned long long gt_region_key(const GtRegion &r) { return ((unsigned long long)gt_bucketOf((uint64_t)r.segid) << 32) | (unsigned)r.f; }
GT_FN void gt_item_order_region(int r, const GtLayout &L) {
  const int nreg = L.hdr->nreg;
  const unsigned long long k = gt_region_key(L.regions[r]);
  int rank = 0;
  for (int j = 0; j < nreg; j++) rank += gt_region_key(L.regions[j]) < k ? 1 : 0;
  L.cands[rank] = GtCand{0, r, L.regions[r].cnt, 0, 0};
}
GT_FN void gt_item_order_chain(int c, const GtLayout &L) {
  const int nreg = L.hdr->nreg, nchain = L.hdr->nchain;
  const int h = L.chains[c].head;
  int rank = nreg;
  for (int j = 0; j < nchain; j++) rank += L.chains[j].head < h ? 1 : 0;
  L.cands[rank] = GtCand{1, c, L.chains[c].m, 0, 0};
}

// candidate -> quadrilateral: builds the edge list (region: members in ascending segment id, clipped by their vote boxes;
// chain: the long segments along the chain) and runs gt_try_quad
template <class W>
GT_FN GtQuad gt_cand_quad(W &w, const GtCand &cd, const GtLS *ls, const int *votes, const GtLayout &L, int n, int iw, int ih, int nentry) {
  GtWork wk = gt_work_carve(L.work + cd.off, cd.m);
  int ne = 0;
  if (cd.type == 0) {
    const GtRegion rg = L.regions[cd.idx];
    int *mem = L.members + rg.off;
    int *sorted = wk.alive;                                        // free until gt_try_quad
    for (int i = w.lane; i < rg.cnt; i += 32) {
      const int v = mem[i];
      int r = 0;
      for (int j = 0; j < rg.cnt; j++) r += mem[j] < v ? 1 : 0;
      sorted[r] = v;
    }
    w.sync();
    for (int base = 0; base < rg.cnt; base += 32) {
      const int j = base + w.lane;
      GtEdge e;
      int has = 0;
      if (j < rg.cnt) {
        const int id = sorted[j];
        has = gt_region_edge(ls[id], id, votes + (size_t)gt_vote_slot(id, rg.segid, nentry) * 5, iw, ih, e);
      }
      const unsigned b = w.ballot(has != 0);
      if (has) {
        int before = 0;
        for (int l = 0; l < w.lane; l++) before += (b >> l) & 1u;
        wk.e0[ne + before] = e;
      }
      int tot = 0;
      for (int l = 0; l < 32; l++) tot += (b >> l) & 1u;
      ne += tot;
    }
    w.sync();
  } else {
    const int head = L.chains[cd.idx].head;
    int steps = 0;
    for (int j = head; j > 0 && steps <= n; j = ls[j].rightPtr, steps++) {       // every lane walks the chain, lane 0 writes
      const GtLS q = ls[j];
      const GtP2 a = {(double)q.x0, (double)q.y0}, b = {(double)q.x1, (double)q.y1};
      if (gt_dist2(a, b) > 32.0 * 32.0) { if (w.lane == 0) wk.e0[ne] = GtEdge{a, b}; ne++; }
    }
    w.sync();
  }
  return gt_try_quad(w, wk, ne, cd.type == 0 ? 0 : 2);
}

#endif
