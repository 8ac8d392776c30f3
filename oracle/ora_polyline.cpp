// ora_polyline.cpp - CPU oracle, Stage C: restatement of oclpolyline.cl and of oclpolyline_execute
// (oclpolyline.c:154-309).  TEST INFRASTRUCTURE ONLY (see rd_oracle.h).
#include <algorithm>
#include <vector>
#include "ora_internal.h"

namespace ora {

typedef ora_ls_t LS_t;                 // oclpolyline.cl:29-39, 56 bytes

struct LSX_t {                         // oclpolyline.cl:41-45, 56 bytes
  int64_t mx00, mx01, mx11, my0, my1;
  int16_t dirSEx, dirSEy, vDirSEx, vDirSEy;
  int32_t distSquSE, padding;
};
static_assert(sizeof(LS_t) == 56, "LS_t must be 56 bytes");
static_assert(sizeof(LSX_t) == 56, "LSX_t must be 56 bytes");

#define MINEDGELEN 1
#define MINNINDEX 4

// oclpolyline.cl:47-59
static inline float distanceSqu(float vx, float vy, float wx, float wy) {
  return (vx - wx) * (vx - wx) + (vy - wy) * (vy - wy);
}

static inline void closestPoint(float vx, float vy, float wx, float wy, float px, float py, float &ox, float &oy) {
  float l2 = distanceSqu(vx, vy, wx, wy);
  if (l2 <= 1e-4f) { ox = vx; oy = vy; return; }
  float t = ((px - vx) * (wx - vx) + (py - vy) * (wy - vy)) / l2;
  if (t < 0.0f) { ox = vx; oy = vy; return; }
  if (t > 1.0f) { ox = wx; oy = wy; return; }
  ox = vx + t * (wx - vx);
  oy = vy + t * (wy - vy);
}

// ---- oclpolyline.cl:66-87 : note the test is != 0 here, > 0 in the oclrect.cl copy ----
static void k_simpleJunction(int32_t *out, const int32_t *in, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      out[p0] = 0;
      if (x <= 0 || y <= 0 || x >= (iw - 1) || y >= (ih - 1)) continue;
      if (in[p0] == 0) continue;
      int count = 1;
      for (int i = 0; i < 8; i++)
        if (in[p0 + RX[i] + RY[i] * iw] != 0) count++;
      out[p0] = count == 1 ? 0 : count;
    }
}

// ---- oclpolyline.cl:89-110 : the 2-px border of `out` is NOT written (stale contents stay) ----
static void k_simpleConnect(int32_t *out, const int32_t *in, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 2; y < ih - 2; y++)
    for (int x = 2; x < iw - 2; x++) {
      const int p0 = y * iw + x;
      out[p0] = in[p0] != 0 ? 1 : 0;
      if (in[p0] != 0) continue;
      if (in[p0 - 2] != 0 && in[p0 - 1] == 2 && in[p0 + 1] == 2 && in[p0 + 2] != 0) out[p0] = 1;
      if (in[p0 - iw * 2] != 0 && in[p0 - iw] == 2 && in[p0 + iw] == 2 && in[p0 + iw * 2] != 0) out[p0] = 1;
      if (in[p0 - iw * 2 - 2] != 0 && in[p0 - iw - 1] == 2 && in[p0 + iw + 1] == 2 && in[p0 + iw * 2 + 2] != 0) out[p0] = 1;
      if (in[p0 - iw * 2 + 2] != 0 && in[p0 - iw + 1] == 2 && in[p0 + iw - 1] == 2 && in[p0 + iw * 2 - 2] != 0) out[p0] = 1;
      if (in[p0 + 2] != 0 && in[p0 + 1] == 2 && in[p0 + iw - 1] == 2 && in[p0 + iw - 2] != 0) out[p0] = 1;
      if (in[p0 - 2] != 0 && in[p0 - 1] == 2 && in[p0 + iw + 1] == 2 && in[p0 + iw + 2] != 0) out[p0] = 1;
      if (in[p0 - iw * 2 + 1] != 0 && in[p0 - iw + 1] == 2 && in[p0 + iw] == 2 && in[p0 + iw * 2] != 0) out[p0] = 1;
      if (in[p0 - iw * 2 - 1] != 0 && in[p0 - iw - 1] == 2 && in[p0 + iw] == 2 && in[p0 + iw * 2] != 0) out[p0] = 1;
    }
}

// ---- oclpolyline.cl:112-124 ----
static void k_stringify(int32_t *out, const int32_t *in, int mod2, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      out[p0] = in[p0];
      if (x <= 0 || y <= 0 || x >= (iw - 1) || y >= (ih - 1)) continue;
      if (((x + y) & 1) != mod2) continue;
      if (in[p0 - iw] != 0 && in[p0 - 1] != 0) out[p0] = 0;
      if (in[p0 - iw] != 0 && in[p0 + 1] != 0) out[p0] = 0;
      if (in[p0 + iw] != 0 && in[p0 - 1] != 0) out[p0] = 0;
      if (in[p0 + iw] != 0 && in[p0 + 1] != 0) out[p0] = 0;
    }
}

// ---- oclpolyline.cl:126-147 ----
static void k_removeBranch(int32_t *out, const int32_t *in, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      out[p0] = 0;
      if (x <= 0 || y <= 0 || x >= (iw - 1) || y >= (ih - 1)) continue;
      if (in[p0] == 0) continue;
      int count = 0;
      for (int i = 0; i < 8; i++)
        if (in[p0 + RX[i] + RY[i] * iw] != 0) count++;
      out[p0] = count <= 2 ? 1 : 0;
    }
}

// ---- oclpolyline.cl:149-167.  Q7: the non-atomic ++ is only ever tested against 0 ----
static void k_countEnds(int32_t *out, const int32_t *junction, const int32_t *label, int iw, int ih) {
  for (int y = 1; y < ih - 1; y++)
    for (int x = 1; x < iw - 1; x++) {
      const int p0 = y * iw + x;
      if (junction[p0] == 2) out[label[p0]]++;
    }
}

static void k_breakLoops(int32_t *edgeinout, int32_t *labelinout, const int32_t *nEnds, int iw, int ih) {
  for (int y = 1; y < ih - 1; y++)
    for (int x = 1; x < iw - 1; x++) {
      const int p0 = y * iw + x;
      if (labelinout[p0] != p0) continue;
      if (nEnds[p0] == 0) { edgeinout[p0] = 0; labelinout[p0] = -1; }
    }
}

// ---- oclpolyline.cl:169-191 ----
static inline void getnp(const int32_t *labelin, int p0, int iw, int &nx, int &ny) {
  const int l = labelin[p0];
  int i;
  for (i = 0; i < 8; i++)
    if (labelin[p0 + RX[i] + RY[i] * iw] == l) break;
  nx = i < 8 ? (p0 + RX[i] + RY[i] * iw) : p0;
  for (i++; i < 8; i++)
    if (labelin[p0 + RX[i] + RY[i] * iw] == l) break;
  ny = i < 8 ? (p0 + RX[i] + RY[i] * iw) : p0;
}

// ---- oclpolyline.cl:193-220 ----
static void k_findEnds0(int32_t *nextout, int32_t *prevout, int32_t *flagout, const int32_t *labelin, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      nextout[p0] = prevout[p0] = flagout[p0] = -1;
      if (x <= 0 || y <= 0 || x >= (iw - 1) || y >= (ih - 1) || labelin[p0] == -1) continue;
      int npx, npy;
      getnp(labelin, p0, iw, npx, npy);
      nextout[p0] = npx;
      prevout[p0] = npy;
      int flag = 0;
      if (npx != p0) {
        int a, b;
        getnp(labelin, npx, iw, a, b);
        if (a == p0) flag |= 1;
      }
      if (npy != p0) {
        int a, b;
        getnp(labelin, npy, iw, a, b);
        if (b == p0) flag |= 2;
      }
      flagout[p0] = flag;
    }
}

// ---- oclpolyline.cl:222-267.  flaginout is updated in place, but a launch only rewrites the two bits
// that launch does not read, so a snapshot-free implementation is race free. ----
static void k_findEnds1(int32_t *nextout, int32_t *prevout, int32_t *flaginout, const int32_t *nextin, const int32_t *previn,
                        const int32_t *labelin, int page, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      nextout[p0] = prevout[p0] = -1;
      if (x <= 0 || y <= 0 || x >= (iw - 1) || y >= (ih - 1) || labelin[p0] == -1) continue;
      const int f0 = __atomic_load_n(&flaginout[p0], __ATOMIC_RELAXED);
      bool revn = page == 0 ? ((f0 & 1) != 0) : ((f0 & 4) != 0);
      bool revp = page == 0 ? ((f0 & 2) != 0) : ((f0 & 8) != 0);
      int nn = nextin[p0], pp = previn[p0];
      for (int i = 0; i < 8; i++) {
        int nn2 = revn ? previn[nn] : nextin[nn];
        int pp2 = revp ? nextin[pp] : previn[pp];
        int nflag = __atomic_load_n(&flaginout[nn], __ATOMIC_RELAXED);
        int pflag = __atomic_load_n(&flaginout[pp], __ATOMIC_RELAXED);
        if (page != 0) { nflag >>= 2; pflag >>= 2; }
        revn = revn ? ((nflag & 2) == 0) : ((nflag & 1) != 0);
        revp = revp ? ((pflag & 1) == 0) : ((pflag & 2) != 0);
        nn = nn2;
        pp = pp2;
      }
      nextout[p0] = nn;
      prevout[p0] = pp;
      int f = f0;
      if (page == 0) {
        f &= 3;
        f |= revn ? 4 : 0;
        f |= revp ? 8 : 0;
      } else {
        f &= (3 << 2);
        f |= revn ? 1 : 0;
        f |= revp ? 2 : 0;
      }
      __atomic_store_n(&flaginout[p0], f, __ATOMIC_RELAXED);
    }
}

// ---- oclpolyline.cl:269-285 ----
static void k_findEnds2(int32_t *numout, int32_t *linkout, const int32_t *nextin, const int32_t *previn, const int32_t *labelin, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      numout[p0] = 0; linkout[p0] = -1;
      if (x <= 0 || y <= 0 || x >= (iw - 1) || y >= (ih - 1)) continue;
      if (labelin[p0] == -1) continue;
      int npx, npy;
      getnp(labelin, p0, iw, npx, npy);
      linkout[p0] = nextin[p0] < previn[p0] ? npx : npy;
      numout[p0] = linkout[p0] == p0 ? 0 : 1;
    }
}

// ---- oclpolyline.cl:287-310 ----
static void k_number(int32_t *numout, int32_t *linkout, const int32_t *numin, const int32_t *linkin, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      numout[p0] = 0; linkout[p0] = -1;
      if (x <= 0 || y <= 0 || x >= (iw - 1) || y >= (ih - 1)) continue;
      if (linkin[p0] == -1) { numout[p0] = numin[p0]; linkout[p0] = linkin[p0]; continue; }
      int no = numin[p0], lo = linkin[p0];
      bool bail = false;
      for (int i = 0; i < 32; i++) {
        if (!(0 < lo && lo < (iw * ih))) { bail = true; break; }
        no += numin[lo];
        lo = linkin[lo];
      }
      if (bail) continue;
      numout[p0] = no;
      linkout[p0] = lo;
    }
}

// ---- oclpolyline.cl:312-355 + oclpolyline.c:170-184 : labelpl (N=12 -> 11 passes) ----
// CANONICAL (Q6): fixed point = smallest index of the component, where two 8-neighbours belong together
// when both numbers are non-zero and differ by at most 1 (compared on number+1, as the kernel does).
static void labelpl(int32_t *label, int32_t *pixinout, int32_t *flags, int iw, int ih) {
  const int n = iw * ih;
  for (int p = 0; p < n; p++) {
    label[p] = p;                       // union-find parent; zero pixels are fixed up below
    pixinout[p] = pixinout[p] == 0 ? 0 : pixinout[p] + 1;
  }
  for (int i = 0; i <= 12 && i < iw; i++) flags[i] = i == 0 ? 1 : 0;
  MinUF uf(label);
  for (int y = 1; y < ih - 1; y++)
    for (int x = 1; x < iw - 1; x++) {
      const int p0 = y * iw + x;
      if (pixinout[p0] == 0) continue;
      for (int i = 0; i < 8; i++) {
        const int p1 = p0 + RX[i] + RY[i] * iw;
        // the adopter must be a non-border non-zero pixel; the neighbour may be any pixel whose label
        // is smaller.  A zero neighbour has pix 0 and label 0; |pix0 - 0| <= 1 needs pix0 == 1, which
        // cannot happen (non-zero pix are number+1 >= 2), so zero pixels never take part.
        if (pixinout[p1] == 0) continue;
        int d = pixinout[p0] - pixinout[p1];
        if (d < 0) d = -d;
        if (d <= 1) uf.unite(p0, p1);
      }
    }
  int comps = 0;
  for (int p = 0; p < n; p++) {          // increasing p: the root (smallest index) of p is already final
    if (pixinout[p] == 0) continue;      // zero pixels were never united with anything
    label[p] = uf.find(p);
    if (label[p] == p) comps++;
  }
  for (int p = 0; p < n; p++)
    if (pixinout[p] == 0) label[p] = 0;  // labelpl_preprocess: label = 0 where the number is 0
  g_stats.labelpl_components = comps;
}

// ---- oclpolyline.cl:357-378 ----
static void k_calcSize(int32_t *out, const int32_t *label, int iw, int ih) {
  for (int p0 = 0; p0 < iw * ih; p0++) {
    int b = label[p0];
    if (b != 0) out[b]++;
  }
}

static void k_filterSize(int32_t *out, const int32_t *labelin, const int32_t *sizein, int sizethre, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int p0 = 0; p0 < iw * ih; p0++) {
    int b = labelin[p0];
    out[p0] = sizein[b] > sizethre ? b : 0;
  }
}

// ---- oclpolyline.cl:380-420.  CANONICAL (Q4): new ids 1..K in raster order of the root pixels
// (the reference: atomic_inc arrival order). ----
static void k_relabel_pass0(int32_t *table, const int32_t *labelin, int iw, int ih) {
  for (int y = 1; y < ih - 1; y++)
    for (int x = 1; x < iw - 1; x++) {
      const int p0 = y * iw + x;
      int g = labelin[p0];
      if (g == 0 || p0 != g) continue;
      if (table[g + 1] == 0) table[g + 1] = ++table[0];
    }
}

static void k_relabel_pass1(int32_t *labelinout, const int32_t *tablein, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      if (x == 0 || y == 0 || x >= iw - 1 || y >= ih - 1) { labelinout[p0] = 0; continue; }
      int g = labelinout[p0];
      if (g == 0) continue;
      labelinout[p0] = tablein[g + 1];
    }
}

static inline bool ls_overflow(int g, int lsListSize) {
  // `g < 0 || lsListSize <= (g+1)*sizeof(LS_t)` : the right-hand side is size_t arithmetic
  return g < 0 || (size_t)lsListSize <= ((size_t)(g + 1)) * sizeof(LS_t);
}

// ---- oclpolyline.cl:439-472 ----
static void k_mkpl_pass0a(LS_t *gp, int lsListSize, const int32_t *numberin, const int32_t *labelin, int32_t *flags, int maxIter, int iw, int ih) {
  for (int x = 0; x < maxIter + 1 && x < iw; x++) flags[x] = x == 0 ? 1 : 0;
  for (int y = 1; y < ih - 1; y++)
    for (int x = 1; x < iw - 1; x++) {
      const int p0 = y * iw + x;
      int g = labelin[p0], n = numberin[p0];
      if (g == 0) continue;
      if (ls_overflow(g, lsListSize)) { g_stats.ls_overflow++; continue; }
      if (n == 1) {
        gp[g].x0 = (float)x; gp[g].y0 = (float)y;
        gp[g].level = 0;
        gp[g].startCount++;
      }
      gp[g].npix++;
      if (n > gp[g].endIndex) gp[g].endIndex = n;
      int32_t *cnt = (int32_t *)gp;
      if (g > *cnt) *cnt = g;
    }
}

// ---- oclpolyline.cl:475-506.  CANONICAL: when several pixels carry n == endIndex the first in raster
// order provides endCoords (the reference: whichever atomic_inc(&endCount) returns 0). ----
static void k_mkpl_pass0b(LS_t *gp, int lsListSize, const int32_t *numberin, const int32_t *labelin, int iw, int ih) {
  for (int y = 1; y < ih - 1; y++)
    for (int x = 1; x < iw - 1; x++) {
      const int p0 = y * iw + x;
      int g = labelin[p0], n = numberin[p0];
      if (g == 0) continue;
      if (ls_overflow(g, lsListSize)) { g_stats.ls_overflow++; continue; }
      if (n == gp[g].endIndex) {
        if (gp[g].startCount == 1 && gp[g].npix >= 2) {
          if (gp[g].endCount++ == 0) {
            gp[g].x1 = (float)x; gp[g].y1 = (float)y;
            gp[g].polyid = labelin[p0];
          }
        } else {
          gp[g].polyid = 0;
        }
      }
    }
}

// ---- oclpolyline.cl:509-540 ----
static void k_mkpl_pass1(LS_t *gp, int lsListSize, int32_t *tmp, const int32_t *labelin, const int32_t *randin, const int32_t *flags, int nIter, int iw, int ih) {
  if (flags[nIter - 1] == 0) return;
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      int g = labelin[p0];
      if (g == 0) continue;
      if (ls_overflow(g, lsListSize)) { g_stats.ls_overflow++; continue; }
      if (gp[g].polyid == 0) continue;
      int x0 = (int)gp[g].x0, y0 = (int)gp[g].y0;
      int x1 = (int)gp[g].x1, y1 = (int)gp[g].y1;
      float cx, cy;
      closestPoint((float)x0, (float)y0, (float)x1, (float)y1, (float)x, (float)y, cx, cy);
      int dist = (int)(hypot_c(cx - x, cy - y) * 65536);
      dist ^= (randin[p0] & 0x1fff);
      tmp[p0] = dist;
      if (dist > gp[g].maxDist) gp[g].maxDist = dist;
    }
}

// ---- oclpolyline.cl:543-615.  `old` is the copy made by the host before this launch (oclpolyline.c:207),
// `nw` the live list.  CANONICAL (Q4/Q8): at most one split per segment and iteration - the arg-max pixel,
// the first in raster order on an exact tie - and the new ids are handed out in raster order of the splitting
// pixels: the reference's atomic_inc arrival order when its work-items run in raster order, so that the ids are the
// ones oracle/_ref/librd_ref.so (the reference's own kernel, sequential schedule) produces. ----
static void k_mkpl_pass2(LS_t *nw, const LS_t *old, int lsListSize, const int32_t *tmp, const int32_t *numberin, const int32_t *labelin,
                         const int32_t *flags, int nIter, float minerror, int iw, int ih) {
  if (flags[nIter - 1] == 0) return;
  const int count = *(const int32_t *)old;
  std::vector<int32_t> winner((size_t)count + 1, -1);
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      int g = labelin[p0];
      if (g == 0) continue;
      if (ls_overflow(g, lsListSize)) continue;
      if (g > count) continue;
      if (old[g].polyid == 0) continue;
      if (tmp[p0] != old[g].maxDist) continue;
      if (winner[g] >= 0) { g_stats.mkpl_ties++; continue; }
      winner[g] = p0;
    }
  std::vector<std::pair<int32_t, int32_t>> order;   // (splitting pixel, segment)
  for (int g = 1; g <= count; g++) if (winner[g] >= 0) order.emplace_back(winner[g], g);
  std::sort(order.begin(), order.end());
  for (const auto &pg : order) {
    const int p0 = pg.first, g = pg.second;
    const int x = p0 % iw, y = p0 / iw, n = numberin[p0];
    const LS_t *gp = old;
    if (gp[g].endIndex - gp[g].startIndex < MINNINDEX - 1) continue;
    if (gp[g].startCount > 1 || gp[g].endCount > 1) continue;
    int maxDist = gp[g].maxDist;
    if (maxDist < ((int)(minerror * 65536))) continue;
    if ((float)maxDist < (minerror * 3 * 65536) &&
        (float)maxDist * maxDist / distanceSqu(gp[g].x0, gp[g].y0, gp[g].x1, gp[g].y1) < 100000.0f) continue;
    if (distanceSqu((float)x, (float)y, gp[g].x0, gp[g].y0) < (MINEDGELEN * MINEDGELEN)) continue;
    if (distanceSqu((float)x, (float)y, gp[g].x1, gp[g].y1) < (MINEDGELEN * MINEDGELEN)) continue;
    int gr = gp[g].rightPtr;
    int32_t *cnt = (int32_t *)nw;
    int gn = (*cnt)++ + 1;
    if (ls_overflow(gn, lsListSize)) { g_stats.ls_overflow++; continue; }
    nw[gn].startIndex = n;
    nw[gn].endIndex = gp[g].endIndex;
    nw[gn].x0 = (float)x;
    nw[gn].y0 = (float)y;
    nw[gn].x1 = gp[g].x1;
    nw[gn].y1 = gp[g].y1;
    nw[gn].leftPtr = g;
    nw[gn].rightPtr = gp[g].rightPtr;
    nw[gn].maxDist = 0;
    nw[gn].polyid = gp[g].polyid;
    nw[gn].level = maxDist;

    nw[g].endIndex = n;
    nw[g].x1 = (float)x;
    nw[g].y1 = (float)y;
    nw[g].rightPtr = gn;
    nw[g].maxDist = 0;

    if (gr != 0) nw[gr].leftPtr = gn;
  }
}

// ---- oclpolyline.cl:618-646 ----
static int k_mkpl_pass3(const LS_t *gp, int lsListSize, const int32_t *numberin, int32_t *labelinout, int32_t *flags, int nIter, int iw, int ih) {
  if (flags[nIter - 1] == 0) return 0;
  int moved = 0;
  for (int p0 = 0; p0 < iw * ih; p0++) {
    int g = labelinout[p0];
    if (g == 0) continue;
    if (ls_overflow(g, lsListSize)) continue;
    if (gp[g].polyid == 0) continue;
    int n = numberin[p0];
    if (gp[g].endIndex < n) {
      labelinout[p0] = gp[g].rightPtr;
      flags[nIter] = 1;
      moved = 1;
    }
  }
  return moved;
}

// ---- oclpolyline.c:186-216 ----
static void mkpl(LS_t *lsList, int32_t *tmpBig, int lsListSize, int32_t *labelinout, const int32_t *numberin, int32_t *randtmp, int32_t *tmp2,
                 int32_t *flags, float minerror, int iw, int ih) {
  const int N = 16;
  k_clear((int32_t *)lsList, (lsListSize + 3) / 4);
  k_mkpl_pass0a(lsList, lsListSize, numberin, labelinout, flags, N, iw, ih);
  k_mkpl_pass0b(lsList, lsListSize, numberin, labelinout, iw, ih);
  k_rand(randtmp, 0, iw * ih);
  for (int i = 0; i < N - 1; i++) {
    k_mkpl_pass1(lsList, lsListSize, tmp2, labelinout, randtmp, flags, i + 1, iw, ih);
    k_copy(tmpBig, (const int32_t *)lsList, (lsListSize + 3) / 4);
    k_mkpl_pass2(lsList, (const LS_t *)tmpBig, lsListSize, tmp2, numberin, labelinout, flags, i + 1, minerror, iw, ih);
    if (k_mkpl_pass3(lsList, lsListSize, numberin, labelinout, flags, i + 1, iw, ih)) g_stats.mkpl_iterations_live = i + 1;
  }
}

// ---- oclpolyline.cl:680-700 ; Q11: one work-item per list entry ----
static void k_refine_pass0(LSX_t *lsx, const LS_t *ls) {
  const int count = *(const int32_t *)ls;
  for (int g = 1; g <= count; g++) {
    if (ls[g].polyid == 0) continue;
    lsx[g].dirSEx = (int16_t)(ls[g].x1 - ls[g].x0);     // convert_short2: truncation (Q18)
    lsx[g].dirSEy = (int16_t)(ls[g].y1 - ls[g].y0);
    lsx[g].vDirSEx = (int16_t)(-lsx[g].dirSEy);
    lsx[g].vDirSEy = lsx[g].dirSEx;
    lsx[g].mx00 = lsx[g].mx01 = lsx[g].mx11 = lsx[g].my0 = lsx[g].my1 = 0;
    lsx[g].distSquSE = lsx[g].dirSEx * lsx[g].dirSEx + lsx[g].dirSEy * lsx[g].dirSEy;
    lsx[g].padding = 0;
  }
}

// ---- oclpolyline.cl:715-750 : 64-bit integer moment sums (order independent) ----
static void k_refine_pass1(LSX_t *lsx, const LS_t *ls, const int32_t *lsIdIn, int iw, int ih) {
  const int count = *(const int32_t *)ls;
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      int g = lsIdIn[p0];
      if (g == 0) continue;
      if (g < 0 || count < g) { g_stats.ls_overflow++; continue; }
      int vx = x - (int)rintf(ls[g].x0), vy = y - (int)rintf(ls[g].y0);      // convert_int2_rte
      int ay = vx * (int)lsx[g].vDirSEx + vy * (int)lsx[g].vDirSEy;
      int ax0 = vx * (int)lsx[g].dirSEx + vy * (int)lsx[g].dirSEy;
      int ax1 = lsx[g].distSquSE;
      lsx[g].mx00 += llrintf((float)ax0 * (float)ax0);                          // convert_long_rte
      lsx[g].mx01 += llrintf((float)ax0 * (float)ax1);
      lsx[g].mx11 += llrintf((float)ax1 * (float)ax1);
      lsx[g].my0 += llrintf((float)ax0 * (float)ay);
      lsx[g].my1 += llrintf((float)ax1 * (float)ay);
    }
}

// ---- oclpolyline.cl:752-770 ----
static void k_refine_pass2(const LSX_t *lsx, LS_t *ls) {
  const int count = *(const int32_t *)ls;
  for (int g = 1; g <= count; g++) {
    if (ls[g].polyid == 0) continue;
    float rdet = (float)lsx[g].mx00 * (float)lsx[g].mx11 - (float)lsx[g].mx01 * (float)lsx[g].mx01;
    if (rdet == 0) continue;
    rdet = (float)(1.0 / (double)rdet);                                         // Q15: `1.0 / rdet` is a double division
    float as0 = ((float)lsx[g].mx11 * (float)lsx[g].my0 - (float)lsx[g].mx01 * (float)lsx[g].my1) * rdet;
    float as1 = ((float)lsx[g].mx00 * (float)lsx[g].my1 - (float)lsx[g].mx01 * (float)lsx[g].my0) * rdet;
    ls[g].x0 += (float)lsx[g].vDirSEx * as1;
    ls[g].y0 += (float)lsx[g].vDirSEy * as1;
    ls[g].x1 += (float)lsx[g].vDirSEx * (as0 + as1);
    ls[g].y1 += (float)lsx[g].vDirSEy * (as0 + as1);
  }
}

// ---- oclpolyline.cl:772-809.  The kernel rewrites the vertex a segment shares with its right neighbour IN PLACE, so what a
// work-item reads depends on which neighbours ran before it.  CANONICAL (Q5): the work-items in id order (the schedule of
// oracle/_ref/librd_ref.so): g sees its own start as already moved by its left neighbour l iff l < g, and its right
// neighbour's end as already moved iff h < g. ----
static void k_refine_pass3(LS_t *ls) {
  const int count = *(const int32_t *)ls;
  for (int g = 1; g <= count; g++) {
    if (ls[g].polyid == 0) continue;
    const int h = ls[g].rightPtr;
    if (h == 0) continue;
    float v0 = ls[g].x0, v1 = ls[g].y0, v2 = ls[g].x1, v3 = ls[g].y1;
    float u0 = ls[h].x0, u1 = ls[h].y0, u2 = ls[h].x1, u3 = ls[h].y1;
    float d = (v2 - v0) * (u3 - u1) - (v3 - v1) * (u2 - u0);
    float mx = (v2 + u0) * 0.5f, my = (v3 + u1) * 0.5f;
    if ((double)fabsf(d) < 1e-6) {                                              // Q15
      ls[g].x1 = ls[h].x0 = mx; ls[g].y1 = ls[h].y0 = my;
      continue;
    }
    float n = (v1 - u1) * (u2 - u0) - (v0 - u0) * (u3 - u1);
    float q = n / d;
    float wx = v0 + q * (v2 - v0), wy = v1 + q * (v3 - v1);
    if (hypot_c(wx - v2, wy - v3) > 10 && hypot_c(wx - u0, wy - u1) > 10) {
      ls[g].x1 = ls[h].x0 = mx; ls[g].y1 = ls[h].y0 = my;
      continue;
    }
    ls[g].x1 = ls[h].x0 = wx; ls[g].y1 = ls[h].y0 = wy;
  }
}

}  // namespace ora

using namespace ora;

extern "C" {

// oclpolyline.c:218-309.  Step numbers are those of SURVEY.md section 10.2.
void ora_polyline_execute(ora_ls_t *lsList, int lsListSize, int32_t *lsIdOut, const int32_t *in, int32_t *tmpBig,
                          int32_t *tmp0, int32_t *tmp1, int32_t *tmp2, int32_t *tmp3, int32_t *tmp4, int32_t *tmp5,
                          float minerror, int sizeThre, int iw, int ih, int stop_step) {
#define STEP(k) do { if (stop_step == (k)) return; } while (0)
  const int n = iw * ih;
  // step 1
  k_simpleJunction(lsIdOut, in, iw, ih);
  k_simpleConnect(tmp2, lsIdOut, iw, ih);
  k_stringify(tmp1, tmp2, 0, iw, ih);
  k_stringify(tmp2, tmp1, 1, iw, ih);
  k_removeBranch(tmp1, tmp2, iw, ih);
  STEP(1);
  // step 2
  label8x(lsIdOut, tmp1, tmp2, 0, iw, ih);
  STEP(2);
  // step 3
  k_simpleJunction(tmp2, tmp1, iw, ih);
  k_clear(tmp3, n);
  k_countEnds(tmp3, tmp2, lsIdOut, iw, ih);
  k_breakLoops(tmp1, lsIdOut, tmp3, iw, ih);
  STEP(3);
  // step 4
  k_findEnds0(tmp0, tmp2, tmpBig, lsIdOut, iw, ih);
  STEP(4);
  // step 5
  k_findEnds1(tmp3, tmp4, tmpBig, tmp0, tmp2, lsIdOut, 0, iw, ih);
  k_findEnds1(tmp0, tmp2, tmpBig, tmp3, tmp4, lsIdOut, 1, iw, ih);
  k_findEnds1(tmp3, tmp4, tmpBig, tmp0, tmp2, lsIdOut, 0, iw, ih);
  k_findEnds1(tmp0, tmp2, tmpBig, tmp3, tmp4, lsIdOut, 1, iw, ih);
  STEP(5);
  // step 6
  k_findEnds2(tmpBig, tmp4, tmp0, tmp2, lsIdOut, iw, ih);
  STEP(6);
  // step 7
  k_number(tmp2, tmp3, tmpBig, tmp4, iw, ih);
  k_number(tmpBig, tmp4, tmp2, tmp3, iw, ih);
  k_number(tmp2, tmp3, tmpBig, tmp4, iw, ih);
  STEP(7);
  // step 8
  k_copy(tmp1, tmp2, n);
  labelpl(tmpBig, tmp1, tmp3, iw, ih);
  STEP(8);
  // step 9
  k_clear(tmp1, n);
  k_calcSize(tmp1, tmpBig, iw, ih);
  k_filterSize(lsIdOut, tmpBig, tmp1, sizeThre, iw, ih);
  STEP(9);
  // step 10 : the clear is launched over 4x the items but its size argument is iw*ih ints (oclpolyline.c:290-291)
  k_clear(tmpBig, n);
  k_relabel_pass0(tmpBig, lsIdOut, iw, ih);
  k_relabel_pass1(lsIdOut, tmpBig, iw, ih);
  STEP(10);
  // step 11
  mkpl(lsList, tmpBig, lsListSize, lsIdOut, tmp2, tmp5, tmp3, tmp4, minerror, iw, ih);
  STEP(11);
  // step 12
  k_refine_pass0((LSX_t *)tmpBig, lsList);
  k_refine_pass1((LSX_t *)tmpBig, lsList, lsIdOut, iw, ih);
  k_refine_pass2((const LSX_t *)tmpBig, lsList);
  k_refine_pass3(lsList);
  g_stats.n_ls = *(const int32_t *)lsList;
#undef STEP
}

}  // extern "C"
