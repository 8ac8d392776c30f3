// ora_rect.cpp - CPU oracle, Stage B + D: restatement of oclrect.cl and of genGPUTask (oclrect.c:235-381).
// TEST INFRASTRUCTURE ONLY (see rd_oracle.h).
#include <vector>
#include "ora_internal.h"

namespace ora {

// ---- oclrect.cl:74-95 ----
static void k_simpleJunction(int32_t *out, const int32_t *in, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      out[p0] = 0;
      if (x <= 0 || y <= 0 || x >= (iw - 1) || y >= (ih - 1)) continue;
      if (!(in[p0] > 0)) continue;
      int count = 1;
      for (int i = 0; i < 8; i++)
        if (in[p0 + RX[i] + RY[i] * iw] > 0) count++;
      out[p0] = count == 1 ? 0 : count;
    }
}

// ---- oclrect.cl:97-121 ----
static void k_simpleConnect(int32_t *out, const int32_t *in, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      out[p0] = 0;
      if (x <= 1 || y <= 1 || x >= (iw - 2) || y >= (ih - 2)) continue;
      out[p0] = in[p0] != 0 ? 1 : 0;
      if (in[p0] != 0) continue;
      if (in[p0 - 1] == 2 && in[p0 + 1] != 0) out[p0] = 1;
      if (in[p0 - 1] != 0 && in[p0 + 1] == 2) out[p0] = 1;
      if (in[p0 - iw] == 2 && in[p0 + iw] != 0) out[p0] = 1;
      if (in[p0 - iw] != 0 && in[p0 + iw] == 2) out[p0] = 1;
      if (in[p0 - iw - 1] == 2 && in[p0 + iw + 1] == 2) out[p0] = 1;
      if (in[p0 - iw + 1] == 2 && in[p0 + iw - 1] == 2) out[p0] = 1;
      if (in[p0 + 1] == 2 && in[p0 + iw - 1] == 2) out[p0] = 1;
      if (in[p0 - 1] == 2 && in[p0 + iw + 1] == 2) out[p0] = 1;
      if (in[p0 - iw + 1] == 2 && in[p0 + iw] == 2) out[p0] = 1;
      if (in[p0 - iw - 1] == 2 && in[p0 + iw] == 2) out[p0] = 1;
    }
}

// ---- oclrect.cl:123-135 ----
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

// ---- oclrect.cl:38-48 ----
static inline uint32_t packlabbl(int l, int a, int b) {
  uint32_t ret = (uint32_t)clampi(b, 0, 1023);
  ret = (ret << 10) | (uint32_t)clampi(a, 0, 1023);
  ret = (ret << 12) | (uint32_t)clampi(l, 0, 4095);
  return ret;
}

#define BLBLURSIZE 4

// ---- oclrect.cl:155-179 ----
static void k_blblur0(uint32_t *out, const int8_t *edge, const uint32_t *in, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      int wsum = 0, oe = edge[y * iw + x] != 0;
      int c0 = 0, c1 = 0, c2 = 0;
      for (int xx = 0; xx >= -BLBLURSIZE; xx--) {
        if (x + xx < 0) break;
        if (x + xx > 0 && edge[y * iw + x + xx] != 0 && edge[y * iw + x + xx - 1] == 0) break;
        if (x + xx > 0 && y < ih - 1 && edge[y * iw + x + xx] == 0 && edge[y * iw + x + xx - 1] != 0 && edge[(y + 1) * iw + x + xx] != 0) break;
        wsum++;
        uint32_t v = in[y * iw + x + xx];
        c0 += v & 4095; c1 += (v >> 12) & 1023; c2 += (v >> 22) & 1023;
      }
      for (int xx = 0; xx <= BLBLURSIZE; xx++) {
        if (x + xx > iw - 1) break;
        if (x + xx < iw - 1 && edge[y * iw + x + xx] == 0 && edge[y * iw + x + xx + 1] != 0) break;
        if (oe && edge[y * iw + x + xx] == 0) break;
        wsum++;
        uint32_t v = in[y * iw + x + xx];
        c0 += v & 4095; c1 += (v >> 12) & 1023; c2 += (v >> 22) & 1023;
      }
      out[y * iw + x] = wsum == 0 ? in[y * iw + x] : packlabbl(c0 / wsum, c1 / wsum, c2 / wsum);
    }
}

// ---- oclrect.cl:181-205 ----
static void k_blblur1(uint32_t *out, const int8_t *edge, const uint32_t *in, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      int wsum = 0, oe = edge[y * iw + x] != 0;
      int c0 = 0, c1 = 0, c2 = 0;
      for (int yy = 0; yy >= -BLBLURSIZE; yy--) {
        if (y + yy < 0) break;
        if (y + yy > 0 && edge[(y + yy) * iw + x] != 0 && edge[(y + yy - 1) * iw + x] == 0) break;
        if (y + yy > 0 && x < iw - 1 && edge[(y + yy) * iw + x] == 0 && edge[(y + yy - 1) * iw + x] != 0 && edge[(y + yy) * iw + x + 1] != 0) break;
        wsum++;
        uint32_t v = in[(y + yy) * iw + x];
        c0 += v & 4095; c1 += (v >> 12) & 1023; c2 += (v >> 22) & 1023;
      }
      for (int yy = 0; yy <= BLBLURSIZE; yy++) {
        if (y + yy > ih - 1) break;
        if (y + yy < ih - 1 && edge[(y + yy) * iw + x] == 0 && edge[(y + yy + 1) * iw + x] != 0) break;
        if (oe && edge[(y + yy) * iw + x] == 0) break;
        wsum++;
        uint32_t v = in[(y + yy) * iw + x];
        c0 += v & 4095; c1 += (v >> 12) & 1023; c2 += (v >> 22) & 1023;
      }
      out[y * iw + x] = wsum == 0 ? in[y * iw + x] : packlabbl(c0 / wsum, c1 / wsum, c2 / wsum);
    }
}

// ---- oclrect.cl:207-216 ----
static void k_quantize(uint32_t *out, const uint32_t *in, int n0, int n1, int n2, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int p0 = 0; p0 < iw * ih; p0++) {
    float l, a, b;
    unpacklab(in[p0], l, a, b);
    out[p0] = packlab(roundf(l * n0) / (float)n0, roundf(a * n1) / (float)n1, roundf(b * n2) / (float)n2);
  }
}

// ---- oclrect.cl:218-244 ----
static void k_despeckle(uint32_t *out, const uint32_t *in, const float *edge, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      out[p0] = in[p0];
      if (edge[p0] < 1e-6f) continue;
      float dist = 1e+10f;
      float l0, a0, b0;
      unpacklab(in[p0], l0, a0, b0);
      for (int yy = -1; yy <= 1; yy++)
        for (int xx = -1; xx <= 1; xx++)
          if (0 <= x + xx && x + xx < iw && 0 <= y + yy && y + yy < ih) {
            const int p1 = (y + yy) * iw + x + xx;
            if (edge[p1] >= 1e-6f) continue;
            float l1, a1, b1;
            unpacklab(in[p1], l1, a1, b1);
            float d = distance3_c(l1 - l0, a1 - a0, b1 - b0);
            if (d < dist) { out[p0] = in[p1]; dist = d; }
          }
    }
}

// ---- oclrect.cl:246-287 : scatter of constants, order independent ----
static void k_mkMergeMask0(int32_t *out, const int32_t *junctionIn, int iw, int ih) {
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      if (junctionIn[y * iw + x] == 0) continue;
      for (int yy = y - 6; yy <= y + 6; yy++)
        for (int xx = x - 6; xx <= x + 6; xx++) {
          if (xx < 0 || iw <= xx || yy < 0 || ih <= yy) continue;
          int dsqu = (yy - y) * (yy - y) + (xx - x) * (xx - x);
          if (16 <= dsqu && dsqu < 36) out[yy * iw + xx] = 1;
        }
    }
}

static void k_mkMergeMask1(int32_t *inout, const int32_t *junctionIn, int iw, int ih) {
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int j = junctionIn[y * iw + x];
      if (j == 2) {
        for (int yy = y - 8; yy <= y + 8; yy++)
          for (int xx = x - 8; xx <= x + 8; xx++) {
            if (xx < 0 || iw <= xx || yy < 0 || ih <= yy) continue;
            int dsqu = (yy - y) * (yy - y) + (xx - x) * (xx - x);
            if (dsqu < 64) inout[yy * iw + xx] = 0;
          }
      } else if (j != 0) {
        for (int yy = y - 4; yy <= y + 4; yy++)
          for (int xx = x - 4; xx <= x + 4; xx++) {
            if (xx < 0 || iw <= xx || yy < 0 || ih <= yy) continue;
            int dsqu = (yy - y) * (yy - y) + (xx - x) * (xx - x);
            if (dsqu < 16) inout[yy * iw + xx] = 0;
          }
      }
    }
}

// ---- oclrect.cl:289-334 + oclrect.c:325-331 : labelxPreprocess + 8 x labelMergeMain ----
// The reference's result is schedule dependent: (i) 8 in-place passes need not converge, (ii) a pixel adopts a neighbour's label only
// if that label is CURRENTLY smaller, and the test `(pix equal || mask[adopter])` is asymmetric, so which trees merge depends on
// the order of the work-items and on transient pointer values, (iii) pixels of the 1-px image border never run the main pass.
// CANONICAL (Q6'): a deterministic fixed point of the same rule.  For a 4-neighbour pair (a, b), b = a+1 or a+iw, with edge[b] <= 0:
//     b may adopt from a  iff  b is not on the image border and (pix[a] == pix[b] || mask[b] != 0)
//     a may adopt from b  iff  a is not on the image border and (pix[a] == pix[b] || mask[a] != 0)
//   - the labelxPreprocess links (every pixel: up neighbour if same colour, else left if same colour) and the pairs that may adopt in
//     BOTH directions are united unconditionally (whatever the order, one of the two labels is the smaller one);
//   - a pair that may adopt in ONE direction only is united when the source's component label (smallest index) is smaller than
//     the adopter's - the reference's `s < g` - evaluated for all such pairs at once on the labels of the round before, for at
//     most ORA_MERGE_ROUNDS rounds (the second one has never enabled anything on the frames of the sweeps: it confirms the fixed point);
// pixels not on the image border get the smallest index of their component, image-border pixels keep their labelxPreprocess value
// except on the top row: a top-row pixel whose lower neighbour has its colour and is no edge pixel lands on the start of its run
// of equal colours (the reference's first pass, row 1 adopting from row 0 left to right, flattens the row-0 chains that way).
// Against the reference's raster-order run (tests/test_ref_device.py, profiles/r04r_*): 0-220 interior pixels of a 640x480 frame
// differ (round 1's rule - unite every pair that may adopt in at least one direction - 58-460, always a coarsening).
#define ORA_MERGE_ROUNDS 2
// The first pass of labelMergeMain, replayed in raster order exactly as the reference's kernel runs it (oclrect.cl:300-334 after
// labelxPreprocess :289-298): image-frame pixels are skipped, a pixel compares the RAW current labels of its neighbours, follows the
// pointers eight times and lowers label[og] and its own label.
static void merge_first_pass(int32_t *label, const int32_t *pix, const int32_t *mask, const int32_t *edge, int iw, int ih) {
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      if (y > 0 && pix[p0] == pix[p0 - iw]) label[p0] = p0 - iw;
      else if (x > 0 && pix[p0] == pix[p0 - 1]) label[p0] = p0 - 1;
      else label[p0] = p0;
    }
  for (int y = 1; y < ih - 1; y++)
    for (int x = 1; x < iw - 1; x++) {
      const int p0 = y * iw + x;
      int g = label[p0];
      const int og = g;
      const bool m = mask[p0] != 0;
      int p1 = p0 - iw, s = label[p1];
      if (s < g && (pix[p0] == pix[p1] || m) && edge[p0] <= 0) g = s;
      p1 = p0 - 1; s = label[p1];
      if (s < g && (pix[p0] == pix[p1] || m) && edge[p0] <= 0) g = s;
      p1 = p0 + 1; s = label[p1];
      if (s < g && (pix[p0] == pix[p1] || m) && edge[p1] <= 0) g = s;
      p1 = p0 + iw; s = label[p1];
      if (s < g && (pix[p0] == pix[p1] || m) && edge[p1] <= 0) g = s;
      for (int j = 0; j < 8; j++) g = label[g];
      if (g != og) {
        if (g < label[og]) label[og] = g;
        if (g < label[p0]) label[p0] = g;
      }
    }
}
// ---- the two forms of the merge labelling.  ora_set_merge_replay(0) (default): the schedule-independent fixed point alone, seeded
// with the preprocess pointers - what the CUDA path computes by default.  ora_set_merge_replay(1): the first pass replayed exactly,
// then the fixed point - the CUDA path under RD_MERGE_REPLAY=1 / rd_set_merge_replay(1). ----
static int g_merge_replay = 0;
static void labelMerge_fixed_point(int32_t *label, const int32_t *pix, const int32_t *mask, const int32_t *edge, int iw, int ih) {
  const int n = iw * ih;
  std::vector<int32_t> init(n);
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      if (y > 0 && pix[p0] == pix[p0 - iw]) init[p0] = p0 - iw;
      else if (x > 0 && pix[p0] == pix[p0 - 1]) init[p0] = p0 - 1;
      else init[p0] = p0;
    }
  for (int p = 0; p < n; p++) label[p] = p;
  MinUF uf(label);
  for (int p = 0; p < n; p++) if (init[p] != p) uf.unite(p, init[p]);
  auto interior = [&](int x, int y) { return x > 0 && y > 0 && x < iw - 1 && y < ih - 1; };
  std::vector<std::pair<int, int>> dir;                    // (adopter, source)
  auto pair_ab = [&](int a, int b, bool ia, bool ib) {
    if (!(edge[b] <= 0)) return;
    const bool same = pix[a] == pix[b];
    const bool b_from_a = ib && (same || mask[b] != 0), a_from_b = ia && (same || mask[a] != 0);
    if (b_from_a && a_from_b) uf.unite(a, b);
    else if (b_from_a) dir.push_back({b, a});
    else if (a_from_b) dir.push_back({a, b});
  };
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int a = y * iw + x;
      if (x + 1 < iw) pair_ab(a, a + 1, interior(x, y), interior(x + 1, y));
      if (y + 1 < ih) pair_ab(a, a + iw, interior(x, y), interior(x, y + 1));
    }
  std::vector<int32_t> root(n);
  for (int round = 0; round < ORA_MERGE_ROUNDS; round++) {
    for (int p = 0; p < n; p++) root[p] = uf.find_compress(p);
    std::vector<size_t> en;
    for (size_t i = 0; i < dir.size(); i++) if (root[dir[i].second] < root[dir[i].first]) en.push_back(i);
    if (en.empty()) break;
    for (size_t i : en) uf.unite(dir[i].first, dir[i].second);
  }
  for (int p = 0; p < n; p++) root[p] = uf.find_compress(p);
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      label[p0] = interior(x, y) ? root[p0] : init[p0];
    }
  // The top row.  An interior pixel q = (x, 1) with the colour of p = (x, 0) starts out pointing at p; when it first adopts - in the
  // reference's first pass, from p itself, whose preprocess label is its left neighbour - it chases the pointers along the top row
  // to the start of p's run of equal colours and drags p along (atomic_min(&label[og], g), og = p).  So in the reference's raster
  // run the top-row pixels sit on the START OF THEIR RUN, not on their left neighbour (99.6 % of the image-frame labels of the
  // reference follow this rule, 70 % the plain preprocess rule).  Left / right / bottom frame pixels are nobody's first pointer.
  int start = 0;
  for (int x = 0; x < iw; x++) {
    if (x == 0 || pix[x] != pix[x - 1]) start = x;
    if (x >= 1 && x < iw - 1 && ih > 2 && start != x && pix[iw + x] == pix[x] && edge[iw + x] <= 0) label[x] = start;
  }
}

// the fixed point seeded with a label plane `first` (pointers towards smaller indices): the plane after the first pass in the replay
// mode; any later state of the reference's label plane in tools/merge_seed_experiment.py
static void merge_fixed_point_from(int32_t *label, const std::vector<int32_t> &first, const int32_t *pix, const int32_t *mask, const int32_t *edge, int iw, int ih);
static void labelMerge_replay(int32_t *label, const int32_t *pix, const int32_t *mask, const int32_t *edge, int iw, int ih) {
  std::vector<int32_t> first((size_t)iw * ih);
  merge_first_pass(first.data(), pix, mask, edge, iw, ih);
  merge_fixed_point_from(label, first, pix, mask, edge, iw, ih);
}
static void merge_fixed_point_from(int32_t *label, const std::vector<int32_t> &first, const int32_t *pix, const int32_t *mask, const int32_t *edge, int iw, int ih) {
  const int n = iw * ih;
  for (int p = 0; p < n; p++) label[p] = p;
  MinUF uf(label);
  for (int p = 0; p < n; p++) if (first[p] != p) uf.unite(p, first[p]);
  auto interior = [&](int x, int y) { return x > 0 && y > 0 && x < iw - 1 && y < ih - 1; };
  std::vector<std::pair<int, int>> dir;                    // (adopter, source)
  auto pair_ab = [&](int a, int b, bool ia, bool ib) {
    if (!(edge[b] <= 0)) return;
    const bool same = pix[a] == pix[b];
    const bool b_from_a = ib && (same || mask[b] != 0), a_from_b = ia && (same || mask[a] != 0);
    if (b_from_a && a_from_b) uf.unite(a, b);
    else if (b_from_a) dir.push_back({b, a});
    else if (a_from_b) dir.push_back({a, b});
  };
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int a = y * iw + x;
      if (x + 1 < iw) pair_ab(a, a + 1, interior(x, y), interior(x + 1, y));
      if (y + 1 < ih) pair_ab(a, a + iw, interior(x, y), interior(x, y + 1));
    }
  std::vector<int32_t> root(n);
  for (int round = 0; round < ORA_MERGE_ROUNDS; round++) {
    for (int p = 0; p < n; p++) root[p] = uf.find_compress(p);
    std::vector<size_t> en;
    for (size_t i = 0; i < dir.size(); i++) if (root[dir[i].second] < root[dir[i].first]) en.push_back(i);
    if (en.empty()) break;
    for (size_t i : en) uf.unite(dir[i].first, dir[i].second);
  }
  for (int p = 0; p < n; p++) root[p] = uf.find_compress(p);
  // interior pixels: the smallest index of their component.  Image-frame pixels never run the main pass; they keep what the first
  // pass left in them, except that one that was still the root of its tree after the first pass has been hooked under a smaller
  // root since: the smallest index of its component.
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      label[p0] = (interior(x, y) || first[p0] == p0) ? root[p0] : first[p0];
    }
}

static void labelMerge(int32_t *label, const int32_t *pix, const int32_t *mask, const int32_t *edge, int iw, int ih) {
  if (g_merge_replay) labelMerge_replay(label, pix, mask, edge, iw, ih);
  else labelMerge_fixed_point(label, pix, mask, edge, iw, ih);
}

// ---- oclrect.cl:336-346 ----
static void k_calcSize(int32_t *out, const int32_t *label, int iw, int ih) {
  for (int p0 = 0; p0 < iw * ih; p0++) {
    int l = label[p0];
    if (l != -1) out[l]++;
  }
}

// ---- oclrect.cl:348-371.  The kernel updates the labels in place, so its outcome depends on the order of the work-items.
// CANONICAL (Q3): raster order - the schedule of the reference run here (oracle/_ref/librd_ref.so), bit-identical to it
// (tests/test_ref_device.py).  A pixel sees the new labels of its NW / N / NE / W neighbours and the old ones elsewhere. ----
static void k_despeckle2(int32_t *labelinout, const int32_t *sizein, int thre, int iw, int ih) {
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      if (sizein[labelinout[p0]] > thre) continue;
      int maxSize = 0, maxLabel = labelinout[p0];
      for (int yy = -1; yy <= 1; yy++)
        for (int xx = -1; xx <= 1; xx++)
          if (0 <= x + xx && x + xx < iw && 0 <= y + yy && y + yy < ih) {
            const int p1 = (y + yy) * iw + x + xx;
            if (sizein[labelinout[p1]] > maxSize) { maxSize = sizein[labelinout[p1]]; maxLabel = labelinout[p1]; }
          }
      labelinout[p0] = maxLabel;
    }
}

// ---- oclrect.cl:373-390 (the `edge` argument is unused by the kernel) ----
static void k_markBoundary(int32_t *out, const int32_t *in, int iw, int ih) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p0 = y * iw + x;
      if (x <= 1 || y <= 1 || x >= iw - 2 || y >= ih - 2) { out[p0] = -1; continue; }
      int nearEdge = 0;
      const int c0 = in[p0];
      for (int yy = -2; yy <= 2; yy++)
        for (int xx = -2; xx <= 2; xx++)
          if (in[p0 + yy * iw + xx] != c0) nearEdge = 1;
      out[p0] = nearEdge ? in[p0] : -1;
    }
}

// ---- oclrect.cl:427-464 : the vote table.  slot = ((lsid*bid) & 0x7fffffff) % nentry, no probing ----
// CANONICAL (Q19): the reference's own rule with its work-items in raster order (the schedule of oracle/_ref/librd_ref.so):
// the first pixel in raster order that hits an empty slot claims it for its lsid (the reference: whoever's
// atomic_cmpxchg lands first), and the very hit that performs the claim is NOT recorded (oclrect.cl:449-456: the
// cmpxchg returns 0, which is not lsid) - so the claiming pixel counts only if its window hits the slot again.
static void k_reduceLS(int32_t *out, const int32_t *boundaryin, const int32_t *lsidin, int iw, int ih, int nentry) {
  for (int y = 1; y < ih - 1; y++)
    for (int x = 1; x < iw - 1; x++) {
      const int lsid = lsidin[y * iw + x];
      if (lsid <= 0) continue;
      for (int yy = -3; yy <= 3; yy++) {
        if (y + yy < 0 || ih <= y + yy) continue;
        for (int xx = -3; xx <= 3; xx++) {
          if (x + xx < 0 || iw <= x + xx) continue;
          const int bid = boundaryin[(y + yy) * iw + x + xx];
          if (bid <= 0) continue;
          const int hash = (int)((((unsigned)lsid * (unsigned)bid) & 0x7fffffffu) % (unsigned)nentry);
          int *e = &out[hash * 5];
          if (e[0] == 0) { e[0] = lsid; g_stats.vote_slots++; continue; }
          if (e[0] != lsid) { g_stats.vote_collisions++; continue; }
          if (iw - x > e[1]) e[1] = iw - x;
          if (x > e[2]) e[2] = x;
          if (ih - y > e[3]) e[3] = ih - y;
          if (y > e[4]) e[4] = y;
        }
      }
    }
}

}  // namespace ora

using namespace ora;

// =====================================================================================================
// L3 object: same 6+6+2 planes and 2 bigs as oclrect_t (oclrect.c:51-53, 120-135), zero-initialised.
// CANONICAL (Q1): device memory the reference never initialises reads as zero on the first frame.
// =====================================================================================================
struct ora_rect {
  int iw, ih;
  int32_t *buf[6], *tmp[6], *iobuf[2], *ioBig[2];
  double t[5];
};

extern "C" {

void ora_rect_simpleJunction(int32_t *out, const int32_t *in, int iw, int ih) { k_simpleJunction(out, in, iw, ih); }
void ora_rect_simpleConnect(int32_t *out, const int32_t *in, int iw, int ih) { k_simpleConnect(out, in, iw, ih); }
void ora_rect_stringify(int32_t *out, const int32_t *in, int mod2, int iw, int ih) { k_stringify(out, in, mod2, iw, ih); }
void ora_rect_blblur0(uint32_t *out, const int8_t *edge, const uint32_t *in, int iw, int ih) { k_blblur0(out, edge, in, iw, ih); }
void ora_rect_blblur1(uint32_t *out, const int8_t *edge, const uint32_t *in, int iw, int ih) { k_blblur1(out, edge, in, iw, ih); }
void ora_rect_quantize(uint32_t *out, const uint32_t *in, int n0, int n1, int n2, int iw, int ih) { k_quantize(out, in, n0, n1, n2, iw, ih); }
void ora_rect_despeckle(uint32_t *out, const uint32_t *in, const float *edge, int iw, int ih) { k_despeckle(out, in, edge, iw, ih); }
void ora_rect_mkMergeMask0(int32_t *out, const int32_t *junction, int iw, int ih) { k_mkMergeMask0(out, junction, iw, ih); }
void ora_rect_mkMergeMask1(int32_t *inout, const int32_t *junction, int iw, int ih) { k_mkMergeMask1(inout, junction, iw, ih); }
/* label holds a label plane of the reference on entry (after any of its passes), the fixed point seeded with it on return */
void ora_rect_labelMerge_seeded(int32_t *label, const int32_t *pix, const int32_t *mask, const int32_t *edge, int iw, int ih) {
  std::vector<int32_t> first(label, label + (size_t)iw * ih);
  merge_fixed_point_from(label, first, pix, mask, edge, iw, ih);
}
void ora_set_merge_replay(int on) { g_merge_replay = on != 0; }
int ora_get_merge_replay(void) { return g_merge_replay; }
void ora_rect_labelMerge_first_pass(int32_t *label, const int32_t *pix, const int32_t *mask, const int32_t *edge, int iw, int ih) { merge_first_pass(label, pix, mask, edge, iw, ih); }
void ora_rect_labelMerge(int32_t *label, const int32_t *pix, const int32_t *mask, const int32_t *edge, int iw, int ih) { labelMerge(label, pix, mask, edge, iw, ih); }
void ora_rect_calcSize(int32_t *out, const int32_t *label, int iw, int ih) { k_calcSize(out, label, iw, ih); }
void ora_rect_despeckle2(int32_t *labelinout, const int32_t *size, int thre, int iw, int ih) { k_despeckle2(labelinout, size, thre, iw, ih); }
void ora_rect_markBoundary(int32_t *out, const int32_t *in, int iw, int ih) { k_markBoundary(out, in, iw, ih); }
void ora_rect_reduceLS(int32_t *out, const int32_t *boundary, const int32_t *lsid, int iw, int ih, int nentry) { k_reduceLS(out, boundary, lsid, iw, ih, nentry); }

ora_rect *ora_rect_create(int iw, int ih) {
  ora_rect *o = (ora_rect *)calloc(1, sizeof(ora_rect));
  o->iw = iw; o->ih = ih;
  const size_t P = (size_t)iw * ih * 4;
  for (int i = 0; i < 6; i++) { o->buf[i] = (int32_t *)calloc(1, P); o->tmp[i] = (int32_t *)calloc(1, P); }
  for (int i = 0; i < 2; i++) { o->iobuf[i] = (int32_t *)calloc(1, P); o->ioBig[i] = (int32_t *)calloc(1, 4 * P); }
  return o;
}

void ora_rect_destroy(ora_rect *o) {
  if (!o) return;
  for (int i = 0; i < 6; i++) { free(o->buf[i]); free(o->tmp[i]); }
  for (int i = 0; i < 2; i++) { free(o->iobuf[i]); free(o->ioBig[i]); }
  free(o);
}

void *ora_rect_buffer(ora_rect *o, const char *name) {
  if (!strncmp(name, "buf", 3) && name[3] >= '0' && name[3] < '6') return o->buf[name[3] - '0'];
  if (!strncmp(name, "tmp", 3) && name[3] >= '0' && name[3] < '6') return o->tmp[name[3] - '0'];
  if (!strncmp(name, "iobuf", 5) && name[5] >= '0' && name[5] < '2') return o->iobuf[name[5] - '0'];
  if (!strncmp(name, "ioBig", 5) && name[5] >= '0' && name[5] < '2') return o->ioBig[name[5] - '0'];
  return NULL;
}

void ora_rect_last_times(ora_rect *o, double out[5]) { for (int i = 0; i < 5; i++) out[i] = o->t[i]; }

// genGPUTask, oclrect.c:235-381.  Step numbers are those of SURVEY.md section 10.1.
void ora_rect_gpu_task(ora_rect *o, const uint8_t *img, int ws, int stop_step) {
  const int iw = o->iw, ih = o->ih, n = iw * ih;
  int32_t **buf = o->buf, **tmp = o->tmp, **iobuf = o->iobuf, **ioBig = o->ioBig;
#define STEP(k) do { if ((k) > 0 && stop_step == (k)) return; } while (0)
  double t0 = now_s();
  o->t[0] = o->t[1] = o->t[2] = o->t[3] = 0;

  // step 0 : memcpy + H2D of P bytes (oclrect.c:239-241); only ws*ih bytes carry the image
  memcpy(iobuf[0], img, (size_t)ws * ih);
  STEP(0);
  // step 1 : bgr2plab
  ora_convert_plab_bgr((uint32_t *)buf[0], (const uint8_t *)iobuf[0], iw, ih, ws);
  STEP(1);
  // step 2 : unpack
  ora_unpack_f_f_f_plab((float *)tmp[0], (float *)tmp[1], (float *)tmp[2], (const uint32_t *)buf[0], iw, ih);
  STEP(2);
  // step 3 : 3 x iirblur r=2, scratch = first plane of ioBig[0], ioBig[1]
  ora_iirblur_f_f((float *)tmp[3], (const float *)tmp[2], (float *)ioBig[0], (float *)ioBig[1], 2, iw, ih);
  ora_iirblur_f_f((float *)tmp[2], (const float *)tmp[1], (float *)ioBig[0], (float *)ioBig[1], 2, iw, ih);
  ora_iirblur_f_f((float *)tmp[1], (const float *)tmp[0], (float *)ioBig[0], (float *)ioBig[1], 2, iw, ih);
  STEP(3);
  // step 4 : pack
  ora_pack_plab_f_f_f((uint32_t *)buf[1], (const float *)tmp[1], (const float *)tmp[2], (const float *)tmp[3], iw, ih);
  STEP(4);
  // step 5 : edgevec
  ora_edgevec_f2_f((float *)ioBig[0], (const float *)tmp[1], iw, ih);
  STEP(5);
  // step 6 : edge magnitude
  ora_edge_f_plab((float *)tmp[0], (const uint32_t *)buf[1], iw, ih);
  STEP(6);
  // step 7 : NMS thinning
  ora_thinthres_f_f_f2((float *)buf[1], (const float *)tmp[0], (const float *)ioBig[0], iw, ih);
  STEP(7);
  // step 8 : threshold + cast -> edge bitmap #1
  ora_threshold_f_f((float *)tmp[0], (const float *)buf[1], 0.0f, 0.0f, 1.0f, n);
  ora_cast_i_f(tmp[1], (const float *)tmp[0], 1.0f, n);
  STEP(8);
  o->t[0] = now_s() - t0; t0 = now_s();

  // step 9 : junction / connect / stringify x2 (oclrect.cl versions)
  k_simpleJunction(buf[2], tmp[1], iw, ih);
  k_simpleConnect(tmp[1], buf[2], iw, ih);
  k_stringify(buf[2], tmp[1], 0, iw, ih);
  k_stringify(tmp[1], buf[2], 1, iw, ih);
  STEP(9);
  // step 10 : label8x bgc=-1
  label8x(buf[2], tmp[1], tmp[0], -1, iw, ih);
  STEP(10);
  // step 11 : calcStrength into buf[3] (NOT cleared: Q1) ; filterStrength 500
  k_rect_calcStrength(buf[3], (const float *)buf[1], buf[2], iw, ih);
  k_rect_filterStrength(buf[2], buf[3], 500, iw, ih);
  STEP(11);
  // step 12 : int edge mask -> i8 mask
  ora_threshold_i_i(tmp[0], buf[2], 0, 0, 1, n);
  ora_cast_c_i((int8_t *)tmp[1], tmp[0], n);
  STEP(12);
  // step 13 : 10 x (blblur0, blblur1)
  k_blblur0((uint32_t *)tmp[0], (const int8_t *)tmp[1], (const uint32_t *)buf[0], iw, ih);
  k_blblur1((uint32_t *)buf[4], (const int8_t *)tmp[1], (const uint32_t *)tmp[0], iw, ih);
  for (int i = 0; i < 9; i++) {
    k_blblur0((uint32_t *)tmp[0], (const int8_t *)tmp[1], (const uint32_t *)buf[4], iw, ih);
    k_blblur1((uint32_t *)buf[4], (const int8_t *)tmp[1], (const uint32_t *)tmp[0], iw, ih);
  }
  STEP(13);
  // step 14 : quantize 24^3 ; despeckle
  k_quantize((uint32_t *)tmp[0], (const uint32_t *)buf[4], 24, 24, 24, iw, ih);
  k_despeckle((uint32_t *)buf[4], (const uint32_t *)tmp[0], (const float *)buf[1], iw, ih);
  STEP(14);
  // step 15 : filterStrength 2500 ; threshold -> strong-edge bitmap in buf[3]
  k_rect_filterStrength(buf[2], buf[3], 2500, iw, ih);
  ora_threshold_i_i(buf[3], buf[2], 0, 0, 1, n);
  STEP(15);
  // step 16 : junction of strong edges ; merge mask
  k_simpleJunction(tmp[0], buf[2], iw, ih);
  k_clear(tmp[1], n);
  k_mkMergeMask0(tmp[1], tmp[0], iw, ih);
  k_mkMergeMask1(tmp[1], tmp[0], iw, ih);
  STEP(16);
  // step 17 : colour-region labels
  labelMerge(buf[5], buf[4], tmp[1], buf[2], iw, ih);
  STEP(17);
  // step 18 : calcSize into tmp[0] which still holds the junction map (Q2) ; despeckle2
  k_calcSize(tmp[0], buf[5], iw, ih);
  k_despeckle2(buf[5], tmp[0], 16, iw, ih);
  STEP(18);
  // step 19 : boundary bands and their components -> segid map
  k_markBoundary(tmp[1], buf[5], iw, ih);
  label8x(iobuf[1], tmp[1], tmp[0], -1, iw, ih);
  STEP(19);
  o->t[1] = now_s() - t0; t0 = now_s();

  // step 20 : polyline
  ora_polyline_execute((ora_ls_t *)ioBig[0], iw * ih * 4 * 4, buf[0], buf[3], ioBig[1],
                       tmp[0], tmp[1], tmp[2], tmp[3], tmp[4], tmp[5], 4.0f, 20, iw, ih, 0);
  STEP(20);
  o->t[2] = now_s() - t0; t0 = now_s();

  // step 21 : vote table
  k_clear(ioBig[1], n * 4);
  k_reduceLS(ioBig[1], iobuf[1], buf[0], iw, ih, iw * ih * 4 / 5);
  o->t[3] = now_s() - t0;
#undef STEP
}

ora_rect_t *ora_rect_cpu_task(ora_rect *o, double tanAOV) {
  double t0 = now_s();
  ora_rect_t *r = ora_tail((const ora_ls_t *)o->ioBig[0], o->iobuf[1], o->ioBig[1], o->iw, o->ih, tanAOV);
  o->t[4] = now_s() - t0;
  return r;
}

ora_rect_t *ora_rect_execute_once(ora_rect *o, const uint8_t *img, int ws, double tanAOV) {
  ora_rect_gpu_task(o, img, ws, 0);
  return ora_rect_cpu_task(o, tanAOV);
}

}  // extern "C"
