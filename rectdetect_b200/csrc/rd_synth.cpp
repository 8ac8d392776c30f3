// rd_synth.cpp - deterministic synthetic frame generator (workload input for tests and bench.py).
//
// Not part of the hot path and not part of the oracle: it only manufactures BGR8 frames of the kind the
// reference's demo programs are pointed at (rect.cpp:68-74, vidrect.cpp:159-166 take camera / file frames):
// a smooth background gradient with K = max(4, iw*ih/76800) filled, projected 3-D rectangles (horizontal
// angle of view 72 degrees, README.md:52-55), +-2 LSB noise and 2x2 supersampled edges.
// Everything derives from one 64-bit seed through splitmix64, so a (seed, iw, ih) triple names a frame.
#include <stdint.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
  }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }   // [0,1)
  double range(double lo, double hi) { return lo + (hi - lo) * uni(); }
  int irange(int lo, int hi) { return lo + (int)(next() % (uint64_t)(hi - lo + 1)); }  // inclusive
};

struct Quad { double x[4], y[4]; };

static bool inside(const Quad &q, double px, double py) {
  // convex quad, either orientation
  int pos = 0, neg = 0;
  for (int i = 0; i < 4; i++) {
    int j = (i + 1) & 3;
    double c = (q.x[j] - q.x[i]) * (py - q.y[i]) - (q.y[j] - q.y[i]) * (px - q.x[i]);
    if (c > 0) pos++; else if (c < 0) neg++;
  }
  return pos == 0 || neg == 0;
}

static bool makeQuad(Rng &rng, int iw, int ih, Quad &q) {
  const double PI = 3.14159265358979323846;
  const double f = (iw / 2.0) / tan(36.0 * PI / 180.0);
  for (int attempt = 0; attempt < 64; attempt++) {
    double depth = rng.range(2.0, 7.0);
    double aspect = rng.range(0.5, 2.0);
    double hgt = rng.range(0.5, 1.6);
    double wid = hgt * aspect;
    double yaw = rng.range(-50.0, 50.0) * PI / 180.0;
    double pitch = rng.range(-50.0, 50.0) * PI / 180.0;
    double roll = rng.range(-35.0, 35.0) * PI / 180.0;
    double cxs = rng.range(0.08, 0.92) * iw, cys = rng.range(0.08, 0.92) * ih;
    double CX = (cxs - iw / 2.0) * depth / f, CY = -(cys - ih / 2.0) * depth / f;
    const double lx[4] = {-wid / 2, wid / 2, wid / 2, -wid / 2};
    const double ly[4] = {hgt / 2, hgt / 2, -hgt / 2, -hgt / 2};
    bool ok = true;
    for (int i = 0; i < 4 && ok; i++) {
      // roll about z, pitch about x, yaw about y
      double x0 = lx[i] * cos(roll) - ly[i] * sin(roll), y0 = lx[i] * sin(roll) + ly[i] * cos(roll), z0 = 0;
      double y1 = y0 * cos(pitch) - z0 * sin(pitch), z1 = y0 * sin(pitch) + z0 * cos(pitch);
      double x2 = x0 * cos(yaw) + z1 * sin(yaw), z2 = -x0 * sin(yaw) + z1 * cos(yaw);
      double X = x2 + CX, Y = y1 + CY, Z = z2 + depth;
      if (Z < 0.5) { ok = false; break; }
      q.x[i] = f * X / Z + iw / 2.0;
      q.y[i] = -f * Y / Z + ih / 2.0;
      if (q.x[i] < 8 || q.x[i] > iw - 9 || q.y[i] < 8 || q.y[i] > ih - 9) ok = false;
    }
    if (!ok) continue;
    for (int i = 0; i < 4 && ok; i++) {
      int j = (i + 1) & 3;
      double dx = q.x[j] - q.x[i], dy = q.y[j] - q.y[i];
      if (dx * dx + dy * dy < 48.0 * 48.0) ok = false;
    }
    if (!ok) continue;
    // reject slivers: both diagonals must be long too, and the quad must be convex
    double cr[4];
    for (int i = 0; i < 4; i++) {
      int j = (i + 1) & 3, k = (i + 2) & 3;
      cr[i] = (q.x[j] - q.x[i]) * (q.y[k] - q.y[j]) - (q.y[j] - q.y[i]) * (q.x[k] - q.x[j]);
    }
    if (!((cr[0] > 0 && cr[1] > 0 && cr[2] > 0 && cr[3] > 0) || (cr[0] < 0 && cr[1] < 0 && cr[2] < 0 && cr[3] < 0))) continue;
    return true;
  }
  return false;
}

}  // namespace

extern "C" {

// Writes a BGR8 frame with row stride ws (>= 3*iw) into out.  If quads_out != NULL it receives up to max_quads
// ground-truth quadrilaterals (8 doubles each: x0,y0,...,x3,y3) in painter's order.  Returns the number of quads.
int rd_synth_frame(uint8_t *out, int iw, int ih, int ws, uint64_t seed, double *quads_out, int max_quads) {
  Rng rng(seed * 0x2545F4914F6CDD1DULL + 0x1234567ULL);
  const int K = (iw * ih / 76800) > 4 ? (iw * ih / 76800) : 4;
  std::vector<float> canvas((size_t)iw * ih * 3);

  double base[3], gx[3], gy[3];
  double grey = rng.range(90, 170);
  for (int c = 0; c < 3; c++) {
    base[c] = grey + rng.range(-18, 18);
    gx[c] = rng.range(-28, 28);
    gy[c] = rng.range(-28, 28);
  }
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++)
      for (int c = 0; c < 3; c++)
        canvas[((size_t)y * iw + x) * 3 + c] = (float)(base[c] + gx[c] * ((double)x / iw - 0.5) + gy[c] * ((double)y / ih - 0.5));

  int nq = 0;
  for (int k = 0; k < K; k++) {
    Quad q;
    if (!makeQuad(rng, iw, ih, q)) continue;
    double col[3];
    for (int tries = 0; tries < 32; tries++) {
      double maxd = 0;
      for (int c = 0; c < 3; c++) { col[c] = rng.range(10, 245); double d = fabs(col[c] - base[c]); if (d > maxd) maxd = d; }
      if (maxd >= 70) break;
    }
    int bx0 = iw, bx1 = 0, by0 = ih, by1 = 0;
    for (int i = 0; i < 4; i++) {
      int fx = (int)floor(q.x[i]), fy = (int)floor(q.y[i]);
      if (fx < bx0) bx0 = fx;
      if (fx + 1 > bx1) bx1 = fx + 1;
      if (fy < by0) by0 = fy;
      if (fy + 1 > by1) by1 = fy + 1;
    }
    if (bx0 < 0) bx0 = 0;
    if (by0 < 0) by0 = 0;
    if (bx1 > iw - 1) bx1 = iw - 1;
    if (by1 > ih - 1) by1 = ih - 1;
    for (int y = by0; y <= by1; y++)
      for (int x = bx0; x <= bx1; x++) {
        int cov = 0;
        for (int sy = 0; sy < 2; sy++)
          for (int sx = 0; sx < 2; sx++)
            if (inside(q, x + 0.25 + 0.5 * sx, y + 0.25 + 0.5 * sy)) cov++;
        if (!cov) continue;
        float a = cov * 0.25f;
        float *px = &canvas[((size_t)y * iw + x) * 3];
        for (int c = 0; c < 3; c++) px[c] = px[c] * (1.0f - a) + (float)col[c] * a;
      }
    if (quads_out && nq < max_quads)
      for (int i = 0; i < 4; i++) { quads_out[nq * 8 + 2 * i] = q.x[i]; quads_out[nq * 8 + 2 * i + 1] = q.y[i]; }
    nq++;
  }

  for (int y = 0; y < ih; y++) {
    uint8_t *row = out + (size_t)y * ws;
    for (int x = 0; x < iw; x++)
      for (int c = 0; c < 3; c++) {
        int noise = rng.irange(-2, 2);
        int v = (int)lrintf(canvas[((size_t)y * iw + x) * 3 + c]) + noise;
        row[x * 3 + c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
      }
    for (int x = iw * 3; x < ws; x++) row[x] = 0;
  }
  return nq;
}

}  // extern "C"
