// Host replay of the first-pass wavefront of the merge labelling (rectdetect_b200/csrc/rd_merge1.cuh, kernels k_m1_pre / k_m1_wave
// of rd_ccl.cu): every row is a lane, ALL rows advance in lock step M1_SKEW pixels behind the row above (the tightest schedule the
// kernel allows; warps further up are only ever further ahead), the label of the pixel above travels through the three-deep
// delay line of the kernel.  mode 0: planes in image layout; mode 1: the same with every access logged, and the replay counts
// HAZARDS: an address read or written by one lane and written by another lane in the same step - the lock step would then depend
// on the order of the lanes inside an instruction; mode 2: planes in the kernel's time-major layout (m1_index).  The result
// (label = min(A, B)) is to be compared with the sequential pass of the oracle.  Test infrastructure.
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../rectdetect_b200/csrc/rd_merge1.cuh"

struct Access { int addr; int lane; bool write; };
static long g_looks = 0;
extern "C" long emu_merge1_looks(void) { const long v = g_looks; g_looks = 0; return v; }
struct LogMem {                                  // M1Linear with a log
  int *A, *B; int iw, p; std::vector<Access> *log; int lane;
  void rd(int a) const { log->push_back({a, lane, false}); }
  void wr(int a) const { log->push_back({a, lane, true}); }
  int look(int q, bool withB) const { g_looks++; rd(2 * q); int v = A[q]; if (withB) { rd(2 * q + 1); const int b = B[q]; if (b < v) v = b; } return v; }
  void setSelf(int v) { wr(2 * p); A[p] = v; }
  void setLeft(int v) { wr(2 * (p - 1)); A[p - 1] = v; }
  void setUp(int v) { wr(2 * (p - iw) + 1); B[p - iw] = v; }
};

extern "C" long emu_merge1(int32_t *label, const int32_t *pix, const int32_t *mask, const int32_t *edge, int iw, int ih, int skew, int mode) {
  const int n = iw * ih;
  const bool tmj = mode == 2, check = mode == 1;
  std::vector<int> A(n), B(n, M1_NONE);
  std::vector<uint8_t> F(n);
  auto at = [&](int q) { return tmj ? m1_index(q, iw, ih) : q; };
  for (int y = 0; y < ih; y++)
    for (int x = 0; x < iw; x++) {
      const int p = y * iw + x;
      const bool in = x > 0 && y > 0 && x < iw - 1 && y < ih - 1;
      int L0;
      F[at(p)] = (uint8_t)m1_record(x, y, iw, ih, (unsigned)pix[p], y > 0 ? (unsigned)pix[p - iw] : 0u, x > 0 ? (unsigned)pix[p - 1] : 0u, in ? (unsigned)pix[p + 1] : 0u,
                                    in ? (unsigned)pix[p + 1 - iw] : 0u, in && mask[p] != 0, in && edge[p] <= 0, in && edge[p + 1] <= 0, L0);
      A[at(p)] = L0;
    }
  std::vector<M1Row> row(ih);
  for (auto &r : row) { r.gleft = 0; r.gleftF = 0u; r.croot[0] = r.croot[1] = -1; }
  // d[y][0..2]: the last three values lane y produced for the row below (final A of the pixel left of the one it just did);
  // like the kernel, lane y takes its `aup` from d[y-1][2] at the start of a step (before anybody pushes), row 1 reads row 0 from memory
  std::vector<unsigned> d(3 * (size_t)ih, 0u), aups(ih, 0u);
  std::vector<Access> log;
  long hazards = 0;
  const int steps = iw + skew * ih + 2;
  for (int t = 0; t < steps; t++) {
    log.clear();
    for (int y = 2; y < ih - 1; y++) aups[y] = d[3 * (y - 1) + 2];
    for (int y = 1; y < ih - 1; y++) {
      const int x = t - skew * y;
      unsigned *dd = &d[3 * y];
      if (x < 0 || x >= iw) { dd[2] = dd[1]; dd[1] = dd[0]; continue; }
      const int p = y * iw + x;
      const unsigned f = F[at(p)];
      unsigned fin = (unsigned)row[y].gleft | row[y].gleftF;
      if (!(f & M1_INT)) { row[y].gleft = A[at(p)]; row[y].gleftF = 0u; }
      else {
        const unsigned aup = (y == 1 || skew != M1_SKEW) ? (unsigned)A[at(p - iw)] : aups[y];
        if (check) { LogMem m{A.data(), B.data(), iw, p, &log, y}; fin = m1_pixel(p, iw, f, aup, m, row[y]); }
        else if (tmj && iw >= M1_BIG && n < (1 << 24)) {
          M1TimeMajor<true> m; m.A = A.data(); m.B = B.data(); m.iw = iw; m.ih = ih; m.dv = m1_div_make(iw);
          m1_row_setup(m, y); m.tm = m.wrap(x + M1_SKEW * (y & 31));
          fin = m1_pixel(p, iw, f, aup, m, row[y]);
        } else if (tmj) {
          M1TimeMajor<false> m; m.A = A.data(); m.B = B.data(); m.iw = iw; m.ih = ih; m.dv = M1Div{0u, 0};
          m1_row_setup(m, y); m.tm = m.wrap(x + M1_SKEW * (y & 31));
          fin = m1_pixel(p, iw, f, aup, m, row[y]);
        }
        else { M1Linear m{A.data(), B.data(), iw, p}; fin = m1_pixel(p, iw, f, aup, m, row[y]); }
      }
      dd[2] = dd[1]; dd[1] = dd[0]; dd[0] = fin;
    }
    if (check) {
      std::sort(log.begin(), log.end(), [](const Access &a, const Access &b) { return a.addr < b.addr; });
      for (size_t i = 0; i < log.size();) {
        size_t j = i;
        while (j < log.size() && log[j].addr == log[i].addr) j++;
        bool w = false, multi = false;
        for (size_t k = i; k < j; k++) { w |= log[k].write; multi |= log[k].lane != log[i].lane; }
        if (w && multi) hazards++;
        i = j;
      }
    }
  }
  // the layout is a bijection; the multiply-high division is exact on the whole frame
  if (tmj && iw >= M1_BIG && n < (1 << 24)) {
    const M1Div dv = m1_div_make(iw);
    for (int q = 0; q < n; q++) if (m1_div(q, dv) != q / iw) return -2;
  }
  if (tmj) {
    std::vector<uint8_t> seen(n, 0);
    for (int p = 0; p < n; p++) { const int i = m1_index(p, iw, ih); if (i < 0 || i >= n || seen[i]) return -1; seen[i] = 1; }
  }
  for (int p = 0; p < n; p++) label[p] = std::min(A[at(p)], B[at(p)]);
  return hazards;
}
