// emu_despeckle2x.cpp - host replay of the exact despeckle2 kernels (rectdetect_b200/csrc/rd_despeckle2.cu).  TEST INFRASTRUCTURE
// (tests/test_emu_kernels.py): kd2_pre and kd2_seq are replayed with the kernels' own building blocks (rd_despeckle2.cuh), the
// 32 lanes of the warp as arrays, the shuffle scan step by step, the chunk carry and the two-row buffer as in the kernel - a
// check of the formulation and of the indexing where no GPU is available (the -m gpu tests then check the kernels themselves).
#include <stddef.h>
#include <vector>
#include "../rectdetect_b200/csrc/rd_despeckle2.cuh"

extern "C" long emu_despeckle2x(int *dst, const int *label, const int *size, int thre, int iw, int ih) {
  const size_t n = (size_t)iw * ih;
  std::vector<int> list(n), recL(n), recS(n), cnt(ih);
  long chunks = 0;
  // kd2_pre: one CTA per row; compaction in ascending x
  for (int y = 0; y < ih; y++) {
    int c = 0;
    for (int x = 0; x < iw; x++) {
      const int l = label[(size_t)y * iw + x];
      if (size[l] > thre) { dst[(size_t)y * iw + x] = l; continue; }
      int bl, bs;
      const int rec = d2_static(x, y, label, size, thre, iw, ih, bl, bs);
      const size_t o = (size_t)y * iw + c++;
      list[o] = rec; recL[o] = bl; recS[o] = bs;
    }
    cnt[y] = c;
  }
  // kd2_seq: one warp, rows top-down, 32 entries per step
  std::vector<int> rbl(2 * (size_t)iw, -12345), rbs(2 * (size_t)iw, -12345);
  int carryL = 0, carryS = 0;
  for (int py = 0; py < ih; py++)
    for (int pc = 0; pc < cnt[py]; pc += D2_CHUNK, chunks++) {
      int bl[32], bs[32], T[32], X[32];
      bool valid[32];
      for (int lane = 0; lane < 32; lane++) {
        valid[lane] = pc + lane < cnt[py];
        const size_t o = (size_t)py * iw + pc + lane;
        const int rec = valid[lane] ? list[o] : 0;
        bl[lane] = valid[lane] ? recL[o] : 0; bs[lane] = valid[lane] ? recS[o] : 0;
        const int x = rec & 0xffff, dyn = (rec >> 20) & 15;
        int code = (rec >> 16) & 15;
        const size_t ab = (size_t)((py + 1) & 1) * iw;
        if (dyn & 1) d2_take(bl[lane], bs[lane], code, rbl[ab + x - 1], rbs[ab + x - 1], 1);
        if (dyn & 2) d2_take(bl[lane], bs[lane], code, rbl[ab + x], rbs[ab + x], 2);
        if (dyn & 4) d2_take(bl[lane], bs[lane], code, rbl[ab + x + 1], rbs[ab + x + 1], 3);
        T[lane] = (dyn & 8) ? d2_threshold(bs[lane], code) : D2_HEAD;
        X[lane] = x;
      }
      if (T[0] != D2_HEAD) {
        if (carryS >= T[0]) { bl[0] = carryL; bs[0] = carryS; }
        T[0] = D2_HEAD;
      }
      for (int d = 1; d < 32; d <<= 1) {
        bool all = true;
        for (int lane = 0; lane < 32; lane++) all &= T[lane] == D2_HEAD;
        if (all) break;
        int nl[32], ns[32], nt[32];
        for (int lane = 0; lane < 32; lane++) {
          nl[lane] = bl[lane]; ns[lane] = bs[lane]; nt[lane] = T[lane];
          if (lane >= d && T[lane] != D2_HEAD) d2_compose(nl[lane], ns[lane], nt[lane], bl[lane - d], bs[lane - d], T[lane - d]);
        }
        for (int lane = 0; lane < 32; lane++) { bl[lane] = nl[lane]; bs[lane] = ns[lane]; T[lane] = nt[lane]; }
      }
      for (int lane = 0; lane < 32; lane++)
        if (valid[lane]) {
          dst[(size_t)py * iw + X[lane]] = bl[lane];
          rbl[(size_t)(py & 1) * iw + X[lane]] = bl[lane]; rbs[(size_t)(py & 1) * iw + X[lane]] = bs[lane];
        }
      carryL = bl[31]; carryS = bs[31];
    }
  return chunks;
}
