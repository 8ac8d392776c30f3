// emu_gtail.cpp - host replay of the GPU tail (rectdetect_b200/csrc/rd_gtail.cu).  TEST INFRASTRUCTURE (tests/test_emu_kernels.py).
// The kernels' logic lives in rd_gtail.cuh; here the per-item kernels run as plain loops and the warp-per-candidate kernel runs with
// its 32 lanes as fibers (ucontext): every warp collective (ballot / shuffle / sync) is a rendezvous - a lane deposits its operand and
// yields, and once all 32 have arrived each reads the result.  What comes out must equal executeCPUTask (rd_tail.cpp / the oracle /
// the reference's own) bit for bit, in the same order - a check of the GPU formulation that needs no GPU.
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>
#include <functional>
#include <vector>
#include "../rectdetect_b200/csrc/rd_gtail.cuh"

namespace {
struct EmuShared {
  ucontext_t sched, ctx[32];
  std::vector<char> stack[32];
  bool done[32];
  union Slot { double d; long long i; } in[2][32];
  int cur;
};
EmuShared *g_sh;
struct GtWarpEmu {
  int lane;
  int ncoll = 0;
  EmuShared::Slot *deposit() { return &g_sh->in[ncoll & 1][lane]; }
  const EmuShared::Slot *rendezvous() { const EmuShared::Slot *r = g_sh->in[ncoll & 1]; ncoll++; swapcontext(&g_sh->ctx[lane], &g_sh->sched); return r; }
  unsigned ballot(bool p) { deposit()->i = p ? 1 : 0; const EmuShared::Slot *r = rendezvous(); unsigned b = 0; for (int l = 0; l < 32; l++) b |= (unsigned)(r[l].i & 1) << l; return b; }
  int shfl(int v, int src) { deposit()->i = v; const EmuShared::Slot *r = rendezvous(); return (int)r[src & 31].i; }
  double shfl(double v, int src) { deposit()->d = v; const EmuShared::Slot *r = rendezvous(); return r[src & 31].d; }
  void sync() { deposit()->i = 0; rendezvous(); }
};
std::function<void(GtWarpEmu &)> *g_body;
void lane_entry(int lane) {
  GtWarpEmu w;
  w.lane = lane;
  (*g_body)(w);
  g_sh->done[lane] = true;
  swapcontext(&g_sh->ctx[lane], &g_sh->sched);
}
void run_warp(std::function<void(GtWarpEmu &)> body) {
  static EmuShared sh;
  g_sh = &sh;
  g_body = &body;
  for (int l = 0; l < 32; l++) {
    sh.stack[l].resize(256 * 1024);
    sh.done[l] = false;
    getcontext(&sh.ctx[l]);
    sh.ctx[l].uc_stack.ss_sp = sh.stack[l].data();
    sh.ctx[l].uc_stack.ss_size = sh.stack[l].size();
    sh.ctx[l].uc_link = &sh.sched;
    makecontext(&sh.ctx[l], (void (*)())lane_entry, 1, l);
  }
  for (;;) {
    bool any = false;
    for (int l = 0; l < 32; l++)
      if (!sh.done[l]) { any = true; swapcontext(&sh.sched, &sh.ctx[l]); }
    if (!any) break;
  }
}
}  // namespace

// returns the number of rectangles written to out (GtRect = rect_t layout), -1 on a scratch error; stats[0..3] = pairs, regions, chains, accepted
extern "C" int emu_gtail(GtRect *out, int cap, const GtLS *ls, const int *segid, const int *votes, int iw, int ih, double tanAOV, int *stats) {
  const size_t npx = (size_t)iw * ih;
  std::vector<int> table(2 * npx, 0);
  const size_t S = 32 * npx + 4096, PS = npx + 65536;
  std::vector<unsigned char> scratch(S), persist(PS);
  int n = *(const int *)ls;
  const int capn = (int)(npx * 16 / sizeof(GtLS)) - 1;
  n = n < 0 ? 0 : (n > capn ? capn : n);
  const GtLayout L = gt_layout(scratch.data(), S, persist.data(), PS, n);
  if (!L.ok) return -1;
  memset(L.hdr, 0, sizeof(GtHdr));
  L.hdr->n = n;
  const int nentry = iw * ih * 4 / 5;
  for (int i = n; i >= 1; i--) gt_item_samples(i, ls, segid, table.data(), L, n, iw, ih);       // (any order: the kernel's is arbitrary)
  for (int p = L.hdr->npairs - 1; p >= 0; p--) gt_item_regions(p, table.data(), L);
  for (int p = L.hdr->npairs - 1; p >= 0; p--) gt_item_members(p, table.data(), L);
  for (int r = 0; r < L.hdr->nreg; r++) gt_item_order_region(r, L);
  for (int c = 0; c < L.hdr->nchain; c++) gt_item_order_chain(c, L);
  for (int p = 0; p < L.hdr->npairs; p++) { table[2 * (size_t)L.pairs[p].segid] = 0; table[2 * (size_t)L.pairs[p].segid + 1] = 0; }
  for (size_t i = 0; i < table.size(); i++) if (table[i] != 0) return -2;
  const int nc = L.hdr->nreg + L.hdr->nchain;
  unsigned long long off = 0;
  for (int c = 0; c < nc; c++) { L.cands[c].off = off; off += gt_align16(gt_work_bytes(L.cands[c].m)); }
  if (off > L.workBytes) return -1;
  L.hdr->ncand = nc;
  for (int c = 0; c < nc; c++) {
    GtQuad lanes[32];
    run_warp([&](GtWarpEmu &w) { lanes[w.lane] = gt_cand_quad(w, L.cands[c], ls, votes, L, n, iw, ih, nentry); });
    for (int l = 1; l < 32; l++) if (memcmp(&lanes[l], &lanes[0], sizeof(GtQuad)) != 0) return -3;      // every lane returns the same quadrilateral
    L.quads[c] = lanes[0];
    if (lanes[0].valid) L.vlist[L.hdr->nvalid++] = c;
  }
  for (int v = 0; v < L.hdr->nvalid; v++)
    for (int mode = 0; mode < 2; mode++) {
      const int c = L.vlist[v];
      GtP3 ray[4];
      gt_pose_setup(L.quads[c], iw, ih, tanAOV, ray);
      L.pose[2 * c + mode] = gt_pose_run(ray, mode);
    }
  int nr = 0;
  for (int c = 0; c < nc; c++) {
    if (!L.quads[c].valid) continue;
    if (nr >= cap) return -4;
    GtP3 ray[4];
    const int first = gt_pose_setup(L.quads[c], iw, ih, tanAOV, ray);
    memset(&out[nr], 0, sizeof(GtRect));
    gt_pose_finish(L.quads[c], first, ray, L.pose[2 * c + 1], L.pose[2 * c], out[nr]);
    nr++;
  }
  if (stats) { stats[0] = L.hdr->npairs; stats[1] = L.hdr->nreg; stats[2] = L.hdr->nchain; stats[3] = L.hdr->nvalid; }
  return nr;
}
