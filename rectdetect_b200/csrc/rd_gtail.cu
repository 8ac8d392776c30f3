// rd_gtail.cu - the host tail of the reference (executeCPUTask, oclrect.c:1049-1226) as CUDA kernels: nothing but the finished
// rectangle list is read back.  The logic lives in rd_gtail.cuh (shared with the host replay of the tests);
// this file holds the kernels and their launch sequence:
//   kt_samples : one thread per line segment - its 15 sample points -> distinct (region, segment) pairs, per-region counters
//                in a direct-address table (two ints per region id), polyline-chain heads and their lengths
//   kt_regions : one thread per pair - the pair that saw its region first opens the region (>= 4 segments)
//   kt_members : one thread per pair - fills the member list of its region
//   kt_order   : one CTA per frame - candidates in the reference's order (ArrayMap bucket order, then chains by head), work
//                storage offsets, table clean-up
//   kt_quad    : one WARP per candidate - edge list, removeShortLS, quick hull, pickExternalLS (lanes test 32 edges at a
//                time), corners, acceptance tests -> quadrilateral
//   kt_pose    : one THREAD per (accepted quadrilateral, objective variant) - the two preconditioned nonlinear-CG runs of
//                poseEstimation (1 176 objective evaluations each, FP64); lanes of a warp run different quadrilaterals
//   kt_finish  : one CTA per frame - picks the better variant, looksLikeAScreen, writes the rect_t list in candidate order
// All arithmetic is IEEE double without FMA contraction, in the order of operations of oclrect.c / vec234.h.
#include "rd_common.cuh"
#include "rd_gtail.cuh"

struct GtWarpDev {
  int lane;
  __device__ __forceinline__ unsigned ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
  __device__ __forceinline__ int shfl(int v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
  __device__ __forceinline__ double shfl(double v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
};

// blob = the read-back record of the frame: 64-byte header, the rect_t list in its first half, the persistent tail state
// (GtHdr, quadrilaterals, pose results) in its second half
struct GtArgs {
  const GtLS *ls; const int *segid; const int *votes; int *table; unsigned char *scratch; unsigned char *blob;
  size_t scratchBytes, blobBytes; int iw, ih, nentry; size_t fs;
};
__device__ __forceinline__ GtLayout gt_layout_of(const GtArgs &a, int n) {
  const size_t half = (a.blobBytes / 2) & ~(size_t)15;
  return gt_layout(a.scratch, a.scratchBytes, a.blob + half, a.blobBytes - half, n);
}
__device__ __forceinline__ GtHdr *gt_hdr_of(const GtArgs &a) { return (GtHdr *)(a.blob + ((a.blobBytes / 2) & ~(size_t)15)); }
__device__ __forceinline__ GtArgs gt_frame(GtArgs a, int z) {
  const size_t o = (size_t)z * a.fs;
  a.ls = (const GtLS *)((const char *)a.ls + o); a.segid = (const int *)((const char *)a.segid + o); a.votes = (const int *)((const char *)a.votes + o);
  a.table = (int *)((char *)a.table + o); a.scratch += o; a.blob += o;
  return a;
}
__device__ __forceinline__ int gt_count(const GtArgs &a) {
  const int n = *(const int *)a.ls, cap = (int)((size_t)a.iw * a.ih * 16 / sizeof(GtLS)) - 1;
  return n < 0 ? 0 : (n > cap ? cap : n);
}

#define GT_TPB 128
__global__ void __launch_bounds__(GT_TPB) kt_samples(GtArgs a0) {
  const GtArgs a = gt_frame(a0, blockIdx.y);
  const int n = gt_count(a);
  const GtLayout L = gt_layout_of(a, n);
  if (blockIdx.x == 0 && threadIdx.x == 0) L.hdr->n = n;
  if (!L.ok) { if (blockIdx.x == 0 && threadIdx.x == 0) L.hdr->err = GT_ERR_SCRATCH; return; }
  for (int i = 1 + blockIdx.x * GT_TPB + threadIdx.x; i <= n; i += gridDim.x * GT_TPB) gt_item_samples(i, a.ls, a.segid, a.table, L, n, a.iw, a.ih);
}
__global__ void __launch_bounds__(GT_TPB) kt_regions(GtArgs a0) {
  const GtArgs a = gt_frame(a0, blockIdx.y);
  const GtLayout L = gt_layout_of(a, gt_count(a));
  if (!L.ok) return;
  const int np = L.hdr->npairs;
  for (int p = blockIdx.x * GT_TPB + threadIdx.x; p < np; p += gridDim.x * GT_TPB) gt_item_regions(p, a.table, L);
}
__global__ void __launch_bounds__(GT_TPB) kt_members(GtArgs a0) {
  const GtArgs a = gt_frame(a0, blockIdx.y);
  const GtLayout L = gt_layout_of(a, gt_count(a));
  if (!L.ok) return;
  const int np = L.hdr->npairs;
  for (int p = blockIdx.x * GT_TPB + threadIdx.x; p < np; p += gridDim.x * GT_TPB) gt_item_members(p, a.table, L);
}
__global__ void __launch_bounds__(256) kt_order(GtArgs a0) {
  const GtArgs a = gt_frame(a0, blockIdx.x);
  const GtLayout L = gt_layout_of(a, gt_count(a));
  if (!L.ok) return;
  const int nreg = L.hdr->nreg, nchain = L.hdr->nchain, np = L.hdr->npairs;
  for (int r = threadIdx.x; r < nreg; r += 256) gt_item_order_region(r, L);
  for (int c = threadIdx.x; c < nchain; c += 256) gt_item_order_chain(c, L);
  for (int p = threadIdx.x; p < np; p += 256) {                  // the table goes back to all-zero for the next frame
    const int s = L.pairs[p].segid;
    a.table[2 * (size_t)s] = 0; a.table[2 * (size_t)s + 1] = 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nc = nreg + nchain;
    unsigned long long off = 0;
    for (int c = 0; c < nc; c++) { L.cands[c].off = off; off += gt_align16(gt_work_bytes(L.cands[c].m)); }
    if (off > L.workBytes) L.hdr->err = GT_ERR_SCRATCH;
    L.hdr->ncand = nc;
  }
}
#define GTQ_WARPS 4
__global__ void __launch_bounds__(GTQ_WARPS * 32) kt_quad(GtArgs a0) {
  const GtArgs a = gt_frame(a0, blockIdx.y);
  const int n = gt_count(a);
  const GtLayout L = gt_layout_of(a, n);
  if (!L.ok || L.hdr->err) return;
  GtWarpDev w;
  w.lane = threadIdx.x & 31;
  const int nc = L.hdr->ncand;
  for (int c = blockIdx.x * GTQ_WARPS + (threadIdx.x >> 5); c < nc; c += gridDim.x * GTQ_WARPS) {
    const GtQuad q = gt_cand_quad(w, L.cands[c], a.ls, a.votes, L, n, a.iw, a.ih, a.nentry);
    if (w.lane == 0) {
      L.quads[c] = q;
      if (q.valid) L.vlist[atomicAdd(&L.hdr->nvalid, 1)] = c;
    }
  }
}
__global__ void __launch_bounds__(64) kt_pose(GtArgs a0, double tanAOV) {
  const GtArgs a = gt_frame(a0, blockIdx.y);
  const GtLayout L = gt_layout_of(a, gt_hdr_of(a)->n);           // (not gt_count: the segment list may belong to the next frame by now)
  if (!L.okPersist || L.hdr->err) return;
  const int nv = L.hdr->nvalid;
  for (int t = blockIdx.x * 64 + threadIdx.x; t < ((nv + 31) >> 5) * 64; t += gridDim.x * 64) {
    // the two variants of one quadrilateral sit 32 threads apart, so a warp runs ONE variant on 32 quadrilaterals
    const int blk = t >> 6, within = t & 63, v = blk * 32 + (within & 31), mode = within < 32 ? 1 : 0;
    if (v >= nv) continue;
    const int c = L.vlist[v];
    GtP3 ray[4];
    gt_pose_setup(L.quads[c], a.iw, a.ih, tanAOV, ray);
    L.pose[2 * c + mode] = gt_pose_run(ray, mode);
  }
}
__global__ void __launch_bounds__(256) kt_finish(GtArgs a0, double tanAOV) {
  const GtArgs a = gt_frame(a0, blockIdx.x);
  const int n = gt_hdr_of(a)->n;
  const GtLayout L = gt_layout_of(a, n);
  int *hdr = (int *)a.blob;                                      // [0] segments, [1] rectangles, [2] error code, [3] candidates
  GtRect *out = (GtRect *)(a.blob + 64);
  const int cap = (int)((((a.blobBytes / 2) & ~(size_t)15) - 64) / sizeof(GtRect));
  __shared__ int wsum[8], base;
  if (!L.okPersist || L.hdr->err) { if (threadIdx.x == 0) { hdr[0] = n; hdr[1] = 0; hdr[2] = L.okPersist ? L.hdr->err : GT_ERR_SCRATCH; hdr[3] = 0; } return; }
  const int nc = L.hdr->ncand;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  for (int c0 = 0; c0 < nc; c0 += 256) {
    const int c = c0 + threadIdx.x;
    const bool valid = c < nc && L.quads[c].valid;
    const unsigned b = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) wsum[wp] = __popc(b);
    __syncthreads();
    int pos = base + __popc(b & ((1u << lane) - 1u));
    for (int k = 0; k < wp; k++) pos += wsum[k];
    if (valid && pos < cap) {
      const GtQuad q = L.quads[c];
      GtP3 ray[4];
      const int first = gt_pose_setup(q, a.iw, a.ih, tanAOV, ray);
      gt_pose_finish(q, first, ray, L.pose[2 * c + 1], L.pose[2 * c], out[pos]);
    }
    __syncthreads();
    if (threadIdx.x == 0) { int t = base; for (int k = 0; k < 8; k++) t += wsum[k]; base = t; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { hdr[0] = n; hdr[1] = base < cap ? base : cap; hdr[2] = base > cap ? GT_ERR_RECTS : 0; hdr[3] = nc; }
}

// The tail for nb frames: rect lists -> blob (64-byte header: segments, rectangles, error code, candidates; rect_t entries behind
// it).  table: 2*iw*ih ints, all zero on entry and on exit.  scratch: scratchBytes of work space.
// phases bit 0: segments -> quadrilaterals (needs ls / segid / votes / table / scratch), bit 1: pose estimation + rect list (needs
// only the record itself and tanAOV: it can be repeated later with another tanAOV while the frame's planes are long gone).
void rd_gtail_run(unsigned char *blob, size_t blobBytes, const linesegment_t *ls, const int *segid, const int *votes, int *table, unsigned char *scratch,
                  size_t scratchBytes, int iw, int ih, double tanAOV, int phases, int nb, size_t fs, cudaStream_t s) {
  GtArgs a;
  a.ls = (const GtLS *)ls; a.segid = segid; a.votes = votes; a.table = table; a.scratch = scratch; a.blob = blob;
  a.scratchBytes = scratchBytes; a.blobBytes = blobBytes; a.iw = iw; a.ih = ih; a.nentry = iw * ih * 4 / 5; a.fs = fs;
  if (phases & 1) {
    RD_CUDA(cudaMemset2DAsync(blob + ((blobBytes / 2) & ~(size_t)15), fs ? fs : sizeof(GtHdr), 0, sizeof(GtHdr), nb, s));
    RD_LAUNCH(kt_samples, dim3(16, nb), GT_TPB, 0, s, a);
    RD_LAUNCH(kt_regions, dim3(32, nb), GT_TPB, 0, s, a);
    RD_LAUNCH(kt_members, dim3(32, nb), GT_TPB, 0, s, a);
    RD_LAUNCH(kt_order, nb, 256, 0, s, a);
    RD_LAUNCH(kt_quad, dim3(32, nb), GTQ_WARPS * 32, 0, s, a);
  }
  if (phases & 2) {
    RD_LAUNCH(kt_pose, dim3(8, nb), 64, 0, s, a, tanAOV);
    RD_LAUNCH(kt_finish, nb, 256, 0, s, a, tanAOV);
  }
}
