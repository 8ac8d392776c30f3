// rd_despeckle2.cu - despeckle2 (oclrect.cl:348-371, oclrect.c:336) with the reference's in-place semantics under its raster
// schedule, exactly, and markBoundary (oclrect.cl:373-390) behind it.  See rd_despeckle2.cuh for the formulation.
//
//   kd2_pre  : fully parallel, one CTA per image row.  Pixels of large regions are final (copied to `dst`); every small-region
//              pixel gets its static record (best known candidate + which causal neighbours are small), compacted per row in
//              ascending x into three list planes, row y at offset y * iw.  rowcnt[y] = number of small pixels of the row.
//   kd2_seq  : the dependent part, one WARP per frame walking the rows top-down, 32 list entries per step.  Everything on the
//              critical path lives in shared memory / registers: the (label, size) pairs the previous row's small pixels ended
//              up with (rowbuf, two rows), the composed maps of a run of small pixels (warp shuffle scan, as many doubling steps
//              as the longest run of the chunk needs), the carry into the next chunk.  The list records are streamed into a
//              shared-memory ring D2_RING chunks ahead with cp.async, so the HBM/L2 latency is off the critical path.
//              Cost: ~0.15 us per chunk, ih + (small pixels)/32 chunks per frame; frames of a batch run side by side.
//   kf_markBoundary : tile kernel on the final labels.
#include "rd_common.cuh"
#include "rd_despeckle2.cuh"
#include "rd_tma.cuh"
#include <mutex>

#define D2P_THREADS 256
__global__ void __launch_bounds__(D2P_THREADS) kd2_pre(int *dst, int *list, int *recL, int *recS, int *rowcnt, const int *label, const int *size, int thre,
                                                       int iw, int ih, size_t fs) {
  rd_batch_y(fs, dst, list, recL, recS, rowcnt, label, size);
  __shared__ int wsum[D2P_THREADS / 32];
  __shared__ int base;
  const int y = blockIdx.x, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const size_t row = (size_t)y * iw;
  for (int x0 = 0; x0 < iw; x0 += D2P_THREADS) {
    const int x = x0 + threadIdx.x;
    bool small = false;
    int rec = 0, bl = 0, bs = 0;
    if (x < iw) {
      const int l = label[row + x];
      small = !(size[l] > thre);
      if (!small) dst[row + x] = l;
      else rec = d2_static(x, y, label, size, thre, iw, ih, bl, bs);
    }
    const unsigned b = __ballot_sync(0xffffffffu, small);
    if (lane == 0) wsum[wp] = __popc(b);
    __syncthreads();
    int off = base;
    for (int w = 0; w < wp; w++) off += wsum[w];
    if (small) {
      const size_t o = row + off + __popc(b & ((1u << lane) - 1u));
      list[o] = rec; recL[o] = bl; recS[o] = bs;
    }
    __syncthreads();
    if (threadIdx.x == 0) { int t = base; for (int w = 0; w < D2P_THREADS / 32; w++) t += wsum[w]; base = t; }
    __syncthreads();
  }
  if (threadIdx.x == 0) rowcnt[y] = base;
}

#define D2_RING 8
#define D2_SMEM_MAX (200 * 1024)
__device__ __forceinline__ void d2_cp_async4(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void d2_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void d2_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(D2_RING - 1) : "memory"); }

// the plain form of the walk: cursors over the per-row lists, one chunk per iteration.  ring [D2_RING][3][32], cnt [ih] (filled), rowbuf [2][iw]
__device__ __noinline__ void d2_seq_generic(int *dst, const int *list, const int *recL, const int *recS, int *ring, const int *cnt, int2 *rowbuf, bool rowbuf_global,
                                            int iw, int ih) {
  const int lane = threadIdx.x;
  int2 *rowbuf_g = rowbuf_global ? rowbuf : (int2 *)NULL;
  // fetch cursor (warp-uniform): the next chunk to stream in is entries [fc, fc + 32) of row fy
  int fy = 0, fc = 0;
  while (fy < ih && cnt[fy] == 0) fy++;
  auto fetch = [&](int stage) {
    if (fy < ih) {
      const int j = fc + lane;
      if (j < cnt[fy]) {
        const size_t o = (size_t)fy * iw + j;
        int *r = ring + stage * 96 + lane;
        d2_cp_async4(r, list + o); d2_cp_async4(r + 32, recL + o); d2_cp_async4(r + 64, recS + o);
      }
      fc += D2_CHUNK;
      if (fc >= cnt[fy]) { fc = 0; do fy++; while (fy < ih && cnt[fy] == 0); }
    }
    d2_cp_commit();
  };
  for (int s = 0; s < D2_RING; s++) fetch(s);
  int py = 0, pc = 0, stage = 0;
  while (py < ih && cnt[py] == 0) py++;
  int carryL = 0, carryS = 0;
  while (py < ih) {
    d2_cp_wait();
    __syncwarp();
    const int n = cnt[py];
    const bool valid = pc + lane < n;
    const int *r = ring + stage * 96 + lane;
    const int rec = valid ? r[0] : 0;
    int bl = valid ? r[32] : 0, bs = valid ? r[64] : 0;
    __syncwarp();
    fetch(stage);
    stage = stage + 1 == D2_RING ? 0 : stage + 1;
    const int x = rec & 0xffff, dyn = (rec >> 20) & 15;
    int code = (rec >> 16) & 15;
    const int2 *above = rowbuf + (size_t)((py + 1) & 1) * iw;
    if (dyn & 1) { const int2 v = above[x - 1]; d2_take(bl, bs, code, v.x, v.y, 1); }
    if (dyn & 2) { const int2 v = above[x]; d2_take(bl, bs, code, v.x, v.y, 2); }
    if (dyn & 4) { const int2 v = above[x + 1]; d2_take(bl, bs, code, v.x, v.y, 3); }
    int T = D2_HEAD;
    if (dyn & 8) T = d2_threshold(bs, code);
    if (lane == 0 && T != D2_HEAD) {                          // the run continues from the previous chunk
      if (carryS >= T) { bl = carryL; bs = carryS; }
      T = D2_HEAD;
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      if (__all_sync(0xffffffffu, T == D2_HEAD)) break;
      const int lL = __shfl_up_sync(0xffffffffu, bl, d), lS = __shfl_up_sync(0xffffffffu, bs, d), lT = __shfl_up_sync(0xffffffffu, T, d);
      if (lane >= d && T != D2_HEAD) d2_compose(bl, bs, T, lL, lS, lT);
    }
    if (valid) {
      dst[(size_t)py * iw + x] = bl;
      rowbuf[(size_t)(py & 1) * iw + x] = make_int2(bl, bs);
    }
    carryL = __shfl_sync(0xffffffffu, bl, 31); carryS = __shfl_sync(0xffffffffu, bs, 31);
    pc += D2_CHUNK;
    if (pc >= n) { pc = 0; do py++; while (py < ih && cnt[py] == 0); }
    if (rowbuf_g) __threadfence_block();
    __syncwarp();
  }
}
// rowbuf_g != NULL: the two-row buffer lives in global memory (frames too wide for shared memory)
__global__ void __launch_bounds__(32) kd2_seq(int *dst, const int *list, const int *recL, const int *recS, const int *rowcnt, int2 *rowbuf_g, int iw, int ih,
                                              size_t fs) {
  rd_batch_x(fs, dst, list, recL, recS, rowcnt);
  if (rowbuf_g) rd_batch_x(fs, rowbuf_g);
  extern __shared__ __align__(16) int d2_smem[];
  int *ring = d2_smem;                                        // [D2_RING][3][32]
  int *cnt = ring + D2_RING * 3 * 32;                         // [ih]
  int2 *rowbuf = rowbuf_g ? rowbuf_g : (int2 *)(cnt + ((ih + 1) & ~1));   // [2][iw]
  for (int i = threadIdx.x; i < ih; i += 32) cnt[i] = rowcnt[i];
  __syncwarp();
  d2_seq_generic(dst, list, recL, recS, ring, cnt, rowbuf, rowbuf_g != NULL, iw, ih);
}

// The same walk, software-pipelined (the version the schedule runs; kd2_seq above stays as the fall-back for frames with more
// chunks than the descriptor table holds).  A chunk's critical path is: the previous chunk's row-buffer stores -> __syncwarp ->
// row-buffer loads -> merge -> vote (-> scan) -> stores; everything else is taken off it:
//   * a table of chunk descriptors (row | chunk-in-row << 16) is built once, in parallel, in shared memory, so the cursors need
//     no loops or row-count look-ups;
//   * the records of chunk k+1 are moved from the cp.async ring to registers while chunk k computes; the ring is filled
//     D2_RING chunks ahead.
#define D2_NCH_MAX 12288
__global__ void __launch_bounds__(32) kd2_seq_pipe(int *dst, const int *list, const int *recL, const int *recS, const int *rowcnt, int iw, int ih, int nch_cap,
                                                   size_t fs) {
  rd_batch_x(fs, dst, list, recL, recS, rowcnt);
  extern __shared__ __align__(16) int d2_smem[];
  int *ring = d2_smem;                                        // [D2_RING][3][32]
  int *cnt = ring + D2_RING * 3 * 32;                         // [ih]
  int2 *rowbuf = (int2 *)(cnt + ((ih + 1) & ~1));             // [2][iw]
  int *desc = (int *)(rowbuf + 2 * (size_t)iw);               // [nch_cap]
  const int lane = threadIdx.x;
  // descriptors: row y contributes ceil(cnt[y] / 32) chunks
  int nch = 0;
  for (int y0 = 0; y0 < ih; y0 += 32) {
    const int y = y0 + lane;
    const int c = y < ih ? rowcnt[y] : 0;
    if (y < ih) cnt[y] = c;
    const int k = (c + 31) >> 5;
    int incl = k;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    const int base = nch + incl - k;
    for (int j = 0; j < k; j++) if (base + j < nch_cap) desc[base + j] = y | (j << 16);
    nch += __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncwarp();
  if (nch > nch_cap) { d2_seq_generic(dst, list, recL, recS, ring, cnt, rowbuf, false, iw, ih); return; }   // more chunks than the table holds
  auto fetch = [&](int k) {                                   // chunk k -> ring stage k % D2_RING
    if (k < nch) {
      const int d = desc[k], y = d & 0xffff, j = ((d >> 16) << 5) + lane;
      if (j < cnt[y]) {
        const size_t o = (size_t)y * iw + j;
        int *r = ring + (k % D2_RING) * 96 + lane;
        d2_cp_async4(r, list + o); d2_cp_async4(r + 32, recL + o); d2_cp_async4(r + 64, recS + o);
      }
    }
    d2_cp_commit();
  };
  for (int k = 0; k < D2_RING; k++) fetch(k);
  // chunk 0 -> registers
  d2_cp_wait();
  __syncwarp();
  int n_y = 0, n_rec = 0, n_bl = 0, n_bs = 0;
  bool n_valid = false;
  auto stage_in = [&](int k) {
    n_valid = false; n_rec = 0; n_bl = 0; n_bs = 0;
    if (k < nch) {
      const int d = desc[k];
      n_y = d & 0xffff;
      n_valid = ((d >> 16) << 5) + lane < cnt[n_y];
      const int *r = ring + (k % D2_RING) * 96 + lane;
      if (n_valid) { n_rec = r[0]; n_bl = r[32]; n_bs = r[64]; }
    }
  };
  stage_in(0);
  int carryL = 0, carryS = 0;
  for (int k = 0; k < nch; k++) {
    const int py = n_y, rec = n_rec;
    int bl = n_bl, bs = n_bs;
    const bool valid = n_valid;
    const int x = rec & 0xffff, dyn = (rec >> 20) & 15;
    int code = (rec >> 16) & 15;
    // critical path, part 1: what the row above ended up with
    const int2 *above = rowbuf + (size_t)((py + 1) & 1) * iw;
    int2 v0 = make_int2(0, 0), v1 = v0, v2 = v0;
    if (dyn & 1) v0 = above[x - 1];
    if (dyn & 2) v1 = above[x];
    if (dyn & 4) v2 = above[x + 1];
    // off the critical path: chunk k+1 -> registers, chunk k+D2_RING -> ring
    asm volatile("cp.async.wait_group %0;" ::"n"(D2_RING - 2) : "memory");
    __syncwarp();
    stage_in(k + 1);
    __syncwarp();
    fetch(k + D2_RING);
    // critical path, part 2
    if (dyn & 1) d2_take(bl, bs, code, v0.x, v0.y, 1);
    if (dyn & 2) d2_take(bl, bs, code, v1.x, v1.y, 2);
    if (dyn & 4) d2_take(bl, bs, code, v2.x, v2.y, 3);
    int T = D2_HEAD;
    if (dyn & 8) T = d2_threshold(bs, code);
    if (lane == 0 && T != D2_HEAD) {                          // the run continues from the previous chunk
      if (carryS >= T) { bl = carryL; bs = carryS; }
      T = D2_HEAD;
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      if (__all_sync(0xffffffffu, T == D2_HEAD)) break;
      const int lL = __shfl_up_sync(0xffffffffu, bl, d), lS = __shfl_up_sync(0xffffffffu, bs, d), lT = __shfl_up_sync(0xffffffffu, T, d);
      if (lane >= d && T != D2_HEAD) d2_compose(bl, bs, T, lL, lS, lT);
    }
    if (valid) {
      rowbuf[(size_t)(py & 1) * iw + x] = make_int2(bl, bs);
      dst[(size_t)py * iw + x] = bl;
    }
    carryL = __shfl_sync(0xffffffffu, bl, 31); carryS = __shfl_sync(0xffffffffu, bs, 31);
    __syncwarp();
  }
}

// The walk with a producer warp (the version the schedule runs).  A single in-order warp pays every dependent latency of its
// instruction stream, also those that are not on the data-dependent path (cursor look-ups, ring bookkeeping, cp.async issue):
// kd2_seq_pipe still needs ~830 cycles per chunk.  Here warp 1 streams the records of chunk k into a ring of D2_STAGES stages
// (cp.async, D2_RING groups in flight, invalid lanes marked -1) and publishes a sequence number; warp 0 only does the dependent part:
// records -> registers, row-buffer look-ups, merge, vote / scan, stores.  Sequence numbers in shared memory (release: fence + store by
// lane 0, acquire: volatile load + fence) - no CTA barrier inside the loop.
#define D2_STAGES 16
__global__ void __launch_bounds__(64) kd2_seq_pc(int *dst, const int *list, const int *recL, const int *recS, const int *rowcnt, int iw, int ih, int nch_cap,
                                                 size_t fs) {
  rd_batch_x(fs, dst, list, recL, recS, rowcnt);
  extern __shared__ __align__(16) int d2_smem[];
  int *ring = d2_smem;                                        // [D2_STAGES][3][32]  (the generic walk uses the first D2_RING stages)
  int *cnt = ring + D2_STAGES * 96;                           // [ih]
  int2 *rowbuf = (int2 *)(cnt + ((ih + 1) & ~1));             // [2][iw]
  int *desc = (int *)(rowbuf + 2 * (size_t)iw);               // [nch_cap]
  __shared__ volatile int prod, cons;
  __shared__ int nch_s;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  if (wp == 0) {
    int nch = 0;
    for (int y0 = 0; y0 < ih; y0 += 32) {
      const int y = y0 + lane;
      const int c = y < ih ? rowcnt[y] : 0;
      if (y < ih) cnt[y] = c;
      const int k = (c + 31) >> 5;
      int incl = k;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
      const int base = nch + incl - k;
      for (int j = 0; j < k; j++) if (base + j < nch_cap) desc[base + j] = y | (j << 16);
      nch += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) { nch_s = nch; prod = 0; cons = 0; }
  }
  __syncthreads();
  const int nch = nch_s;
  if (nch > nch_cap) {                                        // more chunks than the table holds: the plain walk
    if (wp == 0) d2_seq_generic(dst, list, recL, recS, ring, cnt, rowbuf, false, iw, ih);
    return;
  }
  if (wp == 1) {
    // ---------------- producer
    for (int k = 0; k < nch + D2_RING - 1; k++) {
      if (k < nch) {
        while (k - cons >= D2_STAGES) { }                    // stage still in use
        const int d = desc[k], y = d & 0xffff, j = ((d >> 16) << 5) + lane;
        int *r = ring + (k % D2_STAGES) * 96 + lane;
        if (j < cnt[y]) {
          const size_t o = (size_t)y * iw + j;
          d2_cp_async4(r, list + o); d2_cp_async4(r + 32, recL + o); d2_cp_async4(r + 64, recS + o);
        } else {
          r[0] = -1;
        }
      }
      d2_cp_commit();
      if (k >= D2_RING - 1) {                                 // the group of chunk k - (D2_RING - 1) has landed
        d2_cp_wait();
        __syncwarp();
        __threadfence_block();
        if (lane == 0) prod = k - (D2_RING - 1) + 1;
      }
    }
    return;
  }
  // ---------------- consumer
  int carryL = 0, carryS = 0;
  for (int k = 0; k < nch; k++) {
    while (prod <= k) { }
    __threadfence_block();
    const int *r = ring + (k % D2_STAGES) * 96 + lane;
    const int rec = r[0];
    int bl = r[32], bs = r[64];
    const int py = desc[k] & 0xffff;
    __syncwarp();
    if (lane == 0) cons = k + 1;                              // the records are in registers: the stage may be refilled
    const bool valid = rec >= 0;
    const int x = rec & 0xffff, dyn = valid ? (rec >> 20) & 15 : 0;
    int code = (rec >> 16) & 15;
    const int2 *above = rowbuf + (size_t)((py + 1) & 1) * iw;
    int2 v0 = make_int2(0, 0), v1 = v0, v2 = v0;
    if (dyn & 1) v0 = above[x - 1];
    if (dyn & 2) v1 = above[x];
    if (dyn & 4) v2 = above[x + 1];
    if (!valid) { bl = 0; bs = 0; }
    if (dyn & 1) d2_take(bl, bs, code, v0.x, v0.y, 1);
    if (dyn & 2) d2_take(bl, bs, code, v1.x, v1.y, 2);
    if (dyn & 4) d2_take(bl, bs, code, v2.x, v2.y, 3);
    int T = D2_HEAD;
    if (dyn & 8) T = d2_threshold(bs, code);
    if (lane == 0 && T != D2_HEAD) {                          // the run continues from the previous chunk
      if (carryS >= T) { bl = carryL; bs = carryS; }
      T = D2_HEAD;
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      if (__all_sync(0xffffffffu, T == D2_HEAD)) break;
      const int lL = __shfl_up_sync(0xffffffffu, bl, d), lS = __shfl_up_sync(0xffffffffu, bs, d), lT = __shfl_up_sync(0xffffffffu, T, d);
      if (lane >= d && T != D2_HEAD) d2_compose(bl, bs, T, lL, lS, lT);
    }
    if (valid) {
      rowbuf[(size_t)(py & 1) * iw + x] = make_int2(bl, bs);
      dst[(size_t)py * iw + x] = bl;
    }
    carryL = __shfl_sync(0xffffffffu, bl, 31); carryS = __shfl_sync(0xffffffffu, bs, 31);
    __syncwarp();
  }
}

// markBoundary (oclrect.cl:373-390): a pixel keeps its region label if its 5x5 window holds another label; 2-px frame -> -1
#define MB_T 32
#define MB_A 2
#define MB_W (MB_T + 2 * MB_A)
// USE_TMA: interior CTAs fetch the tile as one box of MB_RAWW columns starting at bx - 4 (a multiple of four: rd_tma.cuh)
#define MB_RAWW 40
#define MB_RAWX 4
template <bool USE_TMA>
__global__ void __launch_bounds__(256) kf_markBoundary_t(int *out, const int *in, const __grid_constant__ CUtensorMap map, int iw, int ih, size_t fs) {
  rd_batch_z(fs, out, in);
  __shared__ __align__(128) int sd[USE_TMA ? MB_W * MB_RAWW : MB_W * MB_W];
  __shared__ __align__(8) uint64_t bar;
  const int bx = blockIdx.x * MB_T - MB_A, by = blockIdx.y * MB_T - MB_A;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const bool interior = USE_TMA && bx + MB_A - MB_RAWX >= 0 && by >= 0 && bx + MB_W <= iw && by + MB_W <= ih;
  const int pitch = interior ? MB_RAWW : MB_W, xoff = interior ? MB_RAWX - MB_A : 0;      // tile column t lives at sd[row * pitch + t + xoff]
  if (interior) {
    if (tid == 0) rd_mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
      rd_mbar_expect(&bar, (unsigned)(MB_W * MB_RAWW * 4));
      rd_tma_load3(sd, &map, bx + MB_A - MB_RAWX, by, (int)blockIdx.z, &bar);
    }
    rd_mbar_wait(&bar, 0);
  } else {
    for (int i = tid; i < MB_W * MB_W; i += 256) {
      const int gx = bx + i % MB_W, gy = by + i / MB_W;
      sd[i] = (gx >= 0 && gx < iw && gy >= 0 && gy < ih) ? in[(size_t)gy * iw + gx] : -1;
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int tx = MB_A + threadIdx.x, ty = MB_A + threadIdx.y + k * 8;
    const int gx = bx + tx, gy = by + ty;
    if (gx >= iw || gy >= ih) continue;
    int r = -1;
    if (!(gx <= 1 || gy <= 1 || gx >= iw - 2 || gy >= ih - 2)) {
      const int i = ty * pitch + tx + xoff, c0 = sd[i];
      bool nearEdge = false;
#pragma unroll
      for (int yy = -2; yy <= 2; yy++)
#pragma unroll
        for (int xx = -2; xx <= 2; xx++) nearEdge |= sd[i + yy * pitch + xx] != c0;
      if (nearEdge) r = c0;
    }
    out[(size_t)gy * iw + gx] = r;
  }
}

// dst = despeckle2(label) (dst != label; label is left untouched).  list / recL / recS: scratch planes of iw*ih ints each;
// rowcnt: ih ints; rowbuf: 2*iw int2 of scratch, only used when the frame is too wide for shared memory.
void rd_despeckle2_run(int *dst, const int *label, const int *size, int *list, int *recL, int *recS, int *rowcnt, int2 *rowbuf, int thre, int iw, int ih,
                       int nb, size_t fs, cudaStream_t s) {
  if (iw > 0xffff) exitf(-1, "rectdetect_b200: despeckle2: frames wider than 65535 pixels are not supported\n");
  RD_LAUNCH(kd2_pre, dim3(ih, nb), D2P_THREADS, 0, s, dst, list, recL, recS, rowcnt, label, size, thre, iw, ih, fs);
  const size_t fixed = (size_t)D2_RING * 96 * 4 + (size_t)((ih + 1) & ~1) * 4, full = fixed + (size_t)2 * iw * sizeof(int2);
  static std::mutex mu;
  static bool ready[64] = {false};
  {                                                   // opt in to large dynamic shared memory once per device
    int dev = 0;
    RD_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> g(mu);
    if (dev >= 64 || !ready[dev]) {
      RD_CUDA(cudaFuncSetAttribute(kd2_seq, cudaFuncAttributeMaxDynamicSharedMemorySize, D2_SMEM_MAX));
      RD_CUDA(cudaFuncSetAttribute(kd2_seq_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, D2_SMEM_MAX));
      RD_CUDA(cudaFuncSetAttribute(kd2_seq_pc, cudaFuncAttributeMaxDynamicSharedMemorySize, D2_SMEM_MAX));
      if (dev < 64) ready[dev] = true;
    }
  }
  // descriptor table: room for every row plus a generous share of extra chunks; a frame that turns out to need more (decided on the
  // device) takes the plain walk inside the same kernel; frames too wide / tall for shared memory take kd2_seq
  int nch_cap = 2 * ih + 1024;
  if (nch_cap > D2_NCH_MAX) nch_cap = D2_NCH_MAX;
  const size_t pipe = full + (size_t)nch_cap * 4, pc = pipe + (size_t)(D2_STAGES - D2_RING) * 96 * 4;
  static const bool generic_only = getenv("RD_D2_GENERIC") != NULL, one_warp = getenv("RD_D2_PIPE") != NULL;
  if (!generic_only && !one_warp && pc <= D2_SMEM_MAX) {
    RD_LAUNCH(kd2_seq_pc, nb, 64, pc, s, dst, list, recL, recS, rowcnt, iw, ih, nch_cap, fs);
    return;
  }
  if (!generic_only && pipe <= D2_SMEM_MAX) {
    RD_LAUNCH(kd2_seq_pipe, nb, 32, pipe, s, dst, list, recL, recS, rowcnt, iw, ih, nch_cap, fs);
    return;
  }
  const bool in_smem = full <= D2_SMEM_MAX;
  const size_t smem = in_smem ? full : fixed;
  if (smem > D2_SMEM_MAX) exitf(-1, "rectdetect_b200: despeckle2: frame too tall (%d rows)\n", ih);
  RD_LAUNCH(kd2_seq, nb, 32, smem, s, dst, list, recL, recS, rowcnt, in_smem ? (int2 *)NULL : rowbuf, iw, ih, fs);
}
void rd_markBoundary_run(int *out, const int *in, int iw, int ih, int nb, size_t fs, cudaStream_t s) {
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (rd_tma_ok(in, iw, fs) && rd_tma_make_map(&map, in, CU_TENSOR_MAP_DATA_TYPE_INT32, iw, ih, nb, fs, MB_RAWW, MB_W)) {
    RD_LAUNCH(kf_markBoundary_t<true>, dim3(rd_cdiv(iw, MB_T), rd_cdiv(ih, MB_T), nb), dim3(32, 8), 0, s, out, in, map, iw, ih, fs);
    return;
  }
  RD_LAUNCH(kf_markBoundary_t<false>, dim3(rd_cdiv(iw, MB_T), rd_cdiv(ih, MB_T), nb), dim3(32, 8), 0, s, out, in, map, iw, ih, fs);
}
