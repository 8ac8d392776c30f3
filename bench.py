#!/usr/bin/env python3
"""bench.py - throughput of the rectangle-detection hot path on B200 (one process per GPU).

metric  : Mpix/s = frames/s x iw x ih, BGR frame -> rect_t list (BASELINE.json)
step    : one batch of --frames (default 256) 1280x720 synthetic frames per GPU (config "vidrect 1280x720 synthetic stream, AOV 72,
          batched on 1xB200"); frames of a batch are independent and are sharded across ranks (weak scaling, no data-path
          collective; one gather of the rect lists to rank 0 per step).
value   : frames already resident in HBM; every device stage incl. the device tail (executeCPUTask on the GPU) and the read-back of
          the rect lists, plus the rect-list gather for N > 1.
e2e     : the same through the C-ABI batch call with frames in pinned HOST memory (H2D copies inside the timed region).
roofline: the kernel with the largest share of device time against the measured HBM copy bandwidth.  Per-kernel times come
          from CUDA events around every launch on the launching stream (in the library), in a separate pass through ONE
          pipeline object so that each kernel is timed alone; per-stage (A/B/C/D) totals of the same pass are reported too.
sub-records of the same JSON line: config4 (1920x1080, 512 frames per step over all ranks: strong scaling), config5 (one 3840x2160 frame,
          per-stage roofline), api_stream (the reference's enqueue/poll API, unchanged caller), merge_replay (cost of the exact first-pass
          replay of labelMergeMain, off by default), parity_checked_frames (frames of the
          timed batch compared with the CPU oracle after the timed region).
--impl reference : THE REFERENCE ITSELF on the host cores - oracle/_ref/librd_ref.so = its unmodified host code and its OpenCL C
          kernels compiled as C++ (oracle/Makefile target _ref), one instance per core on a bounded sample of the stream per step;
          the CPU oracle port is timed beside it (cpu_port).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "Mpix/s (frames/s x WxH) 1280x720 rect detect"
TAN_AOV = math.tan(math.radians(36.0))

# algorithmic (compulsory) bytes per pixel of each kernel: planes it must read + planes it must write, 4 B each unless
# noted (DESIGN.md "Kernels").  Used for roofline.achieved of whichever kernel dominates the step.
KERNEL_BYTES_PER_PX = {
    # production schedule
    "kf_bgr2plab4": 7, "kf_bgr2plab1": 7, "kf_iir_h3": 28, "kf_iir_mid3": 40, "kf_iir_v3": 36, "kf_iir_fin3": 52, "kf_edge_thin": 12,
    "kb_strings1": 5, "k_ccl_tile<LinkFn>": 9, "k_ccl_seams": 1, "k_ccl_roots": 1, "k_ccl_flatten": 9, "k_ccl_flatten_list": 9, "k_ccl_flatten_merge": 12,
    "kr_calcStrength": 8, "kf_filter_masks": 9, "kf_blb_extents_s": 3, "kf_blb_stream4<npx>": 10, "kf_blb_stream": 10, "kf_quant_despeckle": 12,
    "kb_junction_mask": 9, "kf_calcSize": 4, "kf_despeckle2_boundary": 8, "kb_strings2": 9, "k_clear4": 4, "k_clear": 4, "k_copy4": 8,
}


def ncu_traffic(kernel):
    """DRAM bytes per FRAME of `kernel` from the committed ncu --set full capture (profiles/ncu_traffic.json,
    dram__bytes_read.sum + dram__bytes_write.sum per launch / frames per launch), or None"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        e = t["kernels"].get(kernel)
        return (float(e["dram_bytes_per_frame"]), t.get("source")) if e else (None, None)
    except Exception:
        return None, None


def synth_batch(iw, ih, first_seed, count, pinned):
    """`count` frames of the synthetic stream (seeds first_seed + i) in one (pinned) host tensor; generated on all host cores"""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from rectdetect_b200.synth import synth_frame, synth_lib
    synth_lib()
    ws = 3 * iw
    t = torch.empty((count, ih, ws), dtype=torch.uint8, pin_memory=pinned)
    a = t.numpy()
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:      # the generator is C++ behind ctypes: the GIL is released
        list(ex.map(lambda i: synth_frame(iw, ih, first_seed + i, out=a[i]), range(count)))
    return t


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md)"""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False
        self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_mpix(iw, ih, seeds, threads=None):
    """CPU restatement of the reference schedule on `threads` host cores (all by default): -> (Mpix/s, cores, seconds)"""
    import oracle_lib as ol
    L = ol.oracle()
    if threads:
        L.ora_set_threads(threads)
    cores = L.ora_get_threads()
    o = ol.OracleRect(iw, ih)
    frames = [ol.synth_frame(iw, ih, s) for s in seeds]
    o.execute_once(frames[0], TAN_AOV)                 # warm-up (page faults, OpenMP pool)
    t0 = time.perf_counter()
    for f in frames:
        o.execute_once(f, TAN_AOV)
    dt = time.perf_counter() - t0
    o.close()
    return len(frames) * iw * ih / dt / 1e6, cores, dt


REF_SO = os.path.join(ROOT, "oracle", "_ref", "librd_ref.so")
REF_KIND = ("the reference itself (oracle/_ref/librd_ref.so: its unmodified host code + its OpenCL C kernels compiled as C++ with g++ -O2, "
            "220 launches per frame, oclrect_enqueueTask/pollTask), one single-threaded instance per host core, the stream split among them")


def cpu_reference_mpix(iw, ih, first_seed, frames, workers=None):
    """THE REFERENCE on the host cores: `workers` processes (all cores by default), each running the reference's own pipeline
    on its share of `frames` frames (tests/ref_worker.py); wall time from the common start to the last "done".
    -> (Mpix/s, workers, seconds)"""
    workers = max(1, min(workers or os.cpu_count() or 1, frames))
    if os.path.exists("/root/reference/oclrect.c"):          # build container: make sure oracle/_ref is current, once, before the workers start
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_ref"], stdout=subprocess.DEVNULL)
    share = [frames // workers + (1 if i < frames % workers else 0) for i in range(workers)]
    procs, seed = [], first_seed
    for n in share:
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "ref_worker.py"), str(iw), str(ih), str(seed), str(n)],
                                      stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, env=dict(os.environ, OMP_NUM_THREADS="1", RD_REF_NO_BUILD="1")))
        seed += n
    for p in procs:
        if p.stdout.readline().strip() != "ready":
            raise RuntimeError("reference worker failed to start")
    t0 = time.perf_counter()
    for p in procs:
        p.stdin.write("go\n")
        p.stdin.flush()
    for p in procs:
        if not p.stdout.readline().startswith("done"):
            raise RuntimeError("reference worker failed")
    dt = time.perf_counter() - t0
    for p in procs:
        p.wait()
    return frames * iw * ih / dt / 1e6, workers, dt


def run_reference(args, rank, world):
    """--impl reference: rank 0 times the reference's own CPU run (oracle/_ref/librd_ref.so; the oracle port when that library
    is absent), the other ranks exit without work"""
    if rank != 0:
        return
    iw, ih = args.w, args.h
    sample = args.ref_frames
    have_ref = os.path.exists(REF_SO)
    run = (lambda first, n: cpu_reference_mpix(iw, ih, first, n)) if have_ref else (lambda first, n: cpu_oracle_mpix(iw, ih, [first + i for i in range(n)]))
    for _ in range(args.warmup):
        run(1000, min(sample, os.cpu_count() or 1))
    vals, secs = [], []
    for s in range(args.steps):
        v, cores, sec = run(1000 + s * sample, sample)
        vals.append(v)
        secs.append(sec)
    dt = float(np.sum(secs))                      # the timed regions (worker start-up and warm-up frames are outside)
    value = sample * iw * ih * args.steps / dt / 1e6
    port_v, port_cores, _ = cpu_oracle_mpix(iw, ih, [1000 + i for i in range(min(sample, 8))])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+i32 (f64 host tail)",
        "data": "synthetic",
        "config": {"workload": "vidrect %dx%d synthetic stream, AOV 72, full imgutil->polyline->rect pipeline" % (iw, ih), "frames_per_step": sample,
                   "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": cores, "kind": "reference" if have_ref else "port",
                         "sample": "%d frames of the workload per step; %s" % (sample, REF_KIND if have_ref else "CPU oracle (OpenMP restatement); oracle/_ref/librd_ref.so not found")},
        "cpu_port": {"value": port_v, "unit": "Mpix/s", "cores": port_cores, "kind": "port",
                     "sample": "the CPU oracle (tuned OpenMP restatement of the same schedule with exact connected components) on the same frames, for comparison"},
        "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_JSON_FD = None


def emit(line):
    """the ONE JSON line of the contract, on the process's real stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # libraries print to stdout behind Python's back (NCCL: "NCCL version ..." from the first communicator): keep fd 1 for the
    # JSON line alone and send everything else to stderr
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--w", type=int, default=1280)
    ap.add_argument("--h", type=int, default=720)
    ap.add_argument("--frames", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--nctx", type=int, default=16, help="pipeline objects (streams) per GPU")
    ap.add_argument("--fpl", type=int, default=8, help="frames per kernel launch")
    ap.add_argument("--ref-frames", type=int, default=os.cpu_count() or 8, help="frames per step of the CPU reference arm (default: one per host core)")
    ap.add_argument("--cpu-frames", type=int, default=2 * (os.cpu_count() or 8), help="frames of the cpu_baseline sample (default: two per host core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline line only: no config4 / config5 / api_stream sub-records")
    ap.add_argument("--config4-frames", type=int, default=512, help="global frames per step of the 1920x1080 sub-record (BASELINE.json config 4)")
    ap.add_argument("--parity-frames", type=int, default=4, help="frames of the timed batch checked against the CPU oracle after the timed region")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import rectdetect_b200 as rd
    from rectdetect_b200 import dist as rdist

    if not torch.cuda.is_available() or rd.device_count() <= local_rank:
        raise SystemExit("bench.py: no CUDA device for rank %d; rectdetect_b200 has no CPU fallback" % rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = "cuda"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_workload(iw, ih, total_frames, first_seed, steps, warmup):
        """one workload: `total_frames` frames per step over all ranks (rank r owns a contiguous shard), value (frames resident
        in HBM) and e2e (pinned host frames, H2D inside the timed region); -> dict (timings are max over ranks)"""
        ws = 3 * iw
        lo, hi = rdist.shard_range(total_frames, world, rank)
        host_frames = synth_batch(iw, ih, first_seed + lo, hi - lo, pinned=True)
        dev_frames = host_frames.to("cuda", non_blocking=False)
        frame_bytes = ih * ws
        batch = rd.Batch(local_rank, iw, ih, nctx=args.nctx, frames_per_launch=args.fpl)

        def gather(rects):
            torch.cuda.set_device(local_rank)
            return rdist.gather_rect_lists(lo, rects, total_frames, device=dev)

        def step(on_device):
            ptr = dev_frames.data_ptr() if on_device else host_frames.data_ptr()
            rects = batch.run(ptr, frame_bytes, ws, hi - lo, TAN_AOV, on_device=on_device)
            return gather(rects)

        from concurrent.futures import ThreadPoolExecutor as _Pool
        gather_pool = _Pool(max_workers=1) if world > 1 else None

        def timed(on_device, nsteps):
            barrier()
            l0 = rd.kernel_launches()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = None
            if world > 1:
                # the gather of step k (NCCL, rank 0 unpacks) runs beside the frames of step k + 1; the last one is inside the timed region too
                fut = None
                for _ in range(nsteps):
                    ptr = dev_frames.data_ptr() if on_device else host_frames.data_ptr()
                    rects = batch.run(ptr, frame_bytes, ws, hi - lo, TAN_AOV, on_device=on_device)
                    if fut is not None:
                        out = fut.result()
                    fut = gather_pool.submit(gather, rects)
                out = fut.result()
            else:
                for _ in range(nsteps):
                    out = step(on_device)
            barrier()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            launches = torch.tensor([rd.kernel_launches() - l0], dtype=torch.int64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dist.all_reduce(launches, op=dist.ReduceOp.SUM)
            return float(t.item()), int(launches.item()), out

        for _ in range(warmup):
            step(True)
            step(False)
        ms_val, launches, rects = timed(True, steps)
        ms_e2e, _, rects_e2e = timed(False, steps)
        wait_ms, list_ms = batch.stage_ms()
        pix = total_frames * iw * ih
        res = {"iw": iw, "ih": ih, "ws": ws, "lo": lo, "hi": hi, "total_frames": total_frames, "frame_bytes": frame_bytes, "host_frames": host_frames, "dev_frames": dev_frames,
               "value": pix * steps / (ms_val * 1e-3) / 1e6, "e2e": pix * steps / (ms_e2e * 1e-3) / 1e6, "ms_val": ms_val / steps, "ms_e2e": ms_e2e / steps,
               "launches": launches, "rects": rects, "rects_e2e": rects_e2e, "wait_ms": wait_ms, "list_ms": list_ms,
               "d2h_bytes": (16 * 1024 * total_frames) if rects is None else sum(max(16 * 1024, 64 + 176 * len(r)) for r in rects)}
        batch.close()
        if gather_pool:
            gather_pool.shutdown()
        return res

    iw, ih, F = args.w, args.h, args.frames
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    main_run = run_workload(iw, ih, F * world, 1000, args.steps, args.warmup)            # config 3: seeds 1000+i, weak scaling
    clocks = sampler.finish() if sampler else None
    value, e2e = main_run["value"], main_run["e2e"]
    total_frames, frame_bytes, ws = main_run["total_frames"], main_run["frame_bytes"], main_run["ws"]
    lo, hi = main_run["lo"], main_run["hi"]
    dev_frames = main_run["dev_frames"]
    rects, rects_e2e = main_run["rects"], main_run["rects_e2e"]

    # per-kernel device time for the roofline of the top kernel: CUDA events around every launch on the launching stream
    # (rd_profile_*), in a separate pass over the same frames through ONE pipeline object, so that each kernel is timed
    # alone on the device (in the headline passes the streams of the pipeline objects overlap) and the event records do not
    # sit inside the headline timings
    def staged_profile(run_once, nframes):
        for _ in range(2):
            run_once()
        torch.cuda.synchronize()
        rd.api.profile_start(None, stages=True)
        run_once()
        torch.cuda.synchronize()
        prof, stage_ms = {}, {}
        for k, (c, ms) in rd.api.profile_stop().items():
            st, name = k.split("/", 1)
            pc, pm = prof.get(name, (0, 0.0))
            prof[name] = (pc + c, pm + ms)
            stage_ms[st] = stage_ms.get(st, 0.0) + ms
        return prof, stage_ms

    solo = rd.Batch(local_rank, iw, ih, nctx=1, frames_per_launch=args.fpl)
    nsolo = min(hi - lo, 4 * args.fpl)
    prof, stage_ms = staged_profile(lambda: solo.run(dev_frames.data_ptr(), frame_bytes, ws, nsolo, TAN_AOV, on_device=True, want_rects=False), nsolo)
    solo.close()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    # per-stage roofline: algorithmic bytes per pixel A 11, B 16, C 8 (+56 B per segment), D 8 (+20 B per vote pair);
    # the list terms are <1 % of a frame and are left out.  T = the device tail (executeCPUTask: FP64 compute on lists, no plane traffic)
    stage_bpp = {"A": 11, "B": 16, "C": 8, "D": 8, "T": 0}

    def stage_table(stage_ms, nframes, w, h):
        out = {}
        for st in sorted(stage_ms):
            us = stage_ms[st] * 1e3 / nframes
            gbs = stage_bpp.get(st, 0) * w * h / (us * 1e-6) / 1e9 if us > 0 and stage_bpp.get(st) else None
            out[st] = {"us_per_frame": round(us, 2), "algorithmic_bytes_per_px": stage_bpp.get(st), "achieved_gbs": round(gbs, 1) if gbs else None,
                       "frac_of_hbm_peak": round(gbs / peak, 4) if gbs else None}
        return out

    # ---- sub-records (BASELINE.json configs 4 and 5, the reference's streaming API, parity of the timed batch) ----
    del main_run["host_frames"], main_run["dev_frames"]
    dev_frames = None
    torch.cuda.empty_cache()
    extras = {}
    if not args.no_extras:
        c4 = run_workload(1920, 1080, args.config4_frames, 2000, 3, 1)                   # config 4: seeds 2000+i, contiguous shards, STRONG scaling
        extras["config4"] = {"workload": "vidrect 1920x1080 synthetic frame batch (seeds 2000+i), %d frames per step over all GPUs in contiguous shards, rect lists gathered to rank 0" % args.config4_frames,
                             "scaling": "strong", "value": c4["value"], "unit": "Mpix/s", "ms_per_step": c4["ms_val"],
                             "e2e": {"value": c4["e2e"], "unit": "Mpix/s", "ms_per_step": c4["ms_e2e"], "h2d_bytes_per_step": c4["total_frames"] * c4["frame_bytes"], "d2h_bytes_per_step": c4["d2h_bytes"]},
                             "pipeline_frac": (43.0 * c4["value"] * 1e6 / 1e9) / (peak * world), "rects_per_step": sum(len(r) for r in c4["rects"]) if c4["rects"] else None, "steps": 3, "warmup": 1}
        del c4
        torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_extras:
        import oracle_lib as ol
        import parity
        # config 5: one 3840x2160 frame through oclrect_executeOnce (pageable host frame in, rect list out), per-stage roofline
        w5, h5 = 3840, 2160
        f5 = ol.synth_frame(w5, h5, 5)
        d5 = rd.Device(local_rank)
        g5 = rd.OclRect(d5, w5, h5)
        lat = []
        for _ in range(6):
            t0 = time.perf_counter()
            r5 = g5.execute_once(f5, TAN_AOV)
            lat.append((time.perf_counter() - t0) * 1e3)
        prof5, stage5 = staged_profile(lambda: g5.run_device(f5, stop_step=0), 1)
        k5 = sum(v[1] for v in prof5.values())
        ms5 = float(np.median(lat[1:]))
        extras["config5"] = {"workload": "rect 3840x2160 single synthetic frame (seed 5), oclrect_executeOnce: host frame in, rect list out", "ms_per_frame": ms5,
                             "value": w5 * h5 / (ms5 * 1e-3) / 1e6, "unit": "Mpix/s", "kernel_ms_per_frame": k5, "rects": len(r5),
                             "pipeline_frac": 43.0 * w5 * h5 / (ms5 * 1e-3) / 1e9 / peak, "stages": stage_table(stage5, 1, w5, h5),
                             "top5_us": [[k, round(v[1] * 1e3, 1)] for k, v in sorted(prof5.items(), key=lambda kv: -kv[1][1])[:5]]}
        g5.close()
        d5.close()
        # the reference's own streaming API, unchanged caller (vidrect.cpp:159-172: enqueue N+1, poll N): one object, and one object per core
        from concurrent.futures import ThreadPoolExecutor
        nstream = 64
        sf = [np.ascontiguousarray(main_frames) for main_frames in synth_batch(iw, ih, 1000, nstream, pinned=False).numpy()]

        def stream(frames, did):
            d = rd.Device(did)
            g = rd.OclRect(d, iw, ih)
            g.execute_once(frames[0], TAN_AOV)                                           # warm-up (first-touch, tanAOV known to the object)
            t0 = time.perf_counter()
            n = 0
            g.enqueue_task(frames[0])
            for i in range(1, len(frames)):
                g.enqueue_task(frames[i])
                n += len(g.poll_task(TAN_AOV))
            n += len(g.poll_task(TAN_AOV))
            dt = time.perf_counter() - t0
            t1 = time.perf_counter()
            g.execute_once(frames[1], TAN_AOV)
            once = time.perf_counter() - t1
            g.close()
            d.close()
            return dt, n, once

        dt1, n1, once1 = stream(sf, local_rank)
        nobj = min(16, os.cpu_count() or 4)
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=nobj) as ex:
            list(ex.map(lambda k: stream(sf[: nstream // 2], local_rank), range(nobj)))
        dtn = time.perf_counter() - t0
        extras["api_stream"] = {"workload": "vidrect 1280x720: oclrect_enqueueTask / oclrect_pollTask, unchanged caller, pageable host frames, carry-over between frames",
                                "one_object": {"value": nstream * iw * ih / dt1 / 1e6, "unit": "Mpix/s", "ms_per_frame": dt1 / nstream * 1e3, "frames": nstream, "rects": n1},
                                "execute_once_ms": once1 * 1e3,
                                "objects_%d" % nobj: {"value": nobj * (nstream // 2 + 2) * iw * ih / dtn / 1e6, "unit": "Mpix/s", "frames_per_object": nstream // 2 + 2,
                                                      "note": "one oclrect_t + queue per host thread, object creation and warm-up inside the timing"}}
        # the first-pass replay of labelMergeMain (rd_set_merge_replay(1) / RD_MERGE_REPLAY=1: the reference's region map bit for bit,
        # include/rectdetect_b200.h): what it costs - one frame through executeOnce, a short batch - and a parity check in that mode
        rd.set_merge_replay(True)
        ol.oracle().ora_set_merge_replay(1)
        try:
            g = rd.OclRect(rd.Device(local_rank), iw, ih)
            g.execute_once(sf[0], TAN_AOV)
            ts = []
            for i in range(6):
                t1 = time.perf_counter()
                rr = g.execute_once(sf[1 + i % 2], TAN_AOV)
                ts.append(time.perf_counter() - t1)
            g.close()
            o = ol.OracleRect(iw, ih)
            o.execute_once(sf[0], TAN_AOV)
            for i in range(6):
                want = o.execute_once(sf[1 + i % 2], TAN_AOV)                              # (same carry-over history as the object above)
            o.close()
            nrep = min(hi - lo, 128)
            dev_rep = synth_batch(iw, ih, 1000, nrep, pinned=False).to("cuda")
            brep = rd.Batch(local_rank, iw, ih, nctx=args.nctx, frames_per_launch=args.fpl)
            for _ in range(2):
                brep.run(dev_rep.data_ptr(), frame_bytes, ws, nrep, TAN_AOV, on_device=True, want_rects=False)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            for _ in range(3):
                brep.run(dev_rep.data_ptr(), frame_bytes, ws, nrep, TAN_AOV, on_device=True, want_rects=False)
            torch.cuda.synchronize()
            dtb = (time.perf_counter() - t1) / 3
            brep.close()
            del dev_rep
            extras["merge_replay"] = {"what": "labelMergeMain with the reference's first pass replayed exactly (off by default)", "execute_once_ms": sorted(ts)[len(ts) // 2] * 1e3,
                                      "batch_value": nrep * iw * ih / dtb / 1e6, "unit": "Mpix/s", "batch_frames": nrep, "rects_identical_to_oracle_in_that_mode": want.tobytes() == rr.tobytes()}
        except Exception as e:                                                          # an optional sub-record must not cost the line
            extras["merge_replay"] = {"error": repr(e)}
        finally:
            rd.set_merge_replay(False)
            ol.oracle().ora_set_merge_replay(0)
        # parity of the timed batch: the first frames of the batch against fresh oracle objects (outside every timed region)
        bad = []
        k = min(args.parity_frames, len(rects) if rects else 0)
        for i in range(k):
            o = ol.OracleRect(iw, ih)
            want = o.execute_once(ol.synth_frame(iw, ih, 1000 + i), TAN_AOV)
            o.close()
            ok, why = parity.rects_close(want, rects[i])
            if not ok:
                bad.append((i, why))
        extras["parity_checked_frames"] = k
        extras["parity_ok"] = not bad
        if bad:
            extras["parity_failures"] = bad[:4]

    if rank == 0:
        top = max(prof.items(), key=lambda kv: kv[1][1])
        tname, (tcnt, tms) = top
        total_kernel_ms = sum(v[1] for v in prof.values())
        bpp = KERNEL_BYTES_PER_PX.get(tname)
        per_launch_ms = tms / tcnt
        fpl_eff = min(args.fpl, nsolo)
        alg_bytes = (bpp * iw * ih * fpl_eff) if bpp else None
        achieved = (alg_bytes / (per_launch_ms * 1e-3) / 1e9) if bpp else None
        traffic_pf, traffic_src = ncu_traffic(tname)
        roofline = {"bound": "hbm", "kernel": tname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                    "traffic": (traffic_pf * fpl_eff) if traffic_pf else None, "traffic_source": traffic_src, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "frames_per_launch": fpl_eff, "avg_launch_us": per_launch_ms * 1e3,
                    "share_of_kernel_time": tms / total_kernel_ms, "kernel_time_us_per_frame": total_kernel_ms * 1e3 / nsolo,
                    "top5_us_per_frame": [[k, round(v[1] * 1e3 / nsolo, 2)] for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:5]],
                    "pipeline_algorithmic_bytes_per_px": 43, "pipeline_frac": (43.0 * value * 1e6 / 1e9) / (peak * world),
                    "pipeline_frac_note": "whole-job algorithmic bytes per second over the HBM peak of the %d GPU(s) used" % world}
        roofline["stages"] = stage_table(stage_ms, nsolo, iw, ih)
        cpu = None
        if not args.no_cpu_baseline and world == 1:      # the CPU baseline is timed at N = 1 only
            if os.path.exists(REF_SO):
                v, cores, secs = cpu_reference_mpix(iw, ih, 1000, args.cpu_frames)
                cpu = {"value": v, "unit": "Mpix/s", "cores": cores, "kind": "reference", "sample": "%d frames of the same stream, %.1f s; %s" % (args.cpu_frames, secs, REF_KIND)}
            else:
                v, cores, secs = cpu_oracle_mpix(iw, ih, [1000 + i for i in range(args.cpu_frames)])
                cpu = {"value": v, "unit": "Mpix/s", "cores": cores, "kind": "port",
                       "sample": "%d frames of the same stream, %.1f s (CPU oracle = OpenMP restatement of the reference's OpenCL kernels + schedule + host tail)" % (args.cpu_frames, secs)}
            pv, pc, ps = cpu_oracle_mpix(iw, ih, [1000 + i for i in range(min(args.cpu_frames, 16))])
            cpu["port"] = {"value": pv, "unit": "Mpix/s", "cores": pc, "kind": "port", "sample": "CPU oracle (OpenMP restatement) on %d frames, %.1f s" % (min(args.cpu_frames, 16), ps)}
            cpu["build"] = "g++ -O2 for baseline x86-64 (no -march), -ffp-contract=off: the bit-exactness build of the reference, not a tuned one"
        nrect = sum(len(r) for r in rects) if rects else 0
        same = rects is not None and rects_e2e is not None and all(a.tobytes() == b.tobytes() for a, b in zip(rects, rects_e2e))
        line = {
            "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main_run["ms_val"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+i32 (f64 device tail)", "data": "synthetic",
            "config": {"workload": "vidrect %dx%d synthetic stream (seeds 1000+i), AOV 72, full imgutil->polyline->rect pipeline incl. executeCPUTask on the device, independent-frame batch "
                                   "(every frame as by a fresh oclrect_t: no carry-over between frames)" % (iw, ih),
                       "frames_per_gpu_per_step": F, "global_frames_per_step": total_frames, "pipelines_per_gpu": args.nctx, "frames_per_launch": args.fpl, "parallelism": "frames x%d" % world, "gather": "none (1 GPU)" if world == 1 else "rect lists of step k gathered to rank 0 over NCCL beside the frames of step k+1 (the last gather inside the timed region)",
                       "l2": "inputs larger than L2 (%d MB of frames + %d working sets of 31 planes per GPU)" % (F * frame_bytes // 2 ** 20, args.nctx * args.fpl)},
            "e2e": {"value": e2e, "unit": "Mpix/s", "h2d_bytes_per_step": total_frames * frame_bytes, "d2h_bytes_per_step": main_run["d2h_bytes"],
                    "ms_per_step": main_run["ms_e2e"]},
            "gpu_launches": main_run["launches"], "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "rects_per_step": nrect, "value_and_e2e_rects_identical": bool(same),
            "host": {"driver_threads": args.nctx, "wait_for_device_ms_last_step": main_run["wait_ms"], "host_list_copy_ms_last_step": main_run["list_ms"], "host_cores": os.cpu_count(),
                     "host_tail": "none: executeCPUTask runs on the device (rd_gtail.cu)"},
        }
        line.update(extras)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
