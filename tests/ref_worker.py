"""One worker of the timed CPU reference arm (bench.py --impl reference / cpu_baseline): THE REFERENCE ITSELF
(oracle/_ref/librd_ref.so, see tests/ref_lib.py) processing a contiguous share of a synthetic stream through its own
oclrect_enqueueTask / oclrect_pollTask pipeline (the vidrect.cpp:159-205 steady state), work-items in raster order on one
host thread.  bench.py starts one worker per host core (frames are independent, so the stream is split among them), waits
until all have warmed up, releases them together and takes the wall time until the last one reports.

usage: python tests/ref_worker.py IW IH FIRST_SEED COUNT   (prints "ready", waits for a line on stdin, prints "done <rects>")
Test / measurement infrastructure only."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_lib as ol  # noqa: E402
import ref_lib as rl  # noqa: E402


def main():
    iw, ih, first, count = (int(a) for a in sys.argv[1:5])
    tan_aov = math.tan(math.radians(36.0))
    rl.set_threads(1)
    frames = [ol.synth_frame(iw, ih, first + i) for i in range(count)]
    r = rl.RefRect(iw, ih)
    r.execute_once(frames[0], tan_aov, iw * 3)                        # warm-up: page faults, lazy binding
    print("ready", flush=True)
    sys.stdin.readline()
    L, nrect = r.L, 0
    L.oclrect_enqueueTask(r.h, frames[0].ctypes.data, iw * 3)
    for i in range(1, count):
        L.oclrect_enqueueTask(r.h, frames[i].ctypes.data, iw * 3)
        nrect += len(rl._rects(L.oclrect_pollTask(r.h, tan_aov)))
    nrect += len(rl._rects(L.oclrect_pollTask(r.h, tan_aov)))
    print("done %d" % nrect, flush=True)


if __name__ == "__main__":
    main()
