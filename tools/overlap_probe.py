"""Does running the geometry of batch i+1 (spatial index, FPS, kNN) on a second stream against the Encoder of batch i
buy anything?  ("overlap".)  Every kernel of the step is sized to own an SM's shared memory
(FPS 197 KB, kNN 190 KB, stage 2 226 KB per CTA), so two of them never share an SM: what can overlap is only the
SMs one kernel leaves idle (FPS: 128 clouds on 148 SMs) and the ragged ends of the persistent kernels.
Prints serial and two-stream ms/step at BASELINE configs[1] (128 clouds x 8192 points)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from ppt_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
tok = bench.make_tokenizer("fp16").to(dev)
B, G, K = bench.BATCH_PER_GPU, bench.N_GROUP, bench.GROUP_SIZE
resident = [h.to(dev) for h in bench.make_host_batches(0, bench.ROTATE, device=dev)]
zeros = torch.zeros(B, dtype=torch.int64, device=dev)
blob, mode = tok.encoder._blob(dev)
STEPS = 40


def geometry(xyz):
    index = ops.spatial_index(xyz)
    _, center = ops.fps(xyz, G, zeros, return_centers=True, index=index)
    return ops.knn_group(xyz, center, K, index=index)


def serial():
    for i in range(STEPS):
        ops.encoder_forward(geometry(resident[i % len(resident)]), blob, mode=mode)


def two_streams(prio):
    main = torch.cuda.current_stream()
    lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
    sg = torch.cuda.Stream(priority=hi if prio else lo)
    sg.wait_stream(main)
    keep, ready = [], []
    with torch.cuda.stream(sg):
        nb = geometry(resident[0])
        ev = torch.cuda.Event(); ev.record(sg)
    ready.append((nb, ev))
    for i in range(STEPS):
        if i + 1 < STEPS:
            with torch.cuda.stream(sg):
                nb2 = geometry(resident[(i + 1) % len(resident)])
                ev2 = torch.cuda.Event(); ev2.record(sg)
            ready.append((nb2, ev2))
        nb, ev = ready[i]
        main.wait_event(ev)
        keep.append((nb, ops.encoder_forward(nb, blob, mode=mode)))
    main.wait_stream(sg)
    return keep


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    torch.cuda.synchronize()
    del out
    return a.elapsed_time(b) / STEPS


print("serial               %.4f ms/step" % timed(serial))
print("two streams          %.4f ms/step" % timed(lambda: two_streams(False)))
print("two streams, geometry at high priority  %.4f ms/step" % timed(lambda: two_streams(True)))
