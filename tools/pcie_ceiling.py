#!/usr/bin/env python
"""Raw pinned-memory copy ceiling of the box: what `e2e` can at best reach.

    python tools/pcie_ceiling.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_ceiling.py

Every rank times cudaMemcpyAsync of the e2e path's own transfer sizes between pinned host memory and its GPU
(D2H 101.4 MB = fp32 tokens + centres of a 128-cloud step, 50.7 MB = fp16 tokens; H2D 12.6 MB = the clouds),
alone and with both directions at once, all ranks concurrently (barrier before, max over ranks after), with the
pinned pages (a) wherever the allocating thread happened to run and (b) on the GPU's own NUMA node
(ppt_b200.hostmem).  Rank 0 prints one JSON line: per-rank and aggregate GB/s."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppt_b200 import hostmem  # noqa: E402

D2H32, D2H16, H2D = 128 * 512 * (384 + 3) * 4, 128 * 512 * 384 * 2 + 128 * 512 * 12, 128 * 8192 * 12
REPS = 20


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    gbuf_out = torch.empty(D2H32, dtype=torch.uint8, device=dev)
    gbuf_in = torch.empty(H2D, dtype=torch.uint8, device=dev)
    s_out, s_in = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn):
        fn()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s_out.wait_event(e0)
        s_in.wait_event(e0)
        for _ in range(REPS):
            fn()
        torch.cuda.current_stream().wait_stream(s_out)
        torch.cuda.current_stream().wait_stream(s_in)
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1) / REPS
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    res = {}
    for placement in ("default", "numa_local"):
        if placement == "default":
            h_out, h_in = torch.empty(D2H32, dtype=torch.uint8).pin_memory(), torch.empty(H2D, dtype=torch.uint8).pin_memory()
        else:
            h_out, h_in = hostmem.pinned_empty((D2H32,), torch.uint8, dev), hostmem.pinned_empty((H2D,), torch.uint8, dev)

        def d2h(n):
            with torch.cuda.stream(s_out):
                h_out[:n].copy_(gbuf_out[:n], non_blocking=True)

        def h2d():
            with torch.cuda.stream(s_in):
                gbuf_in.copy_(h_in, non_blocking=True)

        r = {}
        for name, fn, nbytes in (("d2h_fp32_tokens", lambda: d2h(D2H32), D2H32), ("d2h_fp16_tokens", lambda: d2h(D2H16), D2H16),
                                 ("h2d_clouds", h2d, H2D),
                                 ("both_fp32", lambda: (d2h(D2H32), h2d()), D2H32 + H2D),
                                 ("both_fp16", lambda: (d2h(D2H16), h2d()), D2H16 + H2D)):
            ms = timed(fn)
            r[name] = {"ms": ms, "gbs_per_gpu": nbytes / ms / 1e6, "gbs_aggregate": world * nbytes / ms / 1e6,
                       "steps_per_s_ceiling": 1e3 / ms, "clouds_per_s_ceiling": world * 128 * 1e3 / ms}
        res[placement] = r
    info = {"rank_numa_node": hostmem.gpu_numa_node(dev), "local_cpus": len(hostmem.gpu_local_cpus(dev) or ()),
            "affinity": len(os.sched_getaffinity(0))}
    if world > 1:
        infos = [None] * world
        dist.all_gather_object(infos, info)
        dist.destroy_process_group()
    else:
        infos = [info]
    if rank == 0:
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
        lscpu = subprocess.run("lscpu | grep -i -E 'numa|model name|socket|^CPU\\(s\\)'", shell=True, capture_output=True,
                               text=True).stdout
        print(json.dumps({"n_gpus": world, "bytes": {"d2h_fp32": D2H32, "d2h_fp16": D2H16, "h2d": H2D}, "reps": REPS,
                          "ranks": infos, "copies": res, "topo": topo, "lscpu": lscpu}))


if __name__ == "__main__":
    main()
