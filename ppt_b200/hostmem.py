"""Pinned host memory placed on the NUMA node a GPU hangs off.

The host-to-host path (tokenizer.HostPipeline, bench.py's `e2e`) is bound by PCIe and, with several GPUs per
host, by where the pinned pages live: a DMA into memory of the other socket crosses the inter-socket link and
shares it with every other GPU doing the same.  Linux places pages on the node of the thread that first touches
them (default "local" policy) and cudaHostAlloc touches them while pinning, so binding the calling thread to the
GPU's local CPUs for the duration of the allocation is enough -- no libnuma needed.  The binding is restored
afterwards.  Where the topology cannot be read (no sysfs entry, single node) this is a plain pinned allocation.
"""
import contextlib
import os

import torch


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_local_cpus(device):
    """CPUs of the NUMA node the GPU is attached to (sysfs local_cpulist of its PCI function), or None."""
    try:
        prop = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as f:
            cpus = _parse_cpulist(f.read())
        return cpus or None
    except Exception:
        return None


def gpu_numa_node(device):
    try:
        prop = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            return int(f.read().strip())
    except Exception:
        return None


@contextlib.contextmanager
def bound_to_gpu_node(device):
    """Runs the body with the calling thread restricted to the GPU's local CPUs (intersected with what this
    process is allowed to use); a no-op when that set is unknown or empty."""
    local = gpu_local_cpus(device)
    try:
        before = os.sched_getaffinity(0)
    except Exception:
        before = None
    target = (local & before) if (local and before) else None
    if not target or target == before:
        yield False
        return
    os.sched_setaffinity(0, target)
    try:
        yield True
    finally:
        os.sched_setaffinity(0, before)


def pinned_empty(shape, dtype, device):
    """torch.empty(shape, dtype).pin_memory() with its pages on `device`'s NUMA node."""
    with bound_to_gpu_node(device):
        t = torch.empty(shape, dtype=dtype).pin_memory()
        if t.numel():
            t.view(-1)[:: max(1, 4096 // t.element_size())] = 0  # touch every page while bound
    return t


def pinned_copy(tensor, device):
    """A pinned copy of a CPU tensor on `device`'s NUMA node."""
    out = pinned_empty(tuple(tensor.shape), tensor.dtype, device)
    out.copy_(tensor)
    return out
