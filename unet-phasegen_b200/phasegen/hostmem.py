"""Host-side placement for the end-to-end path (pinned host buffers <-> HBM over PCIe).

A process that feeds one GPU should run, and allocate its pinned staging buffers, on the NUMA node that GPU's PCIe
root hangs off: with one process per GPU on a two-socket host, buffers that land on the other socket make every
host<->device copy cross the inter-socket link, which is what `run_host` is bound by once eight processes copy at once.
Best effort and silent: containers and VMs often hide the topology (numa_node = -1), then nothing is changed.
"""
import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_locality(device_index):
    """(numa_node, set of local CPUs) of a CUDA device from sysfs, or (None, None) when the platform does not say."""
    import torch
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(f"{base}/numa_node").read().strip())
        cpus = _parse_cpulist(open(f"{base}/local_cpulist").read())
    except (OSError, ValueError, AttributeError):
        return None, None
    if node < 0 or not cpus:
        return None, None
    return node, cpus


def bind_to_gpu_node(device_index):
    """Restrict this process to the CPUs local to `device_index` (so that pinned buffers allocated afterwards are
    first-touched on that node).  Returns a short description for logs: {"numa_node": n, "cpus": k} or {"numa_node": None}."""
    node, cpus = gpu_locality(device_index)
    if node is None or not hasattr(os, "sched_setaffinity"):
        return {"numa_node": None}
    try:
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return {"numa_node": node, "cpus": 0, "bound": False}
        os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed), "bound": True}
    except OSError:
        return {"numa_node": node, "bound": False}
