"""NUMA placement of a rank: one process per GPU should run, and allocate its
page-locked transfer buffers, on the CPU socket the GPU hangs off.  With eight
ranks started by torchrun on an unbound box the host ends of all eight PCIe
streams otherwise land on one socket and the field upload / download of the
ranks on the other socket crosses the inter-socket link.

The reference pins nothing (mpirun's binding is the user's business); here the
b200 back end binds itself because it knows its GPU.  ``PLB_NUMA_BIND=0``
switches it off.
"""
import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(pci_bus_id, sysfs="/sys"):
    """NUMA node of a PCI device, or None when the platform does not say
    (virtual machines often report -1)."""
    try:
        with open(os.path.join(sysfs, "bus", "pci", "devices", pci_bus_id,
                               "numa_node")) as f:
            node = int(f.read().strip())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def node_cpus(node, sysfs="/sys"):
    try:
        with open(os.path.join(sysfs, "devices", "system", "node",
                               f"node{node}", "cpulist")) as f:
            return _parse_cpulist(f.read())
    except (OSError, ValueError):
        return set()


def bind_to_gpu(pci_bus_id, sysfs="/sys"):
    """Restricts this process to the CPUs of the GPU's NUMA node (memory it
    allocates afterwards is then node-local by first touch).  Returns the node
    or None if nothing was changed."""
    if os.environ.get("PLB_NUMA_BIND", "1") in ("0", ""):
        return None
    node = gpu_numa_node(pci_bus_id, sysfs)
    if node is None:
        return None
    cpus = node_cpus(node, sysfs)
    try:
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
    except (AttributeError, OSError):
        return None
    return node
