"""File -> GPU sharding.  Files shard with no exchange step: a file (one get_ranges
request) is the unit, nothing is reduced or gathered, so there is no collective on the data
path.  `gpu(file)` = the GPU whose NUMA node matches the block device holding the file; with
one disk / one NUMA node (every box seen so far) it falls back to `file_idx mod n_gpus`.
"""
import os


def _read_int(path, default=-1):
    try:
        with open(path) as f:
            return int(f.read().strip())
    except (OSError, ValueError):
        return default


def numa_node_of_path(path):
    """NUMA node of the block device backing `path`, or -1 (tmpfs, virtio without affinity, ...)."""
    try:
        st = os.stat(path)
    except OSError:
        return -1
    major, minor = os.major(st.st_dev), os.minor(st.st_dev)
    base = f"/sys/dev/block/{major}:{minor}"
    for rel in ("device/numa_node", "../device/numa_node"):
        n = _read_int(os.path.join(base, rel))
        if n >= 0:
            return n
    return -1


def numa_node_of_gpu(pci_bus_id):
    return _read_int(f"/sys/bus/pci/devices/{pci_bus_id.lower()}/numa_node")


def gpu_for_file(file_idx, n_gpus, file_node=-1, gpu_nodes=None):
    """Affinity map: among the GPUs on the file's NUMA node (if known), else all, pick by index."""
    cands = list(range(n_gpus))
    if file_node >= 0 and gpu_nodes:
        local = [g for g in range(n_gpus) if gpu_nodes[g] == file_node]
        if local:
            cands = local
    return cands[file_idx % len(cands)]


def files_for_rank(n_files, rank, world_size, file_nodes=None, gpu_nodes=None):
    """Indices of the files rank `rank` (one process per GPU) reads: a partition of range(n_files)."""
    return [i for i in range(n_files)
            if gpu_for_file(i, world_size, file_nodes[i] if file_nodes else -1, gpu_nodes) == rank]
