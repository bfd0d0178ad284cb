"""Host-side placement for the device->host leg of a bake: bind a rank to the CPUs and memory node next to its GPU
before it allocates page-locked buffers (8 ranks that land their copies in one remote node share one inter-socket link).

The node is looked up in sysfs first (`/sys/bus/pci/devices/<bdf>/numa_node`); VMs often report -1 there, so the
fallback is the "CPU Affinity" / "NUMA Affinity" columns of `nvidia-smi topo -m`.  Nothing here touches the device path.
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess


def _parse_cpulist(spec: str) -> set[int]:
    cpus: set[int] = set()
    for part in spec.replace(" ", "").split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def _pci_bdf(device_index: int) -> str | None:
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        return f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:
        return None


def _sysfs_node(device_index: int):
    bdf = _pci_bdf(device_index)
    if not bdf:
        return None
    try:
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            return node, _parse_cpulist(f.read().strip())
    except Exception:
        return None


def parse_topo_matrix(text: str) -> dict[int, tuple[set[int], int | None]]:
    """`nvidia-smi topo -m` -> {gpu index: (cpu affinity set, numa node or None)}."""
    out: dict[int, tuple[set[int], int | None]] = {}
    header = None
    for line in text.splitlines():
        line = re.sub(r"\x1b\[[0-9;]*m", "", line)
        cols = [c.strip() for c in line.split("\t") if c.strip() != ""]
        if not cols:
            continue
        if header is None and any(c.startswith("CPU Affinity") for c in cols):
            header = cols
            continue
        m = re.match(r"GPU(\d+)$", cols[0])
        if header is None or not m:
            continue
        # data rows carry one more leading column (the row label) than the header
        try:
            ci = header.index("CPU Affinity") + 1
        except ValueError:
            continue
        cpus = _parse_cpulist(cols[ci]) if ci < len(cols) and re.match(r"^[0-9,\- ]+$", cols[ci]) else set()
        node = None
        if "NUMA Affinity" in header:
            ni = header.index("NUMA Affinity") + 1
            if ni < len(cols) and re.match(r"^\d+$", cols[ni]):
                node = int(cols[ni])
        out[int(m.group(1))] = (cpus, node)
    return out


def _topo_node(device_index: int):
    try:
        r = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20)
        row = parse_topo_matrix(r.stdout).get(device_index)
        if row and row[0]:
            return row[1], row[0]
    except Exception:
        pass
    return None


def _set_preferred_node(node: int) -> bool:
    """set_mempolicy(MPOL_PREFERRED, {node}): pages this thread touches (and the driver pins for it) come from `node`."""
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        MPOL_PREFERRED, SYS_set_mempolicy = 1, 238      # x86-64
        mask = ctypes.c_ulong(1 << node)
        return libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(64)) == 0
    except Exception:
        return False


def n_numa_nodes() -> int:
    try:
        return len([d for d in os.listdir("/sys/devices/system/node") if re.match(r"node\d+$", d)])
    except Exception:
        return 1


def bind_to_gpu_numa(device_index: int) -> dict | None:
    """Bind the calling process next to GPU `device_index`; returns what was done (None: nothing to bind to)."""
    found = _sysfs_node(device_index)
    source = "sysfs"
    if found is None:
        found = _topo_node(device_index)
        source = "nvidia-smi topo"
    if found is None:
        return None
    node, cpus = found
    try:
        allowed = os.sched_getaffinity(0)
    except Exception:
        return None
    cpus = set(cpus) & allowed
    if not cpus or cpus == allowed and (node is None or n_numa_nodes() <= 1):
        return {"source": source, "node": node, "cpus": len(cpus), "bound": False,
                "why": "single NUMA node / affinity already covers every allowed CPU"}
    os.sched_setaffinity(0, cpus)
    mem = _set_preferred_node(node) if node is not None and n_numa_nodes() > 1 else False
    return {"source": source, "node": node, "cpus": len(cpus), "bound": True, "mempolicy_preferred": mem}


def describe_topology() -> dict:
    d = {"numa_nodes": n_numa_nodes()}
    try:
        d["allowed_cpus"] = len(os.sched_getaffinity(0))
    except Exception:
        pass
    try:
        r = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20)
        rows = parse_topo_matrix(r.stdout)
        d["gpu_affinity"] = {str(k): {"cpus": len(v[0]), "numa": v[1]} for k, v in sorted(rows.items())}
    except Exception:
        pass
    return d
