"""Host placement helper (simple_mmd_renderer_b200/hostmem.py): `nvidia-smi topo -m` parsing and the decision not to bind
on a single-node host (round 1's bind_to_gpu_numa silently did nothing when sysfs reported numa_node = -1)."""
from simple_mmd_renderer_b200 import hostmem

TOPO_2S = ("\t\x1b[4mGPU0\tGPU1\tGPU2\tNIC0\tCPU Affinity\tNUMA Affinity\tGPU NUMA ID\x1b[0m\n"
           "GPU0\t X \tNV18\tNV18\tSYS\t0-15,64-79\t0\t\tN/A\n"
           "GPU1\tNV18\t X \tNV18\tSYS\t0-15,64-79\t0\t\tN/A\n"
           "GPU2\tNV18\tNV18\t X \tPIX\t32-47,96-111\t2\t\tN/A\n"
           "NIC0\tSYS\tSYS\tPIX\t X \n\nLegend:\n  X = Self\n")
TOPO_1N = "\tGPU0\tCPU Affinity\tNUMA Affinity\tGPU NUMA ID\nGPU0\t X \t0-15\t0\t\tN/A\n"


def test_topo_matrix_is_parsed_per_gpu():
    rows = hostmem.parse_topo_matrix(TOPO_2S)
    assert set(rows) == {0, 1, 2}
    assert rows[0][1] == 0 and rows[2][1] == 2
    assert rows[0][0] == set(range(0, 16)) | set(range(64, 80))
    assert rows[2][0] == set(range(32, 48)) | set(range(96, 112))
    one = hostmem.parse_topo_matrix(TOPO_1N)
    assert one == {0: (set(range(16)), 0)}
    assert hostmem.parse_topo_matrix("no such table") == {}


def test_cpulist_forms():
    assert hostmem._parse_cpulist("0-3, 8,10-11") == {0, 1, 2, 3, 8, 10, 11}
    assert hostmem._parse_cpulist("") == set()


def test_describe_topology_runs_without_a_gpu():
    d = hostmem.describe_topology()
    assert d["numa_nodes"] >= 1
    # no GPU here: nothing to bind to, and the call says so instead of pretending
    assert hostmem.bind_to_gpu_numa(0) is None or isinstance(hostmem.bind_to_gpu_numa(0), dict)
