"""NUMA binding helper (pylabolt_b200/affinity.py) against a fake sysfs tree."""
import os

from pylabolt_b200 import affinity


def _fake_sysfs(root, bus_id, node, cpulists):
    dev = root / "bus" / "pci" / "devices" / bus_id
    dev.mkdir(parents=True)
    (dev / "numa_node").write_text(f"{node}\n")
    for n, cpus in cpulists.items():
        d = root / "devices" / "system" / "node" / f"node{n}"
        d.mkdir(parents=True)
        (d / "cpulist").write_text(cpus + "\n")


def test_cpulist_parsing():
    assert affinity._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert affinity._parse_cpulist("") == set()


def test_numa_node_lookup(tmp_path):
    _fake_sysfs(tmp_path, "0000:1b:00.0", 1, {0: "0-1", 1: "2-3"})
    assert affinity.gpu_numa_node("0000:1b:00.0", str(tmp_path)) == 1
    assert affinity.node_cpus(1, str(tmp_path)) == {2, 3}
    assert affinity.gpu_numa_node("0000:ff:00.0", str(tmp_path)) is None


def test_unknown_node_changes_nothing(tmp_path):
    _fake_sysfs(tmp_path, "0000:1b:00.0", -1, {0: "0-63"})
    before = os.sched_getaffinity(0)
    assert affinity.bind_to_gpu("0000:1b:00.0", str(tmp_path)) is None
    assert os.sched_getaffinity(0) == before


def test_bind_restricts_to_the_node_and_can_be_disabled(tmp_path, monkeypatch):
    allowed = sorted(os.sched_getaffinity(0))
    if len(allowed) < 2:
        return
    half = allowed[:len(allowed) // 2]
    rest = allowed[len(allowed) // 2:]
    _fake_sysfs(tmp_path, "0000:1b:00.0", 0,
                {0: ",".join(map(str, half)), 1: ",".join(map(str, rest))})
    try:
        monkeypatch.setenv("PLB_NUMA_BIND", "0")
        assert affinity.bind_to_gpu("0000:1b:00.0", str(tmp_path)) is None
        assert sorted(os.sched_getaffinity(0)) == allowed
        monkeypatch.setenv("PLB_NUMA_BIND", "1")
        assert affinity.bind_to_gpu("0000:1b:00.0", str(tmp_path)) == 0
        assert sorted(os.sched_getaffinity(0)) == half
    finally:
        os.sched_setaffinity(0, allowed)
