"""The compare-exchange schedules in csrc/cav_sortnet.cuh sort every input (0-1 principle, exhaustive)."""
import itertools
import os
import re

HDR = os.path.join(os.path.dirname(__file__), "..", "gym_collision_avoidance_b200", "csrc", "cav_sortnet.cuh")


def _networks():
    src = open(HDR).read()
    nets = {}
    for m in re.finditer(r"void sortnet_(\d+)\(F &&cex\) \{[^\n]*\n(.*?)\n\}", src, re.S):
        nets[int(m.group(1))] = [(int(a), int(b)) for a, b in re.findall(r"cex\((\d+), (\d+)\);", m.group(2))]
    return nets


def test_networks_sort_all_binary_inputs():
    nets = _networks()
    assert set(nets) >= {2, 3, 4, 5, 6, 7, 8, 9, 11}
    for n, net in nets.items():
        assert all(0 <= i < j < n for i, j in net)
        for bits in itertools.product((0, 1), repeat=n):
            a = list(bits)
            for i, j in net:
                if a[i] > a[j]:
                    a[i], a[j] = a[j], a[i]
            assert a == sorted(a), (n, bits)
