"""pymetis stand-in (test infrastructure): contiguous block partition.

`Subnet.create_partitions` (CDSimulatorComponents.py:556-582) only needs a
membership list; partitions are never read by the step path.
"""


def part_graph(nparts, adjacency=None, **_):
    n = len(adjacency)
    size = max(1, -(-n // nparts))
    membership = [min(i // size, nparts - 1) for i in range(n)]
    return 0, membership
