"""Topology checks shared by the CPU and GPU surface-extraction tests."""
import collections


def _edge_stats(tris):
    und, dirc = collections.Counter(), collections.Counter()
    for a, b, c in tris:
        for x, y in ((a, b), (b, c), (c, a)):
            und[(min(x, y), max(x, y))] += 1
            dirc[(x, y)] += 1
    return und, dirc


def assert_closed_oriented(tris):
    und, dirc = _edge_stats(tris.tolist())
    assert all(v == 2 for v in und.values())                              # watertight: every edge shared by 2 triangles
    assert all(v == 1 and dirc.get((b, a), 0) == 1 for (a, b), v in dirc.items())   # consistently oriented
