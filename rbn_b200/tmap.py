"""Triangular chart container with the storage order of `triangularmap.TMap` as the reference uses it.

The reference imports TMap from the third-party package `triangularmap` (un-pinned, /root/reference/requirements.txt:11)
and only relies on: TMap(arr), TMap.size_from_n, chart[start, end] get/set, .arr and .pretty()
(/root/reference/rbnet/pcfg.py:32,41,215,272,309-310; /root/reference/rbnet/sequential.py:36).  The row order
is pinned by /root/reference/tests/test_pcfg.py:14-19: root first, then spans of decreasing width, left to right.
"""
import math


class TMap:
    """A flat array viewed as the upper triangle of spans (start, end) over a sequence of length n."""

    def __init__(self, arr, *args, **kwargs):
        self.arr = arr
        self.n = self.n_from_size(len(arr))

    @staticmethod
    def size_from_n(n):
        return (n * n + n) // 2

    @staticmethod
    def n_from_size(size):
        n = (math.isqrt(1 + 8 * size) - 1) // 2
        if TMap.size_from_n(n) != size:
            raise ValueError(f"a triangular chart cannot have {size} entries")
        return n

    def _row(self, key):
        start, end = key
        width = end - start
        if start < 0 or width < 1 or end > self.n:
            raise IndexError(f"span {key} outside a chart over {self.n} positions")
        above = self.n - width            # number of complete levels above this span's level
        return above * (above + 1) // 2 + start

    def __getitem__(self, key):
        return self.arr[self._row(key)]

    def __setitem__(self, key, value):
        self.arr[self._row(key)] = value

    def level(self, width):
        """All spans of the given width, left to right."""
        first = self._row((0, width))
        return self.arr[first:first + self.n - width + 1]

    def pretty(self, cell_width=None):
        lines = []
        for width in range(self.n, 0, -1):
            cells = [str(c) for c in self.level(width)]
            pad = cell_width or max(len(c) for c in cells)
            lines.append(" " * ((width - 1) * (pad + 1) // 2) + " ".join(c.center(pad) for c in cells))
        return "\n".join(lines)
