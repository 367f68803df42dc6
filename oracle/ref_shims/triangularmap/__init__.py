"""Stand-in for `triangularmap.TMap` (PyPI, un-pinned in /root/reference/requirements.txt:11, source absent).

TEST INFRASTRUCTURE ONLY.  TMap does no arithmetic: storage + index mapping.  The linear layout is pinned
by the reference's own tests:
  * /root/reference/tests/test_pcfg.py:14-19  (n=4: rows (0,4) | (0,3),(1,4) | (0,2),(1,3),(2,4) | width-1 cells)
  * /root/reference/tests/test_util.py:18-22  (n=3: [0,2] -> row 1)
=> size = n(n+1)/2, depth = n-(end-start), lin = depth(depth+1)/2 + start.
Call sites in the reference: TMap(arr), TMap.size_from_n(n) (pcfg.py:41,215), chart[start,end] get
(pcfg.py:309-310,272) / set (sequential.py:36), .arr (pcfg.py:32), .pretty() (examples only).
"""
import math


class TMap:
    def __init__(self, arr, *args, **kwargs):
        self.arr = arr
        self.n = self.n_from_size(len(arr))

    @staticmethod
    def size_from_n(n):
        return n * (n + 1) // 2

    @staticmethod
    def n_from_size(size):
        n = (math.isqrt(8 * size + 1) - 1) // 2
        if n * (n + 1) // 2 != size:
            raise ValueError(f"{size} is not a triangular number")
        return n

    def linear_index(self, start, end):
        if not (0 <= start < end <= self.n):
            raise IndexError(f"invalid span ({start}, {end}) for n={self.n}")
        depth = self.n - (end - start)
        return depth * (depth + 1) // 2 + start

    def __getitem__(self, item):
        start, end = item
        return self.arr[self.linear_index(start, end)]

    def __setitem__(self, key, value):
        start, end = key
        self.arr[self.linear_index(start, end)] = value

    def pretty(self, *args, **kwargs):
        rows = []
        for depth in range(self.n):
            rows.append("  ".join(str(self.arr[depth * (depth + 1) // 2 + s]) for s in range(depth + 1)))
        return "\n".join(rows)
