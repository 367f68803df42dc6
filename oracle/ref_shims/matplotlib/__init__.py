"""Stand-in for `matplotlib` (only /root/reference/rbnet/util.py:363-377 plotting helpers use it)."""
