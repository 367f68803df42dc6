"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log: per kernel
launch count, summed duration and DRAM bytes.   python tools/summarize_traffic.py <csv> [passes_in_file] [passes_per_step]"""
import collections
import csv
import json
import re
import sys


def load(path):
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ik, iv, im, iu, iid = (hdr.index(x) for x in ('Kernel Name', 'Metric Value', 'Metric Name', 'Metric Unit', 'ID'))
    rows = collections.OrderedDict()
    mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, 'ns': 1, 'us': 1e3, 'ms': 1e6, 'second': 1e9}
    for x in r:
        d = rows.setdefault(int(x[iid]), {'k': x[ik]})
        d[x[im]] = float(x[iv].replace(',', '')) * mult.get(x[iu], 1)
    return rows


def main():
    path = sys.argv[1]
    passes = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
    per_step = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
    rows = load(path)
    names = collections.OrderedDict()
    for d in rows.values():
        k = re.sub(r'^void ', '', re.sub(r'\(.*', '', d['k']))[:56]
        a = names.setdefault(k, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get('gpu__time_duration.sum', 0)
        a[2] += d.get('dram__bytes_read.sum', 0)
        a[3] += d.get('dram__bytes_write.sum', 0)
    tot = [0.0, 0.0, 0.0]
    out = {}
    print(f"{'kernel':58s} {'n':>6s} {'ms':>9s} {'read GB':>9s} {'write GB':>9s}")
    for k, a in sorted(names.items(), key=lambda kv: -kv[1][1]):
        if a[1] / 1e6 > 0.05:
            print(f"{k:58s} {a[0]:6d} {a[1]/1e6:9.2f} {a[2]/1e9:9.2f} {a[3]/1e9:9.2f}")
        out[k] = dict(launches=a[0], ms=a[1] / 1e6, dram_read_GB=a[2] / 1e9, dram_write_GB=a[3] / 1e9)
        for i in range(3):
            tot[i] += a[1 + i]
    scale = per_step / passes
    print(f"total: {tot[0]/1e6:.1f} ms, read {tot[1]/1e9:.1f} GB, write {tot[2]/1e9:.1f} GB in {passes:g} passes")
    print(f"per step ({per_step:g} passes): DRAM {(tot[1]+tot[2])/1e12*scale:.3f} TB (read {tot[1]/1e12*scale:.3f}, write {tot[2]/1e12*scale:.3f})")
    return dict(kernels=out, passes_in_file=passes, passes_per_step=per_step,
                dram_TB_per_step=(tot[1] + tot[2]) / 1e12 * scale, dram_read_TB_per_step=tot[1] / 1e12 * scale,
                dram_write_TB_per_step=tot[2] / 1e12 * scale)


if __name__ == "__main__":
    res = main()
    if len(sys.argv) > 4:
        json.dump(res, open(sys.argv[4], "w"), indent=1)
