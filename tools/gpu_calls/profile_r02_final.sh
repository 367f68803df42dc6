#!/bin/bash
# Round 2, final profiling pass at HEAD (one GPU).  Every profiled command first runs once WITHOUT ncu and must exit 0.
#   launch list of one 128-sequence C3 pass pair (duration + DRAM bytes per launch), ncu --set full of the two largest kernel
#   classes (child gradients, split-sum) and of the float32-chart Gaussian kernels.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
set -u
ARGS="--batch 128 --micro 128 --steps 1 --warmup 1 --skip-e2e --skip-cpu --no-secondary"
python bench.py $ARGS > gpurun_out/r02f_plain_c3.log 2>&1 || { echo "plain c3 run failed"; tail -5 gpurun_out/r02f_plain_c3.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02f_launches_raw.csv python bench.py $ARGS > gpurun_out/r02f_ncu_ll.log 2>&1
echo "launch list rc=$?"
python - <<'P'
import collections, csv
lines = [l for l in open('gpurun_out/r02f_launches_raw.csv') if l.startswith('"')]
r = csv.reader(lines); hdr = next(r)
ik, iv, im, iu, iid = (hdr.index(x) for x in ('Kernel Name', 'Metric Value', 'Metric Name', 'Metric Unit', 'ID'))
mult = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3, 'ns': 1e-3, 'us': 1, 'ms': 1e3, 'second': 1e6}
rows = collections.OrderedDict()
for x in r:
    d = rows.setdefault(int(x[iid]), {'k': x[ik]})
    d[x[im]] = float(x[iv].replace(',', '')) * mult.get(x[iu], 1)
with open('gpurun_out/r02f_launches_c3_b128.csv', 'w') as f:
    f.write('# every kernel launch of bench.py --batch 128 --micro 128 --steps 1 --warmup 1 at the round\'s last kernel commit (two inside+outside passes over 128 sequences, C3 shapes),\n')
    f.write('# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none (cold cache, serialised: compare shares)\n')
    f.write('id,kernel,duration_us,dram_read_MB,dram_write_MB\n')
    for i, d in rows.items():
        f.write(f"{i},{d['k'][:70]},{d.get('gpu__time_duration.sum', 0):.2f},{d.get('dram__bytes_read.sum', 0):.2f},{d.get('dram__bytes_write.sum', 0):.2f}\n")
cls = collections.OrderedDict()
for d in rows.values():
    k = d['k'].split('(')[0][:60]
    c = cls.setdefault(k, [0, 0.0, 0.0, 0.0]); c[0] += 1; c[1] += d.get('gpu__time_duration.sum', 0); c[2] += d.get('dram__bytes_read.sum', 0); c[3] += d.get('dram__bytes_write.sum', 0)
tot = sum(c[2] + c[3] for c in cls.values())
print(f"DRAM bytes of the two passes: {tot / 1e6:.3f} TB -> per 512-sequence step (x2): {2 * tot / 1e6:.3f} TB")
for k, c in sorted(cls.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{c[0]:5d} {c[1] / 1e3:9.2f} ms  rd {c[2] / 1e3:8.1f} GB  wr {c[3] / 1e3:8.1f} GB  {k}")
P
rm -f gpurun_out/r02f_launches_raw.csv
full() {   # full <name> <regex> <skip> <bench args...>
    local name=$1 pat=$2 skip=$3; shift 3
    ncu --set full --clock-control none --import-source on -k "regex:$pat" -s $skip -c 1 -f -o gpurun_out/r02f_full_$name \
        python bench.py "$@" > gpurun_out/r02f_ncu_$name.log 2>&1
    echo "$name rc=$? $(tail -1 gpurun_out/r02f_ncu_$name.log | cut -c1-120)"
    ncu -i gpurun_out/r02f_full_$name.ncu-rep --page raw --csv > gpurun_out/r02f_full_$name.csv 2>/dev/null
    rm -f gpurun_out/r02f_full_$name.ncu-rep
    python tools/extract_ncu.py gpurun_out/r02f_full_$name.csv gpurun_out/r02f_ncu_full_$name.csv "HEAD of round 2, $*" && rm -f gpurun_out/r02f_full_$name.csv
}
full k5 childgrad 208 $ARGS
full k1 splitsum 172 $ARGS
ARGS4="--workload c4f32 --steps 1 --warmup 1 --skip-cpu --no-secondary"
python bench.py $ARGS4 > gpurun_out/r02f_plain_c4f32.log 2>&1 || { echo "plain c4f32 run failed"; exit 1; }
full c4f32_fwd k_gauss_diag8 90 $ARGS4
full c4f32_bwd k_gauss_bwd_diag8 90 $ARGS4
ls -la gpurun_out/r02f_* | head -30
