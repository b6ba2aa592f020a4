"""tools/ncu_stalls.py SOURCE.csv [window=100] [min_samples=90] -- where a kernel's warps wait: the warp
stall samples of an `ncu --set full --import-source on` capture (exported with
`ncu -i X.ncu-rep --page source --csv`), totalled per stall reason, per window of SASS instructions, and
the single instructions that collect the most samples.  Tuning aid; its output is kept under profiles/."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
# (a report with several kernels exports one section per kernel: the first one is analysed, or the
#  first whose name contains the 4th argument)
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
want = sys.argv[4] if len(sys.argv) > 4 else ''
sec = next((j for j in range(len(starts) - 1) if want in rows[starts[j]][1]), 0)
rows = rows[starts[sec]:starts[sec + 1]]
hdr, data = rows[1], [r for r in rows[2:] if len(r) == len(rows[1])]
W = int(sys.argv[2]) if len(sys.argv) > 2 else 100
MIN = int(sys.argv[3]) if len(sys.argv) > 3 else 90
iS, iI, iSrc = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
st = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[iS]) for r in data)
toti = sum(int(r[iI]) for r in data)
print(rows[0][1])
print('SASS instructions %d, samples %d, warp-instructions executed %d' % (len(data), tot, toti))
for i in st:
    s = sum(int(r[i]) for r in data)
    if s > 0.01 * tot:
        print('  %-24s %5.1f %%' % (hdr[i], 100.0 * s / tot))
print('--- windows of %d SASS instructions: share of samples, share of executed instructions, top stall reasons' % W)
for a in range(0, len(data), W):
    ch = data[a:a + W]
    s = sum(int(r[iS]) for r in ch)
    n = sum(int(r[iI]) for r in ch)
    if s > 0.005 * tot:
        top = sorted(st, key=lambda i: -sum(int(r[i]) for r in ch))[:3]
        print('%5d-%5d samples %5.1f %%  instructions %5.1f %%  %s' % (a, a + W - 1, 100.0 * s / tot, 100.0 * n / toti,
              ' '.join('%s=%.0f%%' % (hdr[i][6:], 100.0 * sum(int(r[i]) for r in ch) / max(s, 1)) for i in top)))
print('--- instructions with >= %d samples' % MIN)
for k, r in enumerate(data):
    s = int(r[iS])
    if s >= MIN:
        top = sorted(st, key=lambda i: -int(r[i]))[:2]
        print('%5d %-56s samples %5d (%4.1f %%) executed %9d  %s' % (k, r[iSrc].strip()[:56], s, 100.0 * s / tot, int(r[iI]),
              ' '.join('%s=%s' % (hdr[i][6:], r[i]) for i in top if int(r[i]) > 0)))
