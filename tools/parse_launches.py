"""Print one steady-state step of an ncu launch list (gpu__time_duration.sum CSV)."""
import csv, sys
path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows = list(csv.reader(open(path)))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        h = r; start = i + 1; break
ki = h.index('Kernel Name'); vi = h.index('Metric Value'); idi = h.index('ID')
seq = [(int(r[idi]), r[ki].split('(')[0].replace('void ', '').replace('latte::<unnamed>::', '')[:70], float(r[vi]) / 1000)
       for r in rows[start:] if len(r) > vi]
gi = [i for i, s in enumerate(seq) if 'pair_gemm' in s[1]]
a = gi[which] + 1; b = gi[which + 1] + 1
tot = 0
for s in seq[a:b]:
    print(s[0], s[1], round(s[2], 1)); tot += s[2]
print('sum', round(tot, 1))
