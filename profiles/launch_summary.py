"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name."""
import collections
import csv
import re
import sys

def main(path, nlev=13):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.OrderedDict(); tot = 0; seq = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('b200::', '')
        v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']
        v = v / 1e3 if unit == 'ns' else v * 1e3 if unit == 'ms' else v
        a = agg.setdefault(name, [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v); tot += v
        seq.append((name, v, row['Grid Size']))
    for k, (c, t, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        print(f"{k[:50]:50s} n={c:5d} total={t/1e3:9.3f} ms  avg={t/c:9.1f} us max={mx:9.1f} us  share={t/tot*100:5.1f}%")
    for kn in ('void k_flow<1, 0>', 'void k_flow<0, 0>'):  # forward / backward dataflow sweep: one launch per sweep
        idx = [i for i, (n, _, _) in enumerate(seq) if n == kn][-nlev:]
        print(kn, 'last launches (us, grid):', [(round(seq[i][1], 1), seq[i][2].split(',')[0].strip('(')) for i in idx])

if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 13)
