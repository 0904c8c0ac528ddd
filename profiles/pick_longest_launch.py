"""Index (among the rows of an `ncu --metrics gpu__time_duration.sum --csv` launch list) of the longest launch:
the value to pass as --launch-skip of a second, `--set full` pass with the same -k filter."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
iv = hdr.index("Metric Value")
vals = [float(r[iv].replace(",", "")) for r in rows[h + 1:] if len(r) == len(hdr)]
print(max(range(len(vals)), key=lambda i: vals[i]))
