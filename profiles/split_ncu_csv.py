"""Splits a `ncu --page raw --csv` export that holds several launches into one file per launch.
   python profiles/split_ncu_csv.py in.csv out_a.csv out_b.csv ...   (launch i goes to the i-th output)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, units, body = rows[start], rows[start + 1], rows[start + 2:]
for r, out in zip(body, sys.argv[2:]):
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(hdr)
        w.writerow(units)
        w.writerow(r)
