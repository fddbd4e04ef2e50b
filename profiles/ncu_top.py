"""Summarise an ncu --page source --csv dump: top stall-sampled instructions."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc, iss, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = [r for r in rows[2:] if len(r) > iss and r[iss].strip().isdigit()]
tot = sum(int(r[iss]) for r in data)
print("total samples", tot, "instructions", len(data))
for r in sorted(data, key=lambda r: -int(r[iss]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{int(r[iss]) / tot * 100:5.1f}%  ex={r[iex]:>10}  {r[isrc][:120]}")
