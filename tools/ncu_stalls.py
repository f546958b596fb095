"""Summarise an `ncu --page source --csv` dump: top SASS lines by stall samples.
Usage: ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_stalls.py src.csv [N]"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1]))]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
# the dump may hold several kernels: each starts with a "Kernel Name" row then a header row
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        H = rows[i + 1]
        idx = {h: k for k, h in enumerate(H)}
        j = i + 2
        data = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(H) and rows[j][0] != "Address":
                data.append(rows[j])
            j += 1
        tot = sum(int(r[idx["# Samples"]]) for r in data)
        print("==", name[:90], "| samples", tot)
        stall_cols = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
        for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:top_n]:
            st = {h: int(r[idx[h]]) for h in stall_cols if int(r[idx[h]]) > 0}
            st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
            print(r[idx["Address"]][-5:], r[idx["Source"]][:58].ljust(58), r[idx["# Samples"]].rjust(6),
                  r[idx["Instructions Executed"]].rjust(9), st)
        agg = {h: sum(int(r[idx[h]]) for r in data) for h in stall_cols}
        print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
        i = j
    else:
        i += 1
