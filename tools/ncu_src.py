"""Summarises the source page of an .ncu-rep: top SASS lines by stall samples (usage: ncu_src.py file.ncu-rep [N])."""
import csv, subprocess, sys
rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr, data = rows[hdr_i], rows[hdr_i + 1:]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, "rows", len(data))
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:top_n]
for i in sorted(top):
    r = data[i]
    st = {hdr[j][6:]: int(r[j]) for j in stall if int(r[j]) > 0}
    print(i, r[isamp], r[iex], r[ia].strip()[:70], st)
