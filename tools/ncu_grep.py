import csv, io, re, subprocess, sys
rep, pat = sys.argv[1], re.compile(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    for i, h in enumerate(hdr):
        if pat.search(h) and r[i] not in ("", "0"):
            print(f"  {h:100s} {units[i]:14s} {r[i]}")
