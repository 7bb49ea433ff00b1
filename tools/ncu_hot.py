"""Top stall-sample SASS lines of one kernel in an .ncu-rep (ncu --page source --csv)."""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for r in rows[hi + 1:]:
    if len(r) <= i_ex or r[0] in ("Address", "Kernel Name"):
        if r and r[0] == "Kernel Name" and data: break     # next kernel instance
        continue
    try: data.append((int(r[i_s] or 0), len(data), r[i_src].strip(), int(r[i_ex] or 0)))
    except ValueError: pass
tot = sum(d[0] for d in data) or 1
print(f"{kern}: {tot} stall samples over {len(data)} SASS instructions, {sum(d[3] for d in data)} warp-instructions executed")
for s, idx, src, ex in sorted(data, reverse=True)[:top]:
    print(f"{s:6d} {100*s/tot:5.1f}%  #{idx:4d} ex={ex:8d}  {src[:120]}")
