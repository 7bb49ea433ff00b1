"""Top stall-sample SASS lines of one kernel instance in an .ncu-rep (ncu --page source --csv).

    python tools/ncu_hot.py rep.ncu-rep <substring of the demangled kernel name> [top] [occurrence]
"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
occ = int(sys.argv[4]) if len(sys.argv) > 4 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None:
        cur["rows"].append(r)
sel = [b for b in blocks if kern in b["name"].replace("(int)", "").replace("(bool)", "")]
if not sel:
    print("kernels in report:"); [print("  ", b["name"][:150]) for b in blocks]; sys.exit(1)
b = sel[occ]
hdr = b["hdr"]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for r in b["rows"]:
    try: data.append((int(r[i_s] or 0), len(data), r[i_src].strip(), int(r[i_ex] or 0)))
    except (ValueError, IndexError): pass
tot = sum(d[0] for d in data) or 1
print(f"{b['name'][:120]}\n{tot} stall samples over {len(data)} SASS instructions, {sum(d[3] for d in data)} warp-instructions executed")
for s, idx, src, ex in sorted(data, reverse=True)[:top]:
    print(f"{s:6d} {100*s/tot:5.1f}%  #{idx:4d} ex={ex:8d}  {src[:120]}")
