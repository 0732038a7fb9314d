"""Print the hottest SASS lines (by executed warp instructions and by stall samples) of a kernel in an ncu report.
usage: python tools_ncu_src.py report.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys, io
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several kernels/launches may be concatenated: take the first block
blocks = []
cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
b = blocks[0]
hdr = b["rows"][0]
ci, ce, cs = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
items = []
for i, r in enumerate(b["rows"][1:]):
    try:
        items.append((i, r[ci].strip(), float(r[ce]), float(r[cs])))
    except Exception:
        pass
tot_e = sum(x[2] for x in items); tot_s = sum(x[3] for x in items)
print(b["name"][:100]); print("total warp instr", tot_e, "samples", tot_s, "n_sass", len(items))
print("--- by stall samples")
for i, sx, e, s in sorted(items, key=lambda x: -x[3])[:top]:
    print(f"{i:5d} {100*s/tot_s:5.1f}% smp {100*e/tot_e:5.1f}% exe  {sx[:90]}")
