"""Aggregate warp-stall samples per CUDA source line from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None; hdr = None; out = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; si = hdr.index("# Samples"); ei = hdr.index("Instructions Executed"); continue
    if hdr and len(r) > si and r[0] not in ("", "Line No"):
        try: out.append((int(r[si]), int(r[ei]), cur_file, r[0], r[1].strip()[:110]))
        except ValueError: pass
tot = sum(o[0] for o in out)
print("total samples", tot)
for smp, ex, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{smp:6d} {100*smp/tot:5.1f}%  exec={ex:8d}  {f}:{ln}  {src}")
