"""Per-source-line stall samples from an .ncu-rep (dev tool): python tools/ncu_lines.py rep [launch_idx] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; idx = sys.argv[2] if len(sys.argv) > 2 else "0"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", idx, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; cur_file = ""; lines = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 4 and r[0] == "Line No": hdr = r; si = r.index("Warp Stall Sampling (All Samples)"); ii = r.index("Instructions Executed"); continue
    if hdr and len(r) > si and r[0] not in ("", "Line No"):
        try: lines.append((int(r[si]), int(r[ii]), cur_file, r[0], r[1].strip()[:120]))
        except ValueError: pass
tot = sum(l[0] for l in lines)
print("total samples", tot)
for s, n, f, ln, src in sorted(lines, key=lambda l: -l[0])[:top]:
    print(f"{s:7d} {100*s/max(tot,1):5.1f}% inst={n:8d} {f}:{ln}: {src}")
