"""Development helper: executed warp instructions per CUDA source line from an ncu report.

    python tools/ncu_lines.py <report.ncu-rep> [top]
(ncu -i <report> --page source --csv --print-source cuda,sass; needs -lineinfo at compile time)"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur, hdr, agg, tot = None, None, {}, 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or not r[0]:
        continue  # SASS rows have an empty line number; CUDA rows carry the aggregate of their SASS
    try:
        n = int(r[hdr.index("Instructions Executed")]); ti = int(r[hdr.index("Thread Instructions Executed")])
    except (ValueError, IndexError):
        continue
    if n:
        agg[(cur, int(r[0]))] = (n, ti, r[1].strip()[:100])
        tot += n
print("total warp instructions", tot)
for k, (n, ti, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:4d} {100 * n / tot:5.2f}% lanes {ti / n:5.1f} | {src}")
