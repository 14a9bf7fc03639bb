"""Per-CUDA-source-line summary of an ncu report (needs -lineinfo and --import-source on):
python tools/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h = next(i for i, r in enumerate(rows[:10]) if '# Samples' in r)
hdr = rows[h]; ci = hdr.index('# Samples'); ii = hdr.index('Instructions Executed')
def num(x):
    try: return int(x)
    except Exception: return 0
lines = [r for r in rows[h + 1:] if len(r) > ii and r[0].strip().isdigit() and r[2] == '-']
tot_s = sum(num(r[ci]) for r in lines) or 1; tot_i = sum(num(r[ii]) for r in lines) or 1
print(f"total samples {tot_s}  total warp-instructions {tot_i}")
for r in sorted(lines, key=lambda r: -num(r[ci]))[:top]:
    print(f"{num(r[ci])/tot_s*100:5.1f}% samp {num(r[ii])/tot_i*100:5.1f}% inst | L{r[0]:>4} {r[1].strip()[:118]}")
