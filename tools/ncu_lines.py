"""Per-source-line executed warp instructions of a kernel from a .ncu-rep captured with --import-source on (run here, no GPU).
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[2] == "-" and r[0].isdigit():
        ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
        try: lines.append((int(r[ie]), int(r[isamp] or 0), cur_file, int(r[0]), r[1].strip()[:110]))
        except ValueError: pass
tot = sum(l[0] for l in lines); tots = sum(l[1] for l in lines)
print(f"total warp instructions {tot}, samples {tots}")
for e, smp, f, n, src in sorted(lines, reverse=True)[:top]:
    print(f"{100*e/tot:5.1f}% inst {100*smp/max(tots,1):5.1f}% smp  {f}:{n:<4d} {src}")
