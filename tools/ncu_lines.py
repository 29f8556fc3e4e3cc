"""Per-source-line instruction counts and stall samples from an ncu report (needs -lineinfo and --import-source on).
    python tools/ncu_lines.py gpurun_out/k2.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
out = {}; fname = None; hdr = None; warps = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) == 2 and r[0] == "Function Name":
        if out and "first_fn" in out: break
        continue
    if r and r[0] == "Line No": hdr = r; ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr and len(r) >= len(hdr) and r[0].isdigit():
        try: out[(fname, int(r[0]))] = (int(r[ii] or 0), int(r[si] or 0), r[1].strip())
        except ValueError: pass
tot = sum(v[0] for v in out.values()); ts = sum(v[1] for v in out.values())
w = max(v[0] for v in out.values() if v[0]) if out else 1
print(f"total warp-instructions {tot}, samples {ts}")
cum = 0
for (f, ln), (n, s, src) in sorted(out.items(), key=lambda kv: -kv[1][0])[:top]:
    cum += n
    print(f"{100*n/tot:5.1f}% instr {100*s/max(ts,1):5.1f}% stall  {f}:{ln:<4} {src[:105]}")
print(f"shown {100*cum/tot:.1f}% of instructions")
