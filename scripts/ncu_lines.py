"""Development aid: per-source-line stall samples / instruction counts from an ncu report.
usage: python scripts/ncu_lines.py report.ncu-rep [kernel-substring] [min-fraction]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
def I(x):
    try: return int(x)
    except ValueError: return 0
per, instr, tot, cur, fname, func = {}, {}, {}, None, None, None
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": func = r[1]; continue
    if len(r) < 8 or r[0] == "Line No": continue
    if r[0]:
        cur = (func, fname, I(r[0])); per[cur] = [I(r[4]), r[1][:110]]; tot[func] = tot.get(func, 0) + I(r[4])
    else:
        instr[cur] = instr.get(cur, 0) + I(r[7])
for f in tot:
    if want not in f: continue
    ti = sum(v for k, v in instr.items() if k[0] == f)
    print("==", f[:90], "samples", tot[f], "warp-instr", ti)
    for k, (s_, t) in sorted(per.items(), key=lambda kv: (kv[0][1], kv[0][2])):
        if k[0] == f and s_ > tot[f] * frac:
            print(f"{k[1]}:{k[2]:4d} {100 * s_ / tot[f]:5.1f}% {instr.get(k, 0) / max(ti, 1) * 100:5.1f}%i  {t}")
