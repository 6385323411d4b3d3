"""Top source lines by warp-stall samples from an ncu report (needs -lineinfo, --import-source on)."""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, kernel_seen, hdr, items, total = None, 0, None, [], 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            s = int(d.get("# Samples", "0") or 0)
        except ValueError:
            continue
        total += s
        stall = {k: int(v or 0) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit()}
        top = sorted(stall.items(), key=lambda kv: -kv[1])[:3]
        items.append((s, cur_file, r[0], r[1].strip()[:90], d.get("Instructions Executed", ""), top))
items.sort(key=lambda x: -x[0])
print("total samples", total)
for s, f, ln, src, inst, top in items[:topn]:
    print(f"{s:6d} {100*s/max(total,1):5.1f}%  {f}:{ln:>4}  inst={inst:>8}  {src}   {[(k[6:], v) for k, v in top if v]}")
