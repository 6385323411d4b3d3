"""SASS digest of the product library: per kernel, instruction counts that prove (or disprove) the design claims —
128-bit streaming loads, warp-level CREDUX/REDUX/VOTE/SHFL primitives, block barriers, local-memory (spill) traffic,
fences, atomics, code size.  Works without a GPU (cuobjdump on lib/libposebyte_b200.so).

usage: python tools/sass_digest.py > profiles/r02_sass_digest.md"""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "yolo-pose-cpp_b200", "lib", "libposebyte_b200.so")
PAT = collections.OrderedDict([
    ("LDG 128-bit", r"\bLDG\.[A-Z0-9.]*128"), ("LDG (all)", r"\bLDG\."), ("STG", r"\bSTG\."), ("LDS", r"\bLDS"), ("STS", r"\bSTS"),
    ("CREDUX", r"\bCREDUX"), ("REDUX", r"\bREDUX"), ("VOTE", r"\bVOTE"), ("SHFL", r"\bSHFL"), ("MATCH", r"\bMATCH"),
    ("UBLKCP (bulk copy)", r"\bUBLKCP"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("BAR.SYNC", r"\bBAR\.SYNC"), ("ATOMS", r"\bATOMS"), ("ATOMG/RED", r"\b(ATOMG|RED)\."), ("MEMBAR", r"\bMEMBAR"), ("CCTL", r"\bCCTL"),
    ("LDL (spill)", r"\bLDL"), ("STL (spill)", r"\bSTL"), ("MUFU", r"\bMUFU"), ("FFMA", r"\bFFMA"), ("HMMA/UTCMMA", r"\b(HMMA|UTC.MMA|UTCHMMA|IMMA)"),
])
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); kernels[cur] = []
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        kernels[cur].append(line)
def demangle(n):
    try:
        d = subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip().replace("void ", "").replace("pb::", "").replace("(int)", "").replace("(bool)", "")
        depth, out = 0, ""
        for ch in d:                      # cut at the parameter list, keep the template arguments
            if ch == "<": depth += 1
            if ch == ">": depth -= 1
            if ch == "(" and depth == 0: break
            out += ch
        return out
    except Exception:
        return n
print("# SASS digest — lib/libposebyte_b200.so (sm_100a), `python tools/sass_digest.py`\n")
print("Counts of SASS instructions per kernel (static).  No tensor-core instruction anywhere: this path has no dense contraction.")
print("FFMA only where the source asks for it (`--fmad=false`: no contraction of separate multiplies and adds).\n")
print("| kernel | instr | " + " | ".join(PAT) + " |")
print("|---|---|" + "---|" * len(PAT))
for k, lines in kernels.items():
    if not lines:
        continue
    name = demangle(k)
    if not name.startswith(("pb_", "pose_nms", "nms_legacy", "auction_batch", "kf3", "greedy", "assign_legacy", "letterbox", "pose_distance")) and "kernel" not in name:
        continue
    print(f"| `{name}` | {len(lines)} | " + " | ".join(str(sum(1 for l in lines if re.search(p, l))) for p in PAT.values()) + " |")
