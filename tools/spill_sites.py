"""Where do spills sit relative to the math loop?  usage: python tools/spill_sites.py [name-substring]"""
import subprocess, sys, re
pat = sys.argv[1] if len(sys.argv) > 1 else "tma_kernel"
out = subprocess.run(["cuobjdump", "-sass", "/root/repo/blis_b200/libblis_b200.so"], capture_output=True, text=True).stdout
fn = None; rows = {}
for line in out.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m: fn = m.group(1); rows[fn] = []; continue
    if fn and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line): rows[fn].append(line)
for fn, ls in rows.items():
    d = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    if pat not in d: continue
    math = [i for i, l in enumerate(ls) if re.search(r"\b(DMMA|FFMA2)\b", l)]
    sp = [(i, "STL" if "STL" in l else "LDL") for i, l in enumerate(ls) if re.search(r"\b(STL|LDL)", l)]
    if not math: continue
    # densest math region = main loop: instructions between the first and last backward branch target that contain >50% math
    lo, hi = math[0], math[-1]
    inside = [s for s in sp if lo <= s[0] <= hi]
    print(d.split("(")[0][:110], "| insts", len(ls), "| math", lo, "-", hi, "| spills", len(sp), "inside-math-span", len(inside), [s[0] for s in inside][:12])
