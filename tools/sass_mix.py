"""Instruction mix of the hottest loop of a kernel (dev tool): python tools/sass_mix.py <demangled-substring>
The hot loop is taken as the backward-branch span that contains the most FFMA2/DMMA instructions."""
import re, subprocess, sys, collections
pat = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", "/root/repo/blis_b200/libblis_b200.so"], capture_output=True, text=True).stdout
fn = None; rows = {}
for line in out.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m: fn = m.group(1); rows[fn] = []; continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if fn and m: rows[fn].append((int(m.group(1), 16), m.group(2).strip()))
for fn, ls in rows.items():
    d = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    if pat not in d: continue
    addr2idx = {a: i for i, (a, _) in enumerate(ls)}
    best = None
    for i, (a, ins) in enumerate(ls):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", ins)
        if m:
            t = int(m.group(1), 16)
            if t < a and t in addr2idx:
                j = addr2idx[t]
                n = sum(1 for _, x in ls[j:i + 1] if re.search(r"\b(FFMA2|DMMA|FFMA)\b", x))
                if best is None or n > best[0]: best = (n, j, i)
    if not best: continue
    n, j, i = best
    mix = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x).split()[0].split(".")[0] for _, x in ls[j:i + 1])
    print(d.split("(")[0], "| loop insts", i - j + 1, dict(mix.most_common(12)))
