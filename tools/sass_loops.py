"""Static SASS loop report for one kernel: back-edges, instructions per loop body and their opcode mix.

    cuobjdump -sass file.o | python tools/sass_loops.py <substring of the mangled kernel name> [--dump]

Used to count issue slots per 16-step tile of the scan kernel's warp roles without a GPU.
"""
import re, sys, collections

pat = sys.argv[1]
dump = "--dump" in sys.argv
ins = []          # (addr, text)
cur = False
for line in sys.stdin:
    if "Function :" in line:
        cur = pat in line
        if cur:
            print(line.strip())
        continue
    if not cur:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s*/\*", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_idx = {a: i for i, (a, _) in enumerate(ins)}
print("instructions:", len(ins))
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?`?\(?0x([0-9a-f]+)\)?", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_idx:
            loops.append((addr_idx[tgt], i))
def opname(t):
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    return t.split()[0].split(".")[0]
for s, e in sorted(loops):
    body = ins[s:e + 1]
    h = collections.Counter(opname(t) for _, t in body)
    print(f"loop {ins[s][0]:#x}..{ins[e][0]:#x}: {len(body)} instrs  " + " ".join(f"{k}:{v}" for k, v in h.most_common(24)))
    if dump:
        for a, t in body:
            print(f"    {a:#06x} {t}")
