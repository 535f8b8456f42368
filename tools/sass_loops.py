"""Opcode histogram of the loops of one kernel in a cuobjdump -sass listing.

usage: sass_loops.py <sass.txt> <mangled-name substring> [min_loop_instrs]
Finds backward branches, reports each loop (address range, instruction count) with its opcode histogram."""
import re, sys, collections

def parse(path, key):
    ins, on = [], False
    for line in open(path):
        if "Function :" in line:
            on = key in line
            continue
        if not on:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins

def opcode(text):
    t = text.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0]

def main():
    path, key = sys.argv[1], sys.argv[2]
    minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    ins = parse(path, key)
    print("instructions:", len(ins))
    addr_idx = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        if opcode(t).startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt <= a and tgt in addr_idx and i - addr_idx[tgt] >= minlen:
                    loops.append((addr_idx[tgt], i))
    for (s, e) in loops:
        h = collections.Counter(opcode(t).split(".")[0] for _, t in ins[s:e + 1])
        full = collections.Counter(opcode(t) for _, t in ins[s:e + 1])
        print(f"loop {ins[s][0]:#x}..{ins[e][0]:#x}: {e - s + 1} instrs")
        print("  ", ", ".join(f"{k}:{v}" for k, v in h.most_common()))
        if "-v" in sys.argv:
            print("  ", ", ".join(f"{k}:{v}" for k, v in full.most_common()))
    tot = collections.Counter(opcode(t).split(".")[0] for _, t in ins)
    print("whole kernel:", ", ".join(f"{k}:{v}" for k, v in tot.most_common(40)))

main()
