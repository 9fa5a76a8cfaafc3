"""Instruction mix of the hot loop of a kernel in a built .so (cuobjdump -sass).

    python tools/sass_mix.py <lib.so> <substring of the mangled kernel name> [--dump out.sass]

The hot loop is taken as the innermost backward-branch region that contains the count atomic
(ATOMG.E.ADD) or, failing that, the largest number of FP64 instructions.  Prints counts per opcode class."""
import re
import subprocess
import sys
from collections import Counter


def functions(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, name, res = [], None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                res[name] = cur
            name, cur = m.group(1), []
        elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            m2 = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
            if m2:
                cur.append((int(m2.group(1), 16), m2.group(2).strip()))
    if name:
        res[name] = cur
    return res


def classify(ins):
    op = ins.split()[0]
    if op.startswith("@"):
        op = ins.split()[1]
    base = op.split(".")[0]
    if base in ("DMUL", "DADD", "DFMA", "DSETP", "MUFU"):
        return base if base != "MUFU" else "MUFU"
    if base in ("F2F", "F2I", "I2F", "D2I", "F2FP", "FRND"):
        return "convert"
    if base in ("LDCU", "LDC", "UMOV", "ULDC", "S2UR", "R2UR", "UIADD3", "ULEA", "UISETP", "USEL", "UIMAD", "ULOP3", "USHF", "UPLOP3"):
        return "const/uniform (" + base + ")"
    if base in ("ATOMG", "ATOM", "RED", "REDG", "LDG", "STG", "LD", "ST", "ATOMS", "LDS", "STS"):
        return "memory (" + op + ")"
    if base in ("BRA", "BSSY", "BSYNC", "CALL", "RET", "EXIT", "WARPSYNC", "YIELD", "NOP", "BREAK", "BMOV"):
        return "control"
    return "int/other"


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    fs = functions(lib)
    names = [n for n in fs if pat in n]
    if not names:
        sys.exit(f"no function matching {pat}: {list(fs)[:20]}")
    for n in names:
        ins = fs[n]
        if "--dump" in sys.argv:
            with open(sys.argv[sys.argv.index("--dump") + 1], "w") as f:
                for a, t in ins:
                    f.write(f"/*{a:04x}*/ {t}\n")
        loops = []
        for k, (a, t) in enumerate(ins):
            m = re.search(r"BRA(?:\.\w+)* (?:\w+, )?0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                lo = int(m.group(1), 16)
                body = [(x, y) for x, y in ins if lo <= x <= a]
                loops.append((lo, a, body))
        best = None
        for lo, hi, body in loops:
            nat = sum("ATOMG.E.ADD" in t for _, t in body)
            ndp = sum(t.split()[-0].startswith(("DMUL", "DADD")) or " DMUL" in " " + t or " DADD" in " " + t for _, t in body)
            score = (nat > 0, -len(body)) if nat else (False, ndp)
            if best is None or score > best[0]:
                best = (score, lo, hi, body)
        _, lo, hi, body = best
        c = Counter(classify(t) for _, t in body)
        print(f"{n}\n  hot loop 0x{lo:x}..0x{hi:x}: {len(body)} instructions")
        for k, v in sorted(c.items(), key=lambda kv: -kv[1]):
            print(f"    {v:5d}  {k}")


if __name__ == "__main__":
    main()
