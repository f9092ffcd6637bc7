"""Static instruction budget of the pair kernels from the SASS of the built library (no GPU
needed): instructions between the loop head and the loop branch of the main neighbour loop,
per list entry (includes the rarely taken blocks that sit inside the loop).

    python scripts/sass_budget.py > profiles/r01_sass_budget.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("TMD_B200_LIB") or os.path.join(ROOT, "torchmd_b200", "libtmd_b200.so")


def functions():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, table = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            table[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?)\s*;", line)
        if m and cur:
            table[cur].append((int(m.group(1), 16), m.group(2)))
    return table


def resources():
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    res, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and cur:
            res[cur] = tuple(int(x) for x in m.groups())
    return res


def main_loop(ins):
    """(start, end) indices of the innermost backward branch that contains an LDG.E.128 gather."""
    best = None
    for k, (addr, text) in enumerate(ins):
        m = re.search(r"BRA\s+(?:\S+,\s*)?0x([0-9a-f]+)", text)
        if m and int(m.group(1), 16) < addr:
            start = next(i for i, (a, _) in enumerate(ins) if a >= int(m.group(1), 16))
            body = ins[start : k + 1]
            if any("LDG.E.128" in t for _, t in body) and any("MUFU.RSQ" in t for _, t in body):
                if best is None or (k - start) < (best[1] - best[0]):
                    best = (start, k)
    return best


def summarise(name, ins, res):
    loop = main_loop(ins)
    if loop is None:
        return
    body = [t for _, t in ins[loop[0] : loop[1] + 1]]
    npairs = sum("LDG.E.128" in t for t in body)
    mufu = sum(t.startswith("MUFU") for t in body)
    packed = sum(any(p in t for p in ("FFMA2", "FMUL2", "FADD2")) for t in body)
    spill = sum(("STL" in t or "LDL" in t) for t in body)
    total = len(body)
    r = res.get(name, (0, 0, 0))
    print(f"{name}\n  registers {r[0]}, stack {r[1]} B, static smem {r[2]} B; main loop {total} instructions for {npairs} list entries per lane"
          f" = {total / npairs:.1f} per entry (MUFU {mufu // npairs} per entry, packed fp32x2 instructions {packed}, local-memory ops in the loop: {spill})")


def main():
    table, res = functions(), resources()
    print("Static SASS budget of the pair kernels (cuobjdump of torchmd_b200/libtmd_b200.so, sm_100a; no GPU involved).")
    print("Every lane executes the whole loop body when any lane of its warp has an in-cutoff partner, so the per-entry count is")
    print("the warp-level issue cost per list slot.  k_pair = float separations (default, measured 160 us at 99,999 atoms);")
    print("k_pair_fx = fixed-point separations, k_pair_fx2 = the same with packed fp32x2 arithmetic (both opt-in, not yet measured).  Template args: <ENERGY, PERIODIC, SAFE, MODE> and")
    print("<ENERGY, MODE, SMALLT>.\n")
    for name in sorted(table):
        if "k_pair" in name and ("ILb0E" in name or "k_pair_fx2" in name):
            summarise(name, table[name], res)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        LIB = sys.argv[1]
    main()
