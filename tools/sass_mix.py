"""Dev helper (no GPU needed): static instruction mix of the innermost loops of the hot kernels, from
`cuobjdump -sass` of the built library.  For every kernel the loops (backward branches) with the most instructions
are listed with their opcode histogram - the evidence behind the "instruction-issue bound" statements of DESIGN.md.
usage: python tools/sass_mix.py [out.md]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "msu-latentafis_b200", "lib", "liblatentafis_b200.so")
KERNELS = ["tex_rowmax_kernel", "minu_sim_kernel", "minu_select_kernel", "graph_minu_sparse_kernel",
           "graph_tex_sparse_kernel", "compnet_l1_kernel", "compnet_l234_kernel"]


def loops(sass):
    ins = []
    for l in sass.splitlines():
        m = re.search(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr = {a: i for i, (a, _) in enumerate(ins)}
    out = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr:
                out.append(ins[addr[tgt]:i + 1])
    return out, len(ins)


def opcode(t):
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    return t.split()[0].split(".")[0]


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    names = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    funcs = re.findall(r"Function : (\S+)", names)
    lines = ["# Static instruction mix of the hot kernels' loops (cuobjdump -sass, sm_100a)\n",
             "`python tools/sass_mix.py` - innermost loops (backward branches with no other loop inside) of at least 24 instructions, the largest five per kernel; counts are SASS "
             "instructions per loop iteration.\n"]
    for k in KERNELS:
        f = next((x for x in funcs if k in x and "jobs" not in x and "slow" not in x), None)
        if not f:
            continue
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", f, SO], capture_output=True, text=True).stdout
        ls, total = loops(sass)
        # innermost loops only (no other loop inside), the largest first: these are the steady-state bodies
        rng = [(b[0][0], b[-1][0]) for b in ls]
        inner = [b for b, (lo, hi) in zip(ls, rng)
                 if len(b) >= 24 and not any((lo2 > lo or hi2 < hi) and lo2 >= lo and hi2 <= hi for lo2, hi2 in rng)]
        inner.sort(key=len, reverse=True)
        lines.append(f"\n## {k} ({total} instructions)\n")
        lines.append("| innermost loop, instructions | mix |\n|---|---|")
        seen = set()
        for body in inner[:8]:
            key = (body[0][0], body[-1][0])
            if key in seen or len(seen) >= 5:
                continue
            seen.add(key)
            c = collections.Counter(opcode(t) for _, t in body)
            mix = ", ".join(f"{op} {n}" for op, n in c.most_common(9))
            lines.append(f"| {len(body)} | {mix} |")
    text = "\n".join(lines) + "\n"
    if out_path:
        open(out_path, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
