"""SASS evidence for the hot kernels: disassembles lobpcg_b200/_lib/liblobpcg_b200.so with cuobjdump
and writes, per kernel family, the instruction mix (tensor-pipe, async-copy, shared/global memory, barrier mnemonics) and a
short excerpt of the main loop around the first tensor instruction.

    python tools/sass_summary.py > profiles/sass_r02.md
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "lobpcg_b200" / "_lib" / "liblobpcg_b200.so"
FAMILIES = ["gram_wl_kernel", "tall_nn_persist_kernel", "tall_nn_dmma_kernel", "gram_dmma_kernel", "strip_gram_kernel",
            "gram_zmma_kernel", "tall_nn_zmma_kernel", "oz_gram_kernel", "oz_gram_cluster_kernel", "oz_nn_kernel", "oz_split_kernel", "gram_tc5_tma_kernel", "gram_tc5_kernel", "nn_tc5_kernel", "gram_tf32_kernel",
            "stencil_kernel", "csr_staged_kernel", "csr_win_kernel", "csr_kernel", "residual_kernel", "residual_monitor_kernel"]
KEYS = ["DMMA", "HMMA", "UTCHMMA", "UTCIMMA", "UTCMMA", "UTCBAR", "UCGABAR_ARV", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "LDGSTS", "LDG", "STG", "LDS", "STS",
        "BAR", "DFMA", "DADD", "FFMA", "SYNCS", "WARPSYNC"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    funcs = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        if cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            funcs[cur].append(line)
    dem = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(funcs, dem))
    print("# SASS summary of liblobpcg_b200.so (sm_100a), round 2\n")
    print("`python tools/sass_summary.py` — `cuobjdump -sass` of the shipped library; instruction counts are static (per kernel "
          "instance), the largest instance of each family is shown.\n")
    tot = collections.Counter()
    for f, lines in funcs.items():
        for l in lines:
            m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if m:
                tot[m.group(1).split(".")[0]] += 1
    print("Whole library: " + ", ".join(f"`{k}` {tot[k]}" for k in KEYS if tot[k]) + "\n")
    for fam in FAMILIES:
        cand = [(-names[k].count("false"), ("<double" in names[k] or "double" in names[k]), len(v), k) for k, v in funcs.items() if fam in names[k]]
        if not cand:
            continue
        f = max(cand)[-1]      # the vectorised (16-byte copies) double instance: the one the solver runs
        lines = funcs[f]
        cnt = collections.Counter()
        first = None
        for i, l in enumerate(lines):
            m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if not m:
                continue
            op = m.group(1)
            base = op.split(".")[0]
            cnt[base] += 1
            if first is None and base in (("UTMALDG",) if ("tma" in fam or "cluster" in fam) else ("DMMA", "HMMA", "UTCHMMA", "UTCIMMA", "UTCMMA")):
                first = i
        print(f"## {fam}\n")
        print(f"`{names[f][:160]}`: {len(lines)} instructions, {len(cand)} instance(s) in the library\n")
        print("| " + " | ".join(k for k in KEYS if cnt[k]) + " |")
        print("|" + "---|" * sum(1 for k in KEYS if cnt[k]))
        print("| " + " | ".join(str(cnt[k]) for k in KEYS if cnt[k]) + " |\n")
        if first is not None:
            lo, hi = max(0, first - 6), min(len(lines), first + 14)
            print("```")
            for l in lines[lo:hi]:
                print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l.rstrip())[:150])
            print("```\n")


if __name__ == "__main__":
    main()
