#!/usr/bin/env python3
"""Static instruction mix of the innermost hot loop of a kernel, from `cuobjdump -sass` of the shipped library.

    python profiles/scripts/sass_loop_mix.py ngsf-hmm_b200/libngsfhmm_b200.so 'freq_emission_warpILi8ELi13ELi1' [K]

Finds every backward branch of the function (a loop), takes the INNERMOST loop with the most FP64 instructions, and prints its
mix by pipe.  With K (individuals per lane) it also prints FP64 instructions and flops per individual-pass - the
figures bench.py's roofline uses (FREQ_INSTR_PER_IND_PASS, FREQ_FLOPS_PER_IND_PASS) - so that they can be checked
against the binary without a GPU.  No profiler involved: this is the code, not a measurement."""
import collections
import re
import subprocess
import sys


def functions(so):
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    cur, out = None, {}
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1); out[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur:
            out[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return out


def mnemonic(ins):
    ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
    return ins.split()[0]


def main():
    so, pat = sys.argv[1], sys.argv[2]
    K = int(sys.argv[3]) if len(sys.argv) > 3 else None
    fns = {k: v for k, v in functions(so).items() if pat in k}
    for name, code in fns.items():
        loops = []
        for addr, ins in code:
            m = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?(?:`\(\S+\)|0x([0-9a-f]+))", ins)
            if m and m.group(1) and int(m.group(1), 16) <= addr:
                loops.append((int(m.group(1), 16), addr))
        best, best_n = None, -1
        for lo, hi in loops:
            if any((l2, h2) != (lo, hi) and lo <= l2 and h2 <= hi for l2, h2 in loops):
                continue                                   # not innermost
            body = [mnemonic(i) for a, i in code if lo <= a <= hi]
            n = sum(b.startswith(("DFMA", "DMUL", "DADD")) for b in body)
            if n > best_n:
                best, best_n = (lo, hi, body), n
        print(f"== {name}\n   {len(code)} instructions, {len(loops)} loops")
        if not best:
            continue
        lo, hi, body = best
        mix = collections.Counter(re.sub(r"\..*", "", b) for b in body)
        fp64 = mix["DFMA"] + mix["DMUL"] + mix["DADD"]
        print(f"   hot loop 0x{lo:x}..0x{hi:x}: {len(body)} instructions")
        print("   " + ", ".join(f"{k} {v}" for k, v in mix.most_common()))
        print(f"   FP64 pipe: DFMA {mix['DFMA']} + DMUL {mix['DMUL']} + DADD {mix['DADD']} = {fp64};  "
              f"MUFU {mix['MUFU']}, SHFL {mix['SHFL']}, DSETP {mix['DSETP']}, shared/local loads "
              f"{mix['LDS'] + mix['LDL']}, stores {mix['STS'] + mix['STL']}")
        if K:
            flops = 2 * mix["DFMA"] + mix["DMUL"] + mix["DADD"]
            print(f"   per individual-pass (K = {K} individuals per lane and pass): {fp64 / K:.2f} FP64 instructions, "
                  f"{flops / K:.2f} flop  (whole loop incl. the per-pass tail: lane reduction, odds, stop test)")


if __name__ == "__main__":
    main()
