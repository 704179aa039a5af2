#!/usr/bin/env python3
"""Static look at a kernel's hot loop before spending GPU time on it: opcode histogram and the sum of the
stall counts ptxas encoded (= the fewest cycles ONE warp needs per trip, nothing else on the scheduler).

    python scripts/sass_loop_stats.py <object-or-cubin> <substring of the mangled kernel name>

The hot loop is taken to be the backward conditional branch whose body holds the most FP32 arithmetic (loops that call
out of line or divide the slow way -- cold repair paths -- are skipped).  Control words are decoded
from the second 64-bit word cuobjdump prints for every instruction (stall count: bits 41..44)."""
import collections
import re
import subprocess
import sys


def kernel_sass(obj: str, pattern: str):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout.split("\n")
    keep, on = [], False
    for line in out:
        if "Function :" in line:
            if on:
                break
            on = pattern in line
            if on:
                keep.append(line.strip())
            continue
        if on:
            keep.append(line)
    return keep


def instructions(lines):
    ins, i = [], 0
    while i < len(lines) - 1:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", lines[i])
        if m:
            m2 = re.search(r"/\* (0x[0-9a-f]+) \*/", lines[i + 1])
            hi = int(m2.group(1), 16) if m2 else 0
            ins.append((int(m.group(1), 16), m.group(2), (hi >> 41) & 0xF))
            i += 2
        else:
            i += 1
    return ins


def opcode(text: str) -> str:
    return re.sub(r"^(@!?U?P[0-9T]\s+)?", "", text).split()[0].split(".")[0]


def main():
    obj, pattern = sys.argv[1], sys.argv[2]
    lines = kernel_sass(obj, pattern)
    if not lines:
        sys.exit(f"no kernel matching {pattern!r} in {obj}")
    ins = instructions(lines[1:])
    # candidate loops = backward conditional branches; the hot loop is the one with the most FP32 arithmetic that does
    # not call out of line (a kernel may also carry a cold repair loop with IEEE divisions, which is longer)
    best, best_score = None, -1
    for addr, text, _ in ins:
        m = re.match(r"^@!?U?P\d\s+BRA(?:\.U)?\s+(?:!?UP\d,\s*)?0x([0-9a-f]+)", text)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr:
                body = [opcode(t) for a, t, _ in ins if tgt <= a <= addr]
                if "CALL" in body or "MUFU" in body:
                    continue
                score = sum(body.count(k) for k in ("FADD2", "FMUL2", "FFMA2", "FADD", "FMUL", "FFMA"))
                if score > best_score:
                    best, best_score = (tgt, addr), score
    if best is None:
        sys.exit("no backward conditional branch found")
    loop = [x for x in ins if best[0] <= x[0] <= best[1]]
    hist = collections.Counter(opcode(t) for _, t, _ in loop)
    stalls = collections.Counter()
    for _, t, s in loop:
        stalls[opcode(t)] += s
    print(lines[0])
    print(f"hot loop {best[0]:#x}..{best[1]:#x}: {len(loop)} instructions, encoded stall cycles {sum(s for *_, s in loop)}")
    fp2 = sum(hist[k] for k in ("FADD2", "FMUL2", "FFMA2"))
    fp1 = sum(hist[k] for k in ("FADD", "FMUL", "FFMA"))
    print(f"packed FP32: {fp2}  scalar FP32: {fp1}  FMA-pipe cycles if a packed op holds the pipe for two: {2 * fp2 + fp1 + hist['IMAD']}")
    for op, n in hist.most_common():
        print(f"  {op:12s} {n:5d}   stall cycles {stalls[op]:5d}")


if __name__ == "__main__":
    main()
