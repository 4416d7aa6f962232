#!/usr/bin/env python3
"""Static issue-slot budget of the mode loop of gsf_sum_kernel<D,NC,P,1>, from SASS.

Usage:  python tools/sass_budget.py <binary or .so> [D NC P [DEG]]  (default 3 1 3 5)
        python tools/sass_budget.py <binary or .so> --kernel <mangled-name substring>
                                                    (instruction mix of any kernel's hottest loop)

Finds the hottest innermost loop of the kernel (the one holding the most FP64 instructions)
and prices its body with the issue model measured on B200 (tools/micro/dfma_patterns.cu):

  * an FP64 instruction (DFMA / DADD / DMUL) holds the SMSP issue port for 2 cycles, everything
    else for 1;
  * an FP64 instruction that reads THREE distinct general registers from the register file costs
    a third cycle.  A source is free when it is a constant / uniform / immediate operand, repeats
    another source of the same instruction, or is served by the operand-reuse cache (the previous
    FP64 instruction carried the same register in the same slot with `.reuse`).

Prints the instruction mix, the penalty count and the predicted FP64-pipe utilisation
(2 * #FP64 / cycles) -- a way to compare schedule variants without a GPU.  The model reproduces
the ncu figure of the shipped kernel within about two points (87.8 % measured at C2)."""
import re
import subprocess
import sys


_DUMPS = {}


def sass_dump(path):
    """`cuobjdump -sass` of a binary / shared library (cached per path)"""
    if path not in _DUMPS:
        _DUMPS[path] = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    return _DUMPS[path]


def function_sass(path, mangled_substring):
    """[(address, instruction text)] of the first function whose mangled name contains the substring"""
    lines, on, seen = [], False, False
    for ln in sass_dump(path).splitlines():
        if "Function :" in ln:
            if seen and on:
                break
            on = mangled_substring in ln
            seen = seen or on
            continue
        if on:
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                lines.append((int(m.group(1), 16), m.group(2).strip()))
    return lines


def kernel_sass(path, d, nc, p, deg=5):
    """SASS of gsf_sum_kernel<D, NC, P, L=1, DEG> (DEG: polynomial degree, 5 = throughput, 6 = high)"""
    out = sass_dump(path)
    name = "_ZN3gsf14gsf_sum_kernelILi%dELi%dELi%dELi1ELi%dEEEvNS_7SumArgsE" % (d, nc, p, deg)
    lines, on = [], False
    for ln in out.splitlines():
        if "Function :" in ln:
            on = name in ln
            continue
        if on:
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                lines.append((int(m.group(1), 16), m.group(2).strip()))
    if not lines:
        raise SystemExit("kernel %s not found in %s" % (name, path))
    return lines


FP64 = ("DFMA", "DADD", "DMUL")


def opcode(text):
    t = text.split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]


def hottest_loop(lines):
    addr_index = {a: i for i, (a, _) in enumerate(lines)}
    best = None
    for i, (a, text) in enumerate(lines):
        if opcode(text) != "BRA":
            continue
        m = re.search(r"0x([0-9a-f]+)", text)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a or tgt not in addr_index:
            continue
        body = lines[addr_index[tgt]:i + 1]
        # innermost loops only: no other backward branch inside the body
        inner = False
        for a2, t2 in body[:-1]:
            if opcode(t2) == "BRA":
                m2 = re.search(r"0x([0-9a-f]+)", t2)
                if m2 and int(m2.group(1), 16) <= a2:
                    inner = True
        if inner:
            continue
        n64 = sum(opcode(t) in FP64 for _, t in body)
        if best is None or n64 > best[0]:
            best = (n64, body)
    return best[1]


def sources(text):
    """register sources of an FP64 instruction as (name, reuse) per slot, None for non-register operands"""
    ops = text.split(None, 1)[1] if not text.startswith("@") else text.split(None, 2)[2]
    parts = [x.strip() for x in ops.split(",")][1:]   # drop the destination
    res = []
    for x in parts:
        m = re.match(r"^-?\|?(R\d+)\|?(\.reuse)?$", x)
        res.append((m.group(1), bool(m.group(2))) if m and m.group(1) != "RZ" else None)
    return res


def analyze(path, d=3, nc=1, p=3, deg=5):
    """issue budget of the mode loop: dict(mix, fp64, other, three_register, cycles, pipe_frac, body_len)"""
    body = hottest_loop(kernel_sass(path, d, nc, p, deg))
    mix = {}
    prev = None
    penalties = 0
    for _, text in body:
        op = opcode(text)
        mix[op] = mix.get(op, 0) + 1
        if op in FP64:
            src = sources(text)
            fresh = set()
            for slot, s in enumerate(src):
                if s is None:
                    continue
                cached = prev is not None and slot < len(prev) and prev[slot] is not None \
                    and prev[slot][0] == s[0] and prev[slot][1]
                if not cached:
                    fresh.add(s[0])
            if len(fresh) >= 3:
                penalties += 1
            prev = src
    n64 = sum(v for k, v in mix.items() if k in FP64)
    other = sum(v for k, v in mix.items() if k not in FP64)
    cycles = 2 * n64 + other + penalties
    return {"mix": mix, "fp64": n64, "other": other, "three_register": penalties, "cycles": cycles,
            "pipe_frac": 2.0 * n64 / cycles, "body_len": len(body)}


def loop_mix(path, mangled_substring, hot=("DMMA",) + FP64):
    """instruction mix of the innermost loop of any kernel that holds the most `hot` instructions"""
    lines = function_sass(path, mangled_substring)
    if not lines:
        raise SystemExit("no function matching %s in %s" % (mangled_substring, path))
    addr_index = {a: i for i, (a, _) in enumerate(lines)}
    best = None
    for i, (a, text) in enumerate(lines):
        m = re.search(r"0x([0-9a-f]+)", text) if opcode(text) == "BRA" else None
        if not m or int(m.group(1), 16) >= a or int(m.group(1), 16) not in addr_index:
            continue
        body = lines[addr_index[int(m.group(1), 16)]:i + 1]
        # innermost, except that short spin loops (mbarrier / cp.async waits) may sit inside
        nested = False
        for a2, t2 in body[:-1]:
            m2 = re.search(r"0x([0-9a-f]+)", t2) if opcode(t2) == "BRA" else None
            if m2 and int(m2.group(1), 16) <= a2 and a2 - int(m2.group(1), 16) > 8 * 16:
                nested = True
        if nested:
            continue
        n_hot = sum(opcode(t) in hot for _, t in body)
        if best is None or n_hot > best[0]:
            best = (n_hot, body)
    mix = {}
    for _, t in best[1]:
        mix[opcode(t)] = mix.get(opcode(t), 0) + 1
    return mix


def main():
    path = sys.argv[1]
    if len(sys.argv) >= 4 and sys.argv[2] == "--kernel":   # any kernel: just the mix of its hottest loop
        mix = loop_mix(path, sys.argv[3])
        print("%s  %s  hottest innermost loop: %d instructions" % (path, sys.argv[3], sum(mix.values())))
        print("  mix: " + ", ".join("%s %d" % kv for kv in sorted(mix.items(), key=lambda kv: -kv[1])))
        return
    d, nc, p = (int(x) for x in sys.argv[2:5]) if len(sys.argv) >= 5 else (3, 1, 3)
    deg = int(sys.argv[5]) if len(sys.argv) >= 6 else 5
    r = analyze(path, d, nc, p, deg)
    print("%s  <D=%d NC=%d P=%d DEG=%d>  loop body: %d instructions" % (path, d, nc, p, deg, r["body_len"]))
    print("  mix: " + ", ".join("%s %d" % kv for kv in sorted(r["mix"].items(), key=lambda kv: -kv[1])))
    print("  FP64 %d (x2 cycles)  other %d  three-register FP64 %d  => %d cycles, FP64 pipe %.1f %%"
          % (r["fp64"], r["other"], r["three_register"], r["cycles"], 100.0 * r["pipe_frac"]))


if __name__ == "__main__":
    main()
