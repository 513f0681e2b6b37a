#!/usr/bin/env python
"""Count the SASS instructions per DP cell of the shipped forward kernels.

    python tools/sass_count.py            # writes profiles/sass_counts.json

The steady-state ("FAST") block of k_forward<V,8,false> is the straight-line region that ends
in the block's single STG.E.128 and contains exactly STEPS x 2 hand-off shuffles (SHFL.UP);
of the two such regions in each kernel (SLOW and FAST) the shorter one is the hot loop body.
One block advances STEPS lane-steps = STEPS x K cells x NPAIR alignments per lane.

Pipe classes follow the measured B200/B300 pipe map (guides/B300_MICROARCH.md): IMAD* issue on
the FMA pipe, the DPX / logic / select / add instructions on the ALU pipe, each at 16 lanes per
SMSP per clock; loads, stores and shuffles go to the LSU.
"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "micall-lite_b200", "lib", "libgotoh_b200.so")

ALU = ("VIADDMNMX", "VIMNMX", "VIMNMX3", "VIADD", "LOP3", "SEL", "IADD3", "SHF", "PRMT", "ISETP", "LEA", "MOV",
       "IABS", "PLOP3", "SGXT", "BMSK", "FLO", "POPC", "P2R", "R2P", "FSEL", "CS2R", "IADD")
FMA = ("IMAD", "FFMA", "FMUL", "FADD")
LSU = ("LDS", "STS", "LDG", "STG", "SHFL", "LD", "ST", "LDC", "ATOMG", "REDG", "LDL", "STL")
CTL = ("BRA", "BSSY", "BSYNC", "WARPSYNC", "BAR", "EXIT", "NOP", "CALL", "RET", "YIELD", "BREAK", "BMOV")


def sass_of(pattern):
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append(m.group(1).strip())
    for name, ins in funcs.items():
        if pattern in name:
            return name, ins
    raise SystemExit("kernel %s not found in %s" % (pattern, LIB))


def opcode(ins):
    tok = ins.split()
    op = tok[1] if tok[0].startswith("@") else tok[0]
    return op.split(".")[0]


def classify(op):
    if op in FMA:
        return "fma"
    if op in ALU:
        return "alu"
    if op in LSU:
        return "lsu"
    if op in CTL:
        return "ctl"
    return "other"


def hot_block(ins, steps, shfl="SHFL.UP", per_step=2):
    """Shortest region ending at an STG.E.128 and holding exactly per_step*steps hand-off shuffles."""
    best = None
    for idx, i in enumerate(ins):
        if not (opcode(i) == "STG" and ".128" in i):
            continue
        n, j = 0, idx
        while j >= 0 and n < per_step * steps:
            if ins[j].replace("@!P0 ", "").startswith(shfl) or (" " + shfl) in ins[j]:
                n += 1
            j -= 1
        if n < per_step * steps:
            continue
        # extend backwards over the instructions scheduled before the first shuffle of the block
        # until the previous control-flow instruction (the loop head / branch target)
        while j >= 0 and classify(opcode(ins[j])) != "ctl":
            j -= 1
        region = ins[j + 1:idx + 1]
        if best is None or len(region) < len(best):
            best = region
    return best


def count(kernel_pat, steps, k, npair, shfl="SHFL.UP", per_step=2):
    name, ins = sass_of(kernel_pat)
    region = hot_block(ins, steps, shfl, per_step)
    cells = steps * k * npair
    hist, pipes = {}, {"alu": 0, "fma": 0, "lsu": 0, "ctl": 0, "other": 0}
    for i in region:
        op = opcode(i)
        hist[op] = hist.get(op, 0) + 1
        pipes[classify(op)] += 1
    total = len(region)
    return {"kernel": name, "block_instructions": total, "cells_per_block": cells, "histogram": hist, "pipes": pipes,
            "instr_per_cell": total / cells, "alu_per_cell": pipes["alu"] / cells, "fma_per_cell": pipes["fma"] / cells,
            "lsu_per_cell": pipes["lsu"] / cells}


def main():
    x2 = count("k_forwardINS_5Vec16ELi8ELb0ELb0E", 4, 8, 2)
    x1 = count("k_forwardINS_5Vec32ELi8ELb0ELb0E", 8, 8, 1)
    doc = {"_how": "tools/sass_count.py over cuobjdump -sass of micall-lite_b200/lib/libgotoh_b200.so (steady-state block)",
           "x2": x2, "x1": x1,
           "instr_per_cell_x2": x2["instr_per_cell"], "alu_per_cell_x2": x2["alu_per_cell"], "fma_per_cell_x2": x2["fma_per_cell"],
           "instr_per_cell_x1": x1["instr_per_cell"], "alu_per_cell_x1": x1["alu_per_cell"], "fma_per_cell_x1": x1["fma_per_cell"]}
    # half-warp wavefronts (queries <= 128 wide; C3 runs K = 6)
    try:
        doc["x2_half_k6"] = count("k_forwardINS_5Vec16ELi6ELb0ELb1E", 4, 6, 2)
    except Exception as e:
        doc["x2_half_k6"] = {"error": str(e)}
    # long pairs (C4): strip dataflow kernel, int32 cells, 8 lane-steps x 8 columns per block
    try:
        doc["x1_flow"] = count("k_forward_flowILi8E", 8, 8, 1)
    except Exception as e:
        doc["x1_flow"] = {"error": str(e)}
    # gotoh2 (live aligner) kernels: forward with tie bits, score-only forward, reverse sweep (4 lane-steps x 8 columns)
    for tag, pat, sh, per, npair in (("g2_forward", "k2fILi8ELb0ELb1", "SHFL.UP", 2, 1), ("g2_forward_score_only", "k2fILi8ELb0ELb0", "SHFL.UP", 2, 1),
                                     ("g2_forward_x2", "k2f_x2ILi8E", "SHFL.UP", 2, 2), ("g2_reverse", "k2rILi8ELb0", "SHFL.DOWN", 1, 1),
                                     ("g2_reverse_k3", "k2rILi3ELb0", "SHFL.DOWN", 1, 1), ("g2_reverse_x2_k3", "k2r_x2ILi3E", "SHFL.DOWN", 1, 2)):
        try:
            doc[tag] = count(pat, 4, 3 if tag.endswith("k3") else 8, npair, sh, per)
        except Exception as e:      # a kernel without a 128-bit store in its loop (score-only) has no such block
            doc[tag] = {"error": "no steady-state block with a 128-bit store found (%s)" % e}
    prev = os.path.join(ROOT, "profiles", "sass_counts.json")
    if os.path.exists(prev):
        try:
            old = json.load(open(prev))
            for key in ("ncu_dram_bytes_per_launch", "ncu_dram_bytes_per_cell", "ncu_note", "ncu"):
                if key in old:
                    doc[key] = old[key]
        except Exception:
            pass
    os.makedirs(os.path.dirname(prev), exist_ok=True)
    json.dump(doc, open(prev, "w"), indent=1)
    for tag in ("x2", "x1", "x2_half_k6", "x1_flow", "g2_forward", "g2_forward_score_only", "g2_forward_x2", "g2_reverse", "g2_reverse_k3", "g2_reverse_x2_k3"):
        d = doc[tag]
        if "error" in d:
            print(tag, d["error"])
            continue
        print("%s: %d instr / %d cells = %.2f per cell (alu %.2f, fma %.2f, lsu %.2f)" % (
            tag, d["block_instructions"], d["cells_per_block"], d["instr_per_cell"], d["alu_per_cell"], d["fma_per_cell"], d["lsu_per_cell"]))
        print("   ", sorted(d["histogram"].items(), key=lambda kv: -kv[1]))


if __name__ == "__main__":
    sys.exit(main())
