"""SASS evidence for profiles/: per kernel of the product library the opcode histogram, the memory / atomic / tensor
instructions that define the design (REDG, ATOMS, CREDUX, LDG.128, STG.128, HMMA, FADD.RM, MATCH ...), registers.
    python benchmarks/sass_excerpt.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "shacira_b200", "libshacira_b200.so")
KERNELS = [
    ("tiled forward, 2D image shape (C=1, F=1)", r"latent_fwd_tiled_kernelILi2ELi1ELi1E"),
    ("tiled backward, 2D image shape, decoder gradients", r"latent_bwd_tiled_kernelILi2ELi1ELi1ELb1E"),
    ("lane-pair forward, 3D NeRF shape (C=1, F=4)", r"latent_fwd3d_lp_kernelILi1ELi4E"),
    ("lane-pair backward, 3D NeRF shape", r"latent_bwd3d_lp_kernelILi1ELi4E"),
    ("tile-staged coarse levels of the 3D backward", r"latent_bwd_tiled_kernelILi3ELi1ELi4ELb0E"),
    ("decoder MLP + MSE on the tensor cores (mma.sync TF32)", r"mlp16_tc_step_kernelILi2ELi6ELi2E"),
    ("fused fit kernel: grid forward + MLP / MSE + grid backward (mma.sync TF32, ATOMS fixed point)", r"fit_tile_kernel"),
    ("one-launch optimizer: bit-rate loss + Adam of every group + next SGA sample", r"fit_optimizer_kernel"),
    ("peer-memory all-reduce over NVLink, 8 ranks (ld/st .sys on peer pointers, release/acquire flags)", r"peer_allreduce_kernelILi8ELb0E"),
    ("NVSwitch multicast all-reduce, 8 ranks (multimem.ld_reduce / multimem.st)", r"peer_allreduce_mc_kernelILi8E"),
    ("bit-rate kernel", r"entropy_kernel"),
    ("bit-rate kernel, validation mode through the per-integer table", r"entropy_val_lut_kernel"),
    ("SGA quantiser", r"sga_quantize_kernel"),
]
KEY = ["LDGMC", "MULTIMEM", "MEMBAR", "ERRBAR", "CCTL", "LD", "ST", "ATOM", "REDG", "RED", "ATOMS", "ATOMG", "CREDUX", "REDUX", "LDGSTS", "HMMA", "UTCHMMA", "LDTM", "UTMALDG", "MATCH", "SHFL",
       "LDS", "STS", "LDG", "STG", "F2I", "I2F", "I2FP", "F2F", "FRND", "DMUL", "DFMA", "MUFU", "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)
    for title, pat in KERNELS:
        body = next((f for f in funcs if re.match(r"\S*" + pat, f)), None)
        print("=" * 110)
        print(title)
        if body is None:
            print("  (kernel not found: %s)" % pat)
            continue
        name = body.split("\n", 1)[0].strip()
        print("  " + name)
        m = re.search(re.escape(name) + r".*?\n\s*(REG:\d+[^\n]*)", res, re.S)
        if m:
            print("  " + m.group(1).strip())
        ops = collections.Counter()
        detail = collections.Counter()
        for line in body.split("\n"):
            mm = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
            if not mm:
                continue
            op, suffix = mm.group(1), mm.group(2)
            ops[op] += 1
            if op in KEY:
                detail[op + suffix] += 1
        total = sum(ops.values())
        print("  %d SASS instructions; top opcodes: %s" % (total, ", ".join("%s %d" % kv for kv in ops.most_common(12))))
        print("  design-defining instructions:")
        for k, v in sorted(detail.items(), key=lambda kv: (-kv[1], kv[0])):
            print("    %-44s %d" % (k, v))


if __name__ == "__main__":
    main()
