"""SASS listings of the hot kernels from the built library (cuobjdump), with the control-word column stripped, plus an
opcode histogram per kernel.  Usage: python profiles/tools_sass.py <tag>   ->  profiles/<tag>_sass_<kernel>.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bosot_b200", "libbosot_b200.so")
# kernel label -> substring of the mangled name (resolved against the symbols of the built library)
KERNELS = {
    "scan_kernel": "11scan_kernelE",
    "sot_pair_kernel_R7_table_exp": "15sot_pair_kernelILi7ELb0ELb0ELi1E",   # two-launch path, table-driven exp (the bench kernel)
    "sot_pair_kernel_R2_fused_scan": "15sot_pair_kernelILi2ELb0ELb1ELi1E",  # small pools (10k x 50)
    "sot_pair_kernel_R7_f32": "15sot_pair_kernelILi7ELb0ELb0ELi2E",         # fp32 outputs
    "posterior_ws_kernel_TC64": "19posterior_ws_kernelILi64E",
}
_symbols = [l.split("Function :")[1].strip() for l in subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout.splitlines()
            if "Function :" in l]
KERNELS = {name: next(sym for sym in _symbols if part in sym) for name, part in KERNELS.items()}
tag = sys.argv[1] if len(sys.argv) > 1 else "sass"
for name, sym in KERNELS.items():
    out = subprocess.run(["cuobjdump", "-sass", "-fun", sym, LIB], capture_output=True, text=True).stdout
    lines = [re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/\s*$", "", l.rstrip()) for l in out.splitlines() if re.match(r"^\s+/\*[0-9a-f]{4}\*/", l)]
    ops = collections.Counter()
    for l in lines:
        m = re.match(r"^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_.]+)", l)
        if m:
            ops[m.group(2)] += 1
    path = os.path.join(ROOT, "profiles", "%s_sass_%s.txt" % (tag, name))
    with open(path, "w") as f:
        f.write("# %s (%s), sm_100a, %d instructions (static)\n# opcode histogram:\n" % (name, sym, len(lines)))
        for op, n in ops.most_common():
            f.write("#   %-28s %5d\n" % (op, n))
        f.write("\n".join(lines) + "\n")
    print(path, len(lines), {k: ops[k] for k in ops if any(t in k for t in ("UBLKCP", "DMMA", "VIMNMX.U16x2", "REDUX", "MATCH", "SYNCS", "ATOMS", "MUFU.EX2"))})
