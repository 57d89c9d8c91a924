"""Registers, spills and static shared memory per kernel from the ptxas -v logs of the last build (quip_b200/csrc/build/*.ptxas.log).
    python tools/ptxas_summary.py > profiles/rNN_ptxas_summary.txt"""
import glob
import os
import re
import subprocess

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
print("# ptxas -v per kernel (sm_100a), from quip_b200/csrc/build/*.ptxas.log: registers, spill stores/loads (bytes), static shared memory")
for f in sorted(glob.glob(os.path.join(root, "quip_b200", "csrc", "build", "*.ptxas.log"))):
    L = open(f).read().splitlines()
    for i, l in enumerate(L):
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", l)
        if not m:
            continue
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"gapb200::|\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)
        if "cub::" in name:
            name = re.sub(r"<.*", "<...>", name)
        blob = " ".join(L[i + 1:i + 4])
        regs = re.search(r"Used (\d+) registers", blob)
        sp = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", blob)
        sm = re.search(r"(\d+) bytes smem", blob)
        print("%-18s regs %3s  spill %4s/%-4s  smem %6s  %s" % (os.path.basename(f).replace(".ptxas.log", ""), regs.group(1) if regs else "?",
                                                                 sp.group(1) if sp else "?", sp.group(2) if sp else "?", sm.group(1) if sm else "0", name[:90]))
