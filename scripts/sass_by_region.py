"""Static instruction histogram of the fused kernel by source region (no GPU): nvdisasm --print-line-info of a compiled plan, the
instructions of pdl_main_kernel<true,false> attributed to line ranges of jet_mlp_kernel_mma.cuh (regions are found by their marker
comments).  The layer loops are rolled, so the per-region counts are per layer and tile.
usage: python scripts/sass_by_region.py plan.so"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "pde_learn_b200", "csrc", "jet_mlp_kernel_mma.cuh")).read().splitlines()
marks = [("prologue/helpers", 1)]
for i, l in enumerate(src, 1):
    m = re.search(r"// -{4,} (.*?) -{4,}", l) or re.search(r"// ---- (weight gradient on the tensor pipe|data gradient on the tensor pipe|weight gradient:|data gradient on the FP32)", l)
    if m and i > 400:
        marks.append((m.group(1)[:48], i))
so = sys.argv[1]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, check=True, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, cub)], capture_output=True, text=True).stdout.splitlines()
on = False
region = collections.Counter(); ops = collections.defaultdict(collections.Counter)
cur = None; helper = None
for l in dis:
    if l.startswith(".text."):
        on = "pdl_main_kernelILb1ELb0" in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        f, ln = m.group(1), int(m.group(2))
        if f.endswith("jet_mlp_kernel_mma.cuh"):
            if ln >= marks[1][1]:
                cur = [n for n, s in marks if s <= ln][-1]; helper = None
            else:
                helper = "helper@%d" % ln              # inlined helper: attribute to the enclosing region (last body line seen)
        else:
            helper = os.path.basename(f)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", l)
    if m and cur:
        region[cur] += 1; ops[cur][m.group(1)] += 1
tot = sum(region.values())
for n, s in marks:
    if region[n]:
        print("%5d %5.1f%%  %-50s %s" % (region[n], 100 * region[n] / tot, n, " ".join("%s:%d" % kv for kv in ops[n].most_common(8))))
print("%5d total" % tot)
