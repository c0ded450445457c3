"""Attribute the warp-stall samples / executed instructions of an `ncu --set full --import-source on` capture of pdl_main_kernel to
the regions of warp_body() in jet_mlp_kernel_mma.cuh (no GPU needed): ncu's SASS source page is joined by instruction offset with
`nvdisasm --print-line-info-inline` of the same plan module compiled here; an instruction belongs to the innermost line of its
inline chain that lies inside warp_body (helpers defined outside count for their call site; lambda bodies count as themselves).
usage: python scripts/ncu_by_region.py report.ncu-rep plan.so [--lines N]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUH = os.path.join(ROOT, "pde_learn_b200", "csrc", "jet_mlp_kernel_mma.cuh")
src = open(CUH).read().splitlines()
body0 = next(i for i, l in enumerate(src, 1) if "PDL_DEV void warp_body" in l)
body1 = next(i for i, l in enumerate(src, 1) if "// CTA tail: sum the warps" in l)
marks = [("tile loop head / misc", body0)]
for i, l in enumerate(src, 1):
    if body0 < i < body1:
        m = re.search(r"// -{4,} (.*?) -{4,}", l) or re.search(r"// ---- (weight gradient on the tensor pipe|data gradient on the tensor pipe|weight gradient: W|data gradient on the FP32|warp epilogue)", l)
        if m:
            marks.append((m.group(1)[:52], i))
# finer split of the hidden-layer adjoint: activation adjoint + recompute vs the rest
rep, so = sys.argv[1], sys.argv[2]
nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 0
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, check=True, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(td, cub)], capture_output=True, text=True).stdout.splitlines()
line_of = {}          # offset -> line in warp_body (or None)
on = False; chain = []; chain_open = False
for l in dis:
    if l.startswith(".text."):
        on = "pdl_main_kernelILb1ELb0" in l; chain = []
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        if not chain_open:                       # first location line after an instruction: a new chain (innermost first)
            chain = []
        chain_open = True
        chain.append((m.group(1), int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
    if m:
        chain_open = False
        inside = [ln for f, ln in chain if f.endswith("jet_mlp_kernel_mma.cuh") and body0 <= ln < body1]
        line_of[int(m.group(1), 16)] = inside[0] if inside else None      # innermost line inside warp_body (lambda bodies included)
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
ia, isrc, isamp, iexe = H.index("Address"), H.index("Source"), H.index("# Samples"), H.index("Instructions Executed")
stall_cols = [(c, H.index(c)) for c in ("stall_wait", "stall_selected", "stall_math", "stall_short_sb", "stall_long_sb", "stall_not_selected", "stall_dispatch", "stall_no_inst")]
base = None
reg = collections.defaultdict(lambda: collections.Counter()); per_line = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) <= iexe or not r[ia]:
        continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None:
        base = a
    ln = line_of.get(a - base)
    name = "outside warp_body (staging, CTA tail)" if ln is None else [n for n, s in marks if s <= ln][-1]
    samp, exe = int(r[isamp] or 0), int(r[iexe] or 0)
    for key in (reg[name], tot) + ((per_line[ln],) if ln else ()):
        key["samples"] += samp; key["inst"] += exe
        for c, i in stall_cols:
            key[c] += int(r[i] or 0)
        op = r[isrc].split()[0] if not r[isrc].startswith("@") else r[isrc].split()[1]
        if op.startswith("HMMA"):
            key["hmma"] += exe
print("%-54s %7s %7s %7s | %s" % ("region", "samp%", "inst%", "hmma%", " ".join(c.replace("stall_", "")[:7].rjust(7) for c, _ in stall_cols)))
for name in [n for n, _ in marks] + ["outside warp_body (staging, CTA tail)"]:
    k = reg.get(name)
    if not k or not k["samples"]:
        continue
    print("%-54s %6.1f%% %6.1f%% %6.1f%% | %s" % (name, 100 * k["samples"] / tot["samples"], 100 * k["inst"] / tot["inst"], 100 * k["hmma"] / max(tot["hmma"], 1),
                                            " ".join(("%5.1f%%" % (100 * k[c] / max(k["samples"], 1))).rjust(7) for c, _ in stall_cols)))
print("total samples %d, warp instructions %d, HMMA %d" % (tot["samples"], tot["inst"], tot["hmma"]))
if nlines:
    print("\nhottest lines of warp_body:")
    for ln, k in sorted(per_line.items(), key=lambda kv: -kv[1]["samples"])[:nlines]:
        print("%5d %5.1f%% inst %5.1f%%  %s" % (ln, 100 * k["samples"] / tot["samples"], 100 * k["inst"] / tot["inst"], src[ln - 1].strip()[:110]))
