"""Opcode histogram of the Blackwell-specific instructions in lib/libyolox_b200.so, per kernel family (cuobjdump -sass; runs
without a GPU).  Writes profiles/r02_sass_histogram.md.  The sparse MMA has NO mnemonic of its own in SASS: tcgen05.mma.sp is the
same UTCHMMA with the sparse bit set in the (runtime) instruction descriptor and the metadata column as an extra tmem operand,
so the PTX spelling is shown from a PTX dump of csrc/yx_conv.cu as well."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "coco-dataset-based-light-weight-fast-object-detection-model_b200")
LIB = os.path.join(PKG, "lib", "libyolox_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "LDTM", "STTM", "UTCCP", "SYNCS", "ELECT"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
fam_of = lambda name: re.sub(r"<.*", "", name.split("(")[0]).replace("yx::", "")
counts = collections.defaultdict(collections.Counter)
inst = collections.Counter()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        dem = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = fam_of(dem)
        if "conv_gemm_kernel" in dem:
            cur = "conv_gemm_kernel<...,SP=true>" if dem.rstrip(">)").split(",")[-1].strip().startswith("true") and dem.count(",") >= 4 else "conv_gemm_kernel"
        inst[cur] += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        for key in OPS:
            if op == key or op.startswith(key + "."):
                k = "UTCHMMA.2CTA" if op.startswith("UTCHMMA.2CTA") else ("UTCHMMA" if op.startswith("UTCHMMA") else key)
                counts[cur][k] += 1
                break
cols = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "SYNCS", "ELECT"]
out = ["# SASS opcode histogram of lib/libyolox_b200.so (sm_100a), per kernel family", "",
       "`cuobjdump -sass`, counted over every instantiation of a family (second column).  UTCHMMA = tcgen05.mma (`.2CTA` = cta_group::2),",
       "UTCBAR = tcgen05.commit, UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce-add, LDTM / STTM = tcgen05.ld / st,",
       "SYNCS = mbarrier operations, ELECT = elect.sync.  No HMMA / HGMMA (legacy tensor paths) anywhere in the library.", "",
       "| kernel family | instantiations | " + " | ".join(cols) + " |", "|---|---|" + "---|" * len(cols)]
for fam in sorted(counts, key=lambda f: -sum(counts[f].values())):
    out.append(f"| `{fam}` | {inst[fam]} | " + " | ".join(str(counts[fam][c]) for c in cols) + " |")
legacy = len(re.findall(r"\bHMMA\b|\bHGMMA\b|\bIMMA\b", sass))
out += ["", f"Legacy tensor-core opcodes (HMMA / HGMMA / IMMA) in the whole library: {legacy}.", ""]
# the PTX spelling of the sparse MMA (SASS folds it into UTCHMMA + idesc bit 2)
ptx = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=compute_100a", "-std=c++17", "-DYX_CONV_ACT_SLICE=2", "-diag-suppress", "177",
                      "-ptx", os.path.join(PKG, "csrc", "yx_conv.cu"), "-o", "/dev/stdout"], capture_output=True, text=True).stdout
n_sp = len(re.findall(r"tcgen05\.mma\.sp\.cta_group::1\.kind::f16", ptx))
n_mma = len(re.findall(r"tcgen05\.mma\.cta_group::[12]\.kind::f16", ptx))
out += ["## PTX view of one activation slice of csrc/yx_conv.cu (`nvcc -ptx -DYX_CONV_ACT_SLICE=2`)", "",
        f"`tcgen05.mma.sp.cta_group::1.kind::f16` (2:4 sparse A operand + tensor-memory metadata): {n_sp} sites; "
        f"`tcgen05.mma.cta_group::{{1,2}}.kind::f16`: {n_mma} sites; `tcgen05.st` (metadata -> TMEM): "
        f"{len(re.findall(r'tcgen05.st.sync', ptx))} sites.", ""]
path = os.path.join(ROOT, "profiles", "r02_sass_histogram.md")
open(path, "w").write("\n".join(out) + "\n")
print("\n".join(out[6:]))
