#!/usr/bin/env python
"""Extract the SASS of selected kernels from libmxb200_cuda.so into profiles/sass/ plus a mnemonic census
(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG / UTMAREDG = TMA, .IM2COL = im2col-mode TMA,
LDGMC = multimem.ld_reduce, no HMMA = no legacy mma.sync path)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mxb200", "libmxb200_cuda.so")
OUT = os.path.join(ROOT, "profiles", "sass")
WANT = {   # output name -> regex on the demangled kernel name
    "conv_fwd_im2col_bn128.sass": r"gemm_kernel<128, false, false, __nv_bfloat16, false, 1, 1>",
    "conv_fwd_shifted_taps_bn64.sass": r"gemm_kernel<64, false, false, __nv_bfloat16, false, 1, 3>",
    "conv_dgrad_im2col_2cta.sass": r"gemm_kernel<256, false, true, __nv_bfloat16, false, 2, 1>",
    "conv_wgrad_im2col_bn256.sass": r"gemm_kernel<256, true, true, float, true, 1, 2>",
    "attention_fwd.sass": r"attn::fwd_kernel",
    "attention_bwd.sass": r"attn::bwd_kernel",
    "fused_allreduce_update_nvls.sass": r"ar::fused_kernel<float, __nv_bfloat16, 8>",
}


def main():
    os.makedirs(OUT, exist_ok=True)
    txt = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    parts = re.split(r"(?m)^\s*Function : ", txt)[1:]
    names = subprocess.run(["c++filt"], input="\n".join(p.split("\n", 1)[0].strip() for p in parts), stdout=subprocess.PIPE,
                           text=True).stdout.splitlines()
    census = []
    for fname, rx in WANT.items():
        for name, body in zip(names, parts):
            if re.search(rx, name):
                with open(os.path.join(OUT, fname), "w") as f:
                    f.write("// " + name + "\n" + body)
                ops = collections.Counter(re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body, flags=re.M))
                keys = [k for k in ops if re.match(r"(UTC|LDTM|STTM|UTMA|UBLKCP|LDGMC|HMMA|SYNCS|UTCBAR|REDG|STG\.E\.128\.STRONG)", k)]
                census.append("%s  (%s)\n    %s" % (fname, name[:110], ", ".join("%s x%d" % (k, ops[k]) for k in sorted(keys))))
                break
        else:
            census.append("%s  NOT FOUND (%s)" % (fname, rx))
    with open(os.path.join(ROOT, "profiles", "sass_mnemonics_r2.txt"), "w") as f:
        f.write("\n".join(census) + "\n")
    print("\n".join(census))


if __name__ == "__main__":
    main()
