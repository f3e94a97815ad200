#!/usr/bin/env python
"""profiles/sass_opcodes.txt: per-kernel SASS opcode histogram of libhfg_b200.so (cuobjdump -sass) -- the evidence that the hot
kernels are tcgen05 / TMA code (UTCHMMA, UTCBAR, LDTM, UTMALDG, UTMASTG) and that no legacy mma.sync (HMMA) is present.

    python tools/sass_histogram.py
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "iris_tts_b200", "lib", "libhfg_b200.so")
WANT = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "UTCATOMSWS", "UTMACCTL", "SYNCS", "ELECT", "F2FP", "HADD2.F32",
        "FADD2", "FMUL2", "FFMA2", "ACQBULK", "HMMA", "FFMA"]
KERNELS = ("conv_umma2_kernel|conv_pair_kernel|conv_umma_kernel|logmel_kernel|mrf_combine_kernel|conv_post_mrf_kernel|"
           "mel_to_cl_bf16_kernel|planes_to_raw_kernel|conv_cl_fp32_kernel|conv_post_kernel|transpose_kernel|accum_fp32_kernel|act_split_kernel")


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    out = ["# SASS opcode histogram of iris_tts_b200/lib/libhfg_b200.so (cuobjdump -sass, sm_100a), per kernel instantiation.",
           "# UTCHMMA = tcgen05.mma (kind::f16; loops are rolled, so the count is static code, not issued MMAs), UTCBAR = tcgen05.commit,",
           "# LDTM = tcgen05.ld (TMEM -> registers), UTMALDG / UTMASTG = cp.async.bulk.tensor load / store (TMA), SYNCS = mbarrier ops,",
           "# UTCATOMSWS = tcgen05.alloc/dealloc, F2FP = packed fp32 -> bf16/fp16, HADD2.F32 = fp16 -> fp32, FADD2/FMUL2 = packed fp32 pairs,",
           "# HMMA = legacy mma.sync (expected: none).  Template arguments: conv_umma2<planes, residual, fp16, ragged>, conv_pair<planes, fp16, ragged>.", ""]
    tot = collections.Counter()
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        m = re.search(KERNELS, name)
        short = m.group(0) if m else name[:60]
        targs = []
        if m:
            tail = name[m.end():]
            mt = re.match(r"I((?:L[ib]\d+E)+)E", tail)
            if mt:
                targs = re.findall(r"L[ib](\d+)E", mt.group(1))
        tag = short + (f"<{', '.join(targs)}>" if targs else "")
        c = collections.Counter()
        n = 0
        for line in f.split("\n"):
            mm = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not mm:
                continue
            n += 1
            for w in WANT:
                if mm.group(1).startswith(w):
                    c[w] += 1
                    break
        tot.update(c)
        out.append(f"{tag:40s} {n:6d} instr  " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())))
    out += ["", "TOTAL  " + "  ".join(f"{k}={v}" for k, v in sorted(tot.items())), f"HMMA (legacy mma.sync) instructions: {tot.get('HMMA', 0)}"]
    path = os.path.join(ROOT, "profiles", "sass_opcodes.txt")
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")
    print("\n".join(out[-14:]))


if __name__ == "__main__":
    main()
