"""Writes profiles/sass_r2_summary.txt: opcode histograms of the kernels in build/obj/*.o (cuobjdump -sass) and the
instructions that show which hardware paths they use (bulk-copy engine, mbarriers, setmaxnreg; no tensor cores).
Run after a build: python tools/sass_summary.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = {
    "tube_wide_kernel<double> (conformance, kernels_f64)": ("kernels_f64.o", "_ZN7trm_k6416tube_wide_kernelIdEEvNS_8WideArgsE"),
    "tube_wide_kernel<double> (strict, kernels_f64s)": ("kernels_f64s.o", "_ZN8trm_k64s16tube_wide_kernelIdEEvNS_8WideArgsE"),
    "tube_wide_kernel<float> (kernels_f32)": ("kernels_f32.o", "_ZN7trm_k3216tube_wide_kernelIfEEvNS_8WideArgsE"),
    "src_kernel<double,0> (kernels_f64)": ("kernels_f64.o", "_ZN7trm_k6410src_kernelIdLi0EEEvN3trm7SrcArgsE"),
    "pcm_kernel<double> (kernels_f64)": ("kernels_f64.o", "_ZN7trm_k6410pcm_kernelIdEEvN3trm7PcmArgsE"),
}
MARKERS = ("UBLKCP", "SYNCS", "USETMAXREG", "NANOSLEEP", "MUFU.RCP64H", "HMMA", "UTCMMA", "TCGEN", "LDSM", "DFMA", "DADD", "DMUL")

out = ["SASS of the sm_100a cubins of gnuspeech_b200/lib/libtrm_cuda.so (cuobjdump -sass), round 2: opcode histograms and the",
       "instructions that show which hardware paths the kernels use.  Regenerate: python tools/sass_summary.py", ""]
for name, (obj, sym) in KERNELS.items():
    s = subprocess.run(["cuobjdump", "-sass", "-fun", sym, os.path.join(ROOT, "build", "obj", obj)], stdout=subprocess.PIPE, text=True).stdout
    ops = collections.Counter()
    for ln in s.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[0-9T]+ )?([A-Z0-9_.]+)", ln)
        if m:
            ops[m.group(1)] += 1
    fam = collections.Counter()
    for k, v in ops.items():
        fam[k.split(".")[0]] += v
    out.append("== %s: %d instructions" % (name, sum(ops.values())))
    out.append("   families: " + ", ".join("%s %d" % kv for kv in fam.most_common(24)))
    out.append("   evidence: " + ", ".join("%s x%d" % (k, ops[k]) for k in sorted(ops) if any(t in k for t in MARKERS)))
    out.append("")
out += ["Reading: UBLKCP = bulk-copy (TMA) engine moving control frames / converter windows and coefficient rows into shared",
        "memory; SYNCS.* = mbarrier operations (full / empty ring barriers, copy completion); USETMAXREG.* = setmaxnreg register",
        "redistribution (conformance mode only: donor warpgroup -> recurrence warpgroup); MUFU.RCP64H = hardware reciprocal seed",
        "of the Newton divisions; no tensor-core instructions (HMMA / UTCMMA / tcgen05) anywhere -- no stage of the path is a",
        "contraction."]
open(os.path.join(ROOT, "profiles", "sass_r2_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
