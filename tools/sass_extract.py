"""SASS evidence for the Blackwell-native kernels: per kernel of the built library, the counts of the tensor-core / TMEM /
TMA mnemonics (B200_PROFILING.md "What proves a Blackwell-native kernel") plus the legacy ones that must be absent,
written to profiles/<prefix>_sass_summary.txt, and the SASS of the hot kernels' inner loops to
profiles/<prefix>_sass_<kernel>.txt (the lines carrying those mnemonics with their addresses).
    python tools/sass_extract.py [r02]          # runs cuobjdump here, no GPU needed"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "regione_b200", "_lib", "libregione_b200.so")
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCCP", "SYNCS",
         "FFMA2", "FADD2", "FMNMX3", "MUFU.EX2", "MUFU.TANH", "HMMA", "HGMMA", "LDGSTS"]


def main():
    prefix = sys.argv[1] if len(sys.argv) > 1 else "r02"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            kernels[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            kernels[name].append(line.rstrip())
    out_dir = os.path.join(ROOT, "profiles")
    with open(os.path.join(out_dir, f"{prefix}_sass_summary.txt"), "w") as f:
        f.write(f"# cuobjdump -sass regione_b200/_lib/libregione_b200.so (sm_100a), mnemonic counts per kernel\n")
        f.write("# UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, HMMA/HGMMA = legacy "
                "tensor paths (must be 0)\n")
        for k, lines in kernels.items():
            ops = collections.Counter()
            stg256 = sum(1 for ln in lines if re.search(r"\bSTG\.[A-Z0-9.]*256\b", ln))
            for ln in lines:
                body = ln.split("*/", 1)[1] if "*/" in ln else ln
                mm = re.search(r"\b([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", body.replace("@P", " ").replace("@!P", " "))
                if not mm:
                    continue
                op = mm.group(1)
                for w in WATCH:
                    if op == w or op.startswith(w + ".") or (w == "UTCHMMA.2CTA" and op.startswith("UTCHMMA") and
                                                             ".2CTA" in op):
                        ops[w] += 1
            if not any(ops[w] for w in ("UTCHMMA", "LDTM", "UTMALDG", "FFMA2")):
                continue
            short = k.split("(")[0]
            f.write(f"{short}: {len(lines)} instructions; " + ", ".join(f"{w} {ops[w]}" for w in WATCH if ops[w]) +
                    (f", STG.256 {stg256}" if stg256 else "") + f"; HMMA {ops['HMMA']}, HGMMA {ops['HGMMA']}\n")
            if any(t in short for t in ("attention_kernel<3>", "attention_kernel<0>", "gemm2_kernel<3>", "gemm2_kernel<2>",
                                        "gemm_kernel<0>")):
                fn = re.sub(r"[^A-Za-z0-9]+", "_", short).strip("_")
                with open(os.path.join(out_dir, f"{prefix}_sass_{fn}.txt"), "w") as g:
                    g.write(f"# {k}\n# lines of the SASS carrying tensor-core / TMEM / TMA / packed-math mnemonics\n")
                    for ln in lines:
                        if any(w.split(".")[0] in ln for w in WATCH):
                            g.write(ln + "\n")
    print(open(os.path.join(out_dir, f"{prefix}_sass_summary.txt")).read())


if __name__ == "__main__":
    main()
