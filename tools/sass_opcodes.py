"""Regenerates profiles/r2_sass_opcodes.md: opcode counts per kernel from `cuobjdump -sass` of the built library
(runs anywhere the CUDA toolkit is installed; no GPU needed).  Usage: python tools/sass_opcodes.py > profiles/r2_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "advise_video_ssl_b200", "libavssl_b200.so")
COLS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "UTCBAR", "MUFU.EX2", "FFMA", "FMUL", "FADD", "LDG.E.128", "STG.E.128",
        "ATOMS", "MEMBAR", "ELECT"]


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|avssl::|void ", "", name)
    name = re.sub(r"\(.*$", "", name)                 # argument list
    name = name.replace("(int)", "").replace("(bool)", "").replace("true", "true").replace("false", "false")
    return name.strip()


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_n"] += 1
            for col in COLS:
                if op == col or op.startswith(col + ".") or (col.count(".") and op.startswith(col)):
                    kernels[cur][col] += 1
    names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS opcode counts per kernel (`cuobjdump -sass libavssl_b200.so`, sm_100a, round 2; `python tools/sass_opcodes.py`)\n")
    print("Tensor-core / TMEM / TMA families: `UTCHMMA` = tcgen05.mma (kind::f16 / kind::tf32), `LDTM` / `STTM` = tcgen05.ld / st, `UTMALDG` = TMA tensor\n"
          "load, `UBLKCP` = cp.async.bulk, `SYNCS` = mbarrier operations, `UTCBAR` = tcgen05.commit, `ELECT` = elect.sync (single-lane MMA / TMA issue),\n"
          "`ATOMS` = shared-memory atomics.  `ema_multi_tensor*` must show FMUL / FADD and **no FFMA** (three separately rounded operations: bit-exact\n"
          "with the reference's mul, mul, add).\n")
    print("| kernel | SASS instr | " + " | ".join(COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    rows = sorted(zip(names, kernels.values()), key=lambda kv: (0 if "_tc_kernel" in kv[0] else 1, short(kv[0])))
    for name, c in rows:
        print("| `%s` | %d | " % (short(name), c["_n"]) + " | ".join(str(c[col]) for col in COLS) + " |")


if __name__ == "__main__":
    sys.exit(main())
