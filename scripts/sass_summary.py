"""SASS evidence per kernel of libpeaq_b200.so (runs on the CPU: cuobjdump of the built objects):
instruction mix -- FP64 (DFMA/DMUL/DADD), shared memory, shuffles, TMA bulk copies (UBLKCP) and
mbarrier operations (SYNCS), named barriers, constant-memory loads on the uniform datapath (LDCU
from bank 3, the user's __constant__ data) and DFMAs that take a uniform-register operand,
tensor-core instructions (none expected on this path) -- static counts per kernel.  usage: python scripts/sass_summary.py > profiles/<tag>_sass_summary.txt"""
import collections, glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASSES = [("DFMA", r"^DFMA"), ("DMUL", r"^DMUL"), ("DADD", r"^DADD"), ("MUFU.64H", r"^MUFU\.(RCP|RSQ)64H"),
           ("LDS", r"^LDS"), ("STS", r"^STS"), ("SHFL", r"^SHFL"), ("LDG", r"^LDG"), ("STG", r"^STG"),
           ("LDL/STL (spill)", r"^(LDL|STL)"), ("UBLKCP (TMA bulk)", r"^UBLKCP"), ("SYNCS (mbarrier)", r"^SYNCS"),
           ("LDGSTS (cp.async)", r"^LDGSTS"), ("BAR", r"^BAR"), ("CALL", r"^CALL"),
           ("tensor (HMMA/DMMA/UTC*MMA)", r"^(HMMA|DMMA|IMMA|UTC[A-Z]*MMA|TCGEN)")]

EXTRA = ["LDCU c[0x3] (uniform datapath)", "DFMA with UR operand"]


def collect():
    """{kernel name: {"object": file, "total": n, class: count, ...}} over the built kernel objects"""
    res = collections.OrderedDict()
    for obj in sorted(glob.glob(os.path.join(ROOT, "gstpeaq_b200", "csrc", "build", "*.o"))):
        out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        kernel = None
        counts = collections.OrderedDict()
        for ln in out.splitlines():
            m = re.match(r"\s*Function : (\S+)", ln)
            if m:
                kernel = m.group(1)
                counts[kernel] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
            if m and kernel:
                op = m.group(1)
                counts[kernel]["total"] += 1
                for name, pat in CLASSES:
                    if re.match(pat, op):
                        counts[kernel][name] += 1
                if op.startswith("LDCU") and "c[0x3]" in ln:
                    counts[kernel][EXTRA[0]] += 1
                if op.startswith("DFMA") and re.search(r"\bUR\d+", ln):
                    counts[kernel][EXTRA[1]] += 1
        for kernel, c in counts.items():
            if c["total"] < 40:
                continue
            dem = subprocess.run(["cu++filt", kernel], capture_output=True, text=True).stdout.strip() or kernel
            # "void peaq::<unnamed>::name<(bool)1>(params)" -> "name<1>"
            m = re.search(r"::([A-Za-z_]\w*)(<[^>]*>)?\(", dem)
            short = (m.group(1) + re.sub(r"\((?:bool|int)\)", "", m.group(2) or "")) if m else dem
            res[short] = dict(c, object=os.path.basename(obj))
    return res


def main():
    for short, c in collect().items():
        print("%s  [%s]  %d instructions (%.0f KB)" % (short, c["object"], c["total"], c["total"] * 16 / 1024.))
        names = [n for n, _ in CLASSES] + EXTRA
        print("    " + "  ".join("%s %d" % (n, c[n]) for n in names if c.get(n)))


if __name__ == "__main__":
    main()
