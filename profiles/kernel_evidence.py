"""Static evidence about the compiled kernels (no GPU needed):
   python profiles/kernel_evidence.py > profiles/r2_sass_summary.txt
 - per kernel instantiation: ptxas registers / spills / shared memory (recompiles engine.cu with -Xptxas -v),
 - per kernel instantiation of libstrom_b200.so: SASS opcode counts that show what the kernel is made of
   (256-bit global loads / stores, cp.async staging (LDGSTS), L2 prefetches (CCTL / UBLKPF), programmatic dependent
   launch (PREEXIT / ACQBULK), block barriers, FP64 arithmetic, shared-memory reads), and that no tensor-core or TMA
   tensor instruction is present (the path is a 0.6 flop/byte stream)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "strom_b200", "csrc")
LIB = os.path.join(CSRC, "libstrom_b200.so")
WANT = re.compile(r"prune_chains_kernelI([df])Li([145])ELi(\d)ELb(\d)|prune_quad_kernelI([df])Li([145])|root_loglik_kernelI([df])Li([45])|"
                  r"transition_matrices_kernelI([df])|tiptip_tables_kernelI([df])|fold_roots|pattern")
SHAPES = {"0": "stream", "1": "wide", "2": "narrow", "4": "wide2"}


def pretty(name):
    m = re.search(r"prune_chains_kernelI([df])Li(\d)ELi(\d)ELb(\d)", name)
    if m:
        return "prune_chains<%s,C=%s,%s%s>" % ("f64" if m.group(1) == "d" else "f32", m.group(2), SHAPES[m.group(3)], ",tips-only" if m.group(4) == "1" else "")
    m = re.search(r"prune_quad_kernelI([df])Li(\d)", name)
    if m:
        return "prune_quad<%s,C=%s>" % ("f64" if m.group(1) == "d" else "f32", m.group(2))
    m = re.search(r"root_loglik_kernelI([df])Li(\d)", name)
    if m:
        return "root_loglik<%s,C=%s>" % ("f64" if m.group(1) == "d" else "f32", m.group(2))
    m = re.search(r"(transition_matrices|tiptip_tables)_kernelI([df])", name)
    if m:
        return "%s<%s>" % (m.group(1), "f64" if m.group(2) == "d" else "f32")
    m = re.search(r"\d+([a-z_]+_kernel)", name)
    return m.group(1) if m else name[:50]


def ptxas():
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xptxas", "-v", "-c",
           "-o", "/dev/null", os.path.join(CSRC, "engine.cu"), "-ccbin", "/usr/bin/g++"]
    out = subprocess.run(cmd, capture_output=True, text=True).stderr
    cur, spill, rows = None, "", []
    for l in out.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", l)
        if m:
            cur, spill = m.group(1), ""
        s = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", l)
        if s and cur and int(s.group(2)) >= int((spill or "0/0").split("/")[0]):
            spill = "%s/%s" % (s.group(2), s.group(3))
        u = re.search(r"Used (\d+) registers, used (\d+) barriers(?:, (\d+) bytes cumulative stack size)?, (\d+) bytes smem", l)
        if u and cur and WANT.search(cur):
            rows.append((pretty(cur), u.group(1), spill or "0/0", u.group(4)))
    return rows


def sass():
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for l in out.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            cur = m.group(1) if WANT.search(m.group(1)) else None
            if cur:
                counts[cur] = collections.Counter()
            continue
        if cur:
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
            if m:
                counts[cur][m.group(1)] += 1
    return counts


def main():
    print("# ptxas (-Xptxas -v, sm_100a): registers, spill stores/loads (bytes), static shared memory (bytes)")
    print("%-44s %5s %11s %7s" % ("kernel", "regs", "spill st/ld", "smem"))
    for r in ptxas():
        print("%-44s %5s %11s %7s" % r)
    print()
    print("# SASS opcode counts per kernel of libstrom_b200.so (cuobjdump -sass)")
    cols = [("LDG.256", r"^LDG\..*256"), ("STG.256", r"^STG\..*256"), ("LDG.128", r"^LDG\..*128"), ("LDGSTS", r"^LDGSTS"), ("CCTL", r"^CCTL"),
            ("UBLKPF", r"^UBLKPF"), ("PREEXIT", r"^PREEXIT"), ("ACQBULK", r"^ACQBULK"), ("BAR", r"^BAR"), ("LDS", r"^LDS"), ("STS", r"^STS"),
            ("SHFL", r"^SHFL"), ("DFMA", r"^DFMA"), ("DMUL", r"^DMUL"), ("FFMA", r"^FFMA"), ("MUFU", r"^MUFU"), ("LDL/STL", r"^(LDL|STL)"),
            ("tensor/TMA", r"^(UTC|HMMA|IMMA|DMMA|UTMA|LDTM|STTM|HGMMA)"), ("total", r".")]
    print("%-44s " % "kernel" + " ".join("%8s" % c for c, _ in cols))
    for name, cnt in sass().items():
        row = []
        for _, pat in cols:
            row.append(sum(v for k, v in cnt.items() if re.search(pat, k)))
        print("%-44s " % pretty(name) + " ".join("%8d" % x for x in row))


if __name__ == "__main__":
    main()
