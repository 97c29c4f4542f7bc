#!/usr/bin/env python
"""Executed-instruction histogram by SASS opcode (and by pipe class) of one kernel from an ncu report with source info.
  python tools/ncu_sass_hist.py REPORT.ncu-rep KERNEL_REGEX [TOP]
"""
import csv, io, subprocess, sys, collections

ALU = ("LOP3", "SHF", "PRMT", "VABSDIFF", "VIMNMX", "VIADD", "ISETP", "SEL", "IADD3", "IADD", "LEA", "IMNMX", "FSETP", "PLOP3", "FMNMX", "BMSK", "SGXT", "VHMNMX", "FLO", "IABS", "ICMP")
FMA = ("IMAD", "IDP", "FFMA", "FMUL", "FADD", "HFMA", "FMA")
def cls(op):
    b = op.split(".")[0]
    if b in ("LDS", "STS", "LDG", "STG", "LD", "ST", "ATOMS", "ATOMG", "RED", "LDC", "LDSM", "STSM", "LDGSTS", "ATOM", "LDL", "STL"): return "LSU"
    if b in ("POPC", "MUFU", "I2F", "F2I", "F2F", "I2I", "BREV"): return "XU"
    if b in ("VOTE", "SHFL", "BAR", "BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "NANOSLEEP", "CALL", "RET", "S2R", "CS2R", "R2UR", "S2UR", "SYNCS", "UTMALDG", "ELECT", "MATCH", "DEPBAR", "NOP", "ERRBAR", "MEMBAR", "FENCE", "ACQBULK", "BMOV", "R2P", "P2R", "UBLKCP", "REDUX"): return "CTRL/MISC"
    if b.startswith("U"): return "UNIFORM"
    if any(b.startswith(a) for a in FMA): return "FMA"
    if any(b.startswith(a) for a in ALU) or b in ("MOV",): return "ALU"
    return "OTHER:" + b
def main(path, regex, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "sass", "--csv", "-k", "regex:" + regex], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None; h = collections.Counter(); c = collections.Counter(); tot = 0
    for r in rows:
        if len(r) > 8 and r[0] == "Address":
            hdr = r; ie = r.index("Instructions Executed"); continue
        if hdr and len(r) > ie:
            toks = r[1].split()
            if not toks: continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            op = op.rstrip(";")
            try: n = int(r[ie])
            except ValueError: continue
            base = ".".join(op.split(".")[:2]) if op.split(".")[0] in ("LDS", "STS", "LDG", "STG") else op.split(".")[0]
            h[base] += n; c[cls(op)] += n; tot += n
    print("warp instructions executed: %d" % tot)
    print("| class | share |\n|---|---|")
    for k, v in c.most_common(): print("| %s | %.1f%% |" % (k, 100 * v / tot))
    print("\n| opcode | share |\n|---|---|")
    for k, v in h.most_common(top): print("| %s | %.1f%% |" % (k, 100 * v / tot))
if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
