"""Static instruction mix of an address range of a cuobjdump -sass listing (the hot loop of a kernel).
usage: sass_loop_stats.py file.sass 0c00 19a0 [2500 2530 ...]   (pairs of inclusive hex addresses)"""
import collections, re, sys
ALU = {"LOP3","FSETP","FSEL","FMNMX","FMNMX3","SEL","ISETP","PLOP3","SHF","LEA","IADD3","VIADD","I2FP","POPC","PRMT","MOV","VOTE","VABSDIFF","IABS","FLO","BREV","SGXT","BMSK","P2R","R2P","CS2R","S2R"}
FMA = {"FADD","FMUL","FFMA","IMAD","HFMA2","FADD2","FMUL2","FFMA2"}
XU = {"MUFU","F2I","I2F","FRND","F2F","F2FP"}
def main():
    path = sys.argv[1]
    rng = [(int(sys.argv[i],16), int(sys.argv[i+1],16)) for i in range(2, len(sys.argv), 2)]
    ops = collections.Counter(); pipe = collections.Counter(); n = 0; heavy = 0
    for l in open(path):
        m = re.search(r"/\*([0-9a-f]{4})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
        if not m: continue
        a = int(m.group(1),16)
        if not any(lo <= a <= hi for lo,hi in rng): continue
        full = m.group(3); op = full.split(".")[0]
        ops[op if op != "IMAD" else ("IMAD.WIDE" if ".WIDE" in full else "IMAD.MOV" if ".MOV" in full else "IMAD")] += 1
        n += 1
        if op in ALU: pipe["alu"] += 1
        elif op in FMA:
            pipe["fma"] += 1
            if ".WIDE" in full or full.startswith("IMAD.HI"): heavy += 1
        elif op in XU: pipe["xu"] += 1
        elif op.startswith("U"): pipe["uniform"] += 1
        elif op in ("LDG","LDS","STS","STG","LDC","ATOMS","ATOMG","RED"): pipe["lsu"] += 1
        else: pipe["ctl/other"] += 1
    print(f"{n} instructions; pipes {dict(pipe)}; IMAD.WIDE/HI {heavy} (x4 cycles); ALU x2 = {2*pipe['alu']} cycles")
    print("  " + ", ".join(f"{k} {v}" for k,v in ops.most_common()))
main()
