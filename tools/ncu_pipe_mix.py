#!/usr/bin/env python3
"""Per source line: executed warp-instructions split by issue pipe (ALU / FMA-heavy / LSU / other), spin loops excluded.
Usage: tools/ncu_pipe_mix.py rep.ncu-rep [nblocks]"""
import csv, io, subprocess, sys, collections, re
ALU = {"LOP3","SHF","SEL","ISETP","IADD3","VIADD","LEA","BREV","PRMT","POPC","VIMNMX","VIADDMNMX","FLO","IABS","IMNMX","PLOP3","P2R","R2P","FMNMX","FSEL","FSETP","MOV","IADD","VABSDIFF","VABSDIFF4","IDP","IDP4A","I2I","I2IP"}
FMA = {"IMAD","FFMA","FMUL","FADD","HFMA2","IDP4A_"}
LSU = {"LDS","STS","ATOMS","LDG","STG","LDL","STL","LDC","LDSM","RED","ATOMG","ATOM"}
SPIN = {"SYNCS","NANOSLEEP"}
def main():
    rep = sys.argv[1]; nb = float(sys.argv[2]) if len(sys.argv) > 2 else 4194304.0
    src = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],stdout=subprocess.PIPE,stderr=subprocess.DEVNULL,text=True).stdout
    head=None; fname="?"; cur=None
    per=collections.defaultdict(lambda: collections.Counter())
    tot=collections.Counter(); ops=collections.Counter()
    for r in csv.reader(io.StringIO(src)):
        if not r: continue
        if r[0]=="File Path": fname=r[1].split("/")[-1]; continue
        if r[0]=="Line No": head=r; continue
        if head is None or len(r)!=len(head): continue
        if r[0]!="":
            cur="%s:%s"%(fname,r[0]); curtext=r[1].strip()[:90]; per[cur]["_t"]=curtext; continue
        try: n=float(r[head.index("Instructions Executed")] or 0)
        except ValueError: continue
        s=r[3].strip().split()
        if not s: continue
        op=s[0]
        if op.startswith("@"): op=s[1]
        base=op.rstrip(";").split(".")[0]
        if base in SPIN: cls="spin"
        elif base in ALU: cls="alu"
        elif base in FMA: cls="fma"
        elif base in LSU: cls="lsu"
        else: cls="oth"
        per[cur][cls]+=n; tot[cls]+=n; ops[(cls,base)]+=n
    T=sum(v for k,v in tot.items() if k!="spin")
    print("warp-inst (no spin) %d = %.1f thread-inst/block; alu %.1f%% fma %.1f%% lsu %.1f%% oth %.1f%%; spin %d"%(T,T*32/nb,100*tot["alu"]/T,100*tot["fma"]/T,100*tot["lsu"]/T,100*tot["oth"]/T,tot["spin"]))
    print("ops:", ", ".join("%s %.1f%%"%(b,100*v/T) for (c,b),v in ops.most_common(24) if c!="spin"))
    rows=[(sum(v for k,v in c.items() if k not in("_t","spin")),l,c) for l,c in per.items()]
    rows.sort(key=lambda x:-x[0])
    print("%-24s %6s %6s %6s %6s  (thread-inst per block)"%("line","alu","fma","lsu","oth"))
    for n,l,c in rows[:70]:
        f=32/nb
        print("%-24s %6.1f %6.1f %6.1f %6.1f  %s"%(l,c["alu"]*f,c["fma"]*f,c["lsu"]*f,c["oth"]*f,c["_t"]))
main()
