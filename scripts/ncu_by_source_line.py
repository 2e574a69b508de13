# like byline.py but with nvdisasm -gi inline chains: aggregate by the OUTERMOST frame (kernel body line) and by function-level frames
import csv, io, subprocess, sys, re, collections
rep, skip, disasm, mangled_pat = sys.argv[1:5]
NCH=float(sys.argv[5]); depth=int(sys.argv[6]) if len(sys.argv)>6 else -1
raw = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass","--launch-skip",skip,"--launch-count","1"],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][1][:100])
hdr = rows[1]; iS = hdr.index("Source"); iE = hdr.index("Instructions Executed")
ex = []
for r in rows[2:]:
    if len(r) > iE:
        try: ex.append((r[iS].strip(), int(r[iE])))
        except: pass
ins=[]; on=False; chain=[]; pending=[]
for line in open(disasm):
    if line.startswith(".text."):
        on = mangled_pat in line; chain=[]; continue
    if not on: continue
    m = re.search(r'//## File "([^"]*)", line (\d+)', line)
    if m:
        pending.append((m.group(1).split("/")[-1], int(m.group(2)))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        if pending: chain=pending; pending=[]
        ins.append((tuple(chain), m.group(2).strip()))
if len(ex)==2*len(ins): ex=[(a,(c+ex[i+len(ins)][1])//2) for i,(a,c) in enumerate(ex[:len(ins)])]
n=min(len(ex),len(ins)); print(len(ex),len(ins))
agg=collections.Counter(); ops=collections.defaultdict(collections.Counter)
for i in range(n):
    ch=ins[i][0]
    key=ch[depth] if ch else ("?",0)
    agg[key]+=ex[i][1]
    op=ins[i][1].split()[1] if ins[i][1].startswith('@') else ins[i][1].split()[0]
    ops[key][op.split('.')[0]]+=ex[i][1]
print("total", sum(agg.values())/NCH)
for (f,l),c in sorted(agg.items(), key=lambda x:-x[1])[:70]:
    top=", ".join(f"{o}:{v/NCH:.0f}" for o,v in ops[(f,l)].most_common(6))
    print(f"  {f:20s} {l:5d} {c/NCH:8.1f}   {top}")
