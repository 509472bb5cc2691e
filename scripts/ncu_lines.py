import re,csv,collections,sys,subprocess,os
rep, kregex, obj, mangled_prefix, nwarps_steps = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5])
subprocess.run(f"ncu -i {rep} --page source --csv --kernel-name regex:{kregex} > /tmp/_src.csv 2>/dev/null", shell=True)
os.makedirs('/tmp/cub2',exist_ok=True)
subprocess.run(f"cd /tmp/cub2 && rm -f *.cubin && cuobjdump -xelf all /root/repo/{obj} >/dev/null 2>&1 && nvdisasm --print-line-info *.cubin > all.txt 2>&1", shell=True)
txt=open('/tmp/cub2/all.txt').read().split('\n')
start=None
for i,l in enumerate(txt):
    if l.startswith('\t.section\t.text.'+mangled_prefix): start=i; break
cur=None; insts=[]
for l in txt[start+1:]:
    if l.startswith('\t.section'): break
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)',l)
    if m: cur=(m.group(1).split('/')[-1], int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m: insts.append((m.group(2), cur))
rows=list(csv.reader(open('/tmp/_src.csv')))
hdr=rows[1]; ia=hdr.index("Instructions Executed"); ist=hdr.index("Warp Stall Sampling (All Samples)")
data=rows[2:]
ends=[i for i,r in enumerate(data) if len(r)>0 and r[0]=="Kernel Name"]
if ends: data=data[:ends[0]]        # first captured launch of the kernel only
data=[r for r in data if len(r)>ia]
if len(data)!=len(insts): print("warning: %d profiled instructions vs %d disassembled (device functions called out of line are listed separately)"%(len(data),len(insts)))
n=min(len(data),len(insts)); data=data[:n]; insts=insts[:n]
agg=collections.Counter(); aggs=collections.Counter(); ops=collections.Counter(); opss=collections.Counter()
for k in range(len(data)):
    ex=int(data[k][ia]); st=int(data[k][ist])
    agg[insts[k][1]]+=ex; aggs[insts[k][1]]+=st
    s=insts[k][0]; op=(s.split()[1] if s.startswith('@') else s.split()[0]).split('.')[0]
    ops[op]+=ex; opss[op]+=st
tot=sum(agg.values()); tots=sum(aggs.values())
print("instr per step per warp: %.1f  static %d"%(tot/nwarps_steps, len(insts)))
for op,n in ops.most_common(16): print(f"  {op:8s} {n/nwarps_steps:7.1f}  stall {opss[op]/tots:.3f}")
srcs={}
for key,v in sorted(agg.items(), key=lambda kv:-aggs[kv[0]])[:int(sys.argv[6]) if len(sys.argv)>6 else 40]:
    line=''
    if key:
        f='/root/repo/quandary_b200/csrc/'+key[0]
        if os.path.exists(f):
            if f not in srcs: srcs[f]=open(f).read().split('\n')
            line=srcs[f][key[1]-1].strip()[:100]
    print(f"{v/nwarps_steps:7.1f} {aggs[key]/tots:6.3f} {key} {line}")
