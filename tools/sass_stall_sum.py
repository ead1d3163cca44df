import re,sys,collections
def load(fn):
    lines=open(fn).read().splitlines(); ins=[]; i=0
    while i<len(lines):
        m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/',lines[i])
        if m and i+1<len(lines):
            m2=re.match(r'\s+/\* (0x[0-9a-f]+) \*/',lines[i+1])
            if m2:
                hi=int(m2.group(1),16)
                ins.append((int(m.group(1),16),m.group(2),(hi>>41)&0xf)); i+=2; continue
        i+=1
    return ins
ins=load(sys.argv[1]); lo=int(sys.argv[2],16); hi=int(sys.argv[3],16)
skip=[]
for a in sys.argv[4:]:
    x,y=a.split('-'); skip.append((int(x,16),int(y,16)))
tot=0; n=0; byop=collections.Counter(); cnt=collections.Counter()
for a,t,s in ins:
    if a<lo or a>=hi or any(x<=a<y for x,y in skip): continue
    op=t.split()[0]
    if op.startswith('@'): op=t.split()[1]
    p=op.split('.')
    op=p[0]+('.'+p[1] if p[0] in('LDS','STS','LDG','STG') and len(p)>1 else '')
    tot+=s; n+=1; byop[op]+=s; cnt[op]+=1
print('instructions',n,'sum of stall fields',tot)
for k,v in byop.most_common(30): print('%-10s n=%4d stall=%5d avg=%.2f'%(k,cnt[k],v,v/cnt[k]))
