import re,sys,collections
lo,hi=int(sys.argv[2],16),int(sys.argv[3],16)
c=collections.Counter()
for l in open(sys.argv[1]):
    m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);',l)
    if not m: continue
    a=int(m.group(1),16)
    if a<lo or a>=hi: continue
    ins=m.group(2).split()
    op=ins[0]
    if op.startswith('@'): op=ins[1]
    c[op.split('.')[0] + ('.'+op.split('.')[1] if op.startswith(('LDS','STS','LDG','STG')) and '.' in op else '')]+=1
tot=sum(c.values())
print(tot)
for k,v in c.most_common(40): print(v,k)
