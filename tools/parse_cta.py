import sys,re
runs=[];cur=[]
for l in sys.stdin:
    m=re.match(r'CTA (\d+) entry (\d+) prologue \+(\d+) dep \+(\d+) end \+(\d+)',l)
    if m: cur.append(tuple(int(x) for x in m.groups()))
    elif l.startswith('conv ') or l.startswith('=='):
        if cur: runs.append(cur); cur=[]
        print(l.strip())
if cur: runs.append(cur)
# each python invocation runs reps launches; split by count 148
for r in runs:
    n=148
    for i in range(0,len(r),n):
        g=r[i:i+n]
        e0=min(x[1] for x in g); e1=max(x[1]+x[4] for x in g)
        print(f"  launch: CTAs {len(g)} span {e1-e0} ns; entry spread {max(x[1] for x in g)-e0} ns; prologue avg {sum(x[2] for x in g)/len(g):.0f} ns; dep-wait avg {sum(x[3]-x[2] for x in g)/len(g):.0f}; life avg {sum(x[4] for x in g)/len(g):.0f} min {min(x[4] for x in g)} max {max(x[4] for x in g)}")
