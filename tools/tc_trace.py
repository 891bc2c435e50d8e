import sys,collections
ev=[tuple(map(int,l.split())) for l in open(sys.argv[1]) if l.strip()]
names={1:"x1 wait start",2:"x1 data ready",3:"x2 wait start",4:"x2 data ready",5:"arrive a",6:"arrive b",7:"ab wait start",8:"ab ready",9:"a,b loaded",10:"arrive p",11:"x wait start",12:"x ready",13:"out loaded",14:"output stored",15:"threshold done",16:"stores read",17:"group sync 1",18:"staged",19:"fenced",20:"group sync 2"}
# durations between consecutive stamps, aggregated by (prev,cur)
agg=collections.defaultdict(list)
for (i0,t0),(i1,t1) in zip(ev,ev[1:]):
    agg[(i0,i1)].append(t1-t0)
tot=ev[-1][1]-ev[0][1]
print("stamps",len(ev),"total clk",tot)
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    v2=v[len(v)//10:]  # skip warm-up
    print(f"{names[k[0]]:16s} -> {names[k[1]]:16s} n={len(v):5d} mean={sum(v2)/len(v2):8.1f} share={100*sum(v)/tot:5.1f}%")
