import csv,sys,collections
rows=[r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
t=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    v=float(r[vi].replace(',',''))
    u=r[ui]
    ms = v/1e6 if u in('ns','nsecond') else v/1e3 if u in('us','usecond') else v
    t[r[ki]][0]+=1; t[r[ki]][1]+=ms
tot=sum(v[1] for v in t.values())
print(f"{'kernel':58s} {'launches':>8s} {'total ms':>10s} {'share':>7s}")
for k,v in sorted(t.items(),key=lambda kv:-kv[1][1]):
    print(f"{k[:56]:58s} {v[0]:8d} {v[1]:10.3f} {100*v[1]/tot:6.1f}%")
print(f"{'total':58s} {sum(v[0] for v in t.values()):8d} {tot:10.3f}")
