import json,sys,collections
d=json.load(open(sys.argv[1]))
agg=collections.OrderedDict()
for o in d:
    k=o.get('impl','?'); k=(k.split('>')[0]+'>') if '<' in k else k.split(' ')[0]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=o['ms']
print(sys.argv[1], 'total', round(sum(o['ms'] for o in d),3))
for k,a in sorted(agg.items(), key=lambda x:-x[1][1])[:12]: print(f"   {k:22s} n={a[0]:3d} ms={a[1]:.3f}")
