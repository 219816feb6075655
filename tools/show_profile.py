import json, sys, collections
d=json.load(open(sys.argv[1]))
ops=d['ops']
print("total", d['total_ms'])
agg=collections.defaultdict(lambda:[0,0,0,0])
for r in ops:
    k=r['name'].split('.',1)[1] if '.' in r['name'] else r['name']
    agg[k][0]+=r['ms']; agg[k][1]+=r['flops']; agg[k][2]+=r['bytes']; agg[k][3]+=1
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:24]:
    print(f"{k:28s} n={v[3]:3d} {v[0]:7.3f} ms  {v[1]/(v[0]*1e-3)/1e12 if v[1] else 0:7.1f} TF/s  {v[2]/(v[0]*1e-3)/1e9:8.1f} GB/s")
if len(sys.argv)>2:
    for r in ops:
        if sys.argv[2] in r['name']:
            tf = r['flops']/(r['ms']*1e-3)/1e12 if r['flops'] else 0
            print(f"{r['name']:32s} {r['ms']*1e3:8.1f} us  {tf:7.1f} TF/s {r['bytes']/(r['ms']*1e-3)/1e9:8.1f} GB/s")
