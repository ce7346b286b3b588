import csv,re,sys
path=sys.argv[1]; which=int(sys.argv[2]) if len(sys.argv)>2 else 3
rows=[]
lines=[l for l in open(path) if not l.startswith('==')]
for d in csv.DictReader(lines):
    if d.get('Metric Name')=='gpu__time_duration.sum':
        rows.append((d['Kernel Name'].split('(')[0], float(d['Metric Value'].replace(',',''))/1000.0, d['Grid Size']))
names=[r[0] for r in rows]
starts=[i for i,n in enumerate(names) if 'fbank_logmel' in n]
s,e=starts[which],starts[which+1]
tot=0; agg={}
for i in range(s,e):
    n=re.sub(r'ppv::|_kernel|void ','',names[i]); t=rows[i][1]; tot+=t; agg[n]=agg.get(n,0)+t
    print(f"{i-s:3d} {n:28s} {t:9.1f} {rows[i][2]}")
print(f"# total {tot:.1f} us")
for k,v in sorted(agg.items(), key=lambda x:-x[1]): print(f"# share {k:28s} {v:9.1f} us {100*v/tot:5.1f}%")
