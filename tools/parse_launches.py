"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list."""
import csv, re, sys, collections
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
rows=list(csv.DictReader(lines))
tot=0; out=[]
for i,row in enumerate(rows):
    t=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    t = t/1e3 if u=='ns' else (t*1e3 if u=='ms' else t)
    name=re.sub(r'\(.*','',row['Kernel Name']).replace('void ','')
    out.append((i,name,row['Grid Size'],t)); tot+=t
agg=collections.OrderedDict()
for i,n,g,t in out:
    agg.setdefault(n,[0,0.0]); agg[n][0]+=1; agg[n][1]+=t
if len(sys.argv)>2:
    for i,n,g,t in out: print(i,n,g,round(t,1))
for n,(c,t) in agg.items(): print(f"{n:45s} x{c:3d} {t:9.1f} us  {100*t/tot:5.1f}%")
print("total us", round(tot,1))
