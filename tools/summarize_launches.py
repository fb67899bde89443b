import csv, collections, sys
lines=[l for l in open(sys.argv[1]) if l.startswith('"')]
r=csv.DictReader(lines)
tot=collections.defaultdict(float); cnt=collections.Counter()
for row in r:
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    name=row['Kernel Name'].split('(')[0]
    v=float(row['Metric Value'].replace(',',''))
    unit=row['Metric Unit']
    if unit=='ns': v/=1000
    elif unit=='ms': v*=1000
    tot[name]+=v; cnt[name]+=1
T=sum(tot.values())
print(f"total kernel time {T:.1f} us over {sum(cnt.values())} launches")
for k,v in sorted(tot.items(), key=lambda x:-x[1])[:int(sys.argv[2]) if len(sys.argv)>2 else 100]:
    print(f"{v:10.1f} us {100*v/T:5.1f}%  n={cnt[k]:4d} avg={v/cnt[k]:8.1f} us  {k[:100]}")
# usage: python tools/summarize_launches.py <ncu launch list .csv> [top-N]   (add --seq to print one scan's launches in order)
if '--seq' in sys.argv:
    rows=[row for row in csv.DictReader(lines) if row.get('Metric Name')=='gpu__time_duration.sum']
    st=[i for i,row in enumerate(rows) if row['Kernel Name'].startswith('k_begin_call')]
    if len(st)>=3:
        for row in rows[st[1]:st[2]]:
            v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
            v = v/1000 if u=='ns' else (v*1000 if u=='ms' else v)
            print(f"{v:8.1f} {row['Kernel Name'][:60]}")
