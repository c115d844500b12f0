import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
agg=collections.OrderedDict()
for r in rows[1:]:
    if r[ix['Metric Name']]!='gpu__time_duration.sum': continue
    k=r[ix['Kernel Name']][:60]+' grid='+r[ix['Grid Size']]; v=float(r[ix['Metric Value']].replace(',','')); u=r[ix['Metric Unit']]
    if u in ('ns','nsecond'): v/=1e3
    elif u in ('ms','msecond'): v*=1e3
    agg.setdefault(k,[]).append(v)
for k,v in agg.items(): print(f"{k:90s} n={len(v):3d} mean {sum(v)/len(v):10.1f} us  min {min(v):10.1f} max {max(v):10.1f}")
