import csv, sys
from collections import Counter
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=[r for r in rows[2:] if len(r)==len(hdr) and r[0]!='Address']
ix={h:i for i,h in enumerate(hdr)}
tot_inst=sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_samp=sum(int(r[ix["# Samples"]]) for r in data)
print("total warp-inst",tot_inst,"samples",tot_samp)
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot={h:sum(int(r[ix[h]]) for r in data) for h in stalls}
print({k:v for k,v in sorted(tot.items(), key=lambda kv:-kv[1])[:9]})
c=Counter(); ci=Counter()
for r in data:
    a=int(r[ix["Address"]],16)>>12
    c[a]+=int(r[ix["# Samples"]]); ci[a]+=int(r[ix["Instructions Executed"]])
base=min(c)
print("per 4KB code block: samples, inst")
print(" ".join(f"{a-base:x}:{c[a]}/{ci[a]//1000}k" for a in sorted(c)))
n=int(sys.argv[2]) if len(sys.argv)>2 else 30
top=sorted(data,key=lambda r:-int(r[ix["# Samples"]]))[:n]
for r in top:
    st={h:int(r[ix[h]]) for h in stalls}
    big=sorted(st.items(), key=lambda kv:-kv[1])[:2]
    print(r[ix["Address"]][-5:], r[ix["Source"]][:72].ljust(72), r[ix["# Samples"]], r[ix["Instructions Executed"]], big)
