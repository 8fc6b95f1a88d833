import csv, subprocess, sys
rep=sys.argv[1]; ntop=int(sys.argv[2]) if len(sys.argv)>2 else 50
txt=subprocess.run(["ncu","-i",rep,"--page","source","--print-source","cuda,sass","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(txt.splitlines()))
cur=None;hdr=None;agg={}
for r in rows:
    if len(r)>=2 and r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if len(r)>4 and r[0]=="Line No": hdr=r; continue
    if len(r)<6 or not r[0].isdigit() or hdr is None: continue
    ix={h:i for i,h in enumerate(hdr)}
    try: n=int(r[ix["Instructions Executed"]]); s=int(r[ix["Warp Stall Sampling (All Samples)"]])
    except: continue
    agg[(cur,int(r[0]))]=(n,s,r[1].strip()[:110])
tot=sum(v[0] for v in agg.values())
print("total inst %.2fG"%(tot/1e9))
for (f,l),v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:ntop]:
    print("%6.2fG %5.1f%%  %s:%d  %s"%(v[0]/1e9, 100*v[0]/tot, f,l,v[2]))
