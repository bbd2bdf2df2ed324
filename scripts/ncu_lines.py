"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by source line (first kernel instance)."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
topn=int(sys.argv[2]) if len(sys.argv)>2 else 30
blocks=[]; cur=None
for r in rows:
    if len(r)>=2 and r[0]=="File Path": cur={'file':r[1],'rows':[]}; blocks.append(cur); continue
    if len(r)>=2 and r[0]=="Function Name": cur['func']=r[1]; continue
    if len(r)>3 and r[0]=="Line No": cur['hdr']=r; continue
    if cur is not None and 'hdr' in cur and len(r)==len(cur['hdr']): cur['rows'].append(r)
seen=set(); first=[]
for b in blocks:
    if b['file'] in seen: break
    seen.add(b['file']); first.append(b)
tot=0; lines=[]
for b in first:
    h=b['hdr']; iS=h.index("# Samples"); iAddr=h.index("Address")
    stall_cols=[i for i,x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    for r in b['rows']:
        if r[iAddr]!="-": continue
        try: s=int(r[iS])
        except: continue
        if s>0:
            st={h[i]:int(r[i]) for i in stall_cols if r[i] not in ("","0","-")}
            lines.append((s,b['file'].split('/')[-1],r[0],r[1].strip()[:100],st)); tot+=s
print("total samples",tot)
for s,f,ln,src,st in sorted(lines,key=lambda x:-x[0])[:topn]:
    top=sorted(st.items(),key=lambda kv:-kv[1])[:3]
    print(f"{s:7d} {100*s/tot:5.1f}% {f}:{ln} {src}   {top}")
