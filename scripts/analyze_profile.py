import csv, collections, sys
path = sys.argv[1]; plan_n = sys.argv[2] if len(sys.argv) > 2 else '200'
rows=[r for r in csv.DictReader(open(path)) if r['plan_n']==plan_n]
tot=sum(float(r['ms_total']) for r in rows)
print("total ms", round(tot,2), "ops", len(rows))
agg=collections.OrderedDict()
for r in rows:
    if r['kind']!='gemm': continue
    key=(int(r['M']),int(r['N']),int(r['K']),int(r['block_n']),int(r['stages']),int(r['grid']),r['act'],r['out_fp32'],r['resid'])
    a=agg.setdefault(key,[0,0.0]); a[0]+=int(r['calls']); a[1]+=float(r['ms_total'])
print("%8s %5s %5s %4s %2s %4s act o32 res  calls   ms_tot  us/call  TF/s   GB/s(alg)  floor_us"%("M","N","K","bn","st","grid"))
gt=0
for k,(n,ms) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[3]) if len(sys.argv)>3 else 30]:
    M,N,K,bn,st,grid,act,o32,res=k
    us=ms/n*1e3
    fl=2.0*M*N*K
    by=M*K*2+N*K*2+M*N*(4 if o32=='1' else 2)+(M*N*4 if res=='1' else 0)
    floor=max(fl/1375e12, by/6.4e12)*1e6
    print("%8d %5d %5d %4d %2d %4d  %s   %s   %s  %5d %8.2f %8.1f %6.1f %8.1f %8.1f"%(M,N,K,bn,st,grid,act,o32,res,n,ms,us,fl/us/1e6,by/us/1e3,floor))
agg=collections.OrderedDict()
for r in rows:
    if r['kind']=='gemm': continue
    key=(r['kind'],r['M'],r['i0'],r['i1'],r['i2'],r['i3'],r['i4'])
    a=agg.setdefault(key,[0,0.0]); a[0]+=int(r['calls']); a[1]+=float(r['ms_total'])
for k,(n,ms) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]:
    print(k, n, "%.2f ms"%ms, "%.1f us/call"%(ms/n*1e3))
