import csv, sys, collections
path=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 40
rows=list(csv.reader(open(path)))
cur_file=None; hdr=None; out=[]
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur_file=r[1].split('/')[-1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; continue
    if hdr and r[0] not in ("",) and r[0].isdigit():
        d=dict(zip(hdr,r))
        # duplicate "Source" header: first is cuda source
        try:
            inst=int(r[hdr.index("Instructions Executed")]); samp=int(r[hdr.index("# Samples")])
        except ValueError: continue
        out.append((cur_file,int(r[0]),r[1].strip()[:90],inst,samp))
tot_i=sum(o[3] for o in out); tot_s=sum(o[4] for o in out)
print("total inst",tot_i,"samples",tot_s)
agg=collections.OrderedDict()
for f,l,s,i,sm in out:
    k=(f,l,s); a=agg.setdefault(k,[0,0]); a[0]+=i; a[1]+=sm
for (f,l,s),(i,sm) in sorted(agg.items(), key=lambda x:-x[1][1])[:topn]:
    print(f"{f:16s}:{l:4d} inst {i/tot_i:6.1%} samp {sm/tot_s:6.1%}  {s}")
