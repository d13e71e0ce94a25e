import csv,io,subprocess,sys
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
fname=None;hdr=None;seen=set();rows=[]
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0]=="File Path":
        fname=r[1].split("/")[-1]
        if fname in seen: break
        seen.add(fname)
    elif r[0]=="Line No": hdr=r
    elif hdr and r[0].isdigit():
        try: rows.append((fname,int(r[0]),int(r[hdr.index("# Samples")] or 0),int(r[hdr.index("Instructions Executed")] or 0)))
        except ValueError: pass
tot=sum(x[3] for x in rows); ts=sum(x[2] for x in rows)
nodes=float(sys.argv[2])
phases=eval(sys.argv[3])
acc={}
for f,ln,s,i in rows:
    key='other:'+f
    if f=='scan_small.cu':
        for name,(lo,hi) in phases.items():
            if lo<=ln<=hi: key=name;break
        else: key='small:unmapped'
    acc.setdefault(key,[0,0]); acc[key][0]+=i; acc[key][1]+=s
for k,(i,s) in sorted(acc.items(), key=lambda x:-x[1][0]):
    print(f"{k:28s} {100*i/tot:5.1f}% inst  {i*32/nodes:6.1f} thr-inst/node   {100*s/ts:5.1f}% stall samples")
