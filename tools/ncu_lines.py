import csv, io, subprocess, sys
rep=sys.argv[1]; rx=sys.argv[2]; topn=int(sys.argv[3]) if len(sys.argv)>3 else 25
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass","--kernel-name","regex:"+rx],capture_output=True,text=True).stdout
blocks=[];cur=None
for row in csv.reader(io.StringIO(out)):
    if not row or row[0] in ("Kernel Name","Function Name"): continue
    if row[0] in ("File Name","File Path"):
        cur={"file":row[1],"rows":[],"hdr":None}; blocks.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"]=row
    elif row[0]!="": cur["rows"].append(row)
allrows=[]
for b in blocks:
    h={n:i for i,n in enumerate(b["hdr"])}
    def num(r,k):
        try: return float(r[h[k]].replace(",",""))
        except: return 0.0
    for r in b["rows"]:
        allrows.append((num(r,"Instructions Executed"), num(r,"Thread Instructions Executed"), b["file"].split("/")[-1], r[h["Line No"]], r[1].strip()[:110]))
tot=sum(a[0] for a in allrows)
print("total warp-inst", int(tot))
for f in sorted(set(a[2] for a in allrows)):
    print(f, "%.1f%%"%(100*sum(a[0] for a in allrows if a[2]==f)/tot))
for a in sorted(allrows,reverse=True)[:topn]:
    print("%5.1f%% lanes=%4.1f %s:%s | %s"%(100*a[0]/tot, a[1]/max(a[0],1), a[2], a[3], a[4]))
