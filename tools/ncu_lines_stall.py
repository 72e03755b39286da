import csv, sys, subprocess
rep=sys.argv[1]; reason=sys.argv[2]; topn=int(sys.argv[3]) if len(sys.argv)>3 else 30
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
h=None; fname=''; out=[]
for row in rows:
    if row and row[0]=='File Path': fname=row[1].split('/')[-1]; continue
    if row and row[0]=='Line No': h=row; continue
    if h is None or len(row)<10: continue
    if row[0]:
        def col(name):
            i=[i for i,x in enumerate(h) if x==name][0]
            try: return int(row[i])
            except Exception: return 0
        out.append((col(reason), col('# Samples'), col('Instructions Executed'), fname, row[0], row[1].strip()[:100]))
tot=sum(o[0] for o in out); stot=sum(o[1] for o in out)
print('total', reason, tot, 'of samples', stot)
for n,s,i,f,l,t in sorted(out,reverse=True)[:topn]:
    print(f'{100*n/max(tot,1):5.1f}% of {reason}  {100*s/max(stot,1):5.1f}% samp  {f}:{l}  {t}')
