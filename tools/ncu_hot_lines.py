import csv, sys, subprocess
rep=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 40
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
h=None; fname=''; out=[]
for row in rows:
    if row and row[0]=='File Path': fname=row[1].split('/')[-1]; continue
    if row and row[0]=='Line No': h=row; continue
    if h is None or len(row)<10: continue
    if row[0]:
        ii=[i for i,x in enumerate(h) if x=='Instructions Executed'][0]
        si=[i for i,x in enumerate(h) if x=='# Samples'][0]
        def _i(x):
            try: return int(x)
            except Exception: return 0
        out.append((_i(row[ii]), _i(row[si]), fname, row[0], row[1].strip()[:110]))
tot=sum(o[0] for o in out); stot=sum(o[1] for o in out)
print('total', tot, stot)
for n,s,f,l,t in sorted(out,reverse=True)[:topn]:
    print(f'{100*n/tot:5.1f}% inst {100*s/max(stot,1):5.1f}% samp  {f}:{l}  {t}')
