import csv, sys, subprocess, collections
rep=sys.argv[1]
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
        out.append((fname,int(row[0]),col('# Samples'),col('Instructions Executed')))
tot=sum(o[2] for o in out); toti=sum(o[3] for o in out)
bounds=[int(x) for x in sys.argv[2].split(',')]
names=sys.argv[3].split(',')
reg=collections.Counter(); regi=collections.Counter()
for f,l,s,i in out:
    if f!='sls_solve.cu': reg['other:'+f]+=s; regi['other:'+f]+=i; continue
    k=0
    while k<len(bounds) and l>=bounds[k]: k+=1
    reg[names[k]]+=s; regi[names[k]]+=i
for k,v in reg.most_common(): print(f'{k:40s} {100*v/tot:5.1f}% samples  {100*regi[k]/toti:5.1f}% inst')
