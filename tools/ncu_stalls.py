import csv, sys, subprocess, collections
rep=sys.argv[1]
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
h=None; tot=collections.Counter()
for row in rows:
    if row and row[0]=='Address': h=row; continue
    if h is None or len(row)<len(h): continue
    for i,name in enumerate(h):
        if name.startswith('stall_') and '(Not Issued)' not in name:
            try: tot[name]+=int(row[i])
            except: pass
s=sum(tot.values())
for k,v in tot.most_common(): print(f'{k:28s} {100*v/s:5.1f}%')
