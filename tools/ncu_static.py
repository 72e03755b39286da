import csv, sys, subprocess, collections
rep=sys.argv[1]
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
fname=''; line=None
addr={}   # address -> (file,line)
for row in rows:
    if row and row[0]=='File Path': fname=row[1].split('/')[-1]; continue
    if row and row[0] in ('Line No','Function Name'): continue
    if len(row)<4: continue
    if row[0]: line=int(row[0]); continue
    a=row[2]
    if a.startswith('0x'): addr.setdefault(int(a,16),(fname,line))
alla=sorted(addr)
base=alla[0]
print('instructions', len(alla), 'bytes', (alla[-1]-base+16))
# per-line static counts for sls_solve.cu
cnt=collections.Counter()
for a,(f,l) in addr.items(): cnt[(f,l)]+=1
# contiguous address map: print region boundaries by source line ranges
def region(f,l):
    if f!='sls_solve.cu': return f
    if l<171: return 'helpers(dmma/rcp)'
    if l<289: return 'setup'
    if l<479: return 'forward'
    if l<496: return 'lamT'
    if l<700: return 'backward'
    return 'tail'
reg=collections.Counter()
for a,(f,l) in addr.items(): reg[region(f,l)]+=1
for k,v in reg.most_common(): print(f'{k:24s} {v:6d} instr {v*16/1024:6.1f} KB')
# address span of forward loop (lines 289..479 incl. helpers inlined: approximate by min/max address of forward lines)
for name,lo,hi in (('forward',289,479),('backward',496,700)):
    aa=[a for a,(f,l) in addr.items() if f=='sls_solve.cu' and lo<=l<hi]
    print(name,'address span', (max(aa)-min(aa)+16)/1024,'KB', 'instr in span', sum(1 for a in alla if min(aa)<=a<=max(aa)))
print('top lines by static instruction count:')
for (f,l),v in cnt.most_common(25): print(f'  {v:5d}  {f}:{l}')
