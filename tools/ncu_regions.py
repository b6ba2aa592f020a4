import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
sections=[]; cur=None
for r in rows:
    if r and r[0]=='File Path':
        cur={'file':r[1],'rows':[],'hdr':None}; sections.append(cur)
    elif r and r[0]=='Line No':
        cur['hdr']=r
    elif cur is not None and cur['hdr'] is not None and r and r[0]!='Function Name':
        cur['rows'].append(r)
lines=[]
seenfiles=set()
for s in sections:
    f=s['file'].split('/')[-1]
    if f in seenfiles: continue
    seenfiles.add(f)
    h=s['hdr']; iI=h.index('Instructions Executed'); iN=h.index('# Samples')
    for r in s['rows']:
        if r[0]=='': continue
        try: n=int(r[iI]); smp=int(r[iN])
        except: continue
        lines.append((int(r[0]),f,n,smp))
tot=sum(l[2] for l in lines); st=sum(l[3] for l in lines)
src=open('/root/repo/tess_b200/csrc/tess_image.cuh').read().split('\n')
marks=[]
for i,l in enumerate(src,1):
    if '// ----' in l or l.startswith('template') or l.startswith('TESS_') : marks.append(i)
marks.append(len(src)+1)
steps=float(sys.argv[2]) if len(sys.argv)>2 else 262144.0
print('total warp-inst %d  per warp-step %.0f'%(tot,tot/steps))
for a,b in zip(marks,marks[1:]):
    n=sum(l[2] for l in lines if l[1]=='tess_image.cuh' and a<=l[0]<b); s=sum(l[3] for l in lines if l[1]=='tess_image.cuh' and a<=l[0]<b)
    if n>0.004*tot: print('%4d-%4d %5.1f%% (%5.0f/step) smp %5.1f%%  %s'%(a,b-1,100.0*n/tot,n/steps,100.0*s/st,src[a-1].strip()[:90]))
for f in set(l[1] for l in lines):
    if f!='tess_image.cuh':
        n=sum(l[2] for l in lines if l[1]==f); s=sum(l[3] for l in lines if l[1]==f)
        print('%-30s %5.1f%% (%5.0f/step) smp %5.1f%%'%(f,100.0*n/tot,n/steps,100.0*s/st))
