import csv,sys
rows=list(csv.reader(open(sys.argv[3])))
secs=[];cur=None
for r in rows:
    if r and r[0]=='Kernel Name':
        cur={'hdr':None,'rows':[]};secs.append(cur);continue
    if cur is None: continue
    if cur['hdr'] is None: cur['hdr']=r;continue
    cur['rows'].append(r)
which=int(sys.argv[1])  # usage: ncu -i X.ncu-rep --page source --csv > src.csv; python profiles/src_stalls.py <section> <top n> src.csv
for i,s in enumerate(secs):
    h=s['hdr']; ix={n:j for j,n in enumerate(h)}
    tot=sum(int(r[ix['# Samples']]) for r in s['rows'] if r[ix['# Samples']].isdigit())
    ins=sum(int(r[ix['Instructions Executed']]) for r in s['rows'] if r[ix['Instructions Executed']].isdigit())
    print(i,h[0],h[1][:20],len(s['rows']),'samples',tot,'inst',ins)
s=secs[which]; h=s['hdr']; ix={n:j for j,n in enumerate(h)}
reasons=[n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
agg={n:0 for n in reasons}
for r in s['rows']:
    for n in reasons:
        try: agg[n]+=int(r[ix[n]])
        except: pass
tot=sum(agg.values())
print('--- stall reasons (all samples)')
for n,v in sorted(agg.items(),key=lambda x:-x[1])[:10]: print('  %-28s %6d %.1f%%'%(n,v,100*v/tot))
print('--- top instructions')
top=sorted(s['rows'],key=lambda r:-int(r[ix['# Samples']]))[:int(sys.argv[2])]
for r in top:
    rs=sorted(((int(r[ix[n]]),n) for n in reasons),reverse=True)[:2]
    print('%s %-60s smp %5s exec %7s  %s'%(r[0][-5:],r[1].strip()[:60],r[ix['# Samples']],r[ix['Instructions Executed']],' '.join('%s=%d'%(n[6:],v) for v,n in rs)))
