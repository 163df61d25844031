import sys, json, time, numpy as np
sys.path.insert(0,'.')
from oracle import oracle as O
from bioinformatics_toolkit_b200 import synth
bg=O.DEF_BG
import pickle, os
which = sys.argv[1] if len(sys.argv)>1 else 'c2'
if which=='c2':
    mats = synth.synth_pwms(100, 12)
else:
    mats = synth.config_c4_motifs(200)
cache=f'experiments/thres_{which}.pkl'
if os.path.exists(cache):
    thres=pickle.load(open(cache,'rb'))
else:
    t=time.time()
    thres=[O.pvalue_to_score(1e-5,bg,m) for m in mats]
    print('cdf time',time.time()-t)
    pickle.dump(thres,open(cache,'wb'))
Q=int(sys.argv[2]) if len(sys.argv)>2 else 120
def slot_tables(S, c, NL, NR, D):
    # S: n x 4 score table; returns list of 256-entry int tables for slots rho=-NL..1+NR
    n=S.shape[0]
    d = S.max(axis=1,keepdims=True)-S   # deficits
    scale=Q/D
    tabs=[]
    for rho in range(-NL, 2+NR):
        e=np.zeros(256)
        for j in range(4):
            col=c+4*rho+j
            if 0<=col<n:
                # base j of kmer = bits 2j..2j+1
                bj=(np.arange(256)>>(2*j))&3
                e+=d[col][bj]
        q=np.minimum(Q+1, np.floor(scale*e*(1-1e-12))).astype(int)
        tabs.append(q)
    return tabs
def dist(tab):
    h=np.bincount(tab, minlength=Q+2)/256.0
    return h
def conv_sat(h1,h2):
    r=np.convolve(h1,h2)
    out=np.zeros(Q+2); out[:Q+1]=r[:Q+1]; out[Q+1]=r[Q+1:].sum()
    return out
rows=[]
for mi,(m,th) in enumerate(zip(mats,thres)):
    n=m.shape[0]
    for strand in (0,1):
        mm = m if strand==0 else O.rc_pwm(m)
        S=O.score_table(bg,mm)
        Smax=S.max(axis=1).sum()
        D=Smax-th
        if not np.isfinite(th) or D<=0:
            rows.append((mi,strand,n,D,None,None,None,None)); continue
        best=None
        for c in range(min(0,n-8), max(0,n-8)+1):
            NL=(c+3)//4 if c>0 else 0
            NR=max(0,(n-c-8+3)//4)
            tabs=slot_tables(S,c,NL,NR,D)
            core=conv_sat(dist(tabs[NL]),dist(tabs[NL+1]))
            s1=core[:Q+1].sum()
            # 3-slot core: best of adding left or right neighbour
            s3=None
            if len(tabs)>=3:
                cands=[]
                if NL>0: cands.append(conv_sat(core,dist(tabs[NL-1]))[:Q+1].sum())
                if NR>0: cands.append(conv_sat(core,dist(tabs[NL+2]))[:Q+1].sum())
                s3=min(cands) if cands else s1
            full=dist(tabs[0])
            for t in tabs[1:]: full=conv_sat(full,dist(t))
            cand=full[:Q+1].sum()
            if best is None or s1<best[0]: best=(s1,s3,cand,c,NL,NR)
        rows.append((mi,strand,n,D)+best[:4])
s1s=np.array([r[4] for r in rows if r[4] is not None])
s3s=np.array([r[5] if r[5] is not None else r[4] for r in rows if r[4] is not None])
cands=np.array([r[6] for r in rows if r[4] is not None])
ns=np.array([r[2] for r in rows if r[4] is not None])
print('combos',len(rows),'active',len(s1s))
print('sigma1 mean',s1s.mean(),'median',np.median(s1s),'max',s1s.max())
print('sigma3 mean',s3s.mean(),'median',np.median(s3s),'max',s3s.max())
print('cand mean',cands.mean(),'max',cands.max())
for lo,hi in ((6,9),(10,12),(13,16),(17,22)):
    sel=(ns>=lo)&(ns<=hi)
    if sel.any(): print(lo,hi,sel.sum(),'s1',s1s[sel].mean(),'s3',s3s[sel].mean(),'cand',cands[sel].mean())
Ds=np.array([r[3] for r in rows])
print('D quantiles',np.quantile(Ds[np.isfinite(Ds)],[0,0.1,0.5,0.9,1]))
