"""Experiment behind DESIGN.md 3c: how many lattice columns of a row hold forward / backward values within 2^-thr of the
row maximum (measured on the compiled reference's own lattices, oracle.Reference.stages).  CPU only; needs oracle/_ref."""
import os, sys, numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import Reference
from dynamont_b200 import synth
md=os.path.join(ROOT, 'tests', 'golden', '_models')
def run(model, pore, length, spb, sd_scale=1.0, kind=None, seed=1, dwell='geometric'):
    path=f'{md}/{model}.model'
    nm,ns=synth.native_model(path,pore)
    rna,k=synth.PORE_INFO[pore]
    rng=np.random.default_rng(seed)
    sd=None
    if kind: sd=synth.low_complexity_digits(rng,length,kind,k)
    sig,seq,b=synth.synth_read(rng,nm,ns,k,length,spb,sd_scale=sd_scale,seq_digits=sd,dwell=dwell)
    ref=Reference(path,pore)
    T=sig.size+1
    rows=np.arange(1,T-1,max(1,T//400))
    st=ref.stages(sig,seq,rows=rows)
    R=st['rows']  # [r,4,N]
    Z=st['Zb']
    L2=np.log(2.0)
    out=[]
    for thr in (40,80,126):
        wf=[];wb=[];wp=[];off=[]
        for i,t in enumerate(rows):
            f=np.maximum(R[i,0],R[i,1]); bb=np.maximum(R[i,2],R[i,3])
            fa=np.where(f>f.max()-thr*L2)[0]; ba=np.where(bb>bb.max()-thr*L2)[0]
            wf.append(fa.max()-fa.min()+1); wb.append(ba.max()-ba.min()+1)
            u=np.union1d(fa,ba); wp.append(u.max()-u.min()+1)
        out.append((thr,int(np.max(wf)),int(np.max(wb)),int(np.max(wp)),float(np.mean(wp))))
    # posterior width
    pw=[]
    for i,t in enumerate(rows):
        lp=np.logaddexp(R[i,0]+R[i,2],R[i,1]+R[i,3])-Z
        a=np.where(lp>np.log(1e-9))[0]
        pw.append(a.max()-a.min()+1)
    print(model,pore,length,spb,sd_scale,kind,dwell,'T',T,'widths(thr,maxF,maxB,maxUnion,meanUnion)',out,'postw max',max(pw))
run('synthetic_rna004_9mer','rna004',1000,30)
run('synthetic_rna004_9mer','rna004',3000,30)
run('trained_rna002_5mer','rna002',1000,30)
run('rna002_5mer','rna002',1000,30)
run('trained_rna002_5mer','rna002',1000,30,sd_scale=2.0)
run('rna002_5mer','rna002',1000,9)
run('rna002_5mer','rna002',600,8,sd_scale=1.5,kind='homopolymer')
run('rna002_5mer','rna002',600,8,sd_scale=1.5,kind='dinuc')
run('trained_rna002_5mer','rna002',600,8,sd_scale=1.5,kind='mixed')
