"""CPU study (numpy): how loose is the screening bound of DESIGN.md 4.5? For a final classifier of the config-2
model fixture and 150 samples: cells per sample that survive the in-bag threshold with the current bound, with the
exact cell values as the bound (the floor of any bound-based screen) and with k-het-SNP class bounds."""
import os
import numpy as np, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
d=np.load('tests/golden/c2_model.npz')
coh=bench.make_cohort()
k=0
s0,s1=d['snp_off'][k],d['snp_off'][k+1]; h0,h1=d['hap_off'][k],d['hap_off'][k+1]
snp=d['snpidx'][s0:s1]; ns=len(snp)
freq=d['freq'][h0:h1]; hla=d['hla'][h0:h1]; packed=d['packed'][h0:h1,0]
nh=len(freq); n_hla=int(d['n_hla'])
H=((packed[:,None]>>np.arange(ns,dtype=np.uint64)[None,:])&np.uint64(1)).astype(np.int8)  # [nh, ns]
print('n_snp',ns,'n_hap',nh,'n_hla',n_hla)
T=np.exp(np.arange(0,2*ns+1)*np.log(1e-5)); T[0]=1.0
rng=np.random.default_rng(0)
samples=rng.choice(coh.n_samp, 150, replace=False)
G=coh.geno[:,snp]
al_idx=[np.where(hla==a)[0] for a in range(n_hla)]
tau=2.0**-70
res={'cur':[], 'exact':[], 'k1':[], 'k2':[], 'k3':[], 'k4':[]}
work={k:[] for k in res}
cellpairs=np.zeros((n_hla,n_hla))
for a in range(n_hla):
    for b in range(a,n_hla):
        na,nb=len(al_idx[a]),len(al_idx[b])
        cellpairs[a,b]= na*(na+1)/2 if a==b else na*nb
for s in samples:
    g=G[s]; t1,t2=sorted((coh.h1[s],coh.h2[s]))
    hom0=(g==0); hom2=(g==2); het=(g==1)
    c=(H[:,hom0]==1).sum(1)+(H[:,hom2]==0).sum(1)          # hom mismatches per haplotype
    # exact d(i,j) = c_i + c_j + #het with h_i==h_j
    Hh=H[:,het].astype(np.int32)
    nhet=het.sum()
    same=(Hh@Hh.T)+((1-Hh)@(1-Hh).T)                       # [nh,nh] agreements on het SNPs
    D=c[:,None]+c[None,:]+same
    W=(freq[:,None]*freq[None,:])*T[np.minimum(D,2*ns)]
    P=np.zeros((n_hla,n_hla))
    for a in range(n_hla):
        ia=al_idx[a]
        if len(ia)==0: continue
        for b in range(a,n_hla):
            ib=al_idx[b]
            if len(ib)==0: continue
            blk=W[np.ix_(ia,ib)]
            if a==b: P[a,b]=np.triu(blk,1).sum()*2+np.trace(blk)
            else: P[a,b]=2*blk.sum()
    xref=P[t1,t2]          # ~ (median x_ref/P(true) = 0.999)
    u=freq*T[np.minimum(c,2*ns)]
    U=np.array([u[al_idx[a]].sum() for a in range(n_hla)])
    iu=np.triu_indices(n_hla)
    def count(B):
        need=(B[iu]>=tau*xref)&(B[iu]>0)
        return need.sum(), (cellpairs[iu]*need).sum()
    B=2*np.outer(U,U)
    n,w=count(B); res['cur'].append(n); work['cur'].append(w)
    n,w=count(P); res['exact'].append(n); work['exact'].append(w)
    # class bound with k het SNPs (the first k het SNPs of the sample)
    hidx=np.where(het)[0]
    for kk in (1,2,3,4):
        ks=hidx[:kk]
        if len(ks)<kk: res['k%d'%kk].append(res['cur'][-1]); work['k%d'%kk].append(work['cur'][-1]); continue
        cls=(H[:,ks]*(1<<np.arange(kk))).sum(1)
        nc=1<<kk
        Uc=np.zeros((n_hla,nc))
        np.add.at(Uc,(hla,cls),u)
        # agreements between classes s,t on the k SNPs = kk - popc(s^t)
        M=np.array([[T[kk-bin(s^t).count('1')] for t in range(nc)] for s in range(nc)])
        Bk=2*(Uc@M@Uc.T)
        n,w=count(Bk); res['k%d'%kk].append(n); work['k%d'%kk].append(w)
tot=(cellpairs[np.triu_indices(n_hla)]).sum()
for kx in res:
    print(kx,'cells/sample %.1f'%np.mean(res[kx]),'work frac %.4f'%(np.mean(work[kx])/tot))
