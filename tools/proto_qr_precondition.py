"""NumPy prototype behind the QR-preconditioned SVD path: sweeps and rotations of cyclic one-sided Jacobi on the
rows of a graded matrix A, of its triangular factors L (A = L Q) and L^T, with and without column pivoting."""
import numpy as np, scipy.linalg as sl
rng=np.random.default_rng(0)
def make(R,C,decay):
    u,_=np.linalg.qr(rng.standard_normal((R,R))); v,_=np.linalg.qr(rng.standard_normal((C,R)))
    s=10.0**(-decay*np.arange(R)/R)
    return (u*s)@v.T, s
def jacobi_rows(X,tol=1e-15,max_sweeps=60):
    X=X.copy(); R=X.shape[0]; counts=[]
    for sweep in range(max_sweeps):
        rot=0
        for p in range(R-1):
            for q in range(p+1,R):
                a=X[p]@X[p]; b=X[q]@X[q]; g=X[p]@X[q]
                if g==0 or abs(g)<=tol*np.sqrt(a*b): continue
                rot+=1
                zeta=(b-a)/(2*g); t=np.sign(zeta)/(abs(zeta)+np.sqrt(1+zeta*zeta)) if zeta!=0 else 1.0
                c=1/np.sqrt(1+t*t); s=c*t
                xp=X[p].copy(); X[p]=c*xp-s*X[q]; X[q]=s*xp+c*X[q]
        counts.append(rot)
        if rot==0: break
    return X,counts
R,C=96,192
for decay in (2,8,16):
    A,s=make(R,C,decay)
    _,c0=jacobi_rows(A)
    # LQ of A (rows): A = L Q  <=> A^T = Q^T R, L = R^T
    Q,Rr=np.linalg.qr(A.T); L=Rr.T
    _,c1=jacobi_rows(L)          # rows of L
    _,c1t=jacobi_rows(L.T)       # rows of L^T = columns of L (Drmac-Veselic recommend the transposed factor)
    Qp,Rp,piv=sl.qr(A.T,mode='economic',pivoting=True); Lp=Rp.T
    _,c2=jacobi_rows(Lp); _,c2t=jacobi_rows(Lp.T)
    X,_=jacobi_rows(Lp.T); sv=np.sort(np.linalg.norm(X,axis=1))[::-1]
    print(f"decay 1e-{decay}: plain {len(c0)} sweeps ({sum(c0)} rot) | L {len(c1)} ({sum(c1)}) | L^T {len(c1t)} ({sum(c1t)}) | pivoted L {len(c2)} ({sum(c2)}) | pivoted L^T {len(c2t)} ({sum(c2t)})  rel sv err {np.max(np.abs(sv-s)/s):.1e}")
