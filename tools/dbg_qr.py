import sys, torch, time
sys.path.insert(0,'.')
from tensor_networks_b200.utils import orth_rows_dev, delta_svd_dev
torch.manual_seed(0)
for (c,m,rank) in [(256,4096,256),(256,4096,64),(1024,4096,1024),(1024,4096,64),(96,300,20)]:
    A = torch.randn(c,rank,dtype=torch.float64,device='cuda')@torch.randn(rank,m,dtype=torch.float64,device='cuda')/ (rank*m)**0.5
    M = A.clone()
    Q,R = orth_rows_dev(M)
    I = torch.eye(c,dtype=torch.float64,device='cuda')
    print(c,m,rank,"orth", float((Q@Q.T-I).abs().max()), "resid", float((R.T@Q-A).norm()/A.norm()), "Rfro/Afro", float(R.norm()/A.norm()))
for p in [64,256,1024]:
    X = torch.randn(p,p,dtype=torch.float64,device='cuda')
    t=time.time(); u,s,svt,info = delta_svd_dev(X, 0.0); torch.cuda.synchronize(); dt=time.time()-t
    sref = torch.linalg.svdvals(X)
    print("svd",p,"rank",info["rank"],"max rel err", float(((s-sref).abs()/sref[0]).max()), "fro2", info["fro2"], float((X**2).sum()), "time",dt)
    rec = u@svt
    print("   recon", float((rec-X).norm()/X.norm()), "orthU", float((u.T@u-torch.eye(u.shape[1],device='cuda',dtype=torch.float64)).abs().max()))
