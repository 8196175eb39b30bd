import torch, sys
sys.path.insert(0, ".")
from fragnet_b200 import ops
W = torch.zeros(128, 128, device="cuda")
def run(dh, x):
    _, dW, _ = ops.proj_bwd(x.cuda().contiguous(), W, dh.cuda().contiguous(), False, ops.PRECISION_TF32, want_db=False)
    torch.cuda.synchronize()
    return dW.cpu()
def show(name, got, want):
    nz = got.nonzero()
    print(name, "got nnz", len(nz), "want nnz", int((want != 0).sum()), "max", float(got.abs().max()), "first nz", nz[:6].tolist(), "vals", [float(got[i, j]) for i, j in nz[:6].tolist()])
n = 32
dh = torch.zeros(n, 128); x = torch.zeros(n, 128)
dh[0, 5] = 1; x[0] = torch.arange(1, 129).float()
show("p1", run(dh, x), dh.t() @ x)
dh = torch.zeros(n, 128); x = torch.zeros(n, 128)
dh[3, 40] = 1; x[3] = torch.arange(1, 129).float()
show("p2", run(dh, x), dh.t() @ x)
dh = torch.ones(n, 128); x = torch.zeros(n, 128); x[9, 77] = 2
show("p3", run(dh, x), dh.t() @ x)
dh = torch.randn(64, 128); x = torch.randn(64, 128)
g = run(dh, x); w = dh.t() @ x
print("rand", float((g - w).abs().max()), float(w.abs().max()), g[:2, :4], w[:2, :4])
