import sys, torch
sys.path.insert(0, '/root/repo')
from imgadapointr_b200 import chamfer
from imgadapointr_b200.extensions.chamfer_dist import ChamferDistanceL1, ChamferDistanceL2
dev = torch.device('cuda:0')
def run(name, fn):
    try:
        out = fn(); torch.cuda.synchronize(); print(name, 'ok', out)
    except Exception as e:
        print(name, 'raised', type(e).__name__, str(e)[:80])
run('B=1 N=1 M=1', lambda: [t.tolist() for t in chamfer.forward(torch.rand(1,1,3,device=dev), torch.rand(1,1,3,device=dev))][2:])
r = chamfer.step(torch.rand(1,1,3,device=dev), torch.rand(1,1,3,device=dev), chamfer.LOSS_L1, grad1=True, grad2=True)
print('step 1x1', r.loss.tolist(), r.grad1.tolist())
run('N=0', lambda: [t.shape for t in chamfer.forward(torch.rand(2,0,3,device=dev), torch.rand(2,5,3,device=dev))])
run('M=0', lambda: [t.shape for t in chamfer.forward(torch.rand(2,5,3,device=dev), torch.rand(2,0,3,device=dev))])
run('B=0', lambda: [t.shape for t in chamfer.forward(torch.rand(0,5,3,device=dev), torch.rand(0,7,3,device=dev))])
run('B=3000 small', lambda: float(ChamferDistanceL1()(torch.rand(3000,37,3,device=dev).requires_grad_(True), torch.rand(3000,53,3,device=dev))))
x = torch.rand(2,300,3,device=dev); y = torch.rand(2,500,3,device=dev)
xn = x.clone(); xn[0,5] = float('nan'); xn[1,7,0] = float('inf')
run('NaN/Inf inputs fwd', lambda: [bool(torch.isfinite(t.float()).all()) for t in chamfer.forward(xn, y)])
def nanstep():
    a = xn.clone().requires_grad_(True)
    l = ChamferDistanceL2()(a, y); l.backward(); return float(l), bool(torch.isfinite(a.grad).all())
run('NaN/Inf step', nanstep)
xa = torch.full((2,400,3), float('nan'), device=dev)
run('all-NaN cloud', lambda: [t[0,:3].tolist() for t in chamfer.forward(xa, y)][:1])
def allnan_step():
    a = xa.clone().requires_grad_(True); l = ChamferDistanceL1()(a, y); l.backward(); return float(l)
run('all-NaN step', allnan_step)
big = torch.rand(1, 200000, 3, device=dev); small = torch.rand(1, 3, 3, device=dev)
run('200000 x 3 step', lambda: float(ChamferDistanceL1()(big.requires_grad_(True), small)))
run('3 x 200000 step', lambda: float(ChamferDistanceL1()(small.clone().requires_grad_(True), big.detach())))
print('done')
