import torch, numpy as np
dev='cuda:0'
torch.manual_seed(0)
i = torch.randint(0,1200,(4096,),device=dev).float(); j = torch.randint(0,680,(4096,),device=dev).float()
cx,cy,fx,fy = 599.5,339.5,600.0,600.0
fx2 = 196.71854278974607
c2w = torch.randn(3,4,device=dev)
for f in (fx, fx2):
    d_eager = (i-cx)/f
    d_true = torch.from_numpy(((i.cpu().numpy()-np.float32(cx))/np.float32(f)).astype(np.float32)).to(dev)
    inv = np.float32(1.0)/np.float32(f)
    d_rcp = torch.from_numpy(((i.cpu().numpy()-np.float32(cx))*inv).astype(np.float32)).to(dev)
    print('f',f,'eager==true', bool(torch.equal(d_eager,d_true)), 'eager==rcp', bool(torch.equal(d_eager,d_rcp)))
dirs = torch.stack([(i-cx)/fx, -(j-cy)/fy, -torch.ones_like(i)], -1).reshape(-1,1,3)
prod = dirs*c2w[:3,:3]
s_eager = torch.sum(prod,-1)
p = prod.cpu().numpy()
v1 = (p[...,0]+p[...,1])+p[...,2]
v2 = p[...,0]+(p[...,1]+p[...,2])
v3 = (p[...,0]+p[...,2])+p[...,1]
se = s_eager.cpu().numpy()
print('sum (0+1)+2', np.array_equal(se,v1), ' 0+(1+2)', np.array_equal(se,v2), ' (0+2)+1', np.array_equal(se,v3))
# linspace cpu vs cuda
for n in (32,16,180,102):
    a = torch.linspace(0.,1.,n); b = torch.linspace(0.,1.,n,device=dev).cpu()
    print('linspace',n,'cpu==cuda', bool(torch.equal(a,b)))
a = torch.linspace(0,1199,180); b=torch.linspace(0,1199,180,device=dev).cpu(); print('linspace 0..1199/180', bool(torch.equal(a,b)))
