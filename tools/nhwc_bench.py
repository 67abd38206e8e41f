import sys, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from mcaq_yolo_b200 import ops, fused
B=64
for (C,H) in ((64,80),(128,40),(256,20)):
  for dt in (torch.bfloat16, torch.float32):
    es = 2 if dt==torch.bfloat16 else 4
    nbytes=B*C*H*H*es; nbuf=max(2,int(300e6//nbytes)+1)
    for layout in ("nchw","nhwc"):
        xs=[(torch.randn(B,C,H,H,device='cuda')*2+0.3).to(dt) for _ in range(nbuf)]
        if layout=="nhwc": xs=[x.contiguous(memory_format=torch.channels_last) for x in xs]
        ys=[torch.empty_like(x) for x in xs]
        bm=torch.randint(2,9,(B,H//8 if H>=40 else 5,H//8 if H>=40 else 5),device='cuda').float()
        m=torch.rand(B,H,H,device='cuda')*0.2+0.8
        s,a,keys=ops.reduce_planes(xs[0]); pk=ops.ranges_decode(keys)
        def t(fn):
            for i in range(nbuf): fn(i)
            torch.cuda.synchronize(); g=torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(nbuf): fn(i)
            g.replay(); torch.cuda.synchronize(); ts=[]
            for _ in range(10):
                e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
                e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)*1e3/nbuf)
            return sorted(ts)[5]
        t3=t(lambda i: ops.tile_quantize_ranges(xs[i],bm,pk,None,None,m,out=ys[i]))
        t1=t(lambda i: ops.reduce_planes_into(xs[i],s,a,keys))
        print(f"C={C} H={H} {str(dt)[6:]:8s} {layout}: K1 {t1:6.1f} us {nbytes/t1/1e3:6.0f} GB/s | K3 {t3:6.1f} us {2*nbytes/t3/1e3:6.0f} GB/s", flush=True)
