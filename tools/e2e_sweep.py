import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
import torch, yoxel_voxel_b200 as yv
svo = yv.SVOData.SphereFractal(12); svo.Upload(0)
r = yv.SVORenderer(0); r.SetScene(svo); r.SetResolution(1920,1080); r.SetViewPos((0.5,0.5,0.3)); r.SetViewDir((-1,-1,1.5))
flush = torch.empty(256<<20, dtype=torch.uint8, device='cuda')
ref = None
for chunks, taper in [(0,100),(4,100),(4,80),(5,70),(-1,0)]:
    if chunks < 0: r.SetOption("zero_copy", 1)
    else: r.SetOption("pipeline", chunks); r.SetOption("pipeline_taper", taper)
    ts=[]
    for i in range(60):
        flush.zero_(); torch.cuda.synchronize()
        t0=time.perf_counter(); r.SetViewPos((0.5,0.5,0.3)); img=r.RenderFrame(); ts.append(time.perf_counter()-t0)
    ref = img.copy() if ref is None else ref
    assert (img == ref).all()
    print(chunks, taper, 'e2e ms median %.4f min %.4f'%(np.median(ts[10:])*1e3, min(ts)*1e3), 'launches', r.LastFrameLaunches(), 'dev ms %.4f'%r.LastFrameMs())
