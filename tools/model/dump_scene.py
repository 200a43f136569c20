import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))); os.makedirs('tools/model/_data', exist_ok=True)
import numpy as np, yoxel_voxel_b200 as yv
depth=int(sys.argv[1])
svo=yv.SVOData.SphereFractal(depth)
recs,leaves=svo.packed()
np.ascontiguousarray(recs,np.uint32).tofile('tools/model/_data/recs.bin')
np.ascontiguousarray(leaves,np.uint32).tofile('tools/model/_data/leaves.bin')
W,H=1920,1080
d0,du,dv=yv.init_ray_dir((-1,-1,1.5),(0,0,1),70.0,W,H)
np.concatenate([d0,du,dv]).astype(np.float32).tofile('tools/model/_data/cam.bin')
print(len(recs))
