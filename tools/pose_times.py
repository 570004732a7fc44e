import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
import bench
from agi_lidar_slam_b200 import _cabi
sys.argv=['bench.py']; args=bench.parse(); wl=bench.make_workload(args,0)
dev=torch.device('cuda',0); stream=torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
ctx=_cabi.Context(0,max_scan_points=1<<18,max_down_points=100000,max_map_points=1<<21); ctx.set_stream(stream.cuda_stream)
mp=wl['map']; ctx.map_build(np.concatenate([mp,np.zeros((len(mp),1),np.float32)],1))
bodies=[ctx.scan_preprocess(s['scan'],None,None,0.5)[0] for s in wl['scans']]
flush=torch.empty(384<<20,dtype=torch.uint8,device=dev)
for cold in (True,False):
    out=[]
    for j,b in enumerate(bodies):
        ts=[]
        for r in range(12):
            ctx.scan_upload(b); ctx.state_upload(wl['scans'][j]['x_prior'],wl['P'])
            if cold: flush.fill_(1)
            e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            e0.record(stream); ctx.update_enqueue(0.001,4,False,from_snapshot=True); e1.record(stream)
            x,P,nv,npz=ctx.state_download(); ts.append(e0.elapsed_time(e1))
        out.append((len(b),npz,round(float(np.median(ts[2:]))*1e3,1)))
    print('cold' if cold else 'warm', out)
