import sys, glob, os, importlib
sys.path.insert(0, '/root/repo')
import torch
for path in sorted(glob.glob('/root/repo/h1v2_isaac_b200/lib_var*.so'), key=lambda p: int(p.split('lib_var')[1].split('.')[0])):
    import subprocess
    code = f"""
import sys; sys.path.insert(0,'/root/repo')
import torch
from h1v2_isaac_b200 import _capi
_capi.LIB_PATH = '{path}'
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config
n=32768
sim=H1v2Sim(n, default_config(), seed=1); sim.observe()
acts=[sim.random_actions(i) for i in range(8)]
obs=torch.empty((n,sim.obs_dim),device='cuda'); rew=torch.empty(n,device='cuda'); term=torch.empty(n,dtype=torch.uint8,device='cuda'); trunc=torch.empty(n,dtype=torch.uint8,device='cuda')
for i in range(30): sim.step_into(acts[i%8],obs,rew,term,trunc)
torch.cuda.synchronize()
best=1e9
for rep in range(3):
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(100): sim.step_into(acts[i%8],obs,rew,term,trunc)
    e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1)/100)
print('{os.path.basename(path)}', round(best,4), 'ms', float(rew.mean()))
"""
    print(subprocess.run([sys.executable, '-c', code], capture_output=True, text=True).stdout.strip())
