import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from sleqp_b200 import Fact, problems
p=problems.config(1); cp,ri,v=p.kkt_lower()
vp=torch.from_numpy(v).pin_memory().numpy()
f=Fact(device=0)
for i in range(3):
    print('--- call',i, file=sys.stderr); f.set_matrix(p.N,cp,ri,vp)
