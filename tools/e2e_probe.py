import sys, time, threading; sys.path.insert(0,'.')
import numpy as np
from dream_go_b200 import nn, weights
t=weights.synthetic_network(seed=20261017,num_blocks=9)
net=nn.Network.from_tensors(t,max_batch=256,num_workspaces=4)
f=weights.bernoulli_features(256,seed=1)
for callers in (1,2,3,4):
    s=net.time_e2e(f,400,callers)
    print('native callers',callers,'evals/s %.0f'%(256*400/s), 'us/step %.1f'%(s/400*1e6))
ms,tms,l=net.time_resident(256,200,tower=True,flush_l2=True); print('resident flush ms/step',ms/200, 'tower',tms/200)
ms,tms,l=net.time_resident(256,200,tower=True,flush_l2=False); print('resident noflush ms/step',ms/200, 'tower', tms/200)
