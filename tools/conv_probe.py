import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import pkg, orc
m=pkg(); ctx=m.Context(0)
rng=np.random.default_rng(0)
img=rng.uniform(0.1,1,(100,77)).astype(np.float32)
px=np.array([.1,.5,1,.5,.1,0.05,0.01],np.float32); py=px.copy()
out=ctx.conv2d(img,px,py,direct=True)
ref=orc.direct_convolve2d(img.astype(np.float64),np.outer(px,py).astype(np.float64))
print('err',np.abs(out-ref).max()/np.abs(ref).max())
