import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
os.environ['FRAD_DEBUG_LAUNCH']='1'
from frad_python_b200 import Context
from oracle.frad_oracle import synth_pcm
ctx=Context(0)
pcm=synth_pcm(64*2048/48000,48000,2,seed=5)[:64*2048]
q,tq,_=ctx.p1_encode(pcm,64,2048,2048,16,48000,0.55)
print('max|q|', np.abs(q).max())
out=ctx.p1_decode(q,tq,None,64,2048,2,16,48000)
