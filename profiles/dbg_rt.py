import sys; sys.path.insert(0,'/root/repo')
import numpy as np
import __graft_entry__ as g
pkg=g.load_package(); orc=g.load_oracle()
import cpp_audio_b200.workloads as wl
ctx=pkg.Context(0)
x=wl.c4_streams(1,1.0)
out,stats=pkg.resynth_roundtrip(ctx,x)
ref,rstats=orc.roundtrip(x)
e=np.abs(out-ref)[0]
print("stats",stats,rstats)
for b in range(0,48000,2400):
    print(b, "err L %.2e R %.2e  ref %.3f"%(e[b:b+2400,0].max(), e[b:b+2400,1].max(), np.abs(ref[0,b:b+2400]).max()))
i0=4095
print("first samples out", out[0,i0:i0+6,0], "ref", ref[0,i0:i0+6,0])
print("ratio", out[0,i0+1000:i0+1006,0]/ref[0,i0+1000:i0+1006,0])
# single sine test
n=np.arange(30000); x1=(0.5*np.sin(2*np.pi*1000*n/48000)).astype(np.float32)[None,:]
o1,s1=pkg.resynth_roundtrip(ctx,x1); r1,_=orc.roundtrip(x1)
e1=np.abs(o1-r1)[0]
print("single sine stats",s1,"max err",e1.max(), "at", e1.argmax()//2)
for b in range(4000,30000,2600): print(b,"%.2e"%e1[b:b+2600].max(), "ref %.3f"%np.abs(r1[0,b:b+2600]).max())
print(o1[0,4095:4100,0], r1[0,4095:4100,0])
print(o1[0,8000:8005,0], r1[0,8000:8005,0])
