import sys, numpy as np
sys.path.insert(0,'/root/repo')
import jfnk_b200 as jf
from tests.hostsim.sim import SimBuffers
from tests.test_mesh_parity import _pma2_state, relmax
for N in (51, 300):
    Q,u,u0=_pma2_state(N)
    dt = 1e-4 * ((2.0 / (N - 1)) / 0.04) ** 4
    res={}
    for name,buf,var in (("sim",SimBuffers(),0),("gpu_pt",jf.CudaBuffers(),1),("gpu_march",jf.CudaBuffers(),2)):
        F=jf.PMA2Residual(N=N, dt=dt, buffers=buf, kernel_variant=var); F.set_mesh(Q); F.set_prev(u0)
        vxx,vyy=F.laplace(u)
        res[name]=(F(u),vxx,vyy, F(u0))
    for name in ("gpu_pt","gpu_march"):
        for i,lab in enumerate(("F(u)","vxx","vyy","F(u0)")):
            a=res[name][i].reshape(N,N); b=res["sim"][i].reshape(N,N)
            d=np.abs(a-b)
            loc=np.unravel_index(d.argmax(), d.shape)
            print(N,name,lab,"relmax",d.max()/np.abs(b).max(),"at",loc,"val",b[loc],"absmax",np.abs(b).max(), "at", np.unravel_index(np.abs(b).argmax(), b.shape), "interior relmax", d[8:-8,8:-8].max()/np.abs(b[8:-8,8:-8]).max())
