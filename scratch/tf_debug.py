import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, signals
from oracle.hdrref import CpuStretch
from signalsmith_stretch_b200 import BatchStretch
def rms(a): return float(np.sqrt(np.mean(np.square(a,dtype=np.float64))))
for name in ("config1_12st_44k","config3_7st_ton8k"):
    cfg,C,sr,ratio,kind=signals.CONFIGS[name]
    e=BatchStretch(1); o=CpuStretch('orc'); cfg(e); cfg(o)
    H,B=e.intervalSamples(),e.blockSamples()
    n_calls,co=14,2*H; ci=int(round(co/ratio))
    x=signals.batch(kind,1,C,ci*n_calls,sr)
    errs=[]
    for k in range(n_calls):
        st=o.signal_state()
        for key in ("history","pending","pendingWp","input","prevInput","output","predEnergy"): e.set_state(key, st[key][None])
        xin=x[:,:,k*ci:(k+1)*ci]
        yo=o.process(xin[0],co); yg=np.asarray(e.process(xin,co))[0]
        errs.append(rms(yg-yo))
        # spectral state comparison after the call
        so=o.signal_state(); 
        d_out=np.abs(e.get_state("output")[0]-so["output"]); 
        errs[-1]=(errs[-1], float(d_out.max()), int(d_out.argmax()), float(np.abs(so["output"]).max()))
    print(name); 
    for k,v in enumerate(errs): print("  call %2d rms %.2e  max|dOut| %.2e at bin %d (|out|max %.2e)"%(k,*v))
