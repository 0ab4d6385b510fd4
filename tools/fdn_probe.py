import sys, json, numpy as np, torch
sys.path.insert(0,'/root/repo')
from madronalib_b200 import api, workloads as wl
from madronalib_b200.graph import GraphSpec
api.init(0)
dev=torch.device('cuda',0)
def run(name, w, T, reps=8):
    V=w.n_voices
    g=api.VoiceGraph(w.spec,V); g.set_coefs(w.coef); g.set_state(w.state)
    inp=w.inputs(T); d_in=torch.from_numpy(inp).to(dev); d_out=torch.empty((T,2,V,64),dtype=torch.float32,device=dev)
    sh=torch.cuda.current_stream().cuda_stream
    for _ in range(3): g.process_device(d_in,d_out,None,T,sh)
    ms=[]
    for _ in range(reps):
        g.process_device(d_in,d_out,None,T,sh); ms.append(g.last_kernel_ms())
    print(name, g.kernel_name, "V",V,"T",T,"ms %.4f"%np.mean(ms), "GB/s %.0f"%(4864.0*V*T/np.mean(ms)/1e6)); g.close()
run("full", wl.config_4(16384), 16)
run("one wave (592 CTAs)", wl.config_4(9472), 16)
run("half wave", wl.config_4(4736), 16)
run("two waves exactly", wl.config_4(18944), 16)
# external-input FDN (no generator)
def fdn_in(V):
    g=GraphSpec(); x=g.input(0); fl=g.node("FDN8",x); fr=g.node("FDN8_R",fl); g.output(fl,fr)
    w4=wl.config_4(V)
    fdn=w4.spec.ops.index(20)
    coef=np.ascontiguousarray(w4.coef[w4.spec.coef_slot(fdn):w4.spec.coef_slot(fdn)+32])
    w=wl.Workload("fdn_in",g,V,coef,g.new_state(V)); w.inputs=w4.inputs; return w
run("no generator, full", fdn_in(16384), 16)
run("no generator, one wave", fdn_in(9472), 16)
