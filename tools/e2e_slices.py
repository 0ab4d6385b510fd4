import sys, os, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
from madronalib_b200 import api, workloads as wl
api.init(0)
V,T=65536,64
w=wl.config_a(V)
g=api.VoiceGraph(w.spec,V); g.set_coefs(w.coef); g.set_state(w.state)
h_in=torch.empty((T,1,V,64),dtype=torch.float32).pin_memory(); w.inputs(T,out=h_in.numpy())
h_out=torch.empty((T,1,V,64),dtype=torch.float32).pin_memory()
h_mix=torch.empty((T,1,64),dtype=torch.float32).pin_memory()
for ns in (8,16,24,32,48,64):
    os.environ["MLB_HOST_SLICES"]=str(ns)
    for _ in range(2): g.process_host(h_in.numpy(),T,want_out=True,want_mix=True,out=h_out.numpy(),mix=h_mix.numpy())
    t0=time.perf_counter()
    for _ in range(6): g.process_host(h_in.numpy(),T,want_out=True,want_mix=True,out=h_out.numpy(),mix=h_mix.numpy())
    dt=(time.perf_counter()-t0)/6
    print(ns,"slices: %.2f ms/step  %.3e voice-samples/s"%(dt*1e3, V*T*64/dt), flush=True)
