import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import async_ev_cnn_b200 as P
from async_ev_cnn_b200.engine import EventNetCuda
H,W,B,S=160,224,200,1024
net=EventNetCuda(H,W,P.EFCN_LAYERS,P.xavier_weights(P.EFCN_LAYERS,seed=0),5e-5,0.1,"SAME",n_streams=S,max_events_per_step=2048)
n=48+60
ev=P.synthetic_events("edge",S,n,B,H,W,seed=100)
ev=np.ascontiguousarray(ev.transpose(1,0,2,3)).reshape(n,S*B,3)
off=(np.arange(S+1,dtype=np.int64)*B).astype(np.int32)
evd=torch.from_numpy(ev).cuda(); offd=torch.from_numpy(off).cuda()
for t in range(48): net.step_device(evd[t].data_ptr(),offd.data_ptr(),S*B,None)
torch.cuda.synchronize()
evh=torch.from_numpy(ev[48:]).pin_memory().numpy(); offh=torch.from_numpy(off).pin_memory().numpy()
hb=[torch.empty((S,)+net.head_shape,dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
def timeit(name, fn, k=10):
    fn(0); torch.cuda.synchronize(); net.host_sync()
    t0=time.perf_counter()
    for i in range(1,k+1): fn(i)
    net.host_sync(); torch.cuda.synchronize()
    print("%-40s %.3f ms/step"%(name,(time.perf_counter()-t0)*1e3/k))
timeit("device only", lambda i: net.step_device(evd[40+i].data_ptr(),offd.data_ptr(),S*B,None))
timeit("blocking host", lambda i: net.step_packed(evh[i],offh,out=hb[0]))
timeit("async host", lambda i: net.step_packed_async(evh[i],offh,hb[i&1]))
timeit("async host, stream", lambda i: net.step_packed_async(evh[i],offh,hb[i&1],cuda_stream=torch.cuda.current_stream().cuda_stream))
s2=torch.cuda.Stream()
timeit("async host, side stream", lambda i: net.step_packed_async(evh[i],offh,hb[i&1],cuda_stream=s2.cuda_stream))
t0=time.perf_counter()
for i in range(11,21): net.step_packed_async(evh[i],offh,hb[i&1])
t1=time.perf_counter(); net.host_sync(); t2=time.perf_counter()
print("enqueue %.3f ms/step, drain %.3f ms"%((t1-t0)*100,(t2-t1)*1e3))
import ctypes
from async_ev_cnn_b200 import _native as NN
def nod2h(i):
    NN.check(net._lib.aec_net_step_host_async(net._h, ctypes.c_void_p(evh[i].ctypes.data), ctypes.c_void_p(offh.ctypes.data), int(offh[-1]), None, None))
timeit("async host, no D2H", nod2h)
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for i in range(21,31): net.step_packed_async(evh[i],offh,hb[i&1])
e1.record(); net.host_sync(); torch.cuda.synchronize()
print("async: GPU time on the compute stream %.3f ms/step"%(e0.elapsed_time(e1)/10))
def per_step(name, fn, k=12):
    evs=[torch.cuda.Event(enable_timing=True) for _ in range(k+1)]
    net.host_sync(); torch.cuda.synchronize()
    evs[0].record()
    for i in range(k):
        fn(31+i); evs[i+1].record()
    net.host_sync(); torch.cuda.synchronize()
    print(name, " ".join("%.2f"%evs[i].elapsed_time(evs[i+1]) for i in range(k)))
per_step("device-only per-step ms:", lambda i: net.step_device(evd[40+i].data_ptr(),offd.data_ptr(),S*B,None))
per_step("async per-step ms:      ", lambda i: net.step_packed_async(evh[i],offh,hb[i&1]))
per_step("async no-D2H per-step:  ", nod2h)
