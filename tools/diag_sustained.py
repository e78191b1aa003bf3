"""Developer tool: sustained step rate and clocks over a few seconds of back-to-back steps."""
import sys, time, subprocess, threading, numpy as np, torch, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import async_ev_cnn_b200 as P
from async_ev_cnn_b200.engine import EventNetCuda
H,W,B,S=160,224,200,1024
net=EventNetCuda(H,W,P.EFCN_LAYERS,P.xavier_weights(P.EFCN_LAYERS,seed=0),5e-5,0.1,"SAME",n_streams=S,max_events_per_step=2048)
n=48+64
ev=P.synthetic_events("edge",S,n,B,H,W,seed=100)
ev=np.ascontiguousarray(ev.transpose(1,0,2,3)).reshape(n,S*B,3)
off=(np.arange(S+1,dtype=np.int64)*B).astype(np.int32)
evd=torch.from_numpy(ev).cuda(); offd=torch.from_numpy(off).cuda()
for t in range(48): net.step_device(evd[t].data_ptr(),offd.data_ptr(),S*B,None)
torch.cuda.synchronize()
rows=[]
p=subprocess.Popen(["nvidia-smi","--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown","--format=csv,noheader,nounits","-lms","50"],stdout=subprocess.PIPE,text=True)
def rd():
    for l in p.stdout: rows.append((time.perf_counter(),l.strip()))
threading.Thread(target=rd,daemon=True).start()
time.sleep(0.3)
t_start=time.perf_counter()
for rep in range(12):
    e=[torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    for i in range(64): net.step_device(evd[48+i].data_ptr(),offd.data_ptr(),S*B,None)
    e[1].record(); torch.cuda.synchronize()
    print("t=%.2fs  64 steps: %.3f ms/step"%(time.perf_counter()-t_start, e[0].elapsed_time(e[1])/64))
time.sleep(0.2); p.terminate()
for t,l in rows[::4]: print("%.2f %s"%(t-t_start,l))
