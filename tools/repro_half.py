"""Developer check: the gathered tensor-core kernel on small work lists (one stream), where it cuts 64-site units, for layer shapes with
one and several weight tiles, long K, 1x1 windows and a channel count that is not a multiple of 32.  Each case runs in its own process."""
import sys, os, subprocess
sys.path.insert(0, os.getcwd())
CASES = {
    "c96": "conv1=3,3,1,16 pool1=2,2 conv2=3,3,16,96 pool2=2,2 conv3=1,1,96,20",
    "c256": "conv1=3,3,1,16 pool1=2,2 conv2=3,3,16,256 pool2=2,2 conv3=1,1,256,20",
    "k1152": "conv1=3,3,1,16 pool1=2,2 conv2=3,3,16,128 conv3=3,3,128,128 conv4=1,1,128,20",
    "c110": "conv1=3,3,1,16 pool1=2,2 conv2=3,3,16,128 conv3=1,1,128,110",
    "c512": "conv1=3,3,1,16 pool1=2,2 conv2=3,3,16,128 conv3=1,1,128,512 conv4=1,1,512,110",
}
if len(sys.argv) > 1:
    import numpy as np
    import async_ev_cnn_b200 as P
    from async_ev_cnn_b200.engine import EventNetCuda
    layers = CASES[sys.argv[1]]
    wts = P.xavier_weights(layers, seed=0)
    net = EventNetCuda(32, 48, layers, wts, 5e-5, 0.1, "SAME", n_streams=1)
    ev = P.synthetic_events("uniform", 1, 3, 20, 32, 48, seed=1)
    for t in range(3):
        h = net.step([ev[0, t]])
    print(sys.argv[1], "ok", float(np.abs(h).sum()))
else:
    for k in CASES:
        r = subprocess.run([sys.executable, __file__, k], capture_output=True, text=True, timeout=120)
        print(k, "rc", r.returncode, (r.stdout + r.stderr).strip().splitlines()[-1][:200])
