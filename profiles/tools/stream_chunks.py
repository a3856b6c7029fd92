"""dev: config 5 shape -- many sessions fed 10 ms chunks; ms per chunk step."""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch, speedy_b200 as sb
a = sys.argv[1:]
n = int(a[0]) if len(a) > 0 else 16384
chunks = int(a[1]) if len(a) > 1 else 200
rate, ch = 16000, 1
cf = rate // 100
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); stream = ts.cuda_stream
frames = cf * chunks
d_in = torch.empty((n, frames, ch), dtype=torch.int16, device='cuda')
sb.synth_device(d_in, 0, n, rate, ch, frames, stream=stream)
b = sb.Batch(n, rate, ch, speed=2.5, nonlinear=1.0, feedback=0.1, max_write_frames=cf, out_capacity=frames + 4096)
base = d_in.data_ptr()
def run():
    b.reset(stream)
    for c in range(chunks):
        b.write_device(base + 2 * ch * cf * c, frames, cf, None, stream)
for _ in range(2): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
t0 = time.time(); e0.record(); run(); e1.record(); torch.cuda.synchronize(); t1 = time.time()
ms = e0.elapsed_time(e1)
print("sessions=%d chunks=%d: %.3f ms per 10 ms chunk step (device), %.3f ms wall; real-time factor %.0f audio-s/s" % (n, chunks, ms / chunks, (t1 - t0) * 1e3 / chunks, n * chunks * 0.01 / (ms * 1e-3)))
b.set_profiling(True); run(); torch.cuda.synchronize(); kt = b.kernel_times(); b.set_profiling(False)
print({k: round(v / chunks, 4) for k, v in kt.items()}, "ms per chunk step")
