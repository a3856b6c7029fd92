"""dev: ms per resident step (write_device + flush) as bench.py's value leg does it."""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch, speedy_b200 as sb
a = sys.argv[1:]
n = int(a[0]) if len(a) > 0 else 1024
secs = int(a[1]) if len(a) > 1 else 60
rate = int(a[2]) if len(a) > 2 else 16000
ch = int(a[3]) if len(a) > 3 else 1
frames = rate * secs
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); stream = ts.cuda_stream
d_in = torch.empty((n, frames, ch), dtype=torch.int16, device='cuda')
sb.synth_device(d_in, 0, n, rate, ch, frames, stream=stream)
cap = frames + 4096
b = sb.Batch(n, rate, ch, speed=2.0, nonlinear=1.0, feedback=0.1, max_write_frames=frames, out_capacity=cap)
def step():
    b.reset(stream); b.write_device(d_in, frames, frames, None, stream); b.flush_device(stream)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 8
e0.record()
for _ in range(K): step()
e1.record(); torch.cuda.synchronize()
print("parts=%s buf=%s n=%d secs=%d: %.3f ms/step" % (os.environ.get("SPEEDY_B200_WRITE_PARTS", "-"), os.environ.get("SPEEDY_K4_BUF", "-"), n, secs, e0.elapsed_time(e1) / K))
