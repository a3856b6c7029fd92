"""dev: resident step time (write_device + flush, as bench.py's value leg) under a list of
environment settings, one process.  usage: step_sweep.py [n] [secs] -- "A=1 B=2" "C=3" ..."""
import os, sys
sys.path.insert(0, '.')
import torch, speedy_b200 as sb
a = sys.argv[1:]
cut = a.index('--') if '--' in a else len(a)
n = int(a[0]) if cut > 0 else 1024
secs = int(a[1]) if cut > 1 else 60
speed = float(a[2]) if cut > 2 else 2.0
settings = a[cut + 1:] or ['']
rate, ch = 16000, 1
frames = rate * secs
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); stream = ts.cuda_stream
d_in = torch.empty((n, frames, ch), dtype=torch.int16, device='cuda')
sb.synth_device(d_in, 0, n, rate, ch, frames, stream=stream)
cap = frames + 4096
b = sb.Batch(n, rate, ch, speed=speed, nonlinear=1.0, feedback=0.1, max_write_frames=frames, out_capacity=cap)
def step():
    b.reset(stream); b.write_device(d_in, frames, frames, None, stream); b.flush_device(stream)
for s in settings:
    kv = dict(x.split('=') for x in s.split())
    for k, v in kv.items(): os.environ[k] = v
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 8
    e0.record()
    for _ in range(K): step()
    e1.record(); torch.cuda.synchronize()
    print("n=%d secs=%d speed=%g [%s]: %.3f ms/step" % (n, secs, speed, s, e0.elapsed_time(e1) / K), flush=True)
    for k in kv: os.environ.pop(k, None)
