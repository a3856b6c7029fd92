"""dev: time the kernels of one write+flush (parts=1) and checksum the output.
usage: [SPEEDY_B200_LIB=...] python profiles/tools/kernel_times.py [n] [secs] [rate] [ch] [speed]"""
import os, sys
os.environ.setdefault("SPEEDY_B200_WRITE_PARTS", "1")
sys.path.insert(0, '.')
import numpy as np, torch, speedy_b200 as sb
a = sys.argv[1:]
n = int(a[0]) if len(a) > 0 else 1024
secs = int(a[1]) if len(a) > 1 else 60
rate = int(a[2]) if len(a) > 2 else 16000
ch = int(a[3]) if len(a) > 3 else 1
speed = float(a[4]) if len(a) > 4 else 2.0
frames = rate * secs
d_in = torch.empty((n, frames, ch), dtype=torch.int16, device='cuda')
sb.synth_device(d_in, 0, n, rate, ch, frames)
cap = frames + 4096
b = sb.Batch(n, rate, ch, speed=speed, nonlinear=1.0, feedback=0.1, max_write_frames=frames, out_capacity=cap)
d_out = torch.zeros((n, cap, ch), dtype=torch.int16, device='cuda')
d_cnt = torch.zeros(n, dtype=torch.int32, device='cuda')
best = None
for it in range(4):
    b.reset(); b.set_profiling(True)
    b.write_device(d_in, frames, frames); b.flush_device(); torch.cuda.synchronize()
    kt = b.kernel_times(); b.set_profiling(False)
    if best is None or kt['sonic'] < best['sonic']: best = kt
b.read_device(d_out, cap, d_cnt); torch.cuda.synchronize()
cnt = d_cnt.cpu().numpy().astype(np.int64)
o = d_out.cpu().numpy().astype(np.int64)
w = (np.arange(o.shape[1], dtype=np.int64) % 8191 + 1)[None, :, None]
print("lib=%s thr=%s n=%d rate=%d ch=%d speed=%g | %s | counts_sum=%d checksum=%d" % (
    os.path.basename(os.environ.get("SPEEDY_B200_LIB", "default")), os.environ.get("SPEEDY_K4_THREADS", "-"), n, rate, ch, speed,
    " ".join("%s=%.3f" % kv for kv in best.items()), cnt.sum(), int((o * w).sum())))
