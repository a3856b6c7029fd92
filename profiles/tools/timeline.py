import os, sys
sys.path.insert(0, '.')
import numpy as np, torch, speedy_b200 as sb
n, rate, secs = 1024, 16000, 60
frames = rate * secs
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); stream = ts.cuda_stream
d_in = torch.empty((n, frames, 1), dtype=torch.int16, device='cuda')
sb.synth_device(d_in, 0, n, rate, 1, frames, stream=stream)
b = sb.Batch(n, rate, 1, speed=2.0, nonlinear=1.0, feedback=0.1, max_write_frames=frames, out_capacity=frames + 4096)
for it in range(3):
    b.reset(stream); b.set_profiling(True)
    b.write_device(d_in, frames, frames, None, stream); b.flush_device(stream); torch.cuda.synchronize()
    if it == 2: os.environ['SPEEDY_B200_PROF_DUMP'] = '1'
    kt = b.kernel_times(); b.set_profiling(False)
print(kt)
