import sys, os, time, torch, numpy as np
sys.path.insert(0,'.')
import speedy_b200 as sb
n, rate, secs = 1024, 16000, 60
frames = rate*secs; cap = frames + 4096
d_in = torch.empty((n, frames, 1), dtype=torch.int16, device='cuda'); sb.synth_device(d_in, 0, n, rate, 1, frames)
h_in = torch.empty((n, frames, 1), dtype=torch.int16, pin_memory=True); h_in.copy_(d_in)
h_out = torch.empty((n, cap, 1), dtype=torch.int16, pin_memory=True); h_cnt = torch.zeros(n, dtype=torch.int32)
b = sb.Batch(n, rate, 1, speed=2.0, nonlinear=1.0, feedback=0.1, max_write_frames=frames, out_capacity=cap)
ts = []
for i in range(8):
    t0 = time.perf_counter(); b.process_ptr(h_in, frames, h_out, cap, h_cnt); torch.cuda.synchronize(); ts.append((time.perf_counter()-t0)*1e3)
print('blocks', os.environ.get('SPEEDY_B200_SCATTER_BLOCKS'), 'chunk', os.environ.get('SPEEDY_B200_PROCESS_CHUNK'), 'ms', ' '.join('%.1f' % t for t in ts[2:]), 'min %.2f' % min(ts[2:]), 'sum', int(h_cnt.sum()))
