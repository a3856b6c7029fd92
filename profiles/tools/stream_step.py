"""dev: the 10 ms streaming step (config 5 shape) with and without CUDA-graph replay.
usage: stream_step.py [sessions] [graph: 0|1]"""
import os, sys, time
sys.path.insert(0, '.')
if len(sys.argv) > 2 and sys.argv[2] == '0':
    os.environ['SPEEDY_B200_NO_GRAPH'] = '1'
import numpy as np, torch, speedy_b200 as sb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
rate, chunk, secs = 16000, 160, 6
frames = rate * secs
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); stream = ts.cuda_stream
d_in = torch.empty((n, frames, 1), dtype=torch.int16, device='cuda')
sb.synth_device(d_in, 0, n, rate, 1, frames, stream=stream)
cap = 8192
b = sb.Batch(n, rate, 1, speed=2.5, nonlinear=1.0, feedback=0.1, max_write_frames=chunk, out_capacity=cap)
d_out = torch.empty((n, cap, 1), dtype=torch.int16, device='cuda'); d_cnt = torch.zeros(n, dtype=torch.int32, device='cuda')
stage = torch.empty((n, chunk, 1), dtype=torch.int16, device='cuda')
def run(staged, K=400, warm=100):
    b.reset(stream); lat = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(warm + K):
        if k == warm: torch.cuda.synchronize(); e0.record()
        t0 = time.perf_counter()
        if staged:
            stage.copy_(d_in[:, k * chunk:(k + 1) * chunk]); b.write_device(stage, chunk, chunk, None, stream)
        else:
            b.write_device(d_in[:, k * chunk:], frames, chunk, None, stream)
        b.read_device(d_out, cap, d_cnt, stream)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K
print("n=%d graph=%s: strided input %.3f ms/step, staged input %.3f ms/step" % (n, os.environ.get('SPEEDY_B200_NO_GRAPH') is None, run(False), run(True)))
h_in = torch.empty((n, frames, 1), dtype=torch.int16, pin_memory=True); h_in.copy_(d_in)
h_out = torch.empty((n, cap, 1), dtype=torch.int16, pin_memory=True); h_cnt = torch.zeros(n, dtype=torch.int32)
b.reset(stream); torch.cuda.synchronize(); lat = []
for k in range(500):
    t0 = time.perf_counter(); b.write_ptr(h_in, frames, chunk, k * chunk); b.read_ptr(h_out, cap, h_cnt)
    if k >= 100: lat.append((time.perf_counter() - t0) * 1e3)
lat.sort(); print("host chunks: p50 %.3f p99 %.3f ms" % (lat[len(lat) // 2], lat[int(len(lat) * 0.99)]))
