"""Per-phase cycle table of the Sonic kernel (developer build with -DK4_TIMING: SPEEDY_K4_TIMING=1
SPEEDY_B200_BUILD_OUT=... python -m speedy_b200.build; run with SPEEDY_B200_LIB pointing at it).
Stream 0's leading lane accumulates clock64() deltas per phase; prints a table and writes JSON.
usage: k4_phases.py [streams] [seconds] [speed] [out.json]"""
import sys, json, ctypes as C, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, speedy_b200 as sb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
secs = int(sys.argv[2]) if len(sys.argv) > 2 else 60
speed = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
dst = sys.argv[4] if len(sys.argv) > 4 else None
rate = 16000
frames = rate * secs
d_in = torch.empty((n, frames, 1), dtype=torch.int16, device='cuda')
sb.synth_device(d_in, 0, n, rate, 1, frames)
b = sb.Batch(n, rate, 1, speed=speed, nonlinear=1.0, feedback=0.1, max_write_frames=frames, out_capacity=frames + 4096)
import os
L = sb.lib()
splice = bool(os.environ.get('SPEEDY_K4_PIPELINE'))
dbg = L.speedyDebugK4SpliceCycles if splice else L.speedyDebugK4Cycles
dbg.argtypes = [C.c_void_p, C.c_int]
for it in range(2):
    b.reset(); dbg(None, 1)
    b.write_device(d_in, frames, frames); b.flush_device(); torch.cuda.synchronize()
out = np.zeros(16, np.uint64); dbg(out.ctypes.data, 0)
names = (['wait_window', 'decimate', 'coarse_search', '-', 'fine_search', '-', 'post_record', '-'] if splice else
         ['refill', 'decimate', 'coarse_blocks', 'coarse_pick', 'fine_blocks', 'fine_pick', 'ola', 'copy'])
it = int(out[9]); tot = int(out[8])
table = {'kernel': 'k4_splice (chain warp)' if splice else 'k4_sonic', 'streams': n, 'seconds': secs, 'speed': speed, 'iterations_stream0': it, 'events_stream0': int(out[10]),
         'cycles_total_stream0': tot, 'cycles_per_iteration': tot / max(it, 1), 'phases': {}}
for i, nm in enumerate(names):
    table['phases'][nm] = {'cycles': int(out[i]), 'share': float(out[i]) / tot, 'per_iteration': float(out[i]) / max(it, 1)}
other = tot - int(out[:8].sum())
table['phases']['other'] = {'cycles': other, 'share': other / tot, 'per_iteration': other / max(it, 1)}
table['filler_waited_window_slot'] = int(out[5]); table['not_ready_chunks_short_sum'] = int(out[3]); table['not_ready_need_minus_released_sum'] = int(out[7]); table['filler_blocked_on_raw_slot'] = int(out[15]); table['window_checks'] = int(out[11]); table['window_not_ready'] = int(out[12]); table['queue_full_events'] = int(out[13]); table['mean_records_pending_at_check'] = float(out[14]) / max(int(out[11]), 1)
print(json.dumps(table, indent=1))
if dst:
    json.dump(table, open(dst, 'w'), indent=1)
