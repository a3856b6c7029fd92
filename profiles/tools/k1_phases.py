"""Per-phase cycle table of the tensor-core spectrogram kernel (developer build with -DK1_TIMING:
SPEEDY_B200_EXTRA_FLAGS=-DK1_TIMING SPEEDY_B200_BUILD_OUT=... python -m speedy_b200.build).
usage: SPEEDY_B200_LIB=<that build> python profiles/tools/k1_phases.py [out.json]"""
import os, sys, json, ctypes as C, numpy as np
os.environ.setdefault("SPEEDY_B200_WRITE_PARTS", "1")
sys.path.insert(0, '.')
import torch, speedy_b200 as sb
n, secs, rate = 1024, 60, 16000
frames = rate * secs
d_in = torch.empty((n, frames, 1), dtype=torch.int16, device='cuda')
sb.synth_device(d_in, 0, n, rate, 1, frames)
b = sb.Batch(n, rate, 1, speed=2.0, nonlinear=1.0, feedback=0.1, max_write_frames=frames, out_capacity=frames + 4096)
L = sb.lib(); L.speedyDebugK1Cycles.argtypes = [C.c_void_p, C.c_int]
for it in range(2):
    b.reset(); L.speedyDebugK1Cycles(None, 1); b.write_device(d_in, frames, frames); torch.cuda.synchronize()
out = np.zeros(16, np.uint64); L.speedyDebugK1Cycles(out.ctypes.data, 0)
tiles = int(out[15])
names = {1: "rows: barrier (previous samples free, slot table visible)", 2: "rows: samples (bulk wait or staging)",
         3: "rows: units (pre-emphasis .. split, wait chunk free, stores, arrive)", 5: "issuer: wait accumulator free",
         6: "issuer: wait chunks + 24 MMAs + commits", 8: "epilogue: wait accumulator full", 9: "epilogue: pass 1 (tcgen05.ld, power, log2)",
         10: "epilogue: energy exchange", 11: "epilogue: pass 2 + exchange + store"}
table = {"tiles_cta0": tiles, "phases": {names[i]: float(out[i]) / max(tiles, 1) for i in names}}
print(json.dumps(table, indent=1))
if len(sys.argv) > 1: json.dump(table, open(sys.argv[1], "w"), indent=1)
