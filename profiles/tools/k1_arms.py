"""K1 decision evidence: the spectrogram stage three ways, same box.
  ours     : k1_spectral_* (fused int16 -> |X| -> energy + log-domain spectral difference; nothing
             but 8 bytes per window goes back to HBM), time from speedyBatch profiling
  cufft    : cuFFT batched R2C (torch.fft.rfft on an already zero-padded fp32 [windows, N] matrix)
             + abs: the library bar for the transform ALONE (no int16 conversion, pre-emphasis,
             Hamming, energy or spectral difference, which would be two more passes over HBM)
  gemm     : the DFT as a tensor-core GEMM through cuBLAS: with the window folded about its centre
             (s = v(t) + v(-t), d = v(t) - v(-t)) the real part is S.C and the imaginary part D'.C with
             the same cosine matrix read backwards, so [S; D'] (2 rows per window, K = W/2) x C
             (K x N/2+1); fp16 operands split hi + lo: three products.  Library GEMM alone, operands
             already split and resident, output written to HBM.
usage: k1_arms.py [out.json]"""
import json, os, sys, time
os.environ.setdefault("SPEEDY_B200_WRITE_PARTS", "1")
sys.path.insert(0, '.')
import torch, speedy_b200 as sb

def ev_time(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

def ours(rate, n, secs):
    frames = rate * secs
    d_in = torch.empty((n, frames, 1), dtype=torch.int16, device='cuda')
    sb.synth_device(d_in, 0, n, rate, 1, frames)
    b = sb.Batch(n, rate, 1, speed=2.0, nonlinear=1.0, feedback=0.1, max_write_frames=frames, out_capacity=frames + 4096)
    best = 1e9
    for _ in range(3):
        b.reset(); b.set_profiling(True); b.write_device(d_in, frames, frames); torch.cuda.synchronize()
        best = min(best, b.kernel_times()['spectral']); b.set_profiling(False)
    b.close(); del d_in; torch.cuda.empty_cache()
    return best

def cufft(windows, N):
    chunk = min(windows, 1 << 20)
    x = torch.randn((chunk, N), device='cuda', dtype=torch.float32)
    t = ev_time(lambda: torch.fft.rfft(x, dim=1).abs())
    t_fft = ev_time(lambda: torch.fft.rfft(x, dim=1))
    del x; torch.cuda.empty_cache()
    return t * windows / chunk, t_fft * windows / chunk

def gemm(windows, W, N):
    K = (W // 2 + 15) // 16 * 16
    cols = (N // 2 + 1 + 15) // 16 * 16
    chunk = min(windows, 1 << 20)
    a = torch.randn((2 * chunk, K), device='cuda', dtype=torch.float16)
    bm = torch.randn((K, cols), device='cuda', dtype=torch.float16)
    out = torch.empty((2 * chunk, cols), device='cuda', dtype=torch.float16)
    t = ev_time(lambda: (torch.mm(a, bm, out=out), torch.mm(a, bm, out=out), torch.mm(a, bm, out=out)))
    del a, bm, out; torch.cuda.empty_cache()
    flop = 3 * 2.0 * 2 * windows * K * cols
    return t * windows / chunk, flop, K, cols

res = {"box": torch.cuda.get_device_name(0), "arms": []}
for rate, secs in ((16000, 60), (22050, 10), (48000, 10)):
    n = 1024
    W, N, S = sb.frame_geometry(rate)
    windows = n * ((rate * secs - W - 1) // S + 1)
    t_ours = ours(rate, n, secs)
    t_cufft_abs, t_cufft = cufft(windows, N)
    t_gemm, flop, K, cols = gemm(windows, W, N)
    row = {"rate": rate, "fft": N, "window": W, "windows": windows, "streams": n, "seconds": secs,
           "ours_fused_ms": t_ours, "cufft_r2c_ms": t_cufft, "cufft_r2c_plus_abs_ms": t_cufft_abs,
           "cublas_fp16_3x_gemm_ms": t_gemm, "gemm_shape": [2 * windows, K, cols], "gemm_tflop": flop / 1e12,
           "gemm_tflops_achieved": flop / 1e12 / (t_gemm / 1e3),
           "hbm_bytes_cufft_arm_gb": windows * (N * 4 + (N // 2 + 1) * 8) / 1e9,
           "hbm_bytes_ours_gb": (n * rate * secs * 2 + windows * 8) / 1e9}
    print(row, flush=True)
    res["arms"].append(row)
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
