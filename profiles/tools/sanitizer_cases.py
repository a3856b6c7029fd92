"""small end-to-end runs for compute-sanitizer: a few streams, several rates, write + flush + read"""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import oracle_lib as ol
from gpu_util import gpu_process
for rate, ch, speed in [(16000, 1, 2.0), (48000, 2, 1.5), (22050, 1, 3.5), (16000, 2, 0.7)]:
    pcm = ol.synth(7, 2, rate, ch, rate)  # 1 s
    outs, taps, status = gpu_process(pcm, rate, speed)
    print(rate, ch, speed, [len(o) for o in outs], status)
