"""Turn the ncu artefacts a gpurun call brought back into the committed summaries.

    python profiles/summarize.py r01 gpurun_out/prof_r1_final.ncu-rep gpurun_out/launches_r1.csv \
        gpurun_out/bench_r1.json gpurun_out/bench_r1_ref.json

Writes profiles/<round>_kernels.json (per-kernel metrics from the --set full capture),
profiles/<round>_launches.csv (the launch list: kernel, duration) and
profiles/<round>_launch_shares.json (each kernel's share of the step), and copies the
bench JSON lines next to them.
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
METRICS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct": "issue_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__block_size": "block_size",
    "launch__grid_size": "grid_size",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
    "launch__occupancy_limit_registers": "occupancy_limit_registers_blocks",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_smem_blocks",
    "launch__waves_per_multiprocessor": "waves_per_sm",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
}


def main():
    tag, rep, launches, *bench = sys.argv[1:]
    command = os.environ.get("SUMMARIZE_COMMAND") or (
        "SPEEDY_B200_WRITE_PARTS=1 ncu --set full --clock-control none --import-source on "
        "-k regex:'k4_sonic|k1_dft16|k2_tension' --launch-skip 6 --launch-count 4 python bench.py --steps 1 --warmup 3")
    # (a report, or the `ncu -i report --page raw --csv` export of one made on the GPU box: the reports
    # themselves can be too large to bring back)
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    kernels = []
    for r in rows[2:]:
        k = {"kernel": r[ix["Kernel Name"]]}
        for m, name in METRICS.items():
            if m in ix:
                try:
                    k[name] = float(r[ix[m]].replace(",", ""))
                except ValueError:
                    k[name] = r[ix[m]]
                k[name + "_unit"] = units[ix[m]]
        kernels.append(k)
    json.dump({"source": os.path.basename(rep), "command": command, "kernels": kernels},
              open(os.path.join(HERE, tag + os.environ.get("SUMMARIZE_SUFFIX", "_kernels") + ".json"), "w"), indent=1)
    if launches == "-":  # a kernel capture only
        print(json.dumps([{k: v for k, v in kk.items() if not k.endswith("_unit")} for kk in kernels], indent=1)[:3000])
        return
    # launch list
    lines = [l for l in open(launches) if not l.startswith("==")]
    rows = list(csv.reader(io.StringIO("".join(lines))))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    out, share = [], {}
    for r in rows[1:]:
        if len(r) <= ix["Metric Value"]:
            continue
        name = r[ix["Kernel Name"]]
        val = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        ms = val / 1e6 if unit in ("ns", "nsecond") else (val / 1e3 if unit in ("us", "usecond") else val)
        out.append((r[ix["ID"]], name, ms))
        share[name] = share.get(name, 0.0) + ms
    with open(os.path.join(HERE, tag + "_launches.csv"), "w") as f:
        f.write("id,kernel,duration_ms\n")
        for i, name, ms in out:
            f.write('%s,"%s",%.6f\n' % (i, name, ms))
    total = sum(share.values())
    json.dump({"command": "ncu --metrics gpu__time_duration.sum --clock-control none -c 600 python bench.py "
               "--steps 2 --warmup 3 (cold-cache, serialised: compare shares, not absolutes)",
               "total_ms": total,
               "share": {k: {"ms": v, "frac": v / total} for k, v in sorted(share.items(), key=lambda kv: -kv[1])}},
              open(os.path.join(HERE, tag + "_launch_shares.json"), "w"), indent=1)
    for b in bench:
        shutil.copy(b, os.path.join(HERE, tag + "_" + os.path.basename(b)))
    print(json.dumps([{k: v for k, v in kk.items() if not k.endswith("_unit")} for kk in kernels], indent=1)[:3000])
    print({k: round(v["frac"], 3) for k, v in json.load(open(os.path.join(HERE, tag + "_launch_shares.json")))["share"].items()})


if __name__ == "__main__":
    main()
