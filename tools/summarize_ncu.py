#!/usr/bin/env python3
"""Turns ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_X.csv profiles/NAME.md "<command>"
  python tools/summarize_ncu.py kernel   gpurun_out/prof_X.ncu-rep  profiles/NAME.md "<command>" [frames_per_launch]
  python tools/summarize_ncu.py multi    gpurun_out/prof_X.ncu-rep  profiles/NAME.md "<command>" frames_per_step
      every kernel of one step: launches of the same kernel are summed; writes profiles/<round>_<kernel>_traffic.json per
      kernel and profiles/<round>_int_ops.json (executed thread instructions per frame, for bench.py's int_roofline)
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed.sum", "launch__block_size", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__cycles_active.avg",
]


def launches(src, dst, cmd):
    rows = list(csv.reader(open(src)))
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                v = float(d["Metric Value"].replace(",", ""))
                v = v / 1e3 if d["Metric Unit"] in ("ns", "nsecond") else v
                k = d["Kernel Name"].split("(")[0]
                agg[k][0] += 1
                agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list\n\nCommand (B200, after the same command exited 0 without ncu):\n`{cmd}`\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:80]}` | {v[0]} | {v[1]:.1f} | {v[1] / tot:.4f} |\n")


def kernel(src, dst, cmd, frames):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {h: (r[i], units[i]) for i, h in enumerate(hdr)}
        out.append(d)
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary\n\nCommand (B200): `{cmd}`\n\n")
        for d in out:
            f.write(f"## `{d['Kernel Name'][0][:100]}`\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k][0]} | {d[k][1]} |\n")
            for k, (v, u) in sorted(d.items()):
                if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                    try:
                        if float(v) >= 0.2:
                            f.write(f"| {k} | {v} | {u} |\n")
                    except ValueError:
                        pass
            f.write("\n")
        if frames and out:
            d = out[0]

            def by(x):
                v, u = d[x]
                return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            traffic = by("dram__bytes_read.sum") + by("dram__bytes_write.sum")
            json.dump({"kernel": d["Kernel Name"][0].split("(")[0], "frames_per_launch": frames,
                       "dram_bytes_per_launch": traffic, "dram_bytes_per_frame": traffic / frames, "source": dst},
                      open(dst.replace(".md", "_traffic.json"), "w"), indent=1)


def multi(src, dst, cmd, frames):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3, "msecond": 1.0, "usecond": 1e-3,
             "nsecond": 1e-6, "second": 1e3}
    agg = collections.OrderedDict()
    rnd = dst.split("/")[-1].split("_")[0]
    for r in rows[2:]:
        d = {h: (r[i], units[i]) for i, h in enumerate(hdr)}
        name = d["Kernel Name"][0].split("(")[0].replace("void ", "").split("<")[0]

        def num(k, d=d):
            v, u = d[k]
            return float(v.replace(",", "")) * scale.get(u, 1)
        a = agg.setdefault(name, {"launches": 0, "ms": 0.0, "dram": 0.0, "winst": 0.0, "tinst": 0.0, "first": d})
        a["launches"] += 1
        a["ms"] += num("gpu__time_duration.sum")
        a["dram"] += num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        a["winst"] += num("smsp__inst_executed.sum")
        if "smsp__thread_inst_executed.sum" in d:
            a["tinst"] += num("smsp__thread_inst_executed.sum")
        else:   # warp instructions x average active threads per instruction
            a["tinst"] += num("smsp__inst_executed.sum") * num("smsp__thread_inst_executed_per_inst_executed.ratio")
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary, one step of {frames} frames\n\nCommand (B200, after the same command exited 0 without ncu): `{cmd}`\n\n")
        f.write("Times under ncu are cold-cache and serialised (compare shares).  Launches of one kernel are summed.\n\n")
        f.write("| kernel | launches | ms | DRAM MB/frame | warp inst/frame | thread inst/frame | issue active % | warps active % | regs |\n|---|---|---|---|---|---|---|---|---|\n")
        ops = {}
        for name, a in agg.items():
            d = a["first"]
            f.write(f"| `{name}` | {a['launches']} | {a['ms']:.3f} | {a['dram'] / frames / 1e6:.3f} | {a['winst'] / frames:.0f} | {a['tinst'] / frames:.0f} | "
                    f"{d['smsp__issue_active.avg.pct_of_peak_sustained_active'][0]} | {d['sm__warps_active.avg.pct_of_peak_sustained_active'][0]} | "
                    f"{d['launch__registers_per_thread'][0]} |\n")
            json.dump({"kernel": name, "frames_per_launch": frames, "dram_bytes_per_launch": a["dram"], "dram_bytes_per_frame": a["dram"] / frames,
                       "source": dst}, open(dst.rsplit("/", 1)[0] + f"/{rnd}_{name}_traffic.json", "w"), indent=1)
            ops[name] = {"thread_inst_per_frame": a["tinst"] / frames, "warp_inst_per_frame": a["winst"] / frames}
        f.write("\nFirst launch of every kernel:\n\n")
        for name, a in agg.items():
            d = a["first"]
            f.write(f"## `{d['Kernel Name'][0][:100]}`\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k][0]} | {d[k][1]} |\n")
            for k, (v, u) in sorted(d.items()):
                if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                    try:
                        if float(v) >= 0.2:
                            f.write(f"| {k} | {v} | {u} |\n")
                    except ValueError:
                        pass
            f.write("\n")
    return ops


if __name__ == "__main__":
    mode, src, dst, cmd = sys.argv[1:5]
    if mode == "launches":
        launches(src, dst, cmd)
    elif mode == "multi":
        ops = multi(src, dst, cmd, int(sys.argv[5]))
        ij = dst.rsplit("/", 1)[0] + "/" + dst.split("/")[-1].split("_")[0] + "_int_ops.json"
        try:
            old = json.load(open(ij))
        except Exception:
            old = {"kernels": {}}
        old["kernels"].update(ops)
        old["note"] = "executed thread / warp instructions per frame from ncu --set full (smsp__thread_inst_executed.sum, smsp__inst_executed.sum)"
        json.dump(old, open(ij, "w"), indent=1)
    else:
        kernel(src, dst, cmd, int(sys.argv[5]) if len(sys.argv) > 5 else 0)
