#!/usr/bin/env python
"""Summarises an `ncu --set full` report of the conv kernels of one training step (forward order then backward order):
per-launch duration, DRAM bytes, tensor-pipe / L2 / DRAM utilisation, registers; writes the per-launch DRAM traffic
that bench.py reports as roofline.traffic.
Usage: python tools/summarize_ncu_full.py report.ncu-rep out.txt traffic.json"""
import csv
import io
import json
import subprocess
import sys


def main():
    rep, out_txt, out_json = sys.argv[1:4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def find(name):   # some metrics carry a section prefix ("LTS.TriageCompute.lts__throughput...")
        if name in col:
            return name
        for h in hdr:
            if h.endswith("." + name):
                return h
        raise KeyError(name)

    def val(r, name, scale=1.0):
        name = find(name)
        v = r[col[name]].replace(",", "")
        u = units[col[name]]
        x = float(v) if v else float('nan')
        if name.startswith("dram__bytes_"):
            x *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        if name == "gpu__time_duration.sum":
            x *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
        return x * scale

    lines = ["# round 1 — ncu --set full --clock-control none of the %d conv3d fprop/dgrad launches of ONE training step"
             % len(data),
             "# (`python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-cuda-graph`, "
             "-k regex:conv3d_igemm_kernel|conv3d_slab_kernel -s 104 -c 26)",
             "# forward order (13 launches) then backward order (13 launches); times are cold-cache, serialised",
             "# tensor% = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "",
             "%-28s %9s %10s %10s %8s %7s %7s %5s" % ("kernel", "time_us", "dram_rd_MB", "dram_wr_MB", "tensor%",
                                                       "lts%", "dram%", "regs")]
    tot_t = tot_b = wt = 0.0
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("b2::", "")
        t = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        tp = val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        lts = val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed")
        dr = val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
        regs = int(float(r[col["launch__registers_per_thread"]]))
        lines.append("%-28s %9.1f %10.1f %10.1f %8.2f %7.1f %7.1f %5d" % (name[:28], t, rd / 1e6, wr / 1e6, tp, lts, dr,
                                                                          regs))
        tot_t += t
        tot_b += rd + wr
        wt += tp * t
    lines += ["", "total: %.1f us, DRAM traffic %.1f MB over %d launches = %.1f MB per launch; time-weighted tensor-pipe "
                  "active %.1f %%" % (tot_t, tot_b / 1e6, len(data), tot_b / 1e6 / len(data), wt / tot_t)]
    open(out_txt, "w").write("\n".join(lines) + "\n")
    json.dump({"kernel": "conv3d_igemm_kernel + conv3d_slab_kernel (b2_conv3d_igemm entry point), %d launches per "
                         "training step" % len(data),
               "dram_bytes_per_launch": tot_b / len(data), "launches": len(data),
               "time_weighted_tensor_pipe_active_pct": wt / tot_t, "source": out_txt + " (ncu --set full, round 1)"},
              open(out_json, "w"), indent=1)
    print("\n".join(lines[-3:]))


if __name__ == "__main__":
    main()
