"""Turn the .ncu-rep captures a profile script left in gpurun_out/ into the tracked evidence under profiles/.

    python tools/ncu_summarise.py r02b

For every gpurun_out/<tag>_*.ncu-rep: <name>.raw.csv (ncu --page raw --csv), <name>.details.txt (--page details),
one entry in profiles/<tag>_ncu_summary.json (the metrics DESIGN.md quotes) and — for the scan / finish kernels — the
DRAM traffic per launch in profiles/r02_traffic.json, which bench.py reports as roofline.traffic.
"""
import csv
import glob
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size", "lts__t_sector_hit_rate.pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "smsp__inst_executed.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def traffic_key(name: str):
    """r02b_scan_exact_b1024_shard125k -> ('scan_mma_bf16_kernel', 'rows125000_b1024_clip')"""
    kern = "exact_finish_kernel" if "exact_finish" in name else "scan_mma_bf16_kernel"
    m = re.search(r"_b(\d+)_(clip|gauss|shard(\d+)k)", name)
    if not m:
        return None
    b = int(m.group(1))
    rows = int(m.group(3)) * 1000 if m.group(3) else 1000000
    data = m.group(2) if m.group(2) in ("clip", "gauss") else "clip"
    bk = "ble128" if b <= 128 else "b%d" % b
    return kern, "rows%d_%s_%s" % (rows, bk, data)


def main(tag: str) -> None:
    out_dir = os.path.join(ROOT, "profiles")
    summary = {}
    tp = os.path.join(out_dir, "r02_traffic.json")
    traffic = json.load(open(tp)) if os.path.exists(tp) else {}
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", tag + "_*.ncu-rep"))):
        name = os.path.basename(rep)[:-len(".ncu-rep")]
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        open(os.path.join(out_dir, name + ".raw.csv"), "w").write(raw)
        det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
        open(os.path.join(out_dir, name + ".details.txt"), "w").write(det)
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        ent = {}
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP:
                ent[h] = [v, u]
            if h == "Kernel Name":
                ent["kernel"] = v[:90]
        rd, wr = ent.get("dram__bytes_read.sum"), ent.get("dram__bytes_write.sum")
        if rd and wr:
            ent["traffic_bytes"] = float(rd[0].replace(",", "")) * UNIT[rd[1]] + float(wr[0].replace(",", "")) * UNIT[wr[1]]
            key = traffic_key(name)
            if key:
                traffic.setdefault(key[0], {})[key[1]] = ent["traffic_bytes"]
                if key[1].startswith("rows1000000_ble128"):
                    traffic[key[0]]["rows1000000_ble128_gauss"] = ent["traffic_bytes"]
        summary[name] = ent
        print(name, ent.get("gpu__time_duration.sum"), ent.get("traffic_bytes"),
              ent.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"))
    json.dump(summary, open(os.path.join(out_dir, tag + "_ncu_summary.json"), "w"), indent=1)
    traffic["source"] = ("dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu --set full captures "
                         "(profiles/r02b_*.raw.csv; r02_* for keys not re-captured), per launch")
    json.dump(traffic, open(tp, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02b")
