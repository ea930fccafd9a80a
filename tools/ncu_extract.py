"""Summarise `ncu --set full` reports (read on the CPU box): duration, tensor / DRAM / L2 / L1 utilisation, DRAM bytes,
occupancy and the warp-stall breakdown -> JSON.  Usage: python tools/ncu_extract.py out.json rep1.ncu-rep [rep2 ...]"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_throughput_pct",
    "lts__t_bytes.sum": "l2_bytes", "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_read_bytes",
    "l1tex__m_l1tex2xbar_write_bytes.sum": "sm_to_l2_write_bytes",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__grid_size": "grid", "launch__block_size": "block", "launch__registers_per_thread": "regs",
    "sm__cycles_elapsed.avg": "sm_cycles", "sm__cycles_active.avg": "sm_cycles_active",
    "smsp__inst_executed.sum": "warp_insts",
}


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * m.get(unit, 1)


def extract(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"report": path.split("/")[-1]}
        stalls = {}
        for h, u, v in zip(hdr, units, r):
            if h == "Kernel Name":
                d["kernel"] = v
            elif h in KEYS and v != "":
                v = v.replace(",", "")
                d[KEYS[h]] = to_bytes(v, u) if "byte" in u else float(v)
                if KEYS[h] == "duration_us" and u == "ms":
                    d[KEYS[h]] *= 1e3
                if KEYS[h] == "duration_us" and u == "ns":
                    d[KEYS[h]] /= 1e3
            elif h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v not in ("", "0"):
                stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(float(v), 3)
        d["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
        if "dram_read" in d:
            d["dram_bytes"] = d["dram_read"] + d["dram_write"]
        out.append(d)
    return out


if __name__ == "__main__":
    res = []
    for p in sys.argv[2:]:
        res += extract(p)
    json.dump(res, open(sys.argv[1], "w"), indent=1)
    for d in res:
        print(json.dumps(d))
