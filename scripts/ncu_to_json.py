"""Condenses ncu --set full captures (one launch each) into the JSON bench.py embeds in its `roofline` object, so the
quoted DRAM traffic / issue-slot figures always come from the build they are committed with (scripts/gpu_round.sh writes
profiles/<tag>_traffic.json from the captures of the same pass).

  python scripts/ncu_to_json.py out.json name=path.ncu-rep [name=path.ncu-rep ...]"""
import csv
import json
import subprocess
import sys

M = {"gpu_time_us": ("gpu__time_duration.sum", 1e-3),               # ns -> us
     "dram_read_bytes": ("dram__bytes_read.sum", None), "dram_write_bytes": ("dram__bytes_write.sum", None),
     "issue_active_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
     "sm_throughput_pct": ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
     "tensor_pipe_active_pct": ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1),
     "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1),
     "warp_instructions": ("smsp__inst_executed.sum", 1),
     "registers_per_thread": ("launch__registers_per_thread", 1),
     "lsu_wavefronts_pct": ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1)}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9,
        "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}


def num(s):
    return float(s.replace(",", ""))


def main():
    out_path, kernels = sys.argv[1], {}
    for spec in sys.argv[2:]:
        name, rep = spec.split("=", 1)
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units, vals = rows[0], rows[1], rows[-1]
        rec = {"kernel": vals[hdr.index("Kernel Name")].split("(")[0]}
        for key, (metric, scale) in M.items():
            if metric not in hdr:
                continue
            i = hdr.index(metric)
            v = num(vals[i]) * UNIT.get(units[i], 1)
            rec[key] = v * scale if scale is not None else v
        if "dram_read_bytes" in rec and "dram_write_bytes" in rec:
            rec["dram_bytes"] = rec["dram_read_bytes"] + rec["dram_write_bytes"]
        kernels[name] = rec
    json.dump({"source": "ncu --set full --clock-control none --import-source on, one launch each (scripts/gpu_round.sh); "
                         "dram_bytes = dram__bytes_read.sum + dram__bytes_write.sum per launch",
               "kernels": kernels}, open(out_path, "w"), indent=1)
    print(json.dumps(kernels, indent=1))


if __name__ == "__main__":
    main()
