"""Turns an .ncu-rep (ncu --set full) into a compact per-launch summary for profiles/.

    python tools/summarise_ncu.py gpurun_out/adam_prof.ncu-rep profiles/r01_adam_ncu.txt [json_out]
"""
import csv
import io
import json
import re
import subprocess
import sys

KEYS = [
    ("time_us", "gpu__time_duration.sum"),
    ("dram_read_MB", "dram__bytes_read.sum"),
    ("dram_write_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1tex_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("lts_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_pipe_pct", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor_inst_pct", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("smem_dyn_KB", "launch__shared_mem_per_block_dynamic"),
    ("inst_executed", "smsp__inst_executed.sum"),
    ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("l2_read_MB", "lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_ld.sum"),
    ("sm_clock_MHz", "sm__cycles_elapsed.avg.per_second"),
]
UNIT_SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Kbyte/block": 1.0,
              "byte/block": 1.0 / 1024}


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    tensor_cols = [h for h in hdr if "tensor" in h and "pct_of_peak_sustained_active" in h]
    recs, lines = [], []
    for r in body:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
        rec = {"kernel": name}
        for key, col in KEYS:
            if col in idx:
                v = num(r[idx[col]])
                if v is None:
                    continue
                u = units[idx[col]]
                if key.endswith("_us") or key.endswith("_MB") or key.endswith("_KB"):
                    v *= UNIT_SCALE.get(u, 1.0)
                if key == "sm_clock_MHz":
                    v *= {"Ghz": 1e3, "Mhz": 1.0, "hz": 1e-6}.get(u, 1.0)
                rec[key] = round(v, 3)
        if tensor_cols:
            rec["tensor_pct_max"] = max((num(r[idx[c]]) or 0.0) for c in tensor_cols)
        stalls = []
        for h, i in idx.items():
            m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h)
            if m and num(r[i]) is not None:
                stalls.append((m.group(1), num(r[i])))
        rec["top_stalls"] = [[n, round(v, 2)] for n, v in sorted(stalls, key=lambda x: -x[1])[:4]]
        if "dram_read_MB" in rec and "dram_write_MB" in rec:
            rec["dram_traffic_MB"] = round(rec["dram_read_MB"] + rec["dram_write_MB"], 3)
            if rec.get("time_us"):
                rec["dram_GBps"] = round(rec["dram_traffic_MB"] / rec["time_us"] * 1e3, 1)
        recs.append(rec)
        lines.append(json.dumps(rec))
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none summary of %s (one JSON object per captured launch)\n" % rep.split("/")[-1])
        f.write("\n".join(lines) + "\n")
    if len(sys.argv) > 3:
        json.dump(recs, open(sys.argv[3], "w"), indent=1)
    for rec in recs:
        print(rec["kernel"][:50], rec.get("time_us"), "us dram", rec.get("dram_traffic_MB"), "MB", rec.get("dram_GBps"), "GB/s tensor",
              rec.get("tensor_pct_max"), "stalls", rec["top_stalls"][:2])


if __name__ == "__main__":
    main()
