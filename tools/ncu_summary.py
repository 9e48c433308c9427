"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/<name>.md + .json.

    python tools/ncu_summary.py gpurun_out/r01_k2_cg1.ncu-rep profiles/r01_k2_cg1 "note"
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "lts__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.avg.per_cycle_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__cluster_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    result = {"report": rep, "note": note, "kernels": []}
    md = [f"# ncu summary: {rep}", "", note, ""]
    for vals in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, vals):
            hh = h.split(".Triage")[0] if False else h
            for k in KEYS:
                if hh == k or hh.endswith("." + k):
                    d[k] = (v, u)
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        kern = {"name": name}
        md.append(f"## {name}")
        md.append("")
        md.append("| metric | value | unit |")
        md.append("|---|---|---|")
        for k in KEYS:
            if k in d:
                md.append(f"| {k} | {d[k][0]} | {d[k][1]} |")
                try:
                    kern[k] = float(d[k][0].replace(",", ""))
                    kern[k + "__unit"] = d[k][1]
                except ValueError:
                    kern[k] = d[k][0]
        def b(key):
            if key not in d:
                return None
            return float(d[key][0].replace(",", "")) * UNIT_SCALE.get(d[key][1], 1.0)
        r, w = b("dram__bytes_read.sum"), b("dram__bytes_write.sum")
        if r is not None and w is not None:
            kern["dram_bytes_per_launch"] = r + w
            md.append(f"| **dram bytes per launch (read+write)** | {r + w:.4g} | byte |")
        md.append("")
        result["kernels"].append(kern)
    # stall hot spots
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) > 2:
        h = rows[1]
        if "# Samples" in h:
            i_s, i_src, i_ex = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
            body = [r for r in rows[2:] if len(r) > i_s and r[i_s].isdigit()]
            tot = sum(int(r[i_s]) for r in body) or 1
            md += ["## top stall-sample instructions", "", "| samples | share | warp-insts executed | SASS |", "|---|---|---|---|"]
            for r in sorted(body, key=lambda r: -int(r[i_s]))[:14]:
                md.append(f"| {r[i_s]} | {100 * int(r[i_s]) / tot:.1f}% | {r[i_ex]} | `{r[i_src][:90]}` |")
            md.append("")
    open(out + ".md", "w").write("\n".join(md))
    json.dump(result, open(out + ".json", "w"), indent=1)
    print("\n".join(md[:40]))


if __name__ == "__main__":
    main()
