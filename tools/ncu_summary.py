#!/usr/bin/env python
"""Summarise `ncu --set full` reports (read here, on the CPU box): one markdown section per profiled launch with the
counters the roofline needs, the warp-stall breakdown from the source page, and a JSON file bench.py reads for
`roofline.traffic`.

    python tools/ncu_summary.py --json profiles/r2_ncu_metrics.json --md profiles/r2_ncu_summary.md rep1.ncu-rep rep2.ncu-rep ...
"""
import argparse
import csv
import io
import json
import re
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "elapsed cycles"),
    ("launch__grid_size", "grid size"),
    ("launch__block_size", "block size"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), CTAs/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (shared memory), CTAs/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM bytes"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts (LSU)"),
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def short(name):
    return re.sub(r"\(.*$", "", name).replace("void ", "").replace("mflow::", "").strip()


def to_bytes(v, unit):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(v.replace(",", "")) * f


def stall_breakdown(rep, kernel_regex):
    rows = ncu_csv(rep, "source", ["--kernel-name", f"regex:{kernel_regex}"])
    hdr = next((r for r in rows if "Source" in r and "# Samples" in r), None)
    if hdr is None:
        return None
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: 0 for h in cols}
    top = []
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) != len(hdr):
            if top:
                break      # the next kernel's section
            continue
        if r[ix["Source"]] == "Source":
            break
        for h in cols:
            agg[h] += int(r[ix[h]] or 0)
        top.append((int(r[ix["# Samples"]] or 0), r[ix["Source"]].strip()))
    total = sum(agg.values()) or 1
    top.sort(reverse=True)
    return {k: v / total for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]}, top[:6], sum(t[0] for t in top) or 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+")
    ap.add_argument("--json", default=None)
    ap.add_argument("--md", default=None)
    a = ap.parse_args()
    md, js = [], {}
    for rep in a.reports:
        rows = ncu_csv(rep, "raw")
        if len(rows) < 3:
            print("no data in", rep, file=sys.stderr)
            continue
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        seen = set()
        for r in rows[2:]:
            name = short(r[ix["Kernel Name"]])
            if name in seen:
                continue
            seen.add(name)
            md.append(f"## `{name}`  ({rep.split('/')[-1]})\n")
            md.append("| counter | value |\n|---|---|")
            entry = {}
            for key, label in WANT:
                if key in ix:
                    v, u = r[ix[key]], units[ix[key]]
                    md.append(f"| {label} (`{key}`) | {v} {u} |")
                    try:
                        entry[key] = to_bytes(v, u) if "byte" in u else float(v.replace(",", ""))
                    except ValueError:
                        pass
            if "dram__bytes_read.sum" in entry:
                entry["dram_bytes"] = entry["dram__bytes_read.sum"] + entry.get("dram__bytes_write.sum", 0.0)
                md.append(f"| **DRAM traffic (read + written)** | {entry['dram_bytes'] / 1e6:.1f} MB |")
            sb = stall_breakdown(rep, re.escape(name.split("<")[0]))
            if sb:
                shares, top, nsamp = sb
                md.append("\nWarp-stall samples (all warps): " + ", ".join(f"{k.replace('stall_', '')} {100 * v:.0f}%" for k, v in shares.items()))
                md.append("\nHottest instructions (share of samples):\n")
                for cnt, src in top:
                    md.append(f"* {100 * cnt / nsamp:.1f}%  `{src[:110]}`")
                entry["stalls"] = {k.replace("stall_", ""): round(v, 3) for k, v in shares.items()}
            md.append("")
            js[name] = entry
    if a.md:
        with open(a.md, "w") as f:
            f.write("# ncu --set full summaries (generated by tools/ncu_summary.py)\n\n" + "\n".join(md) + "\n")
    if a.json:
        # short aliases bench.py looks up
        out = dict(js)
        for name, e in js.items():
            if name.startswith("conv_split3_kernel") and "conv" not in out:
                out["conv"] = e
            if name.startswith("wgrad_split3_kernel") and "wgrad" not in out:
                out["wgrad"] = e
            if name.startswith("rqs_planes_apply_kernel") and "rqs_planes_apply" not in out:
                out["rqs_planes_apply"] = e
            if name.startswith("rqs_planes_backward_kernel") and "rqs_planes_backward" not in out:
                out["rqs_planes_backward"] = e
        with open(a.json, "w") as f:
            json.dump(out, f, indent=1)
    print("\n".join(md))


if __name__ == "__main__":
    main()
