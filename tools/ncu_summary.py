#!/usr/bin/env python
"""Turns the ncu artefacts of a round (gpurun_out/) into the small, committed summaries under profiles/.
   python tools/ncu_summary.py <tag>          # reads gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_k_hash.ncu-rep
Runs here (no GPU): `ncu -i` only reads the report."""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaheavy pipe active % (IMAD.WIDE lives here)"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "alu pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction (i-cache)"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall: dispatch"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (memory)"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]


def ncu_csv(rep, page, extra=()):
    r = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True, check=True)
    return list(csv.reader(io.StringIO(r.stdout)))


def launches(tag, md):
    path = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    seq = []
    for r in rows[start + 1:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")[-60:]
        ns = float(r[vi].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
        seq.append((name, r[gi], ns))
    tot = sum(v[1] for v in agg.values())
    md.append(f"## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, {len(seq)} launches, cold-cache, serialised)\n")
    md.append("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:10]:
        md.append(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% |")
    hs = [s for s in seq if "k_hash" in s[0]]
    # the first leaf kernel of a build (the context self-test launches one-block k_hash<2|3> / k_hash_coop<2|3> before it)
    first = next((i for i, s in enumerate(hs) if "3>" in s[0] and "coop" not in s[0] and not s[1].startswith("(1,")), None)
    if first is not None:
        md.append("\nOne build, launch by launch (blocks of 128 threads: 128 hashes per block on `k_hash`, 32 on `k_hash_coop`; from depth 16 on the two "
                  "half-trees alternate, so every level appears twice; serialised by ncu — in a real build the two halves overlap):\n\n"
                  "| kernel | grid | ms |\n|---|---|---:|")
        tot_build = 0.0
        for s in hs[first:first + 80]:
            if s is not hs[first] and "3>" in s[0] and "coop" not in s[0]:
                break
            tot_build += s[2]
            md.append(f"| `{s[0][-14:]}` | {s[1]} | {s[2] / 1e6:.3f} |")
        md.append(f"\nSum of the build's launches: {tot_build / 1e6:.3f} ms (serialised).")
    md.append("")


SECTOR_KEYS = [
    ("lts__t_sectors_op_read.sum", "L2 sectors read"), ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 sectors read for L1"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 sectors, global loads"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum", "L1 sectors hit, global loads"),
    ("dram__sectors_read.sum", "DRAM sectors read"), ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__inst_executed_op_shared_ld.sum", "shared-memory load instructions"),
]


def full(tag, md, suffix="k_hash", title="Top kernels", sectors=False):
    rep = os.path.join(ROOT, "gpurun_out", f"{tag}_{suffix}.ncu-rep")
    if not os.path.exists(rep):
        return
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    md.append(f"## {title}, `ncu --set full --clock-control none --import-source on` ({tag}_{suffix}.ncu-rep)\n")
    for r in rows[2:]:
        md.append(f"### `{r[hdr.index('Kernel Name')][:60]}`\n\n| metric | value |\n|---|---:|")
        for key, label in KEYS + (SECTOR_KEYS if sectors else []):
            if key in hdr and "nan" not in r[hdr.index(key)]:
                i = hdr.index(key)
                md.append(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
        md.append("")
    # dynamic instruction mix from the source page
    src = ncu_csv(rep, "source", ["--print-source", "sass"])
    kern, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            kern.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    seen = set()
    for k in kern:
        if k["name"] in seen:
            continue
        seen.add(k["name"])
        h = k["hdr"]
        i_s, i_e, i_n = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        ops, samp = collections.Counter(), collections.Counter()
        for r in k["rows"]:
            if len(r) < len(h):
                continue
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[i_s].strip())
            op = m.group(2) if m else r[i_s].strip()
            ops[op] += int(r[i_e])
            samp[op] += int(r[i_n])
        tot, ts = sum(ops.values()), max(1, sum(samp.values()))
        md.append(f"Dynamic SASS mix of `{k['name'][:40]}` ({len(k['rows'])} static instructions = {len(k['rows']) * 16 / 1024:.1f} KB):\n")
        md.append("| opcode | share of executed | share of stall samples |\n|---|---:|---:|")
        for op, c in ops.most_common(8):
            md.append(f"| `{op}` | {100 * c / tot:.1f}% | {100 * samp[op] / ts:.1f}% |")
        md.append("")


def main():
    tag = sys.argv[1]
    md = [f"# ncu summary `{tag}` (B200, sm_100a) — generated by tools/ncu_summary.py from gpurun_out/{tag}_*\n"]
    launches(tag, md)
    full(tag, md)
    full(tag, md, "k_coop", "Cooperative kernel (3 lanes per hash): levels of 8192, 4096, ... nodes")
    full(tag, md, "k_lh", "Lead / helper latency kernel (the S-box chain in a warp of its own): levels of <= 888 nodes per half-tree")
    full(tag, md, "k_trace", "Witness-trace fold (k_fold_paths with the state sink): 2^14 paths of the depth-20 tree")
    full(tag, md, "k_tree_trace", "Witness traces from the resident tree (k_trace_tree_paths): 2^14 paths of the depth-20 tree, one thread per (query, level)")
    full(tag, md, "k_lookup", "Low-leaf lookup, prefix array + shared-memory top (k_low_leaf_lookup_fast): 2^20 queries, depth-24 index", sectors=True)
    full(tag, md, "k_lookup_plain", "Low-leaf lookup, plain binary search over the 32-byte keys (k_low_leaf_lookup): the same queries", sectors=True)
    os.makedirs(OUT, exist_ok=True)
    out = os.path.join(OUT, f"{tag}_summary.md")
    open(out, "w").write("\n".join(md) + "\n")
    print(out)


if __name__ == "__main__":
    main()
