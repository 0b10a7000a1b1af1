"""Reduce an `ncu --page raw --csv` dump (one kernel launch) to the metrics the profiles/ summaries quote, and optionally
record the kernel's DRAM traffic in profiles/traffic.json together with the hash of the source that was profiled.
usage: python devtools/ncu_summary.py <raw.csv> <out summary.json> "<command / workload note>"
           [--traffic <kernel key> <chains> <source file> <sha256 of that source when profiled>]
           [--stalls <source.csv> <out stalls.txt>]"""
import csv, json, os, sys, collections

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__issue_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__inst_executed_pipe_xu.sum", "sm__cycles_elapsed.max", "lts__t_bytes.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def read_raw(path):
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    return rows[hdr], rows[hdr + 1], rows[hdr + 2]


def to_bytes(v, u):
    return float(v.replace(",", "")) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u.lower(), 1)


def main():
    raw, out, note = sys.argv[1:4]
    hdr, units, vals = read_raw(raw)
    col = {h: i for i, h in enumerate(hdr)}
    m = {k: [vals[col[k]], units[col[k]]] for k in KEEP if k in col}
    js = dict(kernel=vals[col["Kernel Name"]], command=note, metrics=m)
    json.dump(js, open(out, "w"), indent=1)
    print(out, m.get("gpu__time_duration.sum"))
    a = sys.argv[4:]
    while a:
        if a[0] == "--traffic":
            key, chains, src, sha = a[1:5]; a = a[5:]
            rd = to_bytes(*m["dram__bytes_read.sum"]); wr = to_bytes(*m["dram__bytes_write.sum"])
            path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
            tj = json.load(open(path)) if os.path.exists(path) else {}
            tj[key] = dict(chains=int(chains), dram_bytes_per_launch=rd + wr, dram_bytes_read=rd, dram_bytes_write=wr, source=src,
                           source_sha256=sha, capture=os.path.basename(out))
            json.dump(tj, open(path, "w"), indent=1)
            print(tj[key])
        elif a[0] == "--stalls":
            src_csv, dst = a[1:3]; a = a[3:]
            rows = [r for r in csv.reader(open(src_csv)) if r]
            h = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in " ".join(r) or "Warp Stall Sampling (All Samples)" in r)
            head = rows[h]
            si = head.index("Source")
            ci = next(i for i, x in enumerate(head) if x.startswith("Warp Stall Sampling (All"))
            ii = next((i for i, x in enumerate(head) if x == "Instructions Executed"), None)
            samp = collections.Counter(); inst = collections.Counter(); tot = 0
            for r in rows[h + 1:]:
                if len(r) <= ci:
                    continue
                toks = r[si].split()
                if not toks:
                    continue
                op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
                op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("MUFU", "LDTM", "STTM", "SYNCS", "UTC")) and "." in op else "")
                try:
                    s = int(float(r[ci] or 0))
                except ValueError:
                    continue
                samp[op] += s; tot += s
                if ii is not None:
                    try:
                        inst[op] += int(float(r[ii] or 0))
                    except ValueError:
                        pass
            with open(dst, "w") as f:
                f.write(f"# per-opcode warp stall samples (all), {js['kernel']}\n# {note}\n# opcode  samples  share  warp-instructions-executed\n")
                for op, s in samp.most_common(40):
                    f.write(f"{op:24s} {s:10d} {100.0 * s / max(tot, 1):6.2f}% {inst[op]:14d}\n")
            print(dst, tot)
        else:
            raise SystemExit("unknown option " + a[0])


if __name__ == "__main__":
    main()
