"""Text summary of an ncu --set full report: per kernel the headline metrics, the warp-stall
breakdown and the share of samples between consecutive barriers (= the kernel's phases).
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_name.txt
Run in the build container (needs the `ncu` CLI, no GPU)."""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
STALLS = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_math", "stall_wait", "stall_not_selected",
          "stall_selected", "stall_mio", "stall_no_inst"]


def ncu(rep, page, *extra):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main(rep):
    rows = ncu(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("=" * 100)
        print(name[:160])
        for w in WANT:
            if w in hdr:
                print(f"  {w:75s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
        st = [(num(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for i, h in enumerate(hdr)
              if h.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in h]
        tot = sum(v for v, _ in st) or 1.0
        print("  warp stall sampling: " + "  ".join(f"{h} {100 * v / tot:.1f}%" for v, h in sorted(st, reverse=True)[:9]))
        short = name.split("(")[0].split("<")[0].split("::")[-1].split()[-1]
        src = ncu(rep, "source", "--kernel-name", "regex:" + short)
        if len(src) < 3:
            continue
        sh = src[1]
        iS, iSrc, iI = sh.index("# Samples"), sh.index("Source"), sh.index("Instructions Executed")
        cols = {n: sh.index(n) for n in STALLS if n in sh}
        # several launches of one kernel are concatenated: keep the first copy only
        data, seen_exit = [], False
        for rr in src[2:]:
            if len(rr) < len(sh):
                if data:
                    break
                continue
            data.append(rr)
        total = sum(num(x[iS]) for x in data) or 1.0
        print("  phases (samples between consecutive barriers; share of all samples, instructions, top stalls, top opcodes):")
        seg = dict(n=0.0, inst=0.0, ops={}, st={k: 0.0 for k in cols}, start=0)
        for li, rr in enumerate(data):
            seg["n"] += num(rr[iS]); seg["inst"] += num(rr[iI])
            toks = rr[iSrc].split()
            op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "")
            op = op.split(".")[0]
            seg["ops"][op] = seg["ops"].get(op, 0.0) + num(rr[iI])
            for k, i in cols.items():
                seg["st"][k] += num(rr[i])
            if "BAR" in rr[iSrc] or "EXIT" in rr[iSrc] or li == len(data) - 1:
                if seg["n"] >= 0.004 * total:
                    top = sorted(seg["st"].items(), key=lambda kv: -kv[1])[:3]
                    ops = sorted(seg["ops"].items(), key=lambda kv: -kv[1])[:5]
                    print(f"    sass {seg['start']:5d}-{li:5d}  {100 * seg['n'] / total:5.1f}%  inst {seg['inst']:12.0f}  "
                          + " ".join(f"{k.replace('stall_', '')}={100 * v / total:.1f}" for k, v in top) + "  | "
                          + " ".join(f"{a}:{b / 1e3:.0f}k" for a, b in ops))
                seg = dict(n=0.0, inst=0.0, ops={}, st={k: 0.0 for k in cols}, start=li + 1)


if __name__ == "__main__":
    main(sys.argv[1])
