"""Condense `ncu -i X.ncu-rep --page raw --csv` into the per-kernel summary kept under profiles/ and (optionally)
profiles/traffic.json (dram bytes per launch of the three hot entry points).

    python tools/ncu_summary.py raw.csv profiles/r01h_ncu_full_summary_cfg3_full.csv [--traffic cfg3]
"""
import csv
import json
import sys

COLS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
ENTRY = {"doc_pass_kernel": "ttlda_estep_f64", "sstats_csc_kernel": "ttlda_sstats_csc_f64",
         "theta_refresh_kernel": "ttlda_theta_refresh_f64"}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    raw, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(raw, errors="ignore")))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    head, units, body = rows[h], rows[h + 1], rows[h + 2:]
    idx = [head.index(c) for c in COLS if c in head]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([head[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in body:
            if len(r) == len(head):
                w.writerow([r[i] for i in idx])
    if "--traffic" in sys.argv:
        cfg = sys.argv[sys.argv.index("--traffic") + 1]
        kn, rd, wr = head.index("Kernel Name"), head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum")
        acc = {}
        for r in body:
            if len(r) != len(head):
                continue
            for k, e in ENTRY.items():
                if k in r[kn]:
                    b = float(r[rd]) * UNIT[units[rd]] + float(r[wr]) * UNIT[units[wr]]
                    acc.setdefault(e, []).append(b)
        # the LAST capture of each kernel (steady state: the first E-step launch is the ELBO-only initial pass)
        t = {e: v[-1] for e, v in acc.items()}
        json.dump({cfg: t, "_note": f"dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, "
                                    f"{cfg} full size, {out}"}, open("profiles/traffic.json", "w"), indent=1)
        print(t)


if __name__ == "__main__":
    main()
