"""Condense `ncu -i X.ncu-rep --page raw --csv` into the per-kernel summary kept under profiles/ and (optionally) merge
the DRAM traffic per ENTRY POINT per EM iteration into profiles/traffic.json (what bench.py reports as
`roofline.traffic`).

    python tools/ncu_summary.py raw.csv profiles/r02k_ncu_full_summary_cfg3.csv [--traffic cfg3]

An entry point may launch several kernels (the topic-windowed statistics pass: one pass + carry + fold kernel per
window); its traffic is the SUM over the kernels it launched in ONE iteration — iterations are delimited by the E-step
kernel (`doc_pass_*`), the last complete one is used.
"""
import csv
import json
import os
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
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def entry_of(kernel, windowed):
    if "doc_pass" in kernel:
        return "ttlda_estep_hot_f64" if "hot" in kernel else "ttlda_estep_f64"
    if "sstats_" in kernel:
        return "ttlda_sstats_csc_windowed_f64" if windowed else "ttlda_sstats_csc_f64"
    if "theta_refresh" in kernel:
        return "ttlda_theta_refresh_f64"
    if "mstep_lambda" in kernel or "beta_refresh" in kernel or "colsum_finish" in kernel:
        return "ttlda_mstep_f64"
    return None


def main():
    raw, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(raw, errors="ignore")))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    head, units, body = rows[h], rows[h + 1], [r for r in rows[h + 2:] if len(r) == len(rows[h])]
    idx = [head.index(c) for c in COLS if c in head]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([head[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in body:
            w.writerow([r[i] for i in idx])
    if "--traffic" in sys.argv:
        cfg = sys.argv[sys.argv.index("--traffic") + 1]
        kn, rd, wr = head.index("Kernel Name"), head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum")
        tm = head.index("gpu__time_duration.sum")
        windowed = any("sstats_fold_window" in r[kn] for r in body)
        steps, cur = [], None
        for r in body:
            e = entry_of(r[kn], windowed)
            if e in ("ttlda_estep_f64", "ttlda_estep_hot_f64") and not r[kn].rstrip(")").endswith(", 0>(EStepParams"):
                cur = {}
                steps.append(cur)
            if cur is None or e is None:
                continue
            try:
                b = float(r[rd]) * UNIT[units[rd]] + float(r[wr]) * UNIT[units[wr]]
            except ValueError:
                b = float("nan")
            t = cur.setdefault(e, [0.0, 0.0, 0, []])
            t[3].append((r[kn].split("(")[0], b))
            t[1] += float(r[tm])
            t[2] += 1
        for s_ in steps:  # a launch whose DRAM counters ncu reported as nan takes the mean of the same kernel's other
            for t in s_.values():  # launches inside the same entry point call (the windows of one pass are alike)
                for name, b in t[3]:
                    if b != b:
                        ok = [x for n_, x in t[3] if n_ == name and x == x]
                        b = sum(ok) / len(ok) if ok else float("nan")
                    t[0] += b
        def nlaunch(s):
            return sum(v[2] for v in s.values())

        full = [s for s in steps if any("sstats" in k for k in s)]
        if not full:
            print("no complete iteration captured")
            return
        most = max(nlaunch(s) for s in full)
        last = [s for s in full if nlaunch(s) == most][-1]  # the last COMPLETE iteration (a capture may end mid-way)
        path = os.path.join("profiles", "traffic.json")
        try:
            j = json.load(open(path))
        except Exception:
            j = {}
        j[cfg] = {k: (v[0] if v[0] == v[0] else None) for k, v in last.items()}
        j.setdefault("_sources", {})[cfg] = out
        j["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per ENTRY POINT per EM iteration (summed over the kernels "
                      "the entry point launches), ncu --set full --clock-control none; per configuration, from the capture "
                      "named in _sources")
        json.dump(j, open(path, "w"), indent=1)
        print(cfg, {k: (round(v[0] / 1e9, 3), round(v[1], 3), v[2]) for k, v in last.items()}, "(GB, ms under ncu, launches)")


if __name__ == "__main__":
    main()
