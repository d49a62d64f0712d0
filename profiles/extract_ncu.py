"""Turn an `ncu --set full` report into the per-kernel summary committed under profiles/ and into
profiles/ncu_traffic.json (DRAM bytes per launch, read by bench.py for `roofline.traffic`).

    python profiles/extract_ncu.py gpurun_out/prof_step_k.ncu-rep profiles/r01_ncu_full_v2_summary.txt c2

Runs here (no GPU needed): it only reads the report with `ncu -i ... --page raw --csv`."""
import csv, io, json, os, subprocess, sys
from collections import defaultdict

STAGE_OF = {  # kernel-name substring -> bench.py stage key
    "preprocess_kernel": "preprocess", "emit_kernel": "emit", "rank_sums_kernel": "emit", "render_forward_kernel": "render_fwd",
    "render_backward_kernel": "render_bwd", "gaussian_backward_kernel": "gaussian_bwd", "tile_ranges_kernel": "ranges",
    "adam_geometry_kernel": "adam_geometry", "adam_flat_kernel": "adam_flat", "adam_rest_kernel": "adam_rest",
    "adam_list_kernel": "adam_list", "adam_rest_list_kernel": "adam_rest_list", "lazy_color_kernel": "lazy_color",
    "radix_": "sorts",
}
WANT = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes_per_inst"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("smsp__inst_executed.sum", "warp_inst"),
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3,
         "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}


def main():
    rep, out_txt, config = sys.argv[1], sys.argv[2], sys.argv[3]
    if rep.endswith(".csv"):  # already exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv`): reports of a whole
        raw = open(rep).read()  # step exceed what gpurun copies back
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]

    def col(metric):
        if metric in head:
            return head.index(metric)
        if metric.replace("dram__throughput", "gpu__dram_throughput") in head:
            return head.index(metric.replace("dram__throughput", "gpu__dram_throughput"))
        return None

    cols = {short: col(metric) for metric, short in WANT}
    kcol = head.index("Kernel Name")
    per = defaultdict(list)
    for r in body:
        rec = {}
        for short, c in cols.items():
            if c is None or r[c] in ("", "no data", "n/a"):
                rec[short] = None
                continue
            v = float(r[c].replace(",", ""))
            rec[short] = v * SCALE.get(units[c], 1.0)
        per[r[kcol].split("(")[0]].append(rec)

    lines = ["# %s  (ncu --set full --clock-control none; per-launch averages; cold-cache, serialised replays)" % os.path.basename(rep),
             "%-46s %3s %9s %9s %9s %6s %6s %6s %6s %6s %5s %5s %6s" % ("kernel", "n", "time_us", "dramRdMB", "dramWrMB", "dram%",
                                                                    "l2%", "l1%", "sm%", "issue%", "occ%", "regs", "lanes")]
    traffic = {}
    for k, recs in per.items():
        def avg(key):
            vals = [x[key] for x in recs if x[key] is not None]
            return sum(vals) / len(vals) if vals else float("nan")
        lines.append("%-46s %3d %9.1f %9.2f %9.2f %6.1f %6.1f %6.1f %6.1f %6.1f %5.1f %5.0f %6.1f" % (
            k[:46], len(recs), avg("time_us"), avg("dram_rd") / 1e6, avg("dram_wr") / 1e6, avg("dram_pct"), avg("l2_pct"),
            avg("l1_pct"), avg("sm_pct"), avg("issue_pct"), avg("occupancy_pct"), avg("regs"), avg("lanes_per_inst")))
        for sub, stage in STAGE_OF.items():
            if sub in k:  # kernels of the same stage (e.g. the front and back instantiation) add up
                n_launch = len(recs) if stage in ("sorts", "emit") else 1  # multi-launch stages: totals of the capture
                traffic[stage] = traffic.get(stage, 0) + int(avg("dram_rd") + avg("dram_wr")) * n_launch
                if avg("warp_inst") == avg("warp_inst"):
                    traffic[stage + "_warp_inst"] = traffic.get(stage + "_warp_inst", 0) + int(avg("warp_inst")) * n_launch
        if "Onesweep" in k or "RadixSort" in k:
            traffic.setdefault("_radix_kernels", 0)
            traffic["_radix_kernels"] += int((avg("dram_rd") + avg("dram_wr")) * len(recs))
    open(out_txt, "w").write("\n".join(lines) + "\n")
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json")
    allt = json.load(open(tpath)) if os.path.exists(tpath) else {}
    allt.setdefault(config, {}).update(traffic)
    if "adam_flat_kernel" in per:
        allt[config].pop("adam_geometry", None)
    allt[config]["_source"] = os.path.basename(out_txt)
    json.dump(allt, open(tpath, "w"), indent=1, sort_keys=True)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
