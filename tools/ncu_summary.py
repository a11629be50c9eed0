"""Turns an Nsight Compute report (+ launch list) into the markdown summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep gpurun_out/launches.csv > profiles/ncu_rNN_summary.md
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("launch__registers_per_thread", "registers/thread"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def main():
    rep, launches = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# Nsight Compute summary (`%s`)\n" % rep.split("/")[-1])
    print("`ncu --set full --clock-control none --import-source on`, one launch per kernel, taken from "
          "`bench.py --steps 1 --warmup 3` (128 clouds x 8192 points).  Durations under ncu are cold-cache "
          "and serialised: compare shares, not absolutes.\n")
    print("| kernel | " + " | ".join(n for _, n in METRICS) + " |")
    print("|---|" + "---|" * len(METRICS))
    for d in data:
        name = d[idx["Kernel Name"]].replace("void ", "").replace("<unnamed>::", "")
        name = name.split("(const")[0][:48]
        cells = []
        for m, _ in METRICS:
            if m in idx:
                v, u = d[idx[m]], units[idx[m]]
                try:
                    v = "%.4g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                cells.append("%s %s" % (v, u if u not in ("%", "") else ""))
            else:
                cells.append("-")
        print("| `%s` | " % name + " | ".join(c.strip() for c in cells) + " |")
    if launches:
        lines = [l for l in open(launches) if not l.startswith("==")]
        agg = collections.OrderedDict()
        for r in csv.DictReader(lines):
            if r.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(r["Metric Value"].replace(",", ""))
            v = v / 1e3 if r["Metric Unit"] == "ns" else v * 1e3 if r["Metric Unit"] == "ms" else v
            k = r["Kernel Name"].replace("void ", "").replace("<unnamed>::", "").split("(const")[0][:60]
            agg.setdefault(k, []).append(v)
        tot = sum(sum(v) for v in agg.values())
        print("\n## Launch list (`%s`): every launch of `bench.py --steps 2 --warmup 3`\n" % launches.split("/")[-1])
        print("| kernel | launches | mean us | share of GPU time |")
        print("|---|---|---|---|")
        for k, v in agg.items():
            print("| `%s` | %d | %.1f | %.1f %% |" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))


if __name__ == "__main__":
    main()
