"""Summaries committed under profiles/ from ncu exports.

  python profiles/tools/ncu_summary.py launches <launch_list.csv> <summary_out.csv>
      launch list = `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...` of a bench.py run
  python profiles/tools/ncu_summary.py top <raw_page.csv> <out.json>
      raw page = `ncu -i prof.ncu-rep --page raw --csv` of a `--set full` capture; keeps the metrics the roofline discussion uses
"""
import collections
import csv
import json
import re
import sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.avg.per_second", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("fmb::", "").replace("(int)", "").replace("(bool)", "").strip()


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        full = r[4]
        m = re.match(r"(void )?(fmb::)?(\w+)(<[^>]*>)?", full)
        key = (m.group(3) + (m.group(4) or "")) if m else full[:40]
        key = key.replace("(int)", "").replace("(bool)", "")
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += float(r[-1]) / 1e6          # ns -> ms
    total = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("kernel,launches,total_ms,share_pct,avg_us\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.3f,%.2f,%.1f\n" % (k.replace(",", ";"), n, ms, 100 * ms / total, 1000 * ms / n))


def top(src, dst):
    r = list(csv.reader(open(src)))
    hdr, units = r[0], r[1]
    out = []
    for row in r[2:]:
        d = {"Kernel Name": row[hdr.index("Kernel Name")]}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                d[k] = "%s %s" % (row[i], units[i])
        out.append(d)
    json.dump(out, open(dst, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "top": top}[sys.argv[1]](sys.argv[2], sys.argv[3])
