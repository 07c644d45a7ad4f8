"""Turn the raw ncu pages of tools/r2_ncu.sh (gpurun_out/ncu/*.raw.csv) into the committed evidence:

  profiles/r02_ncu_kernels.csv   one line per captured launch: kernel (template arguments), its role in the cycle,
                                 device time, DRAM bytes, achieved DRAM GB/s = bytes / time, L1 / L2 hit rates,
                                 sectors per load request, threads per instruction, occupancy, registers, smem
  profiles/r02_ncu_traffic.json  DRAM bytes per launch keyed by bench.py's kernel_key (roofline.traffic), with the
                                 sha256 prefix of the libpyamg_b200.so the capture was taken from
  profiles/r02_launch_list_summary.txt   the --metrics gpu__time_duration.sum launch list grouped by kernel
"""
import csv
import json
import os
import re
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "ncu")
OUT = os.path.join(ROOT, "profiles")
OPS = {0: "spmv(restrict)", 1: "residual", 2: "prolong+add", 3: "jacobi", 4: "gs_wave", 5: "block_jacobi"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}
# the order in which the engine meets the (level, op) pairs of AMGB_NCU_SELECT in one V-cycle (tools/r2_ncu.sh)
CYCLE_ORDER = {
    "cycle": [(0, 4), (0, 1), (0, 0), (1, 4), (1, 4), (1, 1), (1, 0), (2, 4), (2, 4), (2, 1), (2, 0), (3, 4), (3, 4),
              (4, 4), (4, 4), (5, 4), (2, 2), (1, 2), (0, 2)],
    "l1_gs_full": [(1, 4)],
    "cfg5": [(0, 5), (0, 1), (0, 0), (1, 5), (0, 2)],
    # the first two SpMV launches of the cfg2 capture belong to the setup (spectral-radius Arnoldi on the device)
    "cfg2": [None, None, (0, 3), (0, 1), (0, 0), (1, 3), (2, 3), (0, 2)],
}


def short(name):
    m = re.match(r"(?:void )?(?:amgb::)?(\w+)(<.*>)?\(", name)
    if not m:
        return name[:60]
    targs = (m.group(2) or "").replace("amgb::", "").replace("TileCfg", "Cfg").replace(" ", "")
    return m.group(1) + targs


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name, default=None):
        i = col.get(name)
        if i is None or r[i] == "":
            return default
        try:
            v = float(r[i].replace(",", ""))
        except ValueError:
            return default
        return v * UNIT.get(units[i], 1.0)
    return body, col, get


def main():
    sha = open(os.path.join(SRC, "so_sha16.txt")).read().strip()
    lines, traffic = [], OrderedDict()
    captures = [
        ("fine", "cfg3 level-0 operator, natural order (isolation)", None),
        ("fine_jacobi_full", "fused Jacobi+residual, level-0 operator (isolation, --set full)", None),
        ("cycle", "cfg3 V-cycle (256^3 RS hierarchy)", "cfg3"),
        ("l1_gs_full", "cfg3 V-cycle, level-1 GS wave (--set full)", "cfg3"),
        ("cfg5", "cfg5 V-cycle (elasticity 300^2)", "cfg5"),
        ("cfg2", "cfg2 V-cycle (SA + Jacobi 2000^2)", "cfg2"),
        ("dense", "cfg5 coarsest level: pinv apply (dense_matvec_kernel)", None),
    ]
    fine_ops = ["spmv", "residual", "jacobi+residual (fused)", "jacobi"]
    for fname, what, cfg in captures:
        path = os.path.join(SRC, fname + ".raw.csv")
        if not os.path.exists(path):
            continue
        body, col, get = load(path)
        # a capture older than the sha file was left over from an earlier call of the script (its step timed out this
        # time): say so instead of attributing it to the current library
        stale = os.path.getmtime(path) < os.path.getmtime(os.path.join(SRC, "so_sha16.txt")) - 60
        sha_here = sha if not stale else "earlier build (capture step timed out in the final call; kernel source unchanged)"
        order = CYCLE_ORDER.get(fname)
        for k, r in enumerate(body):
            name = r[col["Kernel Name"]]
            t_us = get(r, "gpu__time_duration.sum")
            rd, wr = get(r, "dram__bytes_read.sum", 0.0), get(r, "dram__bytes_write.sum", 0.0)
            req = get(r, "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", 0.0)
            sec = get(r, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", 0.0)
            m = re.search(r"<(\d+),\s*(\d+)", name)
            op = int(m.group(2)) if m and ("csr_tile_kernel" in name or "csr_rows_kernel" in name) else None
            mf = re.search(r"csr_tile_flat_kernel<\(?(?:int\))?\s*(\d+)", name)
            if mf:
                op = int(mf.group(1))
            if "block_jacobi" in name:
                op = 5
            level = ""
            if fname == "fine":
                role = fine_ops[k] if k < 4 else ""
                level = 0
            elif fname == "fine_jacobi_full":
                role, level = "jacobi+residual (fused)", 0
            elif fname == "dense":
                role, level = "coarse pinv apply", 5
            else:
                role = OPS.get(op, "") if op is not None else ""
                if order is not None and k < len(order) and order[k] is not None and (op is None or order[k][1] == op):
                    level = order[k][0]
                    if cfg is not None and not stale:
                        key = f"{cfg}:L{level}:{OPS[order[k][1]].split('(')[0]}"
                        ent = traffic.setdefault(key, {"launches": 0, "dram_bytes": 0.0, "time_us": 0.0, "kernel": short(name)})
                        ent["launches"] += 1
                        ent["dram_bytes"] += rd + wr
                        ent["time_us"] += t_us
            lines.append(OrderedDict([
                ("capture", fname), ("what", what), ("launch", k), ("kernel", short(name)), ("level", level), ("role", role),
                ("grid", r[col["Grid Size"]].strip("()").split(",")[0]),
                ("block", r[col["Block Size"]].strip("()").split(",")[0]),
                ("time_us", round(t_us, 3)), ("dram_read_MB", round(rd / 1e6, 3)), ("dram_write_MB", round(wr / 1e6, 3)),
                ("dram_GBps", round((rd + wr) / t_us / 1e3, 1) if t_us else ""),
                ("dram_pct_of_peak", get(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")),
                ("lts_pct_of_peak", get(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed")),
                ("l1_hit_pct", get(r, "l1tex__t_sector_hit_rate.pct")), ("l2_hit_pct", get(r, "lts__t_sector_hit_rate.pct")),
                ("sectors_per_ld_request", round(sec / req, 2) if req else ""),
                ("threads_per_inst", get(r, "smsp__thread_inst_executed_per_inst_executed.ratio")),
                ("warps_active_pct", get(r, "sm__warps_active.avg.pct_of_peak_sustained_active")),
                ("long_scoreboard_per_issue",
                 get(r, "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio")),
                ("regs", get(r, "launch__registers_per_thread")),
                ("smem_dyn_KB", get(r, "launch__shared_mem_per_block_dynamic", 0.0) / 1e3),
                ("so_sha16", sha_here),
            ]))
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "r02_ncu_kernels.csv"), "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(lines[0].keys()))
        w.writeheader()
        w.writerows(lines)
    for ent in traffic.values():
        ent["dram_bytes_per_launch"] = ent["dram_bytes"] / ent["launches"]
        ent["us_per_launch_under_ncu"] = ent["time_us"] / ent["launches"]
    json.dump({"so_sha16": sha, "how": "ncu --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per "
               "launch; launches picked by the engine's AMGB_NCU_SELECT gate (tools/r2_ncu.sh), summarised by "
               "tools/ncu_summarise.py", "kernels": traffic}, open(os.path.join(OUT, "r02_ncu_traffic.json"), "w"), indent=1)
    print(f"{len(lines)} captured launches -> profiles/r02_ncu_kernels.csv, {len(traffic)} keys -> r02_ncu_traffic.json")
    for ln in lines:
        print(f"{ln['capture']:17s} {ln['kernel'][:52]:52s} L{ln['level']!s:2s} {ln['role']:24s} {ln['time_us']:8.1f} us "
              f"{ln['dram_read_MB'] + ln['dram_write_MB']:8.1f} MB {ln['dram_GBps']:>7} GB/s L1 {ln['l1_hit_pct']:.0f}% "
              f"L2 {ln['l2_hit_pct']:.0f}% s/req {ln['sectors_per_ld_request']} thr/inst {ln['threads_per_inst']} "
              f"warps {ln['warps_active_pct']:.0f}% lsb {ln['long_scoreboard_per_issue']}")
    path = os.path.join(SRC, "launches.csv")
    if os.path.exists(path):
        rows = [r for r in csv.reader(ln for ln in open(path) if ln.startswith('"'))]
        ci = {h: i for i, h in enumerate(rows[0])}
        agg = defaultdict(lambda: [0, 0.0])
        for r in rows[1:]:
            if len(r) <= ci["Metric Value"] or r[ci["Metric Name"]] != "gpu__time_duration.sum":
                continue
            v = float(r[ci["Metric Value"]].replace(",", "")) * UNIT.get(r[ci["Metric Unit"]], 1.0)
            a = agg[short(r[ci["Kernel Name"]])]
            a[0] += 1
            a[1] += v
        tot = sum(a[1] for a in agg.values())
        with open(os.path.join(OUT, "r02_launch_list_summary.txt"), "w") as f:
            f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 python bench.py --steps 2 "
                    f"--warmup 3 --configs ''   (first 2500 launches; library {sha}); serialised, cold-cache times: "
                    "compare SHARES, not absolutes\n")
            f.write(f"{'kernel':72s} {'launches':>9s} {'total_us':>12s} {'share':>7s}\n")
            for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"{k[:72]:72s} {a[0]:9d} {a[1]:12.1f} {a[1] / tot:7.3f}\n")
        print(open(os.path.join(OUT, "r02_launch_list_summary.txt")).read())


if __name__ == "__main__":
    main()
