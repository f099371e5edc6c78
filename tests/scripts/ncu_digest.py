#!/usr/bin/env python
"""Tool-derived digests of the committed ncu exports (no hand-typed numbers):
  profiles/r02_ncu_launch_shares.txt   per-kernel share of the launch list (gpu__time_duration.sum pass)
  profiles/r02_ncu_key_metrics.txt     selected rows of the raw page of the --set full capture
  profiles/traffic.json                dram bytes per launch of the dominant kernel (read by bench.py)
CPU-only:  python tests/scripts/ncu_digest.py"""
import collections
import csv
import json
import os
import re

RAW = "r02c_ncu_raw_lstm_f512_h256.csv"   # the capture of the final tree; r02b_ = two rings, before the lean issue loop; r02_ = start of the session
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = os.path.join(ROOT, "profiles")


def launch_shares():
    rows = []
    with open(os.path.join(P, "r02_ncu_launches.csv")) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    tot = collections.OrderedDict()
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1; t[1] += ms
    total = sum(v[1] for v in tot.values())
    with open(os.path.join(P, "r02_ncu_launch_shares.txt"), "w") as f:
        f.write("# per-kernel share of the ncu launch list profiles/r02_ncu_launches.csv (gpu__time_duration.sum, --clock-control none;\n"
                "# command: python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --streams 1; cold-cache, serialised: compare SHARES)\n")
        for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{name:<70} launches {n:4d}  {ms:10.3f} ms  {100 * ms / total:6.2f} %\n")
        f.write(f"{'total':<70} {'':13}  {total:10.3f} ms\n")


def key_metrics():
    rows = list(csv.reader(open(os.path.join(P, RAW))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg",
            "sm__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.max",
            # shared-memory data pipe: operand reads of the tensor core, LSU traffic of the epilogue warps, bulk-copy fills
            "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "memory_l1_wavefronts_shared", "memory_l1_wavefronts_shared_ideal",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
            "l1tex__m_xbar2l1tex_read_bytes_mem_dshared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
    out = {}
    with open(os.path.join(P, "r02_ncu_key_metrics.txt"), "w") as f:
        f.write(f"# selected rows of profiles/{RAW} (ncu --set full --clock-control none, launch 19 of\n"
                "# python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --streams 1 = stage-1 rnn2, F=512, H=256, B=1024, T=300)\n")
        for i, h in enumerate(hdr):
            if any(h == k or h.endswith(k) for k in keys):
                f.write(f"{h:<95} {vals[i]:>20} {units[i]}\n")
                out[h] = (vals[i], units[i])
        # derived: how busy the shared-memory data pipe (one 128-byte wavefront per cycle per SM) is
        def num(k):
            return float(out[k][0].replace(",", ""))
        unit_b = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
        tma = num("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum") * unit_b[out["l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum"][1]] / 128
        dsm = num("l1tex__m_xbar2l1tex_read_bytes_mem_dshared.sum") * unit_b[out["l1tex__m_xbar2l1tex_read_bytes_mem_dshared.sum"][1]] / 128
        tc, lsu = num("l1tex__data_pipe_tc_wavefronts_mem_shared.sum"), num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
        cyc = num("sm__cycles_elapsed.max")
        grid = num("launch__grid_size")
        active = min(148.0, 4 * (33 if num("launch__cluster_dim_x") == 4 else 148 // 4))
        f.write("# derived: 128-byte wavefronts of the shared-memory data pipe per launch: tensor-core operand reads %.3g, LSU (incl. replays) %.3g,\n"
                "#   bulk-copy fills from L2 %.3g, DSMEM in + out 2 x %.3g  =  %.1f %% of (148 SMs x elapsed cycles), %.1f %% of the %d SMs the 33 clusters occupy\n"
                % (tc, lsu, tma, dsm, 100 * (tc + lsu + tma + 2 * dsm) / (148 * cyc), 100 * (tc + lsu + tma + 2 * dsm) / (active * cyc), int(active)))
    def gb(k):
        v, u = out[k]
        v = float(v.replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
    tr = {"tc:v1:F512:H256": {"dram_bytes": gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"),
                              "source": f"profiles/{RAW} (ncu --set full, one launch, round 2, final tree)"}}
    json.dump(tr, open(os.path.join(P, "traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    launch_shares()
    key_metrics()
    print(open(os.path.join(P, "r02_ncu_launch_shares.txt")).read())
    print(open(os.path.join(P, "r02_ncu_key_metrics.txt")).read())
