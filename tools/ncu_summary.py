#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw + source pages) into text: headline metrics and hot SASS regions."""
import csv, subprocess, sys, io

def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']

STALLS = ['stall_barrier', 'stall_branch_resolving', 'stall_dispatch', 'stall_drain', 'stall_lg', 'stall_long_sb',
          'stall_math', 'stall_membar', 'stall_mio', 'stall_misc', 'stall_no_inst', 'stall_not_selected',
          'stall_selected', 'stall_short_sb', 'stall_sleep', 'stall_tex', 'stall_wait']

def traffic(rep, codewords):
    """--traffic N: DRAM bytes per codeword of the (single) captured launch -> JSON on stdout"""
    import json
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    ci = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = float(vals[ci["dram__bytes_read.sum"]]) * scale[units[ci["dram__bytes_read.sum"]]]
    wr = float(vals[ci["dram__bytes_write.sum"]]) * scale[units[ci["dram__bytes_write.sum"]]]
    print(json.dumps({"dram_bytes_per_codeword": (rd + wr) / codewords, "dram_bytes_read": rd, "dram_bytes_write": wr,
                      "codewords": codewords, "kernel": vals[ci["Kernel Name"]],
                      "source": "ncu --set full --clock-control none, tools/scl_prof.sh (5th scl_list launch of tools/scl_perf.py "
                                f"{codewords} 8: detector pairing)"}))


def all_kernels(rep):
    """--all: one block of headline metrics per captured launch"""
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    extra = ['dram__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'launch__grid_size',
             'launch__block_size', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
             'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
             'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
             'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio']
    for vals in rows[2:]:
        print(f"### {vals[ci['Kernel Name']]}  grid {vals[ci['launch__grid_size']]} x block {vals[ci['launch__block_size']]}")
        for h, u, v in zip(hdr, units, vals):
            try:
                big = float(v or 0) > 0.15
            except ValueError:
                big = False
            if h in KEYS or h in extra[:2] or h in extra[4:5] or ('issue_stalled' in h and 'per_issue_active' in h and big):
                print(f"  {h:86s} {u:14s} {v}")
        print()


def main():
    rep = sys.argv[1]
    if len(sys.argv) > 3 and sys.argv[2] == "--traffic":
        return traffic(rep, int(sys.argv[3]))
    if len(sys.argv) > 2 and sys.argv[2] == "--all":
        return all_kernels(rep)
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS or ('issue_stalled' in h and 'per_issue_active' in h and float(v or 0) > 0.1):
            print(f"{h:88s} {u:14s} {v}")
    rows = page(rep, "source")
    hdr, data = rows[1], rows[2:]
    ci = {h: i for i, h in enumerate(hdr)}
    ex = [int(r[ci['Instructions Executed']]) for r in data]
    sm = [int(r[ci['# Samples']]) for r in data]
    tot, totS = sum(ex), sum(sm)
    print(f"\nSASS instructions: {len(data)}  executed: {tot}  samples: {totS}")
    runs, cur = [], None
    for k, r in enumerate(data):
        if cur and abs(ex[k] - cur['ex']) <= 0.02 * max(cur['ex'], 1):
            cur['n'] += 1; cur['s'] += sm[k]; cur['end'] = k
            for key in STALLS:
                cur[key] += int(r[ci[key]])
        else:
            cur = dict(start=k, end=k, ex=ex[k], n=1, s=sm[k])
            for key in STALLS:
                cur[key] = int(r[ci[key]])
            runs.append(cur)
    for r in runs:
        if r['ex'] * r['n'] > 0.005 * tot or r['s'] > 0.01 * totS:
            top = sorted(((r[k], k[6:]) for k in STALLS if r[k] > 0.04 * max(r['s'], 1)), reverse=True)
            print(f"{r['start']:5d}-{r['end']:5d} n={r['n']:4d} exec={r['ex']:>11d} inst%={r['ex']*r['n']/tot*100:5.1f} "
                  f"samp%={r['s']/totS*100:5.1f} " + " ".join(f"{n}={v}" for v, n in top) +
                  f" | {data[r['start']][ci['Source']].strip()[:40]}")
    tots = {k: sum(int(r[ci[k]]) for r in data) for k in STALLS}
    print("stall samples total: " + " ".join(f"{k[6:]}={v} ({v/totS*100:.1f}%)" for k, v in sorted(tots.items(), key=lambda kv: -kv[1]) if v))

if __name__ == "__main__":
    main()
