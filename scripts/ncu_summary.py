#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics + per-code-region breakdown from the
source page.  Usage: python scripts/ncu_summary.py gpurun_out/prof_X.ncu-rep [> profiles/ncu_X_summary.txt]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__t_bytes.sum', 'lts__t_bytes.sum',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum', 'smsp__inst_executed_op_shared_ld.sum']
print("kernel:", rows[2][hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
if len(sys.argv) > 3 and sys.argv[2] == "--traffic-json":
    import json
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    val = lambda k: float(rows[2][hdr.index(k)].replace(",", "")) if k in hdr else None
    tot = sum(float(rows[2][hdr.index(k)]) * scale[rows[1][hdr.index(k)]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    val = lambda k: float(rows[2][hdr.index(k)]) if k in hdr else None
    # executed FP32 flops (thread-level FADD + FMUL + 2 x FFMA, predicated-on; an FFMA2 counts as two thread FFMAs) per sample of
    # the capture (argv[4] = samples in the profiled launch) against the algorithmic flops per sample (argv[5])
    n_samples = float(sys.argv[4]) if len(sys.argv) > 4 else None
    f_alg = float(sys.argv[5]) if len(sys.argv) > 5 else None
    fl = None
    if all(("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % k) in hdr for k in ("fadd", "fmul", "ffma")):
        fl = val("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum") + val("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum") + \
             2.0 * val("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum")
    with open(sys.argv[3], "w") as f:
        json.dump({"dram_bytes_per_launch": tot, "kernel": rows[2][hdr.index("Kernel Name")],
                   "executed_fp32_flops_per_launch": fl,
                   "executed_fp32_flops_per_sample": fl / n_samples if fl and n_samples else None,
                   "executed_over_algorithmic_flops": fl / n_samples / f_alg if fl and n_samples and f_alg else None,
                   "fp32_pipe_active_pct": val("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                   "fma_pipe_inst_pct": val("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                   "alu_pipe_inst_pct": val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                   "issue_slots_busy_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   "warp_execution_efficiency_lanes": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
                   "achieved_occupancy_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
                   "source": "dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` launch, " + rep.split("/")[-1]
                             + " (1920x1080, 64 spp; per-launch traffic does not depend on spp)"}, f, indent=1)
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:80s} {rows[1][i]:>12s} {rows[2][i]}")
stall = [(float(rows[2][i]), h) for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
print("stall reasons (warps per issue-active cycle):", ", ".join(f"{h[34:-24]}={v:.2f}" for v, h in sorted(stall, reverse=True)[:8]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; data = rows[2:]
ia, isrc, iw, it, ismp = (hdr.index(k) for k in ('Address', 'Source', 'Instructions Executed', 'Thread Instructions Executed', '# Samples'))
base = int(data[0][ia], 16)
recs = [(int(r[ia], 16) - base, r[isrc].strip(), int(r[iw]), int(r[it]), int(r[ismp])) for r in data]
totw = sum(r[2] for r in recs); tott = sum(r[3] for r in recs); tots = sum(r[4] for r in recs)
print(f"\nwarp instructions {totw:.4g}  thread instructions {tott:.4g}  active lanes/instr {tott / totw:.2f}")
seg = []; cur = None
for off, s, w, t, smp in recs:
    if cur and abs(cur['w0'] - w) <= 0.15 * max(cur['w0'], w, 1):
        cur['w'] += w; cur['t'] += t; cur['s'] += smp; cur['n'] += 1; cur['end'] = off
    else:
        cur = {'start': off, 'end': off, 'w': w, 't': t, 's': smp, 'n': 1, 'w0': w, 'first': s}; seg.append(cur)
print("code regions (>= 0.6 % of warp instructions):  range  n_instr  warp%  thread%  lanes  stall-samples%  executions/instr")
for s in seg:
    if s['w'] > 0.006 * totw:
        print(f"  {s['start']:#07x}-{s['end']:#07x} n={s['n']:4d} warp%={100 * s['w'] / totw:5.1f} thr%={100 * s['t'] / tott:5.1f} "
              f"lanes={s['t'] / max(1, s['w']):5.1f} stall%={100 * s['s'] / tots:5.1f} exec/inst={s['w0'] / 1e6:8.2f}M  {s['first'][:40]}")
