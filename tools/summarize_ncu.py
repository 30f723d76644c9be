"""Summarise ncu outputs into profiles/: launch list -> per-kernel share table; .ncu-rep -> key metrics per launch.
  python tools/summarize_ncu.py gpurun_out/launches_r01.csv gpurun_out/prof_tc_r01.ncu-rep r01"""
import collections
import csv
import subprocess
import sys

launch_csv, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
lines = [l for l in open(launch_csv) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v)
    k = row['Kernel Name'].split('(')[0][:70]
    agg[k][0] += 1; agg[k][1] += v; tot += v
with open(f'profiles/launches_{tag}.md', 'w') as f:
    f.write(f"# ncu launch list, one full-size ASE minibatch update (B=16384, Ba=4096) -- {tag}\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none` over `tools/profile_minibatch.py` (one minibatch = "
            f"{sum(n for n, _ in agg.values())} launches; cold-cache, serialised: compare SHARES).\n\n| kernel | launches | us | share |\n|---|---:|---:|---:|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f}% |\n")
    f.write(f"| **total** | | {tot:.1f} | |\n")
print(open(f'profiles/launches_{tag}.md').read())

raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units = r[0], r[1]
want = ['Kernel Name', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__cycles_active.avg', 'launch__shared_mem_per_block_dynamic',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard', 'sm__cycles_elapsed.max']
idx = [(h, i) for i, h in enumerate(hdr) if h in want]
with open(f'profiles/ncu_tc_{tag}.md', 'w') as f:
    f.write(f"# ncu --set full, gemm_tc_kernel launches inside one ASE minibatch -- {tag}\n\n")
    for row in r[2:]:
        f.write("```\n")
        for h, i in idx:
            f.write(f"{h:72s} {units[i]:16s} {row[i]}\n")
        f.write("```\n")
print(open(f'profiles/ncu_tc_{tag}.md').read()[:6000])
