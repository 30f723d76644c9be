"""Turn the raw ncu outputs of a gpurun call into the tracked summaries under profiles/ (round 2: persistent CTA-pair kernel).
  python tools/summarize_pair_profiles.py <tag> <launch_csv> <ncu_fwd.ncu-rep> <ncu_dx.ncu-rep> <ncu_dw.ncu-rep>
writes profiles/launches_<tag>.md (+ .csv copy), profiles/ncu_tc_<tag>.md, profiles/roofline_traffic.json"""
import collections
import csv
import json
import re
import shutil
import subprocess
import sys

tag, launch_csv = sys.argv[1], sys.argv[2]
reps = dict(zip(('fwd', 'dx', 'dw'), sys.argv[3:6]))

lines = [l for l in open(launch_csv) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
adam = [i for i, r in enumerate(rows) if 'adam_kernel' in r['Kernel Name']]
mb = rows[adam[-2] + 1:adam[-1] + 1] if len(adam) >= 2 else rows
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in mb:
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    v = v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v)
    k = row['Kernel Name'].split('(')[0].replace('void ', '')[:80]
    agg[k][0] += 1; agg[k][1] += v; tot += v
with open(f'profiles/launches_{tag}.md', 'w') as f:
    f.write(f"# ncu launch list, one full-size ASE minibatch update (B=16384, Ba=4096, gemm_backend 2) -- {tag}\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none` over `tools/profile_minibatch.py 3 2`; the table is the LAST complete minibatch "
            f"(plane scales predicted, {len(mb)} launches). Per-launch times are cold-cache and serialised: compare SHARES.\n\n"
            "| kernel | launches | us | share |\n|---|---:|---:|---:|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f}% |\n")
    f.write(f"| **total** | {len(mb)} | {tot:.1f} | |\n")
shutil.copy(launch_csv, f'profiles/launches_{tag}.csv')

want = ['Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__cluster_dim_x', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second', 'sm__warps_active.avg.pct_of_peak_sustained_active']
what = {'fwd': 'actor layer 2 forward: M=32768 N=1024 K=1024, bias + ReLU, planes-only output + activity bits (512 work items on 74 CTA pairs)',
        'dx': 'its dX GEMM: M=32768 N=1024 K=1024, W read MN-major, activity-bit mask, fused bias-gradient column sums, planes-only output',
        'dw': 'a dW GEMM, both operands MN-major, split-K fp32 RED accumulation'}
traffic = {}
with open(f'profiles/ncu_tc_{tag}.md', 'w') as f:
    f.write(f"# ncu --set full --clock-control none, three gemm_tc2_kernel (persistent CTA pair, cta_group::2) launches inside a warm full-size minibatch -- {tag}\n\n"
            "Captured with `tools/ncu_capture_pair.sh` (kernel-name filter + launch-skip into the third minibatch). ncu replays each launch ~40 times with "
            "cold caches: durations here are NOT bench numbers (CUDA-event timings are in bench_*.json / experiments_*.md).\n")
    for key, rep in reps.items():
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        r = list(csv.reader(raw.splitlines()))
        hdr, units, val = r[0], r[1], r[2]
        f.write(f"\n## {key}: {what[key]}\n\n```\n")
        kn = val[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
        f.write(f"{'Kernel Name':64s} {kn}\n")
        rec = {}
        for h in want:
            if h in hdr:
                i = hdr.index(h); f.write(f"{h:64s} {val[i]} {units[i]}\n"); rec[h] = val[i]
        f.write("```\n")
        mul = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}
        rd = float(rec.get('dram__bytes_read.sum', '0').replace(',', '')) * mul.get(units[hdr.index('dram__bytes_read.sum')], 1.0)
        wr = float(rec.get('dram__bytes_write.sum', '0').replace(',', '')) * mul.get(units[hdr.index('dram__bytes_write.sum')], 1.0)
        traffic[key] = {'grid': rec.get('Grid Size'), 'dram_read_bytes': rd, 'dram_write_bytes': wr, 'us': float(rec.get('gpu__time_duration.sum', '0').replace(',', ''))}
        src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
        rows2 = list(csv.reader(src.splitlines()))
        h2 = rows2[1]; data = rows2[2:]
        iS, isrc = h2.index('Warp Stall Sampling (All Samples)'), h2.index('Source')
        ld = [i for i, rr in enumerate(data) if 'LDTM' in rr[isrc]]
        bars = [i for i, rr in enumerate(data) if 'BAR.SYNC' in rr[isrc]]
        total = sum(int(rr[iS]) for rr in data) or 1
        cuts = [('prologue, TMA producer, MMA issuer (3 of 20 warps)', 0, ld[0] - 80), ('drain loop: wait for a k-block partial, tcgen05.ld, fp32 adds', ld[0] - 80, ld[-1] + 60),
                ('store phase (row-layout math, staging, global stores)', ld[-1] + 60, bars[-2]), ('teardown + out-of-line mbarrier wait loops (idle warps park here)', bars[-2], len(data))]
        f.write("\nwarp samples by phase (source page):\n\n")
        for name, a, b in cuts:
            f.write(f"* {name}: {100 * sum(int(rr[iS]) for rr in data[a:b]) / total:.1f}%\n")
        mn = collections.Counter()
        for rr in data:
            tok = [x for x in rr[isrc].split() if not x.startswith('@')]
            if tok and re.match(r'(UTCHMMA|UTMALDG|LDTM|UTCBAR|SYNCS|REDG|USETMAXREG)', tok[0]):
                mn['.'.join(tok[0].split('.')[:3])] += 1
        f.write("\nSASS mnemonics (static counts): " + ", ".join(f"`{k}` x{v}" for k, v in sorted(mn.items())) + "\n")
dom = traffic.get('fwd', {})
json.dump({'kernel': 'gemm_tc2_kernel<K-major A, K-major B> (actor layer 2 forward, the largest single launch shape)',
           'source': f'profiles/ncu_tc_{tag}.md (ncu --set full --clock-control none)',
           'dram_bytes_per_launch': dom.get('dram_read_bytes', 0) + dom.get('dram_write_bytes', 0), 'launches': traffic},
          open('profiles/roofline_traffic.json', 'w'), indent=1)
print(open(f'profiles/ncu_tc_{tag}.md').read()[:3000])
