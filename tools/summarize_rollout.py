"""profiles/rollout_launches_<tag>.md from the ncu launch list of tools/profile_rollout.py:  python tools/summarize_rollout.py <tag> <csv> [note]"""
import csv, collections, sys
tag, src = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ''
rows = list(csv.DictReader(l for l in open(src, errors='replace') if l.startswith('"')))
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    v = v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v)
    n = r['Kernel Name'].split('(')[0].replace('void ', '')[:90]
    tot[n] += v; cnt[n] += 1
T = sum(tot.values())
with open(f'profiles/rollout_launches_{tag}.md', 'w') as f:
    f.write(f"# ncu launch list of one config-3 rollout (4096 envs x 32 steps, reward pass over 131072 rows, GAE), launched eagerly -- {tag}\n\n"
            "`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none` over `tools/profile_rollout.py` (cudaProfilerStart / Stop around one "
            f"play_steps with the CUDA graph off).  {len(rows)} launches, {T / 1e3:.2f} ms summed (cold-cache, serialised: compare SHARES).  {note}\n\n"
            "| kernel | launches | us | us / launch | share |\n|---|---:|---:|---:|---:|\n")
    for n, v in tot.most_common():
        f.write(f"| `{n}` | {cnt[n]} | {v:.1f} | {v / cnt[n]:.1f} | {100 * v / T:.1f}% |\n")
print(open(f'profiles/rollout_launches_{tag}.md').read()[:600])
