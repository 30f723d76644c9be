"""Summarise an ncu CSV (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch) into a per-kernel table:
launches, mean time, mean DRAM bytes, achieved DRAM GB/s against MEASURED_PEAKS.json hbm_gbs, and -- where given in ALGO below --
the ALGORITHMIC bytes per launch at config-3 sizes and the GB/s they imply.   python tools/summarize_hbm.py in.csv out.md [title]"""
import csv, json, os, sys, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N, H, B, Ba, Z, A, OBS, AMP = 4096, 32, 16384, 4096, 64, 31, 253, 1400
ARENA = 7.04e6
# algorithmic bytes per launch (config 3); see DESIGN.md section 4 for the derivations
ALGO = {
    'obs_build_kernel': N * (17 * 13 + OBS) * 4,
    'amp_obs_build_kernel': N * ((13 + 31 + 31 + 18) * 4 + 140 * 4),            # ring history: one 140-float frame written per env
    'gae_kernel': H * N * (3 * 4 + 1 + 2 * 4),
    'amp_rewards_kernel': H * N * (4 + 2 * Z * 4 + 3 * 4),
    'adam_kernel': ARENA * 7 * 4,
    'gather_rows_kernel': 2 * 4 * (B * (OBS + Z + 3 * A + 5) + 3 * Ba * AMP),
    'policy_sample_kernel': N * A * 4 * 5,
}


def main():
    src, dst = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else 'HBM-bound kernels'
    peak = 6571.9
    try:
        peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    rows = [r for r in csv.reader(l for l in open(src, errors='replace') if l.startswith('"'))]
    hdr = rows[0]
    iK, iM, iV, iU, iID = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('ID')
    per = collections.OrderedDict()
    for r in rows[1:]:
        k = (r[iID], r[iK])
        v = float(r[iV].replace(',', ''))
        u = r[iU].lower()
        m = r[iM]
        if 'time' in m:
            v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3, 'second': 1e6}.get(u, 1.0)
        else:
            v *= {'byte': 1.0, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1.0)
        per.setdefault(k, {})[m] = v
    agg = collections.OrderedDict()
    for (_, name), m in per.items():
        short = name.split('(')[0].split('::')[-1].split('<')[0].strip()
        a = agg.setdefault(short, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += m.get('gpu__time_duration.sum', 0.0)
        a[2] += m.get('dram__bytes_read.sum', 0.0); a[3] += m.get('dram__bytes_write.sum', 0.0)
    tot = sum(a[1] for a in agg.values())
    with open(dst, 'w') as f:
        f.write(f"# {title}\n\nncu `gpu__time_duration.sum`, `dram__bytes_read.sum`, `dram__bytes_write.sum`, `--clock-control none` (per-launch times are cold-cache and "
                f"serialised).  Peak = MEASURED_PEAKS.json `hbm_gbs` = {peak:.1f} GB/s (of measured).  `algo` = algorithmic bytes per launch at config-3 sizes.\n\n")
        f.write("| kernel | launches | mean us | share | DRAM MB/launch (rd + wr) | DRAM GB/s | frac of peak | algo MB | algo GB/s | algo frac |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            n, t, rd, wr = a
            us = t / n
            gbs = (rd + wr) / n / (us * 1e-6) / 1e9 if us > 0 else 0.0
            al = ALGO.get(k)
            f.write(f"| `{k}` | {n} | {us:.1f} | {100 * t / tot:.1f}% | {rd / n / 1e6:.2f} + {wr / n / 1e6:.2f} | {gbs:.0f} | {gbs / peak:.2f} | "
                    + (f"{al / 1e6:.2f} | {al / (us * 1e-6) / 1e9:.0f} | {al / (us * 1e-6) / 1e9 / peak:.2f} |\n" if al else "| | |\n"))
    print(open(dst).read())


if __name__ == '__main__':
    main()
