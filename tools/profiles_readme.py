"""Regenerates the tables of profiles/README.md from the committed evidence files of a round (bench JSON lines, ncu launch list,
ncu metric summary, full-config RMSE lines), so that the README cannot drift from the files it indexes.

  python tools/profiles_readme.py r02 > /tmp/tables.md
"""
import collections
import csv
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def load(path):
    for ln in open(path):
        if ln.startswith("{"):
            return json.loads(ln)
    raise SystemExit(path + ": no JSON line")


def bench_row(tag, path):
    l = load(path)
    r = l["roofline"]
    cpu = l.get("cpu_baseline") or {}
    return (f"| `{os.path.basename(path)}` | {l['config']['workload']} | {l['n_gpus']} | {l['value']:.0f} | {l['ms_per_step']:.1f} | {l['e2e']['value']:.0f} | "
            f"{r['trace_gray_per_s']} | {r['frac']} | {r['bytes_per_ray']} | {r['trace_share_of_step']} / {r['shade_share_of_step']} | "
            f"{cpu.get('value', '—')} ({cpu.get('cores', '—')} thr) | {l['frame_crc32']} | {l['clocks']['sm_mhz']:.0f} {l['clocks']['reasons']} |")


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    out = ["| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        out.append(f"| `{k}` | {cnt[k]} | {v / 1e6:.2f} | {100 * v / T:.1f}% |")
    return "\n".join(out)


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"
    print("### bench lines\n")
    print("| file | workload | GPUs | value Mray/s | ms/step | e2e Mray/s | K2 Gray/s | K2 algorithmic frac | B_ray | trace / shade share | CPU reference Mray/s | frame CRC | SM MHz, throttle reasons |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for path in sorted(glob.glob(os.path.join(P, rnd + "_bench_*.json")) + glob.glob(os.path.join(P, rnd + "_scale_*.json"))):
        if "reference_arm" in path:
            continue
        print(bench_row("", path))
    for path in sorted(glob.glob(os.path.join(P, rnd + "_ncu_launches_*.csv"))):
        print(f"\n### ncu launch list `{os.path.basename(path)}`\n")
        print(launches(path))
    mpath = os.path.join(P, rnd + "_ncu_metrics.json")
    if os.path.exists(mpath):
        m = json.load(open(mpath))["kernels"]
        keys = ["duration", "active_threads_per_warp_inst", "issue_active_pct", "warps_active_pct", "registers", "l1_hit_pct", "l2_hit_pct", "dram_read", "dram_write",
                "stall_long_scoreboard", "stall_wait", "stall_branch", "stall_not_selected", "warp_instructions"]
        print(f"\n### ncu --set full `{os.path.basename(mpath)}`\n")
        print("| capture | " + " | ".join(keys) + " |")
        print("|---|" + "---|" * len(keys))
        for label, ks in m.items():
            k = ks[0]
            print(f"| {label}: `{k.get('kernel', '?').split('(')[0].replace('void ', '')}` | " + " | ".join(
                (f"{k[x]:.3g}" if isinstance(k.get(x), float) else str(k.get(x, "—"))) for x in keys) + " |")
    rpath = os.path.join(P, rnd + "_full_config_rmse.jsonl")
    if os.path.exists(rpath):
        print(f"\n### full-size parity `{os.path.basename(rpath)}`\n")
        print("| configuration | RMSE | bound | max abs diff | bit-identical pixels | non-finite GPU / reference | rays |")
        print("|---|---|---|---|---|---|---|")
        for ln in open(rpath):
            d = json.loads(ln)
            print(f"| {d['config']} | {d['rmse']:.3e} | {d.get('bound', 1e-4):.2e} | {d['max_abs_diff']:.2e} | {d['pixels_bit_identical']} of {d['pixels']} | "
                  f"{d['nonfinite_gpu']} / {d['nonfinite_reference']} | {d['rays']:.3e} |")


if __name__ == "__main__":
    main()
