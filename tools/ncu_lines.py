"""Per-source-line summary of an ncu report captured with `--set full --import-source on` (kernels built with -lineinfo).

usage: python tools/ncu_lines.py report.ncu-rep [kernel-substring] [top-N]

Reads `ncu -i report --page source --csv --print-source cuda,sass`, keeps the first launch whose function name contains
the substring, and prints (a) the share of executed warp instructions and of stall samples per source line, (b) the same
aggregated per file and per 'region' (consecutive source lines are bucketed by the enclosing function-like header found
upwards in the file), (c) the stall-reason totals.  Used for profiles/*_lines.md.
"""
import collections
import csv
import subprocess
import sys


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    per_line = collections.OrderedDict()   # (file, line) -> [inst, samples, text, thread_inst]
    stalls = collections.Counter()
    cur_file, cur_fn, hdr, idx, take = None, None, None, None, False
    seen_fn = set()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            cur_fn = r[1]
            take = want in cur_fn
            continue
        if r[0] == "Line No":
            hdr = r
            idx = {}
            for i, h in enumerate(hdr):
                idx.setdefault(h, i)
            continue
        if not take or hdr is None or len(r) < len(hdr):
            continue
        if r[0] == "":
            continue   # SASS row (already aggregated in its source-line row)
        key = (cur_fn, cur_file, int(r[0]))
        inst = num(r[idx["Instructions Executed"]])
        samp = num(r[idx["# Samples"]])
        tinst = num(r[idx["Thread Instructions Executed"]])
        per_line[key] = [inst, samp, r[1].strip(), tinst]
        for h in hdr:
            if h.startswith("stall_") and "Not Issued" not in h:
                stalls[h] += num(r[idx[h]])
    # keep the first matching function only
    fns = []
    for k in per_line:
        if k[0] not in fns:
            fns.append(k[0])
    if not fns:
        print("no function matches", want)
        return
    fn = fns[0]
    lines = {(k[1], k[2]): v for k, v in per_line.items() if k[0] == fn}
    tot_i = sum(v[0] for v in lines.values()) or 1
    tot_s = sum(v[1] for v in lines.values()) or 1
    tot_t = sum(v[3] for v in lines.values()) or 1
    print(f"kernel: {fn}")
    print(f"warp instructions executed: {tot_i}   stall samples: {tot_s}   avg active threads: {tot_t / tot_i:.1f}")
    print("\nper file:")
    pf = collections.Counter(); ps = collections.Counter()
    for (f, l), v in lines.items():
        pf[f] += v[0]; ps[f] += v[1]
    for f, c in pf.most_common():
        print(f"  {f:24s} {c / tot_i * 100:5.1f}% inst  {ps[f] / tot_s * 100:5.1f}% samples")
    print(f"\ntop {top} lines by stall samples:")
    for (f, l), v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"  {v[1] / tot_s * 100:5.1f}% smp {v[0] / tot_i * 100:5.1f}% inst  {f}:{l:<5d} {v[2][:110]}")
    print(f"\ntop {top} lines by instructions:")
    for (f, l), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {v[0] / tot_i * 100:5.1f}% inst {v[1] / tot_s * 100:5.1f}% smp  {f}:{l:<5d} {v[2][:110]}")
    # buckets of 25 source lines per file
    print("\nper 20-line bucket (>= 1% of samples):")
    bk = collections.Counter(); bi = collections.Counter()
    for (f, l), v in lines.items():
        bk[(f, l // 20 * 20)] += v[1]; bi[(f, l // 20 * 20)] += v[0]
    for (f, b), c in sorted(bk.items()):
        if c / tot_s >= 0.01:
            print(f"  {f}:{b:4d}-{b + 19:<4d} {c / tot_s * 100:5.1f}% smp {bi[(f, b)] / tot_i * 100:5.1f}% inst")
    # note: the stall counters above were summed over every matching function instance; report shares only
    ts = sum(stalls.values()) or 1
    print("\nstall reasons (share of samples, all matching launches):")
    for h, c in stalls.most_common(8):
        print(f"  {h:28s} {c / ts * 100:5.1f}%")


if __name__ == "__main__":
    main()
