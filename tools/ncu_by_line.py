"""Per-source-line instruction counts of one kernel: joins the SASS page of an .ncu-rep (executed instructions per
address) with the line table of the built library (nvdisasm -g), by instruction offset inside the kernel.

    python tools/ncu_by_line.py gpurun_out/x.ncu-rep kmc_run_kernel [top]
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(kernel):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "latticemontecarlo_b200", "liblmc_b200.so")], cwd=d, capture_output=True)
    sass = []
    for cub in sorted(f for f in os.listdir(d) if f.endswith(".cubin")):       # one cubin per translation unit
        text = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
        if kernel in text:
            sass = text.splitlines()
            break
    table, cur, inside = {}, None, False
    for ln in sass:
        if ln.startswith("\t.section\t.text."):
            inside = kernel in ln
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            # keep the innermost "inlined at" chain head only
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    by_source_order = len(sys.argv) > 4            # 4th argument: divisor (e.g. warp-rounds) -> listing in source order, counts / divisor
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = int(rows[2][ia], 16)
    table = line_table(kernel)
    per_line, samples = collections.Counter(), collections.Counter()
    total = 0
    for r in rows[2:]:
        off = int(r[ia], 16) - base
        n = int(r[ie] or 0)
        loc = table.get(off, (None, ""))[0]
        per_line[loc] += n
        samples[loc] += int(r[isamp] or 0)
        total += n
    tot_s = sum(samples.values())
    print("total warp instructions %d, samples %d" % (total, tot_s))
    if by_source_order:
        div = float(sys.argv[4])
        for loc, n in sorted(per_line.items(), key=lambda kv: kv[0] or ("", 0)):
            if n / div < 1.0:
                continue
            src = ""
            if loc:
                try:
                    src = open(os.path.join(ROOT, "latticemontecarlo_b200", "csrc", loc[0])).read().splitlines()[loc[1] - 1].strip()
                except Exception:
                    pass
            print("%8.1f inst %6.2f%% samp  %s:%s  %s" % (n / div, 100.0 * samples[loc] / max(tot_s, 1), loc[0] if loc else "?", loc[1] if loc else "?", src[:110]))
        return
    for loc, n in per_line.most_common(top):
        src = ""
        if loc:
            try:
                src = open(os.path.join(ROOT, "latticemontecarlo_b200", "csrc", loc[0])).read().splitlines()[loc[1] - 1].strip()
            except Exception:
                pass
        print("%6.2f%% inst %6.2f%% samp  %s:%s  %s" % (100.0 * n / total, 100.0 * samples[loc] / max(tot_s, 1), loc[0] if loc else "?", loc[1] if loc else "?", src[:110]))


if __name__ == "__main__":
    main()
