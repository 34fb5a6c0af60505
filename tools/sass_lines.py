"""Join an ncu source-page CSV (per-SASS counters) with nvdisasm -g line info: per source line totals.
usage: python tools/sass_lines.py <prof.ncu-rep> <cubin> <mangled kernel substring>"""
import csv, re, subprocess, sys, collections
rep, cubin, kname = sys.argv[1:4]
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(sass) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":")][0]
lines = []          # (offset, srcline) in order
cur = None
for l in sass[start + 1:]:
    if l.startswith("//-----") or (l.startswith(".text.") and l.rstrip().endswith(":")):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((int(m.group(1), 16), cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [r for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
assert len(data) == len(lines), (len(data), len(lines))
agg = collections.defaultdict(lambda: [0, 0])
tot = [0, 0]
for r, (off, src, txt) in zip(data, lines):
    agg[src][0] += int(r[ie]); agg[src][1] += int(r[isamp]); tot[0] += int(r[ie]); tot[1] += int(r[isamp])
print("total warp-instr %d samples %d" % tuple(tot))
for src, (e, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print("%-22s exec %5.1f%%  samples %5.1f%%" % ("%s:%d" % src if src else "?", 100 * e / tot[0], 100 * s / max(tot[1], 1)))

if len(sys.argv) > 4 and sys.argv[4] == "stalls":
    names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    idx = [hdr.index(h) for h in names]
    per = collections.defaultdict(lambda: collections.Counter())
    for r, (off, src, txt) in zip(data, lines):
        for h, i in zip(names, idx):
            try:
                per[src][h] += int(r[i])
            except Exception:
                pass
    print("--- by samples")
    for src, (e, s_) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        top = ", ".join("%s %d" % (k.replace("stall_", ""), v) for k, v in per[src].most_common(3))
        print("%-22s samples %5.1f%%  exec %5.1f%%   %s" % ("%s:%d" % src if src else "?", 100 * s_ / max(tot[1], 1), 100 * e / tot[0], top))
