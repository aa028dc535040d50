"""Per-source-line stall samples of one kernel: joins `ncu --page source --csv` (SASS view: address, samples,
executed count) with `nvdisasm -g` line markers of the same cubin by instruction offset.
    python tools/ncu_lines.py <ncu source csv> <nvdisasm -g listing> <kernel symbol> [top]"""
import csv, re, sys, collections
src_csv, sass, sym = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
# offset -> line from nvdisasm
line_of, cur, active = {}, None, False
for ln in open(sass):
    if ln.startswith('.text.'):
        active = (ln.strip().rstrip(':') == '.text.' + sym)
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
base = None
samp = collections.Counter(); execd = collections.Counter(); stall = collections.defaultdict(collections.Counter)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    a = int(r[ix['Address']], 16)
    if base is None: base = a
    key = line_of.get(a - base)
    s = int(r[ix['# Samples']] or 0); e = int(r[ix['Instructions Executed']] or 0)
    samp[key] += s; execd[key] += e; tot += s
    for h in stalls:
        v = int(r[ix[h]] or 0)
        if v: stall[key][h[6:]] += v
E = sum(execd.values())
print("total samples %d, warp instructions %.3e" % (tot, E))
for key, s in samp.most_common(top):
    st = ", ".join("%s %.0f%%" % (k, 100.0 * v / max(s, 1)) for k, v in stall[key].most_common(3))
    print("%-22s samples %5.2f%%  instr %5.2f%%  | %s" % ("%s:%d" % key if key else "?", 100.0 * s / tot, 100.0 * execd[key] / E, st))
