"""Stall samples of one kernel aggregated by device sub-function (the $kernel$callee labels of nvdisasm) and by
caller-supplied source-line regions: joins `ncu --page source --csv` with `nvdisasm -g` by instruction offset.
    python tools/ncu_regions.py <ncu source csv> <nvdisasm -g listing> <kernel symbol> file.cu lo-hi:name ..."""
import csv, re, sys, collections
src_csv, sass, sym, srcfile = sys.argv[1:5]
regions = []
for r in sys.argv[5:]:
    rng, _, name = r.partition(':'); lo, hi = rng.split('-'); regions.append((name, int(lo), int(hi)))
info, cur, active, fn = {}, None, False, 'kernel'
for ln in open(sass):
    if ln.startswith('.text.'):
        active = (ln.strip().rstrip(':') == '.text.' + sym); fn = 'kernel'; continue
    if not active: continue
    if ln.startswith('$') and ln.rstrip().endswith(':') and '$_Z' in ln[1:]:
        m = re.search(r'\$_ZN4srcb4fast\d+([a-z_0-9]+?)I', ln); fn = m.group(1) if m else ln.strip()[-40:]; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', ln)
    if m: info[int(m.group(1), 16)] = (fn, cur, m.group(2))
rows = list(csv.reader(open(src_csv))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
base = None
S = collections.Counter(); E = collections.Counter(); ST = collections.defaultdict(collections.Counter)
def region_of(fn, cur):
    if cur and cur[0] == srcfile:
        for name, lo, hi in regions:
            if lo <= cur[1] <= hi: return name
    return fn + ':other'
last_region = {}
for r in rows[2:]:
    if len(r) < len(hdr): continue
    a = int(r[ix['Address']], 16)
    if base is None: base = a
    fn, cur, op = info.get(a - base, ('?', None, '?'))
    # inlined library lines (shuffles, redux ...) inherit the region of the last own-source line of that function
    if cur and cur[0] == srcfile:
        key = region_of(fn, cur); last_region[fn] = key
    else:
        key = last_region.get(fn, fn + ':other')
    s = int(r[ix['# Samples']] or 0); e = int(r[ix['Instructions Executed']] or 0)
    S[key] += s; E[key] += e
    for h in stalls:
        v = int(r[ix[h]] or 0)
        if v: ST[key][h[6:]] += v
tot = sum(S.values()); Et = sum(E.values())
print("total samples %d, warp instructions %.3e" % (tot, Et))
for key, s in S.most_common():
    st = ", ".join("%s %.0f%%" % (k, 100.0 * v / max(s, 1)) for k, v in ST[key].most_common(4))
    print("%-26s samples %5.2f%%  instr %5.2f%% (%.3e)  | %s" % (key, 100.0 * s / tot, 100.0 * E[key] / Et, E[key], st))
