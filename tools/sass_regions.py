"""Static SASS instruction counts per source-line region of one kernel (nvdisasm --print-line-info).
usage: python tools/sass_regions.py obj.o kernel_substring file.cu lo-hi[:name] ..."""
import collections
import re
import subprocess
import sys
import tempfile
import os


def main():
    obj, ksub, src = sys.argv[1:4]
    regions = []
    for r in sys.argv[4:]:
        rng, _, name = r.partition(':')
        lo, hi = rng.split('-')
        regions.append((name or rng, int(lo), int(hi)))
    d = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(d) if f.endswith('.cubin')][0]
    txt = subprocess.run(['nvdisasm', '--print-line-info', os.path.join(d, cub)], capture_output=True, text=True).stdout
    cur = None
    on = False
    cnt = collections.Counter()
    ops = collections.defaultdict(collections.Counter)
    for l in txt.split('\n'):
        if l.startswith('\t.section\t.text.') or l.startswith('.text.') or (l.startswith('$') and l.rstrip().endswith(':') and '$_Z' in l[1:]):
            on = ksub in l
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', l)
        if m:
            cnt[cur] += 1
            ops[cur][m.group(1)] += 1
    print('total', sum(cnt.values()))
    for name, lo, hi in regions:
        tot = 0
        oc = collections.Counter()
        for (f, ln), c in cnt.items():
            if f == src and lo <= ln <= hi:
                tot += c
                oc.update(ops[(f, ln)])
        print(name, tot, oc.most_common(12))
    oth = collections.Counter()
    for (f, ln), c in cnt.items():
        if f != src:
            oth[(f, ln)] += c
    print('other files', oth.most_common(8))


main()
