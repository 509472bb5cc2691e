"""Static count of local-memory (spill) instructions of a kernel by source-line range, from `nvdisasm --print-line-info`."""
import re, collections, sys
def spills(path, kern, fname):
    txt = open(path).read().split('\n')
    start = [i for i, l in enumerate(txt) if l.startswith('\t.section\t.text.' + kern)][0]
    cur = None; cnt = collections.Counter()
    for l in txt[start + 1:]:
        if l.startswith('\t.section'): break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            cur = int(m.group(2)) if m.group(1).endswith(fname) else -1; continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m and ('LDL' in m.group(2) or 'STL' in m.group(2)): cnt[cur] += 1
    return cnt
if __name__ == "__main__":
    path, fname = sys.argv[1], sys.argv[2]
    ranges = [tuple(int(v) for v in a.split('-')) for a in sys.argv[4:]]
    for kern in sys.argv[3].split(','):
        c = spills(path, kern, fname)
        print(kern, "total", sum(c.values()), " ".join("%d-%d:%d" % (a, b, sum(v for k, v in c.items() if a <= k <= b)) for a, b in ranges))
