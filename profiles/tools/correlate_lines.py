import csv, re, sys, collections
rep_csv = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else 'ILb1ELb0'  # template arguments of the captured instantiation
src = open('/root/repo/gnn_tracking_b200/csrc/mlp_tc.cu').read().split('\n')
# offset -> line from nvdisasm
off2line = {}
cur = None
infn = False
for l in open('/tmp/tc_disasm.txt'):
    if '.section' in l and '.text.' in l:
        infn = ('fused_mlp_tc_kernel' + kern) in l
        continue
    m = re.search(r'//## File ".*mlp_tc.cu", line (\d+)', l)
    if m: cur = int(m.group(1)); continue
    if '//## File' in l:
        m2 = re.search(r'//## File "(.*)", line (\d+)', l)
        cur = -1; continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', l)
    if m and infn: off2line[int(m.group(1),16)] = cur
rows = list(csv.reader(open(rep_csv)))
hdr = rows[1]; data = rows[2:]
si = hdr.index('# Samples'); ai = hdr.index('Address')
base = int(data[0][ai],16)
stall_cols = [(h, hdr.index(h)) for h in hdr if h.startswith('stall_') and 'Not' not in h]
per_line = collections.Counter(); per_line_st = collections.defaultdict(collections.Counter)
tot = 0
for r in data:
    n = int(r[si]); tot += n
    off = int(r[ai],16) - base
    ln = off2line.get(off, None)
    per_line[ln] += n
    for h, i in stall_cols:
        v = int(r[i])
        if v: per_line_st[ln][h[6:]] += v
print('total samples', tot)
for ln, n in per_line.most_common(45):
    st = per_line_st[ln].most_common(3)
    text = src[ln-1].strip()[:90] if ln and ln > 0 else str(ln)
    print(f"{n:6d} {100*n/tot:5.1f}%  L{ln}: {text}   {st}")
