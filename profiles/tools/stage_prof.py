import re,csv,collections,sys,subprocess,os
rep=sys.argv[1]; kern=sys.argv[2]  # e.g. ILb1ELb0
os.system(f"cd /tmp/xelf && rm -f *.cubin && cuobjdump -xelf all /root/repo/gnn_tracking_b200/csrc/libgtb200.so >/dev/null 2>&1 && nvdisasm -c -g mlp_tc.sm_100a.cubin > /tmp/tc_disasm.txt 2>/dev/null")
os.system(f"ncu -i {rep} --page source --csv 2>/dev/null > /tmp/tc_srcX.csv; ncu -i {rep} --page raw --csv 2>/dev/null > /tmp/tc_rawX.csv")
off2loc={}
cur=None; infn=False
for l in open('/tmp/tc_disasm.txt'):
    if '.section' in l and '.text.' in l:
        infn = ('fused_mlp_tc_kernel'+kern) in l
        continue
    m=re.search(r'//## File "(.*)", line (\d+)',l)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*);',l)
    if m and infn: off2loc[int(m.group(1),16)]=cur
rows=list(csv.reader(open('/tmp/tc_srcX.csv')))
hdr=rows[1]; data=rows[2:]
ie=hdr.index('Instructions Executed'); ai=hdr.index('Address'); base=int(data[0][ai],16); si=hdr.index('# Samples')
stall_cols=[(h[6:],hdr.index(h)) for h in hdr if h.startswith('stall_') and 'Not' not in h]
per=collections.Counter(); samp=collections.Counter(); st=collections.defaultdict(collections.Counter)
last=0
for r in data:
    loc=off2loc.get(int(r[ai],16)-base)
    if loc and loc[0]=='mlp_tc.cu': last=loc[1]
    per[last]+=int(r[ie]); samp[last]+=int(r[si])
    for h,i in stall_cols:
        if int(r[i]): st[last][h]+=int(r[i])
src=open('/root/repo/gnn_tracking_b200/csrc/mlp_tc.cu').read().split('\n')
marks=[(i+1) for i,l in enumerate(src) if 'TC_PROF(' in l and 'define' not in l]
fn={}
for i,l in enumerate(src):
    for n in ['int tc_copy_slot_row(','void tc_issue_item(const','void tc_issue_mmas(uint32_t','void tc_issue_mmas_k64(','void split_store8','void fused_mlp_tc_kernel']:
        if n in l: fn[n]=i+1
bounds=[(0,'pre')]+[(v,k) for k,v in fn.items()]+[(m,'after L%d %s'%(m,src[m-1].strip())) for m in marks]
bounds.sort()
tiles=7813
agg=collections.Counter(); sagg=collections.Counter(); stagg=collections.defaultdict(collections.Counter)
for ln,n in per.items():
    name=[b for a,b in bounds if ln>=a][-1]
    agg[name]+=n; sagg[name]+=samp[ln]
    for k,v in st[ln].items(): stagg[name][k]+=v
tot=sum(agg.values()); stot=sum(sagg.values())
for a,b in bounds:
    print(f"{b[:46]:46s} instr/tile/warp {agg[b]/tiles/8:6.0f} {100*agg[b]/tot:5.1f}%  samples {100*sagg[b]/stot:5.1f}%  {stagg[b].most_common(3)}")
print('instr per warp per tile', tot/tiles/8, 'samples', stot)
raw=list(csv.reader(open('/tmp/tc_rawX.csv')))
for h,u,v in zip(raw[0],raw[1],raw[2]):
    if h in ('gpu__time_duration.sum','smsp__inst_executed.sum','smsp__cycles_active.avg','sm__issue_active.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'): print(h,v,u)
