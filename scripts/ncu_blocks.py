"""Basic-block summary of a scripts/ncu_sass_flow.py listing (consecutive instructions with equal execution counts): executions,
lanes active, share of instructions and of stall samples.  usage: ncu_blocks.py flow.txt [min inst %]"""
import sys,re
rows=[l.rstrip('\n') for l in open(sys.argv[1])][1:]
thresh=float(sys.argv[2]) if len(sys.argv)>2 else 0.15
blk=[];out=[]
def flush():
    if not blk: return
    ie=blk[0][1]; n=len(blk); thr=sum(b[2] for b in blk)/n; pct=sum(b[3] for b in blk); smp=sum(b[4] for b in blk)
    src=blk[0][5]; ops=' '.join(b[6].split()[0] if not b[6].startswith('@') else b[6].split()[1] for b in blk[:12])
    out.append((blk[0][0],n,ie,thr,pct,smp,src,ops))
prev=None
for l in rows:
    m=re.match(r'\s*(\d+)\s+([\d.]+)% ie=\s*(\d+) thr=\s*([\d.]+) pred=\s*([\d.]+) smp=\s*(\d+) \| (.*?)\s*\| (.*)',l)
    idx,pct,ie,thr,pred,smp,sass,src=int(m[1]),float(m[2]),int(m[3]),float(m[4]),float(m[5]),int(m[6]),m[7],m[8]
    if prev is not None and ie!=prev: flush(); blk=[]
    blk.append((idx,ie,thr,pct,smp,src.split(' < ')[-1] if False else src.split(' < ')[0],sass)); prev=ie
flush()
tot_smp=sum(o[5] for o in out)
for o in out:
    if o[4]>=thresh: print(f'{o[0]:5d} n={o[1]:3d} ie={o[2]:9d} thr={o[3]:5.1f} inst%={o[4]:5.2f} smp%={100*o[5]/tot_smp:5.2f} {o[6]:22s} {o[7][:100]}')
