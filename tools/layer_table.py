"""Developer tool: per-layer table (duration, TFLOP/s, tensor / HBM lower bounds) from an ncu launch list."""
import csv, sys
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches_trunk.csv'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rows=[]
with open(path) as f:
    lines=[l for l in f if not l.startswith('==')]
for row in csv.DictReader(lines):
    if row.get('Metric Name')=='gpu__time_duration.sum':
        rows.append((row['Kernel Name'], row['Grid Size'], float(row['Metric Value'].replace(',','')), row['Metric Unit']))
planes=[64,128,256,512]; blocks=[3,4,6,3]; inpl=64; hw=56
order=[]
for l in range(4):
    for b in range(blocks[l]):
        s=2 if (b==0 and l>0) else 1
        w=planes[l]
        c1=('c1',inpl,w,1,1,hw); c2=('c2',w,w,3,s,hw); hwo=hw//s; c3=('c3',w,w*4,1,1,hwo)
        ex=[c1,c2]
        if b==0: ex.append(('ds',inpl,w*4,1,s,hw))
        ex.append(c3)
        order+= [(l+1,b)+e for e in ex]
        inpl=w*4; hw=hwo
tot=0; ideal=0; i=0; groups={}
for name,grid,val,unit in rows:
    us = val/1000
    if 'maxpool' in name or 'avgpool' in name:
        print(f"{name[:20]:34s} {us:8.1f} us"); tot+=us; groups['pool']=groups.get('pool',0)+us; continue
    if i==0:
        fl=2*B*112*112*64*147; hbm=(B*230*230*8+B*112*112*64*2); tag='stem'; role='stem'
    else:
        l,b,role,cin,cout,k,s,hwin=order[i-1]
        ho=hwin//s
        fl=2*B*ho*ho*cout*cin*k*k
        inb=B*(hwin*hwin if s==1 else ho*ho*(1 if k==1 else 4))*cin*2
        hbm=inb+B*ho*ho*cout*2*(2 if role=='c3' else 1)
        tag=f"L{l}.{b}.{role} {cin}->{cout} k{k}s{s} @{hwin}"
    i+=1
    tf=fl/us/1e6; t_t=fl/1354e6; t_h=hbm/6.55e6
    print(f"{tag:34s} {us:8.1f} us {tf:7.1f} TF/s  tensor-min {t_t:6.1f}  hbm-min {t_h:6.1f}  x{us/max(t_t,t_h):.2f}")
    tot+=us; ideal+=max(t_t,t_h); groups[role]=groups.get(role,0)+us
print("total us %.0f ; sum of per-layer max(tensor,hbm) %.0f"%(tot,ideal)); print({k:round(v) for k,v in groups.items()})
