import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=rows[0]
i_name=hdr.index('Kernel Name'); i_val=hdr.index('Metric Value'); i_unit=hdr.index('Metric Unit'); i_grid=hdr.index('Grid Size')
tot=0; out=[]
for r in rows[1:]:
    v=float(r[i_val].replace(',','')); u=r[i_unit]
    if u=='ns': v/=1e3
    elif u=='ms': v*=1e3
    tot+=v
    nm=r[i_name].split('::')[-1][:22]
    out.append("%s %s %.1f"%(nm,r[i_grid].replace(' ',''),v))
n=int(sys.argv[2]) if len(sys.argv)>2 else len(out)
print(" | ".join(out[:n])); print("total",tot)
