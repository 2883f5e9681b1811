import re, sys
cur=None; rows={}
for line in open(sys.argv[1]):
    if line.startswith('=='): cur=line.split()[1]; rows[cur]=[]; continue
    m=re.match(r'\s+KC=(\S+) TPS=(\S+) MT=(\S+) NACC=(\S+): (\S+) us',line)
    if m: rows[cur].append((m.group(1),m.group(2),m.group(3),m.group(4),float(m.group(5))))
for k,v in rows.items():
    print(k)
    d={}
    for kc,tps,mt,na,us in v: d.setdefault((kc,tps,mt),{})[na]=us
    for key,val in d.items(): print("   KC=%s TPS=%s MT=%s :"%key, "  ".join("N%s=%.0f"%(n,u) for n,u in sorted(val.items())))
