"""Hottest source lines of a kernel in an ncu report: python tools/ncu_hot_lines.py REPORT KERNEL_REGEX [TOP] [inst]"""
import csv,sys,subprocess
rep,kern=sys.argv[1],sys.argv[2]
top=int(sys.argv[3]) if len(sys.argv)>3 else 30
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass','--kernel-name','regex:'+kern],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[2]
iSamp=hdr.index('# Samples'); iInst=hdr.index('Instructions Executed'); iThr=hdr.index('Thread Instructions Executed')
lines=[]
for r in rows[3:]:
    if r and r[0].strip().isdigit():
        try: lines.append((int(r[0]), r[1].strip()[:100], int(r[iSamp]), int(r[iInst]), int(r[iThr])))
        except: pass
tot_s=sum(l[2] for l in lines); tot_i=sum(l[3] for l in lines)
print('total samples',tot_s,'inst',tot_i)
key=2 if len(sys.argv)<5 else 3
for l in sorted(lines,key=lambda x:-x[key])[:top]:
    print(f"{l[0]:4d} samp {100*l[2]/tot_s:5.1f}% inst {100*l[3]/tot_i:5.1f}% thr/inst {l[4]/max(l[3],1):5.1f} | {l[1]}")
