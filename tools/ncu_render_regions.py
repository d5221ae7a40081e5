"""Per-region sampling summary of render_kernel from an ncu report: python tools/ncu_render_regions.py REPORT [render.cu as profiled]"""
import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass','--kernel-name','regex:render_kernel'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[2]
iSamp=hdr.index('# Samples'); iInst=hdr.index('Instructions Executed'); iThr=hdr.index('Thread Instructions Executed')
lines={}
cur=None
files={}
for r in rows:
    if r and r[0]=='File Path': curfile=r[1]
    if r and r[0].strip().isdigit():
        try:
            k=(curfile.split('/')[-1],int(r[0]))
            v=lines.get(k,(0,0,0))
            lines[k]=(v[0]+int(r[iSamp]), v[1]+int(r[iInst]), v[2]+int(r[iThr]))
        except: pass
tot_s=sum(l[0] for l in lines.values()); tot_i=sum(l[1] for l in lines.values())
print('total inst',tot_i)
def rng(f,a,b,name):
    s=sum(v[0] for k,v in lines.items() if k[0]==f and a<=k[1]<=b); i=sum(v[1] for k,v in lines.items() if k[0]==f and a<=k[1]<=b); t=sum(v[2] for k,v in lines.items() if k[0]==f and a<=k[1]<=b)
    print(f"{name:28s} {f}:{a}-{b}: samples {100*s/tot_s:5.1f}% inst {100*i/tot_i:5.1f}% lanes/inst {t/max(i,1):.1f}")
src=open(sys.argv[2] if len(sys.argv)>2 else '/root/repo/optical-flow-2d-data-generation_b200/csrc/render.cu').read().split('\n')  # the source the profiled binary was built from
def find(pat):
    for n,l in enumerate(src,1):
        if pat in l: return n
    raise KeyError(pat)
marks=[('helpers(iround/reflect/mirror)',find('int iround_d')),('Dda2+RowWarp',find('struct Dda2')),('bilinear',find('uint32_t bilinear_rgbx')),('blend',find('uint32_t blend_rgbx')),('mode9 helpers',find('// mode 9: non-rigid warp fields')),('box_hits/bin',find('bool box_hits_tile')),('comp',find('unsigned comp_add')),('kernel head',find('render_kernel(RenderArgs a)')),('pass_setup',find('auto pass_setup')),('bg fetch',find('// ---- background: masks')),('pass loop head',find('// ---- foreground objects in z-order')),('chunk: zero+edges(a)',find('// (3) chunks')),('chunk: edges(b)',find('// (b) threads over (edge, row) items')),('consume masks',find('for (int j = j0; j < j1; ++j) {')),('composite combine',find('const HitObject ho = s_hit[jb.hit];')),('blit',find("// the object's masks are complete")),('flow',find('// ---- flow of the top-most')),('write',find('// ---- write the three blobs')),('end',find('// mode 9 pre-pass'))]
for (n,a),(m,b) in zip(marks,marks[1:]): rng('render.cu',a,b-1,n)
rng('raster_tile.h',1,400,'raster_tile.h')
others=set(k[0] for k in lines)-{'render.cu','raster_tile.h'}
for f in others: rng(f,1,100000,'other')
