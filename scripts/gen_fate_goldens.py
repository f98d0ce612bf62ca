"""Collects, from the reference tree, the FATE md5 sums this repository can check (tests/golden/fate_pixfmts.txt).
Run in the build container only (reads /root/reference); the output file is committed."""
import re, os
R='/root/reference/tests/ref/fate/'
fmts=['rgb24','bgr24','rgba','bgra','argb','abgr','yuv420p','nv12','nv21']
rows=[]
def find(fn, key):
    for i,l in enumerate(open(R+fn),1):
        p=l.split()
        if p and p[0]==key: return i,p[1]
    raise KeyError((fn,key))
for t,size in [('null','352x288'),('copy','352x288'),('vflip','352x288'),('hflip','352x288'),('crop','100x100')]:
    for f in fmts:
        ln,m=find('filter-pixfmts-'+t,f); rows.append((f'tests/ref/fate/filter-pixfmts-{t}:{ln}',t,f,size,1,m))
for f in ('yuv420p','rgb24','bgr24','nv12','nv21'):          # the second scaler reads what the first wrote: needs that format as a SOURCE
    ln,m=find('filter-pixfmts-scale',f); rows.append((f'tests/ref/fate/filter-pixfmts-scale:{ln}','scale',f,'200x100',1,m))
for f in fmts:
    ln,m=find('filter-pixdesc-'+f,'pixdesc-'+f); rows.append((f'tests/ref/fate/filter-pixdesc-{f}:{ln}','pixdesc',f,'352x288',5,m))
with open('/root/repo/tests/golden/fate_pixfmts.txt','w') as o:
    o.write('''# md5 sums the reference tree commits for FATE's filter-pixfmts-* / filter-pixdesc-* tests (tests/fate-run.sh pixfmts(),
# pixdesc()) on the vsynth1 pictures (352x288), for the formats this repository's swscale path writes.  Each is the md5 of a
# NUT stream of rawvideo frames of `scale,format=FMT,FILTER` with flags bicubic+accurate_rnd+bitexact.  FILTER: null/copy =
# the converted picture; vflip / hflip / crop=100:100:100:100 = that picture flipped / cropped (byte moves, done by the test);
# scale = 200:100 through a second scaler reading FMT (yuv420p, nv12 / nv21, and rgb24 / bgr24 through the packed-RGB input readers; the 32-bit
# formats would carry alpha through the scaler: not built); pixdesc = frames 0..4, no filter.
# columns: reference file:line  test  pixel format  output size  frames  md5
''')
    for r in rows: o.write('%-44s %-8s %-8s %-8s %d  %s\n'%r)
print(len(rows))
